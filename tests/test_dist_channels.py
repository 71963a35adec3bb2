"""The N > 1 path is 'one independent RF channel per rank, no data-path collective' (SURVEY.md section 8e): the
ranks exchange the slowest rank's time (MAX) and the samples rendered (SUM), nothing else. World-size-2 gloo run
on CPU of the product's own multi-rank code: bench.py's channel_of_rank / reduce_job / cpu_budget /
bind_to_gpu_numa (which must be a no-op without a GPU), and each rank building the host-side tables of ITS
channels through the C-ABI (htv_tables_create: a different --offset per channel), independently of the other."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_channel_partition_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import bench
        import hacktv_b200 as H
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        assert bench.bind_to_gpu_numa(rank) is None or True            # no GPU here: must not raise
        assert bench.cpu_budget() >= 1
        mine = bench.channel_of_rank(rank, world, 5)
        assert mine == [c for c in range(5) if c %% world == rank]
        # every rank prepares its own channels (host side of htv_init): 8 MHz raster, one offset per channel
        widths = []
        for c in mine:
            t = H.Tables(H.mode_config("i", vfilter=True, offset=(c - 2) * 8000000), 20000000)
            widths.append(int(t.get("geometry")[0]))
            t.close()
        assert widths == [1280] * len(mine), widths
        ms, samples = bench.reduce_job(torch, dist, world, 10.0 + 5.0 * rank, len(mine) * 40960000.0, device="cpu")
        dist.barrier()
        assert ms == 15.0 and samples == 5 * 40960000.0, (ms, samples)
        if rank == 0:
            print("OK", samples / (ms / 1e3) / 1e6)
        dist.destroy_process_group()
    """ % ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK" in r.stdout
