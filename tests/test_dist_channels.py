"""The N > 1 path of bench.py is 'one independent RF channel per rank, no data-path
collective': only the timing (max over ranks) and the sample count (sum) are reduced.
World-size-2 gloo run on CPU of exactly that reduction logic."""
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
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        # each rank owns whole channels: channel c -> rank c %% world (SURVEY.md section 8e)
        channels = [c for c in range(5) if c %% world == rank]
        ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)       # pretend device time
        samples = torch.tensor([len(channels) * 40960000.0], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(samples, op=dist.ReduceOp.SUM)
        dist.barrier()
        if rank == 0:
            assert ms.item() == 15.0 and samples.item() == 5 * 40960000.0
            print("OK", samples.item() / (ms.item() / 1e3) / 1e6)
        dist.destroy_process_group()
    """ % ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK" in r.stdout
