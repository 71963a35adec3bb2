#!/usr/bin/env python
"""bench.py - IQ Msamples/s of the composite-video -> IQ hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one batch: FRAMES video frames (default 64 =
40 000 scan lines = 40.96 M complex samples = 164 MB of int16 IQ, larger than the 126 MB
L2) of BASELINE config 2: PAL System I, 16 Msps, --filter (VSB) + FM mono + NICAM-728 +
colour, built-in test pattern and tone. With N GPUs each rank renders its own independent
RF channel on its own GPU (weak scaling, no collective on the data path); `value` is
the sum over channels divided by the slowest rank's device time.

Keys beyond the base contract:
  roofline      HBM-write roofline of the dominant kernel (k_mod_tma): algorithmic bytes
                (4 B per complex sample) / CUDA-event duration of that kernel, vs the
                measured copy bandwidth in MEASURED_PEAKS.json.
  e2e           the same metric through htv_render_host() with HOST buffers: every
                step uploads its pictures (a live source: one upload per frame) and sound,
                renders, and copies the IQ back to pinned host memory.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/ref_harness) timed on this box's
                host cores on a bounded sample of the same workload (rank 0, N = 1 only).

--impl reference times the reference's own CPU implementation with every host thread it
can use (as many concurrent encoder instances as fit; one instance = 1 + nthreads
pthreads) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODE, RATE, FILTER = "i", 16_000_000, True
WORKLOAD = "PAL-I (-m i) 16 Msps --filter: VSB + FM mono + NICAM-728 + colour, built-in test pattern"
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
REF_THREADS = 3          # main/raster + vfilter + audio (reference video.c:4692: 1 + nthreads)
# dram__bytes_read.sum + dram__bytes_write.sum of one modulator launch from the `ncu --set full` captures
# (ncu flushes L2 between replays, so the L2-resident composite scratch is re-read from HBM and most of the
# output is still in L2 when the kernel ends), per scan line of 1024 samples:
#   k_mod_mma (default; profiles/r01_ncu_mod_mma_raw.txt): 8 192 lines, 21.71 MB read (two byte planes) + 0.11 MB written
#   k_mod_tma (HTV_FIR=scalar; profiles/r01_ncu_mod_raw.txt): 7 232 lines, 38.48 MB read (int32 scratch) + 3.22 MB written
NCU_TRAFFIC_BYTES_PER_LINE = {"mma": 21_822_208 / 8192, "scalar": 41_701_120 / 7232}
FIR = "scalar" if os.environ.get("HTV_FIR") == "scalar" else "mma"
KERNEL = {"mma": "k_mod_mma (video filter on the tensor cores: exact int8 byte-split mma.sync + sound carriers + IQ store)",
          "scalar": "k_mod_tma (scalar video filter + sound carriers + IQ store)"}[FIR]


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md) through
    NVML at ~1 kHz (the region is only milliseconds long; spawning nvidia-smi is too slow).
    Falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.t0 = self.t1 = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            # torch device index -> NVML handle (respects CUDA_VISIBLE_DEVICES through the UUID)
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                if nv:
                    clk = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    self.rows.append((time.perf_counter(), clk, self._max, rs))
                else:
                    q = "clocks.sm,clocks.max.sm"
                    r = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5)
                    p = [x.strip() for x in r.stdout.strip().split(",")]
                    self.rows.append((time.perf_counter(), int(p[0]), int(p[1]), 0))
            except Exception:
                pass
            self._stop.wait(0.001 if nv else 0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self):
        rows = [r for r in self.rows if self.t0 is None or (self.t0 <= r[0] <= self.t1)] or self.rows
        sm = sorted(r[1] for r in rows)
        reasons = set()
        nv = self._nv
        if nv:
            names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                     "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
            for r in rows:
                for k, attr in names.items():
                    bit = getattr(nv, attr, None) or getattr(nv, attr.replace("Event", "Throttle"), 0)
                    if bit and (r[3] & bit):
                        reasons.add(k)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "reasons": sorted(reasons), "samples": len(rows), "source": "nvml" if nv else "nvidia-smi"}


def run_reference_instances(ninst, frames, timeout=900):
    """ninst concurrent reference encoders, each timing `frames` frames after a 2-frame
    warm-up; returns (aggregate Msamples/s, per-instance list, wall seconds)."""
    lines = frames * 625
    cmd = ["timeout", str(timeout), REF_HARNESS, "-m", MODE, "-s", str(RATE), "--skip", "1250", "--lines", str(lines), "--bench"]
    if FILTER:
        cmd.append("--filter")
    t0 = time.time()
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(ninst)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.time() - t0
    per = []
    for o in outs:
        for ln in o.splitlines():
            if ln.startswith("{") and "msamples_per_s" in ln:
                per.append(json.loads(ln)["msamples_per_s"])
    if len(per) != ninst:
        raise RuntimeError("reference harness failed")
    return sum(per), per, wall


def bench_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not os.path.exists(REF_HARNESS):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness was not built (no reference tree at build time)"}))
        return
    ninst = max(1, cores // REF_THREADS)
    frames = args.ref_frames
    for _ in range(args.warmup):
        run_reference_instances(ninst, max(2, frames // 4))
    vals, t = [], 0.0
    for _ in range(args.steps):
        agg, per, wall = run_reference_instances(ninst, frames)
        vals.append(agg)
        t += wall
    v = sum(vals) / len(vals)
    sample = f"{ninst} concurrent reference encoders x {frames} frames ({frames * 625 * 1024 / 1e6:.1f} Msamples each) per step, vid_next_line loop, no sink I/O"
    print(json.dumps({
        "impl": "reference", "metric": "IQ Msamples/s", "value": round(v, 3), "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000 * t / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16 (int32/int64 accumulate)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"{ninst} processes x {REF_THREADS} threads on {cores} host cores"},
        "cpu_baseline": {"value": round(v, 3), "unit": "Msamples/s", "cores": ninst * REF_THREADS, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=64, help="video frames per step")
    ap.add_argument("--ref-frames", type=int, default=100, help="frames per reference instance per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        bench_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import hacktv_b200 as H

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    conf = H.mode_config(MODE, vfilter=FILTER)
    enc = H.Encoder(conf, RATE)            # one RF channel per rank / GPU
    enc.open_test_source()
    nlines = args.frames * enc.lines
    nsamp = nlines * enc.width
    out = torch.empty(nsamp * 2, dtype=torch.int16, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    enc.set_kernel_timing(True)

    # ---- device-resident throughput ("value") ---------------------------------
    for _ in range(args.warmup):
        enc.render(nlines, out.data_ptr(), stream)
    barrier()
    l0 = enc.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    with ClockSampler(local) as clk:
        time.sleep(0.005)
        clk.t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            enc.render(nlines, out.data_ptr(), stream)
        ev1.record()
        barrier()
        clk.t1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = enc.kernel_launches - l0
    clocks = clk.summary()
    # per-launch duration of the dominant kernel, CUDA events on its own stream (untimed extra steps)
    for _ in range(3):
        enc.render(nlines, out.data_ptr(), stream)
        torch.cuda.synchronize()
        kern_ms.append(enc.last_line_kernel_ms())
    kern_lines = enc.last_line_kernel_lines()
    checksum = int(out[:4096].to(torch.int32).sum().item())

    # ---- end to end through the C-ABI with host buffers ("e2e") -----------------
    e2e = None
    if not args.no_e2e:
        e2e_frames = args.frames
        e2e_lines = e2e_frames * enc.lines
        enc2 = H.Encoder(conf, RATE)
        pic = torch.from_numpy(H.test_pattern(enc2.active_width, enc2.active_lines).astype(np.int32)).pin_memory()
        tone = torch.from_numpy(H.test_tone()).pin_memory()
        # a live source: a new picture serial every frame -> one H2D upload per frame
        enc2.open_memory_source(pic.numpy().view(np.uint32)[None], tone.numpy(), audio_block=8192, static_video=False)
        host = torch.empty(e2e_lines * enc2.width * 2, dtype=torch.int16).pin_memory()
        for _ in range(max(1, args.warmup)):
            enc2.render_host_ptr(e2e_lines, host.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            enc2.render_host_ptr(e2e_lines, host.data_ptr())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_samples = e2e_lines * enc2.width
        h2d = e2e_frames * enc2.active_width * enc2.active_lines * 4 + int(e2e_samples / RATE * 32000) * 4
        e2e = {"value": round(world * e2e_samples * args.steps / tt.item() / 1e6, 2), "unit": "Msamples/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": e2e_samples * 4,
               "frames_per_step": e2e_frames, "checksum": int(host[:4096].to(torch.int32).sum().item()),
               "api": "htv_av_memory_open + htv_render_host (C-ABI), pinned host buffers, one picture upload per frame"}
        enc2.close()

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()

    if rank == 0:
        value = world * nsamp * args.steps / (ms_max / 1e3) / 1e6
        peak, peak_src = measured_peak_gbs()
        k_ms = sorted(kern_ms)[len(kern_ms) // 2]
        k_samples = kern_lines * enc.width
        achieved = k_samples * 4 / (k_ms / 1e3) / 1e9 if k_ms > 0 else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_HARNESS):
            v, per, wall = run_reference_instances(1, 150)
            cpu = {"value": round(v, 3), "unit": "Msamples/s", "cores": REF_THREADS, "kind": "reference",
                   "sample": "1 reference encoder (main + vfilter + audio threads), 150 frames = 96 Msamples after a 2-frame warm-up, "
                             "vid_next_line loop, no sink I/O (oracle/_ref/ref_harness)",
                   "host_cores": os.cpu_count()}
        print(json.dumps({
            "metric": "IQ Msamples/s", "value": round(value, 2), "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16 (int32 accumulate; RGB->YUV table built in fp64 at init)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": args.frames, "lines_per_step": nlines,
                       "samples_per_step_per_gpu": nsamp, "parallelism": f"{world} independent RF channel(s), one per GPU, no collectives",
                       "l2": f"each step writes {nsamp * 4 / 1e6:.0f} MB of IQ per GPU (> 126 MB L2); tables are L2-resident by design",
                       "realtime_x": round(value / world / (RATE / 1e6), 1)},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4) if achieved else None,
                         "traffic": int(round(NCU_TRAFFIC_BYTES_PER_LINE[FIR] * kern_lines)) if enc.width == 1024 else None,
                         "kernel": KERNEL, "kernel_ms": round(k_ms, 4),
                         "lines_per_launch": kern_lines, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k_samples * 4,
                         "note": "4 B per complex sample written once; the kernel is issue-slot bound (DESIGN.md section 4), not HBM bound"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "checksum": checksum,
        }))

    enc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
