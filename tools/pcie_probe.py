"""Host<->device copy bandwidth of this box (pinned memory), alone and in both directions at
once: the ceiling for bench.py's e2e figure (164 MB D2H + 123 MB H2D per 64-frame step)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--numa" in sys.argv:
    # as bench.py does before it allocates anything: cores and pinned pages next to the GPU
    import bench
    print(json.dumps({"numa_bound_cpus": bench.bind_to_gpu_numa(0), "numa_memory": bench.prefer_gpu_memory_node(0)}))

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        s1.synchronize(); s2.synchronize()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


import time
def wall(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best

r = {"bytes": n,
     "h2d_gbs": n / wall(h2d) / 1e9, "d2h_gbs": n / wall(d2h) / 1e9,
     "both_each_gbs": n / wall(both) / 1e9}
print(json.dumps(r))

# the encoder's actual pattern: H2D in picture-sized copies (1.9 MB), D2H in 24 MB pieces
fb = 832 * 576 * 4
def h2d_small():
    with torch.cuda.stream(s1):
        for i in range(64):
            d_a[i * fb:(i + 1) * fb].copy_(h_in[:fb], non_blocking=True)
def d2h_pieces():
    with torch.cuda.stream(s2):
        pb = 24 << 20
        for i in range(7):
            h_out[i * pb:(i + 1) * pb].copy_(d_b[i * pb:(i + 1) * pb], non_blocking=True)
def both_small():
    h2d_small(); d2h_pieces()
t_h, t_d, t_b = wall(h2d_small), wall(d2h_pieces), wall(both_small)
print(json.dumps({"h2d_64x1.9MB_ms": t_h * 1e3, "h2d_gbs": 64 * fb / t_h / 1e9, "d2h_7x24MB_ms": t_d * 1e3,
                  "d2h_gbs": 7 * (24 << 20) / t_d / 1e9, "both_ms": t_b * 1e3}))

# round 2's pattern: the step's pictures in 8 contiguous uploads (15.4 MB each) beside the IQ in 8 MB pieces
def h2d_runs():
    with torch.cuda.stream(s1):
        rb = 8 * fb
        for i in range(8):
            d_a[i * rb:(i + 1) * rb].copy_(h_in[i * rb:(i + 1) * rb], non_blocking=True)
def d2h_8mb():
    with torch.cuda.stream(s2):
        pb = 8 << 20
        for i in range(20):
            h_out[i * pb:(i + 1) * pb].copy_(d_b[i * pb:(i + 1) * pb], non_blocking=True)
def both_r2():
    h2d_runs(); d2h_8mb()
t_h, t_d, t_b = wall(h2d_runs), wall(d2h_8mb), wall(both_r2)
print(json.dumps({"h2d_8x15.3MB_ms": t_h * 1e3, "d2h_20x8MB_ms": t_d * 1e3, "both_ms": t_b * 1e3,
                  "floor_for_123MB_up_164MB_down_ms": 1e3 * max(t_h, t_d * 163.84 / 167.77, t_b * 163.84 / 167.77)}))
