"""Device-resident SECAM timing (64 frames per call) + pass statistics."""
import sys, torch
sys.path.insert(0, '.')
import hacktv_b200 as H
for mode, rate in (("l", 16000000), ("l", 13500000), ("l", 20000000)):
    enc = H.Encoder(H.mode_config(mode, vfilter=True), rate); enc.open_test_source()
    n = 64 * enc.lines
    out = torch.empty(n * enc.width * 2, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): enc.render(n, out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): enc.render(n, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(mode, rate, "W", enc.width, "ms per 64 frames", round(ms, 4), "realtime_x", round(64 / 25 / (ms / 1e3), 1))
