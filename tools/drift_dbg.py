"""Carrier drift probe: GPU vs oracle after T seconds for several mode/rate pairs (debug tool)."""
import sys, os, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import hacktv_b200 as H, orc
T = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
for mode, rate, kw in (("m", 13500000, dict(vfilter=True)), ("m", 16000000, dict(vfilter=True)), ("i", 13500000, dict(vfilter=True, nonicam=True)),
                       ("i", 16000000, dict(vfilter=True, nonicam=True)), ("m", 13500000, dict(vfilter=True, nocolour=True))):
    conf = H.mode_config(mode, **kw)
    enc = H.Encoder(conf, rate); enc.open_test_source()
    o = orc.Oracle(conf, rate); o.open_test_source()
    skip = int(T * rate / enc.width)
    scratch = torch.empty(20000 * enc.width * 2, dtype=torch.int16, device='cuda'); st = torch.cuda.current_stream().cuda_stream
    at = 0
    while at < skip:
        n = min(20000, skip - at); enc.render(n, scratch.data_ptr(), st); o.render(n); at += n
    got = enc.render_host(100); want = o.render(100)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    print(mode, rate, kw, f"after {T} s: max {d.max()} exact {(d == 0).mean():.3f}", flush=True)
    enc.close(); o.close()
