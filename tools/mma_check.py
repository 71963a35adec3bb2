#!/usr/bin/env python
"""Tensor-core video filter (k_mod_mma, HTV_FIR=mma) against the scalar TMA modulator
(HTV_FIR=scalar) and the oracle, then the device-resident timing of both. One JSON line per
check into gpurun_out/mma_check.jsonl. Run on the GPU box: python tools/mma_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import hacktv_b200 as H
import orc

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "mma_check.jsonl"), "w")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    LOG.write(s + "\n"); LOG.flush()


def render(sel, mode, rate, nlines, frames=None, audio=None, **kw):
    os.environ["HTV_FIR"] = sel
    conf = H.mode_config(mode, **kw)
    enc = H.Encoder(conf, rate)
    if frames is None:
        enc.open_test_source()
    else:
        enc.set_source(frames, audio)
    got = enc.render_host(nlines)
    enc.close()
    return got


CASES = [
    ("i", 16000000, 1300, dict(vfilter=True, noaudio=True), 0),
    ("i", 16000000, 1300, dict(vfilter=True), 1),
    ("i", 20000000, 700, dict(vfilter=True), 1),
    ("pal", 16000000, 700, dict(vfilter=True), 0),
    ("b", 16000000, 700, dict(vfilter=True), 1),
    ("i", 16000000, 700, dict(vfilter=True, swap_iq=True), 1),
    ("i", 16000000, 700, dict(vfilter=True, level=0.7, volume=2.0), 1),
]
ok_all = True
for mode, rate, nlines, kw, tol in CASES:
    a = render("scalar", mode, rate, nlines, **kw)
    b = render("mma", mode, rate, nlines, **kw)
    conf = H.mode_config(mode, **kw)
    o = orc.Oracle(conf, rate); o.open_test_source()
    want = o.render(nlines); o.close()
    d_sm = np.abs(a.astype(np.int32) - b.astype(np.int32))
    d_o = np.abs(b.astype(np.int32) - want.astype(np.int32))
    ok = bool(d_sm.max() == 0 and d_o.max() <= tol)
    ok_all &= ok
    emit(check="parity", mode=mode, rate=rate, nlines=nlines, kw={k: str(v) for k, v in kw.items()},
         mma_vs_scalar_max=int(d_sm.max()), mma_vs_scalar_nonzero=int(np.count_nonzero(d_sm)),
         mma_vs_oracle_max=int(d_o.max()), tol=tol, ok=ok,
         first_bad=int(np.argmax(d_sm > 0)) if d_sm.max() else -1)

# random pictures: full-range composite values through the byte split
rng = np.random.default_rng(99)
conf = H.mode_config("i", vfilter=True)
e = H.Encoder(conf, 16000000); al, aw = e.active_lines, e.active_width; e.close()
frames = rng.integers(0, 1 << 24, size=(3, al, aw), dtype=np.uint32)
audio = rng.integers(-32768, 32767, size=(40000, 2), dtype=np.int16)
a = render("scalar", "i", 16000000, 1900, frames=frames, audio=audio, vfilter=True)
b = render("mma", "i", 16000000, 1900, frames=frames, audio=audio, vfilter=True)
d = np.abs(a.astype(np.int32) - b.astype(np.int32))
ok = bool(d.max() == 0); ok_all &= ok
emit(check="random pictures", mma_vs_scalar_max=int(d.max()), nonzero=int(np.count_nonzero(d)), ok=ok)

# chunking with the mma path
os.environ["HTV_FIR"] = "mma"
conf = H.mode_config("i", vfilter=True)
x = H.Encoder(conf, 16000000); x.open_test_source(); whole = x.render_host(1500); x.close()
y = H.Encoder(conf, 16000000); y.open_test_source()
parts = np.concatenate([y.render_host(n) for n in (1, 311, 313, 625, 250)]); y.close()
ok = bool(np.array_equal(whole, parts)); ok_all &= ok
emit(check="chunking", ok=ok)

# timing, device resident, 64 frames per call
for name, mode, rate in (("cfg2 i 16M --filter", "i", 16000000), ("cfg5 i 20M --filter", "i", 20000000),
                         ("pal 16M --filter (real low-pass)", "pal", 16000000)):
    for sel in ("scalar", "mma"):
        os.environ["HTV_FIR"] = sel
        enc = H.Encoder(H.mode_config(mode, vfilter=True), rate)
        enc.open_test_source()
        nlines = 64 * enc.lines
        out = torch.empty(nlines * enc.width * 2, dtype=torch.int16, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            enc.render(nlines, out.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for _ in range(K):
            enc.render(nlines, out.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        msps = nlines * enc.width / (ms / 1e3) / 1e6
        emit(check="timing", config=name, fir=sel, ms_per_64_frames=round(ms, 4), msamples_per_s=round(msps, 1),
             realtime_x=round(msps / (rate / 1e6), 1))
        enc.close()
emit(check="summary", ok=bool(ok_all))
sys.exit(0 if ok_all else 1)
