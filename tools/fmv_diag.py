"""Diagnostics for FM video parity on the GPU box: how the GPU output and the oracle differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hacktv_b200 as H
import orc


def pair(mode, rate, n, **kw):
    conf = H.mode_config(mode, **kw)
    e = H.Encoder(conf, rate); o = orc.Oracle(conf, rate)
    e.open_test_source(); o.open_test_source()
    g = e.render_host(n); w = o.render(n); e.close(); o.close()
    return g, w


rate = 20000000
step = 2 * np.pi * 16e6 / 32767 / rate
n = 700
g, w = pair("pal-fm", rate, n)
gc = g.astype(np.float64).reshape(n, -1, 2); gc = gc[..., 0] + 1j * gc[..., 1]
wc = w.astype(np.float64).reshape(n, -1, 2); wc = wc[..., 0] + 1j * wc[..., 1]
rot = np.angle((gc * np.conj(wc)).sum(axis=1))
res = gc * np.exp(-1j * rot)[:, None] - wc
worst = np.maximum(np.abs(res.real), np.abs(res.imag)).max(axis=1)
print("rot/step every 50 lines:", np.round(rot[::50] / step, 1))
print("worst residual per line: hist", np.histogram(worst, bins=[0, 1.5, 3, 6, 12, 25, 50, 100, 1000])[0])
L = int(np.argmax(worst))
print("worst line", L, worst[L])
e = np.angle(gc[L] * np.conj(wc[L])) / step
print("  rot/step in 32-sample blocks:", np.round(e.reshape(-1, 32).mean(axis=1), 1))
# where does the rotation change? per-sample jumps of the (noisy) rotation, thresholded
full = np.unwrap(np.angle((gc * np.conj(wc)).reshape(-1))) / step
blk = full[: (full.size // 16) * 16].reshape(-1, 16).mean(axis=1)
j = np.nonzero(np.abs(np.diff(blk)) > 0.6)[0]
print("blocks(16) where rotation jumps >0.6 step:", len(j), "of", blk.size, "first:", j[:30] * 16, np.round(np.diff(blk)[j[:30]], 1))

g, w = pair("pal-fm", rate, 500, vfilter=True, noaudio=True, offset=-3000000, swap_iq=True)
d = np.abs((g.astype(np.int32) - w.astype(np.int32) + 32768) % 65536 - 32768)
print("offset+swap circular: max", d.max(), "<=1 %.5f" % (d <= 1).mean())
g, w = pair("pal-fm", rate, 500, vfilter=True, noaudio=True, offset=2500000, level=0.5)
d = np.abs(g.astype(np.int32) - w)
print("offset level 0.5: max", d.max(), "<=1 %.5f" % (d <= 1).mean())
