"""GPU parity for the VBI overlay hook ("sparse per-line int16 delta lists added before
the VSB FIR"). The oracle's overlay stage is pinned to the reference's own WSS output on the CPU
(test_oracle_vs_ref.py::test_vbi_overlay_is_where_the_reference_puts_wss); here the CUDA path is compared with
the oracle on overlays placed where the reference's VBI stages put theirs."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _overlays(W, half, seed):
    """Teletext-like data bursts on VBI lines of both fields, a WSS-like line 23 (part of the picture set to
    black first), clear of the line ends."""
    rng = np.random.default_rng(seed)
    out = []
    for line in (7, 8, 19, 20, 23, 320, 321, 333):
        add = np.zeros(W, dtype=np.int16)
        a, b = 150, W - 60
        bits = rng.integers(0, 2, size=(b - a + 7) // 8)
        add[a:b] = (np.repeat(bits, 8)[: b - a] * 9000).astype(np.int16)
        rep = (half, int(W * 0.66), 1234) if line == 23 else (0, 0, 0)
        out.append((line, add, rep))
    out.append((335, None, (200, 400, -777)))              # replace only
    return out


CASES = [
    ("i", 16000000, 1400, dict(vfilter=True), 1),            # TMA modulator
    ("i", 16000000, 1400, dict(vfilter=True, noaudio=True), 0),
    ("pal", 16000000, 700, dict(), 0),                       # real output
    ("m", 13500000, 1100, dict(vfilter=True, noaudio=True), 0),   # 525 lines, W = 858
    ("l", 16000000, 1400, dict(vfilter=True, noaudio=True), 0),   # SECAM: behind the SECAM stage
    ("secam", 16000000, 700, dict(), 0),
    ("pal-fm", 20000000, 700, dict(vfilter=True, noaudio=True), 1),
]


@pytest.mark.parametrize("mode,rate,nlines,kw,tol", CASES)
def test_overlay_parity(built, mode, rate, nlines, kw, tol):
    H = built
    conf = H.mode_config(mode, **kw)
    enc = H.Encoder(conf, rate); o = orc.Oracle(conf, rate)
    ov = _overlays(enc.width, enc.width // 2, 5)
    if enc.lines == 525:
        ov = [(l, a, r) for (l, a, r) in ov if l < 300]
    enc.open_test_source(); o.open_test_source()
    enc.set_vbi_lines(ov)
    for line, add, rep in ov:
        o.add_vbi_line(line, add, rep)
    got = enc.render_host(nlines); want = o.render(nlines)
    plain = H.Encoder(conf, rate); plain.open_test_source()
    base = plain.render_host(nlines); plain.close()
    enc.close(); o.close()
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= tol, f"max |diff| {d.max()}, {np.count_nonzero(d > tol)} out of tolerance"
    assert np.count_nonzero(got != base) > 1000                 # the overlays are really there


def test_overlay_chunking_and_per_frame_pull(built):
    """Overlays survive a call boundary in the middle of their frame, and frames without overlays have none."""
    H = built
    conf = H.mode_config("i", vfilter=True, noaudio=True)
    a = H.Encoder(conf, 16000000); a.open_test_source()
    ov = _overlays(a.width, a.width // 2, 9)
    a.set_vbi_lines(ov, every=2)
    whole = a.render_host(2600); a.close()
    b = H.Encoder(conf, 16000000); b.open_test_source()
    b.set_vbi_lines(ov, every=2)
    parts = np.concatenate([b.render_host(n) for n in (5, 17, 300, 328, 1250, 700)]); b.close()
    assert np.array_equal(whole, parts)
    p = H.Encoder(conf, 16000000); p.open_test_source()
    base = p.render_host(2600); p.close()
    W2 = a.width * 2
    changed = np.nonzero((whole != base).reshape(2600, W2).any(axis=1))[0]
    frames = set((changed // 625).tolist())
    assert frames == {0, 2, 4}, frames                          # frames 1, 3, 5 (1-based) carry overlays


def test_overlay_reaching_the_line_end_survives_a_call_boundary(built):
    """a VBI waveform may run up to the last sample of its line (teletext's raised cosine does, ref
    teletext.c:1069); with the video filter on it then reaches into the first 25 samples of the NEXT line. A
    render call that starts right behind such a line re-rasterises it as its left neighbour and needs its
    overlay again."""
    H = built
    for mode, kw in (("i", dict(vfilter=True, noaudio=True)), ("pal-fm", dict(vfilter=True, noaudio=True))):
        rate = 16000000 if mode == "i" else 20000000
        conf = H.mode_config(mode, **kw)
        a = H.Encoder(conf, rate); a.open_test_source()
        W = a.width
        add = np.zeros(W, dtype=np.int16); add[W - 300:] = 7000    # non-zero up to sample W - 1
        ov = [(20, add, (0, 0, 0)), (21, add, (0, 0, 0)), (333, add, (0, 0, 0))]
        a.set_vbi_lines(ov)
        whole = a.render_host(700); a.close()
        b = H.Encoder(conf, rate); b.open_test_source(); b.set_vbi_lines(ov)
        parts = np.concatenate([b.render_host(n) for n in (20, 1, 1, 311, 367)]); b.close()   # boundaries right behind lines 20, 21, 333
        assert np.array_equal(whole, parts), f"{mode}: {np.count_nonzero(whole != parts)} values differ"
