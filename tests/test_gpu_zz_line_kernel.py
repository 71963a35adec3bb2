"""The fused line kernel (k_line, htv_line.cuh: raster + chroma + video filter + sound carriers + store in one
persistent launch, the default for PAL / NTSC / mono) against the split raster / modulator kernels it replaces
(HTV_PATH=split) and against the oracle. HTV_PATH is read when an encoder is created."""
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _render(H, path, mode, rate, pieces, frames=None, audio=None, **kw):
    old = os.environ.get("HTV_PATH")
    if path:
        os.environ["HTV_PATH"] = path
    else:
        os.environ.pop("HTV_PATH", None)
    try:
        enc = H.Encoder(H.mode_config(mode, **kw), rate)
        if frames is None:
            enc.open_test_source()
        else:
            enc.set_source(frames, audio)
        got = np.concatenate([enc.render_host(n) for n in pieces])
        enc.close()
    finally:
        if old is None:
            os.environ.pop("HTV_PATH", None)
        else:
            os.environ["HTV_PATH"] = old
    return got


CASES = [
    ("i", 16000000, 1300, dict(vfilter=True, noaudio=True), 0),      # W = 1024: 8 tiles, VSB
    ("i", 16000000, 1300, dict(vfilter=True), 1),                    # BASELINE config 2
    ("i", 20000000, 700, dict(vfilter=True), 1),                     # config 5: W = 1280, 10 warps
    ("m", 13500000, 1100, dict(vfilter=True), 1),                    # config 3: W = 858, partial last tile
    ("m", 13500000, 1100, dict(vfilter=True, noaudio=True), 0),
    ("pal", 16000000, 700, dict(), 0),                               # config 1: no filter, real output
    ("pal", 16000000, 700, dict(vfilter=True), 0),                   # real low-pass, no Q taps
    ("i", 16000000, 700, dict(vfilter=True, nocolour=True), 1),      # no chroma planes
    ("i", 16000000, 700, dict(vfilter=True, swap_iq=True, offset=1500000), 2),
    ("l", 16000000, 700, dict(vfilter=True, nocolour=True), 1),      # AM sound + NICAM through the fused kernel
    ("i", 14000000, 700, dict(vfilter=True), 1),                     # W = 896
    # SECAM: k_sec_raster + chain + k_line<SRC> against round 1's raster (k_raster_secam) + scalar modulator (k_mod)
    ("l", 16000000, 1300, dict(vfilter=True), 1),                    # BASELINE config 4
    ("l", 13500000, 900, dict(vfilter=True), 1),                     # W = 864: partial last tile, long FM work list
    ("secam", 16000000, 900, dict(), 0),                             # baseband: chroma chain only, bit-exact
    ("d", 20000000, 700, dict(vfilter=True), 1),                     # W = 1280, FM sound
    ("secam", 15109375, 700, dict(), 0),                             # W = 967: 8 does not divide W (chain tail of 7, scalar row ends)
    ("secam", 15218750, 700, dict(), 0),                             # W = 974: chain tail of 14 samples over two groups
    ("l", 14062500, 700, dict(vfilter=True), 1),                     # W = 900: 4 | W only
]


@pytest.mark.parametrize("mode,rate,nlines,kw,tol", CASES)
def test_fused_equals_split_and_oracle(built, mode, rate, nlines, kw, tol):
    H = built
    fused = _render(H, None, mode, rate, [nlines], **kw)
    split = _render(H, "split", mode, rate, [nlines], **kw)
    d = np.abs(fused.astype(np.int32) - split.astype(np.int32))
    if tol == 0:
        assert d.max() == 0, f"{np.count_nonzero(d)} values differ between the fused and the split kernels"
    else:
        assert d.max() <= tol and (d != 0).mean() < 0.01, (d.max(), (d != 0).mean())
    o = orc.Oracle(H.mode_config(mode, **kw), rate); o.open_test_source()
    want = o.render(nlines); o.close()
    assert np.abs(fused.astype(np.int32) - want.astype(np.int32)).max() <= tol


def test_secam_scalar_notch_equals_tensor_core_notch(built):
    """HTV_FIR=scalar keeps round 1's raster with the scalar luma notch; the default rasters with k_sec_raster (notch and
    baseband low-pass as int8 contractions). Bit-identical on the baseband mode, where no NCO contributes."""
    H = built
    a = _render(H, None, "secam", 16000000, [1000])
    old = os.environ.get("HTV_FIR")
    os.environ["HTV_FIR"] = "scalar"
    try:
        b = _render(H, "split", "secam", 16000000, [1000])
    finally:
        if old is None:
            os.environ.pop("HTV_FIR", None)
        else:
            os.environ["HTV_FIR"] = old
    assert np.array_equal(a, b)


def test_secam_runs_and_calls_are_invisible(built):
    """SECAM: the chain's state (IIR, aliased words) is carried across calls of any size; a persistent raster CTA walks
    a run of rows. Random pictures put many lines on the FM work list."""
    H = built
    rng = np.random.default_rng(5)
    e = H.Encoder(H.mode_config("l", vfilter=True), 16000000); al, aw = e.active_lines, e.active_width; e.close()
    frames = rng.integers(0, 1 << 24, size=(4, al, aw), dtype=np.uint32)
    whole = _render(H, None, "l", 16000000, [2600], frames=frames, vfilter=True)
    parts = _render(H, None, "l", 16000000, [1, 2, 311, 1000, 625, 661], frames=frames, vfilter=True)
    assert np.array_equal(whole, parts)
    split = _render(H, "split", "l", 16000000, [2600], frames=frames, vfilter=True)
    d = np.abs(whole.astype(np.int32) - split.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 0.01, (d.max(), (d != 0).mean())


def test_runs_and_calls_are_invisible(built):
    """A CTA walks a run of consecutive lines and rasters one line more either side; calls of any size - fewer
    lines than CTAs, one line, many frames - give the same stream."""
    H = built
    whole = _render(H, None, "i", 16000000, [9000], vfilter=True)
    parts = _render(H, None, "i", 16000000, [1, 2, 3, 311, 5000, 625, 3058], vfilter=True)
    assert np.array_equal(whole, parts)
    whole = _render(H, None, "pal", 16000000, [2000])
    parts = _render(H, None, "pal", 16000000, [7, 1, 1300, 692])
    assert np.array_equal(whole, parts)


def test_random_pictures_and_loud_audio(built):
    H = built
    rng = np.random.default_rng(21)
    e = H.Encoder(H.mode_config("i", vfilter=True), 16000000); al, aw = e.active_lines, e.active_width; e.close()
    frames = rng.integers(0, 1 << 24, size=(3, al, aw), dtype=np.uint32)
    audio = rng.integers(-32768, 32767, size=(40000, 2), dtype=np.int16)
    a = _render(H, None, "i", 16000000, [1900], frames=frames, audio=audio, vfilter=True, noaudio=True)
    b = _render(H, "split", "i", 16000000, [1900], frames=frames, audio=audio, vfilter=True, noaudio=True)
    assert np.array_equal(a, b)
    conf = H.mode_config("i", vfilter=True)
    a = _render(H, None, "i", 16000000, [1300], frames=frames, audio=audio, vfilter=True)
    o = orc.Oracle(conf, 16000000); o.set_source(frames, audio)
    want = o.render(1300); o.close()
    assert np.abs(a.astype(np.int32) - want.astype(np.int32)).max() <= 1
