"""The tensor-core video filter (k_mod_mma, the default whenever a video filter is on - any line width)
against the scalar TMA modulator it replaces (HTV_FIR=scalar) - bit for bit, the FIR is exact
integer work either way (ref fir.c:564-615) - and against the oracle. HTV_FIR is read when an
encoder is created, so both variants run in one process."""
import os

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _render(H, sel, mode, rate, nlines, frames=None, audio=None, **kw):
    old = os.environ.get("HTV_FIR")
    os.environ["HTV_FIR"] = sel
    try:
        enc = H.Encoder(H.mode_config(mode, **kw), rate)
        if frames is None:
            enc.open_test_source()
        else:
            enc.set_source(frames, audio)
        got = enc.render_host(nlines)
        launches = enc.kernel_launches
        enc.close()
    finally:
        if old is None:
            os.environ.pop("HTV_FIR", None)
        else:
            os.environ["HTV_FIR"] = old
    assert launches > 0
    return got


def _same(a, b, tol):
    """The filter is exact integer work either way: without sound carriers the two selections agree bit for bit.
    HTV_FIR=scalar also selects the split raster / modulator kernels, whose FM carrier takes one sin/cos per
    sample where the fused line kernel rotates (htv_line.cuh): with a carrier the two may round differently on a
    few samples per thousand, never by more than the carrier's own +-1 LSB."""
    if tol == 0:
        assert np.array_equal(a, b), f"{np.count_nonzero(a != b)} samples differ between the two filters"
    else:
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 0.01, (d.max(), (d != 0).mean())


CASES = [
    ("i", 16000000, 1300, dict(vfilter=True, noaudio=True), 0),      # VSB, W = 1024: 8 tiles, 8 warps
    ("i", 16000000, 1300, dict(vfilter=True), 1),                    # BASELINE config 2
    ("i", 20000000, 700, dict(vfilter=True), 1),                     # config 5: W = 1280, 10 tiles, 320 threads
    ("pal", 16000000, 700, dict(vfilter=True), 0),                   # real low-pass: no Q taps, real output
    ("b", 16000000, 700, dict(vfilter=True), 1),
    ("i", 16000000, 700, dict(vfilter=True, swap_iq=True), 1),
    ("i", 16000000, 700, dict(vfilter=True, level=0.7, volume=2.0), 1),
]


@pytest.mark.parametrize("mode,rate,nlines,kw,tol", CASES)
def test_mma_filter_equals_scalar_filter_and_oracle(built, mode, rate, nlines, kw, tol):
    H = built
    scalar = _render(H, "scalar", mode, rate, nlines, **kw)
    mma = _render(H, "mma", mode, rate, nlines, **kw)
    _same(scalar, mma, tol)
    o = orc.Oracle(H.mode_config(mode, **kw), rate); o.open_test_source()
    want = o.render(nlines); o.close()
    d = np.abs(mma.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= tol, f"max |diff| vs oracle = {d.max()}"


def test_default_is_the_tensor_core_filter(built):
    """No HTV_FIR: the default path must be bit-identical to both explicit selections."""
    H = built
    old = os.environ.pop("HTV_FIR", None)
    try:
        enc = H.Encoder(H.mode_config("i", vfilter=True, noaudio=True), 16000000)
        enc.open_test_source(); default = enc.render_host(700); enc.close()
    finally:
        if old is not None:
            os.environ["HTV_FIR"] = old
    assert np.array_equal(default, _render(H, "mma", "i", 16000000, 700, vfilter=True, noaudio=True))
    assert np.array_equal(default, _render(H, "scalar", "i", 16000000, 700, vfilter=True, noaudio=True))


def test_random_pictures_full_range_through_the_byte_split(built):
    """Random colours drive the composite signal over its whole int16 range (negative high bytes,
    every low byte), random audio exercises the sound phase behind the exchange buffer."""
    H = built
    rng = np.random.default_rng(99)
    e = H.Encoder(H.mode_config("i", vfilter=True), 16000000); al, aw = e.active_lines, e.active_width; e.close()
    frames = rng.integers(0, 1 << 24, size=(3, al, aw), dtype=np.uint32)
    audio = rng.integers(-32768, 32767, size=(40000, 2), dtype=np.int16)
    a = _render(H, "scalar", "i", 16000000, 1900, frames=frames, audio=audio, vfilter=True)
    b = _render(H, "mma", "i", 16000000, 1900, frames=frames, audio=audio, vfilter=True)
    _same(a, b, 1)
    a = _render(H, "scalar", "i", 16000000, 700, frames=frames, audio=audio, vfilter=True, noaudio=True)
    b = _render(H, "mma", "i", 16000000, 700, frames=frames, audio=audio, vfilter=True, noaudio=True)
    _same(a, b, 0)


def test_chunking_is_invisible_with_the_mma_filter(built):
    H = built
    whole = _render(H, "mma", "i", 16000000, 1500, vfilter=True)
    old = os.environ.get("HTV_FIR"); os.environ["HTV_FIR"] = "mma"
    try:
        y = H.Encoder(H.mode_config("i", vfilter=True), 16000000); y.open_test_source()
        parts = np.concatenate([y.render_host(n) for n in (1, 311, 313, 625, 250)]); y.close()
    finally:
        if old is None:
            os.environ.pop("HTV_FIR", None)
        else:
            os.environ["HTV_FIR"] = old
    assert np.array_equal(whole, parts)


def test_other_line_widths_take_the_same_path(built):
    """18 Msps: W = 1152 = 9 tiles on 9 warps (288 threads) - the tile/warp split is generic."""
    H = built
    a = _render(H, "scalar", "i", 18000000, 700, vfilter=True, noaudio=True)
    b = _render(H, "mma", "i", 18000000, 700, vfilter=True, noaudio=True)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("mode,rate,nlines,kw", [
    ("m", 13500000, 1100, dict(vfilter=True, noaudio=True)),        # BASELINE config 3 geometry: W = 858
    ("m", 13500000, 1100, dict(vfilter=True)),
    ("i", 13500000, 700, dict(vfilter=True, noaudio=True)),         # W = 864
])
def test_pitched_layout_any_width(built, mode, rate, nlines, kw):
    H = built
    a = _render(H, "scalar", mode, rate, nlines, **kw)
    b = _render(H, "mma", mode, rate, nlines, **kw)
    _same(a, b, 0 if kw.get("noaudio") else 1)
