"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle
on the same inputs. Bit-exact wherever the path is integer-only (video, chroma, VSB,
NICAM); within the +-1 LSB BASELINE.json states wherever an FM/AM sound carrier or the
offset mixer (closed-form NCOs) contributes."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


def _diff(a, b):
    assert a.shape == b.shape
    return np.abs(a.astype(np.int32) - b.astype(np.int32))


def _pair(H, mode, rate, nlines, frames=None, audio=None, **kw):
    conf = H.mode_config(mode, **kw)
    enc = H.Encoder(conf, rate)
    o = orc.Oracle(conf, rate)
    if frames is None and audio is None:
        enc.open_test_source()
        o.open_test_source()
    else:
        enc.set_source(frames, audio)
        o.set_source(frames, audio)
    got = enc.render_host(nlines)
    want = o.render(nlines)
    launches = enc.kernel_launches
    enc.close()
    o.close()
    assert launches > 0
    return got, want


# (mode, rate, lines, overrides, tolerance): BASELINE.json configs 1-5 first
CASES = [
    ("pal", 16000000, 1300, dict(), 0),                                   # config 1
    ("i", 16000000, 1300, dict(vfilter=True), 1),                         # config 2
    ("m", 13500000, 1100, dict(vfilter=True), 1),                         # config 3
    ("i", 20000000, 700, dict(vfilter=True), 1),                          # config 5 (one channel)
    ("i", 16000000, 1300, dict(vfilter=True, noaudio=True), 0),           # video + VSB only: exact
    ("i", 16000000, 700, dict(vfilter=True, nonicam=True), 1),
    ("i", 16000000, 700, dict(), 1),                                      # no filter: no audio lead
    ("i", 16000000, 1300, dict(vfilter=True, noaudio=False, nocolour=True), 1),
    ("m", 13500000, 1100, dict(vfilter=True, noaudio=True), 0),
    ("ntsc", 13500000, 1100, dict(), 0),
    ("pal", 16000000, 700, dict(vfilter=True), 0),                        # real low-pass filter
    ("b", 16000000, 700, dict(vfilter=True), 1),
    ("pal-m", 13500000, 600, dict(vfilter=True), 1),
    ("pal-n", 16000000, 700, dict(vfilter=True), 1),
    ("l", 16000000, 1300, dict(vfilter=True), 1),                         # config 4: SECAM-L, AM sound + NICAM
    ("l", 16000000, 1300, dict(vfilter=True, noaudio=True), 0),           # SECAM chroma is exact
    ("secam", 16000000, 1300, dict(), 0),
    ("d", 16000000, 700, dict(vfilter=True), 1),
    ("secam-i", 16000000, 700, dict(vfilter=True, nonicam=True), 1),
    ("i", 16000000, 700, dict(vfilter=True, swap_iq=True), 1),
    ("i", 16000000, 700, dict(vfilter=True, invert_video=True, noaudio=True), 0),
    ("i", 16000000, 700, dict(vfilter=True, level=0.7, volume=2.0), 1),
]


@pytest.mark.parametrize("mode,rate,nlines,kw,tol", CASES)
def test_test_pattern_parity(built, mode, rate, nlines, kw, tol):
    got, want = _pair(built, mode, rate, nlines, **kw)
    d = _diff(got, want)
    assert d.max() <= tol, f"max |diff| = {d.max()} at {np.argmax(d)}; {np.count_nonzero(d > tol)} samples out of tolerance"
    if tol:
        assert (d == 0).mean() > 0.97, f"only {(d == 0).mean():.4f} of samples exact"


def test_nicam_only_is_exact(built):
    """NICAM is integer-only end to end (frame encoder, DQPSK, pulse shaping, carrier LUT)."""
    H = built
    conf = H.mode_config("i", vfilter=True)
    conf.fm_mono_level = 0.0
    conf.fm_mono_carrier = 0.0
    enc = H.Encoder(conf, 16000000); o = orc.Oracle(conf, 16000000)
    enc.open_test_source(); o.open_test_source()
    got, want = enc.render_host(1300), o.render(1300)
    enc.close(); o.close()
    assert _diff(got, want).max() == 0


def test_random_pictures_and_loud_audio(built):
    """Random frames (a new one every frame) and full-scale noise audio: exercises the
    fp64 RGB->YUV path on arbitrary colours, the limiter's attack, and NICAM companding."""
    H = built
    rng = np.random.default_rng(1234)
    conf = H.mode_config("i", vfilter=True)
    enc = H.Encoder(conf, 16000000)
    frames = rng.integers(0, 1 << 24, size=(3, enc.active_lines, enc.active_width), dtype=np.uint32)
    audio = rng.integers(-32768, 32767, size=(40000, 2), dtype=np.int16)
    enc.close()
    got, want = _pair(H, "i", 16000000, 1900, frames=frames, audio=audio, vfilter=True)
    d = _diff(got, want)
    assert d.max() <= 1, f"max |diff| = {d.max()}; {np.count_nonzero(d > 1)} out of tolerance"


def test_chunking_is_invisible(built):
    """The stream is the same however the caller slices it into render calls."""
    H = built
    conf = H.mode_config("i", vfilter=True)
    a = H.Encoder(conf, 16000000); a.open_test_source()
    whole = a.render_host(1500); a.close()
    b = H.Encoder(conf, 16000000); b.open_test_source()
    parts = np.concatenate([b.render_host(n) for n in (1, 311, 313, 625, 250)]); b.close()
    assert np.array_equal(whole, parts)


def test_next_line_view_matches_batches(built):
    H = built
    conf = H.mode_config("pal")
    a = H.Encoder(conf, 16000000); a.open_test_source()
    ref = a.render_host(700).reshape(700, -1); a.close()
    b = H.Encoder(conf, 16000000); b.open_test_source()
    for k in range(700):
        iq, frame, line = b.next_line()
        assert frame == k // 625 + 1 and line == k % 625 + 1
        assert np.array_equal(iq[0::2], ref[k]) and not iq[1::2].any()
    b.close()


def test_secam_random_pictures_exact(built):
    """SECAM chroma on random pictures (a new one every frame), rendered in uneven pieces so
    the cross-line state (IIR state, the two aliased buffer words) crosses launch boundaries."""
    H = built
    rng = np.random.default_rng(99)
    conf = H.mode_config("l", vfilter=True, noaudio=True)
    enc = H.Encoder(conf, 16000000)
    frames = rng.integers(0, 1 << 24, size=(3, enc.active_lines, enc.active_width), dtype=np.uint32)
    enc.set_source(frames, None)
    got = np.concatenate([enc.render_host(n) for n in (1, 2, 23, 300, 625, 949)])
    enc.close()
    o = orc.Oracle(conf, 16000000)
    o.set_source(frames, None)
    want = o.render(1900)
    o.close()
    assert _diff(got, want).max() == 0


def test_offset_mixer(built):
    """--offset: the complex NCO multiply of ref video.c:3482-3515 incl. its INT16_MAX start-up
    (the first 32767 samples are scaled to ~0). Closed-form NCO: +-1 LSB on top of the sound carriers'."""
    got, want = _pair(built, "i", 16000000, 700, vfilter=True, offset=2000000)
    d = _diff(got, want)
    assert d.max() <= 2, f"max |diff| = {d.max()}"
    assert (d <= 1).mean() > 0.999
    got, want = _pair(built, "i", 16000000, 700, vfilter=True, offset=-3500000, noaudio=True)
    assert _diff(got, want).max() <= 1


def test_memory_source_equals_callback_source(built):
    """htv_av_memory_open (C) and Python callbacks feed the same stream."""
    H = built
    rng = np.random.default_rng(5)
    conf = H.mode_config("i", vfilter=True)
    a = H.Encoder(conf, 16000000)
    frames = rng.integers(0, 1 << 24, size=(2, a.active_lines, a.active_width), dtype=np.uint32)
    audio = rng.integers(-8000, 8000, size=(5000, 2), dtype=np.int16)
    a.set_source(frames, audio, audio_block=777)
    x = a.render_host(1400); a.close()
    b = H.Encoder(conf, 16000000)
    b.open_memory_source(frames, audio, audio_block=1000)
    y = b.render_host(1400); b.close()
    assert np.array_equal(x, y)
