"""Host logic without a GPU: the product's table generation against the oracle's and
against the reference's known-answer values, the test source, and the
C-ABI surface (every symbol the header declares is exported; no compute is attempted)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = ["sync0", "sync1", "sync2", "sync3", "sync4", "sync_off", "burst_win", "chroma_taps", "vsb_itaps",
          "vsb_qtaps", "levels", "nicam_taps", "geometry", "secam_lpf", "secam_notch"]
MODES = [("i", 16000000, True), ("pal", 16000000, False), ("m", 13500000, True), ("l", 16000000, True),
         ("i", 20000000, True), ("b", 16000000, True), ("ntsc", 13500000, False), ("pal", 16000000, True),
         ("pal-n", 16000000, True), ("secam", 16000000, False), ("pal60", 13500000, False)]


@pytest.mark.parametrize("mode,rate,filt", MODES)
def test_product_tables_equal_oracle_tables(built, mode, rate, filt):
    conf = built.mode_config(mode, vfilter=filt)
    t, o = built.Tables(conf, rate), orc.Oracle(conf, rate)
    for name in TABLES:
        a, b = t.get(name), o.table(name)
        assert (a is None) == (b is None), name
        if a is not None:
            assert a.shape == b.shape and np.array_equal(a, b), name
    t.close(); o.close()


def test_known_answers_pal_i_16M(built):
    """Values dumped from the reference's vid_init."""
    t = built.Tables(built.mode_config("i", vfilter=True), 16000000)
    assert list(t.get("geometry")) == [1024, 512, 166, 832, 87, 46]
    assert list(t.get("levels")) == [4653, 17681, 17681, 23265]
    assert list(t.get("sync0")[:8]) == [94, 558, 1508, 2791, 4075, 5025, 5489, 5583]
    assert list(t.get("sync0")[-4:]) == [1747, 711, 150, 3] and len(t.get("sync0")) == 83
    assert list(t.get("sync_off")) == [-3, -3, -3, 509, 509]
    assert [len(t.get(f"sync{i}")) for i in range(5)] == [83, 45, 444, 45, 444]
    assert list(t.get("chroma_taps")) == [3, 34, 249, 1179, 3583, 6978, 8715, 6978, 3583, 1179, 249, 34, 3]
    bw = t.get("burst_win")
    assert list(bw[:10]) == [0, -18, -138, -420, -865, -1409, -1950, -2388, -2663, -2776] and bw[20] == -2792
    assert list(t.get("vsb_itaps")[:26]) == [-1, 1, -8, -12, 3, -23, -3, 61, 3, 74, 165, -46, 46, 64, -431, -182,
                                              -151, -792, 153, 416, -317, 1956, 1975, -236, 6759, 13824]
    assert list(t.get("vsb_qtaps")[:26]) == [2, 2, -3, 7, -12, -34, 0, -50, -57, 74, -8, 56, 310, 43, 108, 340,
                                              -422, -328, -72, -1370, -529, 389, -1464, 2396, 7458, 0]
    nt = t.get("nicam_taps")
    assert len(nt) == 221 and list(nt[107:114]) == [1013, 1024, 1030, 1033, 1030, 1024, 1013] and list(nt[:3]) == [1, 1, 1]


def test_known_answers_ntsc_and_secam(built):
    t = built.Tables(built.mode_config("m", vfilter=True), 13500000)
    assert list(t.get("chroma_taps")) == [4, 70, 622, 2963, 7559, 10329, 7559, 2963, 622, 70, 4]
    assert list(t.get("burst_win")[:9]) == [0, -34, -250, -734, -1427, -2160, -2742, -3063, -3152]
    assert list(t.get("levels")) == [3154, 17740, 18923, 25231]
    t = built.Tables(built.mode_config("l", vfilter=True), 16000000)
    assert list(t.get("geometry"))[4:] == [82, 944]
    assert list(t.get("burst_win")[:10]) == [0, 6, 47, 157, 367, 705, 1195, 1854, 2693, 3719]
    assert list(t.get("secam_lpf")[:8]) == [-9, -68, -64, 400, 1810, 4138, 6456, 7440]
    assert list(t.get("secam_notch")[23:28]) == [4657, 680, 27307, 680, 4657]
    assert list(t.get("levels")) == [21140, 6342, 6342, 1057]


def test_known_answers_yuv(built):
    o = orc.Oracle(built.mode_config("i"), 16000000)
    for rgb, want in {0x000000: (17681, 0, 0), 0xFFFFFF: (4653, 0, 0), 0xBF0000: (14763, 1438, -5999),
                      0x00BF00: (11953, 2824, 5024), 0x0000BF: (16569, -4262, 976), 0x808080: (11141, 0, 0)}.items():
        assert tuple(o.yuv(rgb)) == want
    o.close()


def test_test_source_matches_oracle(built):
    for w, h in ((832, 576), (1039, 576), (715, 480)):
        assert np.array_equal(built.test_pattern(w, h), orc.test_pattern(w, h))
    assert np.array_equal(built.test_tone(), orc.test_tone())
    tone = built.test_tone()
    assert tone.shape == (204800, 2) and not tone[:20480, 0].any() and tone[1:20480, 1].any()


def test_mode_table(built):
    m = built.modes()
    for want in ("i", "pal", "l", "m", "b", "g", "ntsc", "secam", "pal-m", "pal-n"):
        assert want in m
    c, desc = m["i"]
    assert c.lines == 625 and c.colour_mode == 1 and c.modulation == 2 and c.fm_mono_carrier == 5999600
    assert "6.0 MHz FM audio" in desc
    assert m["m"][0].nicam_carrier == 0 and m["l"][0].am_mono_carrier == 6500000


def test_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "hacktv_b200.h")).read()
    names = set(re.findall(r"extern\s+[^;(]*?\b(htv_[a-z0-9_]+)\s*\(", hdr)) | {"htv_modes"}
    assert len(names) > 30
    L = C.CDLL(built.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert built.lib().htv_config_size() == C.sizeof(built.Config) == orc.lib().orc_params_size()


def test_product_never_links_the_oracle(built):
    import subprocess
    out = subprocess.run(["ldd", built.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    nm = subprocess.run(["nm", "-D", built.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in nm


def test_unsupported_configs_fail_loudly(built):
    conf = built.mode_config("pal-fm")
    conf.fm_energy_dispersal = 0.0625    # FM energy dispersal: not on the accelerated path
    with pytest.raises(RuntimeError):
        built.Tables(conf, 20000000)
    conf = built.mode_config("i")
    conf.type = 3                # 819-line raster
    with pytest.raises(RuntimeError):
        built.Tables(conf, 16000000)


def test_encoder_needs_a_gpu_or_says_so(built):
    """No CPU fallback: on a box without CUDA, construction fails with an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        built.Encoder("i", 16000000)


def test_fm_video_tables_match_the_oracle(built):
    """FM video (ref video.c:3678-3740, 4564-4570): pre-emphasis taps per sample rate."""
    for mode, rate, n in (("pal-fm", 20000000, 67), ("pal-fm", 20250000, 67), ("pal-fm", 14000000, 67),
                          ("ntsc-fm", 18000000, 67), ("ntsc-fm", 20250000, 71)):   # 28 Msps: line wider than the kernels' 1536
        conf = built.mode_config(mode, vfilter=True)
        t = built.Tables(conf, rate)
        o = orc.Oracle(conf, rate)
        a, b = t.get("fmv_taps"), o.table("vsb_itaps")
        assert a is not None and len(a) == n and np.array_equal(a, b), (mode, rate)
        o.close()


@pytest.mark.parametrize("mode,rate,prate,filt", [("i", 16000000, 13500000, True), ("i", 20000000, 13500000, False),
                                                  ("i", 13500000, 16000000, True), ("m", 13500000, 9000000, True),
                                                  ("i", 16000000, 14000000, True), ("l", 16000000, 13500000, True)])
def test_pixelrate_resampler_tables_equal_the_oracle(built, mode, rate, prate, filt):
    """htv_tables_create2: the --pixelrate resampler's polyphase taps and geometry (ref fir.c:263-295,
    393-428) against the oracle, which is pinned to the reference's --pixelrate output."""
    import orc
    conf = built.mode_config(mode, vfilter=filt)
    t = built.Tables(conf, rate, prate)
    o = orc.Oracle(conf, rate, prate)
    assert np.array_equal(t.get("rs_taps"), o.table("rs_taps"))
    g, og = t.get("rs_geometry"), o.table("rs_geometry")
    assert np.array_equal(g[:4], og)
    assert g[4] == (2 if filt else 1) * o.width          # sound carriers, offset mixer, passthru: lines of lead
    plain = built.Tables(conf, rate)
    for name in ("vsb_itaps", "vsb_qtaps", "nicam_taps", "levels"):
        a, b = t.get(name), plain.get(name)
        assert (a is None and b is None) or np.array_equal(a, b)
    t.close(); o.close(); plain.close()


def test_pixelrate_pairs_that_are_refused(built):
    with pytest.raises(RuntimeError):
        built.Tables(built.mode_config("i"), 16000000, 13400000)             # line width would vary
    with pytest.raises(RuntimeError):
        built.Tables(built.mode_config("pal-fm"), 20000000, 13500000)        # FM video
    t = built.Tables(built.mode_config("i"), 16000000, 16000000)             # no resampler at all
    assert t.get("rs_taps") is None
    t.close()
