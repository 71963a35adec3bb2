"""The tensor-core video filter's index arithmetic (hacktv_b200/csrc/htv_mma_fir.h) checked on
the CPU: tests/mma_fir_emu.c drives a lane-by-lane model of mma.sync.m16n8k32 (PTX ISA fragment
tables) with the very helpers k_mod_mma uses and compares with the direct 51-tap int32 sum of
the reference (ref fir.c:564-615), including int32 wrap-around with extreme taps."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("mma") / "mma_fir_emu")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "hacktv_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "mma_fir_emu.c")])
    return exe


@pytest.mark.parametrize("W", [1024, 1280, 1152, 1536, 128])
@pytest.mark.parametrize("extreme", [0, 1])
def test_fragment_mapping_reproduces_the_fir(emu, W, extreme):
    for seed in (1, 2, 3):
        out = subprocess.run([emu, str(W), str(seed), str(extreme)], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.startswith("OK %d" % W), out.stdout


@pytest.mark.parametrize("W", [858, 864, 1287, 1024, 1135, 129])
def test_pitched_plane_layout_for_any_width(emu, W):
    """Rows built the way k_raster's pitched branch writes them (own samples + halos in the neighbouring
    rows): NTSC 858, 13.5 Msps PAL 864, the 20.25 Msps NTSC width 1287, odd widths."""
    for seed in (1, 2):
        for extreme in (0, 1):
            out = subprocess.run([emu, str(W), str(seed), str(extreme), "1"], capture_output=True, text=True)
            assert out.returncode == 0 and out.stdout.startswith("OK %d" % W), out.stdout


@pytest.fixture(scope="module")
def lp_emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lp") / "lp_mma_emu")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "hacktv_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "lp_mma_emu.c")])
    return exe


@pytest.mark.parametrize("W", [1024, 1280, 858, 864, 967, 128])
@pytest.mark.parametrize("ntaps", [11, 13, 15, 17])
def test_short_low_pass_and_lane_layout(lp_emu, W, ntaps):
    """k_line's chroma low-pass (11 .. 17 taps by sample rate) and k_sec_raster's 15-tap baseband low-pass: one k-step of
    32, the lane's four samples x = 128 nt + 32 t + g + 8 j in accumulator order - every sample owned once, the sums equal
    the zero-padded direct form including int32 wrap-around."""
    for seed in (1, 2):
        for extreme in (0, 1):
            out = subprocess.run([lp_emu, str(W), str(ntaps), str(seed), str(extreme)], capture_output=True, text=True)
            assert out.returncode == 0 and out.stdout.startswith("OK %d" % W), out.stdout
