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
