"""The C-ABI driven from C, as hacktv itself would: the htv_cli front end (htv_init -> test source ->
htv_next_line -> htv_rf_write into htv_rf_file_open's int16 file sink, ref hacktv.c:1440-1601, rf_file.c) and
the in-process multi-device test (tests/c/multi_device.c: one pthread per encoder, htv_init_on)."""
import json
import os
import subprocess

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "hacktv_b200", "htv_cli")
MULTI = os.path.join(ROOT, "tests", "c", "multi_device")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("args,mode,rate,kw,tol", [
    (["-m", "i", "-s", "16000000", "--filter"], "i", 16000000, dict(vfilter=True), 1),        # complex file
    (["-m", "pal", "-s", "16000000"], "pal", 16000000, dict(), 0),                           # real file: Q dropped
    (["-m", "m", "-s", "13500000", "--filter", "--noaudio"], "m", 13500000, dict(vfilter=True, noaudio=True), 0),
])
def test_cli_file_sink_matches_the_oracle(built, tmp_path, args, mode, rate, kw, tol):
    H = built
    assert os.path.exists(CLI), "htv_cli was not built"
    out = tmp_path / "out.bin"
    lines = 1300
    r = subprocess.run([CLI] + args + ["-o", str(out), "--lines", str(lines), "test"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Sample rate:" in r.stderr
    got = np.fromfile(out, dtype=np.int16)
    o = orc.Oracle(H.mode_config(mode, **kw), rate); o.open_test_source()
    want = o.render(lines); o.close()
    assert got.size == want.size                                  # real modes: one int16 per sample in the file
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= tol


def test_cli_writes_to_stdout(built):
    r = subprocess.run(f"{CLI} -m i -s 16000000 --filter -o - --lines 100 test 2>/dev/null | wc -c", shell=True,
                       capture_output=True, text=True, timeout=120)
    assert int(r.stdout.strip()) == 100 * 1024 * 4


def test_multi_device_pthreads(built):
    """N host threads x N encoders placed with htv_init_on, bit-identical to the single-threaded run; on a
    one-GPU box the encoders share device 0 (still concurrent calls from different threads)."""
    assert os.path.exists(MULTI), "tests/c/multi_device was not built (make -C hacktv_b200/csrc)"
    r = subprocess.run([MULTI, "4", "900"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["mismatching_threads"] == 0 and info["threads"] == 4
