"""GPU parity for the channel combiner (ref --passthru,
_vid_passthru_process video.c:3517-3541): the external-stream form against the oracle
(itself pinned to the reference in test_oracle_vs_ref.py::test_passthru_alignment), the
device-resident forms (htv_render_add, htv_mix_add) against the same sums."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu

# (mode, rate, lines, overrides, external lines, tolerance)
PASSTHRU = [
    ("pal", 16000000, 700, dict(), 800, 0),                                # real output: I only
    ("i", 16000000, 700, dict(vfilter=True, noaudio=True), 800, 0),        # delay 1 line, exact path
    ("i", 20000000, 500, dict(vfilter=True, offset=1250000, level=0.5), 600, 2),   # README recipe, stage 2:
    #   +-1 LSB from the sound carriers, then +-1 from the offset NCO (as test_gpu_parity.test_offset_mixer)
    ("l", 16000000, 700, dict(vfilter=True), 800, 1),
    ("m", 13500000, 600, dict(vfilter=True), 100, 1),                      # external stream ends early
    ("i", 16000000, 9000, dict(vfilter=True, noaudio=True), 8500, 0),      # > one staging buffer (4096 lines)
]


@pytest.mark.parametrize("mode,rate,nlines,kw,ext_lines,tol", PASSTHRU)
def test_passthru_parity(built, mode, rate, nlines, kw, ext_lines, tol):
    H = built
    rng = np.random.default_rng(3)
    conf = H.mode_config(mode, **kw)
    enc = H.Encoder(conf, rate)
    o = orc.Oracle(conf, rate)
    ext = rng.integers(-32768, 32767, size=(ext_lines * enc.width + enc.width // 3, 2), dtype=np.int16)
    assert enc.passthru_delay_lines == (1 if kw.get("vfilter") else 0)
    enc.open_test_source(); o.open_test_source()
    enc.set_passthru(ext); o.set_passthru(ext)
    got = enc.render_host(nlines)
    want = o.render(nlines)
    enc.close(); o.close()
    d = (got.astype(np.int32) - want.astype(np.int32) + 32768) % 65536 - 32768     # the sum may wrap
    assert np.abs(d).max() <= tol, f"max |diff| {np.abs(d).max()}, {np.count_nonzero(d)} values differ"


def test_passthru_must_precede_rendering(built):
    H = built
    enc = H.Encoder(H.mode_config("pal"), 16000000)
    enc.open_test_source()
    enc.render_host(3)
    with pytest.raises(RuntimeError):
        enc.set_passthru(np.zeros((4096, 2), dtype=np.int16))
    enc.close()


def test_two_channels_into_one_buffer(built):
    """The README's recipe (`--offset -6.75e6 --level 0.5 --filter -o - | --offset 1.25e6 --level 0.5
    --passthru /dev/stdin --filter`, README:89-90) without the stream leaving HBM: channel A is rendered
    to device memory, channel B is added into it one line later (the reference's alignment)."""
    import torch
    H = built
    rate, n = 20000000, 400
    ca = H.mode_config("i", vfilter=True, offset=-6750000, level=0.5)
    cb = H.mode_config("i", vfilter=True, offset=1250000, level=0.5)
    a = H.Encoder(ca, rate); a.open_test_source()
    b = H.Encoder(cb, rate); b.open_test_source()
    W = a.width
    buf = torch.zeros((n + 1) * W * 2, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    a.render(n + 1, buf.data_ptr(), st)
    b.render_add(n, buf.data_ptr() + b.passthru_delay_lines * W * 4, st)
    torch.cuda.synchronize()
    got = buf[W * 2:].cpu().numpy()
    a.close(); b.close()

    oa = orc.Oracle(ca, rate); oa.open_test_source()
    sa = oa.render(n + 1); oa.close()
    ob = orc.Oracle(cb, rate); ob.open_test_source()
    ob.set_passthru(sa.reshape(-1, 2))
    want = ob.render(n); ob.close()
    d = (got.astype(np.int32) - want.astype(np.int32) + 32768) % 65536 - 32768
    assert np.abs(d).max() <= 2, np.abs(d).max()                  # +-1 LSB per channel (NCO closed forms)
    assert np.count_nonzero(d) < d.size // 20


@pytest.mark.parametrize("nvalues", [0, 5, 8, 4099, 1 << 20, (1 << 22) + 3])
def test_mix_add_wraps_like_int16(built, nvalues):
    import torch
    H = built
    rng = np.random.default_rng(nvalues)
    x = rng.integers(-32768, 32767, size=nvalues, dtype=np.int16)
    y = rng.integers(-32768, 32767, size=nvalues, dtype=np.int16)
    tx = torch.from_numpy(x).cuda(); ty = torch.from_numpy(y).cuda()
    H.mix_add(tx.data_ptr(), ty.data_ptr(), nvalues, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(tx.cpu().numpy(), (x.astype(np.int32) + y).astype(np.int16))
    assert np.array_equal(ty.cpu().numpy(), y)
