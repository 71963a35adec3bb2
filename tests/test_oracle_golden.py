"""The oracle against the committed golden fixtures (made from the UNMODIFIED reference by
tests/golden/make_golden.py): bit-exact, every BASELINE.json config, start of stream and
>= 10 s in. This is what pins the oracle; it runs without a GPU and without the reference."""
import hashlib
import json
import os

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def _oracle(built, g):
    conf = built.mode_config(g["mode"], vfilter=g["filter"])
    o = orc.Oracle(conf, g["rate"])
    o.open_test_source()
    return o


@pytest.mark.parametrize("name", sorted(GOLD))
def test_first_four_frames_bit_exact(built, name):
    g = GOLD[name]
    o = _oracle(built, g)
    a = o.render(g["a_lines"])
    o.close()
    assert hashlib.sha256(a.tobytes()).hexdigest() == g["a_sha256"]
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    a2 = a.reshape(g["a_lines"], g["values_per_line"])
    assert np.array_equal(a2[z["a_lines"]], z["a"])


# the late window costs ~15 s of oracle time per config: the three sound-carrier types
@pytest.mark.parametrize("name", ["cfg2_i_16M_filter", "cfg3_m_13M5_filter", "cfg4_l_16M_filter"])
def test_after_ten_seconds_bit_exact(built, name):
    g = GOLD[name]
    o = _oracle(built, g)
    chunk = 20000
    left = g["b_skip"]
    while left > 0:                      # the oracle is a stream: render and discard
        n = min(chunk, left)
        o.render(n)
        left -= n
    b = o.render(g["b_lines"])
    o.close()
    assert hashlib.sha256(b.tobytes()).hexdigest() == g["b_sha256"]
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert np.array_equal(b.reshape(g["b_lines"], -1)[z["b_lines"]], z["b"])


GOLD_PR = json.load(open(os.path.join(HERE, "golden", "golden_pixelrate.json")))


@pytest.mark.parametrize("name", sorted(GOLD_PR))
def test_pixelrate_resampler_bit_exact(built, name):
    """--pixelrate: raster at the pixel rate + the reference's polyphase resampler (fixtures from
    tests/golden/make_golden_pixelrate.py)."""
    g = GOLD_PR[name]
    o = orc.Oracle(built.mode_config(g["mode"], vfilter=g["filter"]), g["rate"], g["pixel_rate"])
    o.open_test_source()
    a = o.render(g["lines"])
    o.close()
    assert hashlib.sha256(a.tobytes()).hexdigest() == g["sha256"]
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert np.array_equal(a.reshape(g["lines"], g["values_per_line"])[z["lines"]], z["a"])
