#!/usr/bin/env python
"""Golden fixtures for --pixelrate from the UNMODIFIED reference
(oracle/_ref/ref_harness, zero-heap variant as in make_golden.py): sha256 of the first two
frames and the raw samples of a few lines, written to golden_pixelrate.json / pixelrate_*.npz.

    python tests/golden/make_golden_pixelrate.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

CONFIGS = {
    # name: (mode, sample rate, pixel rate, filter, lines)
    "pixelrate_i_16M_from_13M5_filter": ("i", 16000000, 13500000, True, 1250),
    "pixelrate_l_16M_from_13M5_filter": ("l", 16000000, 13500000, True, 700),
    "pixelrate_m_13M5_from_9M_filter": ("m", 13500000, 9000000, True, 600),
}
KEEP = [0, 1, 2, 22, 23, 24, 311, 312, 313, 599]


def main():
    assert orc.have_ref(), "build the reference first: make -C oracle ref"
    index = {}
    for name, (mode, rate, prate, filt, n) in CONFIGS.items():
        a = orc.run_ref(mode, rate, n, vfilter=filt, extra=("--pixelrate", str(prate)))
        per = a.size // n
        np.savez_compressed(os.path.join(HERE, name + ".npz"), lines=np.array(KEEP), a=a.reshape(n, per)[KEEP])
        index[name] = {"mode": mode, "rate": rate, "pixel_rate": prate, "filter": filt, "lines": n,
                       "sha256": hashlib.sha256(a.tobytes()).hexdigest(), "values_per_line": per}
        print(name, index[name]["sha256"][:16])
    with open(os.path.join(HERE, "golden_pixelrate.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
