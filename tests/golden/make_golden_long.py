#!/usr/bin/env python
"""Long-stream fixtures from the UNMODIFIED reference (oracle/_ref/ref_harness): for BASELINE configs
2, 3 and 4 a set of whole lines out of a frame that starts >= 10 s into the stream (window b of
golden.json) and out of one that starts >= 34 s in - past the wrap of every device-side ring of the
CUDA path (audio 2^20 pairs = 32.8 s, NICAM symbols 2^23 = 23 s, NICAM frames 2^14 = 16.4 s), many
NCO renormalisations, five loops of the 6.4 s test tone (ref av_test.c:156-196).
tests/test_gpu_long_stream.py renders that far on the device and compares.

    python tests/golden/make_golden_long.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

CONFIGS = {
    "cfg2_i_16M_filter": ("i", 16000000, True, 625, ()),
    "cfg3_m_13M5_filter": ("m", 13500000, True, 525, ()),
    "cfg4_l_16M_filter": ("l", 16000000, True, 625, ()),
    # config 3 without its sound carrier: everything left is integer work, so the GPU must match bit for bit at any
    # distance into the stream (the FM carrier of config 3 sits at exactly fs / 3, see tests/test_gpu_long_stream.py)
    "cfg3_m_13M5_filter_noaudio": ("m", 13500000, True, 525, ("--noaudio",)),
}
KEEP = [0, 1, 2, 3, 5, 22, 23, 24, 100, 200, 262, 263, 264, 285, 300, 310, 311, 312, 313, 314, 335, 336, 400, 500, 600, 622, 623, 624]


def main():
    assert orc.have_ref(), "build the reference first: make -C oracle ref"
    index = json.load(open(os.path.join(HERE, "golden_long.json"))) if os.path.exists(os.path.join(HERE, "golden_long.json")) else {}
    only = sys.argv[1:]
    for name, (mode, rate, filt, lpf, extra) in CONFIGS.items():
        if only and name not in only:
            continue
        fps = 25.0 if lpf == 625 else 30000 / 1001
        keep = [l for l in KEEP if l < lpf]
        entry = {"mode": mode, "rate": rate, "filter": filt, "noaudio": "--noaudio" in extra, "lines_per_frame": lpf, "keep": keep, "windows": {}}
        arrays = {}
        for tag, seconds in (("b", 10.0), ("c", 34.0)):
            skip = int(np.ceil(seconds * fps)) * lpf
            x = orc.run_ref(mode, rate, lpf, skip=skip, vfilter=filt, extra=extra, timeout=900)
            x = x.reshape(lpf, -1)
            arrays[tag] = x[keep]
            entry["windows"][tag] = {"skip": skip, "values_per_line": int(x.shape[1])}
            print(name, tag, skip, x.shape)
        np.savez_compressed(os.path.join(HERE, "long_" + name + ".npz"), **arrays)
        index[name] = entry
    with open(os.path.join(HERE, "golden_long.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
