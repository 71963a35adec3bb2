"""Where the end-to-end time goes: the same 64-frame PAL-I step (bench.py's workload) with and
without per-frame picture uploads and with and without the copy back to the host."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hacktv_b200 as H

RATE, FRAMES, STEPS = 16000000, 64, 8
conf = H.mode_config("i", vfilter=True)


def run(live, to_host):
    enc = H.Encoder(conf, RATE)
    pic = torch.from_numpy(H.test_pattern(enc.active_width, enc.active_lines).astype(np.int32)).pin_memory()
    tone = torch.from_numpy(H.test_tone()).pin_memory()
    enc.open_memory_source(pic.numpy().view(np.uint32)[None], tone.numpy(), audio_block=8192, static_video=not live)
    lines = FRAMES * enc.lines
    host = torch.empty(lines * enc.width * 2, dtype=torch.int16).pin_memory()
    dev = torch.empty(lines * enc.width * 2, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def step():
        if to_host:
            enc.render_host_ptr(lines, host.data_ptr())
        else:
            enc.render(lines, dev.data_ptr(), st)
            torch.cuda.synchronize()
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(STEPS):
        step()
    dt = (time.perf_counter() - t0) / STEPS
    enc.close()
    return round(dt * 1e3, 3)


print(json.dumps({"ms_per_64_frames": {
    "device_only_static": run(False, False), "device_only_live_uploads": run(True, False),
    "to_host_static": run(False, True), "to_host_live_uploads": run(True, True)}}))
