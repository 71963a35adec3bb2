"""e2e (htv_render_host, a ring of 8 pinned pictures, every frame uploaded) against the copy-back piece size."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hacktv_b200 as H
RATE, FRAMES, STEPS = 16000000, 64, 10
conf = H.mode_config("i", vfilter=True)
for mb in (4, 8, 12, 16, 24, 32):
    os.environ["HTV_HOST_PIECE_MB"] = str(mb)
    enc = H.Encoder(conf, RATE)
    one = H.test_pattern(enc.active_width, enc.active_lines).astype(np.int32)
    pic = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(one, (8,) + one.shape))).pin_memory()
    tone = torch.from_numpy(H.test_tone()).pin_memory()
    enc.open_memory_source(pic.numpy().view(np.uint32), tone.numpy(), audio_block=8192, static_video=False)
    lines = FRAMES * enc.lines
    host = torch.empty(lines * enc.width * 2, dtype=torch.int16).pin_memory()
    for _ in range(3):
        enc.render_host_ptr(lines, host.data_ptr())
    t0 = time.perf_counter()
    for _ in range(STEPS):
        enc.render_host_ptr(lines, host.data_ptr())
    dt = (time.perf_counter() - t0) / STEPS
    enc.close()
    print(json.dumps({"piece_mb": mb, "ms_per_64_frames": round(dt * 1e3, 3), "msamples_per_s": round(lines * 1024 / dt / 1e6, 1)}), flush=True)
