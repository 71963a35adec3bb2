"""HTV_DEBUG_PIPE timeline of one htv_render_host call (render / copy windows per piece)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hacktv_b200 as H
conf = H.mode_config("i", vfilter=True)
enc = H.Encoder(conf, 16000000)
pic = torch.from_numpy(H.test_pattern(enc.active_width, enc.active_lines).astype(np.int32)).pin_memory()
tone = torch.from_numpy(H.test_tone()).pin_memory()
enc.open_memory_source(pic.numpy().view(np.uint32)[None], tone.numpy(), audio_block=8192, static_video=False)
lines = 64 * enc.lines
host = torch.empty(lines * enc.width * 2, dtype=torch.int16).pin_memory()
for _ in range(3):
    enc.render_host_ptr(lines, host.data_ptr())
os.environ["HTV_DEBUG_PIPE"] = "1"
enc.render_host_ptr(lines, host.data_ptr())
