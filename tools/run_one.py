import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hacktv_b200 as H
mode, rate, filt, frames = sys.argv[1], int(sys.argv[2]), sys.argv[3] == "1", int(sys.argv[4])
enc = H.Encoder(H.mode_config(mode, vfilter=filt), rate); enc.open_test_source()
n = frames * enc.lines
out = torch.empty(n * enc.width * 2, dtype=torch.int16, device="cuda")
for _ in range(2):
    enc.render(n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
