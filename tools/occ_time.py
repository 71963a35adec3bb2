"""Device-resident timing of the BASELINE configs with whatever library HTV_LIB names, and a checksum of one call's
output (the variants must be bit-identical)."""
import sys, zlib, torch
sys.path.insert(0, '.')
import hacktv_b200 as H
for name, mode, rate, filt in (("cfg2 i 16M --filter", "i", 16000000, True), ("cfg5 i 20M --filter", "i", 20000000, True),
                               ("cfg3 m 13.5M --filter", "m", 13500000, True), ("cfg1 pal 16M", "pal", 16000000, False),
                               ("cfg4 l 16M --filter", "l", 16000000, True)):
    enc = H.Encoder(H.mode_config(mode, vfilter=filt), rate); enc.open_test_source()
    n = 64 * enc.lines
    out = torch.empty(n * enc.width * (2 if enc.complex else 1), dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): enc.render(n, out.data_ptr(), st)
    torch.cuda.synchronize()
    crc = zlib.crc32(out.cpu().numpy().tobytes())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): enc.render(n, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fps = 25 if enc.lines == 625 else 30000 / 1001
    print(f"{name:28s} W {enc.width:5d} ms/64 frames {ms:7.4f} realtime_x {64 / fps / (ms / 1e3):8.1f} crc {crc:08x}")
    enc.close()
