HTV_DEBUG=1 python tools/run_one.py l 16000000 1 64 2>&1 | tail -4
python -m pytest tests -m gpu -x -q -k "secam or long or parity or chunk or vbi or dropin or cabi" 2>&1 | tail -6
ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/q_secam.csv python tools/run_one.py l 16000000 1 64 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_secam.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki].split('(')[0][:40], r[vi])
PY
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import hacktv_b200 as H
for mode, rate in (("l", 16000000), ("l", 13500000)):
    enc = H.Encoder(H.mode_config(mode, vfilter=True), rate); enc.open_test_source()
    n = 64 * enc.lines
    out = torch.empty(n * enc.width * 2, dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): enc.render(n, out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): enc.render(n, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(mode, rate, "ms per 64 frames", round(ms, 4), "realtime_x", round(64 / 25 / (ms / 1e3), 1))
PY
