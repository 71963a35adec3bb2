HTV_DEBUG=1 python tools/run_one.py l 16000000 1 64 2>&1 | tail -2
python -m pytest tests -m gpu -x -q -k "secam or long or parity or chunk or vbi or dropin or cabi or pixelrate" 2>&1 | tail -6
ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/q_secam.csv python tools/run_one.py l 16000000 1 64 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_secam.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[len(rows)//2+1:]: print(r[ki].split('(')[0][:40], r[vi])
PY
python tools/sec_time.py
