# A/B of the pre-pass running ahead (HTV_AHEAD) + full GPU test suite + SECAM list
python tools/ab_check.py HTV_AHEAD 0 default --ntsc 2>&1 | grep -v "\"parity\".*\"ok\": true"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
HTV_DEBUG=1 python tools/run_one.py l 16000000 1 64 2>&1 | tail -2
ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/q_secam.csv python tools/run_one.py l 16000000 1 64 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_secam.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[len(rows)//2+1:]: print(r[ki].split('(')[0][:40], r[vi])
PY
