#!/bin/bash
# quick GPU check of a k_line change: A/B parity + timing against the split kernels, then the launch list
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py HTV_PATH split default --ntsc 2>&1 | grep -v "\"parity\".*\"ok\": true" | grep -v "timing.*split"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/q_launches.csv python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
python3 - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[len(rows)//2+1:]:
    try: print(f"{r[ki][:40]:42s} {float(r[vi].replace(',',''))/1000:8.1f} us")
    except: pass
PY
