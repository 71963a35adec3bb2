#!/bin/bash
# Profile one PAL-I 16 Msps --filter step (64 frames) of the shipped build: launch list, ncu --set full of the
# dominant kernel(s) (regex in $1, default k_line), then an untimed-by-profiler bench line.
#   gpurun --timeout 900 -- 'bash tools/profile_step.sh k_line r02a'
set -u
K=${1:-k_line}; TAG=${2:-r02}
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_pal_i_64frames.csv \
	python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_$K \
	python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
ls -la gpurun_out | tail -8
