#!/bin/bash
# The first GPU call after round 1 (everything written after its GPU minutes were spent, plus the
# captures that are stale for the tensor-core build). Each step is bounded; outputs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
echo "== never-run paths: pitched planes (NTSC), W = 1152, --pixelrate device path"
HTV_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_zz_mma_fir.py tests/test_gpu_zz_pixelrate.py -m gpu -q -x \
	> gpurun_out/unvalidated_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/unvalidated_tests.txt; tail -5 gpurun_out/unvalidated_tests.txt
echo "== A/B: HTV_FIR scalar vs mma incl. the widths 128 does not divide"
timeout 200 python tools/ab_check.py HTV_FIR scalar mma --ntsc 2>&1 | tail -12
echo "== A/B: HTV_SIDE default vs split"
timeout 200 python tools/ab_check.py HTV_SIDE default split 2>&1 | tail -12
echo "== launch list of the default build (PAL-I 16 Msps --filter, 64 frames)"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pal_i_64frames.csv \
	python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
echo "== ncu --set full: k_raster and k_mod_mma, one launch each"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 2 -c 1 -f -o gpurun_out/k_raster \
	python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_mod_mma -s 2 -c 1 -f -o gpurun_out/k_mod_mma \
	python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
ls -la gpurun_out
