#!/bin/bash
# Round-2 profile set of the shipped build (PAL-I 16 Msps --filter, 64 frames per call; SECAM-L for the chain):
#   gpurun --timeout 1200 -- 'bash tools/profile_r02.sh'
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== launch lists (gpu__time_duration.sum)"
timeout 150 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r02_launches_pal_i_64frames.csv python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
timeout 150 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r02_launches_secam_l_64frames.csv python tools/run_one.py l 16000000 1 64 > /dev/null 2>&1
echo "== ncu --set full: k_line, k_line_desc_a2, the SECAM raster / chain / output kernels"
timeout 200 $NCU --set full --import-source on -k "regex:^k_line$" -s 1 -c 1 -f -o gpurun_out/r02_k_line python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
timeout 200 $NCU --set full --import-source on -k "regex:^k_line_desc_a2$" -s 1 -c 1 -f -o gpurun_out/r02_k_line_desc_a2 python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on -k "regex:^k_sec_raster$|^k_sec_pass0$|^k_sec_predict$|^k_sec_refine$|^k_sec_fm_list$|^k_sec_out$" -s 6 -c 6 -f -o gpurun_out/r02_secam python tools/run_one.py l 16000000 1 64 > /dev/null 2>&1
echo "== whole step with the caches left alone between kernels (--cache-control none): live DRAM / L2 traffic of every kernel"
timeout 200 $NCU --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -c 40 --csv \
	--log-file gpurun_out/r02_step_traffic_cache_control_none.csv python tools/run_one.py i 16000000 1 64 > /dev/null 2>&1
