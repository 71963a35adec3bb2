# A/B of the pre-pass running ahead (HTV_AHEAD) + SECAM profile
python tools/ab_check.py HTV_AHEAD 0 default 2>&1 | grep -v "\"parity\".*\"ok\": true"
python -m pytest tests -m gpu -x -q -k "vbi or overlay or chunk or parity or dropin or cabi or long" 2>&1 | tail -3
ncu --clock-control none --set full --import-source on -k "regex:^k_sec_fm_list$|^k_sec_out$|^k_raster_secam$|^k_sec_pass0$|^k_sec_refine$" -s 5 -c 5 -f -o gpurun_out/r02_sec4 python tools/run_one.py l 16000000 1 64 > /dev/null 2>&1
