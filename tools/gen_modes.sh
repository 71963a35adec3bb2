#!/bin/sh
# Regenerates hacktv_b200/csrc/htv_modes.c from the reference's mode table.
# Needs /root/reference and oracle/_ref objects (make -C oracle ref). Build container only.
set -e
cd "$(dirname "$0")/.."
R=oracle/_ref/obj
OBJS=$(ls $R/*.o | grep -v -e hacktv.o -e ref_harness.o -e ref_shim_raw.o -e ref_shim_zero.o -e video_b200.o -e video_cpu.o -e hacktv_main.o)
gcc -O1 -w -I/root/reference/src -o /tmp/gen_modes tools/gen_modes.c $OBJS $R/ref_shim_raw.o -lm -pthread
/tmp/gen_modes > hacktv_b200/csrc/htv_modes.c
echo "wrote hacktv_b200/csrc/htv_modes.c"
