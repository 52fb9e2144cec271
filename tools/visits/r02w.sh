#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in "" _p0x8888888888888888 _p0xAAAAAAAA88880000 _p0xAAAAAAAAAAAA0000 _p0xEEEEAAAAAAAA0000 _p0xEEEEEEEEAAAA0000; do
  export SFB200_LIB=$PWD/self_forcing_b200/libsfb200$v.so
  echo "=== variant '$v'"
  SFB_CHECK_TIMEOUT=40 timeout 100 python tools/gpu_report.py attn_chunk attn_sharp > $OUT/r02w_report$v.log 2>&1; tail -1 $OUT/r02w_report$v.log
  SFB_MICROBENCH_TAG=$v timeout 120 python tools/gpu_microbench.py attn_self_S 2>&1 | cut -c1-140
done
