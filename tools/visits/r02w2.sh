#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in "" _p0x8420842084208420 _p0x8888888888880808 _p0x8888888888888080 _p0xA888A888A8888888; do
  export SFB200_LIB=$PWD/self_forcing_b200/libsfb200$v.so
  echo "=== variant '$v'"
  SFB_MICROBENCH_TAG=$v timeout 120 python tools/gpu_microbench.py attn_self_S18720 attn_self_S32760 2>&1 | cut -c1-140
done
