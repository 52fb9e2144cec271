#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for v in "" _p0x8080 _p0x8888 _p0xA4A4 _p0xAAAA; do
  export SFB200_LIB=$PWD/self_forcing_b200/libsfb200$v.so
  echo "=== variant '$v'"
  SFB_CHECK_TIMEOUT=40 timeout 200 python tools/gpu_report.py attn_small attn_chunk attn_split_3way attn_half_split attn_sharp attn_qnorm_cross > $OUT/r02v_report$v.log 2>&1; tail -1 $OUT/r02v_report$v.log
  grep -o '"name": "[a-z_0-9]*", "ok": [a-z]*, "metrics": {"err_rel_l2": [0-9.e-]*' $OUT/r02v_report$v.log | sed 's/"metrics": {//' | tr '\n' ';'; echo
  SFB_MICROBENCH_TAG=$v timeout 120 python tools/gpu_microbench.py attn_self attn_cross 2>&1 | cut -c1-150
done
