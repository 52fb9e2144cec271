#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=60 timeout 900 python tools/gpu_report.py attn > $OUT/r02j_attn_report.log 2>&1; tail -1 $OUT/r02j_attn_report.log
cp $OUT/gpu_report.json $OUT/r02j_attn_report.json
echo "--- half items, cost-weighted"; SFB_MICROBENCH_TAG=_half timeout 300 python tools/gpu_microbench.py attn_ 2>&1 | cut -c1-200
echo "--- padded pairs"; SFB_ATTN_NOHALF=1 SFB_MICROBENCH_TAG=_nohalf timeout 300 python tools/gpu_microbench.py attn_ 2>&1 | cut -c1-200
