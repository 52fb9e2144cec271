#!/bin/bash
# half-item attention: parity of every attention check, then A/B against the padded-pair schedule
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=60 timeout 900 python tools/gpu_report.py attn > $OUT/r02i_attn_report.log 2>&1; tail -3 $OUT/r02i_attn_report.log
cp $OUT/gpu_report.json $OUT/r02i_attn_report.json
echo "--- half items"; SFB_MICROBENCH_TAG=_half timeout 300 python tools/gpu_microbench.py attn_ 2>&1 | cut -c1-260
echo "--- padded pairs"; SFB_ATTN_NOHALF=1 SFB_MICROBENCH_TAG=_nohalf timeout 300 python tools/gpu_microbench.py attn_ 2>&1 | cut -c1-260
