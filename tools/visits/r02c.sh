#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=60 timeout 900 python tools/gpu_report.py --pending gemm_c1 > $OUT/r02c_c1_checks.log 2>&1; echo "c1 checks rc=$?"; grep -E "SUMMARY|\"ok\": false" $OUT/r02c_c1_checks.log | cut -c1-600
SFB_PROBE_BLOCK_N=514 SFB_PROBE_TAG=_c1 timeout 300 python tools/gemm_probe.py > $OUT/r02c_probe_c1.log 2>&1; echo "probe c1 rc=$?"; cut -c1-330 $OUT/r02c_probe_c1.log
SFB_PROBE_BLOCK_N=514 SFB_GEMM_TIMING=1 timeout 300 python tools/gemm_probe.py qkv o_proj cross_q ffn1 ffn2 > $OUT/r02c_timeline_c1.log 2>&1; echo "timeline c1 rc=$?"; grep -v frame_ $OUT/r02c_timeline_c1.log | cut -c1-260
SFB_CHECK_TIMEOUT=60 timeout 900 python tools/gpu_report.py --pending gemm_c2 > $OUT/r02c_c2_checks.log 2>&1; echo "c2 checks rc=$?"; grep -E "SUMMARY|\"ok\": false" $OUT/r02c_c2_checks.log | cut -c1-600
SFB_PROBE_BLOCK_N=515 SFB_PROBE_TAG=_c2 timeout 300 python tools/gemm_probe.py > $OUT/r02c_probe_c2.log 2>&1; echo "probe c2 rc=$?"; cut -c1-330 $OUT/r02c_probe_c2.log
SFB_PROBE_BLOCK_N=515 SFB_GEMM_TIMING=1 timeout 300 python tools/gemm_probe.py qkv o_proj ffn1 ffn2 > $OUT/r02c_timeline_c2.log 2>&1; echo "timeline c2 rc=$?"; grep -v frame_ $OUT/r02c_timeline_c2.log | cut -c1-260
nvidia-smi --query-gpu=name,clocks.sm --format=csv
