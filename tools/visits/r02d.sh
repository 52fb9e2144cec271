#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_PROBE_BLOCK_N=513 SFB_GEMM_TIMING=1 timeout 300 python tools/gemm_probe.py ffn2 o_proj > $OUT/r02d_timeline_sk.log 2>&1; echo "timeline sk rc=$?"; grep -v frame_ $OUT/r02d_timeline_sk.log | cut -c1-260
SFB_PROBE_BLOCK_N=513 SFB_PROBE_TAG=_sk timeout 300 python tools/gemm_probe.py ffn2 o_proj qkv > $OUT/r02d_probe_sk.log 2>&1; echo "probe sk rc=$?"; cut -c1-330 $OUT/r02d_probe_sk.log
