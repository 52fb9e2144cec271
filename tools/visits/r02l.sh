#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python tools/gpu_microbench.py gemm_cross_q gemm_o_proj 2>&1 | cut -c1-220
SFB_GEMM_TIMING=1 timeout 120 python tools/gpu_microbench.py gemm_cross_q_fold gemm_cross_q_stats 2>&1 | grep -A3 "gemm2 timing" | head -24
