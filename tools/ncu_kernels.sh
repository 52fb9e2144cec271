#!/bin/bash
# ncu --set full captures of the two tensor-core kernels at headline shapes (short commands; one GPU).
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/gpu_microbench.py attn_self_S18720 > $OUT/${TAG}_plain_attn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 4 -c 2 -f -o $OUT/${TAG}_attn \
    python tools/gpu_microbench.py attn_self_S18720 > $OUT/${TAG}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
python tools/gpu_microbench.py gemm_qkv gemm_ffn1 gemm_o_proj gemm_ffn2 > $OUT/${TAG}_plain_gemm.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 3 -c 20 -f -o $OUT/${TAG}_gemm \
    python tools/gpu_microbench.py gemm_qkv gemm_ffn1 gemm_o_proj gemm_ffn2 > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la $OUT | grep $TAG
