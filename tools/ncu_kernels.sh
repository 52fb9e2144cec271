#!/bin/bash
# ncu --set full captures of the SHIPPING kernels at the rollout's shapes, one launch each (short commands; one GPU).
# Each capture runs only after the same command exited 0 without ncu.  Reports land in gpurun_out/<tag>_<name>.ncu-rep;
# tools/ncu_summary.py turns them into profiles/<tag>_ncu_kernels.json (read by bench.py for roofline.traffic).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
cap() {   # name, kernel regex, launches to skip, microbench selector...
  local name=$1 regex=$2 skip=$3; shift 3
  python tools/gpu_microbench.py "$@" > $OUT/${TAG}_plain_${name}.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f \
      -o $OUT/${TAG}_${name} python tools/gpu_microbench.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu ${name} rc=$?"
}
cap attn_S4680   attention_fwd_kernel 3 attn_self_S4680
cap attn_S18720  attention_fwd_kernel 3 attn_self_S18720
cap attn_S32760  attention_fwd_kernel 3 attn_self_S32760
cap attn_cross   attention_fwd_kernel 3 attn_cross_S512
cap gemm_qkv     gemm2_bf16_kernel 3 gemm_qkv
cap gemm_o_proj  gemm2_bf16_kernel 3 gemm_o_proj
cap gemm_ffn1    gemm2_bf16_kernel 3 gemm_ffn1
cap gemm_ffn2    gemm2_bf16_kernel 3 gemm_ffn2
cap ln_modulate  "ln_kernel.*Lb0" 3 elementwise
cap ln_affine    "ln_kernel.*Lb1" 3 elementwise
cap rmsnorm      "rmsnorm_kernel" 3 elementwise
cap qk_norm_rope "qk_norm_rope_kernel" 3 elementwise
ls -la $OUT | grep ${TAG}_ | grep ncu-rep
