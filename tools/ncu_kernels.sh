#!/bin/bash
# ncu --set full captures of the SHIPPING kernels at the rollout's shapes, one launch each (short commands; one GPU).
# Each capture runs only after the same command exited 0 without ncu.  Reports land in gpurun_out/<tag>_<name>.ncu-rep;
# tools/ncu_summary.py turns them into profiles/<tag>_ncu_kernels.json (read by bench.py for roofline.traffic).
TAG=${1:-r02}
ONLY=${2:-.}      # regex on the capture names (default: all)
OUT=gpurun_out
mkdir -p $OUT
cap() {   # name, kernel regex, launches to skip, microbench selector...
  local name=$1 regex=$2 skip=$3; shift 3
  python tools/gpu_microbench.py "$@" > $OUT/${TAG}_plain_${name}.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f \
      -o $OUT/${TAG}_${name} python tools/gpu_microbench.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu ${name} rc=$?"
}
# one plain run of everything that is captured below (must exit 0 before any capture)
python tools/gpu_microbench.py attn_ gemm_qkv gemm_o_proj gemm_cross_q gemm_ffn elementwise > $OUT/${TAG}_plain_all.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain_all.log; exit 1; }
capq() {   # like cap, without the per-capture plain run (covered by the run above)
  local name=$1 regex=$2 skip=$3; shift 3
  [[ $name =~ $ONLY ]] || return 0
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f \
      -o $OUT/${TAG}_${name} python tools/gpu_microbench.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu ${name} rc=$?"
}
capq attn_S4680   attention_fwd_kernel 3 attn_self_S4680
capq attn_S18720  attention_fwd_kernel 3 attn_self_S18720
capq attn_S32760  attention_fwd_kernel 3 attn_self_S32760
capq attn_cross   attention_fwd_kernel 3 attn_cross_S512
capq gemm_qkv     gemm2_bf16_kernel 3 gemm_qkv
capq gemm_o_proj_stats  gemm2_bf16_kernel 3 gemm_o_proj_stats
capq gemm_cross_q_fold  gemm2_bf16_kernel 3 gemm_cross_q_fold
capq gemm_ffn1    gemm2_bf16_kernel 3 gemm_ffn1
capq gemm_ffn2    gemm2_bf16_kernel 3 gemm_ffn2
capq ln_modulate  "ln_kernel" 3 elementwise
capq qk_rope_stream "qk_rope_stream_kernel" 3 elementwise
ls -la $OUT | grep ${TAG}_ | grep ncu-rep
