#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/gemm_probe.py > $OUT/r02b_gemm_probe.log 2>&1; echo "probe rc=$?"; cat $OUT/r02b_gemm_probe.log | cut -c1-420
SFB_GEMM_TIMING=1 timeout 300 python tools/gemm_probe.py qkv o_proj cross_q cross_o ffn1 ffn2 > $OUT/r02b_gemm_timeline.log 2>&1; echo "timeline rc=$?"; grep -v "^$" $OUT/r02b_gemm_timeline.log | grep -v frame | cut -c1-300
SFB_CONV_EXACT_N=1 timeout 200 python tools/vae_bench.py --runs 2 --breakdown > $OUT/r02b_vae_exact_n.json 2> $OUT/r02b_vae_exact_n.err; echo "vae exact rc=$?"; head -c 1500 $OUT/r02b_vae_exact_n.json
timeout 200 python tools/vae_bench.py --runs 2 > $OUT/r02b_vae_default.json 2> $OUT/r02b_vae_default.err; echo "vae default rc=$?"; head -c 800 $OUT/r02b_vae_default.json
timeout 1500 python -m pytest tests/test_rollout_gpu.py -x -q -m gpu > $OUT/r02b_pytest_rollout.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r02b_pytest_rollout.log
