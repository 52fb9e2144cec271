#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02e_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r02e_bench.log 2> $OUT/r02e_bench.err; echo "bench rc=$?"; tail -c 6000 $OUT/r02e_bench.log; tail -3 $OUT/r02e_bench.err
SFB_ATTN_TIMING=1 timeout 200 python tools/gpu_microbench.py attn_cross attn_self_S4680 attn_frame > $OUT/r02e_attn_timing.log 2>&1; echo "attn timing rc=$?"; grep -E "attn timing" $OUT/r02e_attn_timing.log | sort | uniq -c | sort -rn | head -12
