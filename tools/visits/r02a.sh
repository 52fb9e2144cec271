#!/bin/bash
# Round-2 first GPU visit: everything round 1 left "pending hardware validation".
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r02a_smi.log 2>&1
SFB_CHECK_TIMEOUT=200 timeout 900 python tools/gpu_report.py --pending > $OUT/r02a_pending.log 2>&1; echo "pending rc=$?"; tail -3 $OUT/r02a_pending.log
cp $OUT/gpu_report.json $OUT/r02a_pending_report.json
timeout 900 python tools/fullsize_parity.py --fp32-yardstick --rollout > $OUT/r02a_fullsize_parity.json 2> $OUT/r02a_fullsize_parity.err; echo "parity rc=$?"; cat $OUT/r02a_fullsize_parity.json; tail -5 $OUT/r02a_fullsize_parity.err
SFB_MICROBENCH_TAG=_r02a timeout 600 python tools/gpu_microbench.py gemm_ attn_ elementwise > $OUT/r02a_microbench.log 2>&1; echo "microbench rc=$?"; cat $OUT/r02a_microbench.log | cut -c1-400
timeout 300 python tools/t5_bench.py > $OUT/r02a_t5_bench.log 2>&1; echo "t5 rc=$?"; tail -5 $OUT/r02a_t5_bench.log
