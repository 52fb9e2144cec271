#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 -s > $OUT/r02q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep "batch-2 vs batch-1" $OUT/r02q_pytest_gpu.log | cut -c1-300; tail -3 $OUT/r02q_pytest_gpu.log
grep -q " passed" $OUT/r02q_pytest_gpu.log && ! grep -q " failed" $OUT/r02q_pytest_gpu.log || exit 1
# launch list of one rollout (eager launches): chunk 5-6 forwards, ~2 forwards' worth of launches
timeout 300 python bench.py --ncu-rollout --no-cuda-graphs > $OUT/r02q_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11000 -c 1300 --csv --log-file $OUT/r02q_launches.csv \
    python bench.py --ncu-rollout --no-cuda-graphs > $OUT/r02q_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
bash tools/ncu_kernels.sh r02 2>&1 | tail -16
