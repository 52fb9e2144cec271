#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02h_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r02h_bench.json 2> $OUT/r02h_bench.err; echo "bench rc=$?"; cat $OUT/r02h_bench.json; tail -3 $OUT/r02h_bench.err
