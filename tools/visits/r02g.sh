#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/r02g_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-vae --no-gpu-eager --no-cpu-baseline > $OUT/r02g_bench_pdl.log 2> $OUT/r02g_bench_pdl.err; echo "bench pdl rc=$?"; python -c "
import json;p=json.loads(open('$OUT/r02g_bench_pdl.log').read().strip().splitlines()[-1]);print('PDL on :',p['value'],p['ms_per_step'],p['clocks'])"; tail -2 $OUT/r02g_bench_pdl.err
SFB_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-vae --no-gpu-eager --no-cpu-baseline > $OUT/r02g_bench_nopdl.log 2> $OUT/r02g_bench_nopdl.err; echo "bench nopdl rc=$?"; python -c "
import json;p=json.loads(open('$OUT/r02g_bench_nopdl.log').read().strip().splitlines()[-1]);print('PDL off:',p['value'],p['ms_per_step'],p['clocks'])"
timeout 600 python tools/long_video_bench.py --runs 2 > $OUT/r02g_long_video.json 2> $OUT/r02g_long_video.err; echo "long video rc=$?"; cat $OUT/r02g_long_video.json; tail -3 $OUT/r02g_long_video.err
