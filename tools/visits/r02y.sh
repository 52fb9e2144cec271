#!/bin/bash
# 2-GPU visit: the whole -m gpu suite (incl. the Ulysses test at world 2) + the N=2 bench line
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_ulysses_gpu.py -x -q --timeout 300 > $OUT/r02y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02y_pytest_gpu.log
grep -q " passed" $OUT/r02y_pytest_gpu.log && ! grep -q " failed" $OUT/r02y_pytest_gpu.log || { grep -n "Error\|assert" $OUT/r02y_pytest_gpu.log | head; exit 1; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/r02y_bench_n2.json 2> $OUT/r02y_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02y_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'e2e',p['e2e']['value'],'clk',p['clocks'])
u=p.get('ulysses') or {}
print('ulysses',{k:u.get(k) for k in ('ms_per_video','frames_per_s_per_video','speedup_vs_one_gpu_in_this_run','strong_scaling_efficiency','finite','error')})
PY
tail -2 $OUT/r02y_bench_n2.err
