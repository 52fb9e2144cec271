#!/bin/bash
# 4-GPU visit: Ulysses parity at world 2 and 4, then the N=4 bench line (data parallel + Ulysses P=4 sub-record)
OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m pytest tests/test_ulysses_gpu.py -x -q --timeout 300 > $OUT/r02z_pytest_ulysses.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02z_pytest_ulysses.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 2 --warmup 3 > $OUT/r02z_bench_n4.json 2> $OUT/r02z_bench_n4.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02z_bench_n4.json').read().strip().splitlines()[-1])
print('N=4 fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'e2e',p['e2e']['value'],'clk',p['clocks'])
u=p.get('ulysses') or {}
print('ulysses',{k:u.get(k) for k in ('parallelism','ms_per_video','frames_per_s_per_video','speedup_vs_one_gpu_in_this_run','strong_scaling_efficiency','finite','error')})
PY
tail -2 $OUT/r02z_bench_n4.err | cut -c1-300
