#!/bin/bash
# 2-GPU visit: Ulysses parity test + the N=2 bench line (data parallel + Ulysses sub-record)
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_ulysses_gpu.py -x -q --timeout 300 > $OUT/r02t_pytest_ulysses.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02t_pytest_ulysses.log
grep -q " passed" $OUT/r02t_pytest_ulysses.log && ! grep -q " failed" $OUT/r02t_pytest_ulysses.log || { tail -30 $OUT/r02t_pytest_ulysses.log; exit 1; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/r02t_bench_n2.json 2> $OUT/r02t_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02t_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'e2e',p['e2e']['value'],'clk',p['clocks'])
print('ulysses',json.dumps(p.get('ulysses'))[:900])
PY
tail -3 $OUT/r02t_bench_n2.err
