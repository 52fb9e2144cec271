#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 > $OUT/r02u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02u_pytest_gpu.log
grep -q " passed" $OUT/r02u_pytest_gpu.log && ! grep -q " failed" $OUT/r02u_pytest_gpu.log || { grep -n "Error\|assert" $OUT/r02u_pytest_gpu.log | head; exit 1; }
timeout 200 python tools/t5_bench.py > $OUT/r02u_t5_bench.json 2> $OUT/r02u_t5_bench.err; echo "t5 rc=$?"; cat $OUT/r02u_t5_bench.json; tail -2 $OUT/r02u_t5_bench.err
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r02u_bench.json 2> $OUT/r02u_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[-1])
print('fps',p['value'],'ms',p['ms_per_step'],'e2e',p['e2e']['value'],'parity',p.get('parity_rel_l2'),p['clocks'])
print({k:(v['launches'],v['ms']) for k,v in p['breakdown'].items() if v['ms']>3})
print('batch2',p.get('throughput_batch2',{}).get('value'),'eager',p.get('gpu_eager_baseline',{}).get('value'),'cpu',p.get('cpu_baseline',{}).get('value'),'vae',p.get('vae_decode',{}).get('ms_per_video'))
print('roofline',p['roofline']['frac'],p['roofline']['achieved'],p['roofline']['traffic'],'gemm',p['roofline_gemm']['frac'])
PY
tail -2 $OUT/r02u_bench.err
