#!/bin/bash
# End-of-round validation on one B200: smoke(), the -m gpu suite, the full bench line
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/r02final_smoke.log | cut -c1-300
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 > $OUT/r02final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02final_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/r02final_bench.json 2> $OUT/r02final_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02final_bench.json').read().strip().splitlines()[-1])
print('fps',p['value'],'ms',p['ms_per_step'],'e2e',p['e2e']['value'],'parity',p.get('parity_rel_l2'),p['clocks'])
print({k:(v['launches'],v['ms']) for k,v in p['breakdown'].items() if v['ms']>0.5})
print('batch2',p.get('throughput_batch2',{}).get('value'),'eager',p.get('gpu_eager_baseline',{}).get('value'),'cpu',p.get('cpu_baseline',{}).get('value'),'vae',p.get('vae_decode',{}).get('ms_per_video'))
print('roofline',p['roofline']['frac'],p['roofline']['achieved'],p['roofline']['traffic'],'gemm',p['roofline_gemm']['frac'],'launches',p['gpu_launches'])
PY
tail -2 $OUT/r02final_bench.err | cut -c1-300
