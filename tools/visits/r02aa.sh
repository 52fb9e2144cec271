#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 300 python tools/gpu_report.py ln_modulate gemm_stats model_forward > $OUT/r02aa_report.log 2>&1; tail -1 $OUT/r02aa_report.log
grep -q "failed: \[\]" $OUT/r02aa_report.log || { grep -v '"ok": true' $OUT/r02aa_report.log | cut -c1-1000; exit 1; }
timeout 120 python tools/gpu_microbench.py elementwise 2>&1 | grep "ln_modulate" | cut -c1-200
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 > $OUT/r02aa_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02aa_pytest_gpu.log
grep -q " passed" $OUT/r02aa_pytest_gpu.log && ! grep -q " failed" $OUT/r02aa_pytest_gpu.log || { grep -n "Error\|assert" $OUT/r02aa_pytest_gpu.log | head; exit 1; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-vae --no-cpu-baseline --no-batch-leg > $OUT/r02aa_$tag.json 2> $OUT/r02aa_$tag.err || { echo "$tag FAILED"; tail -3 $OUT/r02aa_$tag.err; return; }
  python - "$tag" <<'PY'
import json,sys
p=json.loads(open(f'gpurun_out/r02aa_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=p['breakdown']
print(sys.argv[1],'fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'parity',p.get('parity_rel_l2'),'clk',p['clocks']['sm_mhz'],{k:v['ms'] for k,v in b.items() if v['ms']>5})
PY
}
run default A=1
