#!/bin/bash
# every stage bails out early on failure; short timeouts (a hung kernel must not eat the GPU budget)
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 500 python tools/gpu_report.py attn_small attn_cross attn_qnorm > $OUT/r02k_report0.log 2>&1; tail -1 $OUT/r02k_report0.log
grep -q "failed: \[\]" $OUT/r02k_report0.log || { grep -v '"ok": true' $OUT/r02k_report0.log | cut -c1-600; exit 1; }
SFB_CHECK_TIMEOUT=40 timeout 600 python tools/gpu_report.py gemm_stats gemm_lnfold gemm_pair gemm_gate attn_half > $OUT/r02k_report.log 2>&1; grep -v '"ok": true' $OUT/r02k_report.log | cut -c1-1200
grep -q "failed: \[\]" $OUT/r02k_report.log || exit 1
cp $OUT/gpu_report.json $OUT/r02k_report.json
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 > $OUT/r02k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r02k_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-vae --no-cpu-baseline > $OUT/r02k_bench.json 2> $OUT/r02k_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
print('fps',p['value'],'ms',p['ms_per_step'],'e2e',p['e2e']['value'],'parity',p.get('parity_rel_l2'),p['clocks'])
print({k:(v['launches'],v['ms']) for k,v in p['breakdown'].items()})
print({k:v['us_per_launch'] for k,v in p['gemm_shapes'].items()})
PY
tail -3 $OUT/r02k_bench.err
