#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 400 python tools/gpu_report.py qk_norm_rope gemm_stats gemm_seg > $OUT/r02p_report.log 2>&1; tail -1 $OUT/r02p_report.log
grep -q "failed: \[\]" $OUT/r02p_report.log || { grep -v '"ok": true' $OUT/r02p_report.log | cut -c1-1000; exit 1; }
timeout 120 python tools/gpu_microbench.py elementwise 2>&1 | cut -c1-200
timeout 600 python -m pytest tests -x -q -m gpu --timeout 120 > $OUT/r02p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02p_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r02p_bench.json 2> $OUT/r02p_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print('fps',p['value'],'ms',p['ms_per_step'],'e2e',p['e2e']['value'],'parity',p.get('parity_rel_l2'),p['clocks'])
print({k:(v['launches'],v['ms']) for k,v in p['breakdown'].items()})
print('batch2',p.get('throughput_batch2'))
print('roofline',p['roofline']['frac'],p['roofline']['achieved'],'gemm',p['roofline_gemm']['frac'])
PY
tail -3 $OUT/r02p_bench.err
