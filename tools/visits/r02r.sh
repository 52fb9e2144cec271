#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 600 python tools/gpu_report.py attn > $OUT/r02r_attn_report.log 2>&1; tail -1 $OUT/r02r_attn_report.log
grep -q "failed: \[\]" $OUT/r02r_attn_report.log || { grep -v '"ok": true' $OUT/r02r_attn_report.log | cut -c1-800; exit 1; }
cp $OUT/gpu_report.json $OUT/r02r_attn_report.json
SFB_MICROBENCH_TAG=_fp16partials timeout 200 python tools/gpu_microbench.py attn_self attn_frame 2>&1 | cut -c1-170
timeout 400 python bench.py --steps 3 --warmup 3 --no-vae --no-gpu-eager --no-cpu-baseline --no-batch-leg > $OUT/r02r_bench.json 2> $OUT/r02r_bench.err || { echo "bench FAILED"; tail -3 $OUT/r02r_bench.err; exit 1; }
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
b=p['breakdown']
print('fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'clk',p['clocks']['sm_mhz'],'attn_self',b['attention_self']['ms'],'attn_cross',b['attention_cross']['ms'],'gemm',b['gemm']['ms'],'roofline',p['roofline']['frac'])
PY
