#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 600 python tools/gpu_report.py attn > $OUT/r02x_attn_report.log 2>&1; tail -1 $OUT/r02x_attn_report.log
grep -q "failed: \[\]" $OUT/r02x_attn_report.log || { grep -v '"ok": true' $OUT/r02x_attn_report.log | cut -c1-800; exit 1; }
cp $OUT/gpu_report.json $OUT/r02x_attn_report.json
timeout 500 python bench.py --steps 3 --warmup 3 --no-vae --no-gpu-eager --no-cpu-baseline --no-batch-leg > $OUT/r02x_bench.json 2> $OUT/r02x_bench.err || { echo "bench FAILED"; tail -3 $OUT/r02x_bench.err; exit 1; }
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1])
b=p['breakdown']
print('fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'e2e',round(p['e2e']['value'],2),'clk',p['clocks']['sm_mhz'],{k:v['ms'] for k,v in b.items() if v['ms']>5},'roofline',p['roofline']['frac'],'launches',p['gpu_launches'])
PY
timeout 200 python tools/graph_gap_probe.py 2>&1 | tail -1
