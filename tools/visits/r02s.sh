#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 400 python tools/gpu_report.py attn_split attn_few attn_very attn_half attn_qnorm attn_sp > $OUT/r02s_attn_report.log 2>&1; tail -1 $OUT/r02s_attn_report.log
grep -q "failed: \[\]" $OUT/r02s_attn_report.log || { grep -v '"ok": true' $OUT/r02s_attn_report.log | cut -c1-800; exit 1; }
SFB_MICROBENCH_TAG=_combine4 timeout 200 python tools/gpu_microbench.py attn_self attn_frame attn_cross 2>&1 | cut -c1-170
timeout 500 python bench.py --steps 3 --warmup 3 --no-vae --no-gpu-eager --no-cpu-baseline --no-batch-leg > $OUT/r02s_bench.json 2> $OUT/r02s_bench.err || { echo "bench FAILED"; tail -3 $OUT/r02s_bench.err; exit 1; }
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
b=p['breakdown']
print('fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'clk',p['clocks']['sm_mhz'],{k:v['ms'] for k,v in b.items() if v['ms']>5},'roofline',p['roofline']['frac'],p['roofline']['achieved'],'gemm',p['roofline_gemm']['frac'])
print(p['roofline_hbm'])
PY
