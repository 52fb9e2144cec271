#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 500 python tools/gpu_report.py attn > $OUT/r02ad_attn_report.log 2>&1; tail -1 $OUT/r02ad_attn_report.log
grep -q "failed: \[\]" $OUT/r02ad_attn_report.log || { grep -v '"ok": true' $OUT/r02ad_attn_report.log | cut -c1-800; exit 1; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-vae --no-cpu-baseline --no-batch-leg > $OUT/r02ad_$tag.json 2> $OUT/r02ad_$tag.err || { echo "$tag FAILED"; tail -3 $OUT/r02ad_$tag.err; return; }
  python - "$tag" <<'PY'
import json,sys
p=json.loads(open(f'gpurun_out/r02ad_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=p['breakdown']
print(sys.argv[1],'fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'parity',p.get('parity_rel_l2'),'clk',p['clocks']['sm_mhz'],'attn_self',b['attention_self']['ms'],'roofline',round(p['roofline']['frac'],4))
PY
}
run tile160 A=1
run split48 SFB_ATTN_MIN_SPLIT=48
