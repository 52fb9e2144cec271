#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
SFB_CHECK_TIMEOUT=40 timeout 900 python tools/gpu_report.py attn > $OUT/r02o_attn_report.log 2>&1; tail -1 $OUT/r02o_attn_report.log
grep -q "failed: \[\]" $OUT/r02o_attn_report.log || { grep -v '"ok": true' $OUT/r02o_attn_report.log | cut -c1-800; exit 1; }
cp $OUT/gpu_report.json $OUT/r02o_attn_report.json
SFB_MICROBENCH_TAG=_tiles timeout 300 python tools/gpu_microbench.py attn_ 2>&1 | cut -c1-200
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-vae --no-gpu-eager --no-cpu-baseline > $OUT/r02o_$tag.json 2> $OUT/r02o_$tag.err || { echo "$tag FAILED"; tail -3 $OUT/r02o_$tag.err; return; }
  python - "$tag" <<'PY'
import json,sys
p=json.loads(open(f'gpurun_out/r02o_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=p['breakdown']
print(sys.argv[1],'fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'clk',p['clocks']['sm_mhz'],'attn_self',b['attention_self']['ms'],'attn_cross',b['attention_cross']['ms'],'gemm',b['gemm']['ms'])
PY
}
run default A=1
