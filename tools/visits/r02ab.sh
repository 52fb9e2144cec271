#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-vae --no-gpu-eager --no-cpu-baseline --no-batch-leg > $OUT/r02ab_$tag.json 2> $OUT/r02ab_$tag.err || { echo "$tag FAILED"; tail -3 $OUT/r02ab_$tag.err; return; }
  python - "$tag" <<'PY'
import json,sys
p=json.loads(open(f'gpurun_out/r02ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=p['breakdown']
print(sys.argv[1],'fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'clk',p['clocks']['sm_mhz'],{k:v['ms'] for k,v in b.items() if v['ms']>5})
PY
}
run stream A=1
run resident SFB_NO_STREAM_LN=1
run stream2 A=1
run resident2 SFB_NO_STREAM_LN=1
