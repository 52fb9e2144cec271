#!/bin/bash
# same-box A/B of the round's schedule changes: frames/s of bench.py (short form)
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-vae --no-gpu-eager --no-cpu-baseline > $OUT/r02m_$tag.json 2> $OUT/r02m_$tag.err || { echo "$tag FAILED"; tail -3 $OUT/r02m_$tag.err; return; }
  python - "$tag" <<'PY'
import json,sys
p=json.loads(open(f'gpurun_out/r02m_{sys.argv[1]}.json').read().strip().splitlines()[-1])
b=p['breakdown']
print(sys.argv[1],'fps',round(p['value'],2),'ms',round(p['ms_per_step'],1),'clk',p['clocks']['sm_mhz'],'attn_self',b['attention_self']['ms'],'attn_cross',b['attention_cross']['ms'],'gemm',b['gemm']['ms'])
PY
}
run default A=1
run nofold SFB_NO_FOLD=1
run nohalf SFB_ATTN_NOHALF=1
run minsplit32 SFB_ATTN_MIN_SPLIT=32
run default2 A=1
