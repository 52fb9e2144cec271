#!/bin/bash
# refresh the widened rows' numbers with the final kernels: 50-step guided sampler, VAE decode, 801-frame rolling-window video
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/diffusion_bench.py > $OUT/r02_diffusion_sampler_bench.json 2> $OUT/r02_diffusion_sampler_bench.err; echo "sampler rc=$?"; tail -c 600 $OUT/r02_diffusion_sampler_bench.json
timeout 100 python tools/vae_bench.py --runs 2 > $OUT/r02_vae_decode_bench.json 2> $OUT/r02_vae_decode_bench.err; echo "vae rc=$?"; tail -c 400 $OUT/r02_vae_decode_bench.json
timeout 150 python tools/long_video_bench.py --runs 2 > $OUT/r02_long_video_bench.json 2> $OUT/r02_long_video_bench.err; echo "long video rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_long_video_bench.json'))
print([ (r['frames_per_s'], r['steady_chunk_ms'], r['graphs']) for r in d['runs']])
PY
