#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list and full captures of the two tensor-core kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.log 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/${TAG}_bench.log; tail -5 $OUT/${TAG}_bench.err
if [ "${SKIP_WIDENED:-0}" != "1" ]; then
timeout 120 python tools/vae_bench.py --runs 2 --breakdown > $OUT/${TAG}_vae_decode_bench.json 2> $OUT/${TAG}_vae_decode_bench.err; echo "vae bench rc=$?"
timeout 200 python tools/diffusion_bench.py > $OUT/${TAG}_diffusion_sampler_bench.json 2> $OUT/${TAG}_diffusion_sampler_bench.err; echo "sampler bench rc=$?"
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 python bench.py --ncu-rollout > $OUT/${TAG}_plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --ncu-rollout > $OUT/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 300 -c 2 -f -o $OUT/${TAG}_attn \
    python bench.py --ncu-rollout > $OUT/${TAG}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 400 -c 5 -f -o $OUT/${TAG}_gemm \
    python bench.py --ncu-rollout > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
fi
ls -la $OUT
