#!/usr/bin/env python
"""bench.py -- 480x832 frames/s of the Self-Forcing 1.3B 4-step chunk-wise rollout on B200.

A *step* is one full rollout of the headline configuration (BASELINE.json configs[1]): random-init
Wan2.1-T2V-1.3B architecture, 21 latent frames (81 pixel frames) at 60x104 latents, chunks of 3 latent
frames, denoising steps [1000,750,500,250] warped with shift 5, context_noise 0  =  7 chunks x (4 denoise
+ 1 cache-refresh) = 35 KV-cached model forwards, 990.3 TFLOP.

    python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (CUDA kernels via the C ABI)
    python bench.py --impl reference [--gpus N] ...                reference arm: the CPU oracle port

One JSON line on stdout (rank 0).  Keys follow the driver contract:
  value        whole-job frames/s, inputs resident in HBM, CUDA events, max over ranks
  e2e          the same through CausalInferencePipeline.inference with HOST (pinned) noise / text
               embeddings copied in and the latents copied out inside the timed region
  roofline     dominant kernel (tcgen05 self-attention over the KV window): algorithmic FLOPs of its
               launches / their CUDA-event durations measured inside the timed steps, vs MEASURED_PEAKS.json
  cpu_baseline the oracle (CPU restatement of the reference) on this box's host cores, bounded sample
  clocks       nvidia-smi samples taken during the timed region
  vae_decode   extra, outside `value`: the rollout's latents decoded to pixels by the B200 VAE decoder (N = 1)
  gpu_eager_baseline  extra: the oracle run in PyTorch eager mode on the same GPU (cuBLAS + SDPA), the "beat this"
               number of SURVEY.md section 8d (N = 1); carries parity_vs_product (latents rel-L2 of the product's rollout)
  ulysses      extra (N > 1): the same video made by head-parallel groups of 4 / 2 GPUs (strong scaling), with the
               algorithmic all-to-all volume and the NVLink Tx counter of rank 0
N > 1 is data parallel over prompts (one rollout per rank per step, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ------------------------------------------------------------------------------------------------
# workload constants
# ------------------------------------------------------------------------------------------------
NL, C, FFN, NH, HD, T_CTX = 30, 1536, 8960, 12, 128, 512
LAT_FRAMES, PIX_FRAMES, LAT_H, LAT_W = 21, 81, 60, 104
FRAME_TOKENS = (LAT_H // 2) * (LAT_W // 2)          # 1560
DENOISE_STEPS = [1000, 750, 500, 250]
SHIFT = 5.0
METRIC = "480p frames/sec, 1.3B 4-step chunkwise rollout"
UNIT = "frames/s"


def forward_flops(L: int, S: int, layers: int = NL) -> float:
    """Algorithmic FLOPs of one cached forward: linear + self-attn over S + cross-attn over 512."""
    return layers * (12.0 * L * C * C + 4.0 * L * C * FFN + 4.0 * L * S * C + 4.0 * L * T_CTX * C)


def rollout_flops(chunk_frames: int) -> float:
    L = chunk_frames * FRAME_TOKENS
    return sum((len(DENOISE_STEPS) + 1) * forward_flops(L, (i + 1) * L) for i in range(LAT_FRAMES // chunk_frames))


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_ncu_kernels.json")   # written by tools/ncu_summary.py from --set full captures


def _ncu_summary():
    try:
        return json.load(open(NCU_SUMMARY))
    except Exception:
        return None


def _bytes(text: str) -> float:
    val, unit = text.split()
    return float(val.replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def ncu_traffic(kernel_substr: str = "attention_fwd_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the shipping kernel from the committed ncu
    --set full capture -- only if that capture's kernel name matches the kernel launched today, else None."""
    d = _ncu_summary()
    if not d:
        return None, None
    for name, k in d.get("kernels", {}).items():
        if kernel_substr in name and "dram__bytes_read.sum" in k:
            try:
                return _bytes(k["dram__bytes_read.sum"]) + _bytes(k["dram__bytes_write.sum"]), f"{name} ({k.get('shape', '?')})"
            except Exception:
                return None, None
    return None, None


def ncu_tensor_pipe():
    """sm__pipe_tensor_cycles_active (% of active cycles) per captured kernel -- evidence quoted beside the live numbers."""
    d = _ncu_summary()
    if not d:
        return None
    key = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    out = {}
    for name, k in d.get("kernels", {}).items():
        if key in k:
            try:
                out[name] = float(k[key].split()[0])
            except Exception:
                pass
    out["source"] = "profiles/r02_ncu_kernels.json (ncu --set full --clock-control none)"
    return out


def config_dict(chunk_frames: int, world: int = 1, sp_size: int = 1, skip_refresh_tail: bool = False,
                cuda_graphs: bool = True) -> dict:
    """The `config` object of the JSON line -- identical for the product arm and the reference arm."""
    n_videos = world // sp_size
    return {"workload": workload_name(chunk_frames),
            "parallelism": f"dp{world}" if sp_size == 1 else f"dp{n_videos} x ulysses{sp_size} (head-parallel, peer-memory all-to-all)",
            "weights": "random-init 1.3B architecture",
            "frames_per_step_per_gpu": PIX_FRAMES, "forwards_per_step": (LAT_FRAMES // chunk_frames) * 5,
            "l2": "inputs larger than L2 (2.8 GB weights + 6 GB KV cache stream through the 126 MB L2 every forward)",
            "skip_refresh_tail": bool(skip_refresh_tail), "cuda_graphs": bool(cuda_graphs)}


def workload_name(chunk_frames: int) -> str:
    kind = "chunk-wise (3 latent frames/chunk)" if chunk_frames == 3 else f"{chunk_frames} latent frame(s)/block"
    return f"wan2.1-t2v-1.3b self-forcing dmd {kind}, 81 frames 480x832, 4 steps, batch 1 per GPU"


# ------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_uuid=None):
        self.rows = []
        self.proc = None
        cmd = ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"]
        if gpu_uuid:
            cmd += ["-i", gpu_uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def mark(self) -> int:
        return len(self.rows)

    def stop(self, start_row: int = 0, end_row=None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()            # the exact child we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = self.rows[start_row:end_row] or self.rows

        def num(s):
            try:
                return float(s)
            except ValueError:
                return None
        sm = sorted(v for v in (num(r[0]) for r in rows) if v is not None)
        mx = [v for v in (num(r[1]) for r in rows) if v is not None]
        pw = [v for v in (num(r[2]) for r in rows) if v is not None]
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": int(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": int(max(mx)) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU oracle sample (cpu_baseline leg and the --impl reference arm)
# ------------------------------------------------------------------------------------------------
class CpuOracleSample:
    """the reference's CPU path (oracle port) timed on chunk 0 (S = L) AND on one late-chunk
    forward at S = 32760 (the last chunk of the video, attending to the full cache), bf16, all host threads, with
    `layers` of the 30 identical transformer layers (as many as the time budget allows).  The rollout time follows
    from the two measured points: per-layer time is linear in the KV window S (F(L,S) = linear part + attention
    proportional to S), so t(S) is interpolated between them for the 35 (or 105) forwards -- no
    FLOP-based guess of the attention cost.  The oracle is the CPU restatement of the reference's own PyTorch path
    (oracle/causal_wan_oracle.py); the reference checkout itself cannot travel to the GPU box."""

    S_LATE = LAT_FRAMES * FRAME_TOKENS   # 32760

    def __init__(self, chunk_frames: int):
        import torch
        from oracle import causal_wan_oracle as O
        self.torch, self.O = torch, O
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.chunk_frames = chunk_frames
        self.L = chunk_frames * FRAME_TOKENS
        self.x = torch.randn(1, chunk_frames, 16, LAT_H, LAT_W,
                             generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
        self.pe = torch.randn(1, T_CTX, 4096, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
        self.layers = 0

    def prepare(self, layers: int) -> None:
        O, torch = self.O, self.torch
        cfg = O.OracleConfig(dim=C, ffn_dim=FFN, num_heads=NH, num_layers=layers)
        self.w = O.OracleWrapper(O.make_random_params(cfg, seed=0), cfg, SHIFT)
        steps = O.warp_denoising_steps(self.w.scheduler, DENOISE_STEPS)
        self.ts = torch.ones([1, self.chunk_frames], dtype=torch.int64) * steps[0]
        self.kv_first = O.new_kv_cache(cfg, 1, FRAME_TOKENS, torch.bfloat16, "cpu", cache_tokens=self.L)
        # the last chunk's view of the cache: S_LATE - L tokens of earlier (random) context already in place
        self.kv_late = O.new_kv_cache(cfg, 1, FRAME_TOKENS, torch.bfloat16, "cpu", cache_tokens=self.S_LATE)
        g = torch.Generator().manual_seed(5)
        for c in self.kv_late:
            c["k"].copy_(torch.randn(c["k"].shape, generator=g).to(torch.bfloat16))
            c["v"].copy_(torch.randn(c["v"].shape, generator=g).to(torch.bfloat16))
            c["global_end_index"].fill_(self.S_LATE - self.L)
            c["local_end_index"].fill_(self.S_LATE - self.L)
        self.ca = O.new_crossattn_cache(cfg, 1, torch.bfloat16, "cpu")
        self.layers = layers

    def step(self):
        """-> (seconds of the chunk-0 forward, seconds of the S = 32760 forward).  Re-denoising a chunk overwrites its
        cache slot in place (causal_model.py:226-229), so the calls are repeatable."""
        with self.torch.no_grad():
            t0 = time.perf_counter()
            self.w(self.x, self.pe, self.ts, self.kv_first, self.ca, 0)
            t1 = time.perf_counter()
            self.w(self.x, self.pe, self.ts, self.kv_late, self.ca, self.S_LATE - self.L)
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def calibrate(self, seconds_per_step: float) -> None:
        self.prepare(1)
        self.step()                      # first call also fills the cross-attention cache
        t1 = sum(self.step())
        layers = max(1, min(NL, int(seconds_per_step / max(t1, 1e-3))))
        if layers != 1:
            self.prepare(layers)
            self.step()

    def rollout_seconds(self, t_first: float, t_late: float) -> float:
        """Per-forward time interpolated linearly in S between the two measured points, summed over the rollout."""
        scale = NL / self.layers
        if self.S_LATE == self.L:
            return t_first * scale * 5
        slope = (t_late - t_first) / (self.S_LATE - self.L)
        n_chunks = LAT_FRAMES // self.chunk_frames
        return scale * sum((len(DENOISE_STEPS) + 1) * (t_first + slope * (i * self.L)) for i in range(n_chunks))

    def fps(self, sec) -> float:
        return PIX_FRAMES / self.rollout_seconds(*sec)

    def describe(self, sec) -> str:
        t_first, t_late = sec
        fl = forward_flops(self.L, self.L, self.layers) + forward_flops(self.L, self.S_LATE, self.layers)
        return (f"oracle forward of chunk 0 (L=S={self.L}) in {t_first:.2f} s + forward of the last chunk (S={self.S_LATE}) "
                f"in {t_late:.2f} s, {self.layers} of {NL} layers = {fl / 1e12:.2f} TFLOP on {self.threads} threads (bf16, SDPA); "
                f"rollout of {(LAT_FRAMES // self.chunk_frames) * 5} forwards = {self.rollout_seconds(*sec):.1f} s by linear "
                f"interpolation in S between the two measured points")


def run_reference_arm(args) -> None:
    """`--impl reference`: the reference's CPU path (oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = max(1, args.steps + args.warmup)
    per_step = max(2.0, min(20.0, 150.0 / total))
    s = CpuOracleSample(args.chunk_frames)
    s.calibrate(per_step)
    for _ in range(args.warmup):
        s.step()
    times = [s.step() for _ in range(args.steps)]
    sec = (sum(t[0] for t in times) / len(times), sum(t[1] for t in times) / len(times))
    fps = s.fps(sec)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sum(sec) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config_dict(args.chunk_frames, max(1, args.gpus)), "device": "cpu (host cores of the GPU box)",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": s.threads, "kind": "port", "sample": s.describe(sec)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
class _HostTextEncoder:
    """Stands in for the (out-of-scope) UMT5 encoder: the e2e leg keeps the prompt embeddings in pinned
    host memory and copies them to the device on every call -- that is the step's H2D input traffic."""

    def __init__(self, pe, device):
        self.pe, self.device = pe, device

    def __call__(self, text_prompts):
        return {"prompt_embeds": self.pe.to(self.device, non_blocking=True)}


class _NoVAE:
    def decode_to_pixel(self, x, use_cache=False):   # the VAE is after the path
        return x


def run_product_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from self_forcing_b200.ops import CudaOps
    from self_forcing_b200.pipeline import CausalInferencePipeline
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun for N>1", file=sys.stderr)

    ops = CudaOps()
    cf = args.chunk_frames
    gen = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B), timestep_shift=SHIFT, device=dev, init_seed=0, ops=ops)
    # --mode dp (default): data parallel over prompts, every rank its own video (inference.py:45,96-100 in the
    # reference).  --mode ulysses: ranks form head-parallel groups of `sp` GPUs that make ONE video together
    # (self_forcing_b200/ulysses.py); groups are data parallel.  `video` = index of this rank's video.
    sp_size = 1
    if args.mode == "ulysses" and world > 1:
        from self_forcing_b200.ulysses import UlyssesGroup
        sp_size = next(p for p in (4, 2, 1) if world % p == 0 and p <= world)
        groups = [dist.new_group(list(range(g0, g0 + sp_size))) for g0 in range(0, world, sp_size)]
        gen.model.enable_ulysses(UlyssesGroup(groups[rank // sp_size], device=dev))
    video = rank // sp_size
    n_videos = world // sp_size
    pe_host = torch.randn(1, T_CTX, 4096, generator=torch.Generator().manual_seed(1 + video)).to(torch.bfloat16).pin_memory()
    noise_host = torch.randn(1, LAT_FRAMES, 16, LAT_H, LAT_W,
                             generator=torch.Generator().manual_seed(2 + video)).to(torch.bfloat16).pin_memory()
    out_host = torch.empty_like(noise_host).pin_memory()
    pe_dev, noise_dev = pe_host.to(dev), noise_host.to(dev)
    pargs = types.SimpleNamespace(denoising_step_list=DENOISE_STEPS, warp_denoising_step=True, num_frame_per_block=cf,
                                  independent_first_frame=False, context_noise=0, model_kwargs={},
                                  skip_refresh_tail=args.skip_refresh_tail)
    gen.model.use_cuda_graphs = not args.no_cuda_graphs
    enc_dev = lambda text_prompts: {"prompt_embeds": pe_dev}   # noqa: E731
    pipe = CausalInferencePipeline(pargs, dev, generator=gen, text_encoder=enc_dev, vae=_NoVAE())
    torch.manual_seed(1234 + video)      # the re-noise stream must be identical on the ranks of one video

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_step():
        _, lat = pipe.inference(noise_dev, ["synthetic"], return_latents=True)
        return lat

    def host_step():
        pipe.text_encoder = _HostTextEncoder(pe_host, dev)
        noise = noise_host.to(dev, non_blocking=True)
        _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
        out_host.copy_(lat, non_blocking=True)
        torch.cuda.synchronize()
        pipe.text_encoder = enc_dev

    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        pass
    clocks = ClockSampler(uuid) if rank == 0 else None

    if args.ncu_rollout:     # one plain rollout and nothing else: the command profiled under ncu
        resident_step()
        torch.cuda.synchronize()
        if clocks:
            clocks.stop()
        print(json.dumps({"ncu_rollout": "ok", "gpu_launches": ops.launches}))
        return

    for _ in range(args.warmup):
        resident_step()
    # ---- timed region 1: inputs resident in HBM, CUDA events on the launching (current) stream ----
    # (every forward after its first occurrence is replayed as one CUDA graph -- the product default)
    barrier()
    row0 = clocks.mark() if clocks else 0
    launches0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lat = resident_step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launches - launches0
    finite = bool(torch.isfinite(lat.float()).all().item())
    lat_last = lat.clone()

    # ---- timed region 1b: the same K steps with every attention launch bracketed by CUDA events (roofline leg;
    # individual launches cannot be timed inside a graph replay, so this pass launches eagerly).
    # (Event pairs around the ~30 us kernels also time the host whenever it falls behind -- the cross-attention launch read
    # 35 / 47 / 87 us on three boxes; queueing each forward behind 20 ms of cuBLAS work removes the gaps but leaves the
    # chip at the GEMM's lower power-capped clock and read every kernel 10 % slow.  So: plain eager launches; per-kernel GPU
    # times without host gaps are in the ncu launch list, profiles/r02q_ncu_launch_shares.json.)
    ops.start_profile(only={"attention", "attention_sp"})
    barrier()
    for _ in range(args.steps):
        resident_step()
    barrier()
    row1 = clocks.mark() if clocks else 0
    attn_prof = ops.stop_profile()

    # ---- timed region 2: end to end with host buffers ------------------------------------------
    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop(row0, None) if clocks else None

    # ---- per-kernel breakdown of one more (untimed) rollout --------------------------------------
    ops.start_profile()
    resident_step()
    prof = ops.stop_profile()

    def shutdown():
        """Tear the process group down; never let a stuck teardown turn a finished measurement into a hang."""
        if world == 1:
            return
        sys.stdout.flush()
        watchdog = threading.Timer(45.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        gen.model._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()

    ulysses = None
    if world > 1 and args.mode == "dp" and not args.no_ulysses_leg and world % 2 == 0:
        try:
            ulysses = ulysses_leg(args, world, rank, local, dev, ops, cf, ms_total / args.steps)
        except Exception as e:      # the headline line must survive a failure of the extra leg
            ulysses = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world > 1:
        dist.barrier()
    if rank != 0:
        shutdown()
        return

    pk = peaks()
    frames = PIX_FRAMES * n_videos * args.steps
    value = frames / (ms_total / 1e3)
    # roofline of the dominant kernel: self-attention launches (KV window > text length) inside the timed steps
    self_attn = [(tag, ms) for _, tag, ms in attn_prof if tag[3] > T_CTX]
    fl = sum(4.0 * tag[1] * tag[2] * tag[3] * tag[4] * HD for tag, _ in self_attn)
    ms_attn = sum(ms for _, ms in self_attn)
    achieved = fl / (ms_attn * 1e-3) / 1e12 if ms_attn > 0 else 0.0
    groups = {}
    for name, tag, ms in prof:
        key = name
        if name in ("attention", "attention_sp"):
            name = "attention"
            key = "attention_self" if tag[3] > T_CTX else "attention_cross"
        g = groups.setdefault(key, [0, 0.0, 0.0])
        g[0] += 1
        g[1] += ms
        if name == "attention":
            g[2] += 4.0 * tag[1] * tag[2] * tag[3] * tag[4] * HD
        elif name == "gemm":
            g[2] += 2.0 * tag[1] * tag[2] * tag[3]
    kernel_ms = sum(g[1] for g in groups.values())
    breakdown = {k: {"launches": g[0], "ms": round(g[1], 3), "share": round(g[1] / kernel_ms, 4),
                     **({"tflops": round(g[2] / g[1] / 1e9, 1)} if g[2] else {})}
                 for k, g in sorted(groups.items(), key=lambda kv: -kv[1][1])}
    gemm = groups.get("gemm", [0, 0.0, 0.0])
    # HBM-bound kernels: algorithmic bytes (rows read + written, SURVEY.md 8d / DESIGN.md 4) over the event-timed launches
    Lrows = cf * FRAME_TOKENS // sp_size
    hbm_lines = []
    # qk_norm_rope at batch 1: V goes straight from the QKV GEMM into its cache slot, the kernel reads q, k and writes
    # q, k (4 passes); the sequence-parallel form also moves V (6 passes)
    for kname, passes in (("qk_norm_rope", 4), ("qk_norm_rope_sp", 6), ("ln_modulate", 2), ("ln_affine", 2), ("rmsnorm", 2)):
        gk = groups.get(kname)
        if gk and gk[1] > 0:
            nbytes = passes * Lrows * C * 2.0 * gk[0]
            gbs = nbytes / (gk[1] * 1e-3) / 1e9
            hbm_lines.append({"kernel": kname, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                              "frac": gbs / pk["hbm"], "bytes_per_launch": passes * Lrows * C * 2.0,
                              "note": "per-launch CUDA events of the eager breakdown pass (event pairs around ~10 us kernels include host launch gaps; GPU-only times: profiles/r02q_ncu_launch_shares.json)"})
    total_fl = rollout_flops(cf)
    traffic, traffic_src = ncu_traffic()
    # per-shape view of the projections (event-timed eager launches of the breakdown pass)
    gemm_shapes = {}
    for name, tag, ms in prof:
        if name == "gemm":
            g = gemm_shapes.setdefault(f"M{tag[1]}_N{tag[2]}_K{tag[3]}_epi{tag[4] if len(tag) > 4 else '?'}", [0, 0.0, 0.0])
            g[0] += 1
            g[1] += ms
            g[2] += 2.0 * tag[1] * tag[2] * tag[3]
    gemm_shapes = {k: {"launches": g[0], "us_per_launch": round(g[1] / g[0] * 1e3, 2), "tflops": round(g[2] / g[1] / 1e9, 1)}
                   for k, g in sorted(gemm_shapes.items(), key=lambda kv: -kv[1][1]) if g[1] > 0}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak" if sp_size == 1 else "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": config_dict(cf, world, sp_size, args.skip_refresh_tail, gen.model.use_cuda_graphs),
        "per_gpu": value / world,
        "model_tflops": total_fl * n_videos * args.steps / (ms_total / 1e3) / 1e12,
        "model_frac_of_peak": total_fl * n_videos * args.steps / (ms_total / 1e3) / 1e12 / pk["sustained"] / world,
        "e2e": {"value": frames / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": noise_host.numel() * 2 + pe_host.numel() * 2, "d2h_bytes_per_step": out_host.numel() * 2},
        "gpu_launches": launches,
        "roofline": {"kernel": "attention_fwd_kernel (self-attention over the KV window)", "bound": "tensor",
                     "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["sustained"], "traffic": traffic,
                     "traffic_note": (f"dram bytes of ONE launch of {traffic_src} from {os.path.basename(NCU_SUMMARY)}; "
                                      if traffic is not None else "no committed capture of the shipping kernel; ") +
                                     "`achieved` sums all self-attention launches of the timed steps (S = 4680 .. 32760)",
                     "peak_kind": "sustained bf16 cuBLAS, " + pk["source"], "frac_of_burst": achieved / pk["burst"],
                     "launches_timed": len(self_attn), "ms_in_timed_region": ms_attn,
                     "timed_region": "second pass of the same K steps, launched eagerly with CUDA events around each attention launch"},
        "roofline_gemm": {"kernel": "gemm_bf16_kernel (all projections)", "bound": "tensor",
                          "achieved": gemm[2] / gemm[1] / 1e9 if gemm[1] else 0.0, "peak": pk["sustained"],
                          "unit": "TFLOP/s", "frac": (gemm[2] / gemm[1] / 1e9 if gemm[1] else 0.0) / pk["sustained"]},
        "roofline_hbm": hbm_lines, "ncu_tensor_pipe_pct": ncu_tensor_pipe(),
        "breakdown": breakdown, "gemm_shapes": gemm_shapes, "kernel_ms_per_step": kernel_ms, "finite": finite,
        "clocks": clk,
    }
    if ulysses is not None:
        line["ulysses"] = ulysses
    if world == 1 and not args.no_vae:
        # SURVEY.md section 8f rank 1, reported beside the headline (NOT part of `value`, whose metric excludes T5 and
        # the VAE): decode of this rollout's latents to 81 frames 480x832 through B200VAEWrapper (random-init decoder)
        try:
            line["vae_decode"] = vae_decode_leg(ops, dev, lat_last, ms_total / args.steps)
        except Exception as e:   # the headline line must survive a failure of the extra leg
            line["vae_decode"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_batch_leg:
        # BASELINE config 4 (many prompts per GPU): the same rollout with TWO videos per forward.  Weights stream from HBM
        # once for both and every GEMM has 9360 rows (N = 1536: 222 pair tiles = exactly three rounds on 74 CTA pairs).
        try:
            line["throughput_batch2"] = batch_leg(args, gen, pargs, dev, ops, value)
        except Exception as e:
            line["throughput_batch2"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_gpu_eager:
        # "additionally report the reference GPU eager path (cuBLAS + FA2 / SDPA) on the same
        # B200 as the beat-this number".  The reference checkout cannot travel to this box, so its restatement (the
        # oracle, same op sequence in plain PyTorch) runs in eager mode on this GPU with the product's weights and inputs.
        try:
            line["gpu_eager_baseline"] = gpu_eager_leg(gen, cf, pe_dev, noise_dev, value, resident_step)
            pr = line["gpu_eager_baseline"].get("parity_vs_product")
            line["parity_rel_l2"] = pr["latents_rel_l2"] if pr else None
        except Exception as e:   # an extra leg must never take the headline line down
            line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline:
        s = CpuOracleSample(cf)
        s.calibrate(args.cpu_seconds)
        sec = s.step()
        line["cpu_baseline"] = {"value": s.fps(sec), "unit": UNIT, "cores": s.threads, "kind": "port",
                                "sample": s.describe(sec)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"bench_n{world}{'' if sp_size == 1 else '_ulysses'}.json"), "w") as f:
        json.dump(line, f, indent=1)
    print(json.dumps(line), flush=True)
    shutdown()


def batch_leg(args, gen, pargs, dev, ops, fps_batch1: float, batch: int = 2):
    """Frames/s of one GPU making `batch` videos at once (own prompt / noise / cache rows per video), same pipeline call."""
    import torch
    from self_forcing_b200.pipeline import CausalInferencePipeline
    pe = torch.randn(batch, T_CTX, 4096, generator=torch.Generator().manual_seed(11)).to(torch.bfloat16).to(dev)
    noise = torch.randn(batch, LAT_FRAMES, 16, LAT_H, LAT_W, generator=torch.Generator().manual_seed(12)).to(torch.bfloat16).to(dev)
    pipe = CausalInferencePipeline(pargs, dev, generator=gen, text_encoder=lambda text_prompts: {"prompt_embeds": pe}, vae=_NoVAE())
    prompts = ["synthetic"] * batch
    for _ in range(3):                  # eager, capture, first replay
        pipe.inference(noise, prompts, return_latents=True)
    torch.cuda.synchronize()
    steps = max(1, args.steps - 1)
    l0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        _, lat = pipe.inference(noise, prompts, return_latents=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fps = batch * PIX_FRAMES / (ms / 1e3)
    gen.model._graphs.clear()           # give the graphs' private pools back before the next leg
    return {"value": fps, "unit": UNIT, "batch_per_gpu": batch, "ms_per_step": ms, "steps": steps,
            "model_tflops": batch * rollout_flops(args.chunk_frames) / (ms / 1e3) / 1e12,
            "speedup_vs_batch1": fps / fps_batch1, "gpu_launches": ops.launches - l0,
            "finite": bool(torch.isfinite(lat.float()).all().item()),
            "config": {"parallelism": f"dp1 x batch {batch}", "note": "device-timed, inputs resident, CUDA graphs; "
                       "same kernels and pipeline call as the headline, two videos per forward"}}


def nvlink_tx_kib(gpu_index: int):
    """Sum of the NVLink data Tx counters (KiB) of one GPU from `nvidia-smi nvlink -gt d`, or None."""
    try:
        r = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu_index)], capture_output=True, text=True, timeout=20)
        tot, seen = 0, False
        for ln in r.stdout.splitlines():
            if "Data Tx" in ln:
                seen = True
                tot += int(ln.split(":")[-1].strip().split()[0])
        return tot if seen else None
    except Exception:
        return None


def ulysses_leg(args, world: int, rank: int, local: int, dev, ops, cf: int, dp_ms_per_video: float):
    """Strong-scaling mode of SURVEY.md 8e, measured in the same run as the data-parallel line so the driver's SCALE record
    carries it: ranks form head-parallel groups of P = 4 (or 2) GPUs that make ONE video together (peer-memory
    all-to-all inside qk_norm_rope_sp / attention_fwd_sp); groups are data parallel.  Every rank takes part; rank 0
    returns the record."""
    import torch
    import torch.distributed as dist
    from self_forcing_b200.pipeline import CausalInferencePipeline
    from self_forcing_b200.ulysses import UlyssesGroup
    from self_forcing_b200.wrapper import WAN_T2V_1_3B, B200DiffusionWrapper
    P = next(p for p in (4, 2) if world % p == 0)
    groups = [dist.new_group(list(range(g0, g0 + P))) for g0 in range(0, world, P)]
    video = rank // P
    gen = B200DiffusionWrapper(model_config=dict(WAN_T2V_1_3B), timestep_shift=SHIFT, device=dev, init_seed=0, ops=ops)
    gen.model.enable_ulysses(UlyssesGroup(groups[video], device=dev))
    pe = torch.randn(1, T_CTX, 4096, generator=torch.Generator().manual_seed(1 + video)).to(torch.bfloat16).to(dev)
    noise = torch.randn(1, LAT_FRAMES, 16, LAT_H, LAT_W, generator=torch.Generator().manual_seed(2 + video)).to(torch.bfloat16).to(dev)
    pargs = types.SimpleNamespace(denoising_step_list=DENOISE_STEPS, warp_denoising_step=True, num_frame_per_block=cf,
                                  independent_first_frame=False, context_noise=0, model_kwargs={}, skip_refresh_tail=False)
    pipe = CausalInferencePipeline(pargs, dev, generator=gen, text_encoder=lambda text_prompts: {"prompt_embeds": pe}, vae=_NoVAE())
    torch.manual_seed(99 + video)       # identical re-noise stream on the ranks of one video
    for _ in range(3):                  # eager, capture, first replay
        pipe.inference(noise, ["synthetic"], return_latents=True)
    torch.cuda.synchronize()
    dist.barrier()
    tx0 = nvlink_tx_kib(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        _, lat = pipe.inference(noise, ["synthetic"], return_latents=True)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    tx1 = nvlink_tx_kib(local) if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_video = float(t.item()) / args.steps
    finite = bool(torch.isfinite(lat.float()).all().item())
    gen.model._graphs.clear()
    del pipe, gen
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    L = cf * FRAME_TOKENS
    forwards = (LAT_FRAMES // cf) * 5
    a2a = 4.0 * (P - 1) / P * (L / P) * C * 2          # q, k, v out + attention output back, bytes per rank per layer
    rec = {"parallelism": f"dp{world // P} x ulysses{P}", "group_size": P, "videos_in_flight": world // P,
           "ms_per_video": ms_video, "frames_per_s_per_video": PIX_FRAMES / (ms_video / 1e3),
           "frames_per_s_total": PIX_FRAMES * (world // P) / (ms_video / 1e3),
           "speedup_vs_one_gpu_in_this_run": dp_ms_per_video / ms_video, "strong_scaling_efficiency": dp_ms_per_video / ms_video / P,
           "a2a_bytes_per_rank_per_layer_algorithmic": a2a, "a2a_bytes_per_rank_per_video_algorithmic": a2a * NL * forwards,
           "finite": finite, "steps": args.steps,
           "note": "the two all-to-alls per block are st.global on peer-mapped pointers inside the producing kernels, separated by "
                   "flag barriers in peer memory; NCCL only all-gathers the [L/P, 64] head output once per forward"}
    if tx0 is not None and tx1 is not None:
        rec["nvlink_tx_bytes_per_video_rank0"] = (tx1 - tx0) * 1024.0 / args.steps
        rec["nvlink_tx_vs_algorithmic"] = rec["nvlink_tx_bytes_per_video_rank0"] / (a2a * NL * forwards)
    return rec


def gpu_eager_leg(gen, chunk_frames: int, pe_dev, noise_dev, product_fps: float, product_rollout=None) -> dict:
    """One full rollout of the ORACLE (the CPU restatement of the reference's PyTorch path) in eager mode on the GPU:
    cuBLAS for every Linear, torch SDPA for attention, op-by-op elementwise kernels with the reference's float64 RoPE /
    sinusoid / flow->x0, `.item()` index reads per forward -- what the unmodified reference does on a GPU, minus its
    flash_attn varlen packing.  Same random-init weights (the product model's state_dict) and the same inputs.  A
    reported baseline, not part of `value`; the oracle is only ever the thing compared against."""
    import torch
    from oracle import causal_wan_oracle as O
    cfg = O.OracleConfig(dim=C, ffn_dim=FFN, num_heads=NH, num_layers=NL)
    params = {k: v for k, v in gen.model.state_dict().items() if not k.startswith("pose_proj")}
    ow = O.OracleWrapper(params, cfg, SHIFT)
    steps = O.warp_denoising_steps(ow.scheduler, DENOISE_STEPS)
    sync = torch.cuda.synchronize if noise_dev.is_cuda else (lambda: None)
    with torch.no_grad():
        O.rollout(ow, noise_dev, pe_dev, steps, chunk_frames, max_chunks=1)      # warm-up: cuBLAS / SDPA heuristics
        sync()
        torch.manual_seed(4321)       # same re-noise stream (torch.randn_like on the CUDA generator) for both sides
        t0 = time.perf_counter()      # host clock around a synchronised region: eager mode is host-driven by nature
        tr = O.rollout(ow, noise_dev, pe_dev, steps, chunk_frames)
        sync()
        ms = (time.perf_counter() - t0) * 1e3
        parity = None
        if product_rollout is not None:
            # the product's rollout on the same inputs and noise stream: the parity figure the north-star asks for
            torch.manual_seed(4321)
            lat = product_rollout()
            sync()
            d = (lat.double() - tr.latents.double())
            per_chunk = [float(d[:, i:i + chunk_frames].norm() / tr.latents[:, i:i + chunk_frames].double().norm())
                         for i in range(0, lat.shape[1], chunk_frames)]
            parity = {"latents_rel_l2": float(d.norm() / tr.latents.double().norm()), "per_chunk_max": max(per_chunk),
                      "tolerance": 1e-2, "index_exact": tr.index_trace[-1] == (lat.shape[1] * FRAME_TOKENS,) * 2}
    frames = (noise_dev.shape[1] - 1) * 4 + 1
    fps = frames / (ms / 1e3)
    finite = bool(torch.isfinite(tr.latents.float()).all())
    del tr
    if noise_dev.is_cuda:
        torch.cuda.empty_cache()
    return {"value": fps, "unit": UNIT, "ms_per_step": ms, "kind": "port", "finite": finite,
            "what": "oracle (restatement of the reference's PyTorch path) in eager mode on this GPU: cuBLAS GEMMs, torch "
                    "SDPA, op-by-op elementwise kernels, per-forward .item() syncs; same weights and inputs as the product",
            "product_speedup": product_fps / fps, "parity_vs_product": parity}


def vae_decode_leg(ops, dev, latents, rollout_ms: float) -> dict:
    import torch
    from self_forcing_b200.vae import B200VAEWrapper, decode_flops, random_decoder_weights
    wrap = B200VAEWrapper(device=dev, ops=ops)
    sd, shapes = random_decoder_weights(wrap.model)
    wrap.model.load_state_dict(sd)
    wrap.decode_to_pixel(latents[:, :2])                       # warm-up: both frame kinds
    torch.cuda.synchronize()
    before = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    px = wrap.decode_to_pixel(latents)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = decode_flops(wrap.model, shapes, latents.shape[1], latents.shape[3], latents.shape[4])
    pk = peaks()
    return {"roofline": {"kernel": "conv3_implicit_kernel (implicit-GEMM 3x3 convolutions, ~78 % of the decode time)",
                         "bound": "tensor", "achieved": fl / ms / 1e9, "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": fl / ms / 1e9 / pk["sustained"], "peak_kind": "sustained bf16 cuBLAS, " + pk["source"],
                         "note": "algorithmic FLOPs of all convolutions + the middle attention / wall time of the whole "
                                 "decode (norm kernels and launch gaps included); ncu per kernel: profiles/r01o_ncu_conv_kernels.json"},
            "ms_per_video": ms, "frames_per_s": px.shape[1] / (ms / 1e3), "tflop": fl / 1e12, "tflops": fl / ms / 1e9,
            "gpu_launches": ops.launches - before, "out_shape": list(px.shape), "finite": bool(torch.isfinite(px).all()),
            "rollout_plus_decode_frames_per_s": px.shape[1] / ((ms + rollout_ms) / 1e3),
            "note": "latents -> pixels right after the rollout (implicit-GEMM tcgen05 convolutions, channels-last); "
                    "random-init decoder weights; not included in `value`"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-vae", action="store_true", help="skip the extra VAE-decode leg")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the oracle-in-eager-mode-on-the-GPU baseline leg")
    ap.add_argument("--mode", default="dp", choices=["dp", "ulysses"],
                    help="multi-GPU mode: dp = one video per GPU (default), ulysses = one video per group of 2/4 GPUs")
    ap.add_argument("--chunk-frames", type=int, default=3, help="latent frames per block (3 = headline, 1 = frame-wise)")
    ap.add_argument("--skip-refresh-tail", action="store_true",
                    help="skip the unused tail of the clean-context refresh pass (NOT the default: changes the work)")
    ap.add_argument("--no-cuda-graphs", action="store_true", help="launch every kernel eagerly instead of replaying graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch-leg", action="store_true", help="skip the two-videos-per-GPU throughput sub-record")
    ap.add_argument("--no-ulysses-leg", action="store_true", help="N > 1: skip the head-parallel (one video per 4 / 2 GPUs) sub-record")
    ap.add_argument("--ncu-rollout", action="store_true", help="run exactly one rollout (the command captured by ncu)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-oracle sample budget (seconds)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()
