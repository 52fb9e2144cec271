"""CausalDiffusionInferencePipeline -- the 50-step sampler on the cached causal forward: classifier-free guidance
+ flow-matching UniPC, chunk by chunk, with a positive and a negative KV cache.

Same constructor and `inference(noise, text_prompts, input_image, dwpose_data, random_ref_dwpose,
initial_latent=None, return_latents=False, start_frame_index=0)` contract as the reference's
pipeline/causal_diffusion_inference.py:13-540 for text-to-video (`:174-457`).  Image / pose conditioning
(`encode_image`, `dwpose_embedding`, CLIP) is outside the hot path and raises `NotImplementedError` (in the reference
the pose path cannot run with a KV cache anyway).

B200-first differences, none of which changes a result:
  * the conditional and unconditional forwards of a step run as ONE batch-2B forward (weights stream from HBM once,
    every GEMM has twice the rows); the positive / negative caches are the two halves of one [2B, S, H, D] cache and
    stay readable as `kv_cache_pos` / `kv_cache_neg`;
  * guidance, flow -> x0, UniC corrector and UniP predictor are one kernel per step (`sfb_cfg_unipc_step`) instead of
    ~25 elementwise launches; its scalar coefficients come from the host (unipc.py);
  * no per-step `print` of the cache indices (reference :429-430), hence no device -> host sync inside the loop.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .unipc import FlowUniPCMultistepScheduler


class CausalDiffusionInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None, image_encoder=None):
        super().__init__()
        if generator is None:
            from .wrapper import B200DiffusionWrapper
            generator = B200DiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True, device=device)
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be supplied: the UMT5 encoder and the Wan VAE are "
                             "outside the B200 hot path (SURVEY.md section 2, rows 3 and 12)")
        self.device = device
        self.generator, self.text_encoder, self.vae, self.image_encoder = generator, text_encoder, vae, image_encoder
        self.num_train_timesteps = args.num_train_timestep
        self.sampling_steps = int(getattr(args, "sampling_steps", 50))     # reference :66 hard-wires 50
        self.sample_solver = "unipc"
        self.shift = args.timestep_shift
        self.num_transformer_blocks = getattr(self.generator.model, "num_layers", 30)
        self.num_heads = getattr(self.generator.model, "num_heads", 12)
        self.head_dim = getattr(self.generator.model, "head_dim", 128)
        self.frame_seq_length = 1560
        self.kv_cache = None            # [2B, S, H, D] per layer: rows [:B] positive prompt, [B:] negative prompt
        self.crossattn_cache = None
        self.args = args
        self.torch_dtype = torch.bfloat16
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = self.generator.model.local_attn_size
        self.scalar_rounding = getattr(args, "unipc_scalar_rounding", "fp32")
        if getattr(self.generator.model, "_sp", None) is not None:
            raise NotImplementedError("the 50-step sampler runs data-parallel; Ulysses head parallelism is wired for the "
                                      "few-step pipeline only")
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block

    # ---- the reference's per-prompt cache lists, as views of the batched cache ----
    def _half(self, caches, lo: int, hi: int):
        if caches is None:
            return None
        return [{k: (v[lo:hi] if k in ("k", "v") else v) for k, v in c.items()} for c in caches]

    @property
    def kv_cache_pos(self):
        return self._half(self.kv_cache, 0, self._batch)

    @property
    def kv_cache_neg(self):
        return self._half(self.kv_cache, self._batch, 2 * self._batch)

    @property
    def crossattn_cache_pos(self):
        return self._half(self.crossattn_cache, 0, self._batch)

    @property
    def crossattn_cache_neg(self):
        return self._half(self.crossattn_cache, self._batch, 2 * self._batch)

    # ------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], input_image=None, dwpose_data=None,
                  random_ref_dwpose=None, initial_latent: Optional[torch.Tensor] = None, return_latents: bool = False,
                  start_frame_index: Optional[int] = 0):
        if input_image is not None or dwpose_data is not None or random_ref_dwpose is not None:
            raise NotImplementedError("image / pose conditioning is outside the B200 hot path (text-to-video only)")
        batch_size, num_frames, num_channels, height, width = noise.shape
        self.frame_seq_length = (height // 2) * (width // 2)
        if not self.independent_first_frame or (self.independent_first_frame and initial_latent is not None):
            assert num_frames % self.num_frame_per_block == 0
            num_blocks = num_frames // self.num_frame_per_block
        else:
            assert (num_frames - 1) % self.num_frame_per_block == 0
            num_blocks = (num_frames - 1) // self.num_frame_per_block
        num_input_frames = initial_latent.shape[1] if initial_latent is not None else 0
        num_output_frames = num_frames + num_input_frames
        cond = self.text_encoder(text_prompts=text_prompts)["prompt_embeds"]
        uncond = self.text_encoder(text_prompts=[self.args.negative_prompt] * len(text_prompts))["prompt_embeds"]
        both_dict = {"prompt_embeds": torch.cat([cond, uncond], dim=0)}
        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)

        # Step 1: caches (allocated once; reset by rebinding the index tensors like the reference :228-253)
        self._batch = batch_size
        if (self.kv_cache is None or self.kv_cache[0]["k"].shape[0] != 2 * batch_size
                or self.kv_cache[0]["k"].device != noise.device):
            self._initialize_kv_cache(batch_size, noise.dtype, noise.device)
            self._initialize_crossattn_cache(batch_size, noise.dtype, noise.device)
        else:
            for c in self.crossattn_cache:
                c["is_init"] = False
            for c in self.kv_cache:
                c["global_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)
                c["local_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)

        # Step 2: cache the conditioning frames at timestep 0 (reference :256-312)
        current_start_frame = start_frame_index
        cache_start_frame = 0
        if initial_latent is not None:
            if self.independent_first_frame:
                assert (num_input_frames - 1) % self.num_frame_per_block == 0
                num_input_blocks = (num_input_frames - 1) // self.num_frame_per_block
                output[:, :1] = initial_latent[:, :1]
                self._forward_both(initial_latent[:, :1], both_dict, 0.0, current_start_frame)
                current_start_frame += 1
                cache_start_frame += 1
            else:
                assert num_input_frames % self.num_frame_per_block == 0
                num_input_blocks = num_input_frames // self.num_frame_per_block
            for _ in range(num_input_blocks):
                ref = initial_latent[:, cache_start_frame:cache_start_frame + self.num_frame_per_block]
                output[:, cache_start_frame:cache_start_frame + self.num_frame_per_block] = ref
                self._forward_both(ref, both_dict, 0.0, current_start_frame)
                current_start_frame += self.num_frame_per_block
                cache_start_frame += self.num_frame_per_block

        # Step 3: temporal loop over chunks, 50-step guided denoising inside (reference :371-451)
        all_num_frames = [self.num_frame_per_block] * num_blocks
        if self.independent_first_frame and initial_latent is None:
            all_num_frames = [1] + all_num_frames
        for current_num_frames in all_num_frames:
            lo = cache_start_frame - num_input_frames
            latents = noise[:, lo:lo + current_num_frames]
            sample_scheduler = self._initialize_sample_scheduler(noise)
            sample_scheduler.set_begin_index(0)        # skips the timestep lookup (a device -> host read)
            for t in self._timesteps_host:
                flow = self._forward_both(latents, both_dict, float(t), current_start_frame)
                latents = sample_scheduler.step(flow[:batch_size], t, latents, return_dict=False,
                                                model_output_uncond=flow[batch_size:],
                                                guidance_scale=self.args.guidance_scale)[0]
            output[:, cache_start_frame:cache_start_frame + current_num_frames] = latents
            # clean-context refresh of both caches at timestep 0 (reference :431-448)
            self._forward_both(latents, both_dict, 0.0, current_start_frame)
            current_start_frame += current_num_frames
            cache_start_frame += current_num_frames

        video = self.vae.decode_to_pixel(output)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        return (video, output) if return_latents else video

    def _forward_both(self, latents: torch.Tensor, both_dict, t: float, start_frame: int) -> torch.Tensor:
        """Conditional and unconditional forward of the same latents as one batch; returns flow [2B, F, C, H, W]."""
        B, F = latents.shape[:2]
        timestep = torch.full([2 * B, F], t, device=latents.device, dtype=torch.float32)
        flow, _ = self.generator(noisy_image_or_video=torch.cat([latents, latents], dim=0), conditional_dict=both_dict,
                                 timestep=timestep, kv_cache=self.kv_cache, crossattn_cache=self.crossattn_cache,
                                 current_start=start_frame * self.frame_seq_length)
        return flow

    def _initialize_sample_scheduler(self, noise):
        """reference :519-540 (UniPC only; `sample_solver` is hard-wired there, :67)."""
        if self.sample_solver != "unipc":
            raise NotImplementedError("Unsupported solver.")
        ops = getattr(self.generator.model, "ops", None)
        s = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1, use_dynamic_shifting=False,
                                        ops=ops, scalar_rounding=self.scalar_rounding)
        s.set_timesteps(self.sampling_steps, device=noise.device, shift=self.shift)
        self.timesteps = s.timesteps
        self._timesteps_host = list(s._timesteps_host)
        return s

    # ------------------------------------------------------------------------------------
    def _initialize_kv_cache(self, batch_size, dtype, device):
        """reference :466-497, both prompts in one tensor: rows [:B] positive, [B:] negative."""
        if self.local_attn_size != -1:
            kv_cache_size = self.local_attn_size * self.frame_seq_length
        else:
            kv_cache_size = 32760
        self.kv_cache = [{
            "k": torch.zeros([2 * batch_size, kv_cache_size, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "v": torch.zeros([2 * batch_size, kv_cache_size, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "global_end_index": torch.tensor([0], dtype=torch.long, device=device),
            "local_end_index": torch.tensor([0], dtype=torch.long, device=device),
        } for _ in range(self.num_transformer_blocks)]

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        """reference :499-517."""
        self.crossattn_cache = [{
            "k": torch.zeros([2 * batch_size, 512, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "v": torch.zeros([2 * batch_size, 512, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "is_init": False,
        } for _ in range(self.num_transformer_blocks)]
