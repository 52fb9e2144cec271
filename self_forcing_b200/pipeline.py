"""CausalInferencePipeline -- the few-step DMD chunk-wise rollout driver.

Same constructor, attributes and `inference(noise, text_prompts, initial_latent=None,
return_latents=False, profile=False, low_memory=False)` contract as the reference's
pipeline/causal_inference.py:9-312, re-implemented here because the reference file cannot travel
to the GPU box.  (The reference's own, unmodified pipeline also runs on top of
`B200DiffusionWrapper` -- that is the drop-in boundary, see INTEGRATION.md and
tests/test_dropin_reference_pipeline.py.)

Differences, none of which changes a result:
  * no per-step `print` (reference :189) and no `.item()` host syncs inside the loop;
  * `low_memory` is accepted and ignored (180 GB of HBM: model + 6 GB/sample cache stay resident);
  * optional `skip_refresh_tail`: the clean-context refresh pass (:227-235) discards its output, so
    everything after the last layer's K/V append can be skipped.  Off by default.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class CausalInferencePipeline(torch.nn.Module):
    def __init__(self, args, device, generator=None, text_encoder=None, vae=None):
        super().__init__()
        if generator is None:
            from .wrapper import B200DiffusionWrapper
            generator = B200DiffusionWrapper(**getattr(args, "model_kwargs", {}), is_causal=True, device=device)
        if text_encoder is None or vae is None:
            raise ValueError("text_encoder and vae must be supplied: the UMT5 encoder and the Wan VAE are "
                             "outside the B200 hot path (SURVEY.md section 2, rows 3 and 12)")
        self.generator, self.text_encoder, self.vae = generator, text_encoder, vae
        self.scheduler = self.generator.get_scheduler()
        self.denoising_step_list = torch.tensor(args.denoising_step_list, dtype=torch.long)
        if args.warp_denoising_step:   # reference :29-31
            timesteps = torch.cat((self.scheduler.timesteps.cpu(), torch.tensor([0], dtype=torch.float32)))
            self.denoising_step_list = timesteps[1000 - self.denoising_step_list]
        self.num_transformer_blocks = getattr(self.generator.model, "num_layers", 30)
        self.num_heads = getattr(self.generator.model, "num_heads", 12)
        self.head_dim = getattr(self.generator.model, "head_dim", 128)
        self.frame_seq_length = 1560
        self.kv_cache1 = None
        self.crossattn_cache = None
        self.args = args
        self.num_frame_per_block = getattr(args, "num_frame_per_block", 1)
        self.independent_first_frame = args.independent_first_frame
        self.local_attn_size = self.generator.model.local_attn_size
        self.skip_refresh_tail = bool(getattr(args, "skip_refresh_tail", False))
        if self.num_frame_per_block > 1:
            self.generator.model.num_frame_per_block = self.num_frame_per_block

    # ------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, text_prompts: List[str], initial_latent: Optional[torch.Tensor] = None,
                  return_latents: bool = False, profile: bool = False, low_memory: bool = False):
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
        conditional_dict = self.text_encoder(text_prompts=text_prompts)
        output = torch.zeros([batch_size, num_output_frames, num_channels, height, width], device=noise.device,
                             dtype=noise.dtype)
        events = None
        if profile:
            events = dict(init=(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)),
                          diffusion=(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)),
                          vae=(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)), blocks=[])
            events["init"][0].record()

        # Step 1: KV / cross-attention caches (allocated once, reset by rebinding like the reference :112-132)
        want_tokens = self.local_attn_size * self.frame_seq_length if self.local_attn_size != -1 else 32760
        kc0 = None if self.kv_cache1 is None else self.kv_cache1[0]["k"]
        if (kc0 is None or kc0.shape[0] != batch_size or kc0.device != noise.device or kc0.dtype != noise.dtype
                or kc0.shape[1] != want_tokens):      # another batch / device / dtype / resolution: new caches
            self._initialize_kv_cache(batch_size, noise.dtype, noise.device)
            self._initialize_crossattn_cache(batch_size, noise.dtype, noise.device)
        else:
            for c in self.crossattn_cache:
                c["is_init"] = False
            for c in self.kv_cache1:
                c["global_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)
                c["local_end_index"] = torch.tensor([0], dtype=torch.long, device=noise.device)

        # Step 2: cache the conditioning frames, if any (reference :135-169)
        current_start_frame = 0
        if initial_latent is not None:
            zero_t = torch.zeros([batch_size, 1], device=noise.device, dtype=torch.int64)
            if self.independent_first_frame:
                assert (num_input_frames - 1) % self.num_frame_per_block == 0
                num_input_blocks = (num_input_frames - 1) // self.num_frame_per_block
                output[:, :1] = initial_latent[:, :1]
                self._generate(initial_latent[:, :1], conditional_dict, zero_t, current_start_frame, refresh=True)
                current_start_frame += 1
            else:
                assert num_input_frames % self.num_frame_per_block == 0
                num_input_blocks = num_input_frames // self.num_frame_per_block
            for _ in range(num_input_blocks):
                ref = initial_latent[:, current_start_frame:current_start_frame + self.num_frame_per_block]
                output[:, current_start_frame:current_start_frame + self.num_frame_per_block] = ref
                self._generate(ref, conditional_dict, zero_t.expand(batch_size, ref.shape[1]).contiguous(),
                               current_start_frame, refresh=True)
                current_start_frame += self.num_frame_per_block
        if profile:
            events["init"][1].record()
            events["diffusion"][0].record()

        # Step 3: temporal loop over chunks, spatial (denoising) loop inside (reference :176-246)
        all_num_frames = [self.num_frame_per_block] * num_blocks
        if self.independent_first_frame and initial_latent is None:
            all_num_frames = [1] + all_num_frames
        steps = self.denoising_step_list
        for current_num_frames in all_num_frames:
            if profile:
                blk = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                blk[0].record()
            lo = current_start_frame - num_input_frames
            noisy_input = noise[:, lo:lo + current_num_frames]
            for index, current_timestep in enumerate(steps):
                timestep = torch.ones([batch_size, current_num_frames], device=noise.device,
                                      dtype=torch.int64) * current_timestep
                _, denoised_pred = self._generate(noisy_input, conditional_dict, timestep, current_start_frame)
                if index < len(steps) - 1:
                    next_timestep = steps[index + 1]
                    flat = denoised_pred.flatten(0, 1)
                    noisy_input = self.scheduler.add_noise(
                        flat, torch.randn_like(flat),
                        next_timestep * torch.ones([batch_size * current_num_frames], device=noise.device,
                                                   dtype=torch.long)).unflatten(0, denoised_pred.shape[:2])
            output[:, current_start_frame:current_start_frame + current_num_frames] = denoised_pred
            # clean-context cache refresh at t = context_noise (reference :227-235)
            context_timestep = torch.ones_like(timestep) * self.args.context_noise
            self._generate(denoised_pred, conditional_dict, context_timestep, current_start_frame, refresh=True)
            if profile:
                blk[1].record()
                events["blocks"].append(blk)
            current_start_frame += current_num_frames

        if profile:
            events["diffusion"][1].record()
            events["vae"][0].record()
        video = self.vae.decode_to_pixel(output, use_cache=False)
        video = (video * 0.5 + 0.5).clamp(0, 1)
        if profile:
            events["vae"][1].record()
            torch.cuda.synchronize()
            self.last_profile = dict(
                init_ms=events["init"][0].elapsed_time(events["init"][1]),
                diffusion_ms=events["diffusion"][0].elapsed_time(events["diffusion"][1]),
                vae_ms=events["vae"][0].elapsed_time(events["vae"][1]),
                block_ms=[a.elapsed_time(b) for a, b in events["blocks"]])
        return (video, output) if return_latents else video

    def _generate(self, latents, conditional_dict, timestep, start_frame: int, refresh: bool = False):
        kwargs = dict(noisy_image_or_video=latents, conditional_dict=conditional_dict, timestep=timestep,
                      kv_cache=self.kv_cache1, crossattn_cache=self.crossattn_cache,
                      current_start=start_frame * self.frame_seq_length)
        if refresh and self.skip_refresh_tail:
            kwargs["refresh_only"] = True
        return self.generator(**kwargs)

    # ------------------------------------------------------------------------------------
    def _initialize_kv_cache(self, batch_size, dtype, device):
        """Per-layer rolling cache [B, S, H, D] (reference :278-298): S = local_attn_size * frame tokens,
        or 32760 (21 latent frames x 1560 tokens) for global attention."""
        if self.local_attn_size != -1:
            kv_cache_size = self.local_attn_size * self.frame_seq_length
        else:
            kv_cache_size = 32760
        model = self.generator.model
        if getattr(model, "_sp", None) is not None:
            # Ulysses head-parallel group: head-sharded caches in peer-mapped memory (ulysses.py)
            self.kv_cache1 = model.allocate_kv_cache(batch_size, kv_cache_size, dtype, device)
            return
        self.kv_cache1 = [{
            "k": torch.zeros([batch_size, kv_cache_size, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, kv_cache_size, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "global_end_index": torch.tensor([0], dtype=torch.long, device=device),
            "local_end_index": torch.tensor([0], dtype=torch.long, device=device),
        } for _ in range(self.num_transformer_blocks)]

    def _initialize_crossattn_cache(self, batch_size, dtype, device):
        """Per-layer text K/V [B, 512, H, D] (reference :300-312)."""
        self.crossattn_cache = [{
            "k": torch.zeros([batch_size, 512, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "v": torch.zeros([batch_size, 512, self.num_heads, self.head_dim], dtype=dtype, device=device),
            "is_init": False,
        } for _ in range(self.num_transformer_blocks)]
