/* sfb200.h -- C ABI of libsfb200.so: the B200 (sm_100a) kernels behind the Self-Forcing rollout
 * hot path.  Plain pointers, sizes and a CUDA stream handle; no torch types.
 *
 * The reference (alazarteka/Self-Forcing) is pure Python: its "FFI" for this path is the set of
 * torch / flash_attn calls listed below.  Each entry point names the reference call site(s)
 * (file:line under the reference checkout) whose device work it replaces.  INTEGRATION.md shows
 * the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every tensor is bf16 (2-byte) unless stated; "ld*" / "*_stride" are in ELEMENTS;
 *   - every pointer is a device pointer that the caller owns (no allocation inside, no state
 *     kept between calls); rows must be 16-byte aligned (strides multiples of 8 elements);
 *   - `stream` is a cudaStream_t (CUstream) cast to void*; work is enqueued asynchronously;
 *   - return value 0 = success, non-zero = error with text in sfb_last_error() (thread-local);
 *   - one process drives one GPU (torchrun style): kernel attributes are configured once per process on first use.
 */
#ifndef SFB200_H_
#define SFB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_ABI_VERSION 11

const char* sfb_last_error(void);
int sfb_abi_version(void);

/* Epilogues of sfb_gemm_bf16 (the elementwise ops the reference runs after each nn.Linear). */
#define SFB_EPI_BIAS 0      /* y = bf16(acc + bias) */
#define SFB_EPI_GELU 1      /* y = bf16(gelu_tanh(bf16(acc + bias)))        causal_model.py:278        */
#define SFB_EPI_RESIDUAL 2  /* y = bf16(res + bf16(acc + bias))             causal_model.py:324        */
#define SFB_EPI_GATE_RES 3  /* y = bf16(res + bf16(bf16(acc+bias) * gate))  causal_model.py:320,331-332 */
#define SFB_EPI_F32 4       /* y = acc (+ bias) as FLOAT: out0 is a float matrix, ldo0 in floats, one segment,
                               one-CTA tiles (attention logits of the VAE's single-head attention)             */

/* Y[M,N] = epilogue(X[M,K] . W[N,K]^T + bias): tcgen05 GEMM, TMA-fed, fp32 accumulate in TMEM.
 * Replaces nn.Linear (cuBLAS) at wan/modules/causal_model.py:112-114 (q,k,v -- one call with the
 * three weights stacked and up to three output segments of `seg_cols` columns each), :240 (o),
 * :277-279 (ffn), :351/:366 (head), :458-462 (patch/text embed); wan/modules/model.py:172,177-178,193.
 * `gate` row for output row r is gate + ((r + gate_row_offset) / rows_per_gate) * gate_stride (per-frame adaLN
 * gate; gate_row_offset = chunk-global index of row 0 for sequence-parallel callers that hold a slice of the rows).
 * out may alias residual.  block_n: 0 = choose; 64/128/256 = one-CTA tiles of 128 x block_n; 512 = CTA-pair
 * (tcgen05 cta_group::2) tiles of 256 x 256; 515 = pair tiles in clusters of two pairs that share the A operand by
 * TMA multicast (N and seg_cols must be multiples of 256 for both).  workspace / workspace_bytes are ignored (kept
 * for ABI stability). */
int sfb_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, const void* bias,
                  int M, int N, int K, int epilogue,
                  void* out0, long long ldo0, void* out1, long long ldo1, void* out2, long long ldo2, int seg_cols,
                  const void* residual, long long ldr,
                  const void* gate, long long gate_stride, int rows_per_gate, int gate_row_offset,
                  int block_n, void* workspace, long long workspace_bytes, void* stream);

/* sfb_gemm_bf16 with per-row statistics, so that the row-wise norms around a projection need no pass of their own.
 * A statistics record is float[2] = (mean, M2 = sum (x - mean)^2) of one row over SFB_STATS_CHUNK consecutive columns,
 * laid out [row][chunk]; chunks are merged with Chan's parallel-variance formula.  Needs the CTA-pair tiles (M > 128, N and
 * seg_cols multiples of 256; block_n 0 / 512 / 515).
 *   stats_out != NULL: records of the bf16 OUTPUT rows, [M][N / SFB_STATS_CHUNK].
 *   ln_stats  != NULL: records of the INPUT rows x, [M][K / SFB_STATS_CHUNK]; the affine LayerNorm in front of the Linear
 *     (norm3 + cross-attention q, causal_model.py:324 -> model.py:172) is folded into the epilogue:
 *       y[r][n] = bf16(rstd_r * (acc[r][n] - mean_r * ln_sc[n][0]) + ln_sc[n][1])
 *     where w = bf16(W * norm_weight) (per input channel), ln_sc[n] = (sum_c w[n][c], bias[n] + sum_c norm_bias[c] W[n][c])
 *     as float[2]; epilogue SFB_EPI_BIAS with bias == NULL. */
#define SFB_STATS_CHUNK 128
int sfb_gemm_bf16_stats(const void* x, long long ldx, const void* w, long long ldw, const void* bias,
                        int M, int N, int K, int epilogue,
                        void* out0, long long ldo0, void* out1, long long ldo1, void* out2, long long ldo2, int seg_cols,
                        const void* residual, long long ldr,
                        const void* gate, long long gate_stride, int rows_per_gate, int gate_row_offset,
                        int block_n, void* stats_out, const void* ln_stats, const void* ln_sc, float ln_eps, void* stream);

/* Scratch (bytes) for the stream-K schedule of the CTA-pair GEMM (used when whole tiles would leave the last wave
 * badly filled, e.g. 4680 x 1536 outputs = 114 tiles on 74 CTA pairs): caller-owned, zero-initialised once, one launch
 * at a time per workspace.  workspace == NULL disables stream-K. */
long long sfb_gemm_workspace_bytes(void);

/* softmax(q k^T * scale) v without mask, head_dim 128, K/V read in place from a [B,S,H,128] cache
 * window.  Replaces wan/modules/attention.py:32-202 (flash_attn_varlen_func at :136-150) as called
 * from causal_model.py:230-234 (self-attention over the KV window) and model.py:189 (cross-attn). */
int sfb_attention_fwd(const void* q, long long q_row_stride, long long q_batch_stride,
                      const void* k, const void* v, long long kv_row_stride, long long kv_batch_stride,
                      void* out, long long out_row_stride, long long out_batch_stride,
                      int B, int Lq, int Skv, int H, int head_dim, float softmax_scale,
                      void* workspace, long long workspace_bytes, void* stream);

/* sfb_attention_fwd with the WanRMSNorm of the query (model.py:70-86,172) folded in: q is the UN-normalised projection,
 * q_stats its statistics records [B * Lq][q_chunks] over the full channel width (from sfb_gemm_bf16_stats), and the
 * row's factor rsqrt(mean(q^2) + q_eps) multiplies its scores inside the softmax; the caller multiplies the norm's weight
 * into K once per prompt (k' = k * w_q per channel). */
int sfb_attention_fwd_qnorm(const void* q, long long q_row_stride, long long q_batch_stride,
                            const void* k, const void* v, long long kv_row_stride, long long kv_batch_stride,
                            void* out, long long out_row_stride, long long out_batch_stride,
                            int B, int Lq, int Skv, int H, int head_dim, float softmax_scale,
                            const void* q_stats, int q_chunks, float q_eps,
                            void* workspace, long long workspace_bytes, void* stream);

/* Scratch (bytes) sfb_attention_fwd uses to spread long KV windows evenly over all SMs: the (item, KV step)
 * space is cut into one contiguous range per SM and partial (O, max, sum) results of split items are merged
 * by a second small kernel.  The caller owns the buffer; workspace == NULL disables the split. */
long long sfb_attention_workspace_bytes(void);

/* out[l][r][g][:] = bf16(mod[l][g][:] + e[r * e_row_stride + g * e_group_stride + :])
 * adaLN tables for all layers at once: causal_model.py:310 (G=6) and :365 (head, G=2, broadcast e). */
int sfb_modulation_table(const void* mod, const void* e, void* out, int NL, int R, int G, int C,
                         long long e_row_stride, long long e_group_stride, void* stream);

/* y = bf16(bf16(LN(x)) * bf16(1 + scale[g]) + shift[g]), g = (row + row_offset) / rows_per_mod; LN eps, no affine.
 * causal_model.py:315, :327-328, :366 with WanLayerNorm (model.py:89-99). */
int sfb_ln_modulate(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                    const void* shift, const void* scale, long long mod_stride, int rows_per_mod, int row_offset,
                    void* stream);

/* sfb_ln_modulate with the row's (mean, rstd) merged from the statistics records [rows][stats_ld] that the GEMM which
 * produced x wrote (sfb_gemm_bf16_stats, stats_out): the rows are streamed, no reduction pass. */
int sfb_ln_modulate_stats(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                          const void* shift, const void* scale, long long mod_stride, int rows_per_mod, int row_offset,
                          const void* stats, int stats_ld, void* stream);

/* y = bf16(LN(x) * weight + bias)   (norm3, causal_model.py:268-270,324). */
int sfb_ln_affine(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                  const void* weight, const void* bias, void* stream);

/* y = bf16(bf16(x * rsqrt(mean(x^2) + eps)) * weight) over the full width C (WanRMSNorm, model.py:70-86). */
int sfb_rmsnorm(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                const void* weight, void* stream);

/* Fused QK-RMSNorm + 3-D RoPE + KV-cache append (model.py:70-86, causal_model.py:28-56,196-200,
 * 222-229).  q_in/k_in/v_in: [B*L, C] projections.  RoPE tables cos/sin: fp32 [tab_rows, head_dim/2]
 * (columns = frame | height | width ladders).  q -> q_out[b][n], k/v -> k_out/v_out[b][n] where the
 * caller has already offset k_out/v_out to the cache write slot.  v_in == NULL skips the V copy.
 * start_frame_dev != NULL: the chunk's frame offset is read from that device int at run time instead of start_frame
 * (one captured CUDA graph then serves every chunk of a rolling-window video; the caller keeps start_frame + F within
 * the table). */
int sfb_qk_norm_rope(const void* q_in, long long ldq, const void* k_in, long long ldk, const void* v_in, long long ldv,
                     const void* wq, const void* wk, float eps, const float* cos_tab, const float* sin_tab,
                     int tab_rows, int B, int L, int C, int head_dim, int F, int Hh, int Ww, int start_frame,
                     const int* start_frame_dev, void* q_out, long long q_out_row, long long q_out_batch,
                     void* k_out, void* v_out, long long kv_out_row, long long kv_out_batch, void* stream);

/* sfb_qk_norm_rope with the RMS statistics of the q / k rows taken from the statistics records of the QKV projection
 * (sfb_gemm_bf16_stats, stats_out [B*L][stats_ld] records): q's chunks start at record q_chunk0, k's at k_chunk0.  The rows are
 * streamed (no reduction pass), head_dim 128. */
int sfb_qk_norm_rope_stats(const void* q_in, long long ldq, const void* k_in, long long ldk, const void* v_in, long long ldv,
                           const void* wq, const void* wk, float eps, const void* stats, int stats_ld, int q_chunk0,
                           int k_chunk0, const float* cos_tab, const float* sin_tab,
                           int tab_rows, int B, int L, int C, int head_dim, int F, int Hh, int Ww, int start_frame,
                           const int* start_frame_dev,
                           void* q_out, long long q_out_row, long long q_out_batch,
                           void* k_out, void* v_out, long long kv_out_row, long long kv_out_batch, void* stream);

/* ---- Ulysses head-parallel attention for one long video (wan/distributed/xdit_context_parallel.py:66-192) ----
 * Every rank holds a contiguous slice of the chunk's tokens with ALL heads for the token-wise work and ONE head
 * group for attention.  The two all-to-alls per block are plain stores into peer-mapped memory issued by the
 * producing kernels; sfb_peer_barrier is the only synchronisation. */

/* sfb_qk_norm_rope for `rows` tokens starting at chunk token `token_offset`; head group g (C / groups columns) of
 * every row is stored to q_dst[g] / k_dst[g] / v_dst[g] at row (token_offset + local row).  *_dst are HOST arrays of
 * `groups` device pointers (rank g's q buffer / KV-cache write slot as mapped in this process). */
int sfb_qk_norm_rope_sp(const void* q_in, long long ldq, const void* k_in, long long ldk, const void* v_in, long long ldv,
                        const void* wq, const void* wk, float eps, const float* cos_tab, const float* sin_tab,
                        int tab_rows, int rows, int C, int head_dim, int F, int Hh, int Ww, int start_frame,
                        int token_offset, int groups, void* const* q_dst, long long q_dst_row,
                        void* const* k_dst, void* const* v_dst, long long kv_dst_row, void* stream);

/* sfb_attention_fwd for one sample and this rank's H heads over ALL Lq tokens; output rows
 * [d * rows_per_dst, (d+1) * rows_per_dst) are stored to out_dst[d] (rank d's buffer, already offset to this head
 * group's columns) at row (token - d * rows_per_dst). */
int sfb_attention_fwd_sp(const void* q, long long q_row_stride, const void* k, const void* v, long long kv_row_stride,
                         void* const* out_dst, int n_dst, int rows_per_dst, long long out_row_stride,
                         int Lq, int Skv, int H, int head_dim, float softmax_scale,
                         void* workspace, long long workspace_bytes, void* stream);

/* All-ranks barrier over peer-mapped flag words: flag_ptrs[p] = rank p's zero-initialised int32[n + 1] array as
 * mapped here (slot [n] = the rank's own call counter, advanced by the kernel, so the launch is CUDA-graph
 * replayable); every rank must call it the same number of times. */
int sfb_peer_barrier(void* const* flag_ptrs, int rank, int n, void* stream);

/* Rolling-window eviction of the KV cache (reference wan/modules/causal_model.py:212-221: the kept rows after the sink
 * are shifted left by the number of evicted tokens -- `cache[:, sink:sink+keep] = cache[:, sink+ev:...].clone()`), for
 * all layers' K and V tensors at once.  tensors_dev: DEVICE array of n_tensors pointers to [batch, rows, row_bytes]
 * caches of identical geometry.  Rows [src_row, src_row + n_rows) move to [dst_row, ...), dst_row < src_row. */
int sfb_kv_roll(const void* const* tensors_dev, int n_tensors, int batch, long long batch_stride_bytes,
                long long row_bytes, long long dst_row, long long src_row, long long n_rows, void* stream);

/* im2col of Conv3d(k = s = (1,2,2)) (causal_model.py:775-778): x[b][c][f][y][x] with element strides
 * -> out[(b,f,y/2,x/2)][c*4 + (y%2)*2 + x%2]. */
int sfb_patchify(const void* x, long long sb, long long sc, long long sf, long long sy, long long sx,
                 void* out, int B, int Cin, int F, int H, int W, void* stream);

/* sinusoidal_embedding_1d (model.py:15-25) in f64 -> bf16.  t_dtype: 0 f32, 1 i64, 2 f64, 3 bf16. */
int sfb_sinusoid(const void* t, int t_dtype, void* out, int n, int freq_dim, void* stream);

/* Few-row Linear: y = bf16(in(x) . W^T + b), in = identity | bf16(silu(.)).  The time MLPs at
 * causal_model.py:464-467,829-832 (M = B*F rows). */
int sfb_skinny_linear(const void* x, long long ldx, const void* w, long long ldw, const void* bias,
                      void* y, long long ldy, int M, int N, int K, int silu_in, void* stream);

/* unpatchify (causal_model.py:1081-1104) + flow -> x0 in f64 (wan_wrapper.py:204-228), sigma by
 * nearest timestep (first argmin).  flow / x0 are written as contiguous [B,F,Cout,H,W]. */
int sfb_head_finish(const void* head_out, long long ldh,
                    const void* xt, long long xs_b, long long xs_f, long long xs_c, long long xs_y, long long xs_x,
                    const void* timestep, int t_dtype, const float* timesteps, const float* sigmas, int n_tab,
                    void* flow, void* x0, int B, int F, int Cout, int H, int W, void* stream);

/* FlowMatchScheduler.add_noise (scheduler.py:159-176): out = bf16((1-sigma)*x0 + sigma*noise), fp32. */
int sfb_add_noise(const void* x0, const void* noise, const void* timestep, int t_dtype,
                  const float* timesteps, const float* sigmas, int n_tab, void* out, int n_frames, int per_frame,
                  void* stream);

/* One step of the 50-step sampler in one launch (pipeline/causal_diffusion_inference.py:420-428 +
 * wan/utils/fm_solvers_unipc.py:320-323 flow->x0, :549-626 UniC corrector, :404-484 UniP predictor), with the
 * reference's bf16 rounding after every tensor op:
 *   flow   = flow_uncond + g * (flow_cond - flow_uncond)          (flow_uncond NULL: flow = flow_cond, no guidance)
 *   m_out  = sample - sigma * flow                                (x0 prediction of this step)
 *   sample_out = corrector(last_sample, m0, m1, m_out)            (corrector_order 0: copy of `sample`)
 *   prev_out   = predictor(sample_out, m_out, m0)                 (the next latents)
 * All tensors bf16, contiguous, n elements, 16-byte aligned; m0 / m1 = x0 predictions of the previous two steps.
 * m_out may alias m1 and sample_out may alias last_sample.  coef: HOST array of 12 floats
 *   {g, sigma, c_x, c_m0, c_b, 1/c_rk, c_rho0, c_rho_last, p_x, p_m0, p_b, 1/p_rk}
 * where x = sigma_t/sigma_s, m0 = alpha_t*h_phi_1, b = alpha_t*B_h for the corrector (c_) and predictor (p_). */
int sfb_cfg_unipc_step(const void* flow_cond, const void* flow_uncond, const void* sample, const void* last_sample,
                       const void* m0, const void* m1, void* m_out, void* sample_out, void* prev_out, long long n,
                       const float* coef, int corrector_order, int predictor_order, void* stream);

/* ---- Wan VAE decoder (the step right after the rollout: utils/wan_wrapper.py:94-117 -> wan/modules/vae.py:545-593).
 * Activations are CHANNELS-LAST [T, H, W, C] bf16: one voxel = one row of C channels. ---- */

/* Latents in: out[v][:] = conv2_1x1x1(bf16(bf16(z[:, v] / inv_std) + mean)) for the `voxels` positions of one latent
 * frame (wan_wrapper.py:101-102, vae.py:548-554).  z is channels-first with `z_channel_stride` elements between
 * channels; w [16,16], bias/mean/inv_std [16]; out [voxels, 16]. */
int sfb_vae_latent_in(const void* z, long long z_channel_stride, const void* mean, const void* inv_std,
                      const void* w, const void* bias, void* out, int voxels, void* stream);

/* RMS_norm over the channels of every voxel (vae.py:39-54: F.normalize * sqrt(C) * gamma) with the reference's bf16
 * rounding after each op, optionally followed by SiLU (vae.py:195-198).  C multiple of 8, <= 512. */
int sfb_vae_norm_silu(const void* x, long long ldx, const void* gamma, void* y, long long ldy, long long rows, int C,
                      int silu, void* stream);

/* CausalConv3d (vae.py:17-36) and the resample Conv2d (+ nearest 2x upsampling in front, vae.py:75-83) on
 * channels-last frames: x [t_in, H, W, Cin] holds the cached frames followed by the new ones, `t_zero_pad` virtual
 * zero frames stand in front (causal padding not covered by the cache); kernel (kt, ks, ks) with kt in 1..3,
 * ks in {1,3}, "same" spatial zero padding.  w is packed [Cout, kt*ks*ks*Cin] with K order (dt, dh, dw, c).
 * y = bf16(acc + bias) or, with `residual`, bf16(residual + bf16(acc + bias)) (ResidualBlock, vae.py:220).
 * Output rows = (t_in + t_zero_pad - kt + 1) * Ho * Wo voxels; columns [0, seg_cols) go to y0 and, if Cout is
 * 2 * seg_cols, columns [seg_cols, 2 seg_cols) to y1 (the time_conv whose channel halves are alternate frames,
 * vae.py:141-144); seg_cols 0 = Cout.  The gathered operand is staged in `workspace` in row chunks (any size >=
 * 128 rows works; sfb_causal_conv3d_workspace_bytes(rows, ...) = one chunk).  1x1x1 needs no workspace.
 * workspace == NULL selects the implicit-GEMM kernel (ks = 3, upsample2x = 0, one segment, Cin % 16 == 0): the shifted
 * input boxes go from global memory straight into the tensor-core pipeline by TMA, nothing is staged. */
int sfb_causal_conv3d_cl(const void* x, int t_in, int H, int W, int Cin, int t_zero_pad, int upsample2x,
                         const void* w, const void* bias, int Cout, int kt, int ks,
                         const void* residual, long long ldr, void* y0, void* y1, long long ldo, int seg_cols,
                         void* workspace, long long workspace_bytes, void* stream);
long long sfb_causal_conv3d_workspace_bytes(long long rows, int Cin, int kt, int ks);

/* Nearest-neighbour 2x upsampling of H and W (vae.py:57-63), channels-last: x [T, H, W, C] -> y [T, 2H, 2W, C]. */
int sfb_upsample2x_cl(const void* x, void* y, int T, int H, int W, int C, void* stream);

/* p[r][:] = bf16(softmax(scale * s[r][:])) for fp32 score rows s (written by sfb_gemm_bf16 with SFB_EPI_F32, so the
 * logits are never rounded to bf16 -- the single-head attention of vae.py:251-255 is two GEMMs around this kernel). */
int sfb_softmax_rows(const void* s, long long lds, void* p, long long ldp, int rows, int cols, float scale, void* stream);

/* out[c][r] = in[r][c] for a bf16 matrix (V^T for the attention's second GEMM). */
int sfb_transpose_bf16(const void* in, long long ldi, void* out, long long ldo, int R, int C, void* stream);

/* Decoder output [T*HW, ldy >= 3] bf16 (channels-last RGB) -> fp32 [T, 3, HW] clamped to [-1, 1] (wan_wrapper.py:110). */
int sfb_vae_pixel_out(const void* y, int ldy, void* out, int T, long long HW, void* stream);

/* ---- UMT5 text encoder (the step right before the rollout: utils/wan_wrapper.py:38-52 -> wan/modules/t5.py:303-312);
 * its projections are sfb_gemm_bf16 calls.  Written after round 1's GPU budget was spent: not yet validated on hardware. */

/* T5LayerNorm (t5.py:61-66): y = w * bf16(x * rsqrt(mean(x^2) + eps)), statistics in fp32.  C multiple of 8. */
int sfb_t5_rmsnorm(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps, const void* weight,
                   void* stream);

/* p[r][:] = bf16(softmax_fp32(bf16(s[r][:] + bias[r][:]))) with keys whose key_mask[c] == 0 held at finfo(bf16).min
 * (t5.py:103-115: un-scaled logits + relative position bias; key_mask may be NULL).  s, bias, p bf16. */
int sfb_softmax_bias_rows(const void* s, long long lds, const void* bias, long long ldb, const int* key_mask, void* p,
                          long long ldp, int rows, int cols, void* stream);

/* out = bf16(fc1 * gelu(gate)) with the tanh GELU evaluated op by op in bf16 like t5.py:46-50, :136-137. */
int sfb_t5_gated_gelu(const void* fc1, long long ld1, const void* gate, long long ldg, void* out, long long ldo, int rows,
                      int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H_ */
