/*
 * mofa_b200.h -- C ABI of libmofa_b200.so, the sm_100a kernel library behind the MOFA-Video hot path
 * (SVD denoise loop + MOFA-Adapter/FlowControlNet + flow warp).
 *
 * The reference has no FFI layer of its own: its one native boundary is the CuPy launch of the
 * softsplat CUDA string with raw data_ptr()s on torch's current stream
 * (/root/reference/MOFA-Video-Traj/models/softsplat.py:219-226, 340-345).  Every entry point here
 * follows the same contract: raw device pointers owned by the caller (PyTorch's allocator), plain
 * sizes, a caller-supplied cudaStream_t, no device allocation, no synchronisation, int return
 * (0 = ok, negative = error, text via mofa_last_error()).  Each op cites the reference code whose
 * arithmetic it replaces.  All activations are fp16, channels-last ("NHWC": [frames, h*w, C]).
 */
#ifndef MOFA_B200_H
#define MOFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mofa_stream_t; /* cudaStream_t */

#define MOFA_OK 0
#define MOFA_ERR_ARG (-1)
#define MOFA_ERR_CUDA (-2)
#define MOFA_ERR_UNSUPPORTED (-3)

const char* mofa_last_error(void);
int mofa_version(void);
/* number of kernels launched by this library since load / since the last reset (bench "gpu_launches") */
int64_t mofa_launch_count(void);
void mofa_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM family (tcgen05.mma, TMA-fed, TMEM accumulators, persistent tiles).
 *   out[row, n] = epilogue( sum_k A[row, k] * Wt[n, k] )
 * A-operand modes (how the 128-row A tile of each k-block is fetched by TMA):
 *   LINEAR    A is [M, K] row-major (lda); optionally split K = K1 (from a) + K-K1 (from a2):
 *             torch Linear / 1x1 conv / channel-concat (diffusers up-block torch.cat + conv_shortcut).
 *   CONV3X3   A is NHWC [n_img, H, W, C]; implicit GEMM of a 3x3 / stride 1 / pad 1 convolution,
 *             K = 9*C ordered (ky, kx, c); each k-block is a shifted 4-D TMA box, halo zero-filled
 *             by TMA out-of-bounds handling (replaces cuDNN Conv2d inside ResnetBlock2D etc.).
 *   TEMPORAL3 A is [B, T, HW, C]; implicit GEMM of Conv3d kernel (3,1,1) pad (1,0,0) along T,
 *             K = 3*C ordered (kt, c) (TemporalResnetBlock of SpatioTemporalResBlock).
 * Epilogue (fp32):  v = acc + bias[n] + rowbias[row / rows_per_group  (or row % rowbias_mod), n]
 *                   v = act(v)           act: 0 none, 1 SiLU, 2 GEGLU (see below)
 *                   v = alpha * v + beta1 * res1[row, n] + beta2 * res2[row, n]      -> fp16
 * GEGLU (act=2): weight rows are packed per N tile as [bn/2 value rows | bn/2 gate rows]; the
 *   tile writes bn/2 columns  value * gelu_erf(gate)  (diffusers GEGLU); out has N/2 columns.
 * ---------------------------------------------------------------------------------------------- */
#define MOFA_A_LINEAR 0
#define MOFA_A_CONV3X3 1
#define MOFA_A_TEMPORAL3 2

typedef struct mofa_gemm_args {
  int32_t mode;
  int32_t act;            /* 0 none, 1 silu, 2 geglu, 3 relu, 4 sigmoid */
  const void* a;          /* fp16 */
  const void* a2;         /* fp16, LINEAR split-K second source or NULL */
  const void* w;          /* fp16 [N, Ktot] row-major */
  void* out;              /* fp16 [rows, ldc] */
  int64_t ldc;
  /* LINEAR */
  int64_t M, K, K1, lda, lda2;
  /* CONV3X3: NHWC input */
  int32_t n_img, H, W, C;
  /* TEMPORAL3: [B, T, HW, C] (C shared with conv) */
  int32_t B, T, HW;
  int32_t N;              /* weight rows */
  int32_t bn;             /* N tile (multiple of 16, <= 256) */
  const void* bias;       /* fp16 [N] or NULL */
  const void* rowbias;    /* fp16 [groups, ld_rowbias] or NULL */
  int64_t ld_rowbias;
  int64_t rows_per_group;
  int64_t rowbias_mod;    /* > 0: rowbias row = row % rowbias_mod (diffusers 0.24 temporal cross-attention
                             context ordering quirk); 0: rowbias row = row / rows_per_group */
  const void* res1;       /* fp16 [rows, ldr1] or NULL */
  int64_t ldr1;
  const void* res2;
  int64_t ldr2;
  float alpha, beta1, beta2;
  int32_t max_ctas;       /* 0 = one per SM */
  int32_t dilation;       /* CONV3X3 only: tap spacing (0/1 = dense; 2, 4 = CMP's dilated ResNet stages) */
  int32_t ksize;          /* CONV3X3 mode kernel size: 0/3 = 3x3, 5, 7 ("same" padding, K = ksize^2 * C;
                             7x7 = ForegroundMatting heads of the Keypoint adapter) */
  /* GroupNorm statistics of THIS output for the GroupNorm that consumes it next (conv -> GroupNorm -> SiLU -> conv
   * chains of ResnetBlock2D / TemporalResnetBlock): the epilogue adds sum and sum of squares of the fp16-rounded
   * outputs into gn_stats[(row / gn_rows_per_stat) * gn_groups + (gn_c_off + col) / gn_cpg][0..1] (fp32, zeroed by this
   * call), so mofa_groupnorm runs its apply pass only (silu flag bit 1) -- one read of the tensor saved.  NULL = off. */
  float* gn_stats;
  int64_t gn_rows_per_stat;
  int32_t gn_groups;      /* groups of the consuming GroupNorm (32) */
  int32_t gn_cpg;         /* channels per group of the consumer (even) */
  int32_t gn_c_off;       /* channel offset of this tensor inside the consumer's (concatenated) input */
  int32_t gn_pad;
} mofa_gemm_args;

int mofa_gemm(const mofa_gemm_args* args, mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Spatial self-attention (diffusers Attention + AttnProcessor2_0 / F.scaled_dot_product_attention,
 * no mask): qkv is [frames*L, 3*C] fp16 (q | k | v thirds, heads of 64 inside each third);
 * out[frames*L, C].  FlashAttention-style: S=QK^T and PV on tcgen05, online softmax in fp32.
 * ---------------------------------------------------------------------------------------------- */
int mofa_attn_spatial(const void* qkv, void* out, int32_t frames, int32_t L, int32_t heads, float scale,
                      mofa_stream_t stream);

/* Temporal self-attention over T (<=32) tokens per (batch, pixel, head) (TemporalBasicTransformerBlock
 * attn1): qkv [B*T*HW, 3*C] rows ordered (b, t, p); out same row order [B*T*HW, C]. */
int mofa_attn_temporal(const void* qkv, void* out, int32_t B, int32_t T, int32_t HW, int32_t heads, float scale,
                       mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation / elementwise (HBM-bound)
 * ---------------------------------------------------------------------------------------------- */
/* GroupNorm(32 groups)+optional SiLU over channels-last input, optionally over the channel concat of
 * two sources (x1[rows,C1] | x2[rows,C2]) (torch.cat([h, skip],1) + GroupNorm in up blocks).
 * Statistics span `rows_per_stat` consecutive rows (h*w for 2-D GroupNorm, T*h*w for the 5-D temporal
 * one).  stats is a caller workspace of (rows/rows_per_stat)*groups*2 floats. */
int mofa_groupnorm(const void* x1, int32_t C1, const void* x2, int32_t C2, const void* gamma, const void* beta,
                   void* out, int64_t rows, int64_t rows_per_stat, int32_t groups, float eps, int32_t silu,
                   float* stats, mofa_stream_t stream);

/* LayerNorm over C per row; optional pre-add of a per-row-group vector:  x' = x + add[(row / rows_per_group) %
 * add_period]; out = LN(x'); if sum_out != NULL also writes x' (frame position embedding add,
 * TransformerSpatioTemporalModel). */
int mofa_layernorm(const void* x, const void* gamma, const void* beta, void* out, int64_t rows, int32_t C, float eps,
                   const void* add, int64_t rows_per_group, int64_t add_period, void* sum_out, mofa_stream_t stream);

/* out[i] = x[i] + scale * y[i % period]   (period = n for plain axpy; broadcast add otherwise), fp16 */
int mofa_axpy_bcast(const void* x, const void* y, void* out, int64_t n, int64_t period, float scale,
                    mofa_stream_t stream);

/* im2col for 3x3 pad-1 convs with stride s on NHWC fp16: out[n*Ho*Wo, Kpad], K order (ky,kx,c), zero padded */
int mofa_im2col3x3(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride,
                   int32_t Kpad, mofa_stream_t stream);

/* nearest 2x upsample NHWC fp16 (Upsample2D before its conv) */
int mofa_upsample2x(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, mofa_stream_t stream);

/* NCHW <-> NHWC fp16 layout changes at the API boundary; nchw_to_nhwc can write into a channel slice */
int mofa_nchw_to_nhwc(const void* x, void* out, int32_t n_img, int32_t C, int32_t HW, int32_t ldo, int32_t c_off,
                      mofa_stream_t stream);
int mofa_nhwc_to_nchw(const void* x, void* out, int32_t n_img, int32_t C, int32_t HW, int32_t ldi, int32_t c_off,
                      mofa_stream_t stream);

/* small-M linear (M <= 8): out[m,n] = act_out( sum_k act_in(a[m,k]) * w[n,k] + bias[n] ), fp16 io, fp32 acc.
 * act: 0 none, 1 SiLU.  (TimestepEmbedding MLPs, time_emb_proj, collapsed cross-attention vectors.) */
int mofa_linear_small(const void* a, const void* w, const void* bias, void* out, int32_t M, int32_t N, int32_t K,
                      int32_t act_in, int32_t act_out, mofa_stream_t stream);

/* sinusoidal Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[m, dim] fp16 = [cos | sin] */
int mofa_timestep_embedding(const float* t, void* out, int32_t M, int32_t dim, mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Flow warp: softmax-splatting forward, 'avg' mode
 * (/root/reference/MOFA-Video-Traj/models/softsplat.py:232-274, kernel :285-335), for all flows of a
 * clip in one launch, with the adapter's nearest flow pyramid fused
 * (/root/reference/MOFA-Video-Traj/models/svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py:302-319).
 *   feat  fp16 NHWC [hs, ws, C]          first-frame feature at this scale
 *   flow  fp16 NCHW [F, 2, Hf, Wf]       full-resolution flows (the pipeline hands the adapter fp16,
 *                                        pipeline.py:396-397); scale = Hf / hs
 *   acc   fp32 workspace [F, hs, ws, C]  and wsum fp32 [F, hs, ws]  (zeroed by the call)
 *   out   fp16 NHWC [F, hs, ws, C]  =  acc / (wsum + 1e-7)
 * ---------------------------------------------------------------------------------------------- */
int mofa_softsplat_avg(const void* feat, const void* flow, float* acc, float* wsum, void* out, int32_t F, int32_t hs,
                       int32_t ws, int32_t C, int32_t Hf, int32_t Wf, mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused CFG combine + Euler (v-prediction) step + next model input
 * (pipeline.py:449-454, 495-500; scheduling_euler_discrete_karras_fix.py:264-288, 481-520).
 *   noise  fp16 NHWC [2*T, HW, 4] (uncond frames then cond frames)
 *   latents fp32 NCHW [T, 4, HW] in/out (kept in fp32 between steps is NOT what the reference does:
 *           the reference rounds prev_sample to fp16 each step, so `latents_h` fp16 is the state)
 *   latents_h fp16 NCHW [T,4,HW] in/out;  image_latents fp16 NCHW [2, 4, HW] (uncond zeros, cond)
 *   next_in fp16 NHWC [2*T, HW, 8] = cat(latents'/sqrt(sigma_next^2+1), image_latents)
 * If noise == NULL only next_in is produced from the current latents (loop prologue) with sigma.
 * ---------------------------------------------------------------------------------------------- */
int mofa_cfg_euler_step(const void* noise, void* latents_h, const void* image_latents, void* next_in, int32_t T,
                        int32_t HW, float g_min, float g_max, float sigma, float sigma_next, mofa_stream_t stream);
/* Same step with (sigma, sigma_next) read from device memory `sigmas[0..1]` (fp32) at execution time: the launch
 * carries no per-step scalar, so one captured CUDA graph of the whole denoise step (adapter + UNet + this kernel) is
 * replayed for all 25 steps of pipeline.py:447-511 (the timestep reaches mofa_timestep_embedding the same way). */
int mofa_cfg_euler_step_dev(const void* noise, void* latents_h, const void* image_latents, void* next_in, int32_t T,
                            int32_t HW, float g_min, float g_max, const float* sigmas, mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * VAE temporal decoder helpers (diffusers 0.24 TemporalDecoder, called from pipeline.py:194-220).
 * The decoder's convolutions / GroupNorms / temporal convs reuse mofa_gemm / mofa_groupnorm; its single-head
 * d=512 mid-block attention is GEMM -> mofa_softmax_rows -> GEMM.
 * ---------------------------------------------------------------------------------------------- */
/* in-place softmax over each row of x[rows, ld] (first L columns), fp16 storage, fp32 math, L % 8 == 0, L <= 16384 */
int mofa_softmax_rows(void* x, int64_t rows, int32_t L, int64_t ld, mofa_stream_t stream);

/* time_conv_out = Conv3d(3,3,(3,1,1),pad (1,0,0)) over the T frames of a chunk: y fp16 [T, HW, 3] channels-last,
 * w fp32 [3(co),3(ci),3(kt)], b fp32 [3]; writes out_f32 NCHW [T,3,HW] (decoder `sample`) and/or out_u8
 * [T,HW,3] = round(clamp(v/2+0.5,0,1)*255)  (VaeImageProcessor.postprocess, pipeline.py:57-69) */
int mofa_vae_time_conv_out(const void* y, const float* w, const float* b, float* out_f32, void* out_u8, int32_t T,
                           int64_t HW, mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CMP sparse-to-dense flow network helpers (reference:
 * /root/reference/MOFA-Video-Traj/models/cmp/models/{backbone/resnet.py,modules/shallownet.py,modules/decoder.py},
 * utils/visualize_utils.py:6-19).  Convolutions (BatchNorm folded at pack time, ReLU epilogue, dilation 2/4) run on
 * mofa_gemm; these are the streaming pieces around them.  fp16 channels-last everywhere.
 * ---------------------------------------------------------------------------------------------- */
/* general im2col: out[(n,oy,ox), (ky,kx,c)] zero padded to Kpad columns (7x7/s2 stem, 5x5/s2, strided 1x1 / 3x3);
 * pad < 0 selects asymmetric padding (0 before, -pad after): the VAE encoder's F.pad(x,(0,1,0,1)) + stride-2 conv */
int mofa_im2col(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t ksize,
                int32_t stride, int32_t pad, int32_t dilation, int32_t Kpad, mofa_stream_t stream);
/* nn.MaxPool2d (mode 0) / nn.AvgPool2d (mode 1) with kernel ksize, stride, padding */
int mofa_pool2d(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t ksize,
                int32_t stride, int32_t pad, int32_t mode, mofa_stream_t stream);
/* F.interpolate(bilinear, align_corners=True) to Ho x Wo, written into channels [c_off, c_off+C) of rows of ldo */
int mofa_resize_bilinear_ac(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t Ho,
                            int32_t Wo, int32_t ldo, int32_t c_off, mofa_stream_t stream);
/* Fuser.convert_flow: logits fp16 [rows, 2*nbins] -> expected flow fp16 [rows, 2] */
int mofa_cmp_fuser(const void* logits, void* flow, int64_t rows, int32_t nbins, float fmax, mofa_stream_t stream);
/* dst[r, c_off + c] = src[r % period_rows, c]: channel-concat assembly, broadcasting a per-image tensor over frames */
int mofa_copy_cols(const void* src, void* dst, int64_t rows, int32_t C, int64_t period_rows, int32_t ldo,
                   int32_t c_off, mofa_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Keypoint / Hybrid adapter helpers
 * ---------------------------------------------------------------------------------------------- */
/* flow pyramid level written into a channel slice: out[(f,y,x), c_off + c] = fp16(flow[f,c,y*s,x*s] / s), s = Hf/hs
 * (/root/reference/MOFA-Video-Keypoint/models/ldmk_ctrlnet.py:409-417; input of the occlusion hourglass :299-301) */
int mofa_flow_pyramid(const void* flow, void* out, int32_t F, int32_t hs, int32_t ws, int32_t Hf, int32_t Wf,
                      int32_t ldo, int32_t c_off, mofa_stream_t stream);
/* out = a * m + b * (1 - m), m = mask[row % period_rows] fp16: ForegroundMatting blend
 * (/root/reference/MOFA-Video-Keypoint/models/occlusion/hourglass.py:278) and the Hybrid residual blend
 * (/root/reference/MOFA-Video-Hybrid/pipeline/pipeline.py:479-488) */
int mofa_mask_blend(const void* a, const void* b, const void* mask, void* out, int64_t rows, int32_t C,
                    int64_t period_rows, mofa_stream_t stream);
/* nearest F.interpolate(scale_factor = 1/s) on channels-last data (landmark embedding pyramid, ldmk_ctrlnet.py:403-407) */
int mofa_downsample_nearest(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t s,
                            mofa_stream_t stream);
/* Drag-flow post-processing (/root/reference/MOFA-Video-Traj/run_gradio.py:251-255 brush mask, :268-275 nearest resize
 * 384^2 -> HxW with per-axis scaling, :330-333 in-mask / out-mask merge), fp16 NCHW [F, 2, ., .]; brush [Hs, Ws] and
 * flow_out may be NULL */
int mofa_flow_post(const void* flow_in, const void* brush, const void* flow_out, void* out, int32_t F, int32_t Hs,
                   int32_t Ws, int32_t H, int32_t W, mofa_stream_t stream);
/* CLIP-side resize (/root/reference/MOFA-Video-Traj/pipeline/pipeline.py:532-640 _resize_with_antialiasing, called at
 * :123): Gaussian blur + bicubic(align_corners) in one pass, fp32 planes [planes, H, W] -> [planes, Ho, Wo] */
int mofa_resize_antialias(const void* img, void* out, int32_t planes, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                          mofa_stream_t stream);

/* Fused FeedForward of a transformer block (diffusers 0.24 FeedForward(activation_fn="geglu"): GEGLU projection, GELU gate,
 * output Linear; blocks built at unet_spatio_temporal_condition_controlnet.py:169-233 / controlnet_sdv.py:270-309), for
 * C <= 320 (SVD level 0):  out = alpha * (GEGLU(x W1^T + b1) W2^T + b2) + beta1 * res1 + beta2 * res2.
 * x [M, C]; w1_packed [2*hidden, C] and b1_packed [2*hidden] in the GEGLU packing with bn = 128 (per 128 rows: 64 value rows,
 * then their 64 gate rows); w2 [C, hidden]; b2 [C]; out, res1, res2 [M, C] (rows of ldr elements).  The hidden activation
 * stays in tensor memory (second GEMM's A operand), so the [M, hidden] intermediate is never written. */
int mofa_ff_geglu(const void* x, const void* w1_packed, const void* b1_packed, const void* w2, const void* b2, void* out,
                  int64_t M, int32_t C, int32_t hidden, const void* res1, int64_t ldr1, const void* res2, int64_t ldr2,
                  float alpha, float beta1, float beta2, mofa_stream_t stream);
/* Profiling support for the kernel above (no reference counterpart): with MOFA_FF_DEBUG including 2048, block 0 records a
 * clock64() stamp per pipeline stage for its first 64 hidden chunks; this copies the 64 x 16 stamps to host_out. */
int mofa_ff_debug_dump(long long* host_out);

/* Multi-head self-attention for small sequences and head dims other than 64 (CLIP ViT-H/14 image encoder of
 * /root/reference/MOFA-Video-Traj/pipeline/pipeline.py:114-141: 16 heads x 80, 257 tokens): qkv fp16 [n_seq, L, 3*C]
 * (q | k | v per token, C = heads * head_dim), out fp16 [n_seq, L, C]; softmax(q k^T * scale) v in fp32. */
int mofa_attn_small(const void* qkv, void* out, int32_t n_seq, int32_t L, int32_t heads, int32_t head_dim, float scale,
                    mofa_stream_t stream);
/* The same kernel on the temporal layout qkv [B, T, HW, 3*C] -> out [B, T, HW, C]: one sequence of T tokens per (b, pixel).
 * With mofa_attn_small (spatial: n_seq = frames, L = h*w) this is the UNet's attention for checkpoints whose head_dim is not
 * 64 (e.g. the class-default heads (5,10,10,20) of unet_spatio_temporal_condition_controlnet.py:93 -> d = 128);
 * mofa_attn_spatial / mofa_attn_temporal are the d = 64 tensor-core kernels. */
int mofa_attn_small_temporal(const void* qkv, void* out, int32_t B, int32_t T, int32_t HW, int32_t heads,
                             int32_t head_dim, float scale, mofa_stream_t stream);

/* Peer-store gather of the decoded frames (the one kernel -> collective edge of the path).
 * Rank r's decoder tail (mofa_vae_time_conv_out) writes its uint8 frames straight into slot r of a buffer that lives in
 * rank 0's HBM (an IPC-mapped peer pointer passed as `out_u8`): NVLink stores issued by the epilogue itself, no staging
 * copy and no NCCL call.  These three entry points are the synchronisation around it:
 *   mofa_peer_enable  cudaDeviceEnablePeerAccess(current -> peer_device), idempotent
 *   mofa_peer_signal  one thread: __threadfence_system(), then st.release.sys of `value` to `flag` (local or peer memory);
 *                     stream-ordered after the tail kernels, so it publishes all their stores
 *   mofa_peer_wait    n threads spin (ld.acquire.sys, nanosleep) until flags[i] >= value (wrap-safe), at most `timeout_s`
 *                     seconds (then *timed_out = 1 + index of the first silent flag, if timed_out != NULL) */
int mofa_peer_enable(int32_t peer_device);
int mofa_peer_signal(void* flag, uint32_t value, mofa_stream_t stream);
int mofa_peer_wait(const void* flags, int32_t n, uint32_t value, double timeout_s, void* timed_out, mofa_stream_t stream);

/* Sparse motion hints -> dense (flow, mask) planes in front of CMP.
 * mode 0 = get_sparseflow_and_mask_forward (/root/reference/MOFA-Video-Traj/run_gradio.py:61-86): pts float64
 *          [K, Tn, 2] (x, y; entry 0 = start, 1.. = interpolated ends), flow fp32 [Tn-1, H, W, 2] and mask fp32
 *          [Tn-1, H, W], collisions add; sign = -1 for is_backward_flow; B and owner_ws unused.
 * mode 1 = get_sparse_flow + sample_optical_flow (/root/reference/MOFA-Video-Keypoint/utils/utils.py:81-119): pts =
 *          landmarks [B, Tn, K, 2] (fp32, or fp64 when dtype64), flow [B, Tn-1, 2, H, W] in the same dtype, mask uint8
 *          [B, Tn-1, 2, H, W]; assignment, the highest landmark index wins a collision; owner_ws int32 [B, Tn-1, H, W].
 * Outputs are zero-filled here.  The caller validates that mode-0 start points index inside the image (numpy raises
 * IndexError there; negative indices wrap). */
int mofa_sparse_hints(const void* pts, int32_t mode, int32_t dtype64, int32_t B, int32_t Tn, int32_t K, int32_t H,
                      int32_t W, int32_t sign, void* flow, void* mask, int32_t* owner_ws, mofa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MOFA_B200_H */
