/* sta_b200.h — C ABI of libsta_b200.so: the B200 (sm_100a) attention kernels behind the reference's
 * `ldm.modules.attention` hook API.
 *
 * Reference interfaces replaced (paths under /root/reference/attention_optimization/stable-diffusion/):
 *   sta_sattn_fwd / sta_sattn_bwd   CrossAttention.forward with context=None (self-attention `attn1`),
 *                                   ldm/modules/attention.py:175-197 (einsum QK^T * scale -> softmax -> einsum PV)
 *                                   and its autograd backward.
 *   sta_xattn_fwd / sta_xattn_bwd   the (1 + n_obj) `attn2` calls plus the mask-gated alpha-blend of
 *                                   BasicTransformerBlock._forward, ldm/modules/attention.py:278-294, evaluated
 *                                   BEFORE `to_out` (the blend commutes with the affine `to_out`):
 *                                     out[b]     = A_u[b]                                   b <  B  (uncond rows)
 *                                     out[B + b] = A_g[b] + sum_i m_i[b,p] c_i[b] (A_i[b] - A_u[b])      (cond rows)
 *                                   with A_x = softmax(q k_x^T * scale) v_x, and the backward d(out) ->
 *                                   d(q), d(coef) (contexts are frozen, ldm/models/diffusion/ddpm.py:519-523).
 *
 * Rules of the boundary
 *   - Plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise.
 *   - The caller owns all buffers, including workspaces.  The library allocates nothing per call and creates
 *     no streams: work is enqueued on `stream` (a cudaStream_t / CUstream passed as void*), no host sync inside.
 *   - Return 0 on success, non-zero on error; text via sta_last_error() (thread-local).  Unsupported shapes are
 *     errors, never a fallback.  No exceptions cross the boundary.
 *   - fp16 storage, fp32 softmax/accumulation (what torch.autocast("cuda") gives the reference,
 *     scripts/txt2img-gpt.py:310).
 */
#ifndef STA_B200_H_
#define STA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STA_B200_VERSION 104 /* major*100 + minor */

/* return codes */
#define STA_OK 0
#define STA_ERR_BAD_ARG 1
#define STA_ERR_UNSUPPORTED 2
#define STA_ERR_CUDA 3
#define STA_ERR_DEVICE 4

int sta_version(void);
const char* sta_last_error(void);

/* Reads (and optionally clears) the device-side error word that kernels set when a bounded mbarrier wait
 * times out.  Synchronises the device (cudaMemcpy).  *code_out receives the raw word; returns STA_ERR_DEVICE if
 * it was non-zero. */
int sta_device_error(unsigned int* code_out, int clear);

/* ---------------------------------------------------------------------------------------------------------
 * Self-attention (attn1).  Tokens are laid out [batch, n, heads * head_dim] fp16 with arbitrary token/batch
 * strides (in ELEMENTS) so q/k/v may be slices of one fused projection.  head_dim in {40, 80, 160} (SD-v1 UNet) or 512:
 * the single-head mid-block AttnBlock of SD-v1's KL-VAE decoder, ldm/modules/diffusionmodules/model.py:150-202 (a
 * K-dim-pipelined kernel of its own, csrc/sta_sattn_wide.cu; its backward needs no dq_accum, which may be NULL).  Any
 * other head_dim is STA_ERR_UNSUPPORTED; strides and base pointers must keep every row 16-byte aligned.
 * lse (optional, may be NULL): fp32 [batch, heads, n], natural-log-sum-exp of the scaled scores.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* out;  /* fp16 [batch, n, heads*head_dim], token stride o_token_stride */
  float* lse; /* fp32 [batch, heads, n] or NULL */
  int32_t batch, n, heads, head_dim;
  int64_t q_token_stride, q_batch_stride;
  int64_t k_token_stride, k_batch_stride;
  int64_t v_token_stride, v_batch_stride;
  int64_t o_token_stride, o_batch_stride;
  float scale; /* head_dim ** -0.5 in the reference (attention.py:163) */
} sta_sattn_fwd_args;

int sta_sattn_fwd(const sta_sattn_fwd_args* args, void* stream);

typedef struct {
  const void* q;
  const void* k;
  const void* v;
  const void* out;   /* forward output, fp16 */
  const void* d_out; /* fp16, same layout as out (do_token_stride / do_batch_stride) */
  const float* lse;  /* from the forward */
  void* d_q;         /* fp16 [batch, n, heads*head_dim], token stride dqkv_token_stride (below) */
  void* d_k;
  void* d_v;
  float* dq_accum; /* workspace, fp32 [batch, n, heads*head_dim]; zeroed by the library */
  float* delta;    /* workspace, fp32 [batch, heads, n] */
  int32_t batch, n, heads, head_dim;
  int64_t q_token_stride, q_batch_stride;
  int64_t k_token_stride, k_batch_stride;
  int64_t v_token_stride, v_batch_stride;
  int64_t o_token_stride, o_batch_stride;
  int64_t do_token_stride, do_batch_stride;
  float scale;
  /* token stride (elements) shared by d_q / d_k / d_v, batch stride = n * token stride; 0 = dense (heads*head_dim).
   * 3*heads*head_dim lets the three gradients land in one [batch, n, 3C] buffer = d(fused QKV projection). */
  int64_t dqkv_token_stride;
} sta_sattn_bwd_args;

int sta_sattn_bwd(const sta_sattn_bwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused dual cross-attention + mask-gated alpha-blend (attn2 x (1 + n_obj), attention.py:278-294).
 *   q        fp16 [2*B, n, heads*head_dim]  rows [0,B) = unconditional half, [B,2B) = conditional half
 *                                           (c_in = cat([uc, c]), ldm/models/diffusion/plms.py:306)
 *   k_ctx    fp16 [B, 2 + n_obj, ctx_len, heads*head_dim]   to_k of: slot 0 = the prompt's unconditional
 *   v_ctx    (same)                                          context, slot 1 = global prompt context,
 *                                                            slot 2+i = local description i
 *   mask     u8   [B, n_obj, n]     1 inside object i's disc (attention.py:254-261), else 0
 *   coef     f32  [B, n_obj]        alpha_i for this timestep (plms.py:243, weighting_parameter[:, i])
 *   out      fp16 [2*B, n, heads*head_dim]   pre-`to_out` blended attention output
 *   lse      f32  [B, heads, 2 + n_obj, n]   optional (NULL for inference); slot order as k_ctx.  Written for slots
 *                                            0 and 1 everywhere and for slot 2+i on every aligned run of 32 pixels that
 *                                            touches object i's mask (the only places the backward reads it).
 * ctx_len <= 80 (77 for CLIP).  n_obj in [0, 8].
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* q;
  const void* k_ctx;
  const void* v_ctx;
  const uint8_t* mask;
  const float* coef;
  void* out;
  float* lse;
  int32_t prompts; /* B */
  int32_t n, heads, head_dim, n_obj, ctx_len;
  int64_t q_token_stride, q_batch_stride;
  int64_t o_token_stride, o_batch_stride;
  float scale;
} sta_xattn_fwd_args;

int sta_xattn_fwd(const sta_xattn_fwd_args* args, void* stream);

typedef struct {
  const void* q;
  const void* k_ctx;
  const void* v_ctx;
  const uint8_t* mask;
  const float* coef;
  const float* lse;  /* from the forward */
  const void* d_out; /* fp16 [2*B, n, heads*head_dim] */
  void* d_q;         /* fp16 [2*B, n, heads*head_dim] contiguous */
  float* d_coef;     /* fp32 [B, n_obj]; zeroed by the library, then accumulated atomically */
  int32_t prompts;
  int32_t n, heads, head_dim, n_obj, ctx_len;
  int64_t q_token_stride, q_batch_stride;
  int64_t do_token_stride, do_batch_stride;
  float scale;
  /* the forward's `out` (fp16 [2*B, n, heads*head_dim], strides in elements): its unconditional rows ARE A_u, which
   * gives delta_uc = <d_out_c, A_u> for d(coef) without a third accumulator.  Required when n_obj > 0. */
  const void* out;
  int64_t o_token_stride, o_batch_stride;
} sta_xattn_bwd_args;

int sta_xattn_bwd(const sta_xattn_bwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused GroupNorm(32 groups) [+ SiLU] on NHWC fp16 activations, fp32 statistics (reference GroupNorm32 + nn.SiLU:
 * ldm/modules/diffusionmodules/util.py:214-216 with openaimodel.py:206-236, 681-685 and attention.py:317).
 *   x, out  fp16 [batch, hw, channels] (the memory of a channels_last [batch, channels, h, w] tensor)
 *   gamma, beta  f32 [channels]
 *   stats   f32 [batch, 32, 2]  written by the forward (raw sum / sum of squares per group), read by the backward
 * Backward produces d(x) only (the UNet weights are frozen): out = d_x, d_out = upstream gradient, bwd_stats is a
 * f32 [batch, 32, 2] workspace.  channels must be a multiple of 32 and of 8.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const void* d_out; /* backward only */
  const float* gamma;
  const float* beta;
  void* out;
  float* stats;
  float* bwd_stats; /* backward only */
  int32_t batch, hw, channels, silu;
  float eps;
  /* optional fp16 [batch, channels] (dense) added to x BEFORE the statistics (y = GN(x + x_bias)): the ResBlock's
   * conv bias + projected timestep embedding (openaimodel.py:259-268).  NULL = none.  The backward needs the same
   * values again (it normalises x + x_bias); no gradient is produced for x_bias. */
  const void* x_bias;
  int64_t x_bias_stride; /* elements between the rows of x_bias; 0 = channels (dense).  A larger stride lets x_bias be a
                          * column slice of one [batch, sum of channels] table shared by all ResBlocks. */
  /* backward only, optional fp16 [batch, hw, channels]: a second gradient w.r.t. x that the kernel adds to d_x on the fly.
   * x feeds the GroupNorm AND a residual branch in every ResBlock / SpatialTransformer (openaimodel.py:275,
   * attention.py:332-345); this replaces autograd's separate accumulation pass.  NULL = none; may alias out.
   * d_res_stride: elements between its rows; 0 = channels (dense).  A larger stride lets d_res be a channel slice of a
   * wider NHWC tensor (the gradient of the decoder's torch.cat, openaimodel.py:731). */
  const void* d_res;
  int64_t d_res_stride;
  /* torch.cat([h, skip], dim=1) of the decoder (openaimodel.py:731) fused into the GroupNorm that consumes it.
   * Forward: x1 != NULL means x holds channels [0, c_split) as a dense [batch, hw, c_split] tensor and x1 the remaining
   * channels as [batch, hw, channels - c_split]; x_cat (optional) receives the dense concatenation [batch, hw, channels]
   * (the ResBlock's skip branch and the backward need it).  Backward: x is that dense concatenation; out1 != NULL splits
   * d_x the same way (out: [batch, hw, c_split], out1: the rest), so the gradient of the cat is two dense tensors instead
   * of two strided views.  c_split must be a multiple of 8. */
  const void* x1;
  void* x_cat;
  void* out1;
  int32_t c_split;
} sta_groupnorm_args;

int sta_groupnorm_fwd(const sta_groupnorm_args* args, void* stream);
int sta_groupnorm_bwd(const sta_groupnorm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused (bias +) residual add + LayerNorm on token-major fp16 activations [rows, channels].
 * Replaces, inside BasicTransformerBlock._forward (ldm/modules/attention.py:274, 281/297, 299), the chain
 * `attn(...) + x` -> `norm(x)`: under autocast the reference adds in fp16, copies the sum to fp32, normalises in
 * fp32 and copies the fp32 result back to fp16 for the next Linear.
 *   s = x (+ bias) (+ residual)      rounded to fp16 once, written to sum_out
 *   y = LayerNorm(s) * gamma + beta  fp32 statistics over `channels`, written to y in fp16
 *   x         fp16 [rows, channels]   a GEMM output without its bias, or the activation itself
 *   bias      f32  [channels] or NULL;  residual fp16 [rows, channels] or NULL
 *   gamma, beta f32 [channels]; gamma == NULL: no LayerNorm (only sum_out is produced)
 *   sum_out   fp16 [rows, channels] or NULL (NULL allowed when there is neither bias nor residual)
 *   stats     f32  [rows, 2] = (mean, rstd) or NULL (inference)
 * Backward (weights frozen: d(input) only):  d_x = d_sum + LN'(d_y),  d_sum optional (NULL = 0); `xs` is the
 * forward's sum_out (or x when nothing was added).  channels: multiple of 8, <= 2048; rows dense (stride = channels).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const float* bias;
  const void* residual;
  const float* gamma;
  const float* beta;
  void* sum_out;
  void* y;
  float* stats;
  int32_t rows, channels;
  float eps;
} sta_add_layernorm_args;

int sta_add_layernorm_fwd(const sta_add_layernorm_args* args, void* stream);

typedef struct {
  const void* d_y;   /* fp16 [rows, channels] gradient wrt y */
  const void* d_sum; /* fp16 [rows, channels] gradient wrt sum_out (the residual path) or NULL */
  const void* xs;    /* fp16 [rows, channels] the LayerNorm input saved by the forward */
  const float* stats;
  const float* gamma;
  void* d_x; /* fp16 [rows, channels] */
  int32_t rows, channels;
} sta_add_layernorm_bwd_args;

int sta_add_layernorm_bwd(const sta_add_layernorm_bwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * GEGLU gate (ldm/modules/attention.py:42-49: `x, gate = proj(x).chunk(2, -1); x * F.gelu(gate)`, exact erf GELU).
 *   proj   fp16 [rows, 2*inner]   (value | gate) halves of GEGLU.proj's output
 *   fwd:   out fp16 [rows, inner]    = value * gelu(gate)
 *   bwd:   d_out fp16 [rows, inner] -> out fp16 [rows, 2*inner] = d(proj)   (d value | d gate)
 * inner: multiple of 8.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* proj;
  const void* d_out; /* backward only */
  void* out;
  int32_t rows, inner;
} sta_geglu_args;

int sta_geglu_fwd(const sta_geglu_args* args, void* stream);
int sta_geglu_bwd(const sta_geglu_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Nearest-neighbour x2 upsampling of an NHWC fp16 image (`F.interpolate(x, scale_factor=2, mode="nearest")` in
 * Upsample.forward, ldm/modules/diffusionmodules/openaimodel.py:101-118 and model.py:42-58).
 *   fwd: x fp16 [batch, height, width, channels] -> out fp16 [batch, 2*height, 2*width, channels]
 *   bwd: x = d(out) fp16 [batch, 2*height, 2*width, channels] -> out = d(x) fp16 [batch, height, width, channels]
 * (height / width are ALWAYS those of the low-resolution tensor).  channels: multiple of 8.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  void* out;
  int32_t batch, height, width, channels;
} sta_upsample2x_args;

int sta_upsample2x_fwd(const sta_upsample2x_args* args, void* stream);
int sta_upsample2x_bwd(const sta_upsample2x_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Elementwise part of one PLMS / DDIM sampler step (PLMSSampler.p_sample_plms, ldm/models/diffusion/plms.py:296-358):
 * classifier-free guidance of the UNet's two output rows (:304-308), the Adams-Bashforth combination with up to three
 * previous noise estimates (:341-354) and the eta = 0 update of x (:321-338) in ONE launch, and their gradients in one.
 *   eps      f32 [2*B, elems]   UNet output: rows [0,B) unconditional, [B,2B) conditional (elems = C*H*W per prompt)
 *   x        f32 [B, elems]     current latent;   old[k]  f32 [B, elems] = e_{t-1-k} or NULL
 *   e_t      = (1 - guidance) eps_u + guidance eps_c                 (out, kept by the caller as the next old[0])
 *   e'       = w_e e_t + sum_k w_old[k] old[k]                       (AB weights, e.g. 55/24, -59/24, 37/24, -9/24)
 *   x_prev   = a_x x + a_e e'       pred_x0 = p_x x + p_e e'  (optional)
 * Backward: g_x_prev, g_e_t (either may be NULL = 0) -> g_eps [2*B, elems], g_x, g_old[k] (NULL where old[k] was NULL).
 * Everything fp32, dense, 16-byte aligned; elems a multiple of 4.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* eps;
  const void* x;
  const void* old[3];
  void* e_t;
  void* x_prev;
  void* pred_x0;
  int32_t prompts;
  int64_t elems;
  float guidance, w_e;
  float w_old[3];
  float a_x, a_e, p_x, p_e;
} sta_plms_step_args;

int sta_plms_step_fwd(const sta_plms_step_args* args, void* stream);

typedef struct {
  const void* g_x_prev;
  const void* g_e_t;
  void* g_eps;
  void* g_x;
  void* g_old[3];
  int32_t prompts;
  int64_t elems;
  float guidance, w_e;
  float w_old[3];
  float a_x, a_e;
} sta_plms_step_bwd_args;

int sta_plms_step_bwd(const sta_plms_step_bwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Test hook: one tcgen05 GEMM tile with caller-supplied UMMA descriptors (tests/test_probe_gpu.py pins the
 * shared-memory/TMEM operand encodings the kernels above rely on).  Not part of the product path.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* a; /* fp16 row-major [a_tensor_rows, a_cols]; staged as TMA boxes of a_rows x 64 (OOB rows = 0) */
  int32_t a_rows, a_tensor_rows, a_cols, a_in_tmem;
  const void* b; /* fp16 row-major [b_tensor_rows, b_cols]; staged as TMA boxes of b_rows x 64 */
  int32_t b_rows, b_tensor_rows, b_cols;
  uint64_t a_desc_hi, b_desc_hi; /* descriptor templates with a zero start address */
  int32_t nk;
  uint32_t a_off[16], b_off[16]; /* per K-step byte offsets (TMEM column offsets when a_in_tmem) */
  uint32_t idesc;
  int32_t n;          /* accumulator columns to read back */
  float* out;         /* fp32 [128, n] */
  void* smem_dump;    /* optional: raw image of the staging shared memory */
  int32_t dump_bytes;
  int32_t reps;       /* >1: issue the nk-step MMA group `reps` times back to back (timing) */
  int64_t* cycles;    /* optional: SM cycles from first issue to completion of the last MMA */
} sta_probe_args;

int sta_probe_gemm(const sta_probe_args* args, void* stream);

/* Test hook: TMEM load (mode 0: one x32 load in flight, 1: two) / store (mode 2: x16) throughput of one CTA with
 * `warps` warps; *cycles_dev (device int64) receives the SM cycles for `iters` iterations. */
int sta_probe_tmem_bw(int warps, int iters, int mode, long long* cycles_dev, void* stream);

/* Debug hook: cycle counters written by sta_sattn_bwd when STA_DEBUG_FLAGS & 8 (tools/ only). Synchronises. */
int sta_debug_read(long long* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* STA_B200_H_ */
