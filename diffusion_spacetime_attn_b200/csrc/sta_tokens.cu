// sta_tokens.cu — streaming kernels on token-major fp16 activations [rows, channels] of the transformer block:
//
//   sta_add_layernorm_fwd/bwd   s = x (+ bias) (+ residual);  y = LayerNorm(s)      (reference attention.py:274, 281/297, 299:
//                               `attn(norm(x)) + x` followed by the next `norm`; under autocast the reference runs
//                               LayerNorm in fp32 on an fp32 COPY of the fp16 activation and copies the fp32 result back
//                               to fp16 for the next Linear — add + copy + LN + copy = 4 launches, 26 B per element; here
//                               1 launch, 8 B per element, statistics still fp32)
//   sta_geglu_fwd/bwd           out = value * gelu(gate) of GEGLU.proj's [rows, 2*inner] output (attention.py:47-49, exact
//                               erf GELU as F.gelu) and d(proj) from d(out) in one pass (the reference's autograd runs
//                               gelu_backward + 2 mul + cat)
//
// All four are HBM-bound: 16-byte vectors, one warp per row for the LayerNorm (row kept in registers between the
// statistics and the normalisation: one read of the inputs, one write of each output), grid-stride vectors for GEGLU.
// Only d(input) is produced in backward — the UNet weights are frozen during the alpha optimisation.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"
#include <stdlib.h>

namespace sta {

__device__ __forceinline__ void tk_unpack8(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 tk_pack8(const float* f) {
  uint4 o;
  o.x = pack_half2(f[0], f[1]);
  o.y = pack_half2(f[2], f[3]);
  o.z = pack_half2(f[4], f[5]);
  o.w = pack_half2(f[6], f[7]);
  return o;
}

__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

struct LnParams {
  const __half* x;
  const float* bias;
  const __half* res;
  const float* gamma;
  const float* beta;
  __half* sum_out;
  __half* y;
  float* stats;
  // backward
  const __half* dy;
  const __half* dsum;
  __half* dx;
  int rows, c;
  int tpr;  // threads per row: 8, 16 or 32 (a warp holds 32 / tpr rows)
  float eps;
};

constexpr int kLnWarps = 4;

// sum over the tpr consecutive lanes that share a row
__device__ __forceinline__ float row_sum(float v, int tpr) {
  for (int o = tpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// tpr lanes per row (32 / tpr rows per warp); a lane owns vectors sub, sub + tpr, ... (NV of them, the tail predicated
// off).  SD-v1 widths are 40 * 2^k vectors -> tpr = 8 * 2^k, NV = 5: no idle lanes, 5 independent 16-byte loads in flight.
// PRE: prefetch bias / gamma / beta with the row (latency-bound launches: up to ~2 M elements); without it they are fetched
// where they are used, which keeps the kernel at 128 registers for the bandwidth-bound shapes (8192 x 320: 4 blocks per SM
// instead of 2).
template <int NV, bool PRE>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_fwd_kernel(LnParams p) {
  const int tpr = p.tpr, rpw = 32 / tpr;
  const int lane = threadIdx.x & (tpr - 1);
  int row = (blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * rpw + ((threadIdx.x & 31) / tpr);
  const bool live = row < p.rows;  // dead rows still take part in the shuffles
  if (!live) row = p.rows - 1;
  const int vecs = p.c >> 3;
  const long long base = (long long)row * p.c;
  // One memory round trip: every load the kernel needs — the row, the residual, the bias, gamma and beta — is issued before
  // the first use of any of them and before the first store (sum_out may alias nothing, but the compiler cannot know: a store
  // between two loads would serialise the round trips).  The first version fetched the bias after x had arrived and gamma /
  // beta after the reductions: three dependent trips, ~1 us of a 5 us kernel.
  uint4 xv[NV], rv[NV];
  float4 bv[NV][2], gv[NV][2], ev[NV][2];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + tpr * i;
    if (vi < vecs) {
      xv[i] = *reinterpret_cast<const uint4*>(p.x + base + vi * 8);
      if (p.res) rv[i] = *reinterpret_cast<const uint4*>(p.res + base + vi * 8);
      if (PRE && p.bias) {
        bv[i][0] = *reinterpret_cast<const float4*>(p.bias + vi * 8);
        bv[i][1] = *reinterpret_cast<const float4*>(p.bias + vi * 8 + 4);
      }
      if (PRE && p.gamma) {
        gv[i][0] = *reinterpret_cast<const float4*>(p.gamma + vi * 8);
        gv[i][1] = *reinterpret_cast<const float4*>(p.gamma + vi * 8 + 4);
        ev[i][0] = *reinterpret_cast<const float4*>(p.beta + vi * 8);
        ev[i][1] = *reinterpret_cast<const float4*>(p.beta + vi * 8 + 4);
      }
    }
  }
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + tpr * i;
    if (vi < vecs) {
      tk_unpack8(xv[i], v[i]);
      bool rounded = true;
      if (p.res) {
        float r[8];
        tk_unpack8(rv[i], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += r[j];
        rounded = false;
      }
      if (p.bias) {
        float b[8] = {bv[i][0].x, bv[i][0].y, bv[i][0].z, bv[i][0].w, bv[i][1].x, bv[i][1].y, bv[i][1].z, bv[i][1].w};
        if (!PRE) load8f(p.bias + vi * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += b[j];
        rounded = false;
      }
      if (!rounded) {  // the sum is stored in fp16; normalise exactly what the backward will read back
        xv[i] = tk_pack8(v[i]);
        tk_unpack8(xv[i], v[i]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  if (p.sum_out && live && (p.res || p.bias)) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + tpr * i;
      if (vi < vecs) *reinterpret_cast<uint4*>(p.sum_out + base + vi * 8) = xv[i];
    }
  }
  if (!p.gamma) return;
  const float inv_c = 1.f / (float)p.c;
  const float mean = row_sum(s, tpr) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + tpr * i < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(row_sum(q, tpr) * inv_c + p.eps);
  if (p.stats && lane == 0 && live) *reinterpret_cast<float2*>(p.stats + 2 * (long long)row) = make_float2(mean, rstd);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + tpr * i;
    if (vi < vecs && live) {
      float g[8] = {gv[i][0].x, gv[i][0].y, gv[i][0].z, gv[i][0].w, gv[i][1].x, gv[i][1].y, gv[i][1].z, gv[i][1].w};
      float b[8] = {ev[i][0].x, ev[i][0].y, ev[i][0].z, ev[i][0].w, ev[i][1].x, ev[i][1].y, ev[i][1].z, ev[i][1].w};
      if (!PRE) {
        load8f(p.gamma + vi * 8, g);
        load8f(p.beta + vi * 8, b);
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, g[j], b[j]);
      *reinterpret_cast<uint4*>(p.y + base + vi * 8) = tk_pack8(o);
    }
  }
}

// dx = dsum + rstd * (dy*gamma - mean_c(dy*gamma) - xhat * mean_c(dy*gamma*xhat))
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_bwd_kernel(LnParams p) {
  const int tpr = p.tpr, rpw = 32 / tpr;
  const int lane = threadIdx.x & (tpr - 1);
  int row = (blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * rpw + ((threadIdx.x & 31) / tpr);
  const bool live = row < p.rows;
  if (!live) row = p.rows - 1;
  const int vecs = p.c >> 3;
  const long long base = (long long)row * p.c;
  const float2 st = *reinterpret_cast<const float2*>(p.stats + 2 * (long long)row);
  const float mean = st.x, rstd = st.y;
  float g[NV][8], xh[NV][8];
  float s1 = 0.f, s2 = 0.f;
  uint4 dv[NV], xv[NV], sv[NV];  // all loads (the residual-stream gradient too) in flight before the first use
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + tpr * i;
    if (vi < vecs) {
      dv[i] = *reinterpret_cast<const uint4*>(p.dy + base + vi * 8);
      xv[i] = *reinterpret_cast<const uint4*>(p.x + base + vi * 8);
      if (p.dsum) sv[i] = *reinterpret_cast<const uint4*>(p.dsum + base + vi * 8);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + tpr * i;
    if (vi < vecs) {
      float gam[8];
      tk_unpack8(dv[i], g[i]);
      tk_unpack8(xv[i], xh[i]);
      load8f(p.gamma + vi * 8, gam);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mean) * rstd;
        g[i][j] *= gam[j];
        s1 += g[i][j];
        s2 = fmaf(g[i][j], xh[i][j], s2);
      }
    }
  }
  const float inv_c = 1.f / (float)p.c;
  s1 = row_sum(s1, tpr) * inv_c;
  s2 = row_sum(s2, tpr) * inv_c;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + tpr * i;
    if (vi < vecs && live) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - s1 - xh[i][j] * s2);
      if (p.dsum) {
        float d[8];
        tk_unpack8(sv[i], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += d[j];
      }
      *reinterpret_cast<uint4*>(p.dx + base + vi * 8) = tk_pack8(o);
    }
  }
}


#define STA_LN_FWD_DISPATCH(PRE, nv, grid, stream, p)                                          \
  switch (nv) {                                                                                \
    case 1: add_ln_fwd_kernel<1, PRE><<<grid, kLnWarps * 32, 0, stream>>>(p); break;           \
    case 2: add_ln_fwd_kernel<2, PRE><<<grid, kLnWarps * 32, 0, stream>>>(p); break;           \
    case 3: add_ln_fwd_kernel<3, PRE><<<grid, kLnWarps * 32, 0, stream>>>(p); break;           \
    case 4: add_ln_fwd_kernel<4, PRE><<<grid, kLnWarps * 32, 0, stream>>>(p); break;           \
    case 5: add_ln_fwd_kernel<5, PRE><<<grid, kLnWarps * 32, 0, stream>>>(p); break;           \
    case 6: add_ln_fwd_kernel<6, false><<<grid, kLnWarps * 32, 0, stream>>>(p); break;         \
    case 7: add_ln_fwd_kernel<7, false><<<grid, kLnWarps * 32, 0, stream>>>(p); break;         \
    default: add_ln_fwd_kernel<8, false><<<grid, kLnWarps * 32, 0, stream>>>(p); break;        \
  }

#define STA_LN_DISPATCH(kernel, nv, grid, stream, p)                             \
  switch (nv) {                                                                  \
    case 1: kernel<1><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    case 2: kernel<2><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    case 3: kernel<3><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    case 4: kernel<4><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    case 5: kernel<5><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    case 6: kernel<6><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    case 7: kernel<7><<<grid, kLnWarps * 32, 0, stream>>>(p); break;             \
    default: kernel<8><<<grid, kLnWarps * 32, 0, stream>>>(p); break;            \
  }

static int ln_shape_ok(int rows, int channels, const char* who) {
  if (rows < 1 || channels < 8) return fail(STA_ERR_BAD_ARG, "%s: empty shape (%d x %d)", who, rows, channels);
  if (channels % 8 != 0 || channels > 2048)
    return fail(STA_ERR_UNSUPPORTED, "%s: channels %d must be a multiple of 8 and <= 2048", who, channels);
  return STA_OK;
}

// threads per row in {8, 16, 32}: the smallest that keeps <= 8 vectors per thread, preferring the one that wastes
// the fewest lanes (vecs = 40 -> 8 x 5, 80 -> 16 x 5, 160 -> 32 x 5)
static int ln_threads_per_row(int vecs, int* nv_out) {
  int best = 32, best_nv = (vecs + 31) / 32, best_waste = best_nv * 32 - vecs;
  for (int tpr = 16; tpr >= 8; tpr >>= 1) {
    const int nv = (vecs + tpr - 1) / tpr;
    if (nv > 8) break;
    const int waste = nv * tpr - vecs;
    if (waste <= best_waste) { best = tpr; best_nv = nv; best_waste = waste; }
  }
  *nv_out = best_nv;
  return best;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- GEGLU ---------------------------------------------------------------------------------------------------------
struct GegluParams {
  const __half* proj;  // [rows, 2*inner]
  const __half* dout;  // [rows, inner] (backward)
  __half* out;         // fwd [rows, inner]; bwd [rows, 2*inner]
  long long vec_total; // rows * inner / 8   (< 2^31, checked by the launcher)
  int inner_vecs;      // inner / 8
  FastDiv div_inner;   // row = vector index / inner_vecs without an integer division (sta_common.cuh).  The first version
                       // divided a 64-bit index by inner_vecs per vector: ~120 of the 440 SASS instructions of the loop body.
};

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) of the exact-erf GELU (attention.py:49, F.gelu's default), and e = exp(-x^2 / 2) (the
// Gaussian the derivative needs).  libdevice's erff made these kernels ISSUE-bound (ncu: 40 thread-instructions per element,
// issue slots 78 % busy, ALU pipe 64 %, 16.1 us for 63 MB at [8192, 2 x 1280]); this is Abramowitz-Stegun 7.1.26 on the
// complementary side, 0.5 erfc(|x| / sqrt 2) = (a1 t + .. + a5 t^5) e / 2 with t = 1 / (1 + p |x| / sqrt 2): one MUFU.RCP, one
// MUFU.EX2, five FMAs.  Absolute error of Phi, x Phi and Phi + x phi <= 4.3e-7 over [-12, 12] in fp32 arithmetic (checked
// against scipy in fp64) — three orders below the fp16 rounding of the outputs.
__device__ __forceinline__ float gelu_cdf_e(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.f));  // the argument is in [1, 1 + 0.33 |x|]: no range issue
  e = fast_exp2(-0.72134752044448170f * x * x);  // exp(-x^2 / 2)
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float h = p * t * e;  // 0.5 erfc(|x| / sqrt 2)
  return x >= 0.f ? 1.f - h : h;
}
__device__ __forceinline__ float gelu_cdf(float x) {
  float e;
  return gelu_cdf_e(x, e);
}

__global__ void __launch_bounds__(256) geglu_fwd_kernel(GegluParams p) {
  const unsigned int total = (unsigned int)p.vec_total, step = gridDim.x * blockDim.x, iv = (unsigned int)p.inner_vecs;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const unsigned int row = fast_div(i, p.div_inner);
    const __half* pr = p.proj + ((size_t)row * iv + i) * 8;  // (row * 2 iv + (i - row * iv)) * 8
    float a[8], g[8];
    tk_unpack8(*reinterpret_cast<const uint4*>(pr), a);
    tk_unpack8(*reinterpret_cast<const uint4*>(pr + (size_t)iv * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= g[j] * gelu_cdf(g[j]);
    *reinterpret_cast<uint4*>(p.out + (size_t)i * 8) = tk_pack8(a);
  }
}

__global__ void __launch_bounds__(256) geglu_bwd_kernel(GegluParams p) {
  const unsigned int total = (unsigned int)p.vec_total, step = gridDim.x * blockDim.x, iv = (unsigned int)p.inner_vecs;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const unsigned int row = fast_div(i, p.div_inner);
    const size_t off = ((size_t)row * iv + i) * 8;  // (row * 2 iv + (i - row * iv)) * 8
    float a[8], g[8], d[8], da[8], dg[8];
    tk_unpack8(*reinterpret_cast<const uint4*>(p.proj + off), a);
    tk_unpack8(*reinterpret_cast<const uint4*>(p.proj + off + (size_t)iv * 8), g);
    tk_unpack8(*reinterpret_cast<const uint4*>(p.dout + (size_t)i * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float e;
      const float cdf = gelu_cdf_e(g[j], e);
      const float pdf = 0.3989422804014327f * e;
      da[j] = d[j] * g[j] * cdf;
      dg[j] = d[j] * a[j] * fmaf(g[j], pdf, cdf);
    }
    *reinterpret_cast<uint4*>(p.out + off) = tk_pack8(da);
    *reinterpret_cast<uint4*>(p.out + off + (size_t)iv * 8) = tk_pack8(dg);
  }
}

// ---- nearest-neighbour x2 upsampling of an NHWC image (openaimodel.py:101-118, model.py:42-58) ------------------------
// ATen's channels_last kernel moves one element per thread (18 us for a 10 MB output on B200); here one 16-byte vector
// per thread: forward out[b, y, x, :] = in[b, y/2, x/2, :]; backward d_in = sum of the four output gradients (fp32 add).
struct Up2Params {
  const __half* src;
  __half* dst;
  long long vec_total;  // vectors of the tensor the thread index runs over (forward: output, backward: input gradient); < 2^31
  int h, w, cvecs;      // INPUT height / width, channel vectors
  FastDiv div_c, div_x, div_y;  // channel vectors, then the x and y extents of the tensor the index runs over
};

__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(Up2Params p) {
  const unsigned int total = (unsigned int)p.vec_total, step = gridDim.x * blockDim.x;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    unsigned int t, cv, ox, oy, b;
    fast_divmod(i, p.div_c, t, cv);   // div_x = 2 w, div_y = 2 h (output extents)
    fast_divmod(t, p.div_x, t, ox);
    fast_divmod(t, p.div_y, b, oy);
    const size_t in = (((size_t)b * p.h + (oy >> 1)) * p.w + (ox >> 1)) * p.cvecs + cv;
    reinterpret_cast<uint4*>(p.dst)[i] = reinterpret_cast<const uint4*>(p.src)[in];
  }
}

__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(Up2Params p) {
  const unsigned int total = (unsigned int)p.vec_total, step = gridDim.x * blockDim.x;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    unsigned int t, cv, x, y, b;
    fast_divmod(i, p.div_c, t, cv);   // div_x = w, div_y = h (input extents)
    fast_divmod(t, p.div_x, t, x);
    fast_divmod(t, p.div_y, b, y);
    const size_t row = (size_t)2 * p.w * p.cvecs;
    const size_t o = (((size_t)b * 2 * p.h + 2 * y) * 2 * p.w + 2 * x) * p.cvecs + cv;
    const uint4* g = reinterpret_cast<const uint4*>(p.src);
    float a[8], c[8], acc[8];
    tk_unpack8(g[o], acc);
    tk_unpack8(g[o + p.cvecs], a);
    tk_unpack8(g[o + row], c);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += a[j] + c[j];
    tk_unpack8(g[o + row + p.cvecs], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += a[j];
    reinterpret_cast<uint4*>(p.dst)[i] = tk_pack8(acc);
  }
}

static int geglu_grid(long long vec_total) {
  long long blocks = (vec_total + 255) / 256;
  const long long cap = 148LL * 16;  // grid-stride beyond 16 resident-ish CTAs per SM
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace sta

extern "C" int sta_add_layernorm_fwd(const sta_add_layernorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x) return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_fwd: null pointer");
  int rc = ln_shape_ok(a->rows, a->channels, "sta_add_layernorm_fwd");
  if (rc) return rc;
  if (a->gamma && (!a->beta || !a->y)) return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_fwd: gamma without beta / y");
  if ((a->bias || a->residual) && !a->sum_out && !a->gamma)
    return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_fwd: nothing to produce");
  if (!a->gamma && !a->sum_out) return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_fwd: nothing to produce");
  if (!aligned16(a->x) || !aligned16(a->residual) || !aligned16(a->sum_out) || !aligned16(a->y) || !aligned16(a->bias) ||
      !aligned16(a->gamma) || !aligned16(a->beta) || (reinterpret_cast<uintptr_t>(a->stats) & 7u))
    return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_fwd: pointers must be 16-byte aligned");
  LnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.bias = a->bias;
  p.res = reinterpret_cast<const __half*>(a->residual);
  p.gamma = a->gamma; p.beta = a->beta;
  p.sum_out = reinterpret_cast<__half*>(a->sum_out);
  p.y = reinterpret_cast<__half*>(a->y);
  p.stats = a->stats;
  p.rows = a->rows; p.c = a->channels; p.eps = a->eps;
  int nv;
  p.tpr = ln_threads_per_row(a->channels / 8, &nv);
  const int rows_per_block = kLnWarps * (32 / p.tpr);
  const unsigned grid = (unsigned)((a->rows + rows_per_block - 1) / rows_per_block);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static const long long pre_max = getenv("STA_LN_PRE_MAX") ? atoll(getenv("STA_LN_PRE_MAX")) : (2ll << 20);  // A/B knob
  if ((long long)a->rows * a->channels <= pre_max) {
    STA_LN_FWD_DISPATCH(true, nv, grid, s, p);
  } else {
    STA_LN_FWD_DISPATCH(false, nv, grid, s, p);
  }
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_add_layernorm_bwd(const sta_add_layernorm_bwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->d_y || !a->xs || !a->stats || !a->gamma || !a->d_x)
    return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_bwd: null pointer");
  int rc = ln_shape_ok(a->rows, a->channels, "sta_add_layernorm_bwd");
  if (rc) return rc;
  if (!aligned16(a->d_y) || !aligned16(a->d_sum) || !aligned16(a->xs) || !aligned16(a->d_x) || !aligned16(a->gamma) ||
      (reinterpret_cast<uintptr_t>(a->stats) & 7u))
    return fail(STA_ERR_BAD_ARG, "sta_add_layernorm_bwd: pointers must be 16-byte aligned");
  LnParams p{};
  p.dy = reinterpret_cast<const __half*>(a->d_y);
  p.dsum = reinterpret_cast<const __half*>(a->d_sum);
  p.x = reinterpret_cast<const __half*>(a->xs);
  p.stats = const_cast<float*>(a->stats);
  p.gamma = a->gamma;
  p.dx = reinterpret_cast<__half*>(a->d_x);
  p.rows = a->rows; p.c = a->channels;
  int nv;
  p.tpr = ln_threads_per_row(a->channels / 8, &nv);
  const int rows_per_block = kLnWarps * (32 / p.tpr);
  const unsigned grid = (unsigned)((a->rows + rows_per_block - 1) / rows_per_block);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  STA_LN_DISPATCH(add_ln_bwd_kernel, nv, grid, s, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

static sta::FastDiv host_fast_div(unsigned int d) {
  sta::FastDiv f;
  f.d = d;
  sta::make_fast_div_raw(d, &f.magic, &f.shift);
  return f;
}
static void geglu_divisor(sta::GegluParams* p) { p->div_inner = host_fast_div((unsigned int)p->inner_vecs); }

static int geglu_check(const sta_geglu_args* a, bool bwd, const char* who) {
  using namespace sta;
  if (!a || !a->proj || !a->out || (bwd && !a->d_out)) return fail(STA_ERR_BAD_ARG, "%s: null pointer", who);
  if (a->rows < 1 || a->inner < 8) return fail(STA_ERR_BAD_ARG, "%s: empty shape", who);
  if (a->inner % 8 != 0) return fail(STA_ERR_UNSUPPORTED, "%s: inner %d must be a multiple of 8", who, a->inner);
  if ((long long)a->rows * (a->inner / 8) >= (1ll << 31))
    return fail(STA_ERR_UNSUPPORTED, "%s: %d x %d exceeds 2^31 vectors (32-bit indexing)", who, a->rows, a->inner);
  if (!aligned16(a->proj) || !aligned16(a->out) || !aligned16(a->d_out))
    return fail(STA_ERR_BAD_ARG, "%s: pointers must be 16-byte aligned", who);
  return STA_OK;
}

extern "C" int sta_geglu_fwd(const sta_geglu_args* a, void* stream) {
  using namespace sta;
  int rc = geglu_check(a, false, "sta_geglu_fwd");
  if (rc) return rc;
  GegluParams p{};
  p.proj = reinterpret_cast<const __half*>(a->proj);
  p.out = reinterpret_cast<__half*>(a->out);
  p.inner_vecs = a->inner / 8;
  p.vec_total = (long long)a->rows * p.inner_vecs;
  geglu_divisor(&p);
  geglu_fwd_kernel<<<geglu_grid(p.vec_total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_geglu_bwd(const sta_geglu_args* a, void* stream) {
  using namespace sta;
  int rc = geglu_check(a, true, "sta_geglu_bwd");
  if (rc) return rc;
  GegluParams p{};
  p.proj = reinterpret_cast<const __half*>(a->proj);
  p.dout = reinterpret_cast<const __half*>(a->d_out);
  p.out = reinterpret_cast<__half*>(a->out);
  p.inner_vecs = a->inner / 8;
  p.vec_total = (long long)a->rows * p.inner_vecs;
  geglu_divisor(&p);
  geglu_bwd_kernel<<<geglu_grid(p.vec_total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

static int upsample_check(const sta_upsample2x_args* a, const char* who) {
  using namespace sta;
  if (!a || !a->x || !a->out) return fail(STA_ERR_BAD_ARG, "%s: null pointer", who);
  if (a->batch < 1 || a->height < 1 || a->width < 1 || a->channels < 8) return fail(STA_ERR_BAD_ARG, "%s: empty shape", who);
  if (a->channels % 8 != 0) return fail(STA_ERR_UNSUPPORTED, "%s: channels %d must be a multiple of 8", who, a->channels);
  if (!aligned16(a->x) || !aligned16(a->out)) return fail(STA_ERR_BAD_ARG, "%s: pointers must be 16-byte aligned", who);
  return STA_OK;
}

extern "C" int sta_upsample2x_fwd(const sta_upsample2x_args* a, void* stream) {
  using namespace sta;
  int rc = upsample_check(a, "sta_upsample2x_fwd");
  if (rc) return rc;
  Up2Params p{};
  p.src = reinterpret_cast<const __half*>(a->x);
  p.dst = reinterpret_cast<__half*>(a->out);
  p.h = a->height; p.w = a->width; p.cvecs = a->channels / 8;
  p.vec_total = (long long)a->batch * 4 * a->height * a->width * p.cvecs;
  if (p.vec_total >= (1ll << 31)) return fail(STA_ERR_UNSUPPORTED, "sta_upsample2x_fwd: more than 2^31 vectors (32-bit indexing)");
  p.div_c = host_fast_div(p.cvecs); p.div_x = host_fast_div(2 * a->width); p.div_y = host_fast_div(2 * a->height);
  upsample2x_fwd_kernel<<<geglu_grid(p.vec_total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_upsample2x_bwd(const sta_upsample2x_args* a, void* stream) {
  using namespace sta;
  int rc = upsample_check(a, "sta_upsample2x_bwd");
  if (rc) return rc;
  Up2Params p{};
  p.src = reinterpret_cast<const __half*>(a->x);
  p.dst = reinterpret_cast<__half*>(a->out);
  p.h = a->height; p.w = a->width; p.cvecs = a->channels / 8;
  p.vec_total = (long long)a->batch * a->height * a->width * p.cvecs;
  if (p.vec_total >= (1ll << 31)) return fail(STA_ERR_UNSUPPORTED, "sta_upsample2x_bwd: more than 2^31 vectors (32-bit indexing)");
  p.div_c = host_fast_div(p.cvecs); p.div_x = host_fast_div(a->width); p.div_y = host_fast_div(a->height);
  upsample2x_bwd_kernel<<<geglu_grid(p.vec_total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}
