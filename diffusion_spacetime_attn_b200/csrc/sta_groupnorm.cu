// sta_groupnorm.cu — fused GroupNorm(32) [+ SiLU] forward and input-gradient backward for NHWC fp16 activations.
//
// Replaces, per call site, the reference's GroupNorm32 + SiLU chain (ldm/modules/diffusionmodules/util.py:214-216:
// `super().forward(x.float()).type(x.dtype)` followed by nn.SiLU, used by every ResBlock openaimodel.py:206-236, the
// UNet head :681-685 and SpatialTransformer.norm attention.py:317): fp32 copy -> moments -> normalise -> fp16 copy
// -> SiLU = 5 launches and ~6 passes over the activation, here 2 launches and 2 reads + 1 write.  Statistics and
// the affine transform are computed in fp32 exactly like the reference's x.float() path; the result is rounded to
// fp16 once (the reference rounds after the norm and again after SiLU).
//
// These are HBM/L2-bound streaming kernels: one 16-byte vector (8 channels) per thread per row, 4 rows in flight per
// thread, rows strided across the block, per-channel partial sums in registers, per-group sums through ONE pass over
// shared memory (teams of 8 threads per group + shuffles), one global atomic per (block, group).  An optional
// per-(sample, channel) fp32 bias is added to x first: the ResBlock's `conv(x) + bias + emb[:, :, None, None]`
// (openaimodel.py:259-268) folded into the normalisation that follows it.  Only d(x) is produced in backward: the UNet weights are frozen during the alpha optimisation.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace sta {

constexpr int kGnGroups = 32;
constexpr int kGnMaxThreads = 512;  // vecs * lanes rounded up to a warp; channels <= 4096
constexpr int kGnUnroll = 4;        // independent 16-byte loads in flight per thread

struct GnParams {
  const __half* x;    // [B, HW, C] (NHWC); with x1: channels [0, c_split) of the (never materialised) concatenation
  const __half* x1;   // optional [B, HW, C - c_split]: channels [c_split, C) (forward: torch.cat fused into the read)
  __half* xcat;       // forward only, optional [B, HW, C]: the concatenation, written as a side output
  __half* out1;       // backward only, optional: dx is split at c_split into out [B, HW, c_split] and out1 [B, HW, C - c_split]
  int c_split;
  const __half* dy;   // backward only
  const __half* dres; // backward only, optional [B, HW, C]: a second gradient of x (residual branch), added to dx
  long long dres_stride;  // elements between its rows (>= C: it may be a channel slice of a wider NHWC tensor)
  const __half* xb;   // optional fp16 [B, C]: added to x before everything else (conv bias + timestep embedding)
  long long xb_stride;  // elements between the rows of xb (>= C)
  const float* gamma;
  const float* beta;
  __half* out;        // y (forward) or dx (backward)
  float* stats;       // forward: [B, 32, 2] = raw (sum, sumsq) of x + xb
  float* bstats;      // backward: [B, 32, 2] = (sum dxhat, sum dxhat*xhat)
  int batch, hw, c, rows_per_block, silu, lanes;
  float eps;
};

// __fdividef: 2 instructions instead of the ~10 of an IEEE division (these kernels are instruction-bound, ncu: 36
// thread-instructions per element in the forward); its 2-ulp error is far below the fp16 rounding of the result
__device__ __forceinline__ float silu_f(float z) { return __fdividef(z, 1.f + __expf(-z)); }
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = __fdividef(1.f, 1.f + __expf(-z));
  return s * fmaf(z, 1.f - s, 1.f);
}

// This thread's 8 channels [ch0, ch0 + 8) of sample b of a [B, HW, C] activation that may live in two tensors split at
// channel c_split (p1 == nullptr: one dense tensor): pointer to row 0 and the row stride in elements.  c_split is a
// multiple of 8, so a 16-byte vector never straddles the two.
template <typename T>
struct GnCol {
  T* base;
  long long rs;
  __device__ __forceinline__ T* row(int r) const { return base + (long long)r * rs; }
};
template <typename T>
__device__ __forceinline__ GnCol<T> gn_col(T* p0, T* p1, int c_split, int c, int hw, int b, int ch0) {
  if (!p1) return {p0 + ((long long)b * hw) * c + ch0, (long long)c};
  if (ch0 < c_split) return {p0 + ((long long)b * hw) * c_split + ch0, (long long)c_split};
  const int c1 = c - c_split;
  return {p1 + ((long long)b * hw) * c1 + (ch0 - c_split), (long long)c1};
}

// thread -> (row lane, 8-channel vector); threads beyond vecs * lanes (warp padding) only help in the reductions
struct GnMap {
  int vec, lane, vecs, lanes, c0;
  bool active;
  __device__ GnMap(int c, int lanes_) {
    vecs = c >> 3;
    lanes = lanes_;
    vec = threadIdx.x % vecs;
    lane = threadIdx.x / vecs;
    c0 = vec << 3;
    active = lane < lanes;
  }
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_half2(f[0], f[1]);
  o.y = pack_half2(f[2], f[3]);
  o.z = pack_half2(f[4], f[5]);
  o.w = pack_half2(f[6], f[7]);
  return o;
}

__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Block reduction of per-thread, per-channel partials (a[8], b[8]) to per-group sums, then ONE global atomic per
// (block, group, quantity).  Shared-memory float atomics compile to CAS spin loops (ATOMS.CAST.SPIN) and serialised
// the first version of these kernels; here the partials go through shared memory once and are summed by teams of 8
// threads per group with shuffles.
__device__ __forceinline__ void gn_block_reduce(float* part_a, float* part_b, const float* a, const float* b,
                                                const GnMap& m, int c, float* out /* [32][2] of this batch row */) {
  const int cg = c / kGnGroups;
  if (m.active) {
    float4* pa = reinterpret_cast<float4*>(part_a + m.lane * c + m.c0);
    float4* pb = reinterpret_cast<float4*>(part_b + m.lane * c + m.c0);
    pa[0] = make_float4(a[0], a[1], a[2], a[3]);
    pa[1] = make_float4(a[4], a[5], a[6], a[7]);
    pb[0] = make_float4(b[0], b[1], b[2], b[3]);
    pb[1] = make_float4(b[4], b[5], b[6], b[7]);
  }
  __syncthreads();
  const int team = threadIdx.x >> 3, sub = threadIdx.x & 7, teams = blockDim.x >> 3;
  const unsigned mask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const int cnt = m.lanes * cg;
  for (int g = team; g < kGnGroups; g += teams) {
    float sa = 0.f, sb = 0.f;
    for (int idx = sub; idx < cnt; idx += 8) {
      const int lr = idx / cg, j = idx - lr * cg;
      sa += part_a[lr * c + g * cg + j];
      sb += part_b[lr * c + g * cg + j];
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(mask, sa, o);
      sb += __shfl_xor_sync(mask, sb, o);
    }
    if (sub == 0) {
      atomicAdd(&out[g * 2], sa);
      atomicAdd(&out[g * 2 + 1], sb);
    }
  }
}

// ---- forward pass 1: per (b, group) sum and sum of squares ---------------------------------------------------------
__global__ void __launch_bounds__(kGnMaxThreads) gn_stats_kernel(GnParams p) {
  extern __shared__ float gn_smem[];
  float* part_s = gn_smem;
  float* part_q = gn_smem + p.lanes * p.c;
  const GnMap m(p.c, p.lanes);
  const int b = blockIdx.y;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    float xb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p.xb) unpack8(*reinterpret_cast<const uint4*>(p.xb + (long long)b * p.xb_stride + m.c0), xb);
    const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
    const GnCol<const __half> src = gn_col(p.x, p.x1, p.c_split, p.c, p.hw, b, m.c0);
    for (int r = r0 + m.lane; r < r1; r += kGnUnroll * m.lanes) {
      uint4 v[kGnUnroll];
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        const int rr = r + u * m.lanes;
        if (rr < r1) v[u] = *reinterpret_cast<const uint4*>(src.row(rr));
      }
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        if (r + u * m.lanes < r1) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float t = f[i] + xb[i];
            s[i] += t;
            q[i] = fmaf(t, t, q[i]);
          }
        }
      }
    }
  }
  gn_block_reduce(part_s, part_q, s, q, m, p.c, p.stats + (long long)b * kGnGroups * 2);
}

// per-thread constants of the 8 channels it owns
struct GnChan {
  float rstd[8], shx[8], gam[8], bet[8];  // xhat = x * rstd + shx,  shx = (x_bias - mean) * rstd
};

__device__ __forceinline__ void gn_load_chan(const GnParams& p, const GnMap& m, int b, GnChan& k) {
  const int cg = p.c / kGnGroups;
  const float inv_n = 1.f / ((float)p.hw * (float)cg);
  load8(p.gamma + m.c0, k.gam);
  load8(p.beta + m.c0, k.bet);
#pragma unroll
  for (int i = 0; i < 8; ++i) k.shx[i] = 0.f;
  if (p.xb) unpack8(*reinterpret_cast<const uint4*>(p.xb + (long long)b * p.xb_stride + m.c0), k.shx);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (m.c0 + i) / cg;
    const float mean = p.stats[(b * kGnGroups + g) * 2] * inv_n;
    const float var = fmaxf(p.stats[(b * kGnGroups + g) * 2 + 1] * inv_n - mean * mean, 0.f);
    k.rstd[i] = rsqrtf(var + p.eps);
    k.shx[i] = (k.shx[i] - mean) * k.rstd[i];
  }
}

// ---- forward pass 2: y = silu?( (x + xb - mean) * rstd * gamma + beta ) --------------------------------------------
__global__ void __launch_bounds__(kGnMaxThreads) gn_apply_kernel(GnParams p) {
  const GnMap m(p.c, p.lanes);
  const int b = blockIdx.y;
  if (!m.active) return;
  GnChan k;
  gn_load_chan(p, m, b, k);
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = k.rstd[i] * k.gam[i];
    sh[i] = fmaf(k.shx[i], k.gam[i], k.bet[i]);
  }
  const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
  const long long off = ((long long)b * p.hw) * p.c + m.c0;
  const GnCol<const __half> src = gn_col(p.x, p.x1, p.c_split, p.c, p.hw, b, m.c0);
  for (int r = r0 + m.lane; r < r1; r += kGnUnroll * m.lanes) {
    uint4 v[kGnUnroll];
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) {
      const int rr = r + u * m.lanes;
      if (rr < r1) v[u] = *reinterpret_cast<const uint4*>(src.row(rr));
    }
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) {
      const int rr = r + u * m.lanes;
      if (rr < r1) {
        float f[8];
        if (p.xcat) *reinterpret_cast<uint4*>(p.xcat + off + (long long)rr * p.c) = v[u];
        unpack8(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = fmaf(f[i], sc[i], sh[i]);
          f[i] = p.silu ? silu_f(z) : z;
        }
        *reinterpret_cast<uint4*>(p.out + off + (long long)rr * p.c) = pack8(f);
      }
    }
  }
}

// ---- backward pass 1: per (b, group) sum(dxhat) and sum(dxhat * xhat), dxhat = dy * silu'(z) * gamma ----------------
template <bool SILU>
__global__ void __launch_bounds__(kGnMaxThreads) gn_bwd_stats_kernel(GnParams p) {
  extern __shared__ float gn_smem[];
  float* part_1 = gn_smem;
  float* part_2 = gn_smem + p.lanes * p.c;
  const GnMap m(p.c, p.lanes);
  const int b = blockIdx.y;
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    GnChan k;
    gn_load_chan(p, m, b, k);
    const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
    const long long off = ((long long)b * p.hw) * p.c + m.c0;
    constexpr int U = 2;
    for (int r = r0 + m.lane; r < r1; r += U * m.lanes) {
      uint4 vx[U], vd[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = r + u * m.lanes;
        if (rr < r1) {
          vx[u] = *reinterpret_cast<const uint4*>(p.x + off + (long long)rr * p.c);
          vd[u] = *reinterpret_cast<const uint4*>(p.dy + off + (long long)rr * p.c);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u * m.lanes < r1) {
          float f[8], d[8];
          unpack8(vx[u], f);
          unpack8(vd[u], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = fmaf(f[i], k.rstd[i], k.shx[i]);
            float dz = d[i];
            if (SILU) dz *= dsilu_f(fmaf(xh, k.gam[i], k.bet[i]));
            const float dxh = dz * k.gam[i];
            a1[i] += dxh;
            a2[i] = fmaf(dxh, xh, a2[i]);
          }
        }
      }
    }
  }
  gn_block_reduce(part_1, part_2, a1, a2, m, p.c, p.bstats + (long long)b * kGnGroups * 2);
}

// ---- backward pass 2: dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat)) ---------------------------
template <bool SILU>
__global__ void __launch_bounds__(kGnMaxThreads) gn_bwd_apply_kernel(GnParams p) {
  const GnMap m(p.c, p.lanes);
  const int b = blockIdx.y, cg = p.c / kGnGroups;
  if (!m.active) return;
  GnChan k;
  gn_load_chan(p, m, b, k);
  const float inv_n = 1.f / ((float)p.hw * (float)cg);
  float m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (m.c0 + i) / cg;
    m1[i] = p.bstats[(b * kGnGroups + g) * 2] * inv_n;
    m2[i] = p.bstats[(b * kGnGroups + g) * 2 + 1] * inv_n;
  }
  const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
  const long long off = ((long long)b * p.hw) * p.c + m.c0;
  const GnCol<__half> dst = gn_col(p.out, p.out1, p.c_split, p.c, p.hw, b, m.c0);
  constexpr int U = 2;
  for (int r = r0 + m.lane; r < r1; r += U * m.lanes) {
    uint4 vx[U], vd[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = r + u * m.lanes;
      if (rr < r1) {
        vx[u] = *reinterpret_cast<const uint4*>(p.x + off + (long long)rr * p.c);
        vd[u] = *reinterpret_cast<const uint4*>(p.dy + off + (long long)rr * p.c);
        if (p.dres) vr[u] = *reinterpret_cast<const uint4*>(p.dres + ((long long)b * p.hw + rr) * p.dres_stride + m.c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = r + u * m.lanes;
      if (rr < r1) {
        float f[8], d[8];
        unpack8(vx[u], f);
        unpack8(vd[u], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(f[i], k.rstd[i], k.shx[i]);
          float dz = d[i];
          if (SILU) dz *= dsilu_f(fmaf(xh, k.gam[i], k.bet[i]));
          f[i] = k.rstd[i] * (dz * k.gam[i] - m1[i] - xh * m2[i]);
        }
        if (p.dres) {
          unpack8(vr[u], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += d[i];
        }
        *reinterpret_cast<uint4*>(dst.row(rr)) = pack8(f);
      }
    }
  }
}

struct GnLaunch {
  dim3 grid;
  int threads, rows_per_block, lanes;
  size_t smem;  // reduction scratch of the statistics kernels
};

static int gn_launch_shape(const sta_groupnorm_args* a, GnLaunch* L) {
  if (a->channels % kGnGroups != 0 || a->channels % 8 != 0)
    return fail(STA_ERR_UNSUPPORTED, "groupnorm: channels %d must be a multiple of 32 and of 8", a->channels);
  const int vecs = a->channels / 8;
  if (vecs > kGnMaxThreads) return fail(STA_ERR_UNSUPPORTED, "groupnorm: channels %d too large", a->channels);
  int lanes = 256 / vecs;
  if (lanes < 1) lanes = 1;
  L->lanes = lanes;
  L->threads = ((vecs * lanes + 31) / 32) * 32;
  // enough blocks to fill the chip (148 SMs x a few); each thread keeps kGnUnroll rows in flight
  int blocks = (148 * 4 + a->batch - 1) / a->batch;
  int rpb = (a->hw + blocks - 1) / blocks;
  const int quantum = lanes * kGnUnroll;
  rpb = ((rpb + quantum - 1) / quantum) * quantum;
  L->rows_per_block = rpb;
  L->grid = dim3((a->hw + rpb - 1) / rpb, a->batch);
  L->smem = sizeof(float) * 2 * lanes * a->channels;
  return STA_OK;
}

// =====================================================================================================================
// Single-launch kernels on thread-block clusters.  A cluster owns (sample b, a chunk of 1-4 whole groups whose channels
// form whole 16-byte vectors); its CTAs split the rows, every thread keeps its rows in REGISTERS, the per-group sums are
// reduced warp -> CTA (shared memory) -> cluster (distributed shared memory), and the normalisation is applied to the
// registers: ONE launch, the activation is read once, no global atomics and no memset node.  In a captured UNet
// evaluation the two-pass path costs ~12 us per GroupNorm of pure node overhead (memset + 2 dependent kernels); the
// tensors are 1-16 MB, so this matters more than the bytes.  Shapes that do not fit the register budget (the 512^2 VAE
// feature maps, 960 channels at 64^2) take the two-pass kernels above.
// =====================================================================================================================
struct GnClusterParams {
  const __half* x;
  const __half* x1;  // see GnParams: split source (forward), concatenation side output (forward), split dx (backward)
  __half* xcat;
  __half* out1;
  int c_split;
  const __half* dy;
  const __half* dres;  // backward only, optional: added to dx (the gradient of x's other consumer)
  long long dres_stride;  // elements between its rows (>= c)
  const __half* xb;
  long long xb_stride;
  const float* gamma;
  const float* beta;
  __half* out;
  float* stats;   // [B, 32, 2] raw sums of the forward (written by the forward kernel, read by the backward kernel)
  float* bstats;  // backward: [B, 32, 2] written for completeness (same meaning as the two-pass kernels)
  int hw, c, cg;
  int chunk_c;       // channels per cluster: gc groups, a multiple of 8
  int gc;            // groups per cluster
  int vpr;           // 16-byte vectors per row of the chunk
  int lanes;         // row lanes per CTA
  int rows_per_cta;
  int silu;
  float eps;
};

// per-CTA sums of the thread partials a[8], b[8] (8 channels of ONE vector column j) -> red[2*g + {0,1}] for the chunk's
// groups.  A vector spans at most two groups (cg >= 8): each thread first folds its 8 channels into (lo, hi) group
// partials — one float4 in shared memory per thread — then warp g sums the entries that touch group g.
__device__ __forceinline__ void gn_cta_group_sums(float* part, float* red, const float* a, const float* b, bool active,
                                                  int lane_row, int j, const GnClusterParams& p) {
  float4* part4 = reinterpret_cast<float4*>(part);
  {
    const int gl0 = (j * 8) / p.cg;
    const int nb = min(8, (gl0 + 1) * p.cg - j * 8);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // (a_lo, b_lo, a_hi, b_hi)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nb) { v.x += a[i]; v.y += b[i]; }
      else { v.z += a[i]; v.w += b[i]; }
    }
    if (active) part4[lane_row * p.vpr + j] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int g = warp; g < p.gc; g += nwarps) {
    const int j_lo = (g * p.cg) / 8, j_hi = ((g + 1) * p.cg - 1) / 8;  // vector columns touching group g
    float sa = 0.f, sb = 0.f;
    for (int lr = lane; lr < p.lanes; lr += 32) {
      for (int jj = j_lo; jj <= j_hi; ++jj) {
        const float4 v = part4[lr * p.vpr + jj];
        const bool is_lo = (jj * 8) / p.cg == g;
        sa += is_lo ? v.x : v.z;
        sb += is_lo ? v.y : v.w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) { red[2 * g] = sa; red[2 * g + 1] = sb; }
  }
}

// cluster-wide totals of red[0 .. 2*gc) -> tot[0 .. 2*gc) in every CTA's shared memory
__device__ __forceinline__ void gn_cluster_totals(float* red, float* tot, int n_vals) {
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  cluster.sync();  // every CTA's red[] is written
  if ((int)threadIdx.x < n_vals) {
    float acc = 0.f;
    const unsigned cs = cluster.num_blocks();
    for (unsigned r = 0; r < cs; ++r) acc += cluster.map_shared_rank(red, r)[threadIdx.x];
    tot[threadIdx.x] = acc;
  }
  cluster.sync();  // totals visible locally; no CTA may exit (or reuse red[]) while a peer still reads it
}

template <int R>
__global__ void __launch_bounds__(352) gn_cluster_fwd_kernel(GnClusterParams p) {
  extern __shared__ float gn_smem[];
  __shared__ float red[8], tot[8];
  const int j = threadIdx.x % p.vpr, lane_row = threadIdx.x / p.vpr;
  const bool active = lane_row < p.lanes;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int ch0 = chunk * p.chunk_c + j * 8;
  const int r0 = blockIdx.x * p.rows_per_cta, r1 = min(p.hw, r0 + p.rows_per_cta);
  const long long off = ((long long)b * p.hw) * p.c + ch0;
  uint4 v[R];
  float xb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
      if (p.xb) unpack8(*reinterpret_cast<const uint4*>(p.xb + (long long)b * p.xb_stride + ch0), xb);
      const GnCol<const __half> src = gn_col(p.x, p.x1, p.c_split, p.c, p.hw, b, ch0);
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int rr = r0 + lane_row + u * p.lanes;
        if (rr < r1) v[u] = *reinterpret_cast<const uint4*>(src.row(rr));
      }
      if (p.xcat) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const int rr = r0 + lane_row + u * p.lanes;
          if (rr < r1) *reinterpret_cast<uint4*>(p.xcat + off + (long long)rr * p.c) = v[u];
        }
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (r0 + lane_row + u * p.lanes < r1) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float t = f[i] + xb[i];
            s[i] += t;
            q[i] = fmaf(t, t, q[i]);
          }
        }
      }
    }
    gn_cta_group_sums(gn_smem, red, s, q, active, lane_row, j, p);
  }
  float sc[8], sh[8], gam[8], bet[8];
  if (active) {  // issued before the cluster reduction: their L2 latency hides behind its two cluster barriers
    load8(p.gamma + ch0, gam);
    load8(p.beta + ch0, bet);
  }
  gn_cluster_totals(red, tot, 2 * p.gc);
  const int g0 = chunk * p.gc;
  if (blockIdx.x == 0 && (int)threadIdx.x < 2 * p.gc)  // raw sums for the backward (same layout as the two-pass path)
    p.stats[((long long)b * kGnGroups + g0) * 2 + threadIdx.x] = tot[threadIdx.x];
  if (!active) return;
  const float inv_n = 1.f / ((float)p.hw * (float)p.cg);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gl = (j * 8 + i) / p.cg;  // group within the chunk
    const float mean = tot[2 * gl] * inv_n;
    const float var = fmaxf(tot[2 * gl + 1] * inv_n - mean * mean, 0.f);
    sc[i] = rsqrtf(var + p.eps) * gam[i];
    sh[i] = fmaf(xb[i] - mean, sc[i], bet[i]);
  }
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int rr = r0 + lane_row + u * p.lanes;
    if (rr < r1) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = fmaf(f[i], sc[i], sh[i]);
        f[i] = p.silu ? silu_f(z) : z;
      }
      *reinterpret_cast<uint4*>(p.out + off + (long long)rr * p.c) = pack8(f);
    }
  }
}

// Backward.  The kernel is instruction-bound, not bandwidth-bound (the first version spent 89 thread-instructions per
// element, mostly evaluating silu' twice): phase 1 turns each (x, dy) pair into (xhat, dxhat = dy * silu'(z) * gamma),
// accumulates the two group sums and keeps the pair as packed fp16 IN PLACE of the raw data; phase 2 is three FMAs per
// element.  fp16 rounding of xhat / dxhat (2^-11 relative) is below the rounding of the fp16 result itself.
template <int R, bool SILU>
__global__ void __launch_bounds__(640) gn_cluster_bwd_kernel(GnClusterParams p) {
  extern __shared__ float gn_smem[];
  __shared__ float red[8], tot[8];
  const int j = threadIdx.x % p.vpr, lane_row = threadIdx.x / p.vpr;
  const bool active = lane_row < p.lanes;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int ch0 = chunk * p.chunk_c + j * 8;
  const int g0 = chunk * p.gc;
  const int r0 = blockIdx.x * p.rows_per_cta, r1 = min(p.hw, r0 + p.rows_per_cta);
  const long long off = ((long long)b * p.hw) * p.c + ch0;
  const float inv_n = 1.f / ((float)p.hw * (float)p.cg);
  // a vector of 8 channels spans at most two groups (cg >= 8): channels [0, nb) belong to group gl0, the rest to gl0 + 1
  const int gl0 = (j * 8) / p.cg;
  const int nb = min(8, (gl0 + 1) * p.cg - j * 8);
  uint4 vxh[R], vdx[R];  // raw x / dy, then packed xhat / dxhat
  float rs_lo = 0.f, rs_hi = 0.f;
  {
    float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int rr = r0 + lane_row + u * p.lanes;
        if (rr < r1) {
          vxh[u] = *reinterpret_cast<const uint4*>(p.x + off + (long long)rr * p.c);
          vdx[u] = *reinterpret_cast<const uint4*>(p.dy + off + (long long)rr * p.c);
        }
      }
      float gam[8], bet[8], sh[8];  // xhat = x * rs + sh,  sh = (xb - mean) * rs
      load8(p.gamma + ch0, gam);
      load8(p.beta + ch0, bet);
#pragma unroll
      for (int i = 0; i < 8; ++i) sh[i] = 0.f;
      if (p.xb) unpack8(*reinterpret_cast<const uint4*>(p.xb + (long long)b * p.xb_stride + ch0), sh);
      {
        const float* st = p.stats + ((long long)b * kGnGroups + g0 + gl0) * 2;
        const float mean_lo = st[0] * inv_n;
        rs_lo = rsqrtf(fmaxf(st[1] * inv_n - mean_lo * mean_lo, 0.f) + p.eps);
        float mean_hi = 0.f;
        if (nb < 8) {
          mean_hi = st[2] * inv_n;
          rs_hi = rsqrtf(fmaxf(st[3] * inv_n - mean_hi * mean_hi, 0.f) + p.eps);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[i] = (sh[i] - (i < nb ? mean_lo : mean_hi)) * (i < nb ? rs_lo : rs_hi);
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (r0 + lane_row + u * p.lanes < r1) {
          float f[8], d[8];
          unpack8(vxh[u], f);
          unpack8(vdx[u], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = fmaf(f[i], i < nb ? rs_lo : rs_hi, sh[i]);
            float dz = d[i];
            if (SILU) dz *= dsilu_f(fmaf(xh, gam[i], bet[i]));
            const float dxh = dz * gam[i];
            a1[i] += dxh;
            a2[i] = fmaf(dxh, xh, a2[i]);
            f[i] = xh;
            d[i] = dxh;
          }
          vxh[u] = pack8(f);
          vdx[u] = pack8(d);
        }
      }
    }
    gn_cta_group_sums(gn_smem, red, a1, a2, active, lane_row, j, p);
  }
  // the residual-branch gradient is fetched before the cluster reduction (a second DRAM round trip otherwise); a1 / a2 are
  // dead here, so the registers are free
  uint4 vres[R];
  if (p.dres && active) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int rr = r0 + lane_row + u * p.lanes;
      if (rr < r1) vres[u] = *reinterpret_cast<const uint4*>(p.dres + ((long long)b * p.hw + rr) * p.dres_stride + ch0);
    }
  }
  gn_cluster_totals(red, tot, 2 * p.gc);
  if (blockIdx.x == 0 && (int)threadIdx.x < 2 * p.gc && p.bstats)
    p.bstats[((long long)b * kGnGroups + g0) * 2 + threadIdx.x] = tot[threadIdx.x];
  if (!active) return;
  // dx = rs * (dxhat - m1 - xhat * m2) = dxhat * rs - q - xhat * pm,  q = rs * m1,  pm = rs * m2
  const float q_lo = rs_lo * tot[2 * gl0] * inv_n, pm_lo = rs_lo * tot[2 * gl0 + 1] * inv_n;
  const float q_hi = nb < 8 ? rs_hi * tot[2 * gl0 + 2] * inv_n : 0.f, pm_hi = nb < 8 ? rs_hi * tot[2 * gl0 + 3] * inv_n : 0.f;
  const GnCol<__half> dst = gn_col(p.out, p.out1, p.c_split, p.c, p.hw, b, ch0);
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int rr = r0 + lane_row + u * p.lanes;
    if (rr < r1) {
      float f[8], d[8];
      unpack8(vxh[u], f);
      unpack8(vdx[u], d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = fmaf(d[i], i < nb ? rs_lo : rs_hi, -(i < nb ? q_lo : q_hi));
        f[i] = fmaf(-f[i], i < nb ? pm_lo : pm_hi, t);
      }
      if (p.dres) {
        unpack8(vres[u], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += d[i];
      }
      *reinterpret_cast<uint4*>(dst.row(rr)) = pack8(f);
    }
  }
}

template <typename K>
static int gn_cluster_launch(K kernel, const GnClusterParams& p, int cs, int n_chunks, int batch, int threads, size_t smem,
                             cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs, n_chunks, batch);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // co-residency: B200 fits 15 clusters of 8 one-CTA-per-SM blocks (one GPC is short of 16 SMs); a third wave loses
  int max_clusters = 0;
  STA_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
  // two waves still beat the two-pass path (18.8 vs 26.9 us for the [2, 4096, 320] gradient: 16 clusters, 15 fit)
  static const int waves = getenv("STA_GN_MAX_WAVES") ? atoi(getenv("STA_GN_MAX_WAVES")) : 2;
  if (max_clusters * waves < n_chunks * batch) return -1;  // caller falls back to the two-pass kernels
  STA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, p));
  return STA_OK;
}

// Plans and launches the cluster kernel; *launched = false when the shape does not fit (caller falls back to two passes).
static int gn_cluster(const sta_groupnorm_args* a, bool bwd, cudaStream_t s, bool* launched) {
  static const bool off = getenv("STA_GN_TWO_PASS") != nullptr;  // A/B timing and debugging
  *launched = false;
  if (off) return STA_OK;
  const int cg = a->channels / kGnGroups;
  int gc = 1;
  while ((gc * cg) % 8 != 0) gc <<= 1;  // whole groups forming whole 16-byte vectors: gc in {1, 2, 4, 8}
  if (gc > 4 || cg < 8) return STA_OK;  // (the backward's lo / hi split assumes a vector spans <= 2 groups)
  const int chunk_c = gc * cg, vpr = chunk_c / 8, n_chunks = kGnGroups / gc;
  const int t_target = bwd ? 640 : 320, r_max = bwd ? 4 : 8;
  if (vpr > t_target) return STA_OK;
  const int lanes = t_target / vpr;
  const int threads = ((lanes * vpr + 31) / 32) * 32;
  // cluster size: enough CTAs to keep R <= r_max, and (if the rows allow it) >= ~128 CTAs in flight
  int cs = 1;
  while (cs < 8 && ((a->hw + cs - 1) / cs + lanes - 1) / lanes > r_max) cs <<= 1;
  while (cs < 8 && cs * n_chunks * a->batch < 128 && a->hw / (2 * cs) >= lanes) cs <<= 1;
  const int rows_per_cta = (a->hw + cs - 1) / cs;
  const int r = (rows_per_cta + lanes - 1) / lanes;
  if (r > r_max) return STA_OK;
  GnClusterParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.dy = reinterpret_cast<const __half*>(a->d_out);
  p.dres = bwd ? reinterpret_cast<const __half*>(a->d_res) : nullptr;
  p.dres_stride = a->d_res_stride > 0 ? a->d_res_stride : a->channels;
  p.xb = reinterpret_cast<const __half*>(a->x_bias);
  p.xb_stride = a->x_bias_stride > 0 ? a->x_bias_stride : a->channels;
  p.gamma = a->gamma; p.beta = a->beta;
  p.out = reinterpret_cast<__half*>(a->out);
  p.x1 = bwd ? nullptr : reinterpret_cast<const __half*>(a->x1);
  p.xcat = bwd ? nullptr : reinterpret_cast<__half*>(a->x_cat);
  p.out1 = bwd ? reinterpret_cast<__half*>(a->out1) : nullptr;
  p.c_split = a->c_split;
  p.stats = a->stats; p.bstats = a->bwd_stats;
  p.hw = a->hw; p.c = a->channels; p.cg = cg; p.chunk_c = chunk_c; p.gc = gc; p.vpr = vpr; p.lanes = lanes;
  p.rows_per_cta = rows_per_cta; p.silu = a->silu; p.eps = a->eps;
  const size_t smem = sizeof(float4) * (size_t)threads;  // one (lo, hi) partial pair per thread
  int rc;
  if (!bwd) {
    if (r <= 1) rc = gn_cluster_launch(gn_cluster_fwd_kernel<1>, p, cs, n_chunks, a->batch, threads, smem, s);
    else if (r <= 2) rc = gn_cluster_launch(gn_cluster_fwd_kernel<2>, p, cs, n_chunks, a->batch, threads, smem, s);
    else if (r <= 4) rc = gn_cluster_launch(gn_cluster_fwd_kernel<4>, p, cs, n_chunks, a->batch, threads, smem, s);
    else rc = gn_cluster_launch(gn_cluster_fwd_kernel<8>, p, cs, n_chunks, a->batch, threads, smem, s);
  } else if (a->silu) {
    if (r <= 1) rc = gn_cluster_launch(gn_cluster_bwd_kernel<1, true>, p, cs, n_chunks, a->batch, threads, smem, s);
    else if (r <= 2) rc = gn_cluster_launch(gn_cluster_bwd_kernel<2, true>, p, cs, n_chunks, a->batch, threads, smem, s);
    else rc = gn_cluster_launch(gn_cluster_bwd_kernel<4, true>, p, cs, n_chunks, a->batch, threads, smem, s);
  } else {
    if (r <= 1) rc = gn_cluster_launch(gn_cluster_bwd_kernel<1, false>, p, cs, n_chunks, a->batch, threads, smem, s);
    else if (r <= 2) rc = gn_cluster_launch(gn_cluster_bwd_kernel<2, false>, p, cs, n_chunks, a->batch, threads, smem, s);
    else rc = gn_cluster_launch(gn_cluster_bwd_kernel<4, false>, p, cs, n_chunks, a->batch, threads, smem, s);
  }
  if (rc == -1) return STA_OK;  // does not fit one wave
  if (rc) return rc;
  *launched = true;
  return STA_OK;
}

}  // namespace sta

extern "C" int sta_groupnorm_fwd(const sta_groupnorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x || !a->out || !a->gamma || !a->beta || !a->stats) return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: null pointer");
  if (a->batch < 1 || a->hw < 1) return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: empty shape");
  if (a->x_bias && (a->x_bias_stride < 0 || a->x_bias_stride % 8 || (a->x_bias_stride > 0 && a->x_bias_stride < a->channels) ||
                    (reinterpret_cast<uintptr_t>(a->x_bias) & 15u)))
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: x_bias must be 16-byte aligned with a row stride that is 0 or a multiple of 8 >= channels");
  if ((a->x1 || a->x_cat) && (!a->x1 || a->c_split < 8 || a->c_split >= a->channels || a->c_split % 8 ||
                              ((reinterpret_cast<uintptr_t>(a->x1) | reinterpret_cast<uintptr_t>(a->x_cat)) & 15u)))
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: a split source needs x1, 16-byte aligned pointers and 8 <= c_split < channels, c_split %% 8 == 0");
  GnLaunch L;
  int rc = gn_launch_shape(a, &L);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.x1 = reinterpret_cast<const __half*>(a->x1);
  p.xcat = reinterpret_cast<__half*>(a->x_cat);
  p.c_split = a->c_split;
  p.out = reinterpret_cast<__half*>(a->out);
  p.xb = reinterpret_cast<const __half*>(a->x_bias);
  p.xb_stride = a->x_bias_stride > 0 ? a->x_bias_stride : a->channels;
  p.gamma = a->gamma; p.beta = a->beta; p.stats = a->stats;
  p.batch = a->batch; p.hw = a->hw; p.c = a->channels; p.rows_per_block = L.rows_per_block; p.silu = a->silu;
  p.lanes = L.lanes; p.eps = a->eps;
  bool launched = false;
  rc = gn_cluster(a, false, s, &launched);
  if (rc) return rc;
  if (!launched) {
    STA_CUDA_CHECK(cudaMemsetAsync(a->stats, 0, sizeof(float) * a->batch * kGnGroups * 2, s));
    gn_stats_kernel<<<L.grid, L.threads, L.smem, s>>>(p);
    gn_apply_kernel<<<L.grid, L.threads, 0, s>>>(p);
  }
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_groupnorm_bwd(const sta_groupnorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x || !a->d_out || !a->out || !a->gamma || !a->beta || !a->stats || !a->bwd_stats)
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_bwd: null pointer");
  if (a->d_res && ((reinterpret_cast<uintptr_t>(a->d_res) & 15u) || a->d_res_stride < 0 || a->d_res_stride % 8 ||
                   (a->d_res_stride > 0 && a->d_res_stride < a->channels)))
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_bwd: d_res must be 16-byte aligned with a row stride that is 0 or a multiple of 8 >= channels");
  if (a->out1 && (a->c_split < 8 || a->c_split >= a->channels || a->c_split % 8 || (reinterpret_cast<uintptr_t>(a->out1) & 15u)))
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_bwd: a split d_x needs a 16-byte aligned out1 and 8 <= c_split < channels, c_split %% 8 == 0");
  GnLaunch L;
  int rc = gn_launch_shape(a, &L);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.dy = reinterpret_cast<const __half*>(a->d_out);
  p.dres = reinterpret_cast<const __half*>(a->d_res);
  p.dres_stride = a->d_res_stride > 0 ? a->d_res_stride : a->channels;
  p.out = reinterpret_cast<__half*>(a->out);
  p.out1 = reinterpret_cast<__half*>(a->out1);
  p.c_split = a->c_split;
  p.xb = reinterpret_cast<const __half*>(a->x_bias);
  p.xb_stride = a->x_bias_stride > 0 ? a->x_bias_stride : a->channels;
  p.gamma = a->gamma; p.beta = a->beta; p.stats = a->stats; p.bstats = a->bwd_stats;
  p.batch = a->batch; p.hw = a->hw; p.c = a->channels; p.rows_per_block = L.rows_per_block; p.silu = a->silu;
  p.lanes = L.lanes; p.eps = a->eps;
  bool launched = false;
  rc = gn_cluster(a, true, s, &launched);
  if (rc) return rc;
  if (!launched) {
    STA_CUDA_CHECK(cudaMemsetAsync(a->bwd_stats, 0, sizeof(float) * a->batch * kGnGroups * 2, s));
    if (a->silu) {
      gn_bwd_stats_kernel<true><<<L.grid, L.threads, L.smem, s>>>(p);
      gn_bwd_apply_kernel<true><<<L.grid, L.threads, 0, s>>>(p);
    } else {
      gn_bwd_stats_kernel<false><<<L.grid, L.threads, L.smem, s>>>(p);
      gn_bwd_apply_kernel<false><<<L.grid, L.threads, 0, s>>>(p);
    }
  }
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}
