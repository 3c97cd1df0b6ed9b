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

namespace sta {

constexpr int kGnGroups = 32;
constexpr int kGnMaxThreads = 512;  // vecs * lanes rounded up to a warp; channels <= 4096
constexpr int kGnUnroll = 4;        // independent 16-byte loads in flight per thread

struct GnParams {
  const __half* x;    // [B, HW, C] (NHWC)
  const __half* dy;   // backward only
  const __half* xb;   // optional fp16 [B, C]: added to x before everything else (conv bias + timestep embedding)
  const float* gamma;
  const float* beta;
  __half* out;        // y (forward) or dx (backward)
  float* stats;       // forward: [B, 32, 2] = raw (sum, sumsq) of x + xb
  float* bstats;      // backward: [B, 32, 2] = (sum dxhat, sum dxhat*xhat)
  int batch, hw, c, rows_per_block, silu, lanes;
  float eps;
};

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
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
    if (p.xb) unpack8(*reinterpret_cast<const uint4*>(p.xb + (long long)b * p.c + m.c0), xb);
    const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
    const __half* base = p.x + ((long long)b * p.hw) * p.c + m.c0;
    for (int r = r0 + m.lane; r < r1; r += kGnUnroll * m.lanes) {
      uint4 v[kGnUnroll];
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        const int rr = r + u * m.lanes;
        if (rr < r1) v[u] = *reinterpret_cast<const uint4*>(base + (long long)rr * p.c);
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
  float mean[8], rstd[8], gam[8], bet[8], xb[8];
};

__device__ __forceinline__ void gn_load_chan(const GnParams& p, const GnMap& m, int b, GnChan& k) {
  const int cg = p.c / kGnGroups;
  const float inv_n = 1.f / ((float)p.hw * (float)cg);
  load8(p.gamma + m.c0, k.gam);
  load8(p.beta + m.c0, k.bet);
#pragma unroll
  for (int i = 0; i < 8; ++i) k.xb[i] = 0.f;
  if (p.xb) unpack8(*reinterpret_cast<const uint4*>(p.xb + (long long)b * p.c + m.c0), k.xb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (m.c0 + i) / cg;
    const float mean = p.stats[(b * kGnGroups + g) * 2] * inv_n;
    const float var = fmaxf(p.stats[(b * kGnGroups + g) * 2 + 1] * inv_n - mean * mean, 0.f);
    k.mean[i] = mean;
    k.rstd[i] = rsqrtf(var + p.eps);
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
    sh[i] = fmaf(k.xb[i] - k.mean[i], sc[i], k.bet[i]);
  }
  const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
  const long long off = ((long long)b * p.hw) * p.c + m.c0;
  for (int r = r0 + m.lane; r < r1; r += kGnUnroll * m.lanes) {
    uint4 v[kGnUnroll];
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) {
      const int rr = r + u * m.lanes;
      if (rr < r1) v[u] = *reinterpret_cast<const uint4*>(p.x + off + (long long)rr * p.c);
    }
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) {
      const int rr = r + u * m.lanes;
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
}

// ---- backward pass 1: per (b, group) sum(dxhat) and sum(dxhat * xhat), dxhat = dy * silu'(z) * gamma ----------------
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
            const float xh = (f[i] + k.xb[i] - k.mean[i]) * k.rstd[i];
            float dz = d[i];
            if (p.silu) dz *= dsilu_f(fmaf(xh, k.gam[i], k.bet[i]));
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
      const int rr = r + u * m.lanes;
      if (rr < r1) {
        float f[8], d[8];
        unpack8(vx[u], f);
        unpack8(vd[u], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (f[i] + k.xb[i] - k.mean[i]) * k.rstd[i];
          float dz = d[i];
          if (p.silu) dz *= dsilu_f(fmaf(xh, k.gam[i], k.bet[i]));
          f[i] = k.rstd[i] * (dz * k.gam[i] - m1[i] - xh * m2[i]);
        }
        *reinterpret_cast<uint4*>(p.out + off + (long long)rr * p.c) = pack8(f);
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

}  // namespace sta

extern "C" int sta_groupnorm_fwd(const sta_groupnorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x || !a->out || !a->gamma || !a->beta || !a->stats) return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: null pointer");
  if (a->batch < 1 || a->hw < 1) return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: empty shape");
  GnLaunch L;
  int rc = gn_launch_shape(a, &L);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.out = reinterpret_cast<__half*>(a->out);
  p.xb = reinterpret_cast<const __half*>(a->x_bias);
  p.gamma = a->gamma; p.beta = a->beta; p.stats = a->stats;
  p.batch = a->batch; p.hw = a->hw; p.c = a->channels; p.rows_per_block = L.rows_per_block; p.silu = a->silu;
  p.lanes = L.lanes; p.eps = a->eps;
  STA_CUDA_CHECK(cudaMemsetAsync(a->stats, 0, sizeof(float) * a->batch * kGnGroups * 2, s));
  gn_stats_kernel<<<L.grid, L.threads, L.smem, s>>>(p);
  gn_apply_kernel<<<L.grid, L.threads, 0, s>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_groupnorm_bwd(const sta_groupnorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x || !a->d_out || !a->out || !a->gamma || !a->beta || !a->stats || !a->bwd_stats)
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_bwd: null pointer");
  GnLaunch L;
  int rc = gn_launch_shape(a, &L);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.dy = reinterpret_cast<const __half*>(a->d_out);
  p.out = reinterpret_cast<__half*>(a->out);
  p.xb = reinterpret_cast<const __half*>(a->x_bias);
  p.gamma = a->gamma; p.beta = a->beta; p.stats = a->stats; p.bstats = a->bwd_stats;
  p.batch = a->batch; p.hw = a->hw; p.c = a->channels; p.rows_per_block = L.rows_per_block; p.silu = a->silu;
  p.lanes = L.lanes; p.eps = a->eps;
  STA_CUDA_CHECK(cudaMemsetAsync(a->bwd_stats, 0, sizeof(float) * a->batch * kGnGroups * 2, s));
  gn_bwd_stats_kernel<<<L.grid, L.threads, L.smem, s>>>(p);
  gn_bwd_apply_kernel<<<L.grid, L.threads, 0, s>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}
