// sta_groupnorm.cu — fused GroupNorm(32) [+ SiLU] forward and input-gradient backward for NHWC fp16 activations.
//
// Replaces, per call site, the reference's GroupNorm32 + SiLU chain (ldm/modules/diffusionmodules/util.py:214-216:
// `super().forward(x.float()).type(x.dtype)` followed by nn.SiLU, used by every ResBlock openaimodel.py:206-236, the
// UNet head :681-685 and SpatialTransformer.norm attention.py:317): fp32 copy -> moments -> normalise -> fp16 copy
// -> SiLU = 5 launches and ~6 passes over the activation, here 2 launches and 2 reads + 1 write.  Statistics and
// the affine transform are computed in fp32 exactly like the reference's x.float() path; the result is rounded to
// fp16 once (the reference rounds after the norm and again after SiLU).
//
// These are HBM-bound streaming kernels: one 16-byte vector (8 channels) per thread per row, rows strided across the
// block, per-channel partial sums in registers, per-group sums through shared-memory atomics, one global atomic
// per (block, group).  Only d(x) is produced in backward: the UNet weights are frozen during the alpha optimisation.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

constexpr int kGnGroups = 32;

struct GnParams {
  const __half* x;    // [B, HW, C] (NHWC)
  const __half* dy;   // backward only
  const float* gamma;
  const float* beta;
  __half* out;        // y (forward) or dx (backward)
  float* stats;       // forward: [B, 32, 2] = (sum, sumsq) -> overwritten with (mean, rstd) by the apply kernel's reader
  float* bstats;      // backward: [B, 32, 2] = (sum dxhat, sum dxhat*xhat)
  int batch, hw, c, rows_per_block, silu;
  float eps;
};

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}

// thread -> (row lane, 8-channel vector).  blockDim.x = vecs * lanes.
struct GnMap {
  int vec, lane, vecs, lanes, c0;
  __device__ GnMap(int c) {
    vecs = c >> 3;
    lanes = blockDim.x / vecs;
    vec = threadIdx.x % vecs;
    lane = threadIdx.x / vecs;
    c0 = vec << 3;
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

// ---- forward pass 1: per (b, group) sum and sum of squares ---------------------------------------------------------
__global__ void gn_stats_kernel(GnParams p) {
  __shared__ float gsum[kGnGroups], gsq[kGnGroups];
  const GnMap m(p.c);
  const int b = blockIdx.y, cg = p.c / kGnGroups;
  if (threadIdx.x < kGnGroups) { gsum[threadIdx.x] = 0.f; gsq[threadIdx.x] = 0.f; }
  __syncthreads();
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
  if (m.lane < m.lanes) {
    const __half* base = p.x + ((long long)b * p.hw) * p.c + m.c0;
    for (int r = r0 + m.lane; r < r1; r += m.lanes) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (long long)r * p.c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (m.c0 + i) / cg;
      atomicAdd(&gsum[g], s[i]);
      atomicAdd(&gsq[g], q[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x < kGnGroups) {
    atomicAdd(&p.stats[(b * kGnGroups + threadIdx.x) * 2], gsum[threadIdx.x]);
    atomicAdd(&p.stats[(b * kGnGroups + threadIdx.x) * 2 + 1], gsq[threadIdx.x]);
  }
}

// ---- forward pass 2: y = silu?( (x - mean) * rstd * gamma + beta ) -------------------------------------------------
__global__ void gn_apply_kernel(GnParams p) {
  const GnMap m(p.c);
  const int b = blockIdx.y, cg = p.c / kGnGroups;
  if (m.lane >= m.lanes) return;
  const float inv_n = 1.f / ((float)p.hw * (float)cg);
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = m.c0 + i, g = c / cg;
    const float mean = p.stats[(b * kGnGroups + g) * 2] * inv_n;
    const float var = fmaxf(p.stats[(b * kGnGroups + g) * 2 + 1] * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    sc[i] = rstd * p.gamma[c];
    sh[i] = p.beta[c] - mean * sc[i];
  }
  const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
  const long long off = ((long long)b * p.hw) * p.c + m.c0;
  for (int r = r0 + m.lane; r < r1; r += m.lanes) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(p.x + off + (long long)r * p.c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float z = fmaf(f[i], sc[i], sh[i]);
      f[i] = p.silu ? silu_f(z) : z;
    }
    uint4 o;
    o.x = pack_half2(f[0], f[1]);
    o.y = pack_half2(f[2], f[3]);
    o.z = pack_half2(f[4], f[5]);
    o.w = pack_half2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p.out + off + (long long)r * p.c) = o;
  }
}

// ---- backward pass 1: per (b, group) sum(dxhat) and sum(dxhat * xhat), dxhat = dy * silu'(z) * gamma ----------------
__global__ void gn_bwd_stats_kernel(GnParams p) {
  __shared__ float g1[kGnGroups], g2[kGnGroups];
  const GnMap m(p.c);
  const int b = blockIdx.y, cg = p.c / kGnGroups;
  if (threadIdx.x < kGnGroups) { g1[threadIdx.x] = 0.f; g2[threadIdx.x] = 0.f; }
  __syncthreads();
  if (m.lane < m.lanes) {
    const float inv_n = 1.f / ((float)p.hw * (float)cg);
    float mean[8], rstd[8], gam[8], bet[8], a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = m.c0 + i, g = c / cg;
      mean[i] = p.stats[(b * kGnGroups + g) * 2] * inv_n;
      const float var = fmaxf(p.stats[(b * kGnGroups + g) * 2 + 1] * inv_n - mean[i] * mean[i], 0.f);
      rstd[i] = rsqrtf(var + p.eps);
      gam[i] = p.gamma[c];
      bet[i] = p.beta[c];
    }
    const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
    const long long off = ((long long)b * p.hw) * p.c + m.c0;
    for (int r = r0 + m.lane; r < r1; r += m.lanes) {
      float f[8], d[8];
      unpack8(*reinterpret_cast<const uint4*>(p.x + off + (long long)r * p.c), f);
      unpack8(*reinterpret_cast<const uint4*>(p.dy + off + (long long)r * p.c), d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (f[i] - mean[i]) * rstd[i];
        float dz = d[i];
        if (p.silu) dz *= dsilu_f(fmaf(xh, gam[i], bet[i]));
        const float dxh = dz * gam[i];
        a1[i] += dxh;
        a2[i] = fmaf(dxh, xh, a2[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (m.c0 + i) / cg;
      atomicAdd(&g1[g], a1[i]);
      atomicAdd(&g2[g], a2[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x < kGnGroups) {
    atomicAdd(&p.bstats[(b * kGnGroups + threadIdx.x) * 2], g1[threadIdx.x]);
    atomicAdd(&p.bstats[(b * kGnGroups + threadIdx.x) * 2 + 1], g2[threadIdx.x]);
  }
}

// ---- backward pass 2: dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat)) ---------------------------
__global__ void gn_bwd_apply_kernel(GnParams p) {
  const GnMap m(p.c);
  const int b = blockIdx.y, cg = p.c / kGnGroups;
  if (m.lane >= m.lanes) return;
  const float inv_n = 1.f / ((float)p.hw * (float)cg);
  float mean[8], rstd[8], gam[8], bet[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = m.c0 + i, g = c / cg;
    mean[i] = p.stats[(b * kGnGroups + g) * 2] * inv_n;
    const float var = fmaxf(p.stats[(b * kGnGroups + g) * 2 + 1] * inv_n - mean[i] * mean[i], 0.f);
    rstd[i] = rsqrtf(var + p.eps);
    gam[i] = p.gamma[c];
    bet[i] = p.beta[c];
    m1[i] = p.bstats[(b * kGnGroups + g) * 2] * inv_n;
    m2[i] = p.bstats[(b * kGnGroups + g) * 2 + 1] * inv_n;
  }
  const int r0 = blockIdx.x * p.rows_per_block, r1 = min(p.hw, r0 + p.rows_per_block);
  const long long off = ((long long)b * p.hw) * p.c + m.c0;
  for (int r = r0 + m.lane; r < r1; r += m.lanes) {
    float f[8], d[8];
    unpack8(*reinterpret_cast<const uint4*>(p.x + off + (long long)r * p.c), f);
    unpack8(*reinterpret_cast<const uint4*>(p.dy + off + (long long)r * p.c), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (f[i] - mean[i]) * rstd[i];
      float dz = d[i];
      if (p.silu) dz *= dsilu_f(fmaf(xh, gam[i], bet[i]));
      f[i] = rstd[i] * (dz * gam[i] - m1[i] - xh * m2[i]);
    }
    uint4 o;
    o.x = pack_half2(f[0], f[1]);
    o.y = pack_half2(f[2], f[3]);
    o.z = pack_half2(f[4], f[5]);
    o.w = pack_half2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p.out + off + (long long)r * p.c) = o;
  }
}

static int gn_launch_shape(const sta_groupnorm_args* a, dim3* grid, int* threads, int* rows_per_block) {
  if (a->channels % kGnGroups != 0 || a->channels % 8 != 0)
    return fail(STA_ERR_UNSUPPORTED, "groupnorm: channels %d must be a multiple of 32 and of 8", a->channels);
  const int vecs = a->channels / 8;
  if (vecs > 1024) return fail(STA_ERR_UNSUPPORTED, "groupnorm: channels %d too large", a->channels);
  int lanes = 256 / vecs;
  if (lanes < 1) lanes = 1;
  *threads = vecs * lanes;
  // enough blocks to fill the chip (148 SMs x a few) but at least ~8 rows per lane
  int blocks = (148 * 4 + a->batch - 1) / a->batch;
  int rpb = (a->hw + blocks - 1) / blocks;
  if (rpb < lanes * 4) rpb = lanes * 4;
  *rows_per_block = rpb;
  *grid = dim3((a->hw + rpb - 1) / rpb, a->batch);
  return STA_OK;
}

}  // namespace sta

extern "C" int sta_groupnorm_fwd(const sta_groupnorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x || !a->out || !a->gamma || !a->beta || !a->stats) return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: null pointer");
  if (a->batch < 1 || a->hw < 1) return fail(STA_ERR_BAD_ARG, "sta_groupnorm_fwd: empty shape");
  dim3 grid;
  int threads, rpb;
  int rc = gn_launch_shape(a, &grid, &threads, &rpb);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.out = reinterpret_cast<__half*>(a->out);
  p.gamma = a->gamma; p.beta = a->beta; p.stats = a->stats;
  p.batch = a->batch; p.hw = a->hw; p.c = a->channels; p.rows_per_block = rpb; p.silu = a->silu; p.eps = a->eps;
  STA_CUDA_CHECK(cudaMemsetAsync(a->stats, 0, sizeof(float) * a->batch * kGnGroups * 2, s));
  gn_stats_kernel<<<grid, threads, 0, s>>>(p);
  gn_apply_kernel<<<grid, threads, 0, s>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_groupnorm_bwd(const sta_groupnorm_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->x || !a->d_out || !a->out || !a->gamma || !a->beta || !a->stats || !a->bwd_stats)
    return fail(STA_ERR_BAD_ARG, "sta_groupnorm_bwd: null pointer");
  dim3 grid;
  int threads, rpb;
  int rc = gn_launch_shape(a, &grid, &threads, &rpb);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GnParams p{};
  p.x = reinterpret_cast<const __half*>(a->x);
  p.dy = reinterpret_cast<const __half*>(a->d_out);
  p.out = reinterpret_cast<__half*>(a->out);
  p.gamma = a->gamma; p.beta = a->beta; p.stats = a->stats; p.bstats = a->bwd_stats;
  p.batch = a->batch; p.hw = a->hw; p.c = a->channels; p.rows_per_block = rpb; p.silu = a->silu; p.eps = a->eps;
  STA_CUDA_CHECK(cudaMemsetAsync(a->bwd_stats, 0, sizeof(float) * a->batch * kGnGroups * 2, s));
  gn_bwd_stats_kernel<<<grid, threads, 0, s>>>(p);
  gn_bwd_apply_kernel<<<grid, threads, 0, s>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}
