// sta_sattn_bwd.cu — flash self-attention backward for sm_100a: d(out) -> d(q), d(k), d(v).
//
// Backward of attn1 (reference ldm/modules/attention.py:175-197 under autograd; the alpha optimisation
// back-propagates through every self-attention of every UNet evaluation, ldm/models/diffusion/plms.py:276).
//
// Three launches:
//   1. sattn_delta_kernel     delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]            (fp32)
//   2. sattn_bwd_kernel       CTA = one 128-key tile j of one (batch, head), loop over query tiles i:
//          S^T  = K_j Q_i^T,  dP^T = V_j dO_i^T                         (tcgen05 SS, accumulators in TMEM)
//          P^T  = exp2(S^T*scale*log2e - lse_i),  dS^T = scale * P^T o (dP^T - delta_i)   (one thread per key)
//          dV  += P^T dO_i        (TS: P^T packed fp16 in TMEM over S^T;  dO_i is an MN-major smem operand)
//          dK  += dS^T Q_i        (SS: dS^T written by the threads into 128B-swizzled smem, K-major A)
//          dQ_i = dS K_j          (SS: the same smem tile read as an M-major A operand) -> fp32 red.add
//   3. sattn_dq_cast_kernel   dq_accum fp32 -> d_q fp16
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

constexpr int kSBBlockBytes = 128 * 128;

template <int D>
struct SattnBwdCfg {
  static constexpr int DMMA = (D + 15) / 16 * 16;
  static constexpr int NBLK = (D + 63) / 64;
  static constexpr int TILE = NBLK * kSBBlockBytes;
  static constexpr int ST = (NBLK == 3) ? 1 : 2;  // (Q_i, dO_i) ring depth
  // TMEM: S^T [0,128) dP^T [128,256) dV, dK, dQ accumulators of NACC columns each.  For D = 160 three 160-column
  // accumulators do not fit: the head dim is processed in NPASS = 2 column passes split at a 64-column smem
  // block boundary ([0,128) then [128,160)), S^T / dP^T are recomputed per pass, and the dQ accumulator shares
  // columns with S^T / dP^T (ALIAS_DQ).  These layers have N <= 576 tokens, so the extra work is negligible.
  static constexpr int NPASS = (DMMA > 128) ? 2 : 1;
  static constexpr int NACC_MAX = (NPASS == 1) ? DMMA : 128;
  static constexpr bool ALIAS_DQ = NPASS > 1;
  static constexpr int TMEM_DV = 256;
  static constexpr int TMEM_DK = 256 + NACC_MAX;
  static constexpr int TMEM_DQ = ALIAS_DQ ? 0 : 256 + 2 * NACC_MAX;
  static constexpr int DS_BYTES = 2 * kSBBlockBytes;  // dS^T: 128 keys x 128 queries fp16
  static constexpr int SMEM_BYTES = 2 * TILE + 2 * ST * TILE + DS_BYTES + 1024;
  static constexpr int THREADS = 192;
};

struct SattnBwdParams {
  const float* lse;    // [b, h, n]
  const float* delta;  // [b, h, n]
  float* dq_accum;     // [b, n, h*D] fp32
  __half* d_k;
  __half* d_v;         // [b, n, h*D] fp16 contiguous
  int n, heads;
  float scale, scale_log2;
  unsigned int* err;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int D>
__global__ void __launch_bounds__(SattnBwdCfg<D>::THREADS, 1)
sattn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                 const SattnBwdParams p) {
  using Cfg = SattnBwdCfg<D>;
  constexpr int ST = Cfg::ST, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sK = smem;
  unsigned char* sV = sK + Cfg::TILE;
  unsigned char* sQ = sV + Cfg::TILE;          // ring: stage s -> Q at sQ + s*2*TILE, dO right after it
  unsigned char* sDS = sQ + 2 * ST * Cfg::TILE;

  __shared__ uint64_t kv_full, qdo_full[ST], qdo_empty[ST], sdp_full, pds_ready, dq_full, dq_drained;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  __shared__ float s_lse2[128], s_delta[128];  // lse*log2e and delta of the current query tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n = p.n;
  const int T = (n + 127) / 128;

  if (tid == 0) {
    dead = 0;
    mbar_init(&kv_full, 1);
    for (int i = 0; i < ST; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(&sdp_full, 1);
    mbar_init(&pds_ready, 4);
    mbar_init(&dq_full, 1);
    mbar_init(&dq_drained, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(&kv_full, 2 * Cfg::TILE);
      for (int blk = 0; blk < NBLK; ++blk) {
        tma_load_4d(sK + blk * kSBBlockBytes, &tm_k, &kv_full, blk * 64, h, j * 128, b);
        tma_load_4d(sV + blk * kSBBlockBytes, &tm_v, &kv_full, blk * 64, h, j * 128, b);
      }
      for (int it = 0; it < Cfg::NPASS * T; ++it) {
        const int st = it % ST, i = it % T;
        if (!mbar_wait(&qdo_empty[st], ((it / ST) & 1) ^ 1, &dead, p.err, 10)) break;
        mbar_expect_tx(&qdo_full[st], 2 * Cfg::TILE);
        unsigned char* dq = sQ + st * 2 * Cfg::TILE;
        for (int blk = 0; blk < NBLK; ++blk) {
          tma_load_4d(dq + blk * kSBBlockBytes, &tm_q, &qdo_full[st], blk * 64, h, i * 128, b);
          tma_load_4d(dq + Cfg::TILE + blk * kSBBlockBytes, &tm_do, &qdo_full[st], blk * 64, h, i * 128, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);              // K-major, 128B swizzle
      constexpr uint64_t mndesc_hi = umma_desc_hi_sw128(kSBBlockBytes, 1024);  // MN-major, 128-row blocks
      constexpr uint32_t idesc_nt = umma_idesc_f16(128, 128, 0, 0);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), ds_addr = smem_u32(sDS);

      bool ok = mbar_wait(&kv_full, 0, &dead, p.err, 20);
      for (int it = 0; it < Cfg::NPASS * T && ok; ++it) {
        const int st = it % ST, pass = it / T, i = it % T;
        const int nacc = (Cfg::NPASS == 1) ? DMMA : (pass == 0 ? 128 : DMMA - 128);
        const uint32_t col_off = pass * 2 * kSBBlockBytes;               // first 64-column block of this pass
        const uint32_t idesc_acc = umma_idesc_f16(128, nacc, 0, 1);      // A K-major (TMEM or smem), B MN-major
        const uint32_t idesc_dq = umma_idesc_f16(128, nacc, 1, 1);       // A M-major (dS^T read transposed)
        const uint32_t q_addr = smem_u32(sQ + st * 2 * Cfg::TILE), do_addr = q_addr + Cfg::TILE;
        ok = mbar_wait(&qdo_full[st], (it / ST) & 1, &dead, p.err, 21);
        if (ok && Cfg::ALIAS_DQ && it > 0) ok = mbar_wait(&dq_drained, (it - 1) & 1, &dead, p.err, 22);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t off = (k / 4) * kSBBlockBytes + (k % 4) * 32;
          umma_ss(tmem, umma_desc(kdesc_hi, k_addr + off), umma_desc(kdesc_hi, q_addr + off), idesc_nt, k > 0);
        }
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t off = (k / 4) * kSBBlockBytes + (k % 4) * 32;
          umma_ss(tmem + 128, umma_desc(kdesc_hi, v_addr + off), umma_desc(kdesc_hi, do_addr + off), idesc_nt, k > 0);
        }
        umma_commit(&sdp_full);
        ok = mbar_wait(&pds_ready, it & 1, &dead, p.err, 23);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV += P^T dO_i
          umma_ts(tmem + Cfg::TMEM_DV, tmem + k * 8, umma_desc(mndesc_hi, do_addr + col_off + k * 2048), idesc_acc,
                  i > 0 || k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK += dS^T Q_i
          umma_ss(tmem + Cfg::TMEM_DK, umma_desc(kdesc_hi, ds_addr + (k / 4) * kSBBlockBytes + (k % 4) * 32),
                  umma_desc(mndesc_hi, q_addr + col_off + k * 2048), idesc_acc, i > 0 || k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ_i = dS K_j
          umma_ss(tmem + Cfg::TMEM_DQ, umma_desc(mndesc_hi, ds_addr + k * 2048),
                  umma_desc(mndesc_hi, k_addr + col_off + k * 2048), idesc_dq, k > 0);
        umma_commit(&dq_full);
        umma_commit(&qdo_empty[st]);
      }
    }
  } else {
    // ===================================== per-key-row math ==================================
    const int r = ((warp & 3) << 5) + lane;  // row inside the tile: key row for S^T/dP^T, query row for dQ_i
    const int t128 = tid - 64;               // 0..127 among the four math warps
    const int key = j * 128 + r;
    const bool key_ok = key < n;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const float* lse_bh = p.lse + ((long long)b * p.heads + h) * n;
    const float* delta_bh = p.delta + ((long long)b * p.heads + h) * n;
    bool ok = true;
    for (int it = 0; it < Cfg::NPASS * T; ++it) {
      const int pass = it / T, i = it % T;
      const int col0 = pass * 128;                                             // first head-dim column of this pass
      const int ncols = (Cfg::NPASS == 1) ? D : (pass == 0 ? 128 : D - 128);   // columns of this pass
      // stage lse/delta of this query tile for broadcast reads
      {
        const int qi = i * 128 + t128;
        s_lse2[t128] = qi < n ? lse_bh[qi] * 1.4426950408889634f : INFINITY;
        s_delta[t128] = qi < n ? delta_bh[qi] : 0.f;
      }
      named_bar_sync(1, 128);
      ok = mbar_wait_warp(&sdp_full, it & 1, &dead, p.err, 30);
      if (!ok) break;
      tc_fence_after();
      const float* lse2 = s_lse2;
      const float* dlt = s_delta;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t s[32], dp[32];
        tmem_ld32(lane_addr + c0, s);
        tmem_ld32(lane_addr + 128 + c0, dp);
        tmem_ld_wait();
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(s[q]), p.scale_log2, -lse2[c0 + q]));
          float p1 = fast_exp2(fmaf(__uint_as_float(s[q + 1]), p.scale_log2, -lse2[c0 + q + 1]));
          if (!key_ok) { p0 = 0.f; p1 = 0.f; }
          const float d0 = p.scale * p0 * (__uint_as_float(dp[q]) - dlt[c0 + q]);
          const float d1 = p.scale * p1 * (__uint_as_float(dp[q + 1]) - dlt[c0 + q + 1]);
          pk[q >> 1] = pack_half2(p0, p1);
          dk[q >> 1] = pack_half2(d0, d1);
        }
        tmem_st16(lane_addr + (c0 >> 1), pk);
        // dS^T row r, query columns [c0, c0+32): four 16-byte chunks of block c0/64
        unsigned char* blk = sDS + (c0 >> 6) * kSBBlockBytes;
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint4 v = make_uint4(dk[4 * cc], dk[4 * cc + 1], dk[4 * cc + 2], dk[4 * cc + 3]);
          *reinterpret_cast<uint4*>(blk + sw128_offset(r, chunk0 + cc)) = v;
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_ready);
      named_bar_sync(2, 128);  // everyone is done reading s_lse2 / s_delta before the next tile overwrites them
      // ---- drain dQ_i into the fp32 accumulator ----
      ok = mbar_wait_warp(&dq_full, it & 1, &dead, p.err, 31);
      if (!ok) break;
      tc_fence_after();
      {
        const int qrow = i * 128 + r;
        float* dst = p.dq_accum + ((long long)b * n + qrow) * (p.heads * D) + h * D + col0;
#pragma unroll
        for (int c0 = 0; c0 < Cfg::NACC_MAX; c0 += 8) {
          if (c0 >= ncols) break;
          uint32_t o[8];
          tmem_ld8(lane_addr + Cfg::TMEM_DQ + c0, o);
          tmem_ld_wait();
          if (qrow < n) {
            red_add_v4(dst + c0, __uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]));
            red_add_v4(dst + c0 + 4, __uint_as_float(o[4]), __uint_as_float(o[5]), __uint_as_float(o[6]), __uint_as_float(o[7]));
          }
        }
      }
      if (i == T - 1) {
        // ---- end of a pass: dK, dV columns [col0, col0 + ncols) of this key row ----
        __half* dk_row = p.d_k + ((long long)b * n + key) * (p.heads * D) + h * D + col0;
        __half* dv_row = p.d_v + ((long long)b * n + key) * (p.heads * D) + h * D + col0;
#pragma unroll
        for (int c0 = 0; c0 < Cfg::NACC_MAX; c0 += 8) {
          if (c0 >= ncols) break;
          uint32_t a[8], c[8];
        tmem_ld8(lane_addr + Cfg::TMEM_DK + c0, a);
        tmem_ld8(lane_addr + Cfg::TMEM_DV + c0, c);
        tmem_ld_wait();
        if (key_ok) {
          uint4 v;
          v.x = pack_half2(__uint_as_float(a[0]), __uint_as_float(a[1]));
          v.y = pack_half2(__uint_as_float(a[2]), __uint_as_float(a[3]));
          v.z = pack_half2(__uint_as_float(a[4]), __uint_as_float(a[5]));
          v.w = pack_half2(__uint_as_float(a[6]), __uint_as_float(a[7]));
          *reinterpret_cast<uint4*>(dk_row + c0) = v;
          v.x = pack_half2(__uint_as_float(c[0]), __uint_as_float(c[1]));
          v.y = pack_half2(__uint_as_float(c[2]), __uint_as_float(c[3]));
          v.z = pack_half2(__uint_as_float(c[4]), __uint_as_float(c[5]));
          v.w = pack_half2(__uint_as_float(c[6]), __uint_as_float(c[7]));
            *reinterpret_cast<uint4*>(dv_row + c0) = v;
          }
        }
      }
      if (Cfg::ALIAS_DQ) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_drained);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// delta[b,h,i] = <dO[b,i,h,:], O[b,i,h,:]>; one thread per (b, i, h)
template <int D>
__global__ void sattn_delta_kernel(const __half* __restrict__ o, const __half* __restrict__ d_o, float* __restrict__ delta,
                                   int batch, int n, int heads, long long o_ts, long long o_bs, long long do_ts,
                                   long long do_bs) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * n * heads;
  if (idx >= total) return;
  const int h = idx % heads;
  const long long bi = idx / heads;
  const int i = bi % n, b = bi / n;
  const uint4* po = reinterpret_cast<const uint4*>(o + b * o_bs + i * o_ts + h * D);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * do_bs + i * do_ts + h * D);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < D / 8; ++c) {
    const uint4 a = po[c], g = pd[c];
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = __half22float2(ah[t]), y = __half22float2(gh[t]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  delta[((long long)b * heads + h) * n + i] = acc;
}

__global__ void sattn_dq_cast_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  const float4 v = src[idx];
  uint2 o;
  o.x = pack_half2(v.x, v.y);
  o.y = pack_half2(v.z, v.w);
  dst[idx] = o;
}

template <int D>
static int launch_sattn_bwd(const sta_sattn_bwd_args* a, cudaStream_t stream) {
  using Cfg = SattnBwdCfg<D>;
  CUtensorMap tm_q, tm_k, tm_v, tm_do;
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->n, (uint64_t)a->batch};
  const uint32_t box[4] = {64, 1, 128, 1};
  auto mk = [&](CUtensorMap* m, const void* ptr, long long ts, long long bs) {
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)ts * 2, (uint64_t)bs * 2};
    return make_tmap_f16(m, ptr, 4, dims, st, box);
  };
  int rc;
  if ((rc = mk(&tm_q, a->q, a->q_token_stride, a->q_batch_stride))) return rc;
  if ((rc = mk(&tm_k, a->k, a->k_token_stride, a->k_batch_stride))) return rc;
  if ((rc = mk(&tm_v, a->v, a->v_token_stride, a->v_batch_stride))) return rc;
  if ((rc = mk(&tm_do, a->d_out, a->do_token_stride, a->do_batch_stride))) return rc;

  const long long C = (long long)a->heads * D;
  const long long total = (long long)a->batch * a->n * a->heads;
  STA_CUDA_CHECK(cudaMemsetAsync(a->dq_accum, 0, sizeof(float) * a->batch * a->n * C, stream));
  sattn_delta_kernel<D><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(a->out), reinterpret_cast<const __half*>(a->d_out), a->delta, a->batch, a->n,
      a->heads, a->o_token_stride, a->o_batch_stride, a->do_token_stride, a->do_batch_stride);
  STA_CUDA_CHECK(cudaGetLastError());

  SattnBwdParams p;
  p.lse = a->lse;
  p.delta = a->delta;
  p.dq_accum = a->dq_accum;
  p.d_k = reinterpret_cast<__half*>(a->d_k);
  p.d_v = reinterpret_cast<__half*>(a->d_v);
  p.n = a->n;
  p.heads = a->heads;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static bool attr_set = false;
  if (!attr_set) {
    STA_CUDA_CHECK(cudaFuncSetAttribute(sattn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a->n + 127) / 128, a->heads, a->batch);
  sattn_bwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_q, tm_k, tm_v, tm_do, p);
  STA_CUDA_CHECK(cudaGetLastError());

  const long long n4 = (long long)a->batch * a->n * C / 4;
  sattn_dq_cast_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(a->dq_accum),
                                                                        reinterpret_cast<uint2*>(a->d_q), n4);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta

extern "C" int sta_sattn_bwd(const sta_sattn_bwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->q || !a->k || !a->v || !a->out || !a->d_out || !a->lse || !a->d_q || !a->d_k || !a->d_v ||
      !a->dq_accum || !a->delta)
    return fail(STA_ERR_BAD_ARG, "sta_sattn_bwd: null pointer");
  if (a->batch < 1 || a->n < 1 || a->heads < 1) return fail(STA_ERR_BAD_ARG, "sta_sattn_bwd: empty shape");
  if ((a->o_token_stride % 8) || (a->o_batch_stride % 8) || (a->do_token_stride % 8) || (a->do_batch_stride % 8))
    return fail(STA_ERR_UNSUPPORTED, "sta_sattn_bwd: out/d_out rows must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (a->head_dim) {
    case 40: return launch_sattn_bwd<40>(a, s);
    case 80: return launch_sattn_bwd<80>(a, s);
    case 160: return launch_sattn_bwd<160>(a, s);
    default:
      return fail(STA_ERR_UNSUPPORTED, "sta_sattn_bwd: head_dim %d not built (SD-v1 uses 40/80/160)", a->head_dim);
  }
}
