// sta_xattn_bwd.cu — backward of the fused dual cross-attention + alpha-blend: d(out) -> d(q), d(coef).
//
// The contexts are frozen (CLIP text encoder is not trained, reference ldm/models/diffusion/ddpm.py:519-523), so
// no dK/dV is produced; what the alpha optimisation needs (ldm/models/diffusion/plms.py:204-277) is the gradient
// w.r.t. the hidden state (through q) and w.r.t. the per-object blend weights:
//     out_u = A_u,   out_c = A_g + sum_i w_i (A_i - A_u),   w_i = m_i[pix] c_i
//     dA_u = dO_u - sigma dO_c,  dA_g = dO_c,  dA_i = w_i dO_c,      sigma = sum_i w_i
//     dS_x = P_x o (dA_x V_x^T - rowsum(P_x o dA_x V_x^T)),  dQ = scale * sum_x dS_x K_x
//     dc_i = sum_pix m_i <dO_c, A_i - A_u> = sum_pix m_i (delta_i - delta_uc),
//            delta_x = rowsum(P_x o dO_c V_x^T)   (flash-attention's delta term, so A_i is never rebuilt)
//
// CTA = one 128-pixel tile of one (prompt, head); 6 warps: TMA producer, MMA issuer, 4 warps with one thread per
// pixel row.  Tile 0 is the unconditional row (three accumulators: S_u, dO_u V^T, dO_c V^T), tiles 1.. are the
// conditional row's contexts (objects whose mask is empty in this pixel tile are skipped, as in the forward).
// TMEM: four 80-column slots (S/dP pairs A and B, double-buffered across tiles) + the dQ accumulator at 320.
// P is recomputed from the forward's LSE; dS is written back over S as packed fp16 and consumed from TMEM.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

constexpr int kBQBlockBytes = 128 * 128;
constexpr int kBCBlockBytes = 80 * 128;
constexpr int kBMaxObj = 8;

template <int D>
struct XattnBwdCfg {
  static constexpr int DMMA = (D + 15) / 16 * 16;
  static constexpr int NBLK = (D + 63) / 64;
  static constexpr int ST = (NBLK == 3) ? 1 : (NBLK == 2 ? 2 : 3);  // K+V ring depth
  static constexpr bool PREFETCH = ST >= 2;
  static constexpr int QTILE = NBLK * kBQBlockBytes;
  static constexpr int CTILE = NBLK * kBCBlockBytes;
  static constexpr int SMEM_BYTES = 3 * QTILE + 2 * ST * CTILE + 1024;
  static constexpr int THREADS = 192;
  static constexpr int TMEM_DQ = 320;
};

struct XattnBwdParams {
  const uint8_t* mask;
  const float* coef;
  const float* lse;  // [B, heads, 2+n_obj, n]
  __half* d_q;       // [2B, n, heads*D] contiguous
  float* d_coef;     // [B, n_obj]
  int prompts, n, heads, n_obj, ctx_len;
  float scale, scale_log2;
  unsigned int* err;
};

template <int D>
__global__ void __launch_bounds__(XattnBwdCfg<D>::THREADS, 1)
xattn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                 const XattnBwdParams p) {
  using Cfg = XattnBwdCfg<D>;
  constexpr int ST = Cfg::ST, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                  // Q_u, later re-filled with Q_c
  unsigned char* sDOu = sQ + Cfg::QTILE;
  unsigned char* sDOc = sDOu + Cfg::QTILE;
  unsigned char* sK = sDOc + Cfg::QTILE;     // ring of ST stages, each K then V
  unsigned char* sV = sK + ST * Cfg::CTILE;

  __shared__ uint64_t in_full, q2_full, qu_done, kv_full[ST], kv_empty[ST];
  __shared__ uint64_t sdp_full[2], ds_ready, dq_full, dq_drained;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  __shared__ int tile_slot[2 + kBMaxObj];
  __shared__ int n_tiles_s;
  __shared__ float dcoef_s[kBMaxObj];

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, pr = blockIdx.z;
  const int n = p.n, B = p.prompts, n_obj = p.n_obj, n_slots = 2 + p.n_obj;

  if (tid == 0) {
    dead = 0;
    mbar_init(&in_full, 1);
    mbar_init(&q2_full, 1);
    mbar_init(&qu_done, 1);
    for (int i = 0; i < ST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(&sdp_full[0], 1);
    mbar_init(&sdp_full[1], 1);
    mbar_init(&ds_ready, 4);
    mbar_init(&dq_full, 1);
    mbar_init(&dq_drained, 4);
    mbar_fence_init();
  }
  if (tid < kBMaxObj) dcoef_s[tid] = 0.f;
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
      tile_slot[0] = 0;
      tile_slot[1] = 1;
    }
    int cnt = 2;
    for (int i = 0; i < n_obj; ++i) {
      const uint8_t* m = p.mask + ((long long)pr * n_obj + i) * n + q0;
      int any = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int px = lane * 4 + j;
        if (q0 + px < n) any |= m[px];
      }
      if (__any_sync(0xffffffffu, any != 0)) {
        if (lane == 0) tile_slot[cnt] = 2 + i;
        ++cnt;
      }
    }
    if (lane == 0) n_tiles_s = cnt;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int T = n_tiles_s;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      mbar_expect_tx_w(&in_full, 3 * Cfg::QTILE);
      for (int blk = 0; blk < NBLK; ++blk) {
        tma_load_4d_w(sQ + blk * kBQBlockBytes, &tm_q, &in_full, blk * 64, h, q0, pr);
        tma_load_4d_w(sDOu + blk * kBQBlockBytes, &tm_do, &in_full, blk * 64, h, q0, pr);
        tma_load_4d_w(sDOc + blk * kBQBlockBytes, &tm_do, &in_full, blk * 64, h, q0, pr + B);
      }
      bool ok = true;
      for (int t = 0; t < T && ok; ++t) {
        const int st = t % ST, slot = pr * n_slots + tile_slot[t];
        ok = mbar_wait_warp(&kv_empty[st], ((t / ST) & 1) ^ 1, &dead, p.err, 10);
        if (!ok) break;
        mbar_expect_tx_w(&kv_full[st], 2 * Cfg::CTILE);
        for (int blk = 0; blk < NBLK; ++blk) {
          tma_load_4d_w(sK + (st * NBLK + blk) * kBCBlockBytes, &tm_k, &kv_full[st], blk * 64, h, 0, slot);
          tma_load_4d_w(sV + (st * NBLK + blk) * kBCBlockBytes, &tm_v, &kv_full[st], blk * 64, h, 0, slot);
        }
        if (t == 0) {
          // re-fill the Q buffer with the conditional row once S_u = Q_u K_0^T has been computed
          ok = mbar_wait_warp(&qu_done, 0, &dead, p.err, 12);
          if (!ok) break;
          mbar_expect_tx_w(&q2_full, Cfg::QTILE);
          for (int blk = 0; blk < NBLK; ++blk)
            tma_load_4d_w(sQ + blk * kBQBlockBytes, &tm_q, &q2_full, blk * 64, h, q0, pr + B);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);
      constexpr uint64_t mndesc_hi = umma_desc_hi_sw128(kBCBlockBytes, 1024);
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 80, 0, 0);      // [128 x D] x [80 x D]^T
      constexpr uint32_t idesc_dq = umma_idesc_f16(128, DMMA, 0, 1);   // dS[128 x 80] x K[80 x D]
      const uint32_t q_addr = smem_u32(sQ), dou_addr = smem_u32(sDOu), doc_addr = smem_u32(sDOc);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);

      // D[slot] = A[128 x D] (smem, K-major) * Bt[80 x D] (smem, K-major)
      auto issue_nt = [&](int slot, uint32_t a_addr, uint32_t b_addr) {
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t aoff = (k / 4) * kBQBlockBytes + (k % 4) * 32;
          const uint32_t boff = (k / 4) * kBCBlockBytes + (k % 4) * 32;
          umma_ss_w(tmem + slot * 80, umma_desc(kdesc_hi, a_addr + aoff), umma_desc(kdesc_hi, b_addr + boff), idesc_s,
                  k > 0);
        }
      };
      // dQ (+)= dS[slot] (TMEM, packed fp16) * K[80 x D] (smem, MN-major)
      auto issue_dq = [&](int slot, uint32_t kk_addr, bool acc) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
          umma_ts_w(tmem + Cfg::TMEM_DQ, tmem + slot * 80 + k * 8, umma_desc(mndesc_hi, kk_addr + k * 2048), idesc_dq,
                  acc || k > 0);
      };
      auto issue_sdp = [&](int t) {  // conditional-row tile t >= 1 into pair (t & 1)
        const int st = t % ST, s_slot = (t & 1) * 2;
        issue_nt(s_slot, q_addr, k_addr + st * Cfg::CTILE);
        issue_nt(s_slot + 1, doc_addr, v_addr + st * Cfg::CTILE);
        umma_commit_w(&sdp_full[t & 1]);
      };

      bool ok = mbar_wait_warp(&in_full, 0, &dead, p.err, 20) && mbar_wait_warp(&kv_full[0], 0, &dead, p.err, 21);
      if (ok) {
        tc_fence_after();
        issue_nt(0, q_addr, k_addr);
        umma_commit_w(&qu_done);
        issue_nt(1, dou_addr, v_addr);
        issue_nt(2, doc_addr, v_addr);
        umma_commit_w(&sdp_full[0]);
        ok = mbar_wait_warp(&ds_ready, 0, &dead, p.err, 22);
      }
      if (ok) {
        tc_fence_after();
        issue_dq(0, k_addr, false);
        umma_commit_w(&dq_full);
        umma_commit_w(&kv_empty[0]);
        ok = mbar_wait_warp(&q2_full, 0, &dead, p.err, 23) && mbar_wait_warp(&kv_full[1 % ST], (1 / ST) & 1, &dead, p.err, 24);
      }
      if (ok) {
        tc_fence_after();
        issue_sdp(1);
      }
      for (int t = 1; t < T && ok; ++t) {
        const int st = t % ST;
        if (Cfg::PREFETCH && t + 1 < T) {
          ok = mbar_wait_warp(&kv_full[(t + 1) % ST], ((t + 1) / ST) & 1, &dead, p.err, 25);
          if (!ok) break;
          tc_fence_after();
          issue_sdp(t + 1);
        }
        ok = mbar_wait_warp(&ds_ready, t & 1, &dead, p.err, 26);
        if (ok && t == 1) ok = mbar_wait_warp(&dq_drained, 0, &dead, p.err, 27);
        if (!ok) break;
        tc_fence_after();
        issue_dq((t & 1) * 2, k_addr + st * Cfg::CTILE, t > 1);
        if (t == T - 1) umma_commit_w(&dq_full);
        umma_commit_w(&kv_empty[st]);
        if (!Cfg::PREFETCH && t + 1 < T) {
          ok = mbar_wait_warp(&kv_full[(t + 1) % ST], ((t + 1) / ST) & 1, &dead, p.err, 28);
          if (!ok) break;
          tc_fence_after();
          issue_sdp(t + 1);
        }
      }
    }
  } else {
    // ===================================== per-row math ======================================
    const int row = q0 + ((warp & 3) << 5) + lane;
    const bool row_ok = row < n;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const float kLog2e = 1.4426950408889634f;
    const float* lse_row = p.lse + ((long long)pr * p.heads + h) * n_slots * n + row;

    float sigma = 0.f;
    for (int i = 0; i < n_obj; ++i) {
      const float mk = row_ok ? (float)p.mask[((long long)pr * n_obj + i) * n + row] : 0.f;
      sigma += mk * p.coef[pr * n_obj + i];
    }

    auto write_dq = [&](int batch_row) {
      __half* drow = p.d_q + ((long long)batch_row * n + row) * (p.heads * D) + h * D;
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 8) {
        uint32_t o[8];
        tmem_ld8(lane_addr + Cfg::TMEM_DQ + c0, o);
        tmem_ld_wait();
        if (row_ok) {
          uint4 v;
          v.x = pack_half2(__uint_as_float(o[0]), __uint_as_float(o[1]));
          v.y = pack_half2(__uint_as_float(o[2]), __uint_as_float(o[3]));
          v.z = pack_half2(__uint_as_float(o[4]), __uint_as_float(o[5]));
          v.w = pack_half2(__uint_as_float(o[6]), __uint_as_float(o[7]));
          *reinterpret_cast<uint4*>(drow + c0) = v;
        }
      }
    };

    // ---------------- tile 0: unconditional row ----------------
    float delta_uc = 0.f;
    bool ok = mbar_wait_warp(&sdp_full[0], 0, &dead, p.err, 30);
    if (ok) {
      tc_fence_after();
      const float lse2 = row_ok ? lse_row[0] * kLog2e : 0.f;
      float du = 0.f, duc = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t s[16], a[16], c[16];
        tmem_ld16(lane_addr + c0, s);
        tmem_ld16(lane_addr + 80 + c0, a);
        tmem_ld16(lane_addr + 160 + c0, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pp = (c0 + i < p.ctx_len) ? fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lse2)) : 0.f;
          const float dpc = __uint_as_float(c[i]);
          du = fmaf(pp, fmaf(-sigma, dpc, __uint_as_float(a[i])), du);
          duc = fmaf(pp, dpc, duc);
        }
      }
      delta_uc = duc;
#pragma unroll
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t s[16], a[16], c[16], pk[8];
        tmem_ld16(lane_addr + c0, s);
        tmem_ld16(lane_addr + 80 + c0, a);
        tmem_ld16(lane_addr + 160 + c0, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pp = (c0 + i < p.ctx_len) ? fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lse2)) : 0.f;
          const float dp = fmaf(-sigma, __uint_as_float(c[i]), __uint_as_float(a[i]));
          s[i] = __float_as_uint(p.scale * pp * (dp - du));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_half2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
        tmem_st8(lane_addr + (c0 >> 1), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready);
      ok = mbar_wait_warp(&dq_full, 0, &dead, p.err, 31);
    }
    if (ok) {
      tc_fence_after();
      write_dq(pr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_drained);
    }
    // ---------------- tiles 1..T-1: conditional row ----------------
    for (int t = 1; t < T && ok; ++t) {
      const int slot = tile_slot[t];
      const uint32_t s_addr = lane_addr + (t & 1) * 160;
      float w = 1.f, mk = 0.f;
      if (slot >= 2) {
        mk = row_ok ? (float)p.mask[((long long)pr * n_obj + (slot - 2)) * n + row] : 0.f;
        w = mk * p.coef[pr * n_obj + (slot - 2)];
      }
      const float lse2 = row_ok ? lse_row[(long long)slot * n] * kLog2e : 0.f;
      ok = mbar_wait_warp(&sdp_full[t & 1], (t >> 1) & 1, &dead, p.err, 32);
      if (!ok) break;
      tc_fence_after();
      float delta = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t s[16], a[16];
        tmem_ld16(s_addr + c0, s);
        tmem_ld16(s_addr + 80 + c0, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pp = (c0 + i < p.ctx_len) ? fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lse2)) : 0.f;
          delta = fmaf(pp, __uint_as_float(a[i]), delta);
        }
      }
      const float ws = w * p.scale;
#pragma unroll
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t s[16], a[16], pk[8];
        tmem_ld16(s_addr + c0, s);
        tmem_ld16(s_addr + 80 + c0, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pp = (c0 + i < p.ctx_len) ? fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lse2)) : 0.f;
          s[i] = __float_as_uint(ws * pp * (__uint_as_float(a[i]) - delta));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_half2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
        tmem_st8(s_addr + (c0 >> 1), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready);
      if (slot >= 2) {
        float val = mk * (delta - delta_uc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) atomicAdd(&dcoef_s[slot - 2], val);
      }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (ok) ok = mbar_wait_warp(&dq_full, 1, &dead, p.err, 33);
    if (ok) {
      tc_fence_after();
      write_dq(pr + B);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < n_obj && dcoef_s[tid] != 0.f) atomicAdd(&p.d_coef[pr * n_obj + tid], dcoef_s[tid]);
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int D>
static int launch_xattn_bwd(const sta_xattn_bwd_args* a, cudaStream_t stream) {
  using Cfg = XattnBwdCfg<D>;
  CUtensorMap tm_q, tm_do, tm_k, tm_v;
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->n, (uint64_t)a->prompts * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->q_token_stride * 2, (uint64_t)a->q_batch_stride * 2};
    int rc = make_tmap_f16(&tm_q, a->q, 4, dims, st, box);
    if (rc) return rc;
    const uint64_t st2[4] = {2, (uint64_t)D * 2, (uint64_t)a->do_token_stride * 2, (uint64_t)a->do_batch_stride * 2};
    rc = make_tmap_f16(&tm_do, a->d_out, 4, dims, st2, box);
    if (rc) return rc;
  }
  {
    const uint64_t C = (uint64_t)a->heads * D;
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->ctx_len,
                              (uint64_t)a->prompts * (2 + a->n_obj)};
    const uint64_t st[4] = {2, (uint64_t)D * 2, C * 2, C * 2 * (uint64_t)a->ctx_len};
    const uint32_t box[4] = {64, 1, 80, 1};
    int rc = make_tmap_f16(&tm_k, a->k_ctx, 4, dims, st, box);
    if (rc) return rc;
    rc = make_tmap_f16(&tm_v, a->v_ctx, 4, dims, st, box);
    if (rc) return rc;
  }
  XattnBwdParams p;
  p.mask = a->mask;
  p.coef = a->coef;
  p.lse = a->lse;
  p.d_q = reinterpret_cast<__half*>(a->d_q);
  p.d_coef = a->d_coef;
  p.prompts = a->prompts;
  p.n = a->n;
  p.heads = a->heads;
  p.n_obj = a->n_obj;
  p.ctx_len = a->ctx_len;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static bool attr_set = false;
  if (!attr_set) {
    STA_CUDA_CHECK(cudaFuncSetAttribute(xattn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  if (a->n_obj > 0)
    STA_CUDA_CHECK(cudaMemsetAsync(a->d_coef, 0, sizeof(float) * a->prompts * a->n_obj, stream));
  dim3 grid((a->n + 127) / 128, a->heads, a->prompts);
  xattn_bwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_q, tm_do, tm_k, tm_v, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta

extern "C" int sta_xattn_bwd(const sta_xattn_bwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->q || !a->k_ctx || !a->v_ctx || !a->d_out || !a->d_q || !a->lse)
    return fail(STA_ERR_BAD_ARG, "sta_xattn_bwd: null pointer");
  if (a->prompts < 1 || a->n < 1 || a->heads < 1) return fail(STA_ERR_BAD_ARG, "sta_xattn_bwd: empty shape");
  if (a->n_obj < 0 || a->n_obj > kBMaxObj) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: n_obj %d not in [0,%d]", a->n_obj, kBMaxObj);
  if (a->n_obj > 0 && (!a->mask || !a->coef || !a->d_coef)) return fail(STA_ERR_BAD_ARG, "sta_xattn_bwd: mask/coef/d_coef required when n_obj > 0");
  if (a->ctx_len < 1 || a->ctx_len > 80) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: ctx_len %d not in [1,80]", a->ctx_len);
  if (reinterpret_cast<uintptr_t>(a->d_q) & 15) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: d_q must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (a->head_dim) {
    case 40: return launch_xattn_bwd<40>(a, s);
    case 80: return launch_xattn_bwd<80>(a, s);
    case 160: return launch_xattn_bwd<160>(a, s);
    default:
      return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: head_dim %d not built (SD-v1 uses 40/80/160)", a->head_dim);
  }
}
