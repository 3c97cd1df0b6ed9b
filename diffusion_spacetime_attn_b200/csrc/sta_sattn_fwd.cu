// sta_sattn_fwd.cu — flash self-attention forward for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces CrossAttention.forward with context=None (attn1), reference
// ldm/modules/attention.py:175-197: sim = q k^T * scale; softmax(dim=-1); out = attn v — without ever
// writing the [heads, N, N] score tensor to HBM.
//
// CTA = NQ query tiles of 128 rows of one (batch, head); 2 + 4*NQ warps:
//   warp 0        TMA producer: Q tiles once, then K / V tiles of 128 keys through two mbarrier rings
//   warp 1        MMA issuer (one lane): S_r = Q_r K_j^T  (SS, K-major operands)  and  O_r += P_r V_j
//                 (TS: P_r is read from TMEM as packed fp16, V_j is an MN-major smem operand)
//   warps 2..5    softmax of query tile 0: one thread per row, online softmax with lazy rescale
//   warps 6..9    softmax of query tile 1 (NQ = 2)
// TMEM columns: S_0 [0,128) S_1 [128,256) (P_r aliases the first 64 columns of S_r), O_r at 256 + r*DMMA.
// The two query tiles ping-pong on the tensor pipe: while the softmax warps of tile 0 work on S_0 the
// tensor core computes S_1 and O_1 += P_1 V, and vice versa.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

// softmax threads per query row.  Measured on B200 (L0 / L1 / N=9216, us): 1: 111.6 / 22.5 / 434.9   2: 110.6 / 21.4 / 432.1
// — no gain from four instead of two softmax warps per scheduler, so the simpler layout stays the default.
#ifndef STA_FWD_SPLIT
#define STA_FWD_SPLIT 1
#endif
#ifndef STA_POLY_EVERY
#define STA_POLY_EVERY 0  // k > 0: 1 pair in k on the FMA pipe.  Measured on B200 (L0, us): 0: 111.6, 6: 115.7, 4: 117.8, 3: 130.0, 2: 150.7 -> off
#endif

namespace sta {

constexpr int kBlockBytes = 128 * 128;  // one 64-column block of a 128-row tile

template <int D>
struct SattnCfg {
  static constexpr int DMMA = (D + 15) / 16 * 16;   // K of the QK^T MMA and N of the PV MMA
  static constexpr int NBLK = (D + 63) / 64;        // 64-column shared-memory blocks per tile
  static constexpr int NQ = (DMMA <= 128) ? 2 : 1;  // query tiles per CTA (TMEM: 256 + NQ*DMMA <= 512)
  static constexpr int KS = (NBLK == 1) ? 3 : 2;    // K ring depth
  static constexpr int VS = (NBLK == 1) ? 3 : (NBLK == 2 ? 2 : 1);
  static constexpr int TILE_BYTES = NBLK * kBlockBytes;
  static constexpr int SMEM_BYTES = (NQ + KS + VS) * TILE_BYTES + 1024;
  static constexpr int SW = STA_FWD_SPLIT;          // softmax threads per query row (1 or 2)
  static constexpr int THREADS = 64 + 128 * NQ * SW;
  // TMEM columns: S_r at 128 r.  When they fit, the packed-fp16 P_r tiles get their OWN 64 columns (SEP_P) instead
  // of aliasing S_r: the next S_r = Q_r K_{j+1}^T can then be issued as soon as the softmax warps have READ S_r,
  // i.e. it overlaps the exp phase instead of following the P V MMAs.  d=40: 256 + 128 + 96 = 480, d=160: 128 + 64 +
  // 160 = 352 columns; d=80 (256 + 128 + 160 = 544) keeps the aliasing.
  static constexpr bool SEP_P = (128 * NQ + 64 * NQ + NQ * DMMA) <= 512;
  static constexpr int TMEM_P = SEP_P ? 128 * NQ : 0;
  static constexpr int P_STRIDE = SEP_P ? 64 : 128;
  static constexpr int TMEM_O = 128 * NQ + (SEP_P ? 64 * NQ : 0);
};

struct SattnFwdParams {
  __half* out;
  float* lse;
  int n, heads;
  long long o_token_stride, o_batch_stride;
  float scale_log2;  // scale * log2(e)
  unsigned int* err;
  int q_per_cta;  // query tiles handled by one CTA: NQ, or 1 when NQ tiles per CTA would leave SMs idle (small layers)
};

template <int D>
__global__ void __launch_bounds__(SattnCfg<D>::THREADS, 1)
sattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const SattnFwdParams p) {
  using Cfg = SattnCfg<D>;
  constexpr int NQ = Cfg::NQ, KS = Cfg::KS, VS = Cfg::VS, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + NQ * Cfg::TILE_BYTES;
  unsigned char* sV = sK + KS * Cfg::TILE_BYTES;

  __shared__ uint64_t q_full, k_full[KS], k_empty[KS], v_full[VS], v_empty[VS];
  __shared__ uint64_t s_full[NQ], s_consumed[NQ], p_ready[NQ], pv_done[NQ];
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  __shared__ float xmax[2][NQ][2][128], xsum[NQ][2][128];  // row statistics exchanged between the two halves of a row (SW = 2)

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int q0 = blockIdx.x * (128 * p.q_per_cta), h = blockIdx.y, b = blockIdx.z;
  const int n = p.n;
  const int T = (n + 127) / 128;                           // KV tiles
  // second query tile: only with two tiles per CTA, and it may be past the end of the sequence
  const int nq_active = (NQ == 2 && p.q_per_cta == 2 && q0 + 128 < n) ? 2 : 1;

  if (tid == 0) {
    dead = 0;
    mbar_init(&q_full, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < VS; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_consumed[i], 4 * Cfg::SW); mbar_init(&p_ready[i], 4 * Cfg::SW);
      mbar_init(&pv_done[i], 1);
    }
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
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      mbar_expect_tx_w(&q_full, nq_active * Cfg::TILE_BYTES);
      for (int r = 0; r < nq_active; ++r)
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d_w(sQ + (r * NBLK + blk) * kBlockBytes, &tm_q, &q_full, blk * 64, h, q0 + r * 128, b);
      for (int j = 0; j < T; ++j) {
        const int ks = j % KS, vs = j % VS;
        if (!mbar_wait_warp(&k_empty[ks], ((j / KS) & 1) ^ 1, &dead, p.err, 10)) break;
        mbar_expect_tx_w(&k_full[ks], Cfg::TILE_BYTES);
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d_w(sK + (ks * NBLK + blk) * kBlockBytes, &tm_k, &k_full[ks], blk * 64, h, j * 128, b);
        if (!mbar_wait_warp(&v_empty[vs], ((j / VS) & 1) ^ 1, &dead, p.err, 11)) break;
        mbar_expect_tx_w(&v_full[vs], Cfg::TILE_BYTES);
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d_w(sV + (vs * NBLK + blk) * kBlockBytes, &tm_v, &v_full[vs], blk * 64, h, j * 128, b);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // ONE issue warp serves both query tiles (measured: a second issue warp makes the kernel ~10 % slower).  Whole
    // warp, warp-uniform control flow; the single-lane instructions are elected inside the *_w helpers.
    {
      constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);            // K-major operands
      constexpr uint64_t vdesc_hi = umma_desc_hi_sw128(kBlockBytes, 1024);   // MN-major V: LBO = block stride
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, DMMA, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);

      auto issue_qk = [&](int r, int ks) {
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t off = (k / 4) * kBlockBytes + (k % 4) * 32;
          umma_ss_w(tmem + r * 128, umma_desc(kdesc_hi, q_addr + r * Cfg::TILE_BYTES + off),
                  umma_desc(kdesc_hi, k_addr + ks * Cfg::TILE_BYTES + off), idesc_qk, k > 0);
        }
        umma_commit_w(&s_full[r]);
      };
      auto issue_pv = [&](int r, int vs, bool acc) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts_w(tmem + Cfg::TMEM_O + r * DMMA, tmem + Cfg::TMEM_P + r * Cfg::P_STRIDE + k * 8,
                  umma_desc(vdesc_hi, v_addr + vs * Cfg::TILE_BYTES + k * 2048), idesc_pv, acc || k > 0);
        umma_commit_w(&pv_done[r]);
      };
      // warp-uniform "has this phase completed?" (try_wait: a negative answer suspends the warp for a while, which
      // keeps this polling loop off the issue slots of the softmax warp sharing the scheduler)
      auto ready = [&](uint64_t* bar, uint32_t parity) { return __all_sync(0xffffffffu, mbar_try_wait(bar, parity)); };

      bool ok = mbar_wait_warp(&q_full, 0, &dead, p.err, 20) && mbar_wait_warp(&k_full[0], 0, &dead, p.err, 21);
      if (ok) {
        tc_fence_after();
        // Event-driven issue: each query tile r alternates  S_r(j+1) = Q_r K_{j+1}^T  (allowed once the softmax warps
        // have read S_r(j) [SEP_P] / once P_r(j) V_j has been issued [aliased P])  and  O_r += P_r(j) V_j  (once
        // P_r(j) is in TMEM).  Whichever tile has its next operand ready is served first.
        int qk_next[2] = {1, 1}, pv_next[2] = {0, 0};
        int k_rel = 0, v_rel = 0;  // K / V tiles released back to the TMA producer so far
        for (int r = 0; r < nq_active; ++r) issue_qk(r, 0);
        if (nq_active == 1) { qk_next[1] = T; pv_next[1] = T; }
        long long idle = 0;
        while (pv_next[0] < T || pv_next[1] < T) {
          bool progress = false;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (r >= nq_active) continue;
            // service order matters because a failed try_wait suspends the warp for a while: with separate P
            // buffers the latency-critical event is "S_r read" (-> next S_r), with aliased P it is "P_r ready"
            auto try_qk = [&]() {
              const int jq = qk_next[r];
              if (jq < T && (Cfg::SEP_P ? ready(&s_consumed[r], (jq - 1) & 1) : (pv_next[r] >= jq)) &&
                  ready(&k_full[jq % KS], (jq / KS) & 1)) {
                tc_fence_after();
                issue_qk(r, jq % KS);
                qk_next[r] = jq + 1;
                progress = true;
              }
            };
            auto try_pv = [&]() {
              const int jp = pv_next[r];
              if (jp < T && ready(&p_ready[r], jp & 1) && ready(&v_full[jp % VS], (jp / VS) & 1)) {
                tc_fence_after();
                issue_pv(r, jp % VS, jp > 0);
                pv_next[r] = jp + 1;
                progress = true;
              }
            };
            if (Cfg::SEP_P) { try_qk(); try_pv(); } else { try_pv(); try_qk(); }
          }
          // a K (V) tile goes back to the producer once every active query tile has issued its MMA on it
          const int k_done = min(qk_next[0], qk_next[1]), v_done = min(pv_next[0], pv_next[1]);
          while (k_rel < min(k_done, T)) { umma_commit_w(&k_empty[k_rel % KS]); ++k_rel; }
          while (v_rel < min(v_done, T)) { umma_commit_w(&v_empty[v_rel % VS]); ++v_rel; }
          if (progress) {
            idle = 0;
          } else if (++idle > (1ll << 22) || dead) {  // seconds of polling without progress: give up loudly
            if (!dead) { dead = 1; atomicCAS(p.err, 0u, STA_ERR_MBAR_TIMEOUT | 22u); }
            break;
          }
        }
      }
    }
  } else {
    // ===================================== softmax / epilogue ================================
    // SW threads per query row (SW = 2: the 128 score columns of a row are split between two warpgroups, which share the
    // running reference maximum through shared memory — one 256-thread named barrier per key tile — and add their partial
    // row sums once, in the epilogue).  Built to test whether four softmax warps per scheduler hide each other's TMEM-load /
    // MUFU / pack latencies better than two: they do not (see STA_FWD_SPLIT above), the default is one thread per row.
    constexpr int SW = Cfg::SW, CW = 128 / SW, NCH = DMMA / 8;
    const int sw = warp - 2;
    const int r = sw / (4 * SW);      // query tile of this warp
    const int half = (sw >> 2) % SW;  // score columns [CW * half, CW * half + CW) of every key tile
    if (r < nq_active) {
      const int row_in_tile = ((warp & 3) << 5) + lane;
      const int row = q0 + r * 128 + row_in_tile;
      const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
      const uint32_t s_addr = lane_addr + r * 128 + half * CW;
      const uint32_t p_addr = lane_addr + Cfg::TMEM_P + r * Cfg::P_STRIDE + half * (CW / 2);
      const uint32_t o_addr = lane_addr + Cfg::TMEM_O + r * DMMA;
      const int oc0 = half * NCH / SW, oc1 = (half + 1) * NCH / SW;  // 8-column chunks of O owned by this thread
      float m_ref = -INFINITY, l = 0.f;
      bool ok = true;
      for (int j = 0; j < T; ++j) {
        ok = mbar_wait_warp(&s_full[r], j & 1, &dead, p.err, 30);
        if (!ok) break;
        tc_fence_after();
        uint32_t s[CW];
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 32) tmem_ld32(s_addr + c0, s + c0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_consumed[r]);  // S_r is in registers: the tensor core may overwrite it
        const int valid = n - j * 128 - half * CW;  // keys of this slice that exist
        if (valid < CW) {
#pragma unroll
          for (int c = 0; c < CW; ++c)
            if (c >= valid) s[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < CW; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
        }
        float mxl = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (SW == 2) {  // row maximum over both halves (double-buffered exchange: one barrier per key tile)
          xmax[j & 1][r][half][row_in_tile] = mxl;
          named_bar_sync(1 + r, 256);
          mxl = fmaxf(mxl, xmax[j & 1][r][half ^ 1][row_in_tile]);
        }
        const float mx = mxl * p.scale_log2;
        if (j == 0) {
          m_ref = mx;
        } else {
          // lazy rescale: keep the old reference unless the new maximum exceeds it by more than 2^8
          const bool need = mx > m_ref + 8.f;
          if (__any_sync(0xffffffffu, need)) {
            // O_r must be quiescent: P_r(j-1) V_{j-1} has to be complete before the accumulator is rescaled
            ok = mbar_wait_warp(&pv_done[r], (j - 1) & 1, &dead, p.err, 32);
            if (!ok) break;
            tc_fence_after();
            const float f = need ? fast_exp2(m_ref - mx) : 1.f;
            if (need) m_ref = mx;
            l *= f;
            for (int ch = oc0; ch < oc1; ++ch) {
              uint32_t o[8];
              tmem_ld8(o_addr + ch * 8, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st8(o_addr + ch * 8, o);
            }
            tmem_st_wait();
          }
        }
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < CW; c += 2) {  // exp in place: the packed pair (c, c+1) lands in s[c/2]
          // every STA_POLY_EVERY-th pair is evaluated on the FMA pipe instead of the MUFU pipe (see poly_exp2)
          const bool poly = (STA_POLY_EVERY > 0) && ((c >> 1) % (STA_POLY_EVERY > 0 ? STA_POLY_EVERY : 1) == (STA_POLY_EVERY - 1));
          const float x0 = fmaf(__uint_as_float(s[c]), p.scale_log2, -m_ref);
          const float x1 = fmaf(__uint_as_float(s[c + 1]), p.scale_log2, -m_ref);
          const float p0 = poly ? poly_exp2(x0) : fast_exp2(x0);
          const float p1 = poly ? poly_exp2(x1) : fast_exp2(x1);
          l0 += p0;
          l1 += p1;
          s[c >> 1] = pack_half2(p0, p1);
        }
        if (j > 0) {  // the P_r buffer is free once P_r(j-1) V_{j-1} has completed (long done by now)
          ok = mbar_wait_warp(&pv_done[r], (j - 1) & 1, &dead, p.err, 33);
          if (!ok) break;
          tc_fence_after();
        }
#pragma unroll
        for (int c0 = 0; c0 < CW / 2; c0 += 16) tmem_st16(p_addr + c0, s + c0);
        l += l0 + l1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[r]);
      }
      // -------- epilogue: O / l -> fp16, natural-log LSE --------
      ok = __all_sync(0xffffffffu, ok);
      if (SW == 2) {  // total row sum = the two halves' partial sums (taken with the same reference maximum)
        xsum[r][half][row_in_tile] = l;
        named_bar_sync(1 + r, 256);
        l += xsum[r][half ^ 1][row_in_tile];
      }
      if (ok) ok = mbar_wait_warp(&pv_done[r], (T - 1) & 1, &dead, p.err, 31);
      if (ok) {
        tc_fence_after();
        const float inv = 1.f / l;
        __half* orow = p.out + (long long)b * p.o_batch_stride + (long long)row * p.o_token_stride + h * D;
        for (int ch = oc0; ch < oc1; ++ch) {
          if (ch * 8 >= D) break;
          uint32_t o[8];
          tmem_ld8(o_addr + ch * 8, o);
          tmem_ld_wait();
          if (row < n) {
            uint4 v;
            v.x = pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
            v.y = pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
            v.z = pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            v.w = pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
            *reinterpret_cast<uint4*>(orow + ch * 8) = v;
          }
        }
        if (p.lse && row < n && half == 0)
          p.lse[((long long)b * p.heads + h) * n + row] = (m_ref + log2f(l)) * 0.6931471805599453f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int D>
static int launch_sattn_fwd(const sta_sattn_fwd_args* a, cudaStream_t stream) {
  using Cfg = SattnCfg<D>;
  CUtensorMap tm_q, tm_k, tm_v;
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->n, (uint64_t)a->batch};
  const uint32_t box[4] = {64, 1, 128, 1};
  {
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->q_token_stride * 2, (uint64_t)a->q_batch_stride * 2};
    int rc = make_tmap_f16(&tm_q, a->q, 4, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->k_token_stride * 2, (uint64_t)a->k_batch_stride * 2};
    int rc = make_tmap_f16(&tm_k, a->k, 4, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->v_token_stride * 2, (uint64_t)a->v_batch_stride * 2};
    int rc = make_tmap_f16(&tm_v, a->v, 4, dims, st, box);
    if (rc) return rc;
  }
  SattnFwdParams p;
  p.out = reinterpret_cast<__half*>(a->out);
  p.lse = a->lse;
  p.n = a->n;
  p.heads = a->heads;
  p.o_token_stride = a->o_token_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static PerDeviceOnce smem_attr;
  int rc_attr = smem_attr.run([] {
    return cudaFuncSetAttribute(sattn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  if (rc_attr) return rc_attr;
  // Two query tiles per CTA (two softmax warpgroups ping-ponging on one tensor pipe) is the efficient shape when the grid
  // fills the GPU; on the small layers (N = 1024: 64 CTAs) one tile per CTA doubles the number of busy SMs instead.
  static std::atomic<int> sm_count{0};  // every GPU of a box has the same SM count: one query per process
  int num_sms = sm_count.load(std::memory_order_relaxed);
  if (num_sms == 0) {
    int dev = 0;
    num_sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    sm_count.store(num_sms, std::memory_order_relaxed);
  }
  const long long ctas_nq = (long long)((a->n + 128 * Cfg::NQ - 1) / (128 * Cfg::NQ)) * a->heads * a->batch;
  p.q_per_cta = (Cfg::NQ == 2 && ctas_nq < num_sms) ? 1 : Cfg::NQ;
  dim3 grid((a->n + 128 * p.q_per_cta - 1) / (128 * p.q_per_cta), a->heads, a->batch);
  sattn_fwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_q, tm_k, tm_v, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta

extern "C" int sta_sattn_fwd(const sta_sattn_fwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->q || !a->k || !a->v || !a->out) return fail(STA_ERR_BAD_ARG, "sta_sattn_fwd: null pointer");
  if (a->batch < 1 || a->n < 1 || a->heads < 1) return fail(STA_ERR_BAD_ARG, "sta_sattn_fwd: empty shape");
  if ((a->o_token_stride % 8) || (a->o_batch_stride % 8) || (reinterpret_cast<uintptr_t>(a->out) & 15))
    return fail(STA_ERR_UNSUPPORTED, "sta_sattn_fwd: out rows must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (a->head_dim) {
    case 40: return launch_sattn_fwd<40>(a, s);
    case 80: return launch_sattn_fwd<80>(a, s);
    case 160: return launch_sattn_fwd<160>(a, s);
    case 512: return launch_sattn_wide_fwd(a, s);  // KL-VAE mid-block AttnBlock (model.py:150-202)
    default:
      return fail(STA_ERR_UNSUPPORTED, "sta_sattn_fwd: head_dim %d not built (SD-v1 uses 40/80/160, its VAE 512)", a->head_dim);
  }
}
