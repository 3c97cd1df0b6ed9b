// sta_sattn_bwd.cu — flash self-attention backward for sm_100a: d(out) -> d(q), d(k), d(v).
//
// Backward of attn1 (reference ldm/modules/attention.py:175-197 under autograd; the alpha optimisation
// back-propagates through every self-attention of every UNet evaluation, ldm/models/diffusion/plms.py:276).
//
// Three launches:
//   1. sattn_delta_kernel     delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]  (fp32);  dq_accum = 0
//   2. sattn_bwd_kernel       CTA = one 128-key tile j of one (batch, head), loop over query tiles i:
//          S^T  = K_j Q_i^T,  dP^T = V_j dO_i^T                         (tcgen05 SS, accumulators in TMEM)
//          P^T  = exp2(S^T*scale*log2e - lse_i),  dS^T = scale * P^T o (dP^T - delta_i)   (one thread per key)
//          dV  += P^T dO_i        (TS: P^T packed fp16 in TMEM over S^T;  dO_i is an MN-major smem operand)
//          dK  += dS^T Q_i        (SS: dS^T written by the threads into 128B-swizzled smem, K-major A)
//          dQ_i = dS K_j          (SS: the same smem tile read as an M-major A operand) -> fp32 red.add
//   3. sattn_dq_cast_kernel   dq_accum fp32 -> d_q fp16
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include <stdlib.h>

#include "sta_host.h"

namespace sta {

constexpr int kSBBlockBytes = 128 * 128;

template <int D>
struct SattnBwdCfg {
  static constexpr int DMMA = (D + 15) / 16 * 16;
  static constexpr int NBLK = (D + 63) / 64;
  static constexpr int TILE = NBLK * kSBBlockBytes;
  // TMEM: S^T [0,128) dP^T [128,256) dV, dK, dQ accumulators of NACC columns each.  For D = 160 three 160-column
  // accumulators do not fit: the head dim is processed in NPASS = 2 column passes split at a 64-column smem
  // block boundary ([0,128) then [128,160)), S^T / dP^T are recomputed per pass, and the dQ accumulator shares
  // columns with S^T / dP^T (ALIAS_DQ).  These layers have N <= 576 tokens, so the extra work is negligible.
  static constexpr int NPASS = (DMMA > 128) ? 2 : 1;
  static constexpr int NACC_MAX = (NPASS == 1) ? DMMA : 128;
  static constexpr bool ALIAS_DQ = NPASS > 1;
  // MODE 2 (one smem block per tile, d <= 64): dS^T double-buffered, S^T/dP^T of the NEXT query tile are issued
  //         half a tile ahead, so the tensor pipe and the two math warpgroups overlap.
  // MODE 1 / 0: single dS^T buffer; the next tile's S^T/dP^T follow dQ (MODE 0 also waits for the dQ drain).
  static constexpr int MODE = ALIAS_DQ ? 0 : (NBLK == 1 ? 2 : 1);
  static constexpr int ST = (NBLK == 3) ? 1 : 2;            // (Q_i, dO_i) ring depth
  static constexpr int NDS = (MODE == 2) ? 2 : 1;           // dS^T buffers
  static constexpr int TMEM_DV = 256;
  static constexpr int TMEM_DK = 256 + NACC_MAX;
  static constexpr int TMEM_DQ = ALIAS_DQ ? 0 : 256 + 2 * NACC_MAX;
  // MODE 2 also double-buffers the dQ accumulator (256 + 4*48 = 448 columns) so that the math warps drain dQ of
  // tile i-1 AFTER the P/dS math of tile i, off the critical path.
  static constexpr bool PIPE_DRAIN = (MODE == 2) && (256 + 4 * NACC_MAX <= 512);
  static constexpr int DQ_STRIDE = PIPE_DRAIN ? NACC_MAX : 0;
  static constexpr int DS_BYTES = 2 * kSBBlockBytes;  // dS^T: 128 keys x 128 queries fp16
  static constexpr int STAGE_BYTES = PIPE_DRAIN ? 128 * D * 4 : 0;  // fp32 dQ tile staged for the TMA reduce-add
  static constexpr int SMEM_TOTAL = 2 * TILE + 2 * ST * TILE + NDS * DS_BYTES + STAGE_BYTES + 1024;
  // Math warpgroups: query tile i is handled as two half-tiles (64 query columns each, the MMA granularity); with
  // NWG = 4 each half is shared by two warpgroups (32 columns each) so that every scheduler has four math warps to
  // hide the LDS / MUFU / TMEM latencies (with two the math phase ran at 43 % MUFU, 47 % issue utilisation).
  static constexpr int NWG = (MODE == 2) ? 4 : 2;
  static constexpr int CW = 128 / NWG;              // query columns per warpgroup
  static constexpr int CH = (NWG == 4) ? 16 : 32;   // columns per TMEM load / inner chunk (register budget)
  static constexpr int DQ_W0 = (NWG == 4) ? 16 : 24;             // head-dim columns drained by warpgroup 0
  static constexpr int DQ_W1 = (NWG == 4) ? 8 : (D > 24 ? D - 24 : 8);  // ... by every other warpgroup
  // P^T (packed fp16, A operand of dV += P^T dO): over the S^T columns of its half-tile when ONE warpgroup owns the
  // half (NWG = 2); with two warpgroups per half that aliasing would race (one group's P^T lands on S^T columns the
  // other has not read yet), so P^T gets its own 64 columns: 256 + 4*48 + 64 = 512.
  static constexpr int TMEM_P = (NWG == 4) ? 448 : 0;
  static constexpr int P_HALF = (NWG == 4) ? 32 : 64;  // TMEM columns between the P^T of half 0 and half 1
  static_assert(NWG == 2 || 256 + 4 * NACC_MAX + 64 <= 512, "TMEM budget");
  static constexpr int THREADS = 64 + 128 * NWG;  // TMA warp, MMA warp, math warpgroups
};

struct SattnBwdParams {
  const float* lse;    // [b, h, n]
  const float* delta;  // [b, h, n]
  float* dq_accum;     // [b, n, h*D] fp32
  __half* d_k;
  __half* d_v;         // [b, n, h*D] fp16, token stride d_tok (batch stride n * d_tok)
  long long d_tok;
  int n, heads;
  float scale, scale_log2;
  unsigned int* err;
  int dbg;
};

// cycle counters of CTA (0,0,0) when STA_DEBUG_FLAGS & 8: [0..4] math warp 2: wait sdp, compute, wait dq, drain, total;
// [8..10] MMA thread: wait pds, issue, total
__device__ long long g_bwd_dbg[16];
// Ablation / cycle-counter paths (skip dQ accumulation, skip dS stores, identity instead of exp2, counters) exist only in
// a library built with -DSTA_BWD_DEBUG (tools/bwd_cycles.py); in the product they are compiled out: no environment
// variable can change the gradients, and the branches cost no registers.
#ifdef STA_BWD_DEBUG
#define STA_BWD_DBG (p.dbg)
#else
#define STA_BWD_DBG 0
#endif

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Query tile i is processed as two halves of 64 query columns (h = 0, 1).  Math warpgroup h owns half h: one thread per
// key row (TMEM lane), 64 columns of S^T and dP^T.
template <int D>
__global__ void __launch_bounds__(SattnBwdCfg<D>::THREADS, 1)
sattn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dq1,
                 const SattnBwdParams p) {
  using Cfg = SattnBwdCfg<D>;
  constexpr int ST = Cfg::ST, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA, MODE = Cfg::MODE;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sK = smem;
  unsigned char* sV = sK + Cfg::TILE;
  unsigned char* sQ = sV + Cfg::TILE;          // ring: stage s -> Q at sQ + s*2*TILE, dO right after it
  unsigned char* sDS = sQ + 2 * ST * Cfg::TILE;
  float* sStage = reinterpret_cast<float*>(sDS + Cfg::NDS * Cfg::DS_BYTES);  // [128][D] fp32 (PIPE_DRAIN only)

  __shared__ uint64_t kv_full, qdo_full[ST], qdo_empty[ST], sdp_full[2], pds_ready[2], dq_full, dq_drained;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  // -lse*log2e and -delta*scale of the current query tile; double-buffered when there are four warpgroups (no
  // named-barrier ids left for a "done reading" barrier per group, and no need for one then)
  constexpr int NSB = (Cfg::NWG == 4) ? 2 : 1;
  __shared__ __align__(16) float s_lse2_buf[NSB][128], s_delta_buf[NSB][128];

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  // head dim 160: the two accumulator column passes ([0,128) and [128,160)) run in DIFFERENT CTAs (grid.x = key tiles x
  // NPASS): each recomputes S^T / dP^T / P / dS for its key tile, which costs nothing on the idle SMs of these small
  // layers (N <= 576: 32 CTAs before), and halves the serial chain of a CTA.
  const int j = blockIdx.x / Cfg::NPASS, pass_cta = blockIdx.x % Cfg::NPASS, h = blockIdx.y, b = blockIdx.z;
  const int n = p.n;
  const int T = (n + 127) / 128;
  const int total = T;

  if (tid == 0) {
    dead = 0;
    mbar_init(&kv_full, 1);
    for (int i = 0; i < ST; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&pds_ready[i], 2 * Cfg::NWG); }
    mbar_init(&dq_full, 1);
    mbar_init(&dq_drained, 4 * Cfg::NWG);
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
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      mbar_expect_tx_w(&kv_full, 2 * Cfg::TILE);
      for (int blk = 0; blk < NBLK; ++blk) {
        tma_load_4d_w(sK + blk * kSBBlockBytes, &tm_k, &kv_full, blk * 64, h, j * 128, b);
        tma_load_4d_w(sV + blk * kSBBlockBytes, &tm_v, &kv_full, blk * 64, h, j * 128, b);
      }
      for (int it = 0; it < total; ++it) {
        const int st = it % ST, i = it % T;
        if (!mbar_wait_warp(&qdo_empty[st], ((it / ST) & 1) ^ 1, &dead, p.err, 10)) break;
        mbar_expect_tx_w(&qdo_full[st], 2 * Cfg::TILE);
        unsigned char* dq = sQ + st * 2 * Cfg::TILE;
        for (int blk = 0; blk < NBLK; ++blk) {
          tma_load_4d_w(dq + blk * kSBBlockBytes, &tm_q, &qdo_full[st], blk * 64, h, i * 128, b);
          tma_load_4d_w(dq + Cfg::TILE + blk * kSBBlockBytes, &tm_do, &qdo_full[st], blk * 64, h, i * 128, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);              // K-major, 128B swizzle
      constexpr uint64_t mndesc_hi = umma_desc_hi_sw128(kSBBlockBytes, 1024);  // MN-major, 128-row blocks
      constexpr uint32_t idesc_half = umma_idesc_f16(128, 64, 0, 0);           // [128 keys x d] x [64 queries x d]^T
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);

      auto stage_addr = [&](int it) { return smem_u32(sQ + (it % ST) * 2 * Cfg::TILE); };
      // S^T_half = K_j Q_half^T and dP^T_half = V_j dO_half^T; half hf covers query rows [64 hf, 64 hf + 64)
      auto issue_sdp = [&](int it, int hf) {
        const uint32_t q_addr = stage_addr(it) + hf * 8192, do_addr = q_addr + Cfg::TILE;
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t off = (k / 4) * kSBBlockBytes + (k % 4) * 32;
          umma_ss_w(tmem + hf * 64, umma_desc(kdesc_hi, k_addr + off), umma_desc(kdesc_hi, q_addr + off), idesc_half, k > 0);
        }
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t off = (k / 4) * kSBBlockBytes + (k % 4) * 32;
          umma_ss_w(tmem + 128 + hf * 64, umma_desc(kdesc_hi, v_addr + off), umma_desc(kdesc_hi, do_addr + off), idesc_half,
                  k > 0);
        }
        umma_commit_w(&sdp_full[hf]);
      };

      bool ok = mbar_wait_warp(&kv_full, 0, &dead, p.err, 20) && mbar_wait_warp(&qdo_full[0], 0, &dead, p.err, 21);
      if (ok) {
        tc_fence_after();
        issue_sdp(0, 0);
        issue_sdp(0, 1);
      }
      const bool prof = (STA_BWD_DBG & 8) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long m_wait = 0, m_wait_q = 0, m_all = clock64(), mt;
      for (int it = 0; it < total && ok; ++it) {
        const int st = it % ST, pass = pass_cta, i = it;
        const int nacc = (Cfg::NPASS == 1) ? DMMA : (pass == 0 ? 128 : DMMA - 128);
        const uint32_t col_off = pass * 2 * kSBBlockBytes;               // first 64-column block of this pass
        const uint32_t idesc_acc = umma_idesc_f16(128, nacc, 0, 1);      // A K-major (TMEM or smem), B MN-major
        const uint32_t idesc_dq = umma_idesc_f16(128, nacc, 1, 1);       // A M-major (dS^T read transposed)
        const uint32_t q_addr = stage_addr(it), do_addr = q_addr + Cfg::TILE;
        const uint32_t ds_addr = smem_u32(sDS + (it % Cfg::NDS) * Cfg::DS_BYTES);
        const bool more = it + 1 < total;
        for (int hf = 0; hf < 2 && ok; ++hf) {
          mt = clock64();
          ok = mbar_wait_warp(&pds_ready[hf], it & 1, &dead, p.err, 23 + hf);
          m_wait += clock64() - mt;
          if (!ok) break;
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dV += P^T_half dO_half   (K = 64 queries of this half)
            umma_ts_w(tmem + Cfg::TMEM_DV, tmem + Cfg::TMEM_P + hf * Cfg::P_HALF + k * 8,
                    umma_desc(mndesc_hi, do_addr + col_off + hf * 8192 + k * 2048), idesc_acc, i > 0 || hf > 0 || k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dK += dS^T_half Q_half
            umma_ss_w(tmem + Cfg::TMEM_DK, umma_desc(kdesc_hi, ds_addr + hf * kSBBlockBytes + k * 32),
                    umma_desc(mndesc_hi, q_addr + col_off + hf * 8192 + k * 2048), idesc_acc, i > 0 || hf > 0 || k > 0);
          if (hf == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k)  // dQ_i = dS K_j   (M = 128 queries, K = 128 keys)
              umma_ss_w(tmem + Cfg::TMEM_DQ + (it & 1) * Cfg::DQ_STRIDE, umma_desc(mndesc_hi, ds_addr + k * 2048),
                      umma_desc(mndesc_hi, k_addr + col_off + k * 2048), idesc_dq, k > 0);
            umma_commit_w(&dq_full);
            umma_commit_w(&qdo_empty[st]);
          }
          if (more) {
            if (MODE == 2) {
              // next tile's half hf: its S^T/dP^T columns and P^T alias were consumed by the MMAs issued above
              if (hf == 0) {
                mt = clock64();
                ok = mbar_wait_warp(&qdo_full[(it + 1) % ST], ((it + 1) / ST) & 1, &dead, p.err, 25);
                m_wait_q += clock64() - mt;
                if (!ok) break;
                tc_fence_after();
              }
              issue_sdp(it + 1, hf);
            } else if (hf == 1) {
              ok = mbar_wait_warp(&qdo_full[(it + 1) % ST], ((it + 1) / ST) & 1, &dead, p.err, 25);
              if (ok && MODE == 0) ok = mbar_wait_warp(&dq_drained, it & 1, &dead, p.err, 26);
              if (!ok) break;
              tc_fence_after();
              issue_sdp(it + 1, 0);
              issue_sdp(it + 1, 1);
            }
          }
        }
      }
      if (prof) { g_bwd_dbg[8] = m_wait; g_bwd_dbg[9] = m_wait_q; g_bwd_dbg[10] = clock64() - m_all; }
    }
  } else {
    // ===================================== per-key-row math ==================================
    constexpr int NWG = Cfg::NWG, CW = Cfg::CW, CH = Cfg::CH;
    const int g = (warp - 2) >> 2;           // math warpgroup: query columns [g*CW, (g+1)*CW) of the tile
    const int hf = g / (NWG / 2);            // the half-tile (MMA granularity) those columns belong to
    const int r = ((warp & 3) << 5) + lane;  // key row for S^T/dP^T/dK/dV, query row for dQ_i
    const int t128 = tid - 64 - g * 128;     // 0..127 inside the warpgroup
    const int key = j * 128 + r;
    const bool key_ok = key < n;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const float* lse_bh = p.lse + ((long long)b * p.heads + h) * n;
    const float* delta_bh = p.delta + ((long long)b * p.heads + h) * n;
    auto drain_dq = [&](int it_d) {
      const int q0row = it_d * 128, col0 = pass_cta * 128;
      const int ncols = (Cfg::NPASS == 1) ? D : (pass_cta == 0 ? 128 : D - 128);
      const uint32_t src = lane_addr + Cfg::TMEM_DQ + (it_d & 1) * Cfg::DQ_STRIDE;
      if (Cfg::PIPE_DRAIN) {
        // TMEM -> fp32 smem tile -> ONE TMA reduce-add per warpgroup and tile (the per-thread red.global path below
        // costs ~1500 cycles per tile in LSU back-pressure; the bulk reduction is asynchronous).  Warpgroup 0 owns
        // head-dim columns [0, 24), warpgroup 1 columns [24, D): no synchronisation between the warpgroups.
        constexpr int W0 = Cfg::DQ_W0, W1 = Cfg::DQ_W1;
        const int wcols = g == 0 ? W0 : W1, cbase = g == 0 ? 0 : W0 + (g - 1) * W1;
        float* stage = sStage + 128 * cbase;
        if (t128 == 0) bulk_wait_group_read0();  // this warpgroup's previous reduce has finished reading its tile
        named_bar_sync(5 + g, 128);
#pragma unroll
        for (int c0 = 0; c0 < W0; c0 += 8) {
          if (c0 >= wcols) break;
          uint32_t o[8];
          tmem_ld8(src + cbase + c0, o);
          tmem_ld_wait();
          // rows of 64 B (W0 = 16 floats) / 32 B (W1 = 8): thread r of a warp writes row r, so without a swizzle the eight
          // threads of a store phase hit 2 (4) bank groups — 4-way (2-way) conflicts, 5.5 M extra wavefronts per launch in
          // ncu.  The TMA maps use SWIZZLE_64B / SWIZZLE_32B (16-byte chunk index ^= address bits [7,8] / [7]).
          const int ch = c0 >> 2;  // first 16-byte chunk of this 8-float group
          const int sx = (wcols == 16) ? ((r >> 1) & 3) : ((r >> 2) & 1);
          float4* row4 = reinterpret_cast<float4*>(stage + r * wcols);
          row4[ch ^ sx] = make_float4(__uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]));
          row4[(ch + 1) ^ sx] = make_float4(__uint_as_float(o[4]), __uint_as_float(o[5]), __uint_as_float(o[6]), __uint_as_float(o[7]));
        }
        fence_proxy_async_smem();
        named_bar_sync(5 + NWG + g, 128);
        if (t128 == 0 && !(STA_BWD_DBG & 1)) {
          tma_reduce_add_4d(g == 0 ? &tm_dq : &tm_dq1, stage, cbase, h, q0row, b);  // OOB rows are clipped by TMA
          bulk_commit_group();
        }
        return;
      }
      const int qrow = q0row + r;
      float* dst = p.dq_accum + ((long long)b * n + qrow) * (p.heads * D) + h * D + col0;
#pragma unroll
      for (int c0 = 0; c0 < Cfg::NACC_MAX; c0 += 8 * NWG) {
        const int cc = c0 + 8 * g;
        if (cc >= ncols) break;
        uint32_t o[8];
        tmem_ld8(src + cc, o);
        tmem_ld_wait();
        if (qrow < n && !(STA_BWD_DBG & 1)) {
          red_add_v4(dst + cc, __uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]));
          red_add_v4(dst + cc + 4, __uint_as_float(o[4]), __uint_as_float(o[5]), __uint_as_float(o[6]), __uint_as_float(o[7]));
        }
      }
    };
    // per-thread prefetch of one staged statistic: threads 0..63 of a warpgroup own -lse*log2e, 64..127 own -delta*scale
    auto load_stat = [&](int it_n) -> float {
      if (it_n >= total) return 0.f;
      if (t128 >= 2 * CW) return 0.f;
      const int qi = (it_n % T) * 128 + g * CW + (t128 % CW);
      if (t128 < CW) return qi < n ? -lse_bh[qi] * 1.4426950408889634f : -INFINITY;
      return qi < n ? -delta_bh[qi] * p.scale : 0.f;
    };
    float pre_val = load_stat(0);
    bool ok = true;
    const bool prof = (STA_BWD_DBG & 8) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
    long long c_wait_sdp = 0, c_comp = 0, c_wait_dq = 0, c_drain = 0, t_all = clock64(), tt;
    for (int it = 0; it < total; ++it) {
      const int pass = pass_cta, i = it;
      const int col0 = pass * 128;                                             // first head-dim column of this pass
      const int ncols = (Cfg::NPASS == 1) ? D : (pass == 0 ? 128 : D - 128);   // columns of this pass
      // stage lse / delta of this warpgroup's 64 query columns for broadcast reads (values were prefetched from
      // global memory one tile ahead, so the load latency is off the critical path)
      float* s_lse2 = s_lse2_buf[it % NSB];
      float* s_delta = s_delta_buf[it % NSB];
      if (t128 < CW) s_lse2[g * CW + t128] = pre_val;
      else if (t128 < 2 * CW) s_delta[g * CW + t128 - CW] = pre_val;
      named_bar_sync(1 + g, 128);
      pre_val = load_stat(it + 1);
      tt = clock64();
      ok = mbar_wait_warp(&sdp_full[hf], it & 1, &dead, p.err, 30);
      if (!ok) break;
      tc_fence_after();
      c_wait_sdp += clock64() - tt; tt = clock64();
      const float* lse2 = s_lse2 + g * CW;
      const float* dlt = s_delta + g * CW;
      // dS^T block of this half-tile; this warpgroup's columns start at 16-byte chunk (g*CW % 64) / 8 of each row
      unsigned char* ds_blk = sDS + (it % Cfg::NDS) * Cfg::DS_BYTES + hf * kSBBlockBytes;
      const int col_in_half = (g * CW) & 63;
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += CH) {
        uint32_t s[CH], dp[CH];
        if (CH == 32) {
          tmem_ld32(lane_addr + g * CW + c0, s);
          tmem_ld32(lane_addr + 128 + g * CW + c0, dp);
        } else {
          tmem_ld16(lane_addr + g * CW + c0, s);
          tmem_ld16(lane_addr + 128 + g * CW + c0, dp);
        }
        tmem_ld_wait();
        uint32_t pk[CH / 2], dk[CH / 2];
#pragma unroll
        for (int q = 0; q < CH; q += 4) {
          const float4 nl = *reinterpret_cast<const float4*>(lse2 + c0 + q);   // broadcast reads, 4 columns at a time
          const float4 nd = *reinterpret_cast<const float4*>(dlt + c0 + q);
          const float nlv[4] = {nl.x, nl.y, nl.z, nl.w}, ndv[4] = {nd.x, nd.y, nd.z, nd.w};
          float pv[4], dv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = fmaf(__uint_as_float(s[q + e]), p.scale_log2, nlv[e]);
            pv[e] = (STA_BWD_DBG & 4) ? x : fast_exp2(x);
            dv[e] = pv[e] * fmaf(__uint_as_float(dp[q + e]), p.scale, ndv[e]);  // scale * P * (dP - delta)
          }
          pk[q >> 1] = pack_half2(pv[0], pv[1]);
          pk[(q >> 1) + 1] = pack_half2(pv[2], pv[3]);
          dk[q >> 1] = pack_half2(dv[0], dv[1]);
          dk[(q >> 1) + 1] = pack_half2(dv[2], dv[3]);
        }
        if (!key_ok) {  // keys past the end of the sequence (last tile only): P = dS = 0
#pragma unroll
          for (int e = 0; e < CH / 2; ++e) { pk[e] = 0u; dk[e] = 0u; }
        }
        // P^T (packed fp16): query column c of the half-tile -> TMEM column TMEM_P + hf*P_HALF + c/2
        if (CH == 32) tmem_st16(lane_addr + Cfg::TMEM_P + hf * Cfg::P_HALF + ((col_in_half + c0) >> 1), pk);
        else tmem_st8(lane_addr + Cfg::TMEM_P + hf * Cfg::P_HALF + ((col_in_half + c0) >> 1), pk);
        // dS^T row r, query columns [g*CW + c0, +CH): CH/8 16-byte chunks of the half-tile's smem block
        const int chunk0 = (col_in_half + c0) >> 3;
#pragma unroll
        for (int cc = 0; cc < CH / 8; ++cc) {
          uint4 v = make_uint4(dk[4 * cc], dk[4 * cc + 1], dk[4 * cc + 2], dk[4 * cc + 3]);
          if (!(STA_BWD_DBG & 2)) *reinterpret_cast<uint4*>(ds_blk + sw128_offset(r, chunk0 + cc)) = v;
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      const int it_d = Cfg::PIPE_DRAIN ? it - 1 : it;
      if (Cfg::PIPE_DRAIN && it_d >= 0) {
        // Wait for dQ of the previous tile BEFORE releasing this tile to the MMA warp: a parity wait must never fall
        // two phases behind, and completion #it of dq_full cannot happen before pds_ready(it).
        ok = mbar_wait_warp(&dq_full, it_d & 1, &dead, p.err, 31);
        if (!ok) break;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_ready[hf]);
      if (NSB == 1) named_bar_sync(3 + g, 128);  // the warpgroup is done reading its lse / delta staging area
      c_comp += clock64() - tt; tt = clock64();
      // ---- drain dQ into the fp32 accumulator: the warpgroups take alternate 8-column chunks.  With PIPE_DRAIN the
      //      tile drained here is the PREVIOUS one (its MMAs finished while this tile's P/dS math ran) ----
      if (it_d >= 0) {
        if (!Cfg::PIPE_DRAIN) {
          ok = mbar_wait_warp(&dq_full, it_d & 1, &dead, p.err, 31);
          if (!ok) break;
        }
        tc_fence_after();
        c_wait_dq += clock64() - tt; tt = clock64();
        drain_dq(it_d);
      }
      if (Cfg::PIPE_DRAIN && it == total - 1) {  // last tile: nothing left to overlap with
        ok = mbar_wait_warp(&dq_full, it & 1, &dead, p.err, 32);
        if (!ok) break;
        tc_fence_after();
        drain_dq(it);
      }
      if (i == T - 1) {
        // ---- end of a pass: dK, dV columns [col0, col0 + ncols) of this key row (dq_full => all MMAs completed) ----
        __half* dk_row = p.d_k + ((long long)b * n + key) * p.d_tok + h * D + col0;
        __half* dv_row = p.d_v + ((long long)b * n + key) * p.d_tok + h * D + col0;
#pragma unroll
        for (int c0 = 0; c0 < Cfg::NACC_MAX; c0 += 8 * NWG) {
          const int cc = c0 + 8 * g;
          if (cc >= ncols) break;
          uint32_t a[8], c[8];
          tmem_ld8(lane_addr + Cfg::TMEM_DK + cc, a);
          tmem_ld8(lane_addr + Cfg::TMEM_DV + cc, c);
          tmem_ld_wait();
          if (key_ok) {
            uint4 v;
            v.x = pack_half2(__uint_as_float(a[0]), __uint_as_float(a[1]));
            v.y = pack_half2(__uint_as_float(a[2]), __uint_as_float(a[3]));
            v.z = pack_half2(__uint_as_float(a[4]), __uint_as_float(a[5]));
            v.w = pack_half2(__uint_as_float(a[6]), __uint_as_float(a[7]));
            *reinterpret_cast<uint4*>(dk_row + cc) = v;
            v.x = pack_half2(__uint_as_float(c[0]), __uint_as_float(c[1]));
            v.y = pack_half2(__uint_as_float(c[2]), __uint_as_float(c[3]));
            v.z = pack_half2(__uint_as_float(c[4]), __uint_as_float(c[5]));
            v.w = pack_half2(__uint_as_float(c[6]), __uint_as_float(c[7]));
            *reinterpret_cast<uint4*>(dv_row + cc) = v;
          }
        }
      }
      c_drain += clock64() - tt;
      if (MODE == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_drained);
      }
    }
    if (Cfg::PIPE_DRAIN && t128 == 0) bulk_wait_group0();  // all reduce-adds have landed before the kernel ends
    if (prof) {
      g_bwd_dbg[0] = c_wait_sdp; g_bwd_dbg[1] = c_comp; g_bwd_dbg[2] = c_wait_dq; g_bwd_dbg[3] = c_drain;
      g_bwd_dbg[4] = clock64() - t_all;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// delta[b,h,i] = <dO[b,i,h,:], O[b,i,h,:]>, and dq_accum[b,i,h,:] = 0 (the accumulator the main kernel reduce-adds into: no
// separate memset node).  L = 8 / 16 / 32 lanes per (b, i, h) for D = 40 / 80 / 160, one 16-byte vector of O and dO per lane
// (D/8 of the L lanes are active), partial dot products summed with shuffles: every launch is one memory round trip.  The
// first version used one thread per (b, i, h) with D/4 dependent-address loads: 6.9 us for the 4096 (b, i, h) of N = 256.
template <int D>
__global__ void __launch_bounds__(256) sattn_delta_kernel(const __half* __restrict__ o, const __half* __restrict__ d_o,
                                                          float* __restrict__ delta, float* __restrict__ dq_accum, int n,
                                                          int heads, long long o_ts, long long o_bs, long long do_ts,
                                                          long long do_bs) {
  constexpr int L = D <= 64 ? 8 : (D <= 128 ? 16 : 32);
  constexpr int VEC = D / 8;
  // grid.y = batch sample, 32-bit index math inside it: the first version decomposed a 64-bit flat index with four emulated
  // 64-bit divisions (197 of its 312 SASS instructions) and was issue-bound (5.8 us for the 64 x 64 level)
  const unsigned int item = (blockIdx.x * blockDim.x + threadIdx.x) / L;  // (token, head) of this sample
  const int sub = threadIdx.x & (L - 1);
  const unsigned int per_sample = (unsigned int)n * (unsigned int)heads;
  const bool live = item < per_sample;  // dead items still take part in the shuffles
  const unsigned int it = live ? item : per_sample - 1;
  const unsigned int i = it / (unsigned int)heads, h = it - i * (unsigned int)heads;
  const long long b = blockIdx.y;
  float acc = 0.f;
  if (sub < VEC) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + b * o_bs + i * o_ts + h * D + sub * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + b * do_bs + i * do_ts + h * D + sub * 8);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = __half22float2(ah[t]), y = __half22float2(gh[t]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
    if (live) {  // the lane's 8 columns of the fp32 dQ accumulator
      float4* z = reinterpret_cast<float4*>(dq_accum + (((b * n + i) * heads + h) * D) + sub * 8);
      z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int off = L >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (live && sub == 0) delta[(b * heads + h) * n + i] = acc;
}

// dq_accum fp32 [rows, c] dense -> d_q fp16 [rows, c] with row stride d_tok (d_q may be a slice of one d(qkv) buffer)
__global__ void sattn_dq_cast_kernel(const float4* __restrict__ src, __half* __restrict__ dst, unsigned int n4, FastDiv c4,
                                     long long d_tok) {
  const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;  // n4 < 2^31 (checked by the launcher)
  if (idx >= n4) return;
  const float4 v = src[idx];
  uint2 o;
  o.x = pack_half2(v.x, v.y);
  o.y = pack_half2(v.z, v.w);
  unsigned int row, col4;
  fast_divmod(idx, c4, row, col4);  // (a 64-bit idx / c4 was 111 of this kernel's 168 SASS instructions)
  *reinterpret_cast<uint2*>(dst + (long long)row * d_tok + col4 * 4) = o;
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
  CUtensorMap tm_dq, tm_dq1;
  {
    const uint64_t st[4] = {4, (uint64_t)D * 4, (uint64_t)C * 4, (uint64_t)a->n * C * 4};
    const uint32_t bx0[4] = {(uint32_t)Cfg::DQ_W0, 1, 128, 1}, bx1[4] = {(uint32_t)Cfg::DQ_W1, 1, 128, 1};
    // PIPE_DRAIN stages 64-byte / 32-byte rows with the matching shared-memory swizzle (bank-conflict-free stores)
    const int sw0 = (Cfg::PIPE_DRAIN && Cfg::DQ_W0 == 16) ? 64 : 0, sw1 = (Cfg::PIPE_DRAIN && Cfg::DQ_W1 == 8) ? 32 : 0;
    if ((rc = make_tmap_f32_dense(&tm_dq, a->dq_accum, 4, dims, st, bx0, sw0))) return rc;
    if ((rc = make_tmap_f32_dense(&tm_dq1, a->dq_accum, 4, dims, st, bx1, sw1))) return rc;
  }
  {  // delta and the zeroing of dq_accum in one launch
    constexpr int L = D <= 64 ? 8 : (D <= 128 ? 16 : 32);
    const long long per_sample = (long long)a->n * a->heads;
    if (per_sample * L >= (1ll << 31) || a->batch > 65535)
      return fail(STA_ERR_UNSUPPORTED, "sta_sattn_bwd: n * heads = %lld / batch %d exceed the 32-bit index range", per_sample, a->batch);
    const dim3 dgrid((unsigned)((per_sample * L + 255) / 256), (unsigned)a->batch);
    sattn_delta_kernel<D><<<dgrid, 256, 0, stream>>>(
        reinterpret_cast<const __half*>(a->out), reinterpret_cast<const __half*>(a->d_out), a->delta, a->dq_accum, a->n,
        a->heads, a->o_token_stride, a->o_batch_stride, a->do_token_stride, a->do_batch_stride);
  }
  STA_CUDA_CHECK(cudaGetLastError());

  SattnBwdParams p;
  p.lse = a->lse;
  p.delta = a->delta;
  p.dq_accum = a->dq_accum;
  p.d_k = reinterpret_cast<__half*>(a->d_k);
  p.d_v = reinterpret_cast<__half*>(a->d_v);
  p.d_tok = a->dqkv_token_stride > 0 ? a->dqkv_token_stride : C;
  p.n = a->n;
  p.heads = a->heads;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
#ifdef STA_BWD_DEBUG
  p.dbg = getenv("STA_DEBUG_FLAGS") ? atoi(getenv("STA_DEBUG_FLAGS")) : 0;  // tools/bwd_cycles.py builds this variant
#else
  p.dbg = 0;
#endif
  static PerDeviceOnce smem_attr;
  if ((rc = smem_attr.run([] {
        return cudaFuncSetAttribute(sattn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_TOTAL);
      })))
    return rc;
  dim3 grid(((a->n + 127) / 128) * Cfg::NPASS, a->heads, a->batch);
  sattn_bwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_TOTAL, stream>>>(tm_q, tm_k, tm_v, tm_do, tm_dq, tm_dq1, p);
  STA_CUDA_CHECK(cudaGetLastError());

  const long long n4 = (long long)a->batch * a->n * C / 4;
  if (n4 >= (1ll << 31)) return fail(STA_ERR_UNSUPPORTED, "sta_sattn_bwd: %lld gradient vectors exceed the 32-bit index range", n4);
  FastDiv c4;
  c4.d = (unsigned int)(C / 4);
  make_fast_div_raw(c4.d, &c4.magic, &c4.shift);
  sattn_dq_cast_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(a->dq_accum), reinterpret_cast<__half*>(a->d_q), (unsigned int)n4, c4, p.d_tok);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta

extern "C" int sta_debug_read(long long* out, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, sta::g_bwd_dbg, sizeof(long long) * (n < 16 ? n : 16)) == cudaSuccess ? 0 : 3;
}

extern "C" int sta_sattn_bwd(const sta_sattn_bwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->q || !a->k || !a->v || !a->out || !a->d_out || !a->lse || !a->d_q || !a->d_k || !a->d_v || !a->delta ||
      (!a->dq_accum && a->head_dim != 512))  // the 512-wide path owns every gradient element in one CTA: no fp32 accumulator
    return fail(STA_ERR_BAD_ARG, "sta_sattn_bwd: null pointer");
  if (a->batch < 1 || a->n < 1 || a->heads < 1) return fail(STA_ERR_BAD_ARG, "sta_sattn_bwd: empty shape");
  if ((a->o_token_stride % 8) || (a->o_batch_stride % 8) || (a->do_token_stride % 8) || (a->do_batch_stride % 8) ||
      ((reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->d_out) | reinterpret_cast<uintptr_t>(a->dq_accum)) & 15u))
    return fail(STA_ERR_UNSUPPORTED, "sta_sattn_bwd: out / d_out rows and dq_accum must be 16-byte aligned");
  if (a->dqkv_token_stride < 0 || (a->dqkv_token_stride % 8) ||
      (a->dqkv_token_stride > 0 && a->dqkv_token_stride < (int64_t)a->heads * a->head_dim))
    return fail(STA_ERR_BAD_ARG, "sta_sattn_bwd: dqkv_token_stride must be 0 (dense) or a multiple of 8 >= heads*head_dim");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (a->head_dim) {
    case 40: return launch_sattn_bwd<40>(a, s);
    case 80: return launch_sattn_bwd<80>(a, s);
    case 160: return launch_sattn_bwd<160>(a, s);
    case 512: return launch_sattn_wide_bwd(a, s);  // KL-VAE mid-block AttnBlock (model.py:150-202)
    default:
      return fail(STA_ERR_UNSUPPORTED, "sta_sattn_bwd: head_dim %d not built (SD-v1 uses 40/80/160, its VAE 512)", a->head_dim);
  }
}
