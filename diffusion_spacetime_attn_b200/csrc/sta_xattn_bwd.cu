// sta_xattn_bwd.cu — backward of the fused dual cross-attention + alpha-blend: d(out) -> d(q), d(coef).
//
// The contexts are frozen (CLIP text encoder is not trained, reference ldm/models/diffusion/ddpm.py:519-523), so
// no dK/dV is produced; what the alpha optimisation needs (ldm/models/diffusion/plms.py:204-277) is the gradient
// w.r.t. the hidden state (through q) and w.r.t. the per-object blend weights:
//     out_u = A_u,   out_c = A_g + sum_i w_i (A_i - A_u),   w_i = m_i[pix] c_i
//     dA_u = dO_u - sigma dO_c,  dA_g = dO_c,  dA_i = w_i dO_c,      sigma = sum_i w_i
//     dS_x = P_x o (dA_x V_x^T - rowsum(P_x o dA_x V_x^T)),  dQ = scale * sum_x dS_x K_x
//     dc_i = sum_pix m_i <dO_c, A_i - A_u> = sum_pix m_i (delta_i - delta_uc),
//            delta_i  = rowsum(P_i o dO_c V_i^T)   (flash-attention's delta term, so A_i is never rebuilt)
//            delta_uc = <dO_c, A_u> = <dO_c, out_u>  (the forward's unconditional output IS A_u)
//
// Round-2 layout (latency-bound op, see sta_xattn_fwd.cu): CTA = one 128-pixel tile of one (prompt, head);
//   warp 0       control: barrier init, every TMA load, MMA issue, ring refills
//   warps 1..4   math warpgroup 0 (one thread per pixel row)      warps 5..8  math warpgroup 1 (when NBUF == 2)
// Task 0 = unconditional row, tasks 1.. = conditional row's contexts (global, then the objects whose mask is
// non-empty in this tile).  Task t uses TMEM buffer t % NBUF = [S | dP] (2 x 80 columns) and warpgroup t % NBUF:
//   control: S_t = Q K_t^T, dP_t = dA V_t^T (SS)  ->  math: P from the forward's LSE, delta, dS -> packed fp16 over S
//   control: dQ (+)= dS_t K_t (TS, K MN-major)    ->  math: store dQ_u after task 0, dQ_c after the last task
// dA_u = dO_u - sigma dO_c is formed ONCE in shared memory (fp16, in place over the dO_u tile) by warpgroup 0, so the
// unconditional task is an ordinary task (no third accumulator): head dim 40 needs 160 + 2 x 48 = 256 TMEM columns
// and two CTAs are resident per SM; head dims 80 / 160 run one CTA per SM (fewer CTAs than SMs at those levels).
// P is evaluated once per score (kept in registers between the delta pass and the dS pass).
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
  static constexpr int NBUF = (NBLK == 2) ? 2 : 1;  // [S | dP] TMEM buffers = math warpgroups
  static constexpr int ST = (NBLK == 3) ? 1 : 2;    // K and V ring depth
  static constexpr bool SEPQ = NBLK <= 2;           // Q_c has its own tile (else it re-uses Q_u's once S_u is done)
  static constexpr int QTILE = NBLK * kBQBlockBytes;
  static constexpr int CTILE = NBLK * kBCBlockBytes;
  static constexpr int NQ = SEPQ ? 4 : 3;
  static constexpr int SMEM_BYTES = NQ * QTILE + 2 * ST * CTILE + 1024;
  static constexpr int THREADS = 32 * (1 + 4 * NBUF);
  static constexpr int TMEM_DQ = NBUF * 160;  // dQ_u, then dQ_c at + DMMA
  static constexpr int TMEM_COLS = (TMEM_DQ + 2 * DMMA <= 256) ? 256 : 512;
  static constexpr int MIN_CTAS = (TMEM_COLS == 256) ? 2 : 1;
  static_assert(TMEM_DQ + 2 * DMMA <= 512, "TMEM budget");
};

struct XattnBwdParams {
  const uint8_t* mask;
  const float* coef;
  const float* lse;   // [B, heads, 2+n_obj, n]
  const __half* out;  // forward output (only the unconditional rows are read), may be null when n_obj == 0
  __half* d_q;        // [2B, n, heads*D] contiguous
  float* d_coef;      // [B, n_obj]
  int prompts, n, heads, n_obj, ctx_len;
  long long o_token_stride, o_batch_stride;
  float scale, scale_log2;
  unsigned int* err;
};

template <int D>
__global__ void __launch_bounds__(XattnBwdCfg<D>::THREADS, XattnBwdCfg<D>::MIN_CTAS)
xattn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_dq, const XattnBwdParams p) {
  using Cfg = XattnBwdCfg<D>;
  constexpr int ST = Cfg::ST, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA, NBUF = Cfg::NBUF;
  constexpr bool SEPQ = Cfg::SEPQ;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQu = smem;                  // Q_u (head dim 160: re-filled with Q_c once S_u is done)
  unsigned char* sDOu = sQu + Cfg::QTILE;     // dO_u, rewritten in place as dA_u = dO_u - sigma dO_c
  unsigned char* sDOc = sDOu + Cfg::QTILE;
  unsigned char* sQc = SEPQ ? sDOc + Cfg::QTILE : sQu;
  unsigned char* sK = smem + Cfg::NQ * Cfg::QTILE;
  unsigned char* sV = sK + ST * Cfg::CTILE;

  __shared__ uint64_t qu_full, do_full, qc_full, su_done, dau_ready;
  __shared__ uint64_t k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];
  __shared__ uint64_t sdp_full[NBUF], ds_ready[NBUF], dqu_full, dqc_full, list_ready;
  __shared__ unsigned int tile_bits_s;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  __shared__ int tile_slot[2 + kBMaxObj];
  __shared__ int n_tiles_s;
  __shared__ float dcoef_s[kBMaxObj];

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, pr = blockIdx.z;
  const int n = p.n, B = p.prompts, n_obj = p.n_obj, n_slots = 2 + p.n_obj;

  auto load_q = [&](unsigned char* dst, const CUtensorMap* tm, uint64_t* bar, int batch_row) {
    for (int blk = 0; blk < NBLK; ++blk) tma_load_4d_w(dst + blk * kBQBlockBytes, tm, bar, blk * 64, h, q0, batch_row);
  };
  auto load_k = [&](int t, int slot) {
    const int st = t % ST;
    mbar_expect_tx_w(&k_full[st], Cfg::CTILE);
    for (int blk = 0; blk < NBLK; ++blk)
      tma_load_4d_w(sK + (st * NBLK + blk) * kBCBlockBytes, &tm_k, &k_full[st], blk * 64, h, 0, pr * n_slots + slot);
  };
  auto load_v = [&](int t, int slot) {
    const int st = t % ST;
    mbar_expect_tx_w(&v_full[st], Cfg::CTILE);
    for (int blk = 0; blk < NBLK; ++blk)
      tma_load_4d_w(sV + (st * NBLK + blk) * kBCBlockBytes, &tm_v, &v_full[st], blk * 64, h, 0, pr * n_slots + slot);
  };

  if (warp == 0) {
    if (lane == 0) {
      dead = 0;
      tile_bits_s = 0;
      mbar_init(&list_ready, 1);
      mbar_init(&qu_full, 1);
      mbar_init(&do_full, 1);
      mbar_init(&qc_full, 1);
      mbar_init(&su_done, 1);
      mbar_init(&dau_ready, 4);
      for (int i = 0; i < ST; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < NBUF; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&ds_ready[i], 4); }
      mbar_init(&dqu_full, 1);
      mbar_init(&dqc_full, 1);
      mbar_fence_init();
    }
    if (lane < kBMaxObj) dcoef_s[lane] = 0.f;
    __syncwarp();
    mbar_expect_tx_w(&qu_full, Cfg::QTILE);
    load_q(sQu, &tm_q, &qu_full, pr);
    load_k(0, 0);
    mbar_expect_tx_w(&do_full, 2 * Cfg::QTILE);
    load_q(sDOu, &tm_do, &do_full, pr);
    load_q(sDOc, &tm_do, &do_full, pr + B);
    load_v(0, 0);
    if (SEPQ) {
      mbar_expect_tx_w(&qc_full, Cfg::QTILE);
      load_q(sQc, &tm_q, &qc_full, pr + B);
    }
    if (ST >= 2) { load_k(1, 1); load_v(1, 1); }
  } else if (warp == 1) {
    tmem_alloc(&tmem_base_s, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===================================== control warp: TMA + MMA issue =====================================
    constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);
    constexpr uint64_t mndesc_hi = umma_desc_hi_sw128(kBCBlockBytes, 1024);
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 80, 0, 0);      // [128 x D] x [80 x D]^T
    constexpr uint32_t idesc_dq = umma_idesc_f16(128, DMMA, 0, 1);   // dS[128 x 80] x K[80 x D]
    const uint32_t qu_addr = smem_u32(sQu), qc_addr = smem_u32(sQc), dau_addr = smem_u32(sDOu), doc_addr = smem_u32(sDOc);
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);

    // D[col] = A[128 x D] (smem, K-major) * Bt[80 x D] (smem, K-major)
    auto issue_nt = [&](uint32_t col, uint32_t a_addr, uint32_t b_addr) {
#pragma unroll
      for (int k = 0; k < DMMA / 16; ++k) {
        const uint32_t aoff = (k / 4) * kBQBlockBytes + (k % 4) * 32;
        const uint32_t boff = (k / 4) * kBCBlockBytes + (k % 4) * 32;
        umma_ss_w(tmem + col, umma_desc(kdesc_hi, a_addr + aoff), umma_desc(kdesc_hi, b_addr + boff), idesc_s, k > 0);
      }
    };
    // dQ_(t ? c : u) (+)= dS_t (TMEM, packed fp16) * K_t [80 x D] (smem, MN-major)
    auto issue_dq = [&](int t) {
      const int st = t % ST;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        umma_ts_w(tmem + Cfg::TMEM_DQ + (t ? DMMA : 0), tmem + (t % NBUF) * 160 + k * 8,
                  umma_desc(mndesc_hi, k_addr + st * Cfg::CTILE + k * 2048), idesc_dq, (t >= 2) || k > 0);
      umma_commit_w(&k_empty[st]);
    };

    int T = 2, k_next = ST < 2 ? ST : 2, v_next = k_next;  // tasks 0 / 1 always exist; the rest once the list is known
    int sdp_issued = 0, dq_issued = 0;
    bool ok = true;

    auto issue_sdp = [&](int t) {  // conditional-row task t >= 1
      const int st = t % ST;
      ok = mbar_wait_warp(&qc_full, 0, &dead, p.err, 23) &&
           mbar_wait_warp(&k_full[st], (t / ST) & 1, &dead, p.err, 24) &&
           mbar_wait_warp(&v_full[st], (t / ST) & 1, &dead, p.err, 25);
      if (!ok) return;
      tc_fence_after();
      issue_nt((t % NBUF) * 160, qc_addr, k_addr + st * Cfg::CTILE);
      issue_nt((t % NBUF) * 160 + 80, doc_addr, v_addr + st * Cfg::CTILE);
      umma_commit_w(&sdp_full[t % NBUF]);
      umma_commit_w(&v_empty[st]);
    };
    auto refill = [&]() {
      while (ok && k_next < T && k_next - ST < dq_issued) {
        ok = mbar_wait_warp(&k_empty[k_next % ST], (k_next / ST - 1) & 1, &dead, p.err, 10);
        if (ok) load_k(k_next, tile_slot[k_next]);
        ++k_next;
      }
      while (ok && v_next < T && v_next - ST < sdp_issued) {
        ok = mbar_wait_warp(&v_empty[v_next % ST], (v_next / ST - 1) & 1, &dead, p.err, 11);
        if (ok) load_v(v_next, tile_slot[v_next]);
        ++v_next;
      }
    };

    // ---- task 0: S_u = Q_u K_0^T now, dP_u = dA_u V_0^T once warpgroup 0 has formed dA_u ----
    ok = mbar_wait_warp(&qu_full, 0, &dead, p.err, 20) && mbar_wait_warp(&k_full[0], 0, &dead, p.err, 21);
    if (ok) {
      tc_fence_after();
      issue_nt(0, qu_addr, k_addr);
      if (!SEPQ) {
        umma_commit_w(&su_done);
        ok = mbar_wait_warp(&su_done, 0, &dead, p.err, 12);
        if (ok) {
          mbar_expect_tx_w(&qc_full, Cfg::QTILE);
          load_q(sQc, &tm_q, &qc_full, pr + B);
        }
      }
    }
    if (ok) ok = mbar_wait_warp(&list_ready, 0, &dead, p.err, 28);  // built by warpgroup 0 while the loads are in flight
    if (ok) {
      T = n_tiles_s;
      const int first = k_next;
      k_next = v_next = T < ST ? T : ST;
      for (int t = first; t < k_next; ++t) { load_k(t, tile_slot[t]); load_v(t, tile_slot[t]); }
    }
    if (ok) ok = mbar_wait_warp(&dau_ready, 0, &dead, p.err, 22) && mbar_wait_warp(&v_full[0], 0, &dead, p.err, 26);
    if (ok) {
      tc_fence_after();
      issue_nt(80, dau_addr, v_addr);
      umma_commit_w(&sdp_full[0]);
      umma_commit_w(&v_empty[0]);
      sdp_issued = 1;
    }
    while (ok && sdp_issued < T && sdp_issued < NBUF) { refill(); if (ok) issue_sdp(sdp_issued); ++sdp_issued; }

    for (int t = 0; t < T && ok; ++t) {
      refill();
      if (!ok) break;
      ok = mbar_wait_warp(&ds_ready[t % NBUF], (t / NBUF) & 1, &dead, p.err, 27);
      if (!ok) break;
      tc_fence_after();
      issue_dq(t);
      dq_issued = t + 1;
      if (t == 0) umma_commit_w(&dqu_full);
      if (t == T - 1) umma_commit_w(&dqc_full);
      // next [S | dP] pair into the buffer this dQ just consumed (in-order tensor pipe: safe behind issue_dq)
      if (sdp_issued < T && k_next > sdp_issued && v_next > sdp_issued) { issue_sdp(sdp_issued); ++sdp_issued; }
      refill();
      if (ok && sdp_issued < T && sdp_issued <= t + NBUF) { issue_sdp(sdp_issued); ++sdp_issued; }
    }
  } else {
    // ===================================== per-row math ======================================
    const int g = (warp - 1) >> 2;  // math warpgroup
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int r = (quad << 5) + lane;
    const int row = q0 + r;
    const bool row_ok = row < n;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad << 5) << 16);
    const float kLog2e = 1.4426950408889634f;
    const float* lse_row = p.lse + ((long long)pr * p.heads + h) * n_slots * n + row;

    unsigned int bits = 0;
    float sigma = 0.f;
#pragma unroll
    for (int i = 0; i < kBMaxObj; ++i) {
      if (i < n_obj) {
        const bool mk = row_ok && p.mask[((long long)pr * n_obj + i) * n + row] != 0;
        if (mk) { bits |= 1u << i; sigma += p.coef[pr * n_obj + i]; }
      }
    }

    if (g == 0) {
      // task list: tile-level union of the four warps' membership bits (warpgroup 0 covers all 128 pixels)
      const unsigned int wbits = __reduce_or_sync(0xffffffffu, bits);
      if (lane == 0) atomicOr(&tile_bits_s, wbits);
      named_bar_sync(1, 128);
      if (warp == 1 && lane == 0) {
        const unsigned int tb = tile_bits_s;
        int cnt = 2;
        tile_slot[0] = 0;
        tile_slot[1] = 1;
        for (int i = 0; i < n_obj; ++i)
          if ((tb >> i) & 1u) tile_slot[cnt++] = 2 + i;
        n_tiles_s = cnt;
        mbar_arrive(&list_ready);  // release: the list is visible to whoever observes the phase flip
      }
    }

    bool ok = true;
    float delta_uc = 0.f;
    if (n_obj > 0 || g == 0) ok = mbar_wait_warp(&do_full, 0, &dead, p.err, 29);
    if (ok && n_obj > 0) {
      // delta_uc = <dO_c[row], out_u[row]> over this head's D channels (dO_c from the staged tile, out_u from global)
      const __half* orow = p.out + (long long)pr * p.o_batch_stride + (long long)row * p.o_token_stride + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        const uint4 dv = *reinterpret_cast<const uint4*>(sDOc + (c / 8) * kBQBlockBytes + sw128_offset(r, c % 8));
        uint4 ov = make_uint4(0, 0, 0, 0);
        if (row_ok) ov = *reinterpret_cast<const uint4*>(orow + c * 8);
        const __half2* dh = reinterpret_cast<const __half2*>(&dv);
        const __half2* oh = reinterpret_cast<const __half2*>(&ov);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __half22float2(dh[i]), b = __half22float2(oh[i]);
          delta_uc = fmaf(a.x, b.x, delta_uc);
          delta_uc = fmaf(a.y, b.y, delta_uc);
        }
      }
    }
    if (g == 0) {
      // dA_u = dO_u - sigma dO_c, in place over the dO_u tile (skipped by warps whose pixels touch no object)
      if (ok && __any_sync(0xffffffffu, sigma != 0.f)) {
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
          const uint32_t off = (c / 8) * kBQBlockBytes + sw128_offset(r, c % 8);
          uint4 uv = *reinterpret_cast<const uint4*>(sDOu + off);
          const uint4 cv = *reinterpret_cast<const uint4*>(sDOc + off);
          __half2* uh = reinterpret_cast<__half2*>(&uv);
          const __half2* ch = reinterpret_cast<const __half2*>(&cv);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = __half22float2(uh[i]), b = __half22float2(ch[i]);
            uh[i] = __floats2half2_rn(fmaf(-sigma, b.x, a.x), fmaf(-sigma, b.y, a.y));
          }
          *reinterpret_cast<uint4*>(sDOu + off) = uv;
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dau_ready);
    }

    // dQ (fp32, TMEM) -> fp16 -> a dead operand tile in the TMA swizzled layout -> one TMA store per 64 channels
    // (full-line writes instead of 16 bytes per 2*C-byte-strided row per thread)
    auto write_dq = [&](int batch_row, uint32_t col, unsigned char* stage, int bar_id) {
#pragma unroll
      for (int c0 = 0; c0 < DMMA; c0 += 16) {
        uint32_t o[16];
        tmem_ld16(lane_addr + col + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint4 v;
          v.x = pack_half2(__uint_as_float(o[hh * 8 + 0]), __uint_as_float(o[hh * 8 + 1]));
          v.y = pack_half2(__uint_as_float(o[hh * 8 + 2]), __uint_as_float(o[hh * 8 + 3]));
          v.z = pack_half2(__uint_as_float(o[hh * 8 + 4]), __uint_as_float(o[hh * 8 + 5]));
          v.w = pack_half2(__uint_as_float(o[hh * 8 + 6]), __uint_as_float(o[hh * 8 + 7]));
          const int ch = (c0 >> 3) + hh;
          *reinterpret_cast<uint4*>(stage + (ch >> 3) * kBQBlockBytes + sw128_offset(r, ch & 7)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (quad == 0 && lane == 0) {  // one thread issues, commits and waits (bulk groups are per thread)
        for (int blk = 0; blk < NBLK; ++blk) tma_store_4d(&tm_dq, stage + blk * kBQBlockBytes, blk * 64, h, q0, batch_row);
        bulk_commit_group();
        bulk_wait_group_read0();
      }
    };

    int T = 2;  // tasks 0 / 1 always exist; the real count is read once this warpgroup's first task is done
    for (int t = g; t < T && ok; t += NBUF) {
      const int slot = t >= 2 ? tile_slot[t] : t;
      const uint32_t s_addr = lane_addr + (t % NBUF) * 160;
      float w = 1.f;
      bool mk = false, live = true;
      if (slot >= 2) {
        mk = (bits >> (slot - 2)) & 1u;
        w = mk ? p.coef[pr * n_obj + slot - 2] : 0.f;
        live = __any_sync(0xffffffffu, mk);
      }
      const float lse2 = (live && row_ok) ? lse_row[(long long)slot * n] * kLog2e : 0.f;
      ok = mbar_wait_warp(&sdp_full[t % NBUF], (t / NBUF) & 1, &dead, p.err, 32);
      if (!ok) break;
      tc_fence_after();
      float delta = 0.f;
      if (live) {
        uint32_t s[80];
        tmem_ld32(s_addr, s);
        tmem_ld32(s_addr + 32, s + 32);
        tmem_ld16(s_addr + 64, s + 64);
        tmem_ld_wait();
        const int valid = p.ctx_len;
        if (valid >= 76) {  // CLIP's 77 tokens: only the last columns are padding (exp2(-inf) = 0)
#pragma unroll
          for (int c = 76; c < 80; ++c)
            if (c >= valid) s[c] = 0xff800000u;
        } else {
#pragma unroll
          for (int c = 0; c < 76; ++c)
            if (c >= valid) s[c] = 0xff800000u;
#pragma unroll
          for (int c = 76; c < 80; ++c) s[c] = 0xff800000u;
        }
#pragma unroll
        for (int c = 0; c < 80; ++c)
          s[c] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(s[c]), p.scale_log2, -lse2)));
#pragma unroll
        for (int c0 = 0; c0 < 80; c0 += 16) {
          uint32_t a[16];
          tmem_ld16(s_addr + 80 + c0, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) delta = fmaf(__uint_as_float(s[c0 + i]), __uint_as_float(a[i]), delta);
        }
        const float ws = w * p.scale;
#pragma unroll
        for (int c0 = 0; c0 < 80; c0 += 16) {
          uint32_t a[16], pk[8];
          tmem_ld16(s_addr + 80 + c0, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_half2(ws * __uint_as_float(s[c0 + 2 * i]) * (__uint_as_float(a[2 * i]) - delta),
                               ws * __uint_as_float(s[c0 + 2 * i + 1]) * (__uint_as_float(a[2 * i + 1]) - delta));
          tmem_st8(s_addr + (c0 >> 1), pk);
        }
      } else {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c0 = 0; c0 < 40; c0 += 8) tmem_st8(s_addr + c0, z);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready[t % NBUF]);
      if (slot >= 2 && live) {
        float val = mk ? (delta - delta_uc) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) atomicAdd(&dcoef_s[slot - 2], val);
      }
      if (t == 0) {  // dQ_u is complete as soon as the (single) unconditional task's dS has been consumed
        ok = mbar_wait_warp(&dqu_full, 0, &dead, p.err, 31);
        if (ok) {
          tc_fence_after();
          write_dq(pr, Cfg::TMEM_DQ, sDOu, 2);  // the dA_u tile is dead once dP_u has been formed
        }
      }
      if (ok && t < 2) {
        ok = mbar_wait_warp(&list_ready, 0, &dead, p.err, 34);
        if (ok) T = n_tiles_s;
      }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (g == NBUF - 1) {  // the warpgroup that owns task 1 stores the conditional row
      if (ok) ok = mbar_wait_warp(&dqc_full, 0, &dead, p.err, 33);
      if (ok) {
        tc_fence_after();
        write_dq(pr + B, Cfg::TMEM_DQ + DMMA, sDOc, 3);  // every MMA has completed: dO_c is dead
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < n_obj && dcoef_s[tid] != 0.f) atomicAdd(&p.d_coef[pr * n_obj + tid], dcoef_s[tid]);
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

template <int D>
static int launch_xattn_bwd(const sta_xattn_bwd_args* a, cudaStream_t stream) {
  using Cfg = XattnBwdCfg<D>;
  CUtensorMap tm_q, tm_do, tm_k, tm_v, tm_dq;
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->n, (uint64_t)a->prompts * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->q_token_stride * 2, (uint64_t)a->q_batch_stride * 2};
    int rc = make_tmap_f16(&tm_q, a->q, 4, dims, st, box);
    if (rc) return rc;
    const uint64_t st2[4] = {2, (uint64_t)D * 2, (uint64_t)a->do_token_stride * 2, (uint64_t)a->do_batch_stride * 2};
    rc = make_tmap_f16(&tm_do, a->d_out, 4, dims, st2, box);
    if (rc) return rc;
    const uint64_t C = (uint64_t)a->heads * D;
    const uint64_t st3[4] = {2, (uint64_t)D * 2, C * 2, C * 2 * (uint64_t)a->n};  // d_q is dense [2B, n, heads*D]
    rc = make_tmap_f16(&tm_dq, a->d_q, 4, dims, st3, box);
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
  p.out = reinterpret_cast<const __half*>(a->out);
  p.d_q = reinterpret_cast<__half*>(a->d_q);
  p.d_coef = a->d_coef;
  p.prompts = a->prompts;
  p.n = a->n;
  p.heads = a->heads;
  p.n_obj = a->n_obj;
  p.ctx_len = a->ctx_len;
  p.o_token_stride = a->o_token_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static PerDeviceOnce smem_attr;
  int rc = smem_attr.run([] {
    return cudaFuncSetAttribute(xattn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  if (rc) return rc;
  if (a->n_obj > 0)
    STA_CUDA_CHECK(cudaMemsetAsync(a->d_coef, 0, sizeof(float) * a->prompts * a->n_obj, stream));
  dim3 grid((a->n + 127) / 128, a->heads, a->prompts);
  xattn_bwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_q, tm_do, tm_k, tm_v, tm_dq, p);
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
  if (a->n_obj > 0 && (!a->mask || !a->coef || !a->d_coef || !a->out))
    return fail(STA_ERR_BAD_ARG, "sta_xattn_bwd: mask/coef/d_coef/out required when n_obj > 0");
  if (a->ctx_len < 1 || a->ctx_len > 80) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: ctx_len %d not in [1,80]", a->ctx_len);
  if (reinterpret_cast<uintptr_t>(a->d_q) & 15) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: d_q must be 16-byte aligned");
  if (a->n_obj > 0 && ((a->o_token_stride % 8) || (a->o_batch_stride % 8) || (reinterpret_cast<uintptr_t>(a->out) & 15)))
    return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: out rows must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (a->head_dim) {
    case 40: return launch_xattn_bwd<40>(a, s);
    case 80: return launch_xattn_bwd<80>(a, s);
    case 160: return launch_xattn_bwd<160>(a, s);
    default:
      return fail(STA_ERR_UNSUPPORTED, "sta_xattn_bwd: head_dim %d not built (SD-v1 uses 40/80/160)", a->head_dim);
  }
}
