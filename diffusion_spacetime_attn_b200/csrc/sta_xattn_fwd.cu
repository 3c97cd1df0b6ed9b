// sta_xattn_fwd.cu — fused dual (global + per-object local) cross-attention with the mask-gated alpha-blend.
//
// Replaces, in ONE kernel and before `to_out`, what the reference does with 1 + n_obj separate attn2 calls and
// ~5 elementwise launches per object (ldm/modules/attention.py:278-294):
//     out[p]     = A_u                                            (unconditional half, context slot 0)
//     out[B + p] = A_g + sum_i m_i[pix] * c_i * (A_i - A_u)       (conditional half, slots 1 and 2+i)
// A_x = softmax(q k_x^T * scale) v_x over the 77 keys of context x.  The blend is linear in the attention
// outputs, so it is folded into the P operand: P_i rows are scaled by w_i = m_i c_i / l_i and every context of
// the conditional row accumulates into ONE TMEM accumulator; the unconditional accumulator is subtracted with
// weight sum_i w_i in the fp32 epilogue.
//
// The op is microseconds of work per launch (1.6 GFLOP / 11 MB at the 64x64 level): it is bound by the latency
// chain TMA -> MMA -> softmax -> MMA -> store and by the MUFU pipe (one ex2 per score), not by the tensor pipe.
// Round-2 layout, built for that:
//   CTA = one 128-pixel tile of one (prompt, head); 9 warps:
//     warp 0      control: barrier init, ALL TMA loads (issued before anything else: Q_u, K_0, Q_c, K_1, V_0, V_1,
//                 then the live objects), MMA issue (S = Q K^T SS, O += P V TS with P packed fp16 in TMEM and V
//                 MN-major), ring refills
//     warps 1..4  softmax warpgroup 0: contexts ("tasks") 0, 2, 4, ...  + the unconditional row's epilogue
//     warps 5..8  softmax warpgroup 1: tasks 1, 3, 5, ...               + the conditional row's epilogue
//   task 0 = (Q_u, unconditional context), task 1 = (Q_c, global context), tasks 2.. = (Q_c, local context of an
//   object whose mask is non-empty inside this pixel tile).  Task t uses S buffer t & 1: the two warpgroups run their
//   softmaxes concurrently and QK^T of task t + 2 is issued right behind P V of task t (tcgen05.mma executes in
//   issue order, so the S buffer that P_t aliases is safe).
//   TMEM: S0 [0,80) S1 [80,160) (P aliases the first 40 columns of its S), O_u at 160, O_c at 160 + DMMA: 256
//   columns at head dim 40, so TWO CTAs are resident per SM there (288 threads, <= 112 registers, 93 KB smem) and
//   the 256 CTAs of the 64x64 level run as a single wave.
//   A warp whose 32 pixels are all outside object i's mask skips that softmax (P = 0); LSE is written only for
//   the (context, pixel) pairs that were evaluated (the backward applies the same rule).
#include <mutex>

#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

constexpr int kXQBlockBytes = 128 * 128;  // Q block: 128 rows x 64 columns
constexpr int kXCBlockBytes = 80 * 128;   // context block: 80 rows x 64 columns
constexpr int kXMaxObj = 8;

template <int D>
struct XattnCfg {
  static constexpr int DMMA = (D + 15) / 16 * 16;
  static constexpr int NBLK = (D + 63) / 64;
  static constexpr int ST = (NBLK == 1) ? 3 : (NBLK == 2 ? 4 : 2);  // K and V ring depth
  static constexpr int QTILE = NBLK * kXQBlockBytes;
  static constexpr int CTILE = NBLK * kXCBlockBytes;
  static constexpr int SMEM_BYTES = 2 * QTILE + 2 * ST * CTILE + 1024;
  static constexpr int THREADS = 288;
  static constexpr int TMEM_O = 160;
  static constexpr int TMEM_COLS = (TMEM_O + 2 * DMMA <= 256) ? 256 : 512;
  static constexpr int MIN_CTAS = (TMEM_COLS == 256) ? 2 : 1;
};

// Debug-only CTA timeline (tools/xattn_timeline.py builds a separate library with -DSTA_TIMELINE; never in the product).
#ifdef STA_TIMELINE
static long long* g_timeline_fwd = nullptr;
#define STA_TL(i)                                                                                     \
  do {                                                                                                \
    if (p.tl && lane == 0)                                                                            \
      p.tl[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 20 + (i)] = clock64(); \
  } while (0)
#else
#define STA_TL(i) \
  do {            \
  } while (0)
#endif

struct XattnFwdParams {
#ifdef STA_TIMELINE
  long long* tl;
#endif
  const uint8_t* mask;  // [B, n_obj, n]
  const float* coef;    // [B, n_obj]
  __half* out;
  float* lse;  // [B, heads, 2+n_obj, n] or null
  int prompts, n, heads, n_obj, ctx_len;
  long long o_token_stride, o_batch_stride;
  float scale_log2;
  unsigned int* err;
};

template <int D>
__global__ void __launch_bounds__(XattnCfg<D>::THREADS, XattnCfg<D>::MIN_CTAS)
xattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                 const XattnFwdParams p) {
  using Cfg = XattnCfg<D>;
  constexpr int ST = Cfg::ST, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;  // Q_u tile, Q_c tile
  unsigned char* sK = sQ + 2 * Cfg::QTILE;
  unsigned char* sV = sK + ST * Cfg::CTILE;

  __shared__ uint64_t qu_full, qc_full, k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];
  __shared__ uint64_t s_full[2], p_ready[2], ou_full, o_full, list_ready;
  __shared__ unsigned int tile_bits_s;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  __shared__ int tile_slot[2 + kXMaxObj];  // context slot of the t-th task of this CTA
  __shared__ int n_tiles_s;

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, pr = blockIdx.z;
  const int n = p.n, B = p.prompts, n_obj = p.n_obj, n_slots = 2 + p.n_obj;
  if (warp == 0) STA_TL(0);
#ifdef STA_TIMELINE
  if (p.tl && tid == 0) {
    unsigned long long gt;
    unsigned int smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long* e = p.tl + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 20;
    e[16] = (long long)gt;
    e[17] = smid;
  }
#endif

  auto load_k = [&](int t, int slot) {
    const int st = t % ST;
    mbar_expect_tx_w(&k_full[st], Cfg::CTILE);
    for (int blk = 0; blk < NBLK; ++blk)
      tma_load_4d_w(sK + (st * NBLK + blk) * kXCBlockBytes, &tm_k, &k_full[st], blk * 64, h, 0, pr * n_slots + slot);
  };
  auto load_v = [&](int t, int slot) {
    const int st = t % ST;
    mbar_expect_tx_w(&v_full[st], Cfg::CTILE);
    for (int blk = 0; blk < NBLK; ++blk)
      tma_load_4d_w(sV + (st * NBLK + blk) * kXCBlockBytes, &tm_v, &v_full[st], blk * 64, h, 0, pr * n_slots + slot);
  };

  if (warp == 0) {
    if (lane == 0) {
      dead = 0;
      tile_bits_s = 0;
      mbar_init(&qu_full, 1);
      mbar_init(&qc_full, 1);
      mbar_init(&list_ready, 1);
      for (int i = 0; i < ST; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); }
      mbar_init(&ou_full, 1);
      mbar_init(&o_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    // every load that does not depend on the masks goes out before anything else happens in this CTA
    mbar_expect_tx_w(&qu_full, Cfg::QTILE);
    for (int blk = 0; blk < NBLK; ++blk) tma_load_4d_w(sQ + blk * kXQBlockBytes, &tm_q, &qu_full, blk * 64, h, q0, pr);
    load_k(0, 0);
    mbar_expect_tx_w(&qc_full, Cfg::QTILE);
    for (int blk = 0; blk < NBLK; ++blk)
      tma_load_4d_w(sQ + (NBLK + blk) * kXQBlockBytes, &tm_q, &qc_full, blk * 64, h, q0, pr + B);
    load_k(1, 1);  // V_0 / V_1 follow the first two QK^T issues: the operands on the critical path land first
    STA_TL(1);
  } else if (warp == 1) {
    tmem_alloc(&tmem_base_s, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();  // barriers initialised, TMEM address published: nothing slow (no global load) sits in front of it
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) STA_TL(2);

  if (warp == 0) {
    // ===================================== control warp: TMA + MMA issue =====================================
    // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
    constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);
    constexpr uint64_t vdesc_hi = umma_desc_hi_sw128(kXCBlockBytes, 1024);
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, 80, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, DMMA, 0, 1);
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);

    auto issue_qk = [&](int t) {  // S[t & 1] = Q_(t ? c : u) K_t^T
      const int st = t % ST, r = t ? 1 : 0;
#pragma unroll
      for (int k = 0; k < DMMA / 16; ++k) {
        const uint32_t qoff = (k / 4) * kXQBlockBytes + (k % 4) * 32;
        const uint32_t koff = (k / 4) * kXCBlockBytes + (k % 4) * 32;
        umma_ss_w(tmem + (t & 1) * 80, umma_desc(kdesc_hi, q_addr + r * Cfg::QTILE + qoff),
                  umma_desc(kdesc_hi, k_addr + st * Cfg::CTILE + koff), idesc_qk, k > 0);
      }
      umma_commit_w(&s_full[t & 1]);
      umma_commit_w(&k_empty[st]);
    };
    auto issue_pv = [&](int t) {  // O_(t ? c : u) (+)= P_t V_t
      const int st = t % ST, r = t ? 1 : 0;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        umma_ts_w(tmem + Cfg::TMEM_O + r * DMMA, tmem + (t & 1) * 80 + k * 8,
                  umma_desc(vdesc_hi, v_addr + st * Cfg::CTILE + k * 2048), idesc_pv, (t >= 2) || k > 0);
      umma_commit_w(&v_empty[st]);
    };

    bool ok = mbar_wait_warp(&qu_full, 0, &dead, p.err, 20) && mbar_wait_warp(&k_full[0], 0, &dead, p.err, 21);
    STA_TL(3);
    if (ok) {
      tc_fence_after();
      issue_qk(0);
      STA_TL(18);
      ok = mbar_wait_warp(&qc_full, 0, &dead, p.err, 22) && mbar_wait_warp(&k_full[1], 0, &dead, p.err, 23);
    }
    if (ok) {
      tc_fence_after();
      issue_qk(1);
      load_v(0, 0);
      load_v(1, 1);
      // the task list (which objects touch this tile) is built by warp 2 while the first loads are in flight
      ok = mbar_wait_warp(&list_ready, 0, &dead, p.err, 28);
    }
    const int T = ok ? n_tiles_s : 0;  // task 0 -> unconditional row, tasks 1..T-1 -> conditional row
    int k_next = T < ST ? T : ST, v_next = k_next;
    for (int t = 2; t < k_next; ++t) { load_k(t, tile_slot[t]); load_v(t, tile_slot[t]); }
    for (int t = 0; t < T && ok; ++t) {
      // ring refills whose predecessor MMA was issued at least one iteration ago (their completion is imminent)
      while (ok && k_next < T && k_next - ST <= t + 1) {
        ok = mbar_wait_warp(&k_empty[k_next % ST], (k_next / ST - 1) & 1, &dead, p.err, 10);
        if (ok) load_k(k_next, tile_slot[k_next]);
        ++k_next;
      }
      while (ok && v_next < T && v_next - ST <= t - 1) {
        ok = mbar_wait_warp(&v_empty[v_next % ST], (v_next / ST - 1) & 1, &dead, p.err, 11);
        if (ok) load_v(v_next, tile_slot[v_next]);
        ++v_next;
      }
      if (!ok) break;
      ok = mbar_wait_warp(&p_ready[t & 1], (t >> 1) & 1, &dead, p.err, 24) &&
           mbar_wait_warp(&v_full[t % ST], (t / ST) & 1, &dead, p.err, 25);
      if (!ok) break;
      tc_fence_after();
      if (t == 0) STA_TL(4);
      if (t == T - 1) STA_TL(5);
      issue_pv(t);
      if (t == 0) umma_commit_w(&ou_full);
      if (t == T - 1) umma_commit_w(&o_full);
      if (t + 2 < T) {
        ok = mbar_wait_warp(&k_full[(t + 2) % ST], ((t + 2) / ST) & 1, &dead, p.err, 26);
        if (!ok) break;
        tc_fence_after();
        issue_qk(t + 2);
      }
    }
  } else {
    // ===================================== softmax warpgroups / epilogue =====================================
    const int g = (warp - 1) >> 2;  // warpgroup: 0 = even tasks + unconditional output, 1 = odd tasks + conditional output
    const int quad = warp & 3;      // TMEM lane quadrant this warp may access
    const int row = q0 + (quad << 5) + lane;
    const bool row_ok = row < n;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad << 5) << 16);
    const uint32_t s_addr = lane_addr + g * 80;

    // object membership of this pixel (bit i) and sigma = sum_i m_i c_i
    unsigned int bits = 0;
    float sigma = 0.f;
#pragma unroll
    for (int i = 0; i < kXMaxObj; ++i) {
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
    int T = 2;
    for (int t = g; t < T; t += 2) {
      // tasks 0 / 1 (unconditional / global context) always exist; the list is only needed beyond them
      const int slot = t >= 2 ? tile_slot[t] : t;
      float w = 1.f;
      bool live = true;
      if (slot >= 2) {
        const bool mk = (bits >> (slot - 2)) & 1u;
        w = mk ? p.coef[pr * n_obj + slot - 2] : 0.f;
        live = __any_sync(0xffffffffu, mk);
      }
      ok = mbar_wait_warp(&s_full[g], (t >> 1) & 1, &dead, p.err, 30);
      if (!ok) break;
      tc_fence_after();
      if (warp == 1 && t == 0) STA_TL(6);
      if (warp == 5 && t == 1) STA_TL(7);
      if (live) {
        uint32_t s[80];
        tmem_ld32(s_addr, s);
        tmem_ld32(s_addr + 32, s + 32);
        tmem_ld16(s_addr + 64, s + 64);
        tmem_ld_wait();
        const int valid = p.ctx_len;
        if (valid >= 76) {  // CLIP's 77 tokens: only the last columns are padding
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
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 80; c += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
        }
        const float m = fmaxf(mx0, mx1) * p.scale_log2;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c = 0; c < 80; c += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[c]), p.scale_log2, -m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, -m));
          l0 += p0;
          l1 += p1;
          s[c] = __float_as_uint(p0);
          s[c + 1] = __float_as_uint(p1);
        }
        const float l = l0 + l1;
        const float f = __fdividef(w, l);
#pragma unroll
        for (int c0 = 0; c0 < 80; c0 += 16) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_half2(__uint_as_float(s[c0 + 2 * i]) * f, __uint_as_float(s[c0 + 2 * i + 1]) * f);
          tmem_st8(s_addr + (c0 >> 1), pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
        if (warp == 1 && t == 0) STA_TL(8);
        if (warp == 5 && t == 1) STA_TL(9);
        if (p.lse && row_ok)
          p.lse[(((long long)pr * p.heads + h) * n_slots + slot) * n + row] = (m + log2f(l)) * 0.6931471805599453f;
      } else {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c0 = 0; c0 < 40; c0 += 8) tmem_st8(s_addr + c0, z);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
      }
      if (t < 2) {  // first task done: learn how many tasks this tile has
        ok = mbar_wait_warp(&list_ready, 0, &dead, p.err, 33);
        if (!ok) break;
        T = n_tiles_s;
      }
    }
    // -------- epilogue: warpgroup 0 stores the unconditional row, warpgroup 1 the conditional row --------
    ok = __all_sync(0xffffffffu, ok);
    if (ok) ok = g == 0 ? mbar_wait_warp(&ou_full, 0, &dead, p.err, 31) : mbar_wait_warp(&o_full, 0, &dead, p.err, 32);
    if (warp == 1) STA_TL(10);
    if (warp == 5) STA_TL(11);
    if (ok) {
      tc_fence_after();
      // O (fp32, TMEM) -> fp16 -> the warpgroup's own (dead) Q tile in the TMA swizzled layout -> one TMA store per 64
      // channels: full-line writes instead of 16 bytes per 2*C-byte-strided row per thread
      const uint32_t ou_addr = lane_addr + Cfg::TMEM_O;
      const uint32_t oc_addr = ou_addr + DMMA;
      unsigned char* stage = sQ + g * Cfg::QTILE;
      const int r = (quad << 5) + lane;
#pragma unroll
      for (int c0 = 0; c0 < DMMA; c0 += 16) {
        uint32_t u[16], c[16];
        tmem_ld16(ou_addr + c0, u);
        if (g == 1) tmem_ld16(oc_addr + c0, c);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = g == 0 ? __uint_as_float(u[hh * 8 + i])
                          : fmaf(-sigma, __uint_as_float(u[hh * 8 + i]), __uint_as_float(c[hh * 8 + i]));
          uint4 v;
          v.x = pack_half2(o[0], o[1]);
          v.y = pack_half2(o[2], o[3]);
          v.z = pack_half2(o[4], o[5]);
          v.w = pack_half2(o[6], o[7]);
          const int ch = (c0 >> 3) + hh;  // 16-byte chunk index along the head dim
          *reinterpret_cast<uint4*>(stage + (ch >> 3) * kXQBlockBytes + sw128_offset(r, ch & 7)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(2 + g, 128);
      if (quad == 0 && lane == 0) {  // one thread issues, commits and waits (bulk groups are per thread)
        for (int blk = 0; blk < NBLK; ++blk) tma_store_4d(&tm_o, stage + blk * kXQBlockBytes, blk * 64, h, q0, pr + g * B);
        bulk_commit_group();
        bulk_wait_group_read0();
      }
    }
  }
  if (warp == 1) STA_TL(12);
  if (warp == 5) STA_TL(13);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
  if (warp == 1) STA_TL(14);
#ifdef STA_TIMELINE
  if (p.tl && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.tl[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 20 + 15] = (long long)gt;
  }
#endif
}

template <int D>
static int launch_xattn_fwd(const sta_xattn_fwd_args* a, cudaStream_t stream) {
  using Cfg = XattnCfg<D>;
  CUtensorMap tm_q, tm_k, tm_v, tm_o;
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->n, (uint64_t)a->prompts * 2};
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->q_token_stride * 2, (uint64_t)a->q_batch_stride * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_f16(&tm_q, a->q, 4, dims, st, box);
    if (rc) return rc;
    const uint64_t sto[4] = {2, (uint64_t)D * 2, (uint64_t)a->o_token_stride * 2, (uint64_t)a->o_batch_stride * 2};
    rc = make_tmap_f16(&tm_o, a->out, 4, dims, sto, box);
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
  XattnFwdParams p;
#ifdef STA_TIMELINE
  p.tl = g_timeline_fwd;
#endif
  p.mask = a->mask;
  p.coef = a->coef;
  p.out = reinterpret_cast<__half*>(a->out);
  p.lse = a->lse;
  p.prompts = a->prompts;
  p.n = a->n;
  p.heads = a->heads;
  p.n_obj = a->n_obj;
  p.ctx_len = a->ctx_len;
  p.o_token_stride = a->o_token_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static PerDeviceOnce smem_attr;
  int rc = smem_attr.run([] {
    return cudaFuncSetAttribute(xattn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  if (rc) return rc;
  dim3 grid((a->n + 127) / 128, a->heads, a->prompts);
  xattn_fwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_q, tm_k, tm_v, tm_o, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta

#ifdef STA_TIMELINE
extern "C" void sta_debug_timeline_fwd(long long* dev) { sta::g_timeline_fwd = dev; }
#endif

extern "C" int sta_xattn_fwd(const sta_xattn_fwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->q || !a->k_ctx || !a->v_ctx || !a->out) return fail(STA_ERR_BAD_ARG, "sta_xattn_fwd: null pointer");
  if (a->prompts < 1 || a->n < 1 || a->heads < 1) return fail(STA_ERR_BAD_ARG, "sta_xattn_fwd: empty shape");
  if (a->n_obj < 0 || a->n_obj > kXMaxObj) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_fwd: n_obj %d not in [0,%d]", a->n_obj, kXMaxObj);
  if (a->n_obj > 0 && (!a->mask || !a->coef)) return fail(STA_ERR_BAD_ARG, "sta_xattn_fwd: mask/coef required when n_obj > 0");
  if (a->ctx_len < 1 || a->ctx_len > 80) return fail(STA_ERR_UNSUPPORTED, "sta_xattn_fwd: ctx_len %d not in [1,80]", a->ctx_len);
  if ((a->o_token_stride % 8) || (a->o_batch_stride % 8) || (reinterpret_cast<uintptr_t>(a->out) & 15))
    return fail(STA_ERR_UNSUPPORTED, "sta_xattn_fwd: out rows must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (a->head_dim) {
    case 40: return launch_xattn_fwd<40>(a, s);
    case 80: return launch_xattn_fwd<80>(a, s);
    case 160: return launch_xattn_fwd<160>(a, s);
    default:
      return fail(STA_ERR_UNSUPPORTED, "sta_xattn_fwd: head_dim %d not built (SD-v1 uses 40/80/160)", a->head_dim);
  }
}
