// sta_sattn_wide.cu — flash attention with a 512-wide head for sm_100a (tcgen05 + TMEM + TMA): the mid-block
// attention of the KL-VAE decoder.
//
// Replaces AttnBlock.forward of the reference, ldm/modules/diffusionmodules/model.py:150-202 — ONE head of
// d = C = 512 over N = (H/8)^2 tokens (4096 at 512^2, 9216 at 768^2): w = softmax(q k^T / sqrt(C)); h = w v —
// without writing the [N, N] score matrix (32 MB fp16 + 64 MB fp32 at N = 4096, 1 GB at N = 9216 x 2 prompts).
//
// Why this is not an instantiation of sta_sattn_fwd / _bwd: at d = 512 one 128-row operand tile is 128 KB, so only ONE
// full-width tile can live in shared memory, and a 128 x 512 fp32 accumulator is the whole of TMEM.  Hence
//   * every other operand is STREAMED in 16 KB blocks (128 rows x 64 columns, the TMA / UMMA 128B-swizzle atom) through
//     one mbarrier ring, in exactly the order the MMA warp consumes them; a 512-deep contraction is 8 blocks x 4 MMAs;
//   * the output columns are split over CTAs (grid.y = head x column slice); each slice recomputes the scores.
//
// forward   CTA = (128 queries, 256 output columns): Q resident; per key tile j the ring carries K_j (8 blocks) and
//           V_j[:, slice] (4 blocks).  S = Q K_j^T (SS) -> online softmax, one thread per row, lazy rescale -> P packed
//           fp16 in its own TMEM columns -> O += P V_j (TS, two N = 128 MMA groups over adjacent ring slots).
//           S(j+1) is issued as soon as the softmax warps have READ S(j), so it overlaps the exponentials.
// backward  ONE launch, three CTA roles (dK, dQ, dV), CTA = (128-row tile, 256 gradient columns): see the comment above
//           WideBwdCfg.  No atomics, no fp32 accumulator in HBM: each gradient element is owned by one CTA.  Price: the
//           scores are recomputed per column slice and per role.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

constexpr int kWD = 512;                 // head dim = contraction depth of the score GEMMs
constexpr int kWBlk = 128 * 128;         // one ring block: 128 rows x 64 fp16 (128B swizzle)
constexpr int kWNB = kWD / 64;           // blocks per full-width tile

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
// DV = output columns per CTA.  256 (two slices, the scores computed twice) is the efficient shape once the grid fills the
// GPU; one 64 x 64 latent gives only 64 such CTAs, so the launcher picks DV = 128 (128 CTAs, scores computed four times
// but on otherwise idle SMs) whenever the 256-column grid would leave SMs empty.
template <int DV_>
struct WideFwdCfg {
  static constexpr int DV = DV_;
  static constexpr int NSL = kWD / DV;        // column slices
  static constexpr int NVB = DV / 64;         // V blocks per key tile
  static constexpr int RING = 6;              // even: a pair of adjacent blocks (N = 128 operand) never wraps
  static constexpr int SMEM = (kWNB + RING) * kWBlk + 1024;
  static constexpr int TMEM_P = 128, TMEM_O = 256;  // S [0,128)  P [128,192)  O [256,512)
  static constexpr int THREADS = 64 + 128;
};

struct WideFwdParams {
  __half* out;
  float* lse;
  int n, heads;
  long long o_token_stride, o_batch_stride;
  float scale_log2;
  unsigned int* err;
};

template <int DV_>
__global__ void __launch_bounds__(WideFwdCfg<DV_>::THREADS, 1)
sattn_wide_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const WideFwdParams p) {
  using Cfg = WideFwdCfg<DV_>;
  constexpr int RING = Cfg::RING, DV = Cfg::DV;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sR = sQ + kWNB * kWBlk;

  __shared__ uint64_t q_full, r_full[RING], r_empty[RING], s_full, s_consumed, p_ready, pv_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y / Cfg::NSL, slice = blockIdx.y % Cfg::NSL, b = blockIdx.z;
  const int n = p.n;
  const int T = (n + 127) / 128;

  if (tid == 0) {
    dead = 0;
    mbar_init(&q_full, 1);
    for (int i = 0; i < RING; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 1); }
    mbar_init(&s_full, 1);
    mbar_init(&s_consumed, 4);
    mbar_init(&p_ready, 4);
    mbar_init(&pv_done, 1);
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
    mbar_expect_tx_w(&q_full, kWNB * kWBlk);
    for (int blk = 0; blk < kWNB; ++blk) tma_load_4d_w(sQ + blk * kWBlk, &tm_q, &q_full, blk * 64, h, q0, b);
    int idx = 0;
    bool ok = true;
    auto push = [&](const CUtensorMap* tm, int col, int row) {
      const int slot = idx % RING;
      if (ok) ok = mbar_wait_warp(&r_empty[slot], ((idx / RING) & 1) ^ 1, &dead, p.err, 10);
      if (ok) {
        mbar_expect_tx_w(&r_full[slot], kWBlk);
        tma_load_4d_w(sR + slot * kWBlk, tm, &r_full[slot], col, h, row, b);
      }
      ++idx;
    };
    // consumption order of the MMA warp: K(0), then per key tile j: K(j+1), V(j)
    for (int blk = 0; blk < kWNB; ++blk) push(&tm_k, blk * 64, 0);
    for (int j = 0; j < T && ok; ++j) {
      if (j + 1 < T)
        for (int blk = 0; blk < kWNB; ++blk) push(&tm_k, blk * 64, (j + 1) * 128);
      for (int vb = 0; vb < Cfg::NVB; ++vb) push(&tm_v, slice * DV + vb * 64, j * 128);
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);       // K-major operands
    constexpr uint64_t vdesc_hi = umma_desc_hi_sw128(kWBlk, 1024);    // MN-major V: 64-column groups one block apart
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 128, 0, 1);
    const uint32_t q_addr = smem_u32(sQ), r_addr = smem_u32(sR);
    int cidx = 0;
    bool ok = mbar_wait_warp(&q_full, 0, &dead, p.err, 20);
    auto issue_qk = [&]() {  // S = Q K^T: one ring block per 64 columns of the contraction
      for (int blk = 0; blk < kWNB && ok; ++blk) {
        const int slot = cidx % RING;
        ok = mbar_wait_warp(&r_full[slot], (cidx / RING) & 1, &dead, p.err, 21);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_w(tmem, umma_desc(kdesc_hi, q_addr + blk * kWBlk + k * 32),
                    umma_desc(kdesc_hi, r_addr + slot * kWBlk + k * 32), idesc_qk, blk > 0 || k > 0);
        umma_commit_w(&r_empty[slot]);
        ++cidx;
      }
      if (ok) umma_commit_w(&s_full);
    };
    auto issue_pv = [&](bool acc) {  // O[:, 128 g .. 128 g + 128) += P V: two adjacent ring blocks per MMA group
      for (int g = 0; g < Cfg::NVB / 2 && ok; ++g) {
        const int slot = cidx % RING;  // even
        ok = mbar_wait_warp(&r_full[slot], (cidx / RING) & 1, &dead, p.err, 22) &&
             mbar_wait_warp(&r_full[slot + 1], (cidx / RING) & 1, &dead, p.err, 23);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts_w(tmem + Cfg::TMEM_O + g * 128, tmem + Cfg::TMEM_P + k * 8,
                    umma_desc(vdesc_hi, r_addr + slot * kWBlk + k * 2048), idesc_pv, acc || k > 0);
        umma_commit_w(&r_empty[slot]);
        umma_commit_w(&r_empty[slot + 1]);
        cidx += 2;
      }
      if (ok) umma_commit_w(&pv_done);
    };
    if (ok) issue_qk();
    for (int j = 0; j < T && ok; ++j) {
      if (j + 1 < T) {  // the softmax warps hold S(j) in registers: the tensor core may overwrite it
        ok = mbar_wait_warp(&s_consumed, j & 1, &dead, p.err, 24);
        if (!ok) break;
        tc_fence_after();
        issue_qk();
        if (!ok) break;
      }
      ok = mbar_wait_warp(&p_ready, j & 1, &dead, p.err, 25);
      if (!ok) break;
      tc_fence_after();
      issue_pv(j > 0);
    }
  } else {
    // ===================================== softmax / epilogue ================================
    const int row_in_tile = ((warp & 3) << 5) + lane;
    const int row = q0 + row_in_tile;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t p_addr = lane_addr + Cfg::TMEM_P, o_addr = lane_addr + Cfg::TMEM_O;
    float m_ref = -INFINITY, l = 0.f;
    bool ok = true;
    for (int j = 0; j < T; ++j) {
      ok = mbar_wait_warp(&s_full, j & 1, &dead, p.err, 30);
      if (!ok) break;
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) tmem_ld32(lane_addr + c0, s + c0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_consumed);
      const int valid = n - j * 128;  // keys of this tile that exist
      if (valid < 128) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= valid) s[c] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[c]));
        mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      if (j == 0) {
        m_ref = mx;
      } else {
        // lazy rescale: keep the old reference unless the new maximum exceeds it by more than 2^8
        const bool need = mx > m_ref + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          ok = mbar_wait_warp(&pv_done, (j - 1) & 1, &dead, p.err, 32);  // O must be quiescent
          if (!ok) break;
          tc_fence_after();
          const float f = need ? fast_exp2(m_ref - mx) : 1.f;
          if (need) m_ref = mx;
          l *= f;
          for (int ch = 0; ch < DV / 8; ++ch) {
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
      for (int c = 0; c < 128; c += 2) {  // exp in place: the packed pair (c, c+1) lands in s[c/2]
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[c]), p.scale_log2, -m_ref));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, -m_ref));
        l0 += p0;
        l1 += p1;
        s[c >> 1] = pack_half2(p0, p1);
      }
      if (j > 0) {  // the P buffer is free once P(j-1) V(j-1) has completed
        ok = mbar_wait_warp(&pv_done, (j - 1) & 1, &dead, p.err, 33);
        if (!ok) break;
        tc_fence_after();
      }
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) tmem_st16(p_addr + c0, s + c0);
      l += l0 + l1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready);
    }
    // -------- epilogue: O / l -> fp16, natural-log LSE --------
    ok = __all_sync(0xffffffffu, ok);
    if (ok) ok = mbar_wait_warp(&pv_done, (T - 1) & 1, &dead, p.err, 31);
    if (ok) {
      tc_fence_after();
      const float inv = 1.f / l;
      __half* orow = p.out + (long long)b * p.o_batch_stride + (long long)row * p.o_token_stride + h * kWD + slice * DV;
      for (int ch = 0; ch < DV / 8; ++ch) {
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
      if (p.lse && row < n && slice == 0)
        p.lse[((long long)b * p.heads + h) * n + row] = (m_ref + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
// ONE launch, three CTA roles (blockIdx.y = role x head x column slice); a CTA owns a 128-row tile and 256 gradient columns
// and loops over the other side's tiles t:
//   role K (rows = keys j,    X = K_j resident, Y = Q, U = V_j,  W = dO):  S' = K_j Q_t^T, dP' = V_j dO_t^T, dK[:, c] += dS' Q_t[:, c]
//   role Q (rows = queries i, X = Q_i resident, Y = K, U = dO_i, W = V):   S' = Q_i K_t^T, dP' = dO_i V_t^T, dQ[:, c] += dS' K_t[:, c]
//   role V (rows = keys j,    X = K_j resident, Y = Q, W = dO):            S' = K_j Q_t^T,                  dV[:, c] += P' dO_t[:, c]
// P' = exp2(S' scale log2e - lse_query), dS' = scale P' (dP' - delta_query).  dS' goes through a 128B-swizzled smem tile
// (K-major A operand, SS MMA); P' of role V is packed fp16 over the S' columns (A operand from TMEM, TS MMA) and needs no dP'
// at all — which is why dV has its own role: it streams 192 KB per tile pair instead of 448 KB.
// Measured on B200 (N = 4096): the first version (two roles, 128 columns per CTA, two launches) took 415 us and moved
// 3.6 GB from L2 to the SMs at ~9 TB/s — the kernel is bound by the chip's L2 throughput (~6300 B/clk), not by the tensor
// pipe, so the lever is bytes streamed per gradient column: 256-column CTAs and the dP'-free dV role cut them by 39 %.
enum { ROLE_K = 0, ROLE_Q = 1, ROLE_V = 2 };

struct WideBwdCfg {
  static constexpr int NACC = 256;            // gradient columns per CTA
  static constexpr int NSL = kWD / NACC;
  static constexpr int NCB = NACC / 64;       // streamed column-slice blocks per tile
  static constexpr int RING_DS = 4;           // roles K / Q: X tile (128 KB) + dS' tile (32 KB) + 4 ring blocks
  static constexpr int RING_V = 6;            // role V: no dS' tile
  static constexpr int SMEM = (kWNB + 2 + RING_DS) * kWBlk + 1024;
  static constexpr int TMEM_DP = 128, TMEM_ACC = 256;  // S' [0,128) (role V: P' packed over it)  dP' [128,256)  acc [256,512)
  static constexpr int THREADS = 64 + 128;
};
static_assert((kWNB + WideBwdCfg::RING_V) * kWBlk + 1024 <= WideBwdCfg::SMEM, "role V ring");

struct WideBwdParams {
  const float* lse;    // [b, h, n]
  const float* delta;  // [b, h, n]
  __half* d_q;
  __half* d_k;
  __half* d_v;         // [b, n, h*512] fp16, token stride d_tok
  long long d_tok;
  int n, heads;
  float scale, scale_log2;
  unsigned int* err;
};

struct WideBwdShared {
  uint64_t x_full, r_full[WideBwdCfg::RING_V], r_empty[WideBwdCfg::RING_V];
  uint64_t s_full[2], s_consumed, dp_full, pds_ready, acc_done, acc_full;
  uint32_t tmem_base;
  int dead;
  // statistics of the streamed query tile (roles K / V).  Single-buffered: 224 KB of operand tiles + the alignment slack
  // leave ~2 KB of the 227 KB for everything static.
  alignas(16) float lse2[128];  // read four at a time (broadcast float4 loads)
  alignas(16) float dl[128];
};

// Software pipeline (the first version ran  S', dP' -> math -> acc  strictly in sequence: 310 us at N = 4096, the ring idle
// during the math and the math idle during the MMAs):
//   roles K / Q   tensor pipe:  dP'(t) | S'(t+1) | acc(t) | dP'(t+1) ...      (ring blocks in exactly this order)
//                 math warps:   pass 1 (t): P' = exp2(S' ..) kept packed in registers, S' released   — runs under dP'(t)
//                               pass 2 (t): dS' = P' (scale dP' - delta) -> smem                      — runs under S'(t+1)
//   role V        S' is double-buffered in TMEM (columns [0,128) / [128,256), P' packed over its own S'):
//                 tensor pipe:  S'(t+1) | acc(t) ...      math: P'(t) while S'(t+1) streams
// tm_x: resident row-side operand, tm_y: streamed column-side counterpart, tm_u / tm_w: row- / column-side operands of dP'
template <int ROLE>
__device__ __forceinline__ void wide_bwd_body(unsigned char* smem, WideBwdShared& sh, const CUtensorMap* tm_x,
                                              const CUtensorMap* tm_y, const CUtensorMap* tm_u, const CUtensorMap* tm_w,
                                              __half* out, const WideBwdParams& p, int h, int slice) {
  using Cfg = WideBwdCfg;
  constexpr bool HAS_DP = ROLE != ROLE_V;      // dP' and dS' are needed
  constexpr bool COLSTAT = ROLE != ROLE_Q;     // lse / delta belong to the streamed tile's rows (columns of S')
  constexpr int RING = HAS_DP ? Cfg::RING_DS : Cfg::RING_V;
  constexpr int NACC = Cfg::NACC;
  unsigned char* sX = smem;
  unsigned char* sDS = sX + kWNB * kWBlk;
  unsigned char* sR = HAS_DP ? sDS + 2 * kWBlk : sDS;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128, b = blockIdx.z;
  const int n = p.n;
  const int T = (n + 127) / 128;
  const uint32_t tmem = sh.tmem_base;
  int* dead = &sh.dead;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    mbar_expect_tx_w(&sh.x_full, kWNB * kWBlk);
    for (int blk = 0; blk < kWNB; ++blk) tma_load_4d_w(sX + blk * kWBlk, tm_x, &sh.x_full, blk * 64, h, r0, b);
    int idx = 0;
    bool ok = true;
    auto push = [&](const CUtensorMap* tm, int col, int row) {
      const int slot = idx % RING;
      if (ok) ok = mbar_wait_warp(&sh.r_empty[slot], ((idx / RING) & 1) ^ 1, dead, p.err, 10);
      if (ok) {
        mbar_expect_tx_w(&sh.r_full[slot], kWBlk);
        tma_load_4d_w(sR + slot * kWBlk, tm, &sh.r_full[slot], col, h, row, b);
      }
      ++idx;
    };
    for (int blk = 0; blk < kWNB; ++blk) push(tm_y, blk * 64, 0);                                         // S'(0)
    for (int t = 0; t < T && ok; ++t) {
      if (HAS_DP)
        for (int blk = 0; blk < kWNB; ++blk) { push(tm_u, blk * 64, r0); push(tm_w, blk * 64, t * 128); }  // dP'(t)
      if (t + 1 < T)
        for (int blk = 0; blk < kWNB; ++blk) push(tm_y, blk * 64, (t + 1) * 128);                          // S'(t+1)
      for (int s = 0; s < Cfg::NCB; ++s)                                                                  // acc(t) += A B_c
        push(ROLE == ROLE_V ? tm_w : tm_y, slice * NACC + s * 64, t * 128);
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);        // K-major, 128B swizzle
    constexpr uint64_t mndesc_hi = umma_desc_hi_sw128(kWBlk, 1024);    // MN-major, 64-column groups one block apart
    constexpr uint32_t idesc_kk = umma_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t idesc_acc = umma_idesc_f16(128, 128, 0, 1);
    const uint32_t x_addr = smem_u32(sX), ds_addr = smem_u32(sDS), r_addr = smem_u32(sR);
    int cidx = 0;
    bool ok = mbar_wait_warp(&sh.x_full, 0, dead, p.err, 20);
    auto issue_s = [&](int buf) {  // S' = X Y^T into TMEM columns [128 buf, +128)
      for (int blk = 0; blk < kWNB && ok; ++blk) {
        const int slot = cidx % RING;
        ok = mbar_wait_warp(&sh.r_full[slot], (cidx / RING) & 1, dead, p.err, 21);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_w(tmem + buf * 128, umma_desc(kdesc_hi, x_addr + blk * kWBlk + k * 32),
                    umma_desc(kdesc_hi, r_addr + slot * kWBlk + k * 32), idesc_kk, blk > 0 || k > 0);
        umma_commit_w(&sh.r_empty[slot]);
        ++cidx;
      }
      if (ok) umma_commit_w(&sh.s_full[buf]);
    };
    auto issue_dp = [&]() {  // dP' = U W^T: both operands streamed, (U block, W block) in adjacent slots
      for (int blk = 0; blk < kWNB && ok; ++blk) {
        const int slot = cidx % RING;  // even
        ok = mbar_wait_warp(&sh.r_full[slot], (cidx / RING) & 1, dead, p.err, 22) &&
             mbar_wait_warp(&sh.r_full[slot + 1], (cidx / RING) & 1, dead, p.err, 23);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_w(tmem + Cfg::TMEM_DP, umma_desc(kdesc_hi, r_addr + slot * kWBlk + k * 32),
                    umma_desc(kdesc_hi, r_addr + (slot + 1) * kWBlk + k * 32), idesc_kk, blk > 0 || k > 0);
        umma_commit_w(&sh.r_empty[slot]);
        umma_commit_w(&sh.r_empty[slot + 1]);
        cidx += 2;
      }
      if (ok) umma_commit_w(&sh.dp_full);
    };
    // acc[:, 128 g .. +128) += A B_c, K = the 128 streamed rows; A = P' (packed fp16 in TMEM over S' buffer t & 1, role V) or
    // dS' (smem, K-major: 64 streamed rows per block); B_c = two adjacent ring blocks, MN-major
    auto issue_acc = [&](int t) {
      for (int g = 0; g < Cfg::NCB / 2 && ok; ++g) {
        const int slot = cidx % RING;  // even
        ok = mbar_wait_warp(&sh.r_full[slot], (cidx / RING) & 1, dead, p.err, 25) &&
             mbar_wait_warp(&sh.r_full[slot + 1], (cidx / RING) & 1, dead, p.err, 26);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t bdesc = umma_desc(mndesc_hi, r_addr + slot * kWBlk + k * 2048);
          if (ROLE == ROLE_V)
            umma_ts_w(tmem + Cfg::TMEM_ACC + g * 128, tmem + (t & 1) * 128 + k * 8, bdesc, idesc_acc, t > 0 || k > 0);
          else
            umma_ss_w(tmem + Cfg::TMEM_ACC + g * 128, umma_desc(kdesc_hi, ds_addr + (k / 4) * kWBlk + (k % 4) * 32), bdesc,
                      idesc_acc, t > 0 || k > 0);
        }
        umma_commit_w(&sh.r_empty[slot]);
        umma_commit_w(&sh.r_empty[slot + 1]);
        cidx += 2;
      }
      if (ok && HAS_DP) umma_commit_w(&sh.acc_done);  // the dS' tile may be rewritten
    };
    if (ok) issue_s(0);
    for (int t = 0; t < T && ok; ++t) {
      if (HAS_DP) {
        // dP'(t): its TMEM columns were released by pass 2 of tile t-1 (pds_ready(t-1), waited below)
        issue_dp();
        if (!ok) break;
        if (t + 1 < T) {  // S'(t+1) once pass 1 of tile t holds P' in registers
          ok = mbar_wait_warp(&sh.s_consumed, t & 1, dead, p.err, 27);
          if (!ok) break;
          tc_fence_after();
          issue_s(0);
          if (!ok) break;
        }
      } else if (t + 1 < T) {
        // S'(t+1) into the other buffer: its previous content P'(t-1) is read by acc(t-1), issued before (in-order pipe)
        issue_s((t + 1) & 1);
        if (!ok) break;
      }
      ok = mbar_wait_warp(&sh.pds_ready, t & 1, dead, p.err, 24);
      if (!ok) break;
      tc_fence_after();
      issue_acc(t);
    }
    if (ok) umma_commit_w(&sh.acc_full);
  } else {
    // ===================================== per-row math ======================================
    const int r = ((warp & 3) << 5) + lane;  // row of the resident tile = TMEM lane
    const int row = r0 + r;
    const bool row_ok = row < n;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const float* lse_bh = p.lse + ((long long)b * p.heads + h) * n;
    const float* delta_bh = p.delta + ((long long)b * p.heads + h) * n;
    // role Q: the statistics belong to this thread's row; roles K / V: to the streamed tile's rows (columns of S')
    const float my_lse2 = (!COLSTAT && row_ok) ? -lse_bh[row] * 1.4426950408889634f : -INFINITY;
    const float my_dl = (!COLSTAT && row_ok) ? -delta_bh[row] * p.scale : 0.f;
    auto load_lse2 = [&](int t) { const int qi = t * 128 + r; return (t < T && qi < n) ? -lse_bh[qi] * 1.4426950408889634f : -INFINITY; };
    auto load_dl = [&](int t) { const int qi = t * 128 + r; return (HAS_DP && t < T && qi < n) ? -delta_bh[qi] * p.scale : 0.f; };
    float pre_lse2 = COLSTAT ? load_lse2(0) : 0.f, pre_dl = COLSTAT ? load_dl(0) : 0.f;
    bool ok = true;
    for (int t = 0; t < T; ++t) {
      if (COLSTAT) {  // stage the streamed tile's statistics for broadcast reads (prefetched one tile ahead)
        if (t > 0) named_bar_sync(2, 128);  // every thread is done reading the previous tile's statistics
        sh.lse2[r] = pre_lse2;
        if (HAS_DP) sh.dl[r] = pre_dl;
        named_bar_sync(1, 128);
        pre_lse2 = load_lse2(t + 1);
        pre_dl = load_dl(t + 1);
      }
      const int buf = HAS_DP ? 0 : (t & 1);
      ok = mbar_wait_warp(&sh.s_full[buf], (HAS_DP ? t : (t >> 1)) & 1, dead, p.err, 30);
      if (!ok) break;
      tc_fence_after();
      const int valid = n - t * 128;  // streamed rows (columns of S') that exist
      // ---- pass 1: P' = exp2(S' scale log2e - lse), packed fp16 ----
      uint32_t pk[64];
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t s[32];
        tmem_ld32(lane_addr + buf * 128 + c0, s);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float nlv[4];
          if (COLSTAT) {
            const float4 nl = *reinterpret_cast<const float4*>(&sh.lse2[c0 + q]);
            nlv[0] = nl.x; nlv[1] = nl.y; nlv[2] = nl.z; nlv[3] = nl.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) nlv[e] = my_lse2;
          }
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pv[e] = fast_exp2(fmaf(__uint_as_float(s[q + e]), p.scale_log2, nlv[e]));
            if (!COLSTAT && c0 + q + e >= valid) pv[e] = 0.f;  // keys past the end of the sequence
            if (COLSTAT && !row_ok) pv[e] = 0.f;               // key rows past the end of the sequence
          }
          pk[(c0 + q) >> 1] = pack_half2(pv[0], pv[1]);
          pk[((c0 + q) >> 1) + 1] = pack_half2(pv[2], pv[3]);
        }
        // role V: P' over the S' columns this thread has already read: columns [c0/2, c0/2 + 16) of its buffer
        if (ROLE == ROLE_V) tmem_st16(lane_addr + buf * 128 + (c0 >> 1), pk + (c0 >> 1));
      }
      if (ROLE == ROLE_V) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.pds_ready);
        continue;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.s_consumed);  // S'(t) is in registers: S'(t+1) may be issued
      // ---- pass 2: dS' = scale P' (dP' - delta) -> smem ----
      ok = mbar_wait_warp(&sh.dp_full, t & 1, dead, p.err, 32);
      if (ok && t > 0) ok = mbar_wait_warp(&sh.acc_done, (t - 1) & 1, dead, p.err, 33);  // acc(t-1) has read the dS' tile
      if (!ok) break;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t dp[32];
        tmem_ld32(lane_addr + Cfg::TMEM_DP + c0, dp);
        tmem_ld_wait();
        uint32_t dk[16];
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float ndv[4];
          if (COLSTAT) {
            const float4 nd = *reinterpret_cast<const float4*>(&sh.dl[c0 + q]);
            ndv[0] = nd.x; ndv[1] = nd.y; ndv[2] = nd.z; ndv[3] = nd.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) ndv[e] = my_dl;
          }
          const float2 p01 = __half22float2(*reinterpret_cast<const __half2*>(&pk[(c0 + q) >> 1]));
          const float2 p23 = __half22float2(*reinterpret_cast<const __half2*>(&pk[((c0 + q) >> 1) + 1]));
          // scale * P' * (dP' - delta); masked rows / columns have P' = 0
          const float d0 = p01.x * fmaf(__uint_as_float(dp[q]), p.scale, ndv[0]);
          const float d1 = p01.y * fmaf(__uint_as_float(dp[q + 1]), p.scale, ndv[1]);
          const float d2 = p23.x * fmaf(__uint_as_float(dp[q + 2]), p.scale, ndv[2]);
          const float d3 = p23.y * fmaf(__uint_as_float(dp[q + 3]), p.scale, ndv[3]);
          dk[q >> 1] = pack_half2(d0, d1);
          dk[(q >> 1) + 1] = pack_half2(d2, d3);
        }
        // dS' row r, streamed columns [c0, c0 + 32): four 16-byte chunks of block c0 / 64
        unsigned char* ds_blk = sDS + (c0 >> 6) * kWBlk;
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          *reinterpret_cast<uint4*>(ds_blk + sw128_offset(r, chunk0 + cc)) =
              make_uint4(dk[4 * cc], dk[4 * cc + 1], dk[4 * cc + 2], dk[4 * cc + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.pds_ready);
    }
    // -------- epilogue: this row's gradient columns [slice * NACC, +NACC) --------
    ok = __all_sync(0xffffffffu, ok);
    if (ok) ok = mbar_wait_warp(&sh.acc_full, 0, dead, p.err, 31);
    if (ok) {
      tc_fence_after();
      __half* orow = out + ((long long)b * n + row) * p.d_tok + h * kWD + slice * NACC;
      for (int cc = 0; cc < NACC; cc += 8) {
        uint32_t a[8];
        tmem_ld8(lane_addr + Cfg::TMEM_ACC + cc, a);
        tmem_ld_wait();
        if (row_ok) {
          uint4 v;
          v.x = pack_half2(__uint_as_float(a[0]), __uint_as_float(a[1]));
          v.y = pack_half2(__uint_as_float(a[2]), __uint_as_float(a[3]));
          v.z = pack_half2(__uint_as_float(a[4]), __uint_as_float(a[5]));
          v.w = pack_half2(__uint_as_float(a[6]), __uint_as_float(a[7]));
          *reinterpret_cast<uint4*>(orow + cc) = v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(WideBwdCfg::THREADS, 1)
sattn_wide_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                      const WideBwdParams p) {
  using Cfg = WideBwdCfg;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(16) WideBwdShared sh;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  // the two heavy roles first: the light dV CTAs (192 instead of 448 KB streamed per tile pair) fill the tail of the grid
  const int per_role = p.heads * Cfg::NSL;
  const int role = blockIdx.y / per_role, h = (blockIdx.y % per_role) / Cfg::NSL, slice = blockIdx.y % Cfg::NSL;

  if (threadIdx.x == 0) {
    sh.dead = 0;
    mbar_init(&sh.x_full, 1);
    for (int i = 0; i < Cfg::RING_V; ++i) { mbar_init(&sh.r_full[i], 1); mbar_init(&sh.r_empty[i], 1); }
    mbar_init(&sh.s_full[0], 1);
    mbar_init(&sh.s_full[1], 1);
    mbar_init(&sh.s_consumed, 4);
    mbar_init(&sh.dp_full, 1);
    mbar_init(&sh.pds_ready, 4);
    mbar_init(&sh.acc_done, 1);
    mbar_init(&sh.acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh.tmem_base, 512);
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

  if (role == ROLE_K) wide_bwd_body<ROLE_K>(smem, sh, &tm_k, &tm_q, &tm_v, &tm_do, p.d_k, p, h, slice);
  else if (role == ROLE_Q) wide_bwd_body<ROLE_Q>(smem, sh, &tm_q, &tm_k, &tm_do, &tm_v, p.d_q, p, h, slice);
  else wide_bwd_body<ROLE_V>(smem, sh, &tm_k, &tm_q, nullptr, &tm_do, p.d_v, p, h, slice);

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(sh.tmem_base, 512);
}

// delta[b,h,i] = <dO[b,i,h,:], O[b,i,h,:]> over the 512 columns of a head: one warp per (b, i, h), two 16-byte vectors of O
// and dO per lane, all loads issued before the shuffle tree.
__global__ void __launch_bounds__(256) sattn_wide_delta_kernel(const __half* __restrict__ o, const __half* __restrict__ d_o,
                                                               float* __restrict__ delta, long long total, int n, int heads,
                                                               long long o_ts, long long o_bs, long long do_ts,
                                                               long long do_bs) {
  const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= total) return;  // whole warps leave together
  const int h = item % heads;
  const long long bi = item / heads;
  const int i = bi % n;
  const long long b = bi / n;
  const __half* orow = o + b * o_bs + i * o_ts + h * kWD;
  const __half* grow = d_o + b * do_bs + i * do_ts + h * kWD;
  uint4 a[2], g[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    a[t] = *reinterpret_cast<const uint4*>(orow + (t * 32 + lane) * 8);
    g[t] = *reinterpret_cast<const uint4*>(grow + (t * 32 + lane) * 8);
  }
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const __half2* ah = reinterpret_cast<const __half2*>(&a[t]);
    const __half2* gh = reinterpret_cast<const __half2*>(&g[t]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __half22float2(ah[e]), y = __half22float2(gh[e]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[((long long)b * heads + h) * n + i] = acc;
}

static int wide_tmap(CUtensorMap* m, const void* ptr, int heads, int n, int batch, long long ts, long long bs) {
  const uint64_t dims[4] = {(uint64_t)kWD, (uint64_t)heads, (uint64_t)n, (uint64_t)batch};
  const uint64_t st[4] = {2, (uint64_t)kWD * 2, (uint64_t)ts * 2, (uint64_t)bs * 2};
  const uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_f16(m, ptr, 4, dims, st, box);
}

template <int DV>
static int launch_wide_fwd_dv(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v, const WideFwdParams& p,
                              const sta_sattn_fwd_args* a, cudaStream_t stream) {
  using Cfg = WideFwdCfg<DV>;
  static PerDeviceOnce smem_attr;
  int rc;
  if ((rc = smem_attr.run([] {
        return cudaFuncSetAttribute(sattn_wide_fwd_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
      })))
    return rc;
  dim3 grid((a->n + 127) / 128, a->heads * Cfg::NSL, a->batch);
  sattn_wide_fwd_kernel<DV><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(tm_q, tm_k, tm_v, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

int launch_sattn_wide_fwd(const sta_sattn_fwd_args* a, cudaStream_t stream) {
  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  if ((rc = wide_tmap(&tm_q, a->q, a->heads, a->n, a->batch, a->q_token_stride, a->q_batch_stride))) return rc;
  if ((rc = wide_tmap(&tm_k, a->k, a->heads, a->n, a->batch, a->k_token_stride, a->k_batch_stride))) return rc;
  if ((rc = wide_tmap(&tm_v, a->v, a->heads, a->n, a->batch, a->v_token_stride, a->v_batch_stride))) return rc;
  WideFwdParams p;
  p.out = reinterpret_cast<__half*>(a->out);
  p.lse = a->lse;
  p.n = a->n;
  p.heads = a->heads;
  p.o_token_stride = a->o_token_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static std::atomic<int> sm_count{0};  // every GPU of a box has the same SM count: one query per process
  int num_sms = sm_count.load(std::memory_order_relaxed);
  if (num_sms == 0) {
    int dev = 0;
    num_sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    sm_count.store(num_sms, std::memory_order_relaxed);
  }
  const long long ctas256 = (long long)((a->n + 127) / 128) * a->heads * 2 * a->batch;
  if (ctas256 < num_sms) return launch_wide_fwd_dv<128>(tm_q, tm_k, tm_v, p, a, stream);
  return launch_wide_fwd_dv<256>(tm_q, tm_k, tm_v, p, a, stream);
}

int launch_sattn_wide_bwd(const sta_sattn_bwd_args* a, cudaStream_t stream) {
  using Cfg = WideBwdCfg;
  CUtensorMap tm_q, tm_k, tm_v, tm_do;
  int rc;
  if ((rc = wide_tmap(&tm_q, a->q, a->heads, a->n, a->batch, a->q_token_stride, a->q_batch_stride))) return rc;
  if ((rc = wide_tmap(&tm_k, a->k, a->heads, a->n, a->batch, a->k_token_stride, a->k_batch_stride))) return rc;
  if ((rc = wide_tmap(&tm_v, a->v, a->heads, a->n, a->batch, a->v_token_stride, a->v_batch_stride))) return rc;
  if ((rc = wide_tmap(&tm_do, a->d_out, a->heads, a->n, a->batch, a->do_token_stride, a->do_batch_stride))) return rc;

  const long long total = (long long)a->batch * a->n * a->heads;
  sattn_wide_delta_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(a->out), reinterpret_cast<const __half*>(a->d_out), a->delta, total, a->n, a->heads,
      a->o_token_stride, a->o_batch_stride, a->do_token_stride, a->do_batch_stride);
  STA_CUDA_CHECK(cudaGetLastError());

  WideBwdParams p;
  p.lse = a->lse;
  p.delta = a->delta;
  p.d_q = reinterpret_cast<__half*>(a->d_q);
  p.d_k = reinterpret_cast<__half*>(a->d_k);
  p.d_v = reinterpret_cast<__half*>(a->d_v);
  p.d_tok = a->dqkv_token_stride > 0 ? a->dqkv_token_stride : (long long)a->heads * kWD;
  p.n = a->n;
  p.heads = a->heads;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.err = device_error_word();
  static PerDeviceOnce smem_attr;
  if ((rc = smem_attr.run([] {
        return cudaFuncSetAttribute(sattn_wide_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
      })))
    return rc;
  dim3 grid((a->n + 127) / 128, 3 * a->heads * Cfg::NSL, a->batch);  // roles K, Q, V
  sattn_wide_bwd_kernel<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(tm_q, tm_k, tm_v, tm_do, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta
