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
// CTA = one 128-pixel tile of one (prompt, head); 10 warps:
//   warp 0      TMA producer: Q_u and Q_c tiles, then the K/V tiles (80 rows, rows 77..79 zero-filled by TMA)
//               of every ACTIVE context (an object whose mask is empty inside this pixel tile is skipped)
//   warp 1      MMA issuer: S = Q K^T (SS), O += P V (TS, P packed fp16 in TMEM, V MN-major)
//   warps 2..5  softmax + epilogue of the unconditional row (one context)
//   warps 6..9  softmax + epilogue of the conditional row (1 + active objects contexts)
// TMEM columns: S_u [0,96) S_c [96,192) (P aliases S), O_u at 192, O_c at 192 + DMMA.
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
  static constexpr int ST = (NBLK == 3) ? 2 : 3;  // K and V ring depth
  static constexpr int QTILE = NBLK * kXQBlockBytes;
  static constexpr int CTILE = NBLK * kXCBlockBytes;
  static constexpr int SMEM_BYTES = 2 * QTILE + 2 * ST * CTILE + 1024;
  static constexpr int THREADS = 320;
  static constexpr int TMEM_S = 96;
  static constexpr int TMEM_O = 192;
};

struct XattnFwdParams {
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
__global__ void __launch_bounds__(XattnCfg<D>::THREADS, 1)
xattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const XattnFwdParams p) {
  using Cfg = XattnCfg<D>;
  constexpr int ST = Cfg::ST, NBLK = Cfg::NBLK, DMMA = Cfg::DMMA;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + 2 * Cfg::QTILE;
  unsigned char* sV = sK + ST * Cfg::CTILE;

  __shared__ uint64_t q_full, k_full[ST], k_empty[ST], v_full[ST], v_empty[ST];
  __shared__ uint64_t s_full[2], p_ready[2], o_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;
  __shared__ int tile_slot[2 + kXMaxObj];  // context slot of the t-th tile this CTA processes
  __shared__ int n_tiles_s;

  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, pr = blockIdx.z;
  const int n = p.n, B = p.prompts, n_obj = p.n_obj, n_slots = 2 + p.n_obj;

  if (tid == 0) {
    dead = 0;
    mbar_init(&q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (warp == 0) {
    // which objects touch this pixel tile?  (128 mask bytes per object: one 4-byte word per lane)
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
    }
    int cnt = 2;
    if (lane == 0) { tile_slot[0] = 0; tile_slot[1] = 1; }
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
  const int T = n_tiles_s;  // tile 0 -> unconditional row, tiles 1..T-1 -> conditional row

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      mbar_expect_tx_w(&q_full, 2 * Cfg::QTILE);
      for (int r = 0; r < 2; ++r)
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d_w(sQ + (r * NBLK + blk) * kXQBlockBytes, &tm_q, &q_full, blk * 64, h, q0, pr + r * B);
      for (int t = 0; t < T; ++t) {
        const int st = t % ST, slot = pr * n_slots + tile_slot[t];
        if (!mbar_wait_warp(&k_empty[st], ((t / ST) & 1) ^ 1, &dead, p.err, 10)) break;
        mbar_expect_tx_w(&k_full[st], Cfg::CTILE);
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d_w(sK + (st * NBLK + blk) * kXCBlockBytes, &tm_k, &k_full[st], blk * 64, h, 0, slot);
        if (!mbar_wait_warp(&v_empty[st], ((t / ST) & 1) ^ 1, &dead, p.err, 11)) break;
        mbar_expect_tx_w(&v_full[st], Cfg::CTILE);
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d_w(sV + (st * NBLK + blk) * kXCBlockBytes, &tm_v, &v_full[st], blk * 64, h, 0, slot);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    {  // whole warp, warp-uniform control flow; single-lane instructions are elected inside the *_w helpers
      constexpr uint64_t kdesc_hi = umma_desc_hi_sw128(16, 1024);
      constexpr uint64_t vdesc_hi = umma_desc_hi_sw128(kXCBlockBytes, 1024);
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 80, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, DMMA, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);

      auto issue_qk = [&](int r, int st) {
#pragma unroll
        for (int k = 0; k < DMMA / 16; ++k) {
          const uint32_t qoff = (k / 4) * kXQBlockBytes + (k % 4) * 32;
          const uint32_t koff = (k / 4) * kXCBlockBytes + (k % 4) * 32;
          umma_ss_w(tmem + r * Cfg::TMEM_S, umma_desc(kdesc_hi, q_addr + r * Cfg::QTILE + qoff),
                  umma_desc(kdesc_hi, k_addr + st * Cfg::CTILE + koff), idesc_qk, k > 0);
        }
        umma_commit_w(&s_full[r]);
        umma_commit_w(&k_empty[st]);
      };
      auto issue_pv = [&](int r, int st, bool acc) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
          umma_ts_w(tmem + Cfg::TMEM_O + r * DMMA, tmem + r * Cfg::TMEM_S + k * 8,
                  umma_desc(vdesc_hi, v_addr + st * Cfg::CTILE + k * 2048), idesc_pv, acc || k > 0);
        umma_commit_w(&v_empty[st]);
      };

      bool ok = mbar_wait_warp(&q_full, 0, &dead, p.err, 20) && mbar_wait_warp(&k_full[0], 0, &dead, p.err, 21);
      if (ok) {
        tc_fence_after();
        issue_qk(0, 0);
        ok = mbar_wait_warp(&k_full[1 % ST], (1 / ST) & 1, &dead, p.err, 22);
      }
      if (ok) {
        tc_fence_after();
        issue_qk(1, 1 % ST);
        ok = mbar_wait_warp(&p_ready[0], 0, &dead, p.err, 23) && mbar_wait_warp(&v_full[0], 0, &dead, p.err, 24);
      }
      if (ok) {
        tc_fence_after();
        issue_pv(0, 0, false);
        umma_commit_w(&o_full[0]);
      }
      for (int t = 1; t < T && ok; ++t) {
        const int st = t % ST;
        ok = mbar_wait_warp(&p_ready[1], (t - 1) & 1, &dead, p.err, 25) &&
             mbar_wait_warp(&v_full[st], (t / ST) & 1, &dead, p.err, 26);
        if (!ok) break;
        tc_fence_after();
        issue_pv(1, st, t > 1);
        if (t + 1 < T) {
          const int st1 = (t + 1) % ST;
          ok = mbar_wait_warp(&k_full[st1], ((t + 1) / ST) & 1, &dead, p.err, 27);
          if (!ok) break;
          tc_fence_after();
          issue_qk(1, st1);
        } else {
          umma_commit_w(&o_full[1]);
        }
      }
    }
  } else {
    // ===================================== softmax / epilogue ================================
    const int r = (warp - 2) >> 2;  // 0 = unconditional row, 1 = conditional row
    const int row = q0 + ((warp & 3) << 5) + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) << 5) << 16);
    const uint32_t s_addr = lane_addr + r * Cfg::TMEM_S;
    const int t_begin = r == 0 ? 0 : 1, t_end = r == 0 ? 1 : T;
    float sigma = 0.f;  // sum_i m_i c_i of this pixel
    bool ok = true;
    for (int t = t_begin; t < t_end; ++t) {
      const int slot = tile_slot[t];
      float w = 1.f;
      if (slot >= 2) {
        const int i = slot - 2;
        const float mk = row < n ? (float)p.mask[((long long)pr * n_obj + i) * n + row] : 0.f;
        w = mk * p.coef[pr * n_obj + i];
        sigma += w;
      }
      ok = mbar_wait_warp(&s_full[r], (t - t_begin) & 1, &dead, p.err, 30);
      if (!ok) break;
      tc_fence_after();
      uint32_t s[80];
      tmem_ld32(s_addr, s);
      tmem_ld32(s_addr + 32, s + 32);
      tmem_ld16(s_addr + 64, s + 64);
      tmem_ld_wait();
      const int valid = p.ctx_len;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 80; c += 2) {
        if (c >= valid) s[c] = 0xff800000u;
        if (c + 1 >= valid) s[c + 1] = 0xff800000u;
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
      const float f = w / l;
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
      if (lane == 0) mbar_arrive(&p_ready[r]);
      if (p.lse && row < n)
        p.lse[(((long long)pr * p.heads + h) * n_slots + slot) * n + row] = (m + log2f(l)) * 0.6931471805599453f;
    }
    // -------- epilogue --------
    ok = __all_sync(0xffffffffu, ok);
    if (ok) ok = mbar_wait_warp(&o_full[0], 0, &dead, p.err, 31);
    if (ok && r == 1) ok = mbar_wait_warp(&o_full[1], 0, &dead, p.err, 32);
    if (ok) {
      tc_fence_after();
      const uint32_t ou_addr = lane_addr + Cfg::TMEM_O;
      const uint32_t oc_addr = ou_addr + DMMA;
      __half* orow = p.out + (long long)(pr + r * B) * p.o_batch_stride + (long long)row * p.o_token_stride + h * D;
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 8) {
        uint32_t u[8], c[8];
        tmem_ld8(ou_addr + c0, u);
        if (r == 1) tmem_ld8(oc_addr + c0, c);
        tmem_ld_wait();
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = r == 0 ? __uint_as_float(u[i]) : fmaf(-sigma, __uint_as_float(u[i]), __uint_as_float(c[i]));
        if (row < n) {
          uint4 v;
          v.x = pack_half2(o[0], o[1]);
          v.y = pack_half2(o[2], o[3]);
          v.z = pack_half2(o[4], o[5]);
          v.w = pack_half2(o[6], o[7]);
          *reinterpret_cast<uint4*>(orow + c0) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int D>
static int launch_xattn_fwd(const sta_xattn_fwd_args* a, cudaStream_t stream) {
  using Cfg = XattnCfg<D>;
  CUtensorMap tm_q, tm_k, tm_v;
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->heads, (uint64_t)a->n, (uint64_t)a->prompts * 2};
    const uint64_t st[4] = {2, (uint64_t)D * 2, (uint64_t)a->q_token_stride * 2, (uint64_t)a->q_batch_stride * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_f16(&tm_q, a->q, 4, dims, st, box);
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
  static bool attr_set = false;
  if (!attr_set) {
    STA_CUDA_CHECK(cudaFuncSetAttribute(xattn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a->n + 127) / 128, a->heads, a->prompts);
  xattn_fwd_kernel<D><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tm_q, tm_k, tm_v, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

}  // namespace sta

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
