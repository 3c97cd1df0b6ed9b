// sta_probe.cu — single-CTA probe that runs ONE tcgen05 GEMM tile with caller-supplied descriptors.
//
// Test infrastructure for the primitives in sta_common.cuh (exported as sta_probe_gemm): it lets the GPU test
// suite check, against a torch matmul, every operand flavour the attention kernels rely on — K-major and
// MN-major 128B-swizzled shared-memory operands as TMA writes them (including zero-filled out-of-bounds
// columns/rows), A taken from TMEM as packed fp16, multi-block K, N that is not a multiple of 64 — and dump the
// raw shared-memory image so the swizzle formula (sw128_offset) is pinned too.
#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

struct ProbeDev {
  const __half* a;
  int a_rows, a_cols, a_in_tmem;
  int b_rows, b_cols;
  unsigned long long a_desc_hi, b_desc_hi;
  int nk;
  unsigned int a_off[16], b_off[16];
  unsigned int idesc;
  int n;
  float* out;
  unsigned char* smem_dump;
  int dump_bytes;
  unsigned int* err;
  int reps;
  long long* cycles;
};

constexpr int kProbeOperandBytes = 64 * 1024;

__global__ void __launch_bounds__(128, 1)
probe_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, ProbeDev p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + kProbeOperandBytes;
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int dead;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    dead = 0;
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const int a_boxes = p.a_in_tmem ? 0 : (p.a_cols + 63) / 64;
  const int b_boxes = (p.b_cols + 63) / 64;
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, (uint32_t)(a_boxes * p.a_rows * 128 + b_boxes * p.b_rows * 128));
    for (int j = 0; j < a_boxes; ++j) tma_load_2d(sA + j * p.a_rows * 128, &tm_a, &bar_tma, j * 64, 0);
    for (int j = 0; j < b_boxes; ++j) tma_load_2d(sB + j * p.b_rows * 128, &tm_b, &bar_tma, j * 64, 0);
  }
  if (p.a_in_tmem) {
    // thread = row; pack consecutive K pairs into one 32-bit TMEM column, A region starts at column 256
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + 256;
    const __half* row = p.a + (size_t)tid * p.a_cols;
    for (int c0 = 0; c0 < p.a_cols / 2; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int k = 2 * (c0 + j);
        float lo = k < p.a_cols ? __half2float(row[k]) : 0.f;
        float hi = k + 1 < p.a_cols ? __half2float(row[k + 1]) : 0.f;
        r[j] = pack_half2(lo, hi);
      }
      tmem_st8(lane_base + c0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    if (mbar_wait(&bar_tma, 0, &dead, p.err, 1)) {
      tc_fence_after();
      // descriptors are built before the timed region so that the loop below is (predicated) MMA issue only
      uint64_t adesc[16], bdesc[16];
      uint32_t atm[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        bdesc[k] = umma_desc(p.b_desc_hi, smem_u32(sB) + p.b_off[k]);
        adesc[k] = umma_desc(p.a_desc_hi, smem_u32(sA) + p.a_off[k]);
        atm[k] = tmem + 256 + p.a_off[k];
      }
      const bool alt = p.dump_bytes == -1;
      const int reps = p.reps > 1 ? p.reps : 1;
      const int nk = p.nk;
      const long long t0 = clock64();
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k < nk) {
            const uint32_t d = tmem + (alt ? (k & 1) * 128 : 0);
            const uint32_t acc = (k > (alt ? 1 : 0)) || rep > 0;
            if (p.a_in_tmem) umma_ts(d, atm[k], bdesc[k], p.idesc, acc);
            else umma_ss(d, adesc[k], bdesc[k], p.idesc, acc);
          }
        }
      }
      umma_commit(&bar_mma);
      if (p.cycles) {
        mbar_wait(&bar_mma, 0, &dead, p.err, 3);
        p.cycles[0] = clock64() - t0;
      }
    }
  }
  __syncwarp();
  if (mbar_wait(&bar_mma, 0, &dead, p.err, 2)) {
    tc_fence_after();
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < p.n; c += 8) {
      uint32_t r[8];
      tmem_ld8(lane_base + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < p.n) p.out[(size_t)tid * p.n + c + j] = __uint_as_float(r[j]);
    }
  }
  if (p.smem_dump)
    for (int i = tid; i < p.dump_bytes; i += 128) p.smem_dump[i] = smem[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace sta

extern "C" int sta_probe_gemm(const sta_probe_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->b || !a->out) return fail(STA_ERR_BAD_ARG, "sta_probe_gemm: null argument");
  if (a->nk < 1 || a->nk > 16) return fail(STA_ERR_BAD_ARG, "sta_probe_gemm: nk must be in [1,16]");
  if (a->a_rows > 256 || a->b_rows > 256 || ((a->a_cols + 63) / 64) * a->a_rows * 128 > kProbeOperandBytes ||
      ((a->b_cols + 63) / 64) * a->b_rows * 128 > kProbeOperandBytes)
    return fail(STA_ERR_UNSUPPORTED, "sta_probe_gemm: operand does not fit the 64 KiB staging area");
  CUtensorMap tm_a, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  {
    uint64_t dims[2] = {(uint64_t)a->b_cols, (uint64_t)a->b_tensor_rows};
    uint64_t str[2] = {2, (uint64_t)a->b_cols * 2};
    uint32_t box[2] = {64, (uint32_t)a->b_rows};
    int rc = make_tmap_f16(&tm_b, a->b, 2, dims, str, box);
    if (rc) return rc;
  }
  if (!a->a_in_tmem) {
    uint64_t dims[2] = {(uint64_t)a->a_cols, (uint64_t)a->a_tensor_rows};
    uint64_t str[2] = {2, (uint64_t)a->a_cols * 2};
    uint32_t box[2] = {64, (uint32_t)a->a_rows};
    int rc = make_tmap_f16(&tm_a, a->a, 2, dims, str, box);
    if (rc) return rc;
  } else if (a->a_rows != 128) {
    return fail(STA_ERR_BAD_ARG, "sta_probe_gemm: A in TMEM needs 128 rows");
  }
  ProbeDev p;
  p.a = reinterpret_cast<const __half*>(a->a);
  p.a_rows = a->a_rows; p.a_cols = a->a_cols; p.a_in_tmem = a->a_in_tmem;
  p.b_rows = a->b_rows; p.b_cols = a->b_cols;
  p.a_desc_hi = a->a_desc_hi; p.b_desc_hi = a->b_desc_hi;
  p.nk = a->nk;
  for (int i = 0; i < 16; ++i) { p.a_off[i] = a->a_off[i]; p.b_off[i] = a->b_off[i]; }
  p.idesc = a->idesc;
  p.n = a->n;
  p.out = a->out;
  p.smem_dump = reinterpret_cast<unsigned char*>(a->smem_dump);
  p.dump_bytes = a->dump_bytes;
  p.err = device_error_word();
  p.reps = a->reps;
  p.cycles = reinterpret_cast<long long*>(a->cycles);
  const int smem_bytes = 2 * kProbeOperandBytes + 1024;
  STA_CUDA_CHECK(cudaFuncSetAttribute(probe_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_gemm_kernel<<<1, 128, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tm_a, tm_b, p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

// ---------------------------------------------------------------------------------------------------------
// TMEM load/store throughput probe: `warps` warps (multiple of 4) each read (or write) `iters` x 32 columns of
// their 32 lanes; reports SM cycles for the whole CTA.  bytes = warps * 32 lanes * 32 cols * 4 B * iters.
// ---------------------------------------------------------------------------------------------------------
namespace sta {
__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(int iters, int mode, long long* cycles, unsigned int* sink) {
  __shared__ uint32_t tmem_base_s;
  __shared__ long long t_begin;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) << 5) << 16) + (warp >> 2) * 64;
  uint32_t acc = 0;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      tmem_ld32(base + (it & 1) * 32, r);
      tmem_ld_wait();
      acc += r[0] ^ r[13] ^ r[31];
    } else if (mode == 1) {  // two loads in flight
      uint32_t r2[32];
      tmem_ld32(base, r);
      tmem_ld32(base + 32, r2);
      tmem_ld_wait();
      acc += r[0] ^ r[31] ^ r2[7] ^ r2[31];
    } else {
      tmem_st16(base + (it & 1) * 16, r);
      tmem_st_wait();
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0xdeadbeef) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}
}  // namespace sta

extern "C" int sta_probe_tmem_bw(int warps, int iters, int mode, long long* cycles_dev, void* stream) {
  using namespace sta;
  if (warps < 4 || warps > 16 || (warps % 4)) return fail(STA_ERR_BAD_ARG, "warps must be 4, 8, 12 or 16");
  tmem_bw_kernel<<<1, warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(iters, mode, cycles_dev, device_error_word());
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}
