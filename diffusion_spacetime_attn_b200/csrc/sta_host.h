// sta_host.h — host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map
// construction.  No libcuda link dependency: cuTensorMapEncodeTiled is resolved through the runtime.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/sta_b200.h"  // return codes STA_OK / STA_ERR_*

namespace sta {

// thread-local last-error text ("message via sta_last_error() (thread-local)")
char* last_error_buf();
int fail(int code, const char* fmt, ...);


#define STA_CUDA_CHECK(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return ::sta::fail(STA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// Runs `f` (a cudaError_t-returning callable, e.g. cudaFuncSetAttribute for a kernel's dynamic shared memory) once per
// DEVICE, thread-safely: function attributes are per device, and the launchers may be called from several host threads.
struct PerDeviceOnce {
  std::atomic<uint64_t> done[4];  // one bit per device ordinal (256 devices)
  std::mutex mu;
  PerDeviceOnce() { for (auto& d : done) d.store(0); }
  template <typename F>
  int run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(STA_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    const int w = (dev >> 6) & 3;
    const uint64_t bit = 1ull << (dev & 63);
    if (done[w].load(std::memory_order_acquire) & bit) return STA_OK;
    std::lock_guard<std::mutex> lock(mu);
    if (done[w].load(std::memory_order_acquire) & bit) return STA_OK;
    e = f();
    if (e != cudaSuccess) return fail(STA_ERR_CUDA, "per-device kernel attribute: %s", cudaGetErrorString(e));
    done[w].fetch_or(bit, std::memory_order_release);
    return STA_OK;
  }
};

// one 4-byte device word per process that kernels bump on an mbarrier timeout (never freed)
unsigned int* device_error_word();

// Tiled fp16 tensor map with 128B swizzle.  dims/strides innermost first; strides in BYTES for dims 1..rank-1.
// box[0] must be <= 64 elements (128 bytes).  OOB elements are filled with zeros.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

// Tiled fp32 tensor map WITHOUT swizzle (dense box rows), used as the destination of TMA reduce-add.
int make_tmap_f32_dense(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes = 0);  // 0 (dense rows), 32 or 64: smem-side XOR swizzle

// magic number of sta_common.cuh's fast_div for the divisor d >= 1 (valid for dividends < 2^31)
struct FastDiv;
inline void make_fast_div_raw(unsigned int d, unsigned int* magic, unsigned int* shift) {
  unsigned int sh = 0;
  while ((1ull << sh) < d) ++sh;
  *shift = sh;
  *magic = (unsigned int)((((1ull << 32) * ((1ull << sh) - d)) / d + 1) & 0xffffffffull);
}

// head dim 512 (KL-VAE mid-block AttnBlock): sta_sattn_wide.cu, reached through sta_sattn_fwd / sta_sattn_bwd
int launch_sattn_wide_fwd(const sta_sattn_fwd_args* a, cudaStream_t stream);
int launch_sattn_wide_bwd(const sta_sattn_bwd_args* a, cudaStream_t stream);

}  // namespace sta
