// sta_host.cu — error plumbing, tensor-map construction, library-level C-ABI entry points.
#include <stdarg.h>

#include <mutex>

#include "../../include/sta_b200.h"
#include "sta_common.cuh"
#include "sta_host.h"

namespace sta {

unsigned int* device_error_word() {
  static unsigned int* word = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (cudaMalloc(&word, sizeof(unsigned int)) != cudaSuccess) { word = nullptr; return; }
    cudaMemset(word, 0, sizeof(unsigned int));
  });
  return word;
}

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-map cache ("the library allocates nothing persistent except cached CUtensorMaps keyed by
// (ptr, shape)").  cuTensorMapEncodeTiled costs ~1-2 us of host time per map and the attention launchers build 4-6 maps per
// call; eager callers hit the same (pointer, geometry) over and over (q/k/v views of the cached-allocator's blocks), CUDA
// graphs do not care.  A tensor map is a pure function of its encode arguments, so a stale entry cannot exist: the key IS
// the full argument list.  Small direct-mapped table, one mutex.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct TmapKey {
  const void* base;
  int rank, kind;
  uint64_t dims[5], strides[5];
  uint32_t box[5];
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot {
  bool used = false;
  TmapKey key;
  CUtensorMap map;
};
constexpr int kTmapSlots = 1024;
TmapSlot g_tmap_cache[kTmapSlots];
std::mutex g_tmap_mu;

TmapKey make_key(const void* base, int rank, int kind, const uint64_t* dims, const uint64_t* strides, const uint32_t* box) {
  TmapKey k;
  memset(&k, 0, sizeof(k));  // padding bytes take part in the comparison
  k.base = base;
  k.rank = rank;
  k.kind = kind;
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.strides[i] = strides[i]; k.box[i] = box[i]; }
  return k;
}
size_t hash_key(const TmapKey& k) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
  uint64_t h = 1469598103934665603ull;  // FNV-1a
  for (size_t i = 0; i < sizeof(TmapKey); ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return static_cast<size_t>(h % kTmapSlots);
}
bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  const TmapSlot& s = g_tmap_cache[hash_key(k)];
  if (s.used && s.key == k) { *out = s.map; return true; }
  return false;
}
void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  TmapSlot& s = g_tmap_cache[hash_key(k)];
  s.used = true;
  s.key = k;
  s.map = m;
}
}  // namespace

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(STA_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return fail(STA_ERR_UNSUPPORTED, "tensor base %p is not 16-byte aligned", base);
  const TmapKey key = make_key(base, rank, /*kind=*/0, dims, strides_bytes, box);
  if (tmap_lookup(key, out)) return STA_OK;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i];
      if (strides_bytes[i] % 16 != 0)
        return fail(STA_ERR_UNSUPPORTED, "stride[%d]=%llu bytes is not a multiple of 16", i,
                    (unsigned long long)strides_bytes[i]);
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(STA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  tmap_store(key, *out);
  return STA_OK;
}

int make_tmap_f32_dense(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(STA_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(STA_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed with CUresult %d", (int)r);
  return STA_OK;
}

}  // namespace sta

extern "C" {

int sta_version(void) { return STA_B200_VERSION; }

const char* sta_last_error(void) { return sta::last_error_buf(); }

int sta_device_error(unsigned int* code_out, int clear) {
  unsigned int v = 0;
  unsigned int* w = sta::device_error_word();
  if (!w) return sta::fail(STA_ERR_CUDA, "could not allocate the device error word");
  STA_CUDA_CHECK(cudaMemcpy(&v, w, sizeof(v), cudaMemcpyDeviceToHost));
  if (code_out) *code_out = v;
  if (clear && v != 0) {
    STA_CUDA_CHECK(cudaMemset(w, 0, sizeof(v)));
  }
  if (v != 0) return sta::fail(STA_ERR_DEVICE, "device error word 0x%x (mbarrier wait timed out, id %u)", v, v & 0xff);
  return STA_OK;
}

}  // extern "C"
