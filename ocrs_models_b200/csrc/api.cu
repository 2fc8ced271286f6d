// Library-wide C-ABI plumbing: error string, version, device query.
#include "common.cuh"
#include "tma_util.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <unordered_map>

static thread_local char g_err[512] = "";

void ocrs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void ocrs_count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// ---- CUtensorMap cache (process lifetime, guarded by a mutex; the library allocates nothing else) ----
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
struct MapKey {
  unsigned long long v[14];
};
std::mutex g_map_mutex;
std::unordered_map<std::string, CUtensorMap> g_maps;
}  // namespace

int ocrs_get_tensor_map(CUtensorMap* out, const void* ptr, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box, int swizzle128) {
  OCRS_CHECK_ARG(rank >= 2 && rank <= 4, "tensor map: rank %d unsupported", rank);
  MapKey k;
  memset(&k, 0, sizeof(k));
  int dev = 0;
  cudaGetDevice(&dev);
  k.v[0] = (unsigned long long)(uintptr_t)ptr;
  k.v[1] = ((unsigned long long)rank << 32) | ((unsigned long long)dev << 8) | (unsigned long long)swizzle128;
  for (int i = 0; i < rank; ++i) {
    k.v[2 + i] = dims[i];
    k.v[6 + i] = i + 1 < rank ? strides_bytes[i] : 0;
    k.v[10 + i] = box[i];
  }
  const std::string key((const char*)&k, sizeof(k));
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn enc = get_encode();
  OCRS_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled is unavailable in this driver");
  cuuint64_t d[4], st[3];
  cuuint32_t b[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)ptr, d, st, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   // haloed boxes start 16 bytes before a 128-byte line and are 160 bytes wide: promoting every row to whole
                   // 128-byte lines would fetch 256-384 bytes for 160 (measured: DRAM reads 2x the algorithmic bytes)
                   // (no measurable effect on the kernels' time either way: they are bound by instruction issue)
                   CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OCRS_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u", (int)r, rank,
                 dims[0], dims[1], box[0], box[1]);
  if (g_maps.size() > 4096) g_maps.clear();  // bound the cache if a caller never reuses addresses
  g_maps.emplace(key, m);
  *out = m;
  return 0;
}

int ocrs_plane_map(CUtensorMap* out, const float* base, long long ss, int N, int C, int H, int W, int bw, int bh, int bc) {
  const unsigned long long dims[4] = {(unsigned long long)W, (unsigned long long)H, (unsigned long long)C, (unsigned long long)N};
  const unsigned long long strides[3] = {(unsigned long long)W * 4, (unsigned long long)H * W * 4, (unsigned long long)ss * 4};
  const unsigned box[4] = {(unsigned)bw, (unsigned)bh, (unsigned)bc, 1u};
  return ocrs_get_tensor_map(out, base, 4, dims, strides, box, 0);
}

extern "C" {
// Kernels launched by this library since load (every entry point counts its own launches).
long long ocrs_launch_count(void) { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
const char* ocrs_last_error(void) { return g_err; }
int ocrs_version(void) { return 100; }
// Compute capability major*10+minor of the current device (100 on B200), or <0 on error.
int ocrs_device_arch(void) {
  int dev = 0, maj = 0, min = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return maj * 10 + min;
}
}
