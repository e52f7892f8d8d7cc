// Library runtime: error string, launch counter, device properties.  No global mutable state beyond these.
#include "common.cuh"
#include "../../include/ga_sm100.h"
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string.h>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void ga_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void ga_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int ga_check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    ga_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return GA_ERR_LAUNCH;
  }
  return GA_OK;
}

int ga_num_sms() {
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = sms[dev & 63];
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}
bool ga_first_on_device(GaPerDevice& s) {
  int dev = 0;
  cudaGetDevice(&dev);
  bool& d = s.done[dev & 63];
  if (d) return false;
  d = true;
  return true;
}

extern "C" int ga_version(void) { return 100; }
// message of the calling thread's last failed call, copied into the caller's buffer (always NUL-terminated); returns its length
extern "C" int ga_last_error(char* buf, size_t n) {
  const size_t len = strlen(g_err);
  if (buf && n) {
    const size_t c = len < n - 1 ? len : n - 1;
    memcpy(buf, g_err, c);
    buf[c] = 0;
  }
  return (int)len;
}

// bytes of caller-owned scratch an entry point needs for the given shape (kernels never allocate)
extern "C" long long ga_workspace_bytes(int op, long long a, long long b, long long c, long long d) {
  switch (op) {
    case GA_WS_DWCONV7_BWD: return (long long)ga_dwconv7_bwd_parts((int)a, (int)b, (int)c, (int)d) * 50 * d * (long long)sizeof(float);
    case GA_WS_COLSTATS: return (long long)ga_colstats_parts(a, (int)b) * 2 * b * (long long)sizeof(float);
    case GA_WS_LAYERNORM_BWD: return (long long)ga_layernorm_bwd_parts(a, (int)b) * 2 * b * (long long)sizeof(float);
    default: ga_set_error("ga_workspace_bytes: unknown op %d", op); return -1;
  }
}
extern "C" long long ga_launch_count(void) { return g_launches.load(); }

// ---- tensor-map cache (immutable entries, mutex guarded; SURVEY.md 8b threading contract) -------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_mu;
struct MapKey {
  const void* ptr; uint64_t dims[5]; uint64_t strides[4]; uint32_t box[5]; int rank, dtype, swz; int pad;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = (const uint64_t*)&k;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

int ga_tensor_map(CUtensorMap_st* out, int dtype, int rank, const void* ptr, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      ga_set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(err));
      return GA_ERR_UNSUPPORTED;
    }
    g_encode = (EncodeTiledFn)fn;
  }
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rank = rank; key.dtype = dtype; key.swz = swizzle128;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return GA_OK; }
  cuuint64_t d[5], st[4]; cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUtensorMap m;
  CUresult r = g_encode(&m, dtype == GA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                        const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ga_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu stride1 %llu box %u %u %u ptr %p", (int)r,
                 rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                 (unsigned long long)strides_bytes[0], box[0], box[1], rank > 2 ? box[2] : 0, ptr);
    return GA_ERR_ALIGN;
  }
  if (g_maps.size() > 16384) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return GA_OK;
}
