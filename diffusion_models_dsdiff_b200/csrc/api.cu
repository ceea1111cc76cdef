// Library plumbing: error text, launch counter, driver entry point for TMA descriptor encoding.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int resolve_driver() {
  if (g_encode) return 0;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return 3;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

int encode_tmap(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                const uint32_t *box, const uint32_t *elem_strides, bool swizzle128) {
  if (int rc = resolve_driver()) return rc;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  // 16-bit elements either way (fp16 / bf16): the type only matters for OOB fill, which is zero.
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx,
                        es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] "
              "estr [%u,%u,%u,%u] stride0 %llu base %p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, elem_strides[0],
              rank > 1 ? elem_strides[1] : 0, rank > 2 ? elem_strides[2] : 0, rank > 3 ? elem_strides[3] : 0,
              (unsigned long long)strides_bytes[0], base);
    return 4;
  }
  return 0;
}

namespace {
std::mutex g_tmap_mu;
std::unordered_map<std::string, CUtensorMap> g_tmap_cache;
}  // namespace

// Descriptor encoding costs microseconds on the host; layers are launched with the same buffers every step
// (always under CUDA-graph replay, usually also eagerly), so descriptors are cached by their full definition.
int cached_tmap(CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides,
                const uint32_t *box, const uint32_t *estr, bool swizzle128) {
  std::string key(reinterpret_cast<const char *>(&base), sizeof(base));
  key.append(reinterpret_cast<const char *>(&rank), sizeof(rank));
  key.append(reinterpret_cast<const char *>(dims), sizeof(uint64_t) * rank);
  key.append(reinterpret_cast<const char *>(strides), sizeof(uint64_t) * (rank - 1));
  key.append(reinterpret_cast<const char *>(box), sizeof(uint32_t) * rank);
  key.append(reinterpret_cast<const char *>(estr), sizeof(uint32_t) * rank);
  key.push_back(swizzle128 ? 1 : 0);
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) {
    *out = it->second;
    return 0;
  }
  if (int rc = encode_tmap(out, base, rank, dims, strides, box, estr, swizzle128)) return rc;
  if (g_tmap_cache.size() > 200000) g_tmap_cache.clear();
  g_tmap_cache.emplace(std::move(key), *out);
  return 0;
}

}  // namespace b200

extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" const char *b200_last_error(void) { return b200::g_err; }
extern "C" uint64_t b200_launch_count(void) { return b200::g_launches.load(); }

extern "C" int b200_init(void) {
  using namespace b200;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  B200_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B200_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B200_CHECK(major == 10, "libb200diff is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)",
             dev, major, minor);
  return resolve_driver();
}
