// Host-side TMA descriptor cache (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, so the
// library has no link-time dependency on libcuda).  The cache is immutable per key and guarded
// by a mutex: it is the only global state of libmmi_b200.
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace mmi {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    cudaGetLastError();
  }
  return fn;
}

bool encode_available() { return get_encode() != nullptr; }

struct MapKey {
  const void* ptr; uint64_t inner, outer, stride; uint32_t box_inner, box_outer; int swizzle;
  uint64_t batch;   // 0: 2-D map
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner &&
           box_outer == o.box_outer && swizzle == o.swizzle && batch == o.batch;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.stride); mix(k.box_inner); mix(k.box_outer); mix((uint64_t)k.swizzle); mix(k.batch);
    return h;
  }
};

static bool get_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t batch, uint64_t stride_elems, uint32_t box_inner,
                    uint32_t box_outer, CUtensorMapSwizzle swizzle, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, stride_elems, box_inner, box_outer, (int)swizzle, batch};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tma: cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t gdim[3] = {inner, outer, batch};
  cuuint64_t gstride[2] = {stride_elems * 2, outer * stride_elems * 2};
  cuuint32_t box[3] = {box_inner, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, batch ? 3 : 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tma: cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u swizzle=%d", (int)r, ptr,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride_elems, box_inner, box_outer, (int)swizzle);
    return false;
  }
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return true;
}

bool get_tensor_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride_elems, uint32_t box_inner,
                    uint32_t box_outer, CUtensorMapSwizzle swizzle, CUtensorMap* out) {
  return get_map(ptr, inner, outer, 0, stride_elems, box_inner, box_outer, swizzle, out);
}
bool get_tensor_map_3d(const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t stride_elems, uint32_t box_inner,
                       uint32_t box_rows, CUtensorMapSwizzle swizzle, CUtensorMap* out) {
  return get_map(ptr, inner, rows, batch, stride_elems, box_inner, box_rows, swizzle, out);
}

}  // namespace tc
}  // namespace mmi
