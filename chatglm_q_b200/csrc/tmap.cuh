// Host-side TMA descriptor (CUtensorMap) construction with a small cache.  The driver entry
// point is resolved through the runtime (cudaGetDriverEntryPoint) so the library does not link
// libcuda directly.  A descriptor depends only on (pointer, shape, box, swizzle): re-using a cached
// one for a recycled pointer with the same geometry is therefore always valid.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cgq {

struct TmapKey {
  const void* ptr;
  uint64_t dim0, dim1;      // elements; dim0 innermost
  uint64_t stride1_bytes;   // byte stride of dim1
  uint32_t box0, box1;
  int dtype;                // CUtensorMapDataType
  int swizzle;              // CUtensorMapSwizzle
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && dim0 == o.dim0 && dim1 == o.dim1 && stride1_bytes == o.stride1_bytes &&
           box0 == o.box0 && box1 == o.box1 && dtype == o.dtype && swizzle == o.swizzle;
  }
};

// Returns CGQ_OK and fills *out, or a CGQ_ERR_* code (error text set).
int get_tmap_2d(const TmapKey& key, CUtensorMap* out);

}  // namespace cgq
