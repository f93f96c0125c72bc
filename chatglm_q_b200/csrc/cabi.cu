// C-ABI entry points (include/cgq.h): argument validation, kernel selection, error text,
// TMA descriptor cache.  No torch types anywhere in this library.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include <atomic>

#include "common.cuh"
#include "tmap.cuh"

namespace cgq {

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return CGQ_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static thread_local void* g_trace = nullptr;
void* take_trace_buffer() {
  void* t = g_trace;
  g_trace = nullptr;
  return t;
}

static int check_device() {
  static int ok[64] = {0};  // 0 unknown, 1 ok, -1 bad
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return CGQ_OK;
  if (ok[dev] == 0) {
    int major = 0;
    CGQ_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    ok[dev] = (major == 10) ? 1 : -1;
  }
  if (ok[dev] < 0) {
    set_error("cgq: device %d is not sm_100 (this library only carries sm_100a code)", dev);
    return CGQ_ERR_UNSUPPORTED;
  }
  return CGQ_OK;
}

// ------------------------------------------------------------------ tensor-map cache
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.dim0);
    mix(k.dim1);
    mix(k.stride1_bytes);
    mix((static_cast<uint64_t>(k.box0) << 32) | k.box1);
    mix((static_cast<uint64_t>(k.dtype) << 8) | static_cast<uint64_t>(k.swizzle));
    return h;
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int get_tmap_2d(const TmapKey& key, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return CGQ_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cgq: cuTensorMapEncodeTiled not available from the driver");
    return CGQ_ERR_CUDA;
  }
  cuuint64_t dims[2] = {key.dim0, key.dim1};
  cuuint64_t strides[1] = {key.stride1_bytes};
  cuuint32_t box[2] = {key.box0, key.box1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, static_cast<CUtensorMapDataType>(key.dtype), 2, const_cast<void*>(key.ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   static_cast<CUtensorMapSwizzle>(key.swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cgq: cuTensorMapEncodeTiled failed (%d) ptr=%p dims=(%llu,%llu) stride=%llu box=(%u,%u)",
              static_cast<int>(r), key.ptr, (unsigned long long)key.dim0,
              (unsigned long long)key.dim1, (unsigned long long)key.stride1_bytes, key.box0,
              key.box1);
    return CGQ_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return CGQ_OK;
}

// ------------------------------------------------------------------ validation shared by both GEMMs
static int check_common(const char* fn, const void* A, int64_t lda, const void* Wq,
                        const void* scale, const void* C, int64_t ldc, int M, int N, int K,
                        int dtype) {
  if (M < 0 || N <= 0 || K <= 0 || lda < K || ldc < N) {
    set_error("%s: bad shape M=%d N=%d K=%d lda=%lld ldc=%lld", fn, M, N, K, (long long)lda,
              (long long)ldc);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16 && dtype != CGQ_DTYPE_BF16) {
    set_error("%s: unsupported dtype code %d (0=f16, 1=bf16)", fn, dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  if (M > 0 && (A == nullptr || C == nullptr)) {
    set_error("%s: null A/C", fn);
    return CGQ_ERR_MISALIGNED;
  }
  if (Wq == nullptr || scale == nullptr) {
    set_error("%s: null weight/scale", fn);
    return CGQ_ERR_MISALIGNED;
  }
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C) |
       reinterpret_cast<uintptr_t>(scale)) & 1) {
    set_error("%s: A/C/scale must be 2-byte aligned", fn);
    return CGQ_ERR_MISALIGNED;
  }
  return CGQ_OK;
}

static int check_workspace(const char* fn, void* ws, size_t bytes) {
  if (ws == nullptr || bytes < kWorkspaceBytes) {
    set_error("%s: workspace %p / %zu bytes, need %zu zero-initialised bytes", fn, ws, bytes,
              kWorkspaceBytes);
    return CGQ_ERR_WORKSPACE;
  }
  return CGQ_OK;
}

}  // namespace cgq

using namespace cgq;

extern "C" int cgq_version(void) { return (0 << 16) | 2; }
extern "C" const char* cgq_last_error(void) { return g_err; }
extern "C" size_t cgq_workspace_bytes(void) { return kWorkspaceBytes; }
extern "C" void cgq_debug_trace(void* device_buffer) { g_trace = device_buffer; }
extern "C" int cgq_set_decode_arith(int arith) { return default_w4_arith(arith); }

// The shape-general CUDA-core kernels are a correctness net for shapes the TMA kernels cannot take (N % 16 != 0,
// misaligned pointers, missing workspace), not a path a real layer should ever run on: AUTO counts every time it takes
// them, and refuses when told to (CGQ_FORBID_SIMPLE=1 or cgq_forbid_simple(1)).
namespace {
std::atomic<unsigned long long> g_simple_auto{0};
std::atomic<int> g_forbid_simple{[] {
  const char* e = getenv("CGQ_FORBID_SIMPLE");
  return (e != nullptr && atoi(e) != 0) ? 1 : 0;
}()};
int take_simple(const char* fn, int M, int N, int K) {
  if (g_forbid_simple.load() != 0) {
    set_error("%s: M=%d N=%d K=%d is not taken by the TMA / tcgen05 kernels (needs N %% 16 == 0, 16-byte aligned "
              "pointers, lda %% 8 == 0, the workspace) and the CUDA-core kernel is forbidden (CGQ_FORBID_SIMPLE)",
              fn, M, N, K);
    return CGQ_ERR_BAD_SHAPE;
  }
  g_simple_auto.fetch_add(1);
  return CGQ_OK;
}
}  // namespace
extern "C" unsigned long long cgq_simple_fallback_count(void) { return g_simple_auto.load(); }
extern "C" int cgq_forbid_simple(int on) { return g_forbid_simple.exchange(on != 0 ? 1 : 0); }

extern "C" int cgq_w4a16_gemm_ex(const void* A, int64_t lda, const uint8_t* Wq, const void* scale,
                                 const void* bias, void* C, int64_t ldc, int M, int N, int K,
                                 int group, int dtype, void* workspace, size_t workspace_bytes,
                                 void* stream, int impl) {
  const char* fn = "cgq_w4a16_gemm";
  int rc = check_common(fn, A, lda, Wq, scale, C, ldc, M, N, K, dtype);
  if (rc != CGQ_OK) return rc;
  if (group != 32 || K % 32 != 0) {
    set_error("%s: group must be 32 and divide K (group=%d, K=%d)", fn, group, K);
    return CGQ_ERR_BAD_SHAPE;
  }
  rc = check_device();
  if (rc != CGQ_OK) return rc;
  if (M == 0) return CGQ_OK;
  GemmArgs a{A, lda, Wq, scale, bias, C, ldc, M, N, K, dtype, workspace,
             static_cast<cudaStream_t>(stream)};
  if (impl == CGQ_IMPL_AUTO) {
    if (M <= 8 && w4_gemv_supported(a) && workspace != nullptr && workspace_bytes >= kWorkspaceBytes)
      impl = CGQ_IMPL_GEMV;
    else if (M > 8 && w4_tc_supported(a))
      impl = CGQ_IMPL_TC;
    else {
      rc = take_simple(fn, M, N, K);
      if (rc != CGQ_OK) return rc;
      impl = CGQ_IMPL_SIMPLE;
    }
  }
  switch (impl) {
    case CGQ_IMPL_SIMPLE:
      return launch_w4_simple(a);
    case CGQ_IMPL_GEMV_UMMA: {
      if (M != 1 || !w4_gemv_supported(a)) {
        set_error("%s: tcgen05 decode kernel needs M==1, N%%16==0, 16-byte aligned pointers", fn);
        return CGQ_ERR_MISALIGNED;
      }
      bool taken = false;
      rc = launch_w4_gemv_umma(a, &taken);
      if (rc != CGQ_OK) return rc;
      if (!taken) {
        set_error("%s: shape too large for the tcgen05 decode kernel's shared-memory rings", fn);
        return CGQ_ERR_BAD_SHAPE;
      }
      return CGQ_OK;
    }
    case CGQ_IMPL_TC:
      if (!w4_tc_supported(a)) {
        set_error("%s: tcgen05 kernel needs K%%32==0, N%%16==0, 16-byte aligned pointers, lda%%8==0", fn);
        return CGQ_ERR_MISALIGNED;
      }
      if (workspace_bytes < kWorkspaceBytes) a.workspace = nullptr;   // no split-K without the partial-tile workspace
      return launch_w4_tc(a);
    case CGQ_IMPL_GEMV:
    case CGQ_IMPL_GEMV_EXACT:
    case CGQ_IMPL_GEMV_SUBNORMAL:
    case CGQ_IMPL_GEMV_IMMA:
      if (M > 8 || !w4_gemv_supported(a)) {
        set_error("%s: GEMV kernel needs M<=8, N%%16==0, 16-byte aligned pointers, lda%%8==0", fn);
        return CGQ_ERR_MISALIGNED;
      }
      rc = check_workspace(fn, workspace, workspace_bytes);
      if (rc != CGQ_OK) return rc;
      return launch_w4_gemv(a, impl == CGQ_IMPL_GEMV_EXACT       ? W4_ARITH_EXACT
                               : impl == CGQ_IMPL_GEMV_SUBNORMAL ? W4_ARITH_SUBNORMAL
                               : impl == CGQ_IMPL_GEMV_IMMA      ? W4_ARITH_IMMA
                                                                 : W4_ARITH_DEFAULT);
    default:
      set_error("%s: unknown impl %d", fn, impl);
      return CGQ_ERR_BAD_SHAPE;
  }
}

extern "C" int cgq_w4a16_gemm(const void* A, int64_t lda, const uint8_t* Wq, const void* scale,
                              const void* bias, void* C, int64_t ldc, int M, int N, int K,
                              int group, int dtype, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return cgq_w4a16_gemm_ex(A, lda, Wq, scale, bias, C, ldc, M, N, K, group, dtype, workspace,
                           workspace_bytes, stream, CGQ_IMPL_AUTO);
}

extern "C" int cgq_prefetch_next_w4(const uint8_t* Wq, const void* scale, int N, int K) {
  if (Wq == nullptr) {
    set_next_w4_hint(nullptr, nullptr, 0, 0);
    return CGQ_OK;
  }
  if (scale == nullptr || N <= 0 || K <= 0 || K % 32 != 0 || N % 16 != 0 ||
      ((reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(scale)) & 15)) {
    set_error("cgq_prefetch_next_w4: needs 16-byte aligned Wq / scale, N%%16==0, K%%32==0 (N=%d K=%d)", N, K);
    return CGQ_ERR_MISALIGNED;
  }
  set_next_w4_hint(Wq, scale, N, K);
  return CGQ_OK;
}

extern "C" int cgq_handover_next(const uint32_t* wait_ctr, uint32_t wait_count, uint32_t* signal_ctr) {
  if (((reinterpret_cast<uintptr_t>(wait_ctr) | reinterpret_cast<uintptr_t>(signal_ctr)) & 3) ||
      (wait_ctr != nullptr && wait_count == 0)) {
    set_error("cgq_handover_next: counters must be 4-byte aligned and wait_count > 0 with a wait counter");
    return CGQ_ERR_MISALIGNED;
  }
  set_w4_handover(wait_ctr, wait_count, signal_ctr);
  return CGQ_OK;
}
extern "C" int cgq_w4_gemv_tiles(int N) { return N > 0 ? w4_gemv_tiles(N) : 0; }

extern "C" int cgq_w4a16_gemv_fused(const void* A, const uint8_t* Wq, const void* scale,
                                    const void* bias, const void* resid, void* C, int N, int K,
                                    int group, int dtype, int prologue, const void* norm_w,
                                    float eps, void* stream) {
  const char* fn = "cgq_w4a16_gemv_fused";
  const int a_len = prologue == CGQ_PRO_SILU_GATE ? 2 * K : K;
  int rc = check_common(fn, A, a_len, Wq, scale, C, N, 1, N, K, dtype);
  if (rc != CGQ_OK) return rc;
  if (group != 32 || K % 32 != 0) {
    set_error("%s: group must be 32 and divide K (group=%d, K=%d)", fn, group, K);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (prologue != CGQ_PRO_NONE && prologue != CGQ_PRO_RMSNORM && prologue != CGQ_PRO_SILU_GATE) {
    set_error("%s: unknown prologue %d", fn, prologue);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (prologue == CGQ_PRO_RMSNORM &&
      (norm_w == nullptr || (reinterpret_cast<uintptr_t>(norm_w) & 15))) {
    set_error("%s: CGQ_PRO_RMSNORM needs a 16-byte aligned norm weight", fn);
    return CGQ_ERR_MISALIGNED;
  }
  rc = check_device();
  if (rc != CGQ_OK) return rc;
  GemmArgs a{A, a_len, Wq, scale, bias, C, N, 1, N, K, dtype, nullptr,
             static_cast<cudaStream_t>(stream)};
  if (!w4_gemv_supported(a)) {
    set_error("%s: needs N%%16==0 and 16-byte aligned A / Wq / scale", fn);
    return CGQ_ERR_MISALIGNED;
  }
  GemvFused fu{prologue, norm_w, eps, resid};
  return launch_w4_gemv_fused(a, fu);
}

extern "C" int cgq_w8a16_gemv_fused(const void* A, const int8_t* Wq, const void* scale, const void* bias,
                                    const void* resid, void* C, int N, int K, int dtype, int prologue,
                                    const void* norm_w, float eps, void* stream) {
  const char* fn = "cgq_w8a16_gemv_fused";
  const int a_len = prologue == CGQ_PRO_SILU_GATE ? 2 * K : K;
  int rc = check_common(fn, A, a_len, Wq, scale, C, N, 1, N, K, dtype);
  if (rc != CGQ_OK) return rc;
  if (prologue != CGQ_PRO_NONE && prologue != CGQ_PRO_RMSNORM && prologue != CGQ_PRO_SILU_GATE) {
    set_error("%s: unknown prologue %d", fn, prologue);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (prologue == CGQ_PRO_RMSNORM && (norm_w == nullptr || (reinterpret_cast<uintptr_t>(norm_w) & 15))) {
    set_error("%s: CGQ_PRO_RMSNORM needs a 16-byte aligned norm weight", fn);
    return CGQ_ERR_MISALIGNED;
  }
  rc = check_device();
  if (rc != CGQ_OK) return rc;
  GemmArgs a{A, a_len, Wq, scale, bias, C, N, 1, N, K, dtype, nullptr, static_cast<cudaStream_t>(stream)};
  if (!w8_gemv_supported(a)) {
    set_error("%s: needs K%%16==0 and 16-byte aligned A / Wq", fn);
    return CGQ_ERR_MISALIGNED;
  }
  GemvFused fu{prologue, norm_w, eps, resid};
  return launch_w8_gemv_fused(a, fu);
}

extern "C" int cgq_w8a16_gemm_ex(const void* A, int64_t lda, const int8_t* Wq, const void* scale,
                                 const void* bias, void* C, int64_t ldc, int M, int N, int K,
                                 int dtype, void* workspace, size_t workspace_bytes, void* stream,
                                 int impl) {
  const char* fn = "cgq_w8a16_gemm";
  int rc = check_common(fn, A, lda, Wq, scale, C, ldc, M, N, K, dtype);
  if (rc != CGQ_OK) return rc;
  rc = check_device();
  if (rc != CGQ_OK) return rc;
  if (M == 0) return CGQ_OK;
  GemmArgs a{A, lda, Wq, scale, bias, C, ldc, M, N, K, dtype, workspace,
             static_cast<cudaStream_t>(stream)};
  if (impl == CGQ_IMPL_AUTO) {
    if (M <= 8 && w8_gemv_supported(a) && workspace != nullptr && workspace_bytes >= kWorkspaceBytes)
      impl = CGQ_IMPL_GEMV;
    else if (M > 8 && w8_tc_supported(a))
      impl = CGQ_IMPL_TC;
    else {
      rc = take_simple(fn, M, N, K);
      if (rc != CGQ_OK) return rc;
      impl = CGQ_IMPL_SIMPLE;
    }
  }
  switch (impl) {
    case CGQ_IMPL_SIMPLE:
      return launch_w8_simple(a);
    case CGQ_IMPL_TC:
      if (!w8_tc_supported(a)) {
        set_error("%s: tcgen05 kernel needs K%%16==0, 16-byte aligned pointers, lda%%8==0", fn);
        return CGQ_ERR_MISALIGNED;
      }
      if (workspace_bytes < kWorkspaceBytes) a.workspace = nullptr;
      return launch_w8_tc(a);
    case CGQ_IMPL_GEMV:
      if (M > 8 || !w8_gemv_supported(a)) {
        set_error("%s: GEMV kernel needs M<=8, K%%16==0, 16-byte aligned pointers, lda%%8==0", fn);
        return CGQ_ERR_MISALIGNED;
      }
      rc = check_workspace(fn, workspace, workspace_bytes);
      if (rc != CGQ_OK) return rc;
      return launch_w8_gemv(a);
    default:
      set_error("%s: unknown impl %d", fn, impl);
      return CGQ_ERR_BAD_SHAPE;
  }
}

extern "C" int cgq_w8a16_gemm(const void* A, int64_t lda, const int8_t* Wq, const void* scale,
                              const void* bias, void* C, int64_t ldc, int M, int N, int K,
                              int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  return cgq_w8a16_gemm_ex(A, lda, Wq, scale, bias, C, ldc, M, N, K, dtype, workspace,
                           workspace_bytes, stream, CGQ_IMPL_AUTO);
}
