// Host-side plumbing shared by every C-ABI entry point: last-error string, device check,
// SM count cache and the TMA tensor-map encoder (driver entry point fetched at run time so the
// library links against libcudart only and loads on a GPU-less build box).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cudaTypedefs.h>

#include "common.cuh"

namespace b200ssl {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static int g_pdl = 0;  // measured on B200 inside the captured step: 57.7 ms with, 56.9 ms without -> off by default
bool pdl_enabled() { return g_pdl != 0; }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  B200SSL_CHECK(fn != nullptr, -3, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  B200SSL_CHECK(rank >= 1 && rank <= 5, -2, "tensor map rank %d unsupported", rank);
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstrides[i - 1] = strides_bytes[i];
      B200SSL_CHECK(strides_bytes[i] % 16 == 0, -2, "tensor map stride %llu B not 16B aligned (dim %d)",
                    static_cast<unsigned long long>(strides_bytes[i]), i);
    }
  }
  const CUtensorMapDataType dt =
      elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstrides, gbox,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200SSL_CHECK(r == CUDA_SUCCESS, -3,
                "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu box %u,%u)",
                static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
                static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
  return 0;
}

}  // namespace b200ssl

extern "C" const char* b200ssl_last_error(void) { return b200ssl::g_last_error; }

extern "C" int b200ssl_version(void) { return 100; }

// 1 = launch the GEMM / LayerNorm / attention kernels with programmatic dependent launch, so each one's set-up
// overlaps the tail of its predecessor in the stream; 0 (default: the overlap did not pay, see above) = plain order.
extern "C" int b200ssl_set_pdl(int on) {
  b200ssl::g_pdl = on ? 1 : 0;
  return 0;
}

// 0 when the current device can run the library (compute capability 10.x), negative otherwise.
extern "C" int b200ssl_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    b200ssl::set_last_error("no CUDA device is visible");
    return -1;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    b200ssl::set_last_error("b200ssl kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return -1;
  }
  return 0;
}
