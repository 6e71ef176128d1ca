// clm_api.cu — error string, version, device probe and the TMA descriptor encoder.
#include <stdarg.h>
#include <string.h>

#include "clm_common.cuh"

static thread_local char g_err[1024] = "";

void clm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* clm_last_error(void) { return g_err; }
extern "C" int clm_version(void) { return 100; }

extern "C" int clm_device_check(void) {
  int dev = 0;
  CLM_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CLM_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    clm_set_error("clm_b200 needs an sm_100a device (B200); found %s sm_%d%d", prop.name,
                  prop.major, prop.minor);
    return CLM_ERR_UNSUPPORTED;
  }
  return CLM_OK;
}

int clm_num_sms() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
  cached = n;
  return n;
}

// cuTensorMapEncodeTiled lives in libcuda; fetch it through the runtime so the library has
// no link-time dependency on the driver (it must load on a CPU-only box for the symbol test).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr) {
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int clm_make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                          uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    clm_set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return CLM_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    clm_set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch "
                  "(base=%p ld=%llu)", base, (unsigned long long)ld);
    return CLM_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (box_cols * 2 == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (box_cols * 2 == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (box_cols * 2 != 128) {
    clm_set_error("unsupported TMA box width %u", box_cols);
    return CLM_ERR_INVALID;
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    clm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu "
                  "box=%ux%u)", (int)r, (unsigned long long)rows, (unsigned long long)cols,
                  (unsigned long long)ld, box_cols, box_rows);
    return CLM_ERR_CUDA;
  }
  return CLM_OK;
}
