// clm_api.cu — error string, version, device probe and the TMA descriptor encoder.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include <stdlib.h>

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

int clm_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CLM_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on;
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

int clm_make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     int elem_bytes, uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    clm_set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return CLM_ERR_CUDA;
  }
  if (elem_bytes != 2 && elem_bytes != 4) {
    clm_set_error("TMA element size %d unsupported (bf16 = 2, fp32 = 4)", elem_bytes);
    return CLM_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * elem_bytes) % 16 != 0) {
    clm_set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch "
                  "(base=%p ld=%llu)", base, (unsigned long long)ld);
    return CLM_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const uint32_t row_bytes = box_cols * static_cast<uint32_t>(elem_bytes);
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (row_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (row_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (row_bytes != 128) {
    clm_set_error("unsupported TMA box width %u x %d bytes", box_cols, elem_bytes);
    return CLM_ERR_INVALID;
  }
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    clm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu "
                  "box=%ux%u elem=%d)", (int)r, (unsigned long long)rows, (unsigned long long)cols,
                  (unsigned long long)ld, box_cols, box_rows, elem_bytes);
    return CLM_ERR_CUDA;
  }
  return CLM_OK;
}

int clm_make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                          uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  return clm_make_tmap_2d(map, base, rows, cols, ld, 2, box_cols, box_rows);
}

// bf16 [d2][d1][d0] with byte strides of dims 1 and 2; box {box0 (64 = one 128-byte swizzled row), box1, 1}
int clm_make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    clm_set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return CLM_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || stride1_bytes % 16 != 0 || stride2_bytes % 16 != 0 ||
      box0 * 2 != 128) {
    clm_set_error("3-D TMA operand: 16-byte alignment / pitch and a 64-element box are required");
    return CLM_ERR_INVALID;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    clm_set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
    return CLM_ERR_CUDA;
  }
  return CLM_OK;
}

// ---------------------------------------------------------------------------------------
// launch accounting / event profiler
// ---------------------------------------------------------------------------------------
namespace {
struct ProfRec {
  int kind;
  double flops, bytes;
  cudaEvent_t a, b;
};
std::atomic<unsigned long long> g_launches{0};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfRec> g_recs;
thread_local ProfRec* g_open = nullptr;
thread_local ProfRec g_cur;
}  // namespace

void clm_prof_begin(int kind, double flops, double bytes, cudaStream_t s) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_open = nullptr;
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  g_cur.kind = kind;
  g_cur.flops = flops;
  g_cur.bytes = bytes;
  if (cudaEventCreate(&g_cur.a) != cudaSuccess || cudaEventCreate(&g_cur.b) != cudaSuccess) return;
  cudaEventRecord(g_cur.a, s);
  g_open = &g_cur;
}

void clm_prof_end(cudaStream_t s) {
  if (!g_open) return;
  cudaEventRecord(g_cur.b, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_recs.push_back(g_cur);
  g_open = nullptr;
}

extern "C" unsigned long long clm_launch_count(void) { return g_launches.load(); }

// a CUDA graph captured over library calls launches its kernels again on every replay without passing
// through the library: the host side adds the captured launch count per replay so the total stays true
extern "C" void clm_launch_count_add(long long n) { g_launches.fetch_add(static_cast<unsigned long long>(n)); }

extern "C" int clm_prof_is_enabled(void) { return g_prof_on.load() ? 1 : 0; }

extern "C" int clm_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  g_prof_on.store(on != 0);
  return CLM_OK;
}

extern "C" int clm_prof_summary(int kind, double* ms, double* flops, double* bytes, int* launches) {
  CLM_REQUIRE(kind >= 0 && kind < CLM_K_COUNT && ms && flops && bytes && launches,
              "clm_prof_summary: bad argument");
  CLM_CUDA_CHECK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double t = 0, f = 0, b = 0;
  int n = 0;
  for (auto& r : g_recs) {
    if (r.kind != kind) continue;
    float e = 0.f;
    CLM_CUDA_CHECK(cudaEventElapsedTime(&e, r.a, r.b));
    t += e;
    f += r.flops;
    b += r.bytes;
    ++n;
  }
  *ms = t; *flops = f; *bytes = b; *launches = n;
  return CLM_OK;
}

// Per-launch records (the launch list): out[4*i .. 4*i+3] = {kind, flops, bytes, ms}.
// Returns the number of records available (may exceed max_records; only max_records are written).
extern "C" int clm_prof_records(double* out, int max_records) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int i = 0;
  for (auto& r : g_recs) {
    if (out && i < max_records) {
      float e = 0.f;
      cudaEventElapsedTime(&e, r.a, r.b);
      out[4 * i + 0] = r.kind; out[4 * i + 1] = r.flops; out[4 * i + 2] = r.bytes; out[4 * i + 3] = e;
    }
    ++i;
  }
  return i;
}

// The launch list with start stamps: out[3*i .. 3*i+2] = {kind, start offset in ms from the first recorded
// launch's start event, duration in ms}.  Lets a caller see the gaps BETWEEN kernels of a step (start of launch
// i+1 minus end of launch i), i.e. what the sum of kernel durations does not account for.
extern "C" int clm_prof_timeline(double* out, int max_records) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int i = 0;
  for (auto& r : g_recs) {
    if (out && i < max_records) {
      float st = 0.f, e = 0.f;
      cudaEventElapsedTime(&st, g_recs.front().a, r.a);
      cudaEventElapsedTime(&e, r.a, r.b);
      out[3 * i + 0] = r.kind; out[3 * i + 1] = st; out[3 * i + 2] = e;
    }
    ++i;
  }
  return i;
}
