// clm_elementwise.cu — the HBM-bound kernels of the encoder: LayerNorm, L2-normalise,
// token/patch embedding, pooling.  One warp owns one row; rows are read and written with
// 128-bit accesses that are contiguous across the warp; statistics are fp32 and use the
// two-pass (mean, then centred variance) form so they track torch's nn.LayerNorm closely.
#include <stdlib.h>

#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int kWarpsPerBlock = 8;

// One row of `NV*128` floats held as NV float4 per lane (lane-interleaved: coalesced).
template <int NV>
struct Row {
  float4 v[NV];
  __device__ __forceinline__ void load(const float* p) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = p4[j * 32 + lane_id()];
  }
  // the bf16 residual stream: 4 bf16 (8 bytes) per lane and vector, same column ownership as the fp32 form
  __device__ __forceinline__ void load_bf16(const __nv_bfloat16* p) {
    const uint2* p2 = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint2 u = p2[j * 32 + lane_id()];
      v[j].x = __uint_as_float(u.x << 16); v[j].y = __uint_as_float(u.x & 0xffff0000u);
      v[j].z = __uint_as_float(u.y << 16); v[j].w = __uint_as_float(u.y & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void add(const float* p) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 a = __ldg(p4 + j * 32 + lane_id());
      v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
    }
  }
  // (x - mean) * rstd * gamma + beta, in place
  __device__ __forceinline__ void layernorm(const float* gamma, const float* beta, float eps) {
    constexpr float inv_n = 1.0f / (NV * 128);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 g = __ldg(g4 + j * 32 + lane_id());
      const float4 b = __ldg(b4 + j * 32 + lane_id());
      v[j].x = (v[j].x - mean) * rstd * g.x + b.x;
      v[j].y = (v[j].y - mean) * rstd * g.y + b.y;
      v[j].z = (v[j].z - mean) * rstd * g.z + b.z;
      v[j].w = (v[j].w - mean) * rstd * g.w + b.w;
    }
  }
  __device__ __forceinline__ void store_f32(float* p) const {
    float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int j = 0; j < NV; ++j) p4[j * 32 + lane_id()] = v[j];
  }
  __device__ __forceinline__ void store_bf16(__nv_bfloat16* p) const {
    uint2* p2 = reinterpret_cast<uint2*>(p);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      uint2 o;
      o.x = pack_bf16x2(v[j].x, v[j].y);
      o.y = pack_bf16x2(v[j].z, v[j].w);
      p2[j * 32 + lane_id()] = o;
    }
  }
};

// kH16: the residual stream (x / h below) is bf16 instead of fp32
template <int NV, bool kH16 = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows, float eps) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int D = NV * 128;
  Row<NV> r;
  if (kH16) r.load_bf16(static_cast<const __nv_bfloat16*>(x) + static_cast<size_t>(row) * D);
  else r.load(static_cast<const float*>(x) + static_cast<size_t>(row) * D);
  r.layernorm(gamma, beta, eps);
  r.store_bf16(y + static_cast<size_t>(row) * D);
}

// (mean, rstd) of every row of a bf16 residual stream: what a GEMM with the LayerNorm folded in (clm_gemm_ln_epi)
// needs of the row -- half the traffic of the LayerNorm pass it replaces (no normalised copy is written)
template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
row_stats_kernel(const __nv_bfloat16* __restrict__ h, float2* __restrict__ stats, int rows, float eps, int reverse) {
  pdl_wait();
  pdl_trigger();
  // reverse: the first blocks take the LAST rows, so the pass ends on rows 0.. and leaves those in the L2 -- the rows the
  // GEMM that follows reads first (and it starts on the rows the reduce-add GEMM before it touched last)
  int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (reverse) row = rows - 1 - row;
  constexpr int D = NV * 128;
  constexpr float inv_n = 1.0f / D;
  Row<NV> r;
  r.load_bf16(h + static_cast<size_t>(row) * D);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += (r.v[j].x + r.v[j].y) + (r.v[j].z + r.v[j].w);
  const float mean = warp_sum(s) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = r.v[j].x - mean, b = r.v[j].y - mean, c = r.v[j].z - mean, d = r.v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
  if (lane_id() == 0) stats[row] = make_float2(mean, rstd);
}

template <int NV, bool kH16 = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pool_ln_kernel(const void* __restrict__ h, const int32_t* __restrict__ row_idx,
               const float* __restrict__ gamma, const float* __restrict__ beta,
               __nv_bfloat16* __restrict__ y, int batch, int tokens, float eps) {
  const int b = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (b >= batch) return;
  constexpr int D = NV * 128;
  const int t = row_idx ? row_idx[b] : 0;
  Row<NV> r;
  const size_t off = (static_cast<size_t>(b) * tokens + t) * D;
  if (kH16) r.load_bf16(static_cast<const __nv_bfloat16*>(h) + off);
  else r.load(static_cast<const float*>(h) + off);
  r.layernorm(gamma, beta, eps);
  r.store_bf16(y + static_cast<size_t>(b) * D);
}

template <int NV, bool kH16 = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
vision_embed_ln_kernel(const float* __restrict__ patch_out, const float* __restrict__ class_emb,
                       const float* __restrict__ pos_emb, const float* __restrict__ gamma,
                       const float* __restrict__ beta, void* __restrict__ h, int batch, int np,
                       float eps) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int tokens = np + 1;
  if (row >= batch * tokens) return;
  constexpr int D = NV * 128;
  const int b = row / tokens;
  const int t = row - b * tokens;
  Row<NV> r;
  if (t == 0) r.load(class_emb);
  else r.load(patch_out + (static_cast<size_t>(b) * np + (t - 1)) * D);
  r.add(pos_emb + static_cast<size_t>(t) * D);
  r.layernorm(gamma, beta, eps);
  if (kH16) r.store_bf16(static_cast<__nv_bfloat16*>(h) + static_cast<size_t>(row) * D);
  else r.store_f32(static_cast<float*>(h) + static_cast<size_t>(row) * D);
}

template <int NV, bool kH16 = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_text_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tok_emb,
                  const float* __restrict__ pos_emb, void* __restrict__ h, int rows, int tokens,
                  int vocab) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int D = NV * 128;
  const int t = row % tokens;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  Row<NV> r;
  r.load(tok_emb + static_cast<size_t>(id) * D);
  r.add(pos_emb + static_cast<size_t>(t) * D);
  if (kH16) r.store_bf16(static_cast<__nv_bfloat16*>(h) + static_cast<size_t>(row) * D);
  else r.store_f32(static_cast<float*>(h) + static_cast<size_t>(row) * D);
}

// eos_pos[b] = first t with ids[b,t] == eos_id, else 0 (argmax of an all-zero mask)
__global__ void eos_pos_kernel(const int32_t* __restrict__ ids, int32_t* __restrict__ eos_pos,
                               int batch, int tokens, int eos_id) {
  const int b = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (b >= batch) return;
  int first = 0x7fffffff;
  for (int t = lane_id(); t < tokens; t += 32)
    if (ids[static_cast<size_t>(b) * tokens + t] == eos_id) first = min(first, t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  if (lane_id() == 0) eos_pos[b] = (first == 0x7fffffff) ? 0 : first;
}

// generic-width L2 normalise (dim multiple of 4): warp per row, no epsilon (matches the reference)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
              int rows, int dim) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * dim);
  const int n4 = dim >> 2;
  float s = 0.f;
  for (int i = lane_id(); i < n4; i += 32) {
    const float4 a = x4[i];
    s += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
  }
  const float nrm = sqrtf(warp_sum(s));
  float4* y4 = reinterpret_cast<float4*>(y + static_cast<size_t>(row) * dim);
  uint2* yb2 = yb ? reinterpret_cast<uint2*>(yb + static_cast<size_t>(row) * dim) : nullptr;
  for (int i = lane_id(); i < n4; i += 32) {
    float4 a = x4[i];
    a.x /= nrm; a.y /= nrm; a.z /= nrm; a.w /= nrm;
    y4[i] = a;
    if (yb2) {
      uint2 o;
      o.x = pack_bf16x2(a.x, a.y);
      o.y = pack_bf16x2(a.z, a.w);
      yb2[i] = o;
    }
  }
}

// seeker query fusion: v = wa*a (+ wb*b), out = v/||v|| (seeker_service.py:146-157); warp per row
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
fuse_normalize_kernel(const float* __restrict__ a, float wa, const float* __restrict__ b, float wb,
                      float* __restrict__ y, __nv_bfloat16* __restrict__ yb, int rows, int dim) {
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* a4 = reinterpret_cast<const float4*>(a + static_cast<size_t>(row) * dim);
  const float4* b4 = b ? reinterpret_cast<const float4*>(b + static_cast<size_t>(row) * dim) : nullptr;
  const int n4 = dim >> 2;
  auto fused = [&](int i) {
    float4 v = a4[i];
    // the reference computes w*e per modality and then sums (python sum: 0 + w_t*e_t + w_i*e_i):
    // separate roundings of the two products, no FMA contraction
    v.x = __fmul_rn(wa, v.x); v.y = __fmul_rn(wa, v.y); v.z = __fmul_rn(wa, v.z); v.w = __fmul_rn(wa, v.w);
    if (b4) {
      const float4 u = b4[i];
      v.x = __fadd_rn(v.x, __fmul_rn(wb, u.x)); v.y = __fadd_rn(v.y, __fmul_rn(wb, u.y));
      v.z = __fadd_rn(v.z, __fmul_rn(wb, u.z)); v.w = __fadd_rn(v.w, __fmul_rn(wb, u.w));
    }
    return v;
  };
  float s = 0.f;
  for (int i = lane_id(); i < n4; i += 32) {
    const float4 v = fused(i);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  const float nrm = sqrtf(warp_sum(s));
  float4* y4 = reinterpret_cast<float4*>(y + static_cast<size_t>(row) * dim);
  uint2* yb2 = yb ? reinterpret_cast<uint2*>(yb + static_cast<size_t>(row) * dim) : nullptr;
  for (int i = lane_id(); i < n4; i += 32) {
    float4 v = fused(i);
    v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm;
    y4[i] = v;
    if (yb2) {
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      yb2[i] = o;
    }
  }
}

// im2col for a stride==kernel conv: out[(b*g*g + gy*g + gx), c*P*P + py*P + px] =
// pix[b, c, gy*P+py, gx*P+px].  One thread produces 8 consecutive output columns (16 B).
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ pix, __nv_bfloat16* __restrict__ out, int batch, int image,
              int patch, int kpad) {
  const int grid_w = image / patch;
  const int k = 3 * patch * patch;
  const int chunks_per_row = kpad >> 3;
  const long long total = static_cast<long long>(batch) * grid_w * grid_w * chunks_per_row;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % chunks_per_row);
  const long long prow = idx / chunks_per_row;
  const int gx = static_cast<int>(prow % grid_w);
  const int gy = static_cast<int>((prow / grid_w) % grid_w);
  const int b = static_cast<int>(prow / (grid_w * grid_w));
  // one division per thread, not per element: (channel, patch row, patch column) of the chunk's first column,
  // then walk; a patch width that is a multiple of 8 never wraps inside a chunk and loads two aligned float4
  float f[8];
  const int pp = patch * patch;
  int col = chunk * 8;
  int c = col / pp;
  const int rem = col - c * pp;
  int py = rem / patch;
  int px = rem - py * patch;
  auto row_ptr = [&](int cc, int yy) {
    return pix + ((static_cast<size_t>(b) * 3 + cc) * image + (gy * patch + yy)) * image + gx * patch;
  };
  if ((patch & 7) == 0 && (image & 3) == 0 && (reinterpret_cast<uintptr_t>(pix) & 15) == 0 && col + 8 <= k) {
    const float4* src = reinterpret_cast<const float4*>(row_ptr(c, py) + px);
    const float4 lo = __ldg(src), hi = __ldg(src + 1);
    f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w;
    f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
  } else {
    const float* base = (col < k) ? row_ptr(c, py) : pix;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f[i] = (col + i < k) ? __ldg(base + px) : 0.f;
      if (++px == patch) {
        px = 0;
        if (++py == patch) { py = 0; ++c; }
        if (col + i + 1 < k) base = row_ptr(c, py);
      }
    }
  }
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  reinterpret_cast<uint4*>(out)[idx] = o;
}

inline int blocks_for(int rows) { return (rows + kWarpsPerBlock - 1) / kWarpsPerBlock; }

#define CLM_DISPATCH_DIM(dim, KERNEL_CALL)                                   \
  switch (dim) {                                                             \
    case 512: { constexpr int NV = 4; KERNEL_CALL; break; }                  \
    case 768: { constexpr int NV = 6; KERNEL_CALL; break; }                  \
    case 1024: { constexpr int NV = 8; KERNEL_CALL; break; }                 \
    case 256: { constexpr int NV = 2; KERNEL_CALL; break; }                  \
    case 128: { constexpr int NV = 1; KERNEL_CALL; break; }                  \
    default:                                                                 \
      clm_set_error("unsupported width %d (supported: 128,256,512,768,1024)", dim); \
      return CLM_ERR_UNSUPPORTED;                                            \
  }

}  // namespace

extern "C" int clm_layernorm_ex(const void* x, int x_dtype, const float* gamma, const float* beta, void* y_bf16,
                                int rows, int dim, float eps, void* stream) {
  CLM_REQUIRE(x && gamma && beta && y_bf16 && rows >= 0, "clm_layernorm: bad argument");
  CLM_REQUIRE(x_dtype == CLM_OUT_F32 || x_dtype == CLM_OUT_BF16, "clm_layernorm: x_dtype must be CLM_OUT_F32 / CLM_OUT_BF16");
  if (rows == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool h16 = x_dtype == CLM_OUT_BF16;
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (h16 ? 4.0 : 6.0) * rows * dim, s);
  if (h16) {
    CLM_DISPATCH_DIM(dim, (clm_launch_pdl(layernorm_kernel<NV, true>, dim3(blocks_for(rows)), dim3(kWarpsPerBlock * 32),
                                          0, s, x, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16), rows, eps)));
  } else {
    CLM_DISPATCH_DIM(dim, (clm_launch_pdl(layernorm_kernel<NV, false>, dim3(blocks_for(rows)), dim3(kWarpsPerBlock * 32),
                                          0, s, x, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16), rows, eps)));
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

int clm_row_stats_dir(const void* h_bf16, float* stats, int rows, int dim, float eps, void* stream, int reverse);

extern "C" int clm_row_stats(const void* h_bf16, float* stats, int rows, int dim, float eps, void* stream) {
  return clm_row_stats_dir(h_bf16, stats, rows, dim, eps, stream, 0);
}

// reverse: block order only (see row_stats_kernel); the tower alternates the direction of consecutive kernels
int clm_row_stats_dir(const void* h_bf16, float* stats, int rows, int dim, float eps, void* stream, int reverse) {
  CLM_REQUIRE(h_bf16 && stats && rows >= 0, "clm_row_stats: bad argument");
  CLM_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 7) == 0, "clm_row_stats: stats must be 8-byte aligned");
  if (rows == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 2.0 * rows * dim + 8.0 * rows, s);
  CLM_DISPATCH_DIM(dim, (clm_launch_pdl(row_stats_kernel<NV>, dim3(blocks_for(rows)), dim3(kWarpsPerBlock * 32), 0, s,
                                        static_cast<const __nv_bfloat16*>(h_bf16), reinterpret_cast<float2*>(stats),
                                        rows, eps, reverse)));
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16,
                             int rows, int dim, float eps, void* stream) {
  return clm_layernorm_ex(x, CLM_OUT_F32, gamma, beta, y_bf16, rows, dim, eps, stream);
}

extern "C" int clm_l2norm(const float* x, float* y, void* y_bf16_or_null, int rows, int dim,
                          void* stream) {
  CLM_REQUIRE(x && y && rows >= 0 && dim > 0 && dim % 4 == 0, "clm_l2norm: bad argument");
  if (rows == 0) return CLM_OK;
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * rows * dim, static_cast<cudaStream_t>(stream));
  l2norm_kernel<<<blocks_for(rows), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, y, static_cast<__nv_bfloat16*>(y_bf16_or_null), rows, dim);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_fuse_normalize(const float* a, float wa, const float* b_or_null, float wb, float* out,
                                  void* out_bf16_or_null, int rows, int dim, void* stream) {
  if (rows == 0) return CLM_OK;
  CLM_REQUIRE(a && out && rows > 0 && dim > 0 && dim % 4 == 0, "clm_fuse_normalize: bad argument");
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (b_or_null ? 12.0 : 8.0) * rows * dim, static_cast<cudaStream_t>(stream));
  fuse_normalize_kernel<<<blocks_for(rows), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      a, wa, b_or_null, wb, out, static_cast<__nv_bfloat16*>(out_bf16_or_null), rows, dim);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_embed_text(const int32_t* ids, const float* tok_emb, const float* pos_emb,
                              float* h, int32_t* eos_pos, int batch, int tokens, int dim, int vocab,
                              int eos_id, void* stream) {
  return clm_embed_text_ex(ids, tok_emb, pos_emb, h, CLM_OUT_F32, eos_pos, batch, tokens, dim, vocab, eos_id, stream);
}

extern "C" int clm_embed_text_ex(const int32_t* ids, const float* tok_emb, const float* pos_emb,
                                 void* h, int h_dtype, int32_t* eos_pos, int batch, int tokens, int dim, int vocab,
                                 int eos_id, void* stream) {
  CLM_REQUIRE(ids && tok_emb && pos_emb && h && batch >= 0 && tokens > 0, "clm_embed_text: bad argument");
  CLM_REQUIRE(h_dtype == CLM_OUT_F32 || h_dtype == CLM_OUT_BF16, "clm_embed_text: h_dtype must be CLM_OUT_F32 / CLM_OUT_BF16");
  if (batch == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows = batch * tokens;
  const bool h16 = h_dtype == CLM_OUT_BF16;
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (h16 ? 6.0 : 8.0) * rows * dim, s);
  if (h16) {
    CLM_DISPATCH_DIM(dim, (embed_text_kernel<NV, true><<<blocks_for(rows), kWarpsPerBlock * 32, 0, s>>>(
                              ids, tok_emb, pos_emb, h, rows, tokens, vocab)));
  } else {
    CLM_DISPATCH_DIM(dim, (embed_text_kernel<NV, false><<<blocks_for(rows), kWarpsPerBlock * 32, 0, s>>>(
                              ids, tok_emb, pos_emb, h, rows, tokens, vocab)));
  }
  if (eos_pos)
    eos_pos_kernel<<<blocks_for(batch), kWarpsPerBlock * 32, 0, s>>>(ids, eos_pos, batch, tokens,
                                                                     eos_id);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_patch_im2col(const float* pixel_values, void* patches_bf16, int batch, int image,
                                int patch, int kpad, void* stream) {
  CLM_REQUIRE(pixel_values && patches_bf16 && batch >= 0 && patch > 0 && image % patch == 0 &&
                  kpad % 64 == 0 && kpad >= 3 * patch * patch,
              "clm_patch_im2col: bad argument");
  if (batch == 0) return CLM_OK;
  const int g = image / patch;
  const long long total = static_cast<long long>(batch) * g * g * (kpad / 8);
  const int blocks = static_cast<int>((total + 255) / 256);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 12.0 * batch * image * image + 16.0 * total,
                 static_cast<cudaStream_t>(stream));
  im2col_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pixel_values, static_cast<__nv_bfloat16*>(patches_bf16), batch, image, patch, kpad);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_vision_embed_ln(const float* patch_out, const float* class_emb,
                                   const float* pos_emb, const float* gamma, const float* beta,
                                   float* h, int batch, int np, int dim, float eps, void* stream) {
  return clm_vision_embed_ln_ex(patch_out, class_emb, pos_emb, gamma, beta, h, CLM_OUT_F32, batch, np, dim, eps,
                                stream);
}

extern "C" int clm_vision_embed_ln_ex(const float* patch_out, const float* class_emb,
                                      const float* pos_emb, const float* gamma, const float* beta,
                                      void* h, int h_dtype, int batch, int np, int dim, float eps, void* stream) {
  CLM_REQUIRE(patch_out && class_emb && pos_emb && gamma && beta && h && batch >= 0 && np > 0,
              "clm_vision_embed_ln: bad argument");
  CLM_REQUIRE(h_dtype == CLM_OUT_F32 || h_dtype == CLM_OUT_BF16,
              "clm_vision_embed_ln: h_dtype must be CLM_OUT_F32 / CLM_OUT_BF16");
  if (batch == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows = batch * (np + 1);
  const bool h16 = h_dtype == CLM_OUT_BF16;
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (h16 ? 6.0 : 8.0) * rows * dim, s);
  if (h16) {
    CLM_DISPATCH_DIM(dim, (vision_embed_ln_kernel<NV, true><<<blocks_for(rows), kWarpsPerBlock * 32, 0, s>>>(
                              patch_out, class_emb, pos_emb, gamma, beta, h, batch, np, eps)));
  } else {
    CLM_DISPATCH_DIM(dim, (vision_embed_ln_kernel<NV, false><<<blocks_for(rows), kWarpsPerBlock * 32, 0, s>>>(
                              patch_out, class_emb, pos_emb, gamma, beta, h, batch, np, eps)));
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_pool_ln(const float* h, const int32_t* row_idx_or_null, const float* gamma,
                           const float* beta, void* y_bf16, int batch, int tokens, int dim, float eps,
                           void* stream) {
  return clm_pool_ln_ex(h, CLM_OUT_F32, row_idx_or_null, gamma, beta, y_bf16, batch, tokens, dim, eps, stream);
}

extern "C" int clm_pool_ln_ex(const void* h, int h_dtype, const int32_t* row_idx_or_null, const float* gamma,
                              const float* beta, void* y_bf16, int batch, int tokens, int dim, float eps,
                              void* stream) {
  CLM_REQUIRE(h && gamma && beta && y_bf16 && batch >= 0 && tokens > 0, "clm_pool_ln: bad argument");
  CLM_REQUIRE(h_dtype == CLM_OUT_F32 || h_dtype == CLM_OUT_BF16, "clm_pool_ln: h_dtype must be CLM_OUT_F32 / CLM_OUT_BF16");
  if (batch == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool h16 = h_dtype == CLM_OUT_BF16;
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (h16 ? 4.0 : 6.0) * batch * dim, s);
  if (h16) {
    CLM_DISPATCH_DIM(dim, (pool_ln_kernel<NV, true><<<blocks_for(batch), kWarpsPerBlock * 32, 0, s>>>(
                              h, row_idx_or_null, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16),
                              batch, tokens, eps)));
  } else {
    CLM_DISPATCH_DIM(dim, (pool_ln_kernel<NV, false><<<blocks_for(batch), kWarpsPerBlock * 32, 0, s>>>(
                              h, row_idx_or_null, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16),
                              batch, tokens, eps)));
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}
