// clm_train.cu — the kernels the LoRA training step needs on top of the encoder's forward kernels
// (reference scripts/train_lora.py:83-108 loss, :170-211 step loop).
//
// The forward of a training step reuses clm_gemm_epi / clm_layernorm / clm_attention; every backward
// contraction that is GEMM-shaped (dgrad through the frozen weights, the LoRA down/up projections and their
// weight gradients) is again clm_gemm_epi on transposed operands.  What is new lives here:
//   * QuickGELU forward on a stored pre-activation and its backward,
//   * LayerNorm backward (input gradient only: gamma / beta are frozen), accumulating into the fp32
//     gradient of the residual stream and refreshing its bf16 shadow (the next dgrad GEMM's operand),
//   * batched bf16 transposes / casts (operands of the weight-gradient GEMMs, bf16 shadows of the masters),
//   * attention backward (recomputes P from q, k; CUDA cores: at the reference's training shapes -- ViT-B/32,
//     50 / 77 tokens, batch 8 -- a head is a 50 x 50 problem and the step is launch bound),
//   * the symmetric InfoNCE loss with its gradient with respect to the un-normalised features,
//   * a fused clip-by-global-norm + AdamW update over the flat LoRA parameter buffer.
// Only gradients of the LoRA factors exist: the base model is frozen (models/lora_adapter.py:46-56).
#include <math.h>

#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int kWarps = 8;
constexpr unsigned kSumsqBlocks = 1184;  // partial sums of the gradient norm (8 blocks per SM)

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------
// QuickGELU on a stored pre-activation:  g = z * sigmoid(1.702 z);  dz = dg * s * (1 + 1.702 z (1 - s))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quickgelu_fwd_kernel(const uint4* __restrict__ z, uint4* __restrict__ g, size_t n8) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 a = z[i];
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(quick_gelu(bf16_lo(w[k])), quick_gelu(bf16_hi(w[k])));
  g[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ float quick_gelu_grad(float z) {
  const float s = 1.0f / (1.0f + __expf(-1.702f * z));
  return s * (1.0f + 1.702f * z * (1.0f - s));
}

__global__ void __launch_bounds__(256)
quickgelu_bwd_kernel(const uint4* __restrict__ dg, const uint4* __restrict__ z, uint4* __restrict__ dz, size_t n8) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 a = z[i], d = dg[i];
  const uint32_t zw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    o[k] = pack_bf16x2(bf16_lo(dw[k]) * quick_gelu_grad(bf16_lo(zw[k])),
                       bf16_hi(dw[k]) * quick_gelu_grad(bf16_hi(zw[k])));
  dz[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, input gradient only.  One warp per row, the row in registers (NV float4 per lane).
//   xhat = (x - mean) rstd;  dyg = dy * gamma;  dx = rstd (dyg - mean(dyg) - xhat mean(dyg xhat))
//   dres(row) = (accumulate ? dres(row) : 0) + dx;  dres_bf16(row) = bf16(dres(row))
// row_idx != NULL: gather mode of the pooled rows -- item b reads dy row b and works on row
// b * tokens + row_idx[b] of x / dres (row_idx_is_zero: row b * tokens, the class token).
// ------------------------------------------------------------------------------------------------
template <int NV, bool kDyF32>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     float* __restrict__ dres, __nv_bfloat16* __restrict__ dres_bf16, int rows, float eps,
                     int accumulate, const int32_t* __restrict__ row_idx, int tokens, int gather) {
  const int r = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  constexpr int D = NV * 128;
  constexpr float inv_n = 1.0f / D;
  const int lane = lane_id();
  const size_t xr = gather ? (static_cast<size_t>(r) * tokens + (row_idx ? row_idx[r] : 0)) : static_cast<size_t>(r);
  const float4* x4 = reinterpret_cast<const float4*>(x + xr * D);
  float4 v[NV], g[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = x4[j * 32 + lane];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float4 d;
    if (kDyF32) {
      d = reinterpret_cast<const float4*>(static_cast<const float*>(dy) + static_cast<size_t>(r) * D)[j * 32 + lane];
    } else {
      const uint2 w = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dy) + static_cast<size_t>(r) * D)[j * 32 + lane];
      d = make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
    }
    const float4 gm = __ldg(g4 + j * 32 + lane);
    g[j] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
    v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;
    c1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    c2 += (g[j].x * v[j].x + g[j].y * v[j].y) + (g[j].z * v[j].z + g[j].w * v[j].w);
  }
  c1 = warp_sum(c1) * inv_n;
  c2 = warp_sum(c2) * inv_n;
  float4* o4 = reinterpret_cast<float4*>(dres + xr * D);
  uint2* ob = dres_bf16 ? reinterpret_cast<uint2*>(dres_bf16 + xr * D) : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float4 o;
    o.x = rstd * (g[j].x - c1 - v[j].x * c2);
    o.y = rstd * (g[j].y - c1 - v[j].y * c2);
    o.z = rstd * (g[j].z - c1 - v[j].z * c2);
    o.w = rstd * (g[j].w - c1 - v[j].w * c2);
    if (accumulate) {
      const float4 a = o4[j * 32 + lane];
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    o4[j * 32 + lane] = o;
    if (ob) ob[j * 32 + lane] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
  }
}

// ------------------------------------------------------------------------------------------------
// out[b][c][r] = bf16(scale * in[b][r][c]):  batched transpose through a padded 32 x 32 shared tile
// ------------------------------------------------------------------------------------------------
template <bool kInF32>
__global__ void __launch_bounds__(256)
transpose_kernel(const void* __restrict__ in, long long ld_in, long long bs_in, int rows, int cols,
                 __nv_bfloat16* __restrict__ out, long long ld_out, long long bs_out, float scale) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      const size_t off = static_cast<size_t>(b) * bs_in + static_cast<size_t>(r) * ld_in + c;
      v = kInF32 ? static_cast<const float*>(in)[off] : __bfloat162float(static_cast<const __nv_bfloat16*>(in)[off]);
    }
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (c < cols && r < rows)
      out[static_cast<size_t>(b) * bs_out + static_cast<size_t>(c) * ld_out + r] = __float2bfloat16_rn(scale * tile[tx][ty + 8 * k]);
  }
}

// bf16 -> bf16 transpose of a [rows, cols] matrix with 32-bit global accesses on both sides: a 64 x 64 tile is
// staged as halfwords (row stride 66: the column walk of the write-out is at most 2-way bank conflicted), and every
// thread assembles output words from two input rows.  cols, ld_in, ld_out even, rows padded by the caller's ld_out.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ld_in, int rows, int cols,
                      __nv_bfloat16* __restrict__ out, long long ld_out) {
  __shared__ uint16_t tile[64][66];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = r0 + wy + 8 * k, c = c0 + 2 * lane;
    uint32_t w = 0;
    if (r < rows && c < cols) w = *reinterpret_cast<const uint32_t*>(in + static_cast<size_t>(r) * ld_in + c);
    *reinterpret_cast<uint32_t*>(&tile[wy + 8 * k][2 * lane]) = w;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + wy + 8 * k, r = r0 + 2 * lane;
    if (c < cols && r < rows) {  // rows beyond `rows` inside the pair were staged as zeros
      const uint32_t w = static_cast<uint32_t>(tile[2 * lane][wy + 8 * k]) |
                         (static_cast<uint32_t>(tile[2 * lane + 1][wy + 8 * k]) << 16);
      *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(c) * ld_out + r) = w;
    }
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4, float scale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 a = in[i];
  out[i] = make_uint2(pack_bf16x2(scale * a.x, scale * a.y), pack_bf16x2(scale * a.z, scale * a.w));
}

// ------------------------------------------------------------------------------------------------
// Attention backward.  head_dim 64, T <= 384.  Per (batch, head), with c = 1/8:
//   S = c Q K^T (+ causal mask), P = softmax(S), dP = dO V^T, delta_i = sum_j P_ij dP_ij,
//   dS = c P (dP - delta),  dQ = dS K,  dK = dS^T Q,  dV = P^T dO.
// Kernel 1 (grid: 32-query blocks x heads): a warp owns a query row -- lane <-> key for S / dP (K, V rows in
// padded shared memory: stride 33 words, conflict free), lane <-> two output dims for dQ -- and leaves P and dS
// (bf16) TRANSPOSED in a scratch buffer [head][key][query] so that kernel 2 (grid: 64-key blocks x heads; a warp
// owns a key, lane <-> two dims) streams them with coalesced loads.  No atomics: bit-for-bit deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kAttMaxT = 384;
constexpr int kAttMaxKK = kAttMaxT / 32;
constexpr int kRowStrideW = 33;  // words per staged 64-wide bf16 row (32 data + 1 pad)

__device__ __forceinline__ void stage_rows(uint32_t* dst, const __nv_bfloat16* src, size_t ld, int T) {
  // [T][64] bf16 from global (row stride ld elements) into shared rows of kRowStrideW words
  for (int idx = threadIdx.x; idx < T * 32; idx += blockDim.x) {
    const int r = idx >> 5, w = idx & 31;
    dst[r * kRowStrideW + w] = reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(r) * ld)[w];
  }
}

__global__ void __launch_bounds__(256)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                   __nv_bfloat16* __restrict__ dqkv, __nv_bfloat16* __restrict__ scratch, int T, int Tp,
                   int heads, int causal) {
  extern __shared__ uint32_t smem_u[];
  const int D = heads * 64;
  const size_t ld = 3 * static_cast<size_t>(D);
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
  const int i0 = blockIdx.x * 32;
  uint32_t* Ks = smem_u;                            // [T][33]
  uint32_t* Vs = Ks + T * kRowStrideW;              // [T][33]
  __nv_bfloat16* Pst = reinterpret_cast<__nv_bfloat16*>(Vs + T * kRowStrideW);  // [32][Tp + 2]
  const int pst_ld = Tp + 2;
  __nv_bfloat16* dSst = Pst + 32 * pst_ld;          // [32][Tp + 2]
  float* dsrow = reinterpret_cast<float*>(dSst + 32 * pst_ld);  // [8 warps][Tp]
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld + h * 64;
  stage_rows(Ks, base + D, ld, T);
  stage_rows(Vs, base + 2 * D, ld, T);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* myds = dsrow + warp * Tp;
  const int nkk = (T + 31) / 32;
  for (int rr = warp; rr < 32; rr += kWarps) {
    const int i = i0 + rr;
    if (i >= T) {
      for (int j = lane; j < Tp; j += 32) {
        Pst[rr * pst_ld + j] = __float2bfloat16_rn(0.f);
        dSst[rr * pst_ld + j] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    // the query row and its output gradient, replicated in every lane
    const uint32_t qw = reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(i) * ld)[lane];
    const uint32_t ow = reinterpret_cast<const uint32_t*>(dout + (static_cast<size_t>(b) * T + i) * D + h * 64)[lane];
    float q[64], go[64];
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const uint32_t a = __shfl_sync(0xffffffffu, qw, d), c = __shfl_sync(0xffffffffu, ow, d);
      q[2 * d] = bf16_lo(a); q[2 * d + 1] = bf16_hi(a);
      go[2 * d] = bf16_lo(c); go[2 * d + 1] = bf16_hi(c);
    }
    float s[kAttMaxKK], dp[kAttMaxKK];
    float m = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < kAttMaxKK; ++kk) {
      s[kk] = -INFINITY;
      dp[kk] = 0.f;
      if (kk < nkk) {
        const int j = kk * 32 + lane;
        if (j < T && (!causal || j <= i)) {
          const uint32_t* kr = Ks + j * kRowStrideW;
          const uint32_t* vr = Vs + j * kRowStrideW;
          float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
          for (int w = 0; w < 32; ++w) {
            const uint32_t kw = kr[w], vw = vr[w];
            a0 = fmaf(q[2 * w], bf16_lo(kw), a0);
            a1 = fmaf(q[2 * w + 1], bf16_hi(kw), a1);
            c0 = fmaf(go[2 * w], bf16_lo(vw), c0);
            c1 = fmaf(go[2 * w + 1], bf16_hi(vw), c1);
          }
          s[kk] = (a0 + a1) * 0.125f;
          dp[kk] = c0 + c1;
          m = fmaxf(m, s[kk]);
        }
      }
    }
    m = warp_max(m);
    float l = 0.f;
#pragma unroll
    for (int kk = 0; kk < kAttMaxKK; ++kk) {
      if (kk < nkk) {
        s[kk] = (s[kk] == -INFINITY) ? 0.f : __expf(s[kk] - m);
        l += s[kk];
      }
    }
    const float inv_l = 1.0f / warp_sum(l);
    float delta = 0.f;
#pragma unroll
    for (int kk = 0; kk < kAttMaxKK; ++kk) {
      if (kk < nkk) {
        s[kk] *= inv_l;  // P
        delta = fmaf(s[kk], dp[kk], delta);
      }
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int kk = 0; kk < kAttMaxKK; ++kk) {
      if (kk < nkk) {
        const int j = kk * 32 + lane;
        const float ds = 0.125f * s[kk] * (dp[kk] - delta);
        if (j < Tp) {
          Pst[rr * pst_ld + j] = __float2bfloat16_rn(s[kk]);
          dSst[rr * pst_ld + j] = __float2bfloat16_rn(ds);
          myds[j] = ds;
        }
      }
    }
    __syncwarp();
    // dQ[i][2 lane, 2 lane + 1] = sum_j dS_j K[j][..]
    float d0 = 0.f, d1 = 0.f;
    const int jend = causal ? (i + 1) : T;
    for (int j = 0; j < jend; ++j) {
      const float ds = myds[j];
      const uint32_t kw = Ks[j * kRowStrideW + lane];
      d0 = fmaf(ds, bf16_lo(kw), d0);
      d1 = fmaf(ds, bf16_hi(kw), d1);
    }
    reinterpret_cast<uint32_t*>(dqkv + (static_cast<size_t>(b) * T + i) * ld + h * 64)[lane] = pack_bf16x2(d0, d1);
    __syncwarp();
  }
  __syncthreads();
  // transposed write-out: scratch[(bh * 2 + which) * Tp + j][i0 .. i0 + 31]
  __nv_bfloat16* pt = scratch + static_cast<size_t>(bh) * 2 * Tp * Tp;
  __nv_bfloat16* dst = pt + static_cast<size_t>(Tp) * Tp;
  for (int j = warp; j < T; j += kWarps) {
    pt[static_cast<size_t>(j) * Tp + i0 + lane] = Pst[lane * pst_ld + j];
    dst[static_cast<size_t>(j) * Tp + i0 + lane] = dSst[lane * pst_ld + j];
  }
}

__global__ void __launch_bounds__(256)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                    __nv_bfloat16* __restrict__ dqkv, const __nv_bfloat16* __restrict__ scratch, int T, int Tp,
                    int heads, int causal) {
  extern __shared__ uint32_t smem_u[];
  const int D = heads * 64;
  const size_t ld = 3 * static_cast<size_t>(D);
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
  uint32_t* Qs = smem_u;                  // [T][33]
  uint32_t* Os = Qs + T * kRowStrideW;    // [T][33] (dO)
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * ld + h * 64;
  stage_rows(Qs, base, ld, T);
  stage_rows(Os, dout + static_cast<size_t>(b) * T * D + h * 64, D, T);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const __nv_bfloat16* pt = scratch + static_cast<size_t>(bh) * 2 * Tp * Tp;
  const __nv_bfloat16* dst = pt + static_cast<size_t>(Tp) * Tp;
  const int j0 = blockIdx.x * 64;
  for (int jj = warp; jj < 64; jj += kWarps) {
    const int j = j0 + jj;
    if (j >= T) break;
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    const int ibeg = causal ? (j & ~31) : 0;  // P_ij = 0 for i < j under the causal mask
    for (int ib = ibeg; ib < T; ib += 32) {
      const float pv = __bfloat162float(pt[static_cast<size_t>(j) * Tp + ib + lane]);
      const float dv = __bfloat162float(dst[static_cast<size_t>(j) * Tp + ib + lane]);
      const int n = min(32, T - ib);
      for (int ii = 0; ii < n; ++ii) {
        const float p = __shfl_sync(0xffffffffu, pv, ii), ds = __shfl_sync(0xffffffffu, dv, ii);
        const uint32_t qw = Qs[(ib + ii) * kRowStrideW + lane], ow = Os[(ib + ii) * kRowStrideW + lane];
        k0 = fmaf(ds, bf16_lo(qw), k0);
        k1 = fmaf(ds, bf16_hi(qw), k1);
        v0 = fmaf(p, bf16_lo(ow), v0);
        v1 = fmaf(p, bf16_hi(ow), v1);
      }
    }
    uint32_t* row = reinterpret_cast<uint32_t*>(dqkv + (static_cast<size_t>(b) * T + j) * ld + h * 64);
    row[D / 2 + lane] = pack_bf16x2(k0, k1);
    row[D + lane] = pack_bf16x2(v0, v1);
  }
}

// ------------------------------------------------------------------------------------------------
// C[m][n] = alpha * sum_k A[m sam + k sak] * B[k sbk + n sbn]   (fp32, CUDA cores; the N x N logits of the loss
// and their two gradient products -- /temperature amplifies the logits 14x, so they stay in fp32)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
             long long sbn, float* __restrict__ C, int ldc, int M, int N, int K, float alpha) {
  __shared__ float As[16][65], Bs[16][65];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      const int kk = idx & 15, mm = idx >> 4;
      const int m = m0 + mm, n = n0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
      Bs[kk][mm] = (n < N && k < K) ? B[k * sbk + n * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; bb[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[static_cast<size_t>(m) * ldc + n] = alpha * acc[i][j];
    }
}

// n = x / ||x|| per row (no epsilon, as the reference's loss: train_lora.py:95-96) and 1 / ||x||
__global__ void __launch_bounds__(kWarps * 32)
normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ n, float* __restrict__ inv_norm, int rows, int dim) {
  const int r = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + static_cast<size_t>(r) * dim;
  float s = 0.f;
  for (int c = lane_id(); c < dim; c += 32) s = fmaf(xr[c], xr[c], s);
  const float inv = 1.0f / sqrtf(warp_sum(s));
  for (int c = lane_id(); c < dim; c += 32) n[static_cast<size_t>(r) * dim + c] = xr[c] * inv;
  if (lane_id() == 0) inv_norm[r] = inv;
}

// log-sum-exp of every row (which = 0) and every column (which = 1) of the N x N logits
__global__ void __launch_bounds__(kWarps * 32)
lse_kernel(const float* __restrict__ L, float* __restrict__ lse_row, float* __restrict__ lse_col, int N) {
  const int w = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (w >= 2 * N) return;
  const int which = w >= N, r = which ? w - N : w;
  const long long s0 = which ? 1 : N, s1 = which ? N : 1;  // element (r, c) of the walk at r * s0 + c * s1
  float m = -INFINITY;
  for (int c = lane_id(); c < N; c += 32) m = fmaxf(m, L[r * s0 + c * s1]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane_id(); c < N; c += 32) s += expf(L[r * s0 + c * s1] - m);
  s = warp_sum(s);
  if (lane_id() == 0) (which ? lse_col : lse_row)[r] = m + logf(s);
}

// loss = mean_i (lse_row_i - L_ii) / 2 + mean_j (lse_col_j - L_jj) / 2   (train_lora.py:101-106); one block, fixed
// summation order (the value is reproducible bit for bit)
__global__ void __launch_bounds__(256)
loss_value_kernel(const float* __restrict__ L, const float* __restrict__ lse_row, const float* __restrict__ lse_col,
                  int N, float loss_scale, float* __restrict__ loss_out) {
  __shared__ float red[8];
  float part = 0.f;
  for (int i = threadIdx.x; i < N; i += 256) {
    const float l = L[static_cast<size_t>(i) * N + i];
    part += (lse_row[i] - l) + (lse_col[i] - l);
  }
  part = warp_sum(part);
  if (lane_id() == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    *loss_out = v * loss_scale / (2.0f * N);
  }
}

// G = dloss/dL * loss_scale / temperature, in place over L:  dloss/dL_ij = (softmax_row + softmax_col - 2 [i = j]) / (2N)
__global__ void __launch_bounds__(256)
loss_grad_kernel(float* __restrict__ L, const float* __restrict__ lse_row, const float* __restrict__ lse_col, int N,
                 float inv_temp, float loss_scale) {
  const size_t total = static_cast<size_t>(N) * N;
  const float gs = loss_scale * inv_temp / (2.0f * N);
  for (size_t idx = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; idx < total; idx += static_cast<size_t>(gridDim.x) * 256) {
    const int i = static_cast<int>(idx / N), j = static_cast<int>(idx - static_cast<size_t>(i) * N);
    const float l = L[idx];
    float g = expf(l - lse_row[i]) + expf(l - lse_col[j]);
    if (i == j) g -= 2.0f;
    L[idx] = g * gs;
  }
}

// dx = (dn - n (n . dn)) / ||x||  per row: the backward of x / ||x||
__global__ void __launch_bounds__(kWarps * 32)
normalize_bwd_kernel(const float* __restrict__ n, const float* __restrict__ inv_norm, const float* __restrict__ dn,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16, int rows, int dim) {
  const int r = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const size_t o = static_cast<size_t>(r) * dim;
  float s = 0.f;
  for (int c = lane_id(); c < dim; c += 32) s = fmaf(n[o + c], dn[o + c], s);
  s = warp_sum(s);
  const float inv = inv_norm[r];
  for (int c = lane_id(); c < dim; c += 32) {
    const float v = (dn[o + c] - n[o + c] * s) * inv;
    dx[o + c] = v;
    if (dx_bf16) dx_bf16[o + c] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Optimizer: sum of squares of the masked gradient, then clip-by-global-norm + AdamW in one pass
// (torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW semantics; train_lora.py:141,190-193).
// mult[i] is 0 for structural zeros of the fused LoRA layouts (padding columns, off-diagonal blocks of B_cat),
// otherwise the factor that turns the GEMM's raw product into the gradient of the master entry (1 for A, the
// LoRA scaling for B).  hyper (device): {lr, 1 / (1 - beta1^t), 1 / sqrt(1 - beta2^t), unused}.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, const float* __restrict__ mult, size_t n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256) {
    const float v = g[i] * mult[i];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane_id() == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    partial[blockIdx.x] = v;
  }
}

// out[0] = sum of the per-block partials in a fixed order (no atomics: the clip factor, and with it every
// parameter, is reproducible bit for bit)
__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  s = warp_sum(s);
  if (lane_id() == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    *out = v;
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ mult, float* __restrict__ m,
             float* __restrict__ v, size_t n, const float* __restrict__ hyper, const float* __restrict__ sumsq,
             float max_norm, float beta1, float beta2, float eps, float wd) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float mu = mult[i];
  if (mu == 0.f) return;  // structural zero: stays exactly zero
  const float lr = hyper[0], bc1 = hyper[1], bc2s = hyper[2];
  float clip = 1.0f;
  if (max_norm > 0.f) {
    const float norm = sqrtf(*sumsq);
    clip = fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const float gi = g[i] * mu * clip;
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  pi -= lr * bc1 * mi / (sqrtf(vi) * bc2s + eps);
  p[i] = pi;
}

// ------------------------------------------------------------------------------------------------
// LoRA weight gradients for FEW token rows (the reference's batch of 8: 400 / 616 rows), where the tensor-core
// route -- four transposes and two deep-K GEMMs -- is six launches of a few microseconds each:
//   gB[f][c] += sum_r dy[r][f] t[r][c]   (f < n_out)        gA[f][c] += sum_r x[r][f] u[r][c]   (f < n_in)
// One launch, CUDA cores, fp32 accumulation: a CTA (128 threads, 4 x 4 outputs each) owns 32 features x 64 LoRA columns of one of
// the two products and walks its token rows in chunks of 64 staged in shared memory, the next chunk's loads in flight.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
lora_wgrad_small_kernel(const __nv_bfloat16* __restrict__ dy, int ld_dy, int n_out, const __nv_bfloat16* __restrict__ t,
                        int ld_t, const __nv_bfloat16* __restrict__ x, int ld_x, int n_in,
                        const __nv_bfloat16* __restrict__ u, int ld_u, int cols, int rows, float* __restrict__ gB,
                        float* __restrict__ gA) {
  constexpr int kChunk = 64;  // token rows per staged chunk
  __shared__ __align__(16) float fs[kChunk][36], cs[kChunk][68];  // 16-byte aligned rows: 128-bit reads in the product loop
  const int blocks_b = (n_out + 31) / 32;
  const int cblocks = cols / 64;
  const int fb = blockIdx.x / cblocks, cb = blockIdx.x - fb * cblocks;
  const bool is_b = fb < blocks_b;
  const __nv_bfloat16* F = is_b ? dy : x;
  const __nv_bfloat16* Cm = is_b ? t : u;
  const int ldf = is_b ? ld_dy : ld_x, ldc = is_b ? ld_t : ld_u;
  const int nf = is_b ? n_out : n_in;
  const int f0 = (is_b ? fb : fb - blocks_b) * 32, c0 = cb * 64;
  float* out = is_b ? gB : gA;
  // gridDim.y CTAs share the token rows of a tile (contiguous slices, multiples of the chunk) and add their partial
  // sums with atomics; gridDim.y == 1: one CTA per tile, plain read-modify-write, bit-for-bit reproducible
  const int per = ((rows + gridDim.y - 1) / gridDim.y + kChunk - 1) / kChunk * kChunk;
  const int r_lo = blockIdx.y * per, r_hi = min(rows, r_lo + per);
  if (r_lo >= r_hi) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 4 columns x 4 features per thread: 16 x 8 threads
  const int w = threadIdx.x & 31, rbase = threadIdx.x >> 5;  // staging: bf16 pair w of rows rbase + 4 k
  const bool f_ok = w < 16 && f0 + 2 * w < nf;
  float acc[4][4] = {};
  uint32_t fw[16], cw[16];
  // every load of a chunk is issued before anything waits on one, and the next chunk's loads are in flight while
  // this one is multiplied: a CTA walks its rows alone, so the L2 latency would otherwise be paid once per chunk
  auto load_chunk = [&](int r0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int r = r0 + rbase + 4 * k;
      fw[k] = cw[k] = 0u;
      if (r < r_hi) {
        if (f_ok) fw[k] = __ldg(reinterpret_cast<const uint32_t*>(F + static_cast<size_t>(r) * ldf + f0 + 2 * w));
        cw[k] = __ldg(reinterpret_cast<const uint32_t*>(Cm + static_cast<size_t>(r) * ldc + c0 + 2 * w));
      }
    }
  };
  load_chunk(r_lo);
  for (int r0 = r_lo; r0 < r_hi; r0 += kChunk) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int rr = rbase + 4 * k;
      if (w < 16) { fs[rr][2 * w] = bf16_lo(fw[k]); fs[rr][2 * w + 1] = bf16_hi(fw[k]); }
      cs[rr][2 * w] = bf16_lo(cw[k]); cs[rr][2 * w + 1] = bf16_hi(cw[k]);
    }
    __syncthreads();
    if (r0 + kChunk < r_hi) load_chunk(r0 + kChunk);
#pragma unroll 8
    for (int rr = 0; rr < kChunk; ++rr) {
      const float4 av = *reinterpret_cast<const float4*>(&fs[rr][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&cs[rr][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + ty * 4 + i;
    if (f < nf) {
      float* o = out + static_cast<size_t>(f) * cols + c0 + tx * 4;
      if (gridDim.y == 1) {
        float4 v = *reinterpret_cast<float4*>(o);
        v.x += acc[i][0]; v.y += acc[i][1]; v.z += acc[i][2]; v.w += acc[i][3];
        *reinterpret_cast<float4*>(o) = v;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(o + j, acc[i][j]);
      }
    }
  }
}

#define CLM_TRAIN_DISPATCH_DIM(dim, CALL)                                     \
  switch (dim) {                                                              \
    case 128: { constexpr int NV = 1; CALL; break; }                          \
    case 256: { constexpr int NV = 2; CALL; break; }                          \
    case 512: { constexpr int NV = 4; CALL; break; }                          \
    case 768: { constexpr int NV = 6; CALL; break; }                          \
    case 1024: { constexpr int NV = 8; CALL; break; }                         \
    default:                                                                  \
      clm_set_error("unsupported width %d (supported: 128,256,512,768,1024)", dim); \
      return CLM_ERR_UNSUPPORTED;                                             \
  }

inline int attn_tp(int tokens) { return (tokens + 31) / 32 * 32; }

}  // namespace

// clm_attention_bwd.cu: the tcgen05 kernel for sequences of at most 128 tokens
int clm_attention_bwd_tc_supported(int tokens);
int clm_attention_bwd_tc_launch(const void* qkv, const void* dout, void* dqkv, int batch, int tokens, int heads,
                                int causal, cudaStream_t stream);

extern "C" int clm_quickgelu_fwd(const void* z_bf16, void* g_bf16, long long n, void* stream) {
  CLM_REQUIRE(z_bf16 && g_bf16 && n >= 0 && n % 8 == 0, "clm_quickgelu_fwd: bad argument (n must be a multiple of 8)");
  if (n == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 4.0 * n, s);
  const size_t n8 = static_cast<size_t>(n) / 8;
  quickgelu_fwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(
      static_cast<const uint4*>(z_bf16), static_cast<uint4*>(g_bf16), n8);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_quickgelu_bwd(const void* dg_bf16, const void* z_bf16, void* dz_bf16, long long n, void* stream) {
  CLM_REQUIRE(dg_bf16 && z_bf16 && dz_bf16 && n >= 0 && n % 8 == 0, "clm_quickgelu_bwd: bad argument");
  if (n == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 6.0 * n, s);
  const size_t n8 = static_cast<size_t>(n) / 8;
  quickgelu_bwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(
      static_cast<const uint4*>(dg_bf16), static_cast<const uint4*>(z_bf16), static_cast<uint4*>(dz_bf16), n8);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_layernorm_bwd(const void* dy, int dy_is_f32, const float* x, const float* gamma, float* dres,
                                 void* dres_bf16_or_null, int rows, int dim, float eps, int accumulate,
                                 const int32_t* row_idx_or_null, int tokens, int gather, void* stream) {
  CLM_REQUIRE(dy && x && gamma && dres && rows >= 0 && (!gather || tokens > 0), "clm_layernorm_bwd: bad argument");
  if (rows == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (dy_is_f32 ? 4.0 : 2.0) * rows * dim + (accumulate ? 14.0 : 10.0) * rows * dim, s);
  const int blocks = (rows + kWarps - 1) / kWarps;
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(dres_bf16_or_null);
  if (dy_is_f32) {
    CLM_TRAIN_DISPATCH_DIM(dim, (layernorm_bwd_kernel<NV, true><<<blocks, kWarps * 32, 0, s>>>(
                                    dy, x, gamma, dres, sh, rows, eps, accumulate, row_idx_or_null, tokens, gather)));
  } else {
    CLM_TRAIN_DISPATCH_DIM(dim, (layernorm_bwd_kernel<NV, false><<<blocks, kWarps * 32, 0, s>>>(
                                    dy, x, gamma, dres, sh, rows, eps, accumulate, row_idx_or_null, tokens, gather)));
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_transpose_to_bf16(const void* in, int in_is_f32, long long ld_in, long long batch_stride_in,
                                     int rows, int cols, void* out_bf16, long long ld_out,
                                     long long batch_stride_out, int batch, float scale, void* stream) {
  CLM_REQUIRE(in && out_bf16 && rows >= 0 && cols >= 0 && batch >= 0 && ld_in >= cols && ld_out >= rows,
              "clm_transpose_to_bf16: bad argument");
  if (rows == 0 || cols == 0 || batch == 0) return CLM_OK;
  CLM_REQUIRE(batch <= 65535 && (rows + 31) / 32 <= 65535, "clm_transpose_to_bf16: too many tiles");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, (in_is_f32 ? 6.0 : 4.0) * rows * cols * batch, s);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out_bf16);
  if (!in_is_f32 && batch == 1 && scale == 1.0f && cols % 2 == 0 && ld_in % 2 == 0 && ld_out % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 3) == 0) {
    transpose_bf16_kernel<<<dim3((cols + 63) / 64, (rows + 63) / 64), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(in), ld_in, rows, cols, ob, ld_out);
    CLM_CUDA_CHECK(cudaGetLastError());
    return CLM_OK;
  }
  const dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  if (in_is_f32)
    transpose_kernel<true><<<grid, 256, 0, s>>>(in, ld_in, batch_stride_in, rows, cols,
                                                static_cast<__nv_bfloat16*>(out_bf16), ld_out, batch_stride_out, scale);
  else
    transpose_kernel<false><<<grid, 256, 0, s>>>(in, ld_in, batch_stride_in, rows, cols,
                                                 static_cast<__nv_bfloat16*>(out_bf16), ld_out, batch_stride_out, scale);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_cast_to_bf16(const float* in, void* out_bf16, long long n, float scale, void* stream) {
  CLM_REQUIRE(in && out_bf16 && n >= 0 && n % 4 == 0, "clm_cast_to_bf16: bad argument (n must be a multiple of 4)");
  if (n == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 6.0 * n, s);
  const size_t n4 = static_cast<size_t>(n) / 4;
  cast_bf16_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const float4*>(in), static_cast<uint2*>(out_bf16), n4, scale);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_lora_wgrad_small(const void* dy_bf16, int ld_dy, int n_out, const void* t_bf16, int ld_t,
                                    const void* x_bf16, int ld_x, int n_in, const void* u_bf16, int ld_u, int cols,
                                    int rows, float* grad_b, float* grad_a_t, int deterministic, void* stream) {
  CLM_REQUIRE(dy_bf16 && t_bf16 && x_bf16 && u_bf16 && grad_b && grad_a_t, "clm_lora_wgrad_small: null argument");
  CLM_REQUIRE(rows >= 0 && n_out > 0 && n_in > 0 && cols > 0 && cols % 64 == 0 && n_out % 2 == 0 && n_in % 2 == 0 &&
                  ld_dy % 2 == 0 && ld_t % 2 == 0 && ld_x % 2 == 0 && ld_u % 2 == 0 && ld_dy >= n_out && ld_x >= n_in &&
                  ld_t >= cols && ld_u >= cols,
              "clm_lora_wgrad_small: bad shape (cols a multiple of 64, feature counts and leading dims even)");
  if (rows == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = ((n_out + 31) / 32 + (n_in + 31) / 32) * (cols / 64);
  // few tiles: up to four CTAs share a tile's token rows (atomic adds, arrival order) unless reproducibility is asked for
  int splits = 1;
  if (!deterministic) {
    splits = clm_num_sms() / blocks;
    if (splits > 4) splits = 4;
    if (splits > (rows + 127) / 128) splits = (rows + 127) / 128;
    if (splits < 1) splits = 1;
  }
  ProfScope prof(CLM_K_GEMM, 2.0 * rows * cols * (static_cast<double>(n_out) + n_in),
                 2.0 * rows * (static_cast<double>(n_out) + n_in + 2 * cols) + 8.0 * cols * (static_cast<double>(n_out) + n_in), s);
  lora_wgrad_small_kernel<<<dim3(blocks, splits), 128, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy_bf16), ld_dy, n_out, static_cast<const __nv_bfloat16*>(t_bf16), ld_t,
      static_cast<const __nv_bfloat16*>(x_bf16), ld_x, n_in, static_cast<const __nv_bfloat16*>(u_bf16), ld_u, cols, rows,
      grad_b, grad_a_t);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" size_t clm_attention_bwd_scratch_bytes(int batch, int tokens, int heads) {
  if (batch <= 0 || tokens <= 0 || heads <= 0) return 0;
  const size_t tp = attn_tp(tokens);
  return static_cast<size_t>(batch) * heads * 2 * tp * tp * sizeof(__nv_bfloat16);
}

extern "C" int clm_attention_bwd(const void* qkv_bf16, const void* dout_bf16, void* dqkv_bf16, void* scratch,
                                 size_t scratch_bytes, int batch, int tokens, int heads, int causal, void* stream) {
  CLM_REQUIRE(qkv_bf16 && dout_bf16 && dqkv_bf16 && scratch && batch >= 0 && heads > 0,
              "clm_attention_bwd: bad argument");
  CLM_REQUIRE(tokens >= 1 && tokens <= kAttMaxT, "clm_attention_bwd: tokens=%d must be in [1,%d]", tokens, kAttMaxT);
  if (batch == 0) return CLM_OK;
  CLM_REQUIRE(scratch_bytes >= clm_attention_bwd_scratch_bytes(batch, tokens, heads),
              "clm_attention_bwd: scratch of %zu bytes too small (need %zu)", scratch_bytes,
              clm_attention_bwd_scratch_bytes(batch, tokens, heads));
  CLM_REQUIRE(static_cast<long long>(batch) * heads <= 65535, "clm_attention_bwd: batch * heads > 65535");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (clm_attention_bwd_tc_supported(tokens))
    return clm_attention_bwd_tc_launch(qkv_bf16, dout_bf16, dqkv_bf16, batch, tokens, heads, causal, s);
  const int T = tokens, Tp = attn_tp(tokens);
  const size_t smem1 = static_cast<size_t>(2) * T * kRowStrideW * 4 + static_cast<size_t>(2) * 32 * (Tp + 2) * 2 +
                       static_cast<size_t>(kWarps) * Tp * 4;
  const size_t smem2 = static_cast<size_t>(2) * T * kRowStrideW * 4;
  static bool attr_done = false;
  if (!attr_done) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const double flops = 10.0 * batch * heads * static_cast<double>(T) * T * 64;  // five T x T x 64 products
  const double bytes = 2.0 * batch * T * heads * 64 * (3 + 1 + 3);
  const __nv_bfloat16* qkv = static_cast<const __nv_bfloat16*>(qkv_bf16);
  const __nv_bfloat16* dout = static_cast<const __nv_bfloat16*>(dout_bf16);
  __nv_bfloat16* dqkv = static_cast<__nv_bfloat16*>(dqkv_bf16);
  __nv_bfloat16* scr = static_cast<__nv_bfloat16*>(scratch);
  {
    ProfScope prof(CLM_K_ATTENTION, 0.6 * flops, bytes, s);
    attn_bwd_dq_kernel<<<dim3((T + 31) / 32, batch * heads), 256, smem1, s>>>(qkv, dout, dqkv, scr, T, Tp, heads, causal);
    CLM_CUDA_CHECK(cudaGetLastError());
  }
  {
    ProfScope prof(CLM_K_ATTENTION, 0.4 * flops, bytes, s);
    attn_bwd_dkv_kernel<<<dim3((T + 63) / 64, batch * heads), 256, smem2, s>>>(qkv, dout, dqkv, scr, T, Tp, heads, causal);
    CLM_CUDA_CHECK(cudaGetLastError());
  }
  return CLM_OK;
}

extern "C" size_t clm_clip_loss_workspace_bytes(int n, int dim) {
  if (n <= 0 || dim <= 0) return 0;
  const size_t nn = static_cast<size_t>(n);
  // n_i, n_t, dn_i, dn_t [n, dim]; logits [n, n]; inv norms (2n), lse (2n)
  return (4 * nn * dim + nn * nn + 4 * nn + 64) * sizeof(float);
}

extern "C" int clm_clip_loss(const float* feat_i, const float* feat_t, int n, int dim, float temperature,
                             float loss_scale, float* loss_out, float* dfeat_i, float* dfeat_t, void* dfeat_i_bf16,
                             void* dfeat_t_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  CLM_REQUIRE(feat_i && feat_t && loss_out && workspace && n > 0 && dim > 0 && temperature > 0.f,
              "clm_clip_loss: bad argument");
  CLM_REQUIRE(workspace_bytes >= clm_clip_loss_workspace_bytes(n, dim), "clm_clip_loss: workspace too small");
  CLM_REQUIRE((dfeat_i == nullptr) == (dfeat_t == nullptr), "clm_clip_loss: give both gradients or neither");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t nd = static_cast<size_t>(n) * dim;
  float* ni = static_cast<float*>(workspace);
  float* nt = ni + nd;
  float* dni = nt + nd;
  float* dnt = dni + nd;
  float* L = dnt + nd;
  float* inv_i = L + static_cast<size_t>(n) * n;
  float* inv_t = inv_i + n;
  float* lse_r = inv_t + n;
  float* lse_c = lse_r + n;
  const int rb = (n + kWarps - 1) / kWarps;
  const dim3 gnn((n + 63) / 64, (n + 63) / 64), gnd((dim + 63) / 64, (n + 63) / 64);
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * nd, s);
    normalize_rows_kernel<<<rb, kWarps * 32, 0, s>>>(feat_i, ni, inv_i, n, dim);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * nd, s);
    normalize_rows_kernel<<<rb, kWarps * 32, 0, s>>>(feat_t, nt, inv_t, n, dim);
  }
  {
    ProfScope prof(CLM_K_GEMM, 2.0 * n * n * dim, 8.0 * nd + 4.0 * n * n, s);
    sgemm_kernel<<<gnn, 256, 0, s>>>(ni, dim, 1, nt, 1, dim, L, n, n, n, dim, 1.0f / temperature);  // L = n_i n_t^T / T
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * n * n, s);
    lse_kernel<<<(2 * n + kWarps - 1) / kWarps, kWarps * 32, 0, s>>>(L, lse_r, lse_c, n);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 12.0 * n, s);
    loss_value_kernel<<<1, 256, 0, s>>>(L, lse_r, lse_c, n, loss_scale, loss_out);
  }
  if (dfeat_i) {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * n * n, s);
    const size_t total = static_cast<size_t>(n) * n;
    const unsigned blocks = static_cast<unsigned>(total / 256 + 1 < 1184 ? total / 256 + 1 : 1184);
    loss_grad_kernel<<<blocks, 256, 0, s>>>(L, lse_r, lse_c, n, 1.0f / temperature, loss_scale);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  if (!dfeat_i) return CLM_OK;
  {
    ProfScope prof(CLM_K_GEMM, 2.0 * n * n * dim, 8.0 * nd + 4.0 * n * n, s);
    sgemm_kernel<<<gnd, 256, 0, s>>>(L, n, 1, nt, dim, 1, dni, dim, n, dim, n, 1.0f);  // dn_i = G n_t
  }
  {
    ProfScope prof(CLM_K_GEMM, 2.0 * n * n * dim, 8.0 * nd + 4.0 * n * n, s);
    sgemm_kernel<<<gnd, 256, 0, s>>>(L, 1, n, ni, dim, 1, dnt, dim, n, dim, n, 1.0f);  // dn_t = G^T n_i
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 14.0 * nd, s);
    normalize_bwd_kernel<<<rb, kWarps * 32, 0, s>>>(ni, inv_i, dni, dfeat_i, static_cast<__nv_bfloat16*>(dfeat_i_bf16), n, dim);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 14.0 * nd, s);
    normalize_bwd_kernel<<<rb, kWarps * 32, 0, s>>>(nt, inv_t, dnt, dfeat_t, static_cast<__nv_bfloat16*>(dfeat_t_bf16), n, dim);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_adamw_step(float* params, const float* grads, const float* grad_mult, float* exp_avg,
                              float* exp_avg_sq, long long n, const float* hyper_dev, float* sumsq_scratch,
                              float max_grad_norm, float beta1, float beta2, float eps, float weight_decay,
                              void* stream) {
  CLM_REQUIRE(params && grads && grad_mult && exp_avg && exp_avg_sq && hyper_dev && sumsq_scratch && n >= 0,
              "clm_adamw_step: bad argument");
  if (n == 0) return CLM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t nn = static_cast<size_t>(n);
  const unsigned blocks = static_cast<unsigned>((nn + 255) / 256 < kSumsqBlocks ? (nn + 255) / 256 : kSumsqBlocks);
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 8.0 * nn, s);
    grad_sumsq_kernel<<<blocks, 256, 0, s>>>(grads, grad_mult, nn, sumsq_scratch + 1);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 4.0 * blocks, s);
    sumsq_final_kernel<<<1, 256, 0, s>>>(sumsq_scratch + 1, static_cast<int>(blocks), sumsq_scratch);
  }
  {
    ProfScope prof(CLM_K_ELEMENTWISE, 0.0, 32.0 * nn, s);
    adamw_kernel<<<static_cast<unsigned>((nn + 255) / 256), 256, 0, s>>>(params, grads, grad_mult, exp_avg, exp_avg_sq, nn,
                                                                         hyper_dev, sumsq_scratch, max_grad_norm, beta1,
                                                                         beta2, eps, weight_decay);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}
