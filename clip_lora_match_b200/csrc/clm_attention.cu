// clm_attention.cu — fused attention for CLIP ViT towers on tcgen05 / TMEM.
//
// CLIP sequences are short (T = 50 / 77 / 197 / 257), so a whole score row block fits in
// tensor memory: one CTA handles one (batch, head, 128-query tile) and
//   1. TMA-loads Q (128x64), K (Tk x 64) and V (Tk x 64) as 128-byte-swizzled boxes straight
//      out of the fused QKV activation [B*T, 3D];
//   2. S = Q K^T with tcgen05.mma (K-major A and B), S stays in TMEM (T columns of fp32);
//   3. each of the 128 threads owns one query row: reads its S row from TMEM twice (max, then
//      exp/sum — fp32 softmax, modeling_clip.py:274), writes P as bf16 into shared memory in
//      the K-major SWIZZLE_128B layout UMMA expects;
//   4. O = P V with tcgen05.mma, V consumed as an MN-major B operand (no transpose pass);
//   5. O rows are scaled by 1/rowsum, converted to bf16 and stored (128 contiguous bytes/row).
// No score or probability ever reaches HBM.  Keys beyond T (the next sequence's rows, or TMA
// zero fill) are masked in step 3; query rows beyond T are computed but never stored.
#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int kThreads = 128;
constexpr int kHeadDim = 64;
constexpr int kBoxRows = 64;
constexpr int kBoxBytes = kBoxRows * kHeadDim * 2;  // 8 KiB

// MN-major SWIZZLE_128B descriptor (V: rows = keys (K dim), 64 contiguous head-dim elements (N)).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO: single 64-wide MN atom, unused
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: 8 key rows (8 x 128 B)
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool kCausal>
__global__ void __launch_bounds__(kThreads)
attention_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_tail,
                 __nv_bfloat16* __restrict__ out, int T, int H, int Tp, int Tk, int mtiles,
                 uint32_t tmem_cols, int region_a_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                       // 128 x 64 bf16
  uint8_t* sK = smem + 2 * kBoxBytes;       // Tp x 64 bf16
  uint8_t* sP = smem;                       // aliases Q,K once S is complete: 128 x Tk bf16
  uint8_t* sV = smem + region_a_bytes;      // Tp x 64 bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Tp * 128);
  uint64_t* bar_load = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int mt = blockIdx.x % mtiles;
  const int bh = blockIdx.x / mtiles;
  const int h = bh % H;
  const int b = bh / H;
  const int D = H * kHeadDim;
  const int row_base = b * T;
  const int n64 = Tp / kBoxRows;          // full 64-row boxes of K / V
  const int n16 = (Tp % kBoxRows) / 16;   // 16-row boxes covering the ragged tail

  if (tid == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_tail);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(2 * kBoxBytes + 2 * Tp * 128));
    for (int i = 0; i < 2; ++i)
      tma_load_2d(sQ + i * kBoxBytes, &map_qkv, bar_load, h * kHeadDim,
                  row_base + mt * 128 + i * kBoxRows);
    for (int i = 0; i < n64; ++i)
      tma_load_2d(sK + i * kBoxBytes, &map_qkv, bar_load, D + h * kHeadDim, row_base + i * kBoxRows);
    for (int i = 0; i < n16; ++i)
      tma_load_2d(sK + n64 * kBoxBytes + i * 2048, &map_tail, bar_load, D + h * kHeadDim,
                  row_base + n64 * kBoxRows + i * 16);
    for (int i = 0; i < n64; ++i)
      tma_load_2d(sV + i * kBoxBytes, &map_qkv, bar_load, 2 * D + h * kHeadDim, row_base + i * kBoxRows);
    for (int i = 0; i < n16; ++i)
      tma_load_2d(sV + n64 * kBoxBytes + i * 2048, &map_tail, bar_load, 2 * D + h * kHeadDim,
                  row_base + n64 * kBoxRows + i * 16);
  }
  mbar_wait(bar_load, 0);

  // ---- S = Q K^T ------------------------------------------------------------------------
  if (tid == 0) {
    tc_fence_after();
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t k_addr = smem_u32(sK);
    for (int n0 = 0; n0 < Tp; n0 += 256) {
      const int nn = (Tp - n0) < 256 ? (Tp - n0) : 256;
      const uint32_t idesc = umma_idesc_bf16(128, nn, 0, 0);
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k) {
        const uint64_t da = umma_desc_sw128(q_addr + k * 32, 1024);
        const uint64_t db = umma_desc_sw128(k_addr + n0 * 128 + k * 32, 1024);
        umma_bf16_ss(tmem + n0, da, db, idesc, k != 0 ? 1u : 0u);
      }
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // ---- softmax: thread = query row ------------------------------------------------------
  const int r = tid;            // row inside the tile == TMEM lane
  const int qi = mt * 128 + r;  // query position in the sequence
  int valid = T;
  if (kCausal) valid = (qi + 1 < T) ? qi + 1 : T;
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int nchunks = (Tp + 31) / 32;
  constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

  // pass 1: row maximum.  Only chunks that cross `valid` need per-element masking.
  float mx = -INFINITY;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(trow + c * 32, v);
    tmem_ld_wait();
    if (c * 32 + 32 <= valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
  }
  // pass 2: p = 2^(s*c - mx*c) (one FFMA + one MUFU), row sum, P -> shared memory as bf16
  const float neg_mx = -mx * kScaleLog2e;
  float sum = 0.f;
  uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(trow + c * 32, v);
    tmem_ld_wait();
    float p[32];
    if (c * 32 + 32 <= valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        p[i] = fast_exp2(fmaf(__uint_as_float(v[i]), kScaleLog2e, neg_mx));
        sum += p[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = fast_exp2(fmaf(__uint_as_float(v[i]), kScaleLog2e, neg_mx));
        p[i] = (c * 32 + i < valid) ? e : 0.f;
        sum += p[i];
      }
    }
    uint8_t* pblk = prow + (c >> 1) * 16384;  // 64-column k-block
    const int j0 = (c & 1) * 4;               // first 16-byte chunk inside the 128-byte row
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      uint4 o;
      o.x = pack_bf16x2(p[8 * jj + 0], p[8 * jj + 1]);
      o.y = pack_bf16x2(p[8 * jj + 2], p[8 * jj + 3]);
      o.z = pack_bf16x2(p[8 * jj + 4], p[8 * jj + 5]);
      o.w = pack_bf16x2(p[8 * jj + 6], p[8 * jj + 7]);
      *reinterpret_cast<uint4*>(pblk + (((j0 + jj) ^ (r & 7)) << 4)) = o;
    }
  }
  fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the UMMA operand reads
  tc_fence_before();
  __syncthreads();

  // ---- O = P V ---------------------------------------------------------------------------
  if (tid == 0) {
    tc_fence_after();
    const uint32_t p_addr = smem_u32(sP);
    const uint32_t v_addr = smem_u32(sV);
    constexpr uint32_t idesc = umma_idesc_bf16(128, kHeadDim, 0, 1);
    const int ksteps = Tp / 16;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = umma_desc_sw128(p_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 1024);
      const uint64_t db = umma_desc_sw128_mn(v_addr + ks * 2048);
      umma_bf16_ss(tmem, da, db, idesc, ks != 0 ? 1u : 0u);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 1);
  tc_fence_after();

  // ---- epilogue ---------------------------------------------------------------------------
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(trow + c * 32, v);
    tmem_ld_wait();
    if (qi < T) {
      uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(row_base + qi) * D +
                                           h * kHeadDim + c * 32);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[8 * jj + 0]) * inv, __uint_as_float(v[8 * jj + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[8 * jj + 2]) * inv, __uint_as_float(v[8 * jj + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[8 * jj + 4]) * inv, __uint_as_float(v[8 * jj + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[8 * jj + 6]) * inv, __uint_as_float(v[8 * jj + 7]) * inv);
        o4[jj] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace

int clm_attention_launch(const void* qkv, void* out, int batch, int tokens, int heads, int causal,
                         cudaStream_t stream) {
  CLM_REQUIRE(qkv && out && batch >= 0 && tokens > 0 && heads > 0, "clm_attention: bad argument");
  CLM_REQUIRE(tokens <= 512, "clm_attention: tokens=%d > 512 unsupported (CLIP uses <= 257)", tokens);
  if (batch == 0) return CLM_OK;
  const int T = tokens;
  const int D = heads * kHeadDim;
  const int Tp = (T + 15) / 16 * 16;
  const int Tk = (T + 63) / 64 * 64;
  const int mtiles = (T + 127) / 128;
  const int tp32 = (Tp + 31) / 32 * 32;
  uint32_t tmem_cols = 64;
  while (static_cast<int>(tmem_cols) < tp32) tmem_cols <<= 1;
  const int qk_bytes = 2 * kBoxBytes + Tp * 128;
  const int p_bytes = (Tk / 64) * 16384;
  const int region_a = ((qk_bytes > p_bytes ? qk_bytes : p_bytes) + 1023) / 1024 * 1024;
  const int smem_bytes = region_a + Tp * 128 + 64 + 1024;

  CUtensorMap map;
  int rc = clm_make_tmap_bf16_2d(&map, qkv, static_cast<uint64_t>(batch) * T, 3ull * D, 3ull * D,
                                 kHeadDim, kBoxRows);
  if (rc) return rc;
  CUtensorMap map_tail;
  rc = clm_make_tmap_bf16_2d(&map_tail, qkv, static_cast<uint64_t>(batch) * T, 3ull * D, 3ull * D,
                             kHeadDim, 16);
  if (rc) return rc;
  const long long grid = static_cast<long long>(batch) * heads * mtiles;
  CLM_REQUIRE(grid < 2147483647LL, "clm_attention: grid too large");
  // algorithmic work: QK^T and PV at the true T (causal not discounted)
  ProfScope prof(CLM_K_ATTENTION, 4.0 * batch * heads * static_cast<double>(T) * T * kHeadDim,
                 2.0 * batch * T * 4.0 * D, stream);
  if (causal) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attention_kernel<true><<<static_cast<int>(grid), kThreads, smem_bytes, stream>>>(
        map, map_tail, static_cast<__nv_bfloat16*>(out), T, heads, Tp, Tk, mtiles, tmem_cols, region_a);
  } else {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attention_kernel<false><<<static_cast<int>(grid), kThreads, smem_bytes, stream>>>(
        map, map_tail, static_cast<__nv_bfloat16*>(out), T, heads, Tp, Tk, mtiles, tmem_cols, region_a);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_attention(const void* qkv_bf16, void* out_bf16, int batch, int tokens, int heads,
                             int causal, void* stream) {
  return clm_attention_launch(qkv_bf16, out_bf16, batch, tokens, heads, causal,
                              static_cast<cudaStream_t>(stream));
}
