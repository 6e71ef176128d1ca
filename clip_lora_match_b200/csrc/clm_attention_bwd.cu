// clm_attention_bwd.cu — attention backward on tcgen05 for sequences of at most 128 tokens: the shapes of the
// reference's own training configuration (config/lora_config.yaml: ViT-B/32 -> 50 vision tokens, 77 text tokens;
// scripts/train_lora.py:170-187).  Longer sequences take the CUDA-core kernels of clm_train.cu.
//
// Per (batch, head), c = 1/8, everything recomputed from q, k, v and dO (nothing but qkv is kept by the forward):
//   S  = Q K^T      dP  = dO V^T      (lanes = queries)     P = softmax(c S + mask),  delta = rowsum(P dP),
//   S' = K Q^T      dP' = V dO^T      (lanes = keys)        dS = c P (dP - delta)
//   dQ = dS K       dV = P'^T-form: P' dO        dK = dS' Q
// The four score products are SS-form UMMAs from the TMA-staged 128 x 64 tiles (K-major, 128-byte swizzle).  The
// three output products need their left operand with the CONTRACTED index along the columns: dQ contracts over
// keys (dS, lanes = queries), dV and dK contract over queries (P', dS' with lanes = keys) -- which is why the
// scores are formed twice, once per orientation: each orientation is exponentiated in place, written back to
// TMEM as packed bf16 and consumed as a TS-form A operand (as P is in the forward kernel), with K / dO / Q as
// MN-major shared-memory B operands.  No transposition through shared memory, nothing round-trips through HBM.
// The transposed orientation needs the per-QUERY statistics (log-sum-exp and delta) per COLUMN: the four
// query-lane warps publish them in shared memory before the four key-lane warps start.
// TMEM (512 columns): four score regions of Tp columns (Tp = 64 / 96 / 128), packed dS / P' / dS' in the lower
// halves of their own regions, the three 64-column accumulators behind them or in regions that are dead by then.
// One CTA per SM, one item at a time (no pipelining across items yet: the tensor pipe idles during the
// exponentials; at these sizes the step is launch bound anyway).
#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int kThreads = 288;  // warps 0-3: query-lane softmax / dQ epilogue, 4-7: key-lane / dK, dV epilogue, 8: TMA + MMA
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;

struct Params {
  int T, Tp, heads, causal, num_items;
  int col_s, col_dp, col_st, col_dpt, col_dq, col_dv, col_dk;
};

// MN-major SWIZZLE_128B descriptor: rows = the contracted index, 64 contiguous elements = the N index
__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]: A is bf16 packed two per 32-bit TMEM column (lane = row)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   __nv_bfloat16* __restrict__ dqkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = smem + kTileBytes;
  uint8_t* v_s = smem + 2 * kTileBytes;
  uint8_t* o_s = smem + 3 * kTileBytes;  // dO
  float* lse_s = reinterpret_cast<float*>(smem + 4 * kTileBytes);  // [128] per query: max + log2(sum), log2 domain
  float* delta_s = lse_s + 128;                                     // [128] per query: rowsum(P dP)
  uint64_t* bars = reinterpret_cast<uint64_t*>(delta_s + 128);      // full, mma1, mma2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = lane_id();
  const int D = p.heads * 64;
  const size_t ld = 3 * static_cast<size_t>(D);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int T = p.T, Tp = p.Tp;
  const int nch = Tp / 32;  // 32-column chunks of a score row
  const uint32_t idesc_s = umma_idesc_bf16(128, Tp, 0, 0);
  const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
  uint32_t ph = 0;

  for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ph ^= 1) {
    const int b = it / p.heads, h = it - b * p.heads;
    const int row0 = b * T;
    if (warp == 8) {
      if (elect_one_sync()) {  // uniform operands: UTCHMMA straight from uniform registers
        mbar_arrive_expect_tx(&bars[0], 4 * kTileBytes);
        tma_load_2d(q_s, &map_qkv, &bars[0], h * 64, row0);
        tma_load_2d(k_s, &map_qkv, &bars[0], D + h * 64, row0);
        tma_load_2d(v_s, &map_qkv, &bars[0], 2 * D + h * 64, row0);
        tma_load_2d(o_s, &map_do, &bars[0], h * 64, row0);
        mbar_wait(&bars[0], ph);
        tc_fence_after();
        const uint32_t qa = smem_u32(q_s), ka = smem_u32(k_s), va = smem_u32(v_s), oa = smem_u32(o_s);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S = Q K^T
          umma_bf16_ss(tmem + p.col_s, umma_desc_sw128(qa + k * 32, 1024), umma_desc_sw128(ka + k * 32, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP = dO V^T
          umma_bf16_ss(tmem + p.col_dp, umma_desc_sw128(oa + k * 32, 1024), umma_desc_sw128(va + k * 32, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S' = K Q^T
          umma_bf16_ss(tmem + p.col_st, umma_desc_sw128(ka + k * 32, 1024), umma_desc_sw128(qa + k * 32, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP' = V dO^T
          umma_bf16_ss(tmem + p.col_dpt, umma_desc_sw128(va + k * 32, 1024), umma_desc_sw128(oa + k * 32, 1024), idesc_s, k != 0);
        umma_commit(&bars[1]);
      }
      __syncwarp();
    } else {
      mbar_wait(&bars[1], ph);
      tc_fence_after();
      uint32_t sv[32], dv[32], pk[16];
      if (warp < 4) {
        // ---- query-lane orientation: statistics, then dS (packed bf16) over the lower half of the dP region
        const int i = (warp & 3) * 32 + lane;
        const int jmax = p.causal ? min(T, i + 1) : T;  // valid keys: j < jmax
        const uint32_t srow = tmem + lane_base + p.col_s, drow = tmem + lane_base + p.col_dp;
        float m = -INFINITY;
        for (int c = 0; c < nch; ++c) {
          tmem_ld_32x32b_x32(srow + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) m = (c * 32 + e < jmax) ? fmaxf(m, __uint_as_float(sv[e])) : m;
        }
        const float neg_m = -m * kScaleLog2e;
        float l = 0.f, acc = 0.f;
        for (int c = 0; c < nch; ++c) {
          tmem_ld_32x32b_x32(srow + c * 32, sv);
          tmem_ld_32x32b_x32(drow + c * 32, dv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float ex = (c * 32 + e < jmax) ? fast_exp2(fmaf(__uint_as_float(sv[e]), kScaleLog2e, neg_m)) : 0.f;
            l += ex;
            acc = fmaf(ex, __uint_as_float(dv[e]), acc);
          }
        }
        const float inv_l = 1.0f / l;
        const float delta = acc * inv_l;
        const float neg_lse = neg_m - log2f(l);  // P = 2^(s c log2e + neg_lse)
        lse_s[i] = neg_lse;
        delta_s[i] = delta;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // statistics visible to the key-lane warps
        for (int c = 0; c < nch; ++c) {
          tmem_ld_32x32b_x32(srow + c * 32, sv);
          tmem_ld_32x32b_x32(drow + c * 32, dv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float d0 = 0.f, d1 = 0.f;
            if (c * 32 + e < jmax)
              d0 = 0.125f * fast_exp2(fmaf(__uint_as_float(sv[e]), kScaleLog2e, neg_lse)) * (__uint_as_float(dv[e]) - delta);
            if (c * 32 + e + 1 < jmax)
              d1 = 0.125f * fast_exp2(fmaf(__uint_as_float(sv[e + 1]), kScaleLog2e, neg_lse)) * (__uint_as_float(dv[e + 1]) - delta);
            pk[e / 2] = pack_bf16x2(d0, d1);
          }
          tmem_st_x16(drow + c * 16, pk);  // columns of dP chunk c/2: already consumed
        }
        tmem_st_wait();
      } else {
        // ---- key-lane orientation: P' and dS' (packed bf16) over the lower halves of the S' and dP' regions
        const int j = (warp & 3) * 32 + lane;
        const uint32_t srow = tmem + lane_base + p.col_st, drow = tmem + lane_base + p.col_dpt;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t pk2[16];
        for (int c = 0; c < nch; ++c) {
          tmem_ld_32x32b_x32(srow + c * 32, sv);
          tmem_ld_32x32b_x32(drow + c * 32, dv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            const int i0 = c * 32 + e;
            if (j < T && i0 < T && (!p.causal || j <= i0)) {
              p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), kScaleLog2e, lse_s[i0]));
              d0 = 0.125f * p0 * (__uint_as_float(dv[e]) - delta_s[i0]);
            }
            if (j < T && i0 + 1 < T && (!p.causal || j <= i0 + 1)) {
              p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), kScaleLog2e, lse_s[i0 + 1]));
              d1 = 0.125f * p1 * (__uint_as_float(dv[e + 1]) - delta_s[i0 + 1]);
            }
            pk[e / 2] = pack_bf16x2(p0, p1);
            pk2[e / 2] = pack_bf16x2(d0, d1);
          }
          tmem_st_x16(srow + c * 16, pk);
          tmem_st_x16(drow + c * 16, pk2);
        }
        tmem_st_wait();
      }
    }
    tc_fence_before();
    __syncthreads();  // every packed operand is in TMEM, every score column is dead
    if (warp == 8) {
      if (elect_one_sync()) {  // uniform operands: UTCHMMA straight from uniform registers
        tc_fence_after();
        const uint32_t qa = smem_u32(q_s), ka = smem_u32(k_s), oa = smem_u32(o_s);
        const int nks = Tp / 16;
        for (int ks = 0; ks < nks; ++ks)  // dQ = dS K (contract over keys)
          umma_ts(tmem + p.col_dq, tmem + p.col_dp + ks * 8, desc_sw128_mn(ka + ks * 2048), idesc_o, ks != 0);
        for (int ks = 0; ks < nks; ++ks)  // dV = P' dO (contract over queries)
          umma_ts(tmem + p.col_dv, tmem + p.col_st + ks * 8, desc_sw128_mn(oa + ks * 2048), idesc_o, ks != 0);
        for (int ks = 0; ks < nks; ++ks)  // dK = dS' Q
          umma_ts(tmem + p.col_dk, tmem + p.col_dpt + ks * 8, desc_sw128_mn(qa + ks * 2048), idesc_o, ks != 0);
        umma_commit(&bars[2]);
      }
      __syncwarp();
    } else {
      mbar_wait(&bars[2], ph);
      tc_fence_after();
      const int r = (warp & 3) * 32 + lane;  // query (warps 0-3) or key (warps 4-7)
      uint32_t v16[16];
      const int nout = warp < 4 ? 1 : 2;
      for (int o = 0; o < nout; ++o) {
        const int col = warp < 4 ? p.col_dq : (o == 0 ? p.col_dk : p.col_dv);
        const int sect = warp < 4 ? 0 : (o == 0 ? 1 : 2);  // dq | dk | dv third of the output row
        uint4* dst = reinterpret_cast<uint4*>(dqkv + (static_cast<size_t>(row0) + r) * ld + sect * D + h * 64);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32b_x16(tmem + lane_base + col + c * 16, v16);
          tmem_ld_wait();
          if (r < T) {
            uint4 a, bq;
            a.x = pack_bf16x2(__uint_as_float(v16[0]), __uint_as_float(v16[1]));
            a.y = pack_bf16x2(__uint_as_float(v16[2]), __uint_as_float(v16[3]));
            a.z = pack_bf16x2(__uint_as_float(v16[4]), __uint_as_float(v16[5]));
            a.w = pack_bf16x2(__uint_as_float(v16[6]), __uint_as_float(v16[7]));
            bq.x = pack_bf16x2(__uint_as_float(v16[8]), __uint_as_float(v16[9]));
            bq.y = pack_bf16x2(__uint_as_float(v16[10]), __uint_as_float(v16[11]));
            bq.z = pack_bf16x2(__uint_as_float(v16[12]), __uint_as_float(v16[13]));
            bq.w = pack_bf16x2(__uint_as_float(v16[14]), __uint_as_float(v16[15]));
            dst[2 * c] = a;
            dst[2 * c + 1] = bq;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // accumulators drained, tiles free for the next item's TMA
    tc_fence_after();
  }
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace

// 1 if the tcgen05 kernel handles this shape
int clm_attention_bwd_tc_supported(int tokens) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CLM_ATTN_BWD_TC");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && tokens >= 1 && tokens <= 128;
}

int clm_attention_bwd_tc_launch(const void* qkv, const void* dout, void* dqkv, int batch, int tokens, int heads,
                                int causal, cudaStream_t stream) {
  const int D = heads * 64;
  Params p;
  p.T = tokens;
  p.Tp = tokens <= 64 ? 64 : (tokens <= 96 ? 96 : 128);
  p.heads = heads;
  p.causal = causal;
  p.num_items = batch * heads;
  p.col_s = 0;
  p.col_dp = p.Tp;
  p.col_st = 2 * p.Tp;
  p.col_dpt = 3 * p.Tp;
  if (p.Tp == 64) {         // 256 score columns, accumulators behind them
    p.col_dq = 256; p.col_dv = 320; p.col_dk = 384;
  } else if (p.Tp == 96) {  // 384 score columns; dQ in the dead S region
    p.col_dq = 0; p.col_dv = 384; p.col_dk = 448;
  } else {                  // 512 score columns; S is dead, the upper half of dP is dead
    p.col_dq = 0; p.col_dv = 64; p.col_dk = 128 + 64;
  }
  CUtensorMap mq, mo;
  int rc;
  const uint64_t rows = static_cast<uint64_t>(batch) * tokens;
  if ((rc = clm_make_tmap_bf16_2d(&mq, qkv, rows, 3ull * D, 3ull * D, 64, 128))) return rc;
  if ((rc = clm_make_tmap_bf16_2d(&mo, dout, rows, D, D, 64, 128))) return rc;
  const size_t smem = 4 * kTileBytes + 2 * 128 * sizeof(float) + 64 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done = true;
  }
  const int grid = p.num_items < clm_num_sms() ? p.num_items : clm_num_sms();
  // four score products + three output products, all on 128-row tiles (the hardware's work, not the algorithm's)
  const double flops = 10.0 * batch * heads * static_cast<double>(tokens) * tokens * 64;
  const double bytes = 2.0 * batch * tokens * heads * 64 * (3 + 1 + 3);
  ProfScope prof(CLM_K_ATTENTION, flops, bytes, stream);
  attn_bwd_tc_kernel<<<grid, kThreads, smem, stream>>>(mq, mo, static_cast<__nv_bfloat16*>(dqkv), p);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}
