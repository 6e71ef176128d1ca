// clm_attention.cu — fused attention for CLIP ViT towers on tcgen05 / TMEM.
//
// CLIP sequences are short (T = 50 / 77 / 197 / 257), so the whole score row block of a
// 128-query tile fits in tensor memory and no online softmax is needed.  The kernel is
// persistent (one CTA per SM) and warp-specialised; a work item is one (batch, head):
//
//   warp 0      TMA producer: Q, K, V of the next items (128-byte-swizzled boxes straight out of
//               the fused QKV activation [B*T, 3D]) into a ring of shared-memory stages.
//               K and V are loaded ONCE per (batch, head) and shared by all its query tiles.
//   warp 1      MMA issuer.  Per 128-query tile t:  S = Q K^T  (SS form, K-major A and B) into
//               TMEM slot t % 2;  later  O = P V  (TS form: P is read from TMEM, V is an MN-major
//               shared-memory operand, so neither P nor V^T is ever materialised in smem/HBM).
//   warps 2-5 / 6-9   two softmax groups that alternate tiles (ping-pong): thread = query row;
//               pass 1 row max, pass 2 p = 2^(s*c - max*c) (one FFMA + one MUFU.EX2), row sum,
//               P written back to TMEM as packed bf16 over the columns S no longer needs;
//               then O is read from TMEM, scaled by 1/rowsum and stored as bf16.
//
// While group A runs softmax on tile t, the tensor core computes S(t+1) and group B starts on it;
// the TMA warp is one or more items ahead, so HBM latency is off the critical path.
// TMEM slot layout (columns): [0,Tp) S fp32 -> [0,Tp/2) P bf16x2 (aliases S) ; O fp32 at o_off.
// Softmax is fp32 (modeling_clip.py:274); scale 1/sqrt(64) (modeling_clip.py:271).
#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int kThreads = 320;
constexpr int kHeadDim = 64;
constexpr int kMaxStages = 6;

struct AttnParams {
  int T, H, Tp, mtiles, num_items, stages, stage_bytes, nslots, slot_cols, o_off;
};

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

// D[tmem] (+)= A[tmem] * B[smem]: A is bf16 packed two per 32-bit TMEM column (lane = row).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <bool kCausal>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map16,
                 __nv_bfloat16* __restrict__ out, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* stage_full = bars;                     // [kMaxStages]
  uint64_t* stage_empty = bars + kMaxStages;       // [kMaxStages]
  uint64_t* s_full = bars + 2 * kMaxStages;        // [2] S ready (MMA commit)
  uint64_t* p_full = s_full + 2;                   // [2] P written (128 softmax threads)
  uint64_t* o_full = s_full + 4;                   // [2] O ready (MMA commit)
  uint64_t* slot_free = s_full + 6;                // [2] O drained (128 softmax threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T, H = p.H, Tp = p.Tp, D = p.H * kHeadDim;
  const int kv_bytes = Tp * 128;  // one of Q / K / V in a stage

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map64);
    tma_prefetch_desc(&map16);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&stage_full[s], 1);
      mbar_init(&stage_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&o_full[s], 1);
      mbar_init(&slot_free[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int n64 = Tp / 64, n16 = (Tp % 64) / 16;
      int st = 0;
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const int b = it / H, h = it % H;
        const int row_base = b * T;
        mbar_wait(&stage_empty[st], ph ^ 1);
        uint8_t* base = smem + st * p.stage_bytes;
        mbar_arrive_expect_tx(&stage_full[st], static_cast<uint32_t>(3 * kv_bytes));
        for (int part = 0; part < 3; ++part) {  // 0 = Q, 1 = K, 2 = V
          uint8_t* dst = base + part * kv_bytes;
          const int col = part * D + h * kHeadDim;
          for (int i = 0; i < n64; ++i)
            tma_load_2d(dst + i * 8192, &map64, &stage_full[st], col, row_base + i * 64);
          for (int i = 0; i < n16; ++i)
            tma_load_2d(dst + n64 * 8192 + i * 2048, &map16, &stage_full[st], col,
                        row_base + n64 * 64 + i * 16);
        }
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc_pv = umma_idesc_bf16(128, kHeadDim, 0, 1);
    int st = 0;
    uint32_t ph = 0;
    int t = 0;                 // running tile counter of this CTA
    int prev_slot = -1, prev_use = 0, prev_stage = 0, prev_last = 0;
    uint32_t prev_v_addr = 0;

    auto issue_pv = [&]() {
      // O(prev) = P(prev) V : P from TMEM (written by the softmax group), V MN-major from smem
      mbar_wait(&p_full[prev_slot], prev_use & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sbase = tmem + static_cast<uint32_t>(prev_slot * p.slot_cols);
        const int ksteps = Tp / 16;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_bf16_ts(sbase + p.o_off, sbase + ks * 8, umma_desc_sw128_mn(prev_v_addr + ks * 2048),
                       idesc_pv, ks != 0 ? 1u : 0u);
        umma_commit(&o_full[prev_slot]);
        if (prev_last) umma_commit(&stage_empty[prev_stage]);  // stage reusable once these MMAs retire
      }
      __syncwarp();
      prev_slot = -1;
    };

    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      mbar_wait(&stage_full[st], ph);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(smem + st * p.stage_bytes);
      const uint32_t k_addr = q_addr + kv_bytes;
      const uint32_t v_addr = k_addr + kv_bytes;
      for (int mt = 0; mt < p.mtiles; ++mt, ++t) {
        const int slot = t % p.nslots;
        const int use = t / p.nslots;
        if (p.nslots == 1 && prev_slot >= 0) issue_pv();  // single slot: O(t-1) must drain first
        mbar_wait(&slot_free[slot], (use & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sbase = tmem + static_cast<uint32_t>(slot * p.slot_cols);
          for (int n0 = 0; n0 < Tp; n0 += 256) {
            const int nn = (Tp - n0) < 256 ? (Tp - n0) : 256;
            const uint32_t idesc = umma_idesc_bf16(128, nn, 0, 0);
#pragma unroll
            for (int k = 0; k < kHeadDim / 16; ++k)
              umma_bf16_ss(sbase + n0, umma_desc_sw128(q_addr + mt * 16384 + k * 32, 1024),
                           umma_desc_sw128(k_addr + n0 * 128 + k * 32, 1024), idesc, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[slot]);
        }
        __syncwarp();
        if (prev_slot >= 0) issue_pv();
        prev_slot = slot; prev_use = use; prev_stage = st; prev_v_addr = v_addr;
        prev_last = (mt == p.mtiles - 1);
      }
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
    if (prev_slot >= 0) issue_pv();
  } else {
    // ================= softmax groups =================
    const int g = (warp - 2) >> 2;       // group 0: warps 2-5, group 1: warps 6-9
    const int q = warp & 3;              // TMEM lane quarter of this warp
    const int r = q * 32 + lane;         // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int nchunks = (Tp + 31) / 32;
    constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    int t = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const int b = it / H, h = it % H;
      for (int mt = 0; mt < p.mtiles; ++mt, ++t) {
        // two TMEM slots: groups ping-pong (tile t -> group t % 2, slot t % 2).  A single slot
        // (Tp > 256) serialises tiles, so one group takes them all: a group must never wait on a
        // barrier phase more than one ahead of the one it last consumed.
        if ((p.nslots == 2 ? (t & 1) : 0) != g) continue;
        const int slot = t % p.nslots;
        const uint32_t par = static_cast<uint32_t>((t / p.nslots) & 1);
        const uint32_t trow = tmem + lane_off + static_cast<uint32_t>(slot * p.slot_cols);
        const int qi = mt * 128 + r;                     // query position in the sequence
        const bool warp_live = mt * 128 + q * 32 < T;    // does this warp own any real row?
        int valid = T;
        if (kCausal) valid = (qi + 1 < T) ? qi + 1 : T;

        mbar_wait(&s_full[slot], par);
        tc_fence_after();
        float sum = 1.f;
        if (warp_live) {
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
          const float neg_mx = -mx * kScaleLog2e;
          sum = 0.f;
          for (int c = 0; c < nchunks; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(trow + c * 32, v);
            tmem_ld_wait();
            float e[32];
            if (c * 32 + 32 <= valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                e[i] = fast_exp2(fmaf(__uint_as_float(v[i]), kScaleLog2e, neg_mx));
                sum += e[i];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = fast_exp2(fmaf(__uint_as_float(v[i]), kScaleLog2e, neg_mx));
                e[i] = (c * 32 + i < valid) ? x : 0.f;
                sum += e[i];
              }
            }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
            tmem_st_32x32b_x16(trow + c * 16, pk);  // P aliases the S columns already consumed
          }
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&p_full[slot]);

        mbar_wait(&o_full[slot], par);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        if (warp_live) {
          tmem_ld_32x32b_x32(trow + p.o_off, o0);
          tmem_ld_32x32b_x32(trow + p.o_off + 32, o1);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&slot_free[slot]);  // the slot can take S(t + nslots) while we store
        if (warp_live && qi < T) {
          const float inv = 1.0f / sum;
          uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(b * T + qi) * D + h * kHeadDim);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o0[8 * jj + 0]) * inv, __uint_as_float(o0[8 * jj + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o0[8 * jj + 2]) * inv, __uint_as_float(o0[8 * jj + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o0[8 * jj + 4]) * inv, __uint_as_float(o0[8 * jj + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o0[8 * jj + 6]) * inv, __uint_as_float(o0[8 * jj + 7]) * inv);
            o4[jj] = o;
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o1[8 * jj + 0]) * inv, __uint_as_float(o1[8 * jj + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o1[8 * jj + 2]) * inv, __uint_as_float(o1[8 * jj + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o1[8 * jj + 4]) * inv, __uint_as_float(o1[8 * jj + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o1[8 * jj + 6]) * inv, __uint_as_float(o1[8 * jj + 7]) * inv);
            o4[4 + jj] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

int clm_attention_launch(const void* qkv, void* out, int batch, int tokens, int heads, int causal,
                         cudaStream_t stream) {
  CLM_REQUIRE(qkv && out && batch >= 0 && tokens > 0 && heads > 0, "clm_attention: bad argument");
  CLM_REQUIRE(tokens <= 512, "clm_attention: tokens=%d > 512 unsupported (CLIP uses <= 257)", tokens);
  if (batch == 0) return CLM_OK;
  const int T = tokens;
  const int D = heads * kHeadDim;
  AttnParams p;
  p.T = T;
  p.H = heads;
  p.Tp = (T + 15) / 16 * 16;
  p.mtiles = (T + 127) / 128;
  const long long items = static_cast<long long>(batch) * heads;
  CLM_REQUIRE(items < 2147483647LL, "clm_attention: too many (batch, head) items");
  p.num_items = static_cast<int>(items);
  // a stage holds Q, K, V ([Tp,64] bf16 each); the last query tile's 128-row UMMA window may run
  // past Q into K/V (finite values, rows never stored), so a stage is at least mtiles*16 KiB
  int stage_bytes = 3 * p.Tp * 128;
  if (stage_bytes < p.mtiles * 16384) stage_bytes = p.mtiles * 16384;
  p.stage_bytes = stage_bytes;
  const int smem_budget = 227 * 1024 - 1024 - 256;
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  CLM_REQUIRE(stages >= 1, "clm_attention: tokens=%d needs %d bytes of shared memory per stage", T,
              stage_bytes);
  p.stages = stages;
  // TMEM slot: S needs round_up(Tp,32) fp32 columns; P (bf16x2) reuses the first Tp/2; O (64
  // columns) sits at the next multiple of 64 past P
  p.o_off = ((p.Tp / 2) + 63) / 64 * 64;
  const int s_cols = (p.Tp + 31) / 32 * 32;
  const int need = (s_cols > p.o_off + 64) ? s_cols : p.o_off + 64;
  CLM_REQUIRE(need <= 512, "clm_attention: tokens=%d needs %d TMEM columns", T, need);
  p.nslots = need <= 256 ? 2 : 1;
  p.slot_cols = p.nslots == 2 ? 256 : 512;
  const int smem_bytes = stages * stage_bytes + 256 + 1024;

  CUtensorMap map64, map16;
  int rc = clm_make_tmap_bf16_2d(&map64, qkv, static_cast<uint64_t>(batch) * T, 3ull * D, 3ull * D,
                                 kHeadDim, 64);
  if (rc) return rc;
  rc = clm_make_tmap_bf16_2d(&map16, qkv, static_cast<uint64_t>(batch) * T, 3ull * D, 3ull * D,
                             kHeadDim, 16);
  if (rc) return rc;
  const int grid = p.num_items < clm_num_sms() ? p.num_items : clm_num_sms();
  // algorithmic work: QK^T and PV at the true T (causal not discounted)
  ProfScope prof(CLM_K_ATTENTION, 4.0 * batch * heads * static_cast<double>(T) * T * kHeadDim,
                 2.0 * batch * T * 4.0 * D, stream);
  if (causal) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attention_kernel<true><<<grid, kThreads, smem_bytes, stream>>>(
        map64, map16, static_cast<__nv_bfloat16*>(out), p);
  } else {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attention_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(
        map64, map16, static_cast<__nv_bfloat16*>(out), p);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_attention(const void* qkv_bf16, void* out_bf16, int batch, int tokens, int heads,
                             int causal, void* stream) {
  return clm_attention_launch(qkv_bf16, out_bf16, batch, tokens, heads, causal,
                              static_cast<cudaStream_t>(stream));
}
