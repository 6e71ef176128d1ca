// clm_gemm.cu — persistent, warp-specialised tcgen05/TMEM GEMM fed by TMA.
//
//   out[M,N] = epi( A[M,K]·W[N,K]^T (+ A2[M,K2]·W2[N,K2]^T) + bias ) (+ residual)
//
// Both operands are K-major bf16 (activations [rows, K]; nn.Linear weights [out, in]), so
// neither needs a transpose: TMA drops 64-column (128-byte) swizzled boxes into shared
// memory and tcgen05.mma reads them through K-major SWIZZLE_128B descriptors.
//
// CTA = 320 threads:
//   warp 0      TMA producer (one elected lane)          smem ring: full[]/empty[] mbarriers
//   warp 1      TMEM allocator + MMA issuer (one lane)   accumulator ring: tmem_full[]/tmem_empty[]
//   warps 2..9  epilogue: tcgen05.ld -> bias/QuickGELU/residual -> global stores.  Warp w reads
//               TMEM lane quarter (w % 4) and column half ((w - 2) / 4); the fp32 residual of the
//               next 32-column chunk is prefetched while the current chunk is converted/stored.
// The accumulator is double buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  The grid is persistent: min(tiles, #SM) CTAs stride over
// the tile list (n fastest, so CTAs that run together share A rows through L2).
//
// The optional (A2, W2) pair is the unmerged LoRA update folded in as extra K blocks of the
// SAME accumulator: y = x W^T + (x A^T)(s B)^T  (K4/K6 in §2b).
#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int BM = 128;
constexpr int kEpiWarps = 8;
constexpr int BK = 64;
constexpr int kNumThreads = 320;
constexpr int kAccStages = 2;

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = kAccStages * BN;  // 512 / 256 / 128: powers of two
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiWarps * 4096 /*epilogue staging*/;
};

struct EpiParams {
  void* out;
  const float* bias;
  const float* residual;
  int ldo;
  int ldr;
  int out_f32;
  int act;
};

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
            int M, int N, int kb_main, int kb_ext, EpiParams ep) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + BM - 1) / BM;
  const int num_tiles = m_tiles * n_tiles;
  const int kb_total = kb_main + kb_ext;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (kb_ext > 0) {
      tma_prefetch_desc(&map_a2);
      tma_prefetch_desc(&map_b2);
    }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
          if (kb < kb_main) {
            tma_load_2d(sa, &map_a, &full[stage], kb * BK, m0);
            tma_load_2d(sb, &map_b, &full[stage], kb * BK, n0);
          } else {
            tma_load_2d(sa, &map_a2, &full[stage], (kb - kb_main) * BK, m0);
            tma_load_2d(sb, &map_b2, &full[stage], (kb - kb_main) * BK, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t b_addr = a_addr + C::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing 16 bf16 (32 B) along K inside the 128-B swizzle atom
            const uint64_t da = umma_desc_sw128(a_addr + k * 32, 1024);
            const uint64_t db = umma_desc_sw128(b_addr + k * 32, 1024);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                          // smem slot free when MMAs done
          if (kb == kb_total - 1) umma_commit(&tmem_full[acc]);  // accumulator ready
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue warps =================
    // TMEM gives each lane one accumulator ROW (32 fp32 columns per tcgen05.ld).  Storing that
    // directly would touch 32 different cache lines per warp instruction, so each warp transposes
    // its 32x32 chunk through a private 4 KiB shared-memory buffer (16-byte pieces XOR-swizzled by
    // row: conflict-free both ways) and does bias / activation / residual / stores in the
    // COALESCED domain: lane l owns columns 4*(l%8)..+3 of rows (l/8) + 4*i, i = 0..7, so every
    // global instruction covers 4 full 128-byte row segments.
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which half of the tile's columns this warp drains
    constexpr int kChunks = BN / 64;    // 32-column chunks per warp
    uint8_t* stage = smem + C::kStages * C::kStageBytes + 256 + (warp - 2) * 4096;
    const int piece = lane & 7;         // 16-byte piece (4 fp32 columns) inside the 128-byte chunk row
    const int rsub = lane >> 3;         // row offset inside each group of 4 rows
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM + q * 32;
      const int n0 = (tile % n_tiles) * BN + half * (BN / 2);
      float4 rcur[8], rnxt[8];  // residual of the current / next chunk (coalesced layout)
      auto load_residual = [&](float4 (&dst)[8], int col0) {
        const int col = col0 + piece * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = m0 + i * 4 + rsub;
          dst[i] = (row < M && col < N)
                       ? *reinterpret_cast<const float4*>(ep.residual + static_cast<size_t>(row) * ep.ldr + col)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (ep.residual) load_residual(rcur, n0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        const int col0 = n0 + c * 32;
        if (c + 1 < kChunks && ep.residual) load_residual(rnxt, col0 + 32);
        tmem_ld_wait();
        // row-per-lane -> shared (piece j of row `lane` lands in slot j ^ (lane & 7))
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col = col0 + piece * 4;
        const bool col_ok = col < N;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias && col_ok) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + rsub;
          float4 f = *reinterpret_cast<const float4*>(stage + rl * 128 + ((piece ^ (rl & 7)) << 4));
          f.x += b.x; f.y += b.y; f.z += b.z; f.w += b.w;
          if (ep.act == CLM_EPI_QUICKGELU) {
            f.x = quick_gelu(f.x); f.y = quick_gelu(f.y); f.z = quick_gelu(f.z); f.w = quick_gelu(f.w);
          }
          if (ep.residual) {
            const float4 r = rcur[i];
            f.x += r.x; f.y += r.y; f.z += r.z; f.w += r.w;
          }
          const int row = m0 + rl;
          if (row < M && col_ok) {
            if (ep.out_f32) {
              *reinterpret_cast<float4*>(static_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col) = f;
            } else {
              *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col) =
                  make_uint2(pack_bf16x2(f.x, f.y), pack_bf16x2(f.z, f.w));
            }
          }
        }
        if (ep.residual) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rcur[i] = rnxt[i];
        }
        __syncwarp();  // the staging buffer is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma2,
                const CUtensorMap& mb2, int M, int N, int kb_main, int kb_ext, const EpiParams& ep,
                cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < clm_num_sms() ? tiles : clm_num_sms();
  gemm_kernel<BN><<<grid, kNumThreads, C::kSmemBytes, stream>>>(ma, mb, ma2, mb2, M, N, kb_main,
                                                                kb_ext, ep);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

}  // namespace

// Internal entry used by the tower code as well (same translation unit boundary as the C-ABI).
int clm_gemm_launch(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                    const void* A2, int lda2, const void* W2, int ldw2, int K2, void* out, int ldo,
                    int out_dtype, const float* bias, const float* residual, int ldr, int epilogue,
                    cudaStream_t stream) {
  CLM_REQUIRE(A && W && out, "clm_gemm_epi: null operand");
  CLM_REQUIRE(M > 0 && N > 0 && K > 0, "clm_gemm_epi: bad shape M=%d N=%d K=%d", M, N, K);
  CLM_REQUIRE(N % 8 == 0, "clm_gemm_epi: N=%d must be a multiple of 8", N);
  CLM_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0,
              "clm_gemm_epi: leading dims must be multiples of 8 (lda=%d ldw=%d ldo=%d)", lda, ldw,
              ldo);
  CLM_REQUIRE(lda >= K && ldw >= K && ldo >= N, "clm_gemm_epi: leading dim smaller than extent");
  CLM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "clm_gemm_epi: out not 16-B aligned");
  CLM_REQUIRE(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "clm_gemm_epi: bias align");
  CLM_REQUIRE(!residual || ((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && ldr % 4 == 0),
              "clm_gemm_epi: residual alignment");
  const bool has_ext = (A2 != nullptr && W2 != nullptr && K2 > 0);
  if (has_ext) {
    CLM_REQUIRE(lda2 % 8 == 0 && ldw2 % 8 == 0 && lda2 >= K2 && ldw2 >= K2,
                "clm_gemm_epi: bad extension leading dims");
  }
  const int BN = (N >= 256 && N % 256 == 0) ? 256 : ((N >= 128 && N % 128 == 0) ? 128 : (N > 128 ? 256 : (N > 64 ? 128 : 64)));
  CUtensorMap ma, mb, ma2, mb2;
  int rc;
  if ((rc = clm_make_tmap_bf16_2d(&ma, A, M, K, lda, BK, BM))) return rc;
  if ((rc = clm_make_tmap_bf16_2d(&mb, W, N, K, ldw, BK, BN))) return rc;
  if (has_ext) {
    if ((rc = clm_make_tmap_bf16_2d(&ma2, A2, M, K2, lda2, BK, BM))) return rc;
    if ((rc = clm_make_tmap_bf16_2d(&mb2, W2, N, K2, ldw2, BK, BN))) return rc;
  } else {
    ma2 = ma;
    mb2 = mb;
  }
  EpiParams ep;
  ep.out = out;
  ep.bias = bias;
  ep.residual = residual;
  ep.ldo = ldo;
  ep.ldr = ldr;
  ep.out_f32 = (out_dtype == CLM_OUT_F32);
  ep.act = epilogue;
  const int kb_main = (K + BK - 1) / BK;
  const int kb_ext = has_ext ? (K2 + BK - 1) / BK : 0;
  const double flops = 2.0 * M * N * (static_cast<double>(K) + (has_ext ? K2 : 0));
  const double bytes = 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) +
                       static_cast<double>(M) * N * (ep.out_f32 ? 4 : 2) +
                       (residual ? 4.0 * M * N : 0.0);
  ProfScope prof(CLM_K_GEMM, flops, bytes, stream);
  switch (BN) {
    case 256: return launch_gemm<256>(ma, mb, ma2, mb2, M, N, kb_main, kb_ext, ep, stream);
    case 128: return launch_gemm<128>(ma, mb, ma2, mb2, M, N, kb_main, kb_ext, ep, stream);
    default: return launch_gemm<64>(ma, mb, ma2, mb2, M, N, kb_main, kb_ext, ep, stream);
  }
}

extern "C" int clm_gemm_epi(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                            const void* A2, int lda2, const void* W2, int ldw2, int K2, void* out,
                            int ldo, int out_dtype, const float* bias, const float* residual,
                            int ldr, int epilogue, void* stream) {
  return clm_gemm_launch(A, lda, W, ldw, M, N, K, A2, lda2, W2, ldw2, K2, out, ldo, out_dtype, bias,
                         residual, ldr, epilogue, static_cast<cudaStream_t>(stream));
}
