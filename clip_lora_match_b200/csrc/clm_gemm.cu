// clm_gemm.cu — persistent, warp-specialised tcgen05/TMEM GEMM fed by TMA.
//
//   out[M,N] = epi( A[M,K]·W[N,K]^T (+ A2[M,K2]·W2[N,K2]^T) + bias ) (+ residual)
//
// Both operands are K-major bf16 (activations [rows, K]; nn.Linear weights [out, in]), so
// neither needs a transpose: TMA drops 64-column (128-byte) swizzled boxes into shared
// memory and tcgen05.mma reads them through K-major SWIZZLE_128B descriptors.
//
// Two variants of one kernel (template kCtas):
//   kCtas = 2  CTA pair (cluster of 2, tcgen05 cta_group::2): a 256 x 256 tile per pair; each CTA
//              TMA-loads its own 128 rows of A and HALF of the B tile, the leader CTA's MMA thread
//              issues UMMA M=256 that reads both CTAs' shared memory and writes both CTAs' TMEM.
//              Per-SM operand traffic drops by a third (32 KiB instead of 48 KiB per k-block) and
//              the freed shared memory gives a 6-deep ring.  Used when N % 256 == 0.
//   kCtas = 1  single CTA, 128 x BN tile (BN = 256 / 128 / 64): everything else (LoRA
//              down-projection N = 64, tiny shapes).
//
// CTA = 320 threads:
//   warp 0      TMA producer (one elected lane)          smem ring: full[]/empty[] mbarriers
//   warp 1      TMEM allocator + MMA issuer (one lane)   accumulator ring: tmem_full[]/tmem_empty[]
//   warps 2..9  epilogue.  Warp w reads TMEM lane quarter (w % 4), column half ((w - 2) / 4):
//               tcgen05.ld (lane = accumulator row) -> bias / QuickGELU in registers -> 128-byte
//               swizzled rows in a private 4 KiB staging buffer -> ONE TMA tile store per 32x64
//               (bf16) / 32x32 (fp32) slab.  The in-place residual form (out == residual) issues a
//               TMA reduce-add instead, so the fp32 residual stream is updated inside the L2 and
//               never travels to the SM.  Ragged M / N edges are clipped by the TMA unit.
// The accumulator is double buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  The grid is persistent and strides over the tile list (n
// fastest, so CTAs that run together share A rows through L2).
//
// The optional (A2, W2) pair is the unmerged LoRA update folded in as extra K blocks of the
// SAME accumulator: y = x W^T + (x A^T)(s B)^T  (K4/K6 in §2b).
#include <stdlib.h>

#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int BM = 128;  // rows per CTA
constexpr int kEpiWarps = 8;
constexpr int BK = 64;
constexpr int kNumThreads = 320;
constexpr int kAccStages = 2;
// Warp roles.  CLM_GEMM_CTL_HI=1 puts the two control warps (TMA producer, MMA issuer) at the HIGHEST warp ids
// of the CTA and the epilogue warps at 0..7: the sub-partition arbiter prefers the highest eligible warp id, so
// the single-thread issue loops are not queued behind epilogue warps that are busy with bias / QuickGELU math.
// Measured in the ViT-L/14 step (tools/step_profile.py, same-call A/B, twice): 88.40 / 88.83 ms -> 88.21 / 87.75 ms.
#ifndef CLM_GEMM_CTL_HI
#define CLM_GEMM_CTL_HI 1
#endif
// bf16 epilogues: branch-free, one-piece-ahead loads of the per-column constants (0 = the piece-by-piece form)
#ifndef CLM_GEMM_EPI_PREFETCH
#define CLM_GEMM_EPI_PREFETCH 1
#endif
__device__ float g_zero_f32[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
constexpr int kTmaWarp = CLM_GEMM_CTL_HI ? 8 : 0;
constexpr int kMmaWarp = CLM_GEMM_CTL_HI ? 9 : 1;
constexpr int kEpiWarp0 = CLM_GEMM_CTL_HI ? 0 : 2;

template <int BN, int kCtas>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / kCtas) * BK * 2;  // this CTA's share of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kStageBytes == 49152) ? 4 : (kStageBytes == 32768 ? 6 : 8);
  static constexpr int kTmemCols = kAccStages * BN;  // 512 / 256 / 128: powers of two
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 1024 /*align*/ + kEpiWarps * 4096 /*epilogue staging*/ + 256 /*barriers*/;
};

// epilogue variants (template parameter kEpi)
constexpr int kEpiLegacy = 0;     // per-thread global loads/stores; residual != out or bf16 out + residual
constexpr int kEpiStoreBf16 = 1;  // TMA store of bf16 tiles
constexpr int kEpiStoreF32 = 2;   // TMA store of fp32 tiles
constexpr int kEpiReduceF32 = 3;  // TMA reduce-add of fp32 tiles: out (== residual) += acc + bias
constexpr int kEpiReduceBf16 = 4; // TMA reduce-add of bf16 tiles into a bf16 residual stream (out == residual, bf16):
                                  // the L2 adds bf16(acc + bias) to the stored bf16 value, one rounding per update

struct EpiParams {
  void* out;
  const float* bias;
  const float* residual;
  int ldo;
  int ldr;
  int out_f32;
  int act;
  // split-K (reduce-add epilogue only): a tile's k-blocks are cut into ksplit ranges of kb_per blocks, each range a
  // work unit of its own that adds its partial product into out; range 0 also adds the bias.  1 = off.
  int ksplit;
  int kb_per;
  // LayerNorm folded into the GEMM (clm_gemm_ln_epi): A is the raw bf16 residual stream, W carries gamma, and the
  // epilogue applies the per-row statistics.  ln_mode 1: out = rstd (acc - mean * col_s) + bias;
  // ln_mode 2 (LoRA down-projection, to be scaled by rstd later): out = (acc - mean * col_s) + bias.  row_stats = NULL: off.
  const float2* row_stats;
  const float* col_s;
  int ln_mode;
  // walk the tile list from its end (last rows first): a scheduling hint for L2 reuse between consecutive kernels
  // (CLM_EPI_REVERSE), no effect on the result
  int reverse;
};

__device__ __forceinline__ float ln_fix(float v, float b, float s, float neg_mu, float alpha) {
  return fmaf(alpha, fmaf(neg_mu, s, v), b);
}

// ---- cta_group::2 helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32  remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64  _, [remAddr32];\n\t"
      "}"
      :
      : "r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int BN, int kCtas, int kEpi>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
            const __grid_constant__ CUtensorMap map_out, int M, int N, int kb_main, int kb_ext,
            EpiParams ep) {
  using C = Cfg<BN, kCtas>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + C::kStages * C::kStageBytes;  // 1024-byte aligned: kEpiWarps x 4 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kEpiWarps * 4096);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

  // warp index through a shuffle: provably warp-uniform, so the role dispatch below is a uniform branch and ptxas
  // keeps the single-thread issue loops on the uniform datapath (see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCtas == 2) ? cluster_ctarank() : 0u;  // position inside the CTA pair
  const int worker = (kCtas == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_workers = (kCtas == 2) ? (gridDim.x >> 1) : gridDim.x;
  constexpr int TM = BM * kCtas;  // tile rows per worker
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + TM - 1) / TM;
  const int ksplit = ep.ksplit;
  const int num_tiles = m_tiles * n_tiles * ksplit;  // work units: (tile, k range), k range fastest
  const int kb_total = kb_main + kb_ext;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (kb_ext > 0) {
      tma_prefetch_desc(&map_a2);
      tma_prefetch_desc(&map_b2);
    }
    if (kEpi != kEpiLegacy) tma_prefetch_desc(&map_out);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps * kCtas);  // one arrive per epilogue warp of the pair
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (kCtas == 2) {
      tmem_alloc_2sm(tmem_base_slot, C::kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_base_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);  // (uniform for the same reason)
  pdl_wait();     // everything above overlapped the tail of the previous kernel; its outputs are visible from here on
  pdl_trigger();

  if (warp == kTmaWarp) {
    // ================= TMA producer (every CTA loads its own A rows and its share of B) ======
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = worker; unit < num_tiles; unit += num_workers) {
        const int tile = unit / ksplit, ks = unit - tile * ksplit;
        const int kb_lo = ks * ep.kb_per, kb_hi = (ksplit == 1 || kb_lo + ep.kb_per > kb_total) ? kb_total : kb_lo + ep.kb_per;
        const int tpos = ep.reverse ? (m_tiles * n_tiles - 1 - tile) : tile;  // reverse: last rows first
        const int m0 = (tpos / n_tiles) * TM + static_cast<int>(rank) * BM;
        const int n0 = (tpos % n_tiles) * BN + static_cast<int>(rank) * (BN / kCtas);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          const CUtensorMap* ma = kb < kb_main ? &map_a : &map_a2;
          const CUtensorMap* mb = kb < kb_main ? &map_b : &map_b2;
          const int kc = (kb < kb_main ? kb : kb - kb_main) * BK;
          if (kCtas == 2) {
            // the leader's barrier collects the bytes of BOTH CTAs
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
            tma_load_2d_2sm(sa, ma, &full[stage], kc, m0);
            tma_load_2d_2sm(sb, mb, &full[stage], kc, n0);
          } else {
            mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
            tma_load_2d(sa, ma, &full[stage], kc, m0);
            tma_load_2d(sb, mb, &full[stage], kc, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, 0, 0);
      // K-major SWIZZLE_128B operand descriptor (umma_desc_sw128 with SBO = 1024): the high word is constant, the low
      // word is (1 << 16) | (shared-memory address >> 4)
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = worker; unit < num_tiles; unit += num_workers) {
        const int ks = unit % ksplit;
        const int kb_lo = ks * ep.kb_per, kb_hi = (ksplit == 1 || kb_lo + ep.kb_per > kb_total) ? kb_total : kb_lo + ep.kb_per;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          // Every lane computes the (warp-uniform) operands; one elected lane issues.  With the operands provably
          // uniform the four tcgen05.mma of a k-block are four UTCHMMA on uniform registers; under `if (lane == 0)`
          // ptxas wrapped EACH of them (and each commit) in an ELECT / 5 x R2UR / branch loop -- ~150 instructions per
          // k-block on the one warp whose issue rate paces the whole GEMM (the issuer was never seen waiting for
          // operands or for TMEM, profiles/r2_ncu_l14_layer_gemms_bf16stream.md).
          const uint32_t a_lo = (1u << 16) | (((smem_base + static_cast<uint32_t>(stage) * C::kStageBytes) & 0x3FFFFu) >> 4);
          const uint32_t b_lo = a_lo + (C::kABytes >> 4);
          const bool last_kb = (kb == kb_hi - 1);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing 16 bf16 (32 B = 2 descriptor units) along K inside the 128-B swizzle atom
              const uint64_t da = (static_cast<uint64_t>(kDescHi) << 32) | (a_lo + 2u * k);
              const uint64_t db = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo + 2u * k);
              if (kCtas == 2) umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb != kb_lo || k != 0) ? 1u : 0u);
              else umma_bf16_ss(tmem_d, da, db, idesc, (kb != kb_lo || k != 0) ? 1u : 0u);
            }
            if (kCtas == 2) {
              umma_commit_2sm(&empty[stage]);                 // frees the slot in both CTAs
              if (last_kb) umma_commit_2sm(&tmem_full[acc]);  // accumulators ready (both)
            } else {
              umma_commit(&empty[stage]);
              if (last_kb) umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (kEpi != kEpiLegacy) {
    // ================= epilogue warps (TMA store / reduce path) =================
    // tcgen05.ld hands each lane one accumulator ROW; a slab is 128 bytes of that row (64 bf16 or
    // 32 fp32 columns).  Lane r writes its 8 16-byte pieces to staging row r with piece p at slot
    // p ^ (r & 7) -- the SWIZZLE_128B pattern of the output tensor map, and conflict-free for the
    // eight lanes of every shared-memory wavefront.  One elected lane then hands the 32-row slab to
    // the TMA unit; the buffer is reused once the unit has READ it (wait_group.read).
    constexpr bool kF32 = (kEpi == kEpiStoreF32 || kEpi == kEpiReduceF32);
    constexpr int kSlabCols = kF32 ? 32 : 64;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;  // which half of the tile's columns this warp drains
    // BN = 64 with bf16 output: one 64-column slab, drained by the half-0 warps only
    constexpr bool kSplit = (BN / 2 >= kSlabCols);
    constexpr int kWarpCols = kSplit ? BN / 2 : BN;
    constexpr int kSlabs = kWarpCols / kSlabCols;
    const bool active = kSplit || half == 0;
    const uint32_t stg = smem_u32(staging + (warp - kEpiWarp0) * 4096);
    const uint32_t my_row = stg + static_cast<uint32_t>(lane) * 128u;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = worker; unit < num_tiles; unit += num_workers) {
      const int tile = unit / ksplit;
      const float* bias = (unit - tile * ksplit == 0) ? ep.bias : nullptr;  // the first k range of a tile adds the bias
      const int tpos = ep.reverse ? (m_tiles * n_tiles - 1 - tile) : tile;
      const int m0 = (tpos / n_tiles) * TM + static_cast<int>(rank) * BM + q * 32;
      const int n0 = (tpos % n_tiles) * BN + (kSplit ? half * (BN / 2) : 0);
      const bool rows_live = active && m0 < M;
      const bool ln = ep.row_stats != nullptr;  // warp-uniform
      float ln_mu = 0.f, ln_alpha = 1.f;  // ln_mu holds -mean; loaded while the tile's MMAs are still running
      if (ln && rows_live && m0 + lane < M) {  // tcgen05.ld: lane = accumulator row
        const float2 st = __ldg(ep.row_stats + m0 + lane);
        ln_mu = -st.x;
        if (ep.ln_mode == 1) ln_alpha = st.y;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + (kSplit ? half * (BN / 2) : 0));
#pragma unroll
      for (int sl = 0; sl < kSlabs; ++sl) {
        const int col0 = n0 + sl * kSlabCols;
        const bool live = rows_live && col0 < N;  // warp-uniform
        uint32_t pk[32];
        if (kF32) {
          uint32_t v[32];
          if (live) {
            tmem_ld_32x32b_x32(taddr + sl * kSlabCols, v);
            tmem_ld_wait();
          }
          if (sl == kSlabs - 1) {  // accumulator fully drained: hand the TMEM stage back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (kCtas == 2) mbar_arrive_cta(&tmem_empty[acc], 0);
              else mbar_arrive(&tmem_empty[acc]);
            }
          }
          if (live) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias && col0 + 4 * p < N) b = __ldg(reinterpret_cast<const float4*>(bias + col0 + 4 * p));
              float x0, x1, x2, x3;
              if (ln) {
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col0 + 4 * p < N) cs = __ldg(reinterpret_cast<const float4*>(ep.col_s + col0 + 4 * p));
                x0 = ln_fix(__uint_as_float(v[4 * p + 0]), b.x, cs.x, ln_mu, ln_alpha);
                x1 = ln_fix(__uint_as_float(v[4 * p + 1]), b.y, cs.y, ln_mu, ln_alpha);
                x2 = ln_fix(__uint_as_float(v[4 * p + 2]), b.z, cs.z, ln_mu, ln_alpha);
                x3 = ln_fix(__uint_as_float(v[4 * p + 3]), b.w, cs.w, ln_mu, ln_alpha);
              } else {
                x0 = __uint_as_float(v[4 * p + 0]) + b.x; x1 = __uint_as_float(v[4 * p + 1]) + b.y;
                x2 = __uint_as_float(v[4 * p + 2]) + b.z; x3 = __uint_as_float(v[4 * p + 3]) + b.w;
              }
              if (ep.act == CLM_EPI_QUICKGELU) {
                x0 = quick_gelu(x0); x1 = quick_gelu(x1); x2 = quick_gelu(x2); x3 = quick_gelu(x3);
              }
              pk[4 * p + 0] = __float_as_uint(x0); pk[4 * p + 1] = __float_as_uint(x1);
              pk[4 * p + 2] = __float_as_uint(x2); pk[4 * p + 3] = __float_as_uint(x3);
            }
          }
        } else {
          uint32_t v0[32], v1[32];
          if (live) {
            tmem_ld_32x32b_x32(taddr + sl * kSlabCols, v0);
            tmem_ld_32x32b_x32(taddr + sl * kSlabCols + 32, v1);
            tmem_ld_wait();
          }
          if (sl == kSlabs - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (kCtas == 2) mbar_arrive_cta(&tmem_empty[acc], 0);
              else mbar_arrive(&tmem_empty[acc]);
            }
          }
#if CLM_GEMM_EPI_PREFETCH
          if (live) {
            // Per-column constants (bias, col_sums) without a branch and one piece AHEAD of the arithmetic: clamped
            // column index (N % 8 == 0), a zero vector stands in for an absent array, and ln_fix with mean = 0,
            // alpha = 1, s = 0 is exactly acc + bias.  The loads of piece p + 1 are in flight while piece p is computed
            // (piece by piece under `if (bias && col < N)` every piece stalled on its own broadcast loads: a third of the
            // epilogue warps' busy time in the stall samples).
            const float* bsrc = bias ? bias : g_zero_f32;
            const float* csrc = ln ? ep.col_s : g_zero_f32;
            const int bstep = bias ? 1 : 0, cstep = ln ? 1 : 0;
            auto col_of = [&](int p) { const int c = col0 + 8 * p; return c < N ? c : N - 8; };
            float4 nb0, nb1, nc0, nc1;
            {
              const int c = col_of(0);
              nb0 = __ldg(reinterpret_cast<const float4*>(bsrc + bstep * c)); nb1 = __ldg(reinterpret_cast<const float4*>(bsrc + bstep * c + 4));
              nc0 = __ldg(reinterpret_cast<const float4*>(csrc + cstep * c)); nc1 = __ldg(reinterpret_cast<const float4*>(csrc + cstep * c + 4));
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) {  // piece p = columns col0 + 8p .. + 7
              const float4 b0 = nb0, b1 = nb1, c0 = nc0, c1 = nc1;
              if (p + 1 < 8) {
                const int c = col_of(p + 1);
                nb0 = __ldg(reinterpret_cast<const float4*>(bsrc + bstep * c)); nb1 = __ldg(reinterpret_cast<const float4*>(bsrc + bstep * c + 4));
                nc0 = __ldg(reinterpret_cast<const float4*>(csrc + cstep * c)); nc1 = __ldg(reinterpret_cast<const float4*>(csrc + cstep * c + 4));
              }
              const uint32_t* src = (p < 4) ? &v0[8 * p] : &v1[8 * (p - 4)];
              float x0 = ln_fix(__uint_as_float(src[0]), b0.x, c0.x, ln_mu, ln_alpha);
              float x1 = ln_fix(__uint_as_float(src[1]), b0.y, c0.y, ln_mu, ln_alpha);
              float x2 = ln_fix(__uint_as_float(src[2]), b0.z, c0.z, ln_mu, ln_alpha);
              float x3 = ln_fix(__uint_as_float(src[3]), b0.w, c0.w, ln_mu, ln_alpha);
              float x4 = ln_fix(__uint_as_float(src[4]), b1.x, c1.x, ln_mu, ln_alpha);
              float x5 = ln_fix(__uint_as_float(src[5]), b1.y, c1.y, ln_mu, ln_alpha);
              float x6 = ln_fix(__uint_as_float(src[6]), b1.z, c1.z, ln_mu, ln_alpha);
              float x7 = ln_fix(__uint_as_float(src[7]), b1.w, c1.w, ln_mu, ln_alpha);
              if (ep.act == CLM_EPI_QUICKGELU) {
                x0 = quick_gelu(x0); x1 = quick_gelu(x1); x2 = quick_gelu(x2); x3 = quick_gelu(x3);
                x4 = quick_gelu(x4); x5 = quick_gelu(x5); x6 = quick_gelu(x6); x7 = quick_gelu(x7);
              }
              pk[4 * p + 0] = pack_bf16x2(x0, x1); pk[4 * p + 1] = pack_bf16x2(x2, x3);
              pk[4 * p + 2] = pack_bf16x2(x4, x5); pk[4 * p + 3] = pack_bf16x2(x6, x7);
            }
          }
#else
          if (live) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {  // piece p = columns col0 + 8p .. + 7
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (bias && col0 + 8 * p < N) {
                b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * p));
                b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 8 * p + 4));
              }
              const uint32_t* src = (p < 4) ? &v0[8 * p] : &v1[8 * (p - 4)];
              float x0, x1, x2, x3, x4, x5, x6, x7;
              if (ln) {
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                if (col0 + 8 * p < N) {
                  c0 = __ldg(reinterpret_cast<const float4*>(ep.col_s + col0 + 8 * p));
                  c1 = __ldg(reinterpret_cast<const float4*>(ep.col_s + col0 + 8 * p + 4));
                }
                x0 = ln_fix(__uint_as_float(src[0]), b0.x, c0.x, ln_mu, ln_alpha);
                x1 = ln_fix(__uint_as_float(src[1]), b0.y, c0.y, ln_mu, ln_alpha);
                x2 = ln_fix(__uint_as_float(src[2]), b0.z, c0.z, ln_mu, ln_alpha);
                x3 = ln_fix(__uint_as_float(src[3]), b0.w, c0.w, ln_mu, ln_alpha);
                x4 = ln_fix(__uint_as_float(src[4]), b1.x, c1.x, ln_mu, ln_alpha);
                x5 = ln_fix(__uint_as_float(src[5]), b1.y, c1.y, ln_mu, ln_alpha);
                x6 = ln_fix(__uint_as_float(src[6]), b1.z, c1.z, ln_mu, ln_alpha);
                x7 = ln_fix(__uint_as_float(src[7]), b1.w, c1.w, ln_mu, ln_alpha);
              } else {
                x0 = __uint_as_float(src[0]) + b0.x; x1 = __uint_as_float(src[1]) + b0.y;
                x2 = __uint_as_float(src[2]) + b0.z; x3 = __uint_as_float(src[3]) + b0.w;
                x4 = __uint_as_float(src[4]) + b1.x; x5 = __uint_as_float(src[5]) + b1.y;
                x6 = __uint_as_float(src[6]) + b1.z; x7 = __uint_as_float(src[7]) + b1.w;
              }
              if (ep.act == CLM_EPI_QUICKGELU) {
                x0 = quick_gelu(x0); x1 = quick_gelu(x1); x2 = quick_gelu(x2); x3 = quick_gelu(x3);
                x4 = quick_gelu(x4); x5 = quick_gelu(x5); x6 = quick_gelu(x6); x7 = quick_gelu(x7);
              }
              pk[4 * p + 0] = pack_bf16x2(x0, x1); pk[4 * p + 1] = pack_bf16x2(x2, x3);
              pk[4 * p + 2] = pack_bf16x2(x4, x5); pk[4 * p + 3] = pack_bf16x2(x6, x7);
            }
          }
#endif
        }
        if (live) {
          if (lane == 0) bulk_wait_read<0>();  // the previous slab has left the staging buffer
          __syncwarp();
#pragma unroll
          for (int p = 0; p < 8; ++p)
            st_shared_v4(my_row + ((static_cast<uint32_t>(p) ^ sw) << 4), pk[4 * p], pk[4 * p + 1],
                         pk[4 * p + 2], pk[4 * p + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (kEpi == kEpiReduceF32 || kEpi == kEpiReduceBf16) tma_reduce_add_2d(&map_out, stg, col0, m0);
            else tma_store_2d(&map_out, stg, col0, m0);
            bulk_commit();
          }
        }
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();  // every tile of this warp has reached global memory
  } else {
    // ================= epilogue warps (legacy per-thread path) =================
    // TMEM gives each lane one accumulator ROW (32 fp32 columns per tcgen05.ld).  Storing that
    // directly would touch 32 different cache lines per warp instruction, so each warp transposes
    // its 32x32 chunk through a private 4 KiB shared-memory buffer (16-byte pieces XOR-swizzled by
    // row: conflict-free both ways) and does bias / activation / residual / stores in the
    // COALESCED domain: lane l owns columns 4*(l%8)..+3 of rows (l/8) + 4*i, i = 0..7, so every
    // global instruction covers 4 full 128-byte row segments.
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;   // which half of the tile's columns this warp drains
    constexpr int kChunks = BN / 64;    // 32-column chunks per warp
    uint8_t* stage = staging + (warp - kEpiWarp0) * 4096;
    const int piece = lane & 7;         // 16-byte piece (4 fp32 columns) inside the 128-byte chunk row
    const int rsub = lane >> 3;         // row offset inside each group of 4 rows
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      const int tpos = ep.reverse ? (m_tiles * n_tiles - 1 - tile) : tile;
      const int m0 = (tpos / n_tiles) * TM + static_cast<int>(rank) * BM + q * 32;
      const int n0 = (tpos % n_tiles) * BN + half * (BN / 2);
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
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_cta(&tmem_empty[acc], 0);  // the leader's MMA thread waits on it
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    if (kCtas == 2) tmem_dealloc_2sm(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int BN, int kCtas, int kEpi>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma2,
                const CUtensorMap& mb2, const CUtensorMap& mo, int M, int N, int kb_main, int kb_ext,
                const EpiParams& ep, cudaStream_t stream) {
  using C = Cfg<BN, kCtas>;
  static bool attr_set = false;
  if (!attr_set) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN, kCtas, kEpi>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ((M + BM * kCtas - 1) / (BM * kCtas)) * ((N + BN - 1) / BN) * ep.ksplit;  // work units
  const int max_workers = clm_num_sms() / kCtas;
  const int workers = tiles < max_workers ? tiles : max_workers;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * kCtas);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = clm_pdl_enabled() ? 2 : 1;
  CLM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, kCtas, kEpi>, ma, mb, ma2, mb2, mo, M, N, kb_main,
                                    kb_ext, ep));
  return CLM_OK;
}

// which kernel instantiation the last clm_gemm_epi call of this thread selected (tests assert it)
thread_local int g_last_variant = 0;

template <int kEpi>
int launch_gemm_bn(bool pair, int BN, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma2,
                   const CUtensorMap& mb2, const CUtensorMap& mo, int M, int N, int kb_main, int kb_ext,
                   const EpiParams& ep, cudaStream_t stream) {
  g_last_variant = (pair ? 256 : BN) * 100 + (pair ? 2 : 1) * 10 + kEpi;
  if (pair) return launch_gemm<256, 2, kEpi>(ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
  switch (BN) {
    case 256: return launch_gemm<256, 1, kEpi>(ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
    case 128: return launch_gemm<128, 1, kEpi>(ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
    default: return launch_gemm<64, 1, kEpi>(ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
  }
}

}  // namespace

// Internal entry used by the tower code as well (same translation unit boundary as the C-ABI).
int clm_gemm_launch(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                    const void* A2, int lda2, const void* W2, int ldw2, int K2, void* out, int ldo,
                    int out_dtype, const float* bias, const float* residual, int ldr, int epilogue,
                    cudaStream_t stream, const float* row_stats, const float* col_sums, int ln_mode) {
  CLM_REQUIRE(A && W && out, "clm_gemm_epi: null operand");
  if (row_stats) {
    CLM_REQUIRE(col_sums && (ln_mode == 1 || ln_mode == 2), "clm_gemm_ln_epi: col_sums and ln_mode 1 / 2 are required");
    CLM_REQUIRE(!residual, "clm_gemm_ln_epi: no residual with the folded LayerNorm epilogue");
    CLM_REQUIRE((reinterpret_cast<uintptr_t>(row_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(col_sums) & 15) == 0,
                "clm_gemm_ln_epi: row_stats must be 8-byte and col_sums 16-byte aligned");
  }
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
  static int force_1cta = -1;  // CLM_GEMM_1CTA=1 forces the single-CTA kernel (A/B measurements)
  if (force_1cta < 0) {
    const char* e = getenv("CLM_GEMM_1CTA");
    force_1cta = (e && e[0] == '1') ? 1 : 0;
  }
  // CTA-pair kernel for the big GEMMs; single-CTA kernel for narrow / ragged N and tiny M
  bool pair = !force_1cta && (N % 256 == 0) && (M > 128);
  int BN = pair ? 256
                : ((N >= 256 && N % 256 == 0) ? 256
                   : ((N >= 128 && N % 128 == 0) ? 128 : (N > 128 ? 256 : (N > 64 ? 128 : 64))));
  // Small M (the reference's one-item-at-a-time calls: 197 or 77 rows): a grid of 256-wide tiles would leave
  // most SMs idle and every tile walks the whole K loop, so take the narrowest tile that divides N until at
  // least half of the SMs have one.  Launch-bound regime: what counts is the latency of one tile.
  {
    const long long sms = clm_num_sms();
    const long long tiles_now = pair ? static_cast<long long>((M + 255) / 256) * (N / 256)
                                     : static_cast<long long>((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    if (tiles_now * (pair ? 2 : 1) * 2 <= sms) {
      for (int bn : {256, 128, 64}) {
        if (N % bn != 0) continue;
        pair = false;
        BN = bn;
        if (static_cast<long long>((M + BM - 1) / BM) * (N / bn) * 2 >= sms) break;
      }
    }
  }
  const int b_box_rows = pair ? BN / 2 : BN;
  CUtensorMap ma, mb, ma2, mb2;
  int rc;
  if ((rc = clm_make_tmap_bf16_2d(&ma, A, M, K, lda, BK, BM))) return rc;
  if ((rc = clm_make_tmap_bf16_2d(&mb, W, N, K, ldw, BK, b_box_rows))) return rc;
  if (has_ext) {
    if ((rc = clm_make_tmap_bf16_2d(&ma2, A2, M, K2, lda2, BK, BM))) return rc;
    if ((rc = clm_make_tmap_bf16_2d(&mb2, W2, N, K2, ldw2, BK, b_box_rows))) return rc;
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
  ep.act = epilogue & CLM_EPI_QUICKGELU;
  ep.row_stats = reinterpret_cast<const float2*>(row_stats);
  ep.col_s = col_sums;
  ep.ln_mode = ln_mode;
  ep.reverse = (epilogue & CLM_EPI_REVERSE) ? 1 : 0;
  // epilogue variant: TMA tile stores, or a TMA reduce-add for the in-place residual update; the
  // per-thread legacy path only for residual != out / bf16 out + residual (CLM_GEMM_EPI=legacy forces it)
  static int force_legacy = -1;
  if (force_legacy < 0) {
    const char* e = getenv("CLM_GEMM_EPI");
    force_legacy = (e && e[0] == 'l') ? 1 : 0;
  }
  int epi = ep.out_f32 ? kEpiStoreF32 : kEpiStoreBf16;
  if (residual) {
    // residual == out (same address, same pitch) is the in-place update of the residual stream, in the stream's own
    // type: fp32, or bf16 when out_dtype says so (the pointer is then a bf16 buffer despite its declared type)
    const bool in_place = residual == static_cast<const float*>(out) && ldr == ldo;
    epi = in_place ? (ep.out_f32 ? kEpiReduceF32 : kEpiReduceBf16) : kEpiLegacy;
  }
  if (force_legacy && epi != kEpiReduceBf16 && !row_stats) epi = kEpiLegacy;  // (the legacy path reads an fp32 residual)
  CUtensorMap mo = ma;
  if (epi != kEpiLegacy) {
    if ((rc = clm_make_tmap_2d(&mo, out, M, N, ldo, ep.out_f32 ? 4 : 2, ep.out_f32 ? 32 : 64, 32))) return rc;
    ep.residual = nullptr;  // the reduce-add reads the residual inside the L2
  }
  const int kb_main = (K + BK - 1) / BK;
  const int kb_ext = has_ext ? (K2 + BK - 1) / BK : 0;
  // Split-K for deep, narrow products with the in-place reduce-add epilogue (the LoRA weight gradients of the
  // training step: out [features, 64] += dy^T t over tens of thousands of token rows -- 4 to 18 tiles on 148 SMs):
  // cut K so that every SM has a work unit, but keep at least 8 k-blocks per unit.
  ep.ksplit = 1;
  ep.kb_per = kb_main + kb_ext;
  if ((epi == kEpiReduceF32) && (epilogue & CLM_EPI_SPLIT_K)) {
    const long long sms = clm_num_sms();
    const long long tiles = pair ? static_cast<long long>((M + 255) / 256) * (N / 256)
                                 : static_cast<long long>((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const long long ctas = tiles * (pair ? 2 : 1);
    const int kb_total = kb_main + kb_ext;
    if (ctas * 2 <= sms && kb_total >= 32) {
      long long want = sms / ctas;
      if (want > kb_total / 8) want = kb_total / 8;
      if (want > 1) {
        ep.kb_per = static_cast<int>((kb_total + want - 1) / want);
        ep.ksplit = (kb_total + ep.kb_per - 1) / ep.kb_per;
      }
    }
  }
  const double flops = 2.0 * M * N * (static_cast<double>(K) + (has_ext ? K2 : 0));
  const double bytes = 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) +
                       static_cast<double>(M) * N * (ep.out_f32 ? 4 : 2) +
                       (residual ? (epi == kEpiReduceBf16 ? 2.0 : 4.0) * M * N : 0.0);
  ProfScope prof(CLM_K_GEMM, flops, bytes, stream);
  switch (epi) {
    case kEpiStoreBf16:
      return launch_gemm_bn<kEpiStoreBf16>(pair, BN, ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
    case kEpiStoreF32:
      return launch_gemm_bn<kEpiStoreF32>(pair, BN, ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
    case kEpiReduceF32:
      return launch_gemm_bn<kEpiReduceF32>(pair, BN, ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
    case kEpiReduceBf16:
      return launch_gemm_bn<kEpiReduceBf16>(pair, BN, ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
    default:
      return launch_gemm_bn<kEpiLegacy>(pair, BN, ma, mb, ma2, mb2, mo, M, N, kb_main, kb_ext, ep, stream);
  }
}

extern "C" int clm_last_gemm_variant(void) { return g_last_variant; }

extern "C" int clm_gemm_epi(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                            const void* A2, int lda2, const void* W2, int ldw2, int K2, void* out,
                            int ldo, int out_dtype, const float* bias, const float* residual,
                            int ldr, int epilogue, void* stream) {
  return clm_gemm_launch(A, lda, W, ldw, M, N, K, A2, lda2, W2, ldw2, K2, out, ldo, out_dtype, bias,
                         residual, ldr, epilogue, static_cast<cudaStream_t>(stream), nullptr, nullptr, 0);
}

extern "C" int clm_gemm_ln_epi(const void* H, int ldh, const void* Wg, int ldw, int M, int N, int K,
                               const void* A2, int lda2, const void* W2, int ldw2, int K2, void* out,
                               int ldo, int out_dtype, const float* bias, const float* row_stats,
                               const float* col_sums, int ln_mode, int epilogue, void* stream) {
  CLM_REQUIRE(row_stats && col_sums, "clm_gemm_ln_epi: row_stats / col_sums are required");
  return clm_gemm_launch(H, ldh, Wg, ldw, M, N, K, A2, lda2, W2, ldw2, K2, out, ldo, out_dtype, bias,
                         nullptr, 0, epilogue, static_cast<cudaStream_t>(stream), row_stats, col_sums, ln_mode);
}
