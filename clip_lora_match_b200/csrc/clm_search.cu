// clm_search.cu — brute-force cosine top-k over the embedding index.
//
// Pass 1 (search_kernel): scores = Q · E^T on tcgen05 from the bf16 shadow of the index.
//   The mainloop is the same TMA -> smem ring -> tcgen05.mma -> TMEM pipeline as clm_gemm.cu
//   (128 queries x 256 index rows per accumulator, double buffered).  The epilogue never
//   writes a score: each epilogue thread owns one query row, streams its 256 scores per tile
//   out of TMEM and keeps the kc best (score, row) pairs of its work unit in shared memory
//   (replace-the-minimum list; a running threshold rejects ~all scores with one FMNMX tree
//   and one compare per 32).  A work unit is (query tile, index split); units are ordered
//   split-major so the CTAs that run together sweep the SAME index rows and share them in L2:
//   the index crosses HBM about once per query batch (src/embedding/search.py:96,99 fused).
//
// Pass 2 (merge_kernel): per query, select the kc best of splits*kc candidates, re-score them
//   exactly in fp32 against the fp32 master rows, order by (score desc, id asc), emit top k.
//   bf16 scores only ever *nominate*; every returned score is an fp32 dot product, which is
//   what makes the ids match the fp32 reference.
#include <math.h>
#include <stdlib.h>

#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int BM = 128;   // queries per tile
constexpr int BN = 256;   // index rows per tile
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kAccStages = 2;
constexpr int kABytes = BM * BK * 2;
constexpr int kMaxStages = 6;
constexpr int kTmemCols = kAccStages * BN;
// Warp roles: the four epilogue warps are warps 0..3 (TMEM lane quarter = warp id), the two control warps sit
// ABOVE them (the sub-partition arbiter prefers the highest eligible warp id, so the single-thread TMA / MMA
// issue loops are not queued behind an epilogue warp that shares their sub-partition).  CLM_SEARCH_CTL_HI=0
// restores the old order (control warps 0, 1) for A/B builds.
#ifndef CLM_SEARCH_CTL_HI
#define CLM_SEARCH_CTL_HI 1
#endif
constexpr int kTmaWarp = CLM_SEARCH_CTL_HI ? 4 : 0;
constexpr int kMmaWarp = CLM_SEARCH_CTL_HI ? 5 : 1;
// per-CTA stage: its 128 query rows + its share of the 256-row index tile (all of it, or half in a CTA pair)
template <int kCtas>
struct SCfg {
  static constexpr int kBBytes = (BN / kCtas) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
};

struct SearchParams {
  int nq, n, kblocks, q_tiles, splits, tiles_n, kc, stages;  // q_tiles: tiles of BM * kCtas queries
  int a_bytes;  // bytes of the query tile actually loaded per k-block (64-row box when nq <= 64)
  int kb;         // a list holding kb entries proves a lower bound of the query's kb-th best score (kb = k)
  float* thr_io;  // per-query running lower bound of the kb-th best score, shared by all units (or null)
  float margin;   // scores above (shared bound - margin) are kept: see clm_search_topk in include/clm_b200.h
  // Per-query score histogram shared by ALL work units (or null): every kept candidate with score >= base[q]
  // counts in bin (score - base[q]) / kHistStep (the top bin is open ended).  kb candidates counted at or above a
  // bin edge prove that the query's kb-th best score is at least that edge -- over all rows any unit has seen,
  // not only the rows of one unit -- which is what tightens the shared bound early on small shards.
  const float* hist_base;
  unsigned int* hist;
  float* cand_score;
  int32_t* cand_id;
};

// ---- cta_group::2 helpers (same protocol as clm_gemm.cu) --------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                     int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint)
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

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// State of one thread's candidate list (the list itself is in shared memory, element j at sc[j * BM]).
constexpr int kHistBins = 32;
constexpr float kHistStep = 1.0f / 256.0f;

struct ListState {
  int cnt;
  unsigned int* hist;  // this query's histogram (or null)
  float hbase;
  float thr;        // a score must exceed this to enter the list: max(own, gbound - margin)
  float own;        // minimum of this list once it is full (-inf before)
  float runmin;     // minimum of the entries appended so far (while the list is filling)
  float gbound;     // the shared per-query bound as last seen / published by this thread
  float margin;
};

__device__ __forceinline__ void publish_bound(ListState& st, float m, float* gthr) {
  if (gthr && m > st.gbound) {  // float max through the integer atomics
    st.gbound = m;
    if (m >= 0.f) atomicMax(reinterpret_cast<int*>(gthr), __float_as_int(m));
    else atomicMin(reinterpret_cast<unsigned int*>(gthr), __float_as_uint(m));
  }
}

// restore the min-heap property below node j (children 2j+1, 2j+2; element j at [j * BM])
__device__ __forceinline__ void sift_down(float* my_sc, int32_t* my_id, int kc, int j, float s, int id) {
  for (;;) {
    int c = 2 * j + 1;
    if (c >= kc) break;
    float cs = my_sc[c * BM];
    if (c + 1 < kc) {
      const float rs = my_sc[(c + 1) * BM];
      if (rs < cs) { cs = rs; ++c; }
    }
    if (cs >= s) break;
    my_sc[j * BM] = cs;
    my_id[j * BM] = my_id[c * BM];
    j = c;
  }
  my_sc[j * BM] = s;
  my_id[j * BM] = id;
}

// Insert (s, id) into a thread's list of the kc best.  While the list fills, entries are appended; once kb of
// them exist their minimum is a lower bound of the query's kb-th best score and is published to the shared
// bound.  When the list becomes full it is turned into a min-heap (root = the list minimum = the thread's own
// threshold); from then on an insertion replaces the root and sifts down: O(log kc) shared-memory accesses
// instead of a rescan of all kc slots.  Deliberately NOT inlined: the scan loop tests 32 scores per chunk and an
// inlined copy per score made the epilogue 32x larger than the instruction cache likes (ncu: ~45 % of the warp
// samples of the Q = 4096 scan sat on instruction fetch after these blocks).
__device__ __noinline__ ListState list_insert(ListState st, float s, int id, float* my_sc, int32_t* my_id, int kc,
                                              int kb, float* gthr) {
  if (st.cnt < kc) {
    my_sc[st.cnt * BM] = s;
    my_id[st.cnt * BM] = id;
    ++st.cnt;
    st.runmin = fminf(st.runmin, s);
    if (st.cnt == kb) publish_bound(st, st.runmin, gthr);
    if (st.cnt == kc) {
      for (int j = kc / 2 - 1; j >= 0; --j) sift_down(my_sc, my_id, kc, j, my_sc[j * BM], my_id[j * BM]);
      st.own = my_sc[0];
      if (kc >= kb) publish_bound(st, st.own, gthr);
    }
  } else {  // s > own (the caller's threshold): replace the minimum
    sift_down(my_sc, my_id, kc, 0, s, id);
    st.own = my_sc[0];
    if (kc >= kb) publish_bound(st, st.own, gthr);
  }
  if (st.hist && s >= st.hbase) {
    int b = static_cast<int>((s - st.hbase) * (1.0f / kHistStep));
    atomicAdd(st.hist + (b < kHistBins - 1 ? b : kHistBins - 1), 1u);
  }
  st.thr = fmaxf(st.own, st.gbound - st.margin);
  return st;
}

// kCtas = 2: a CTA pair (cluster of 2, cta_group::2) scores 256 queries x 256 index rows per
// accumulator; each CTA loads its own 128 queries and HALF of the index tile, so the operand bytes
// pulled from L2 per FLOP drop by a third (the Q >= 256 regime is L2-to-SM-bandwidth bound).
template <int kCtas>
__global__ void __launch_bounds__(kThreads, 1)
search_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_e,
              SearchParams p) {
  constexpr int kStageBytes = SCfg<kCtas>::kStageBytes;
  constexpr int kBBytes = SCfg<kCtas>::kBBytes;
  constexpr int TM = BM * kCtas;  // queries per work unit
  const uint32_t rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const int worker = (kCtas == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_workers = (kCtas == 2) ? (gridDim.x >> 1) : gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* list_sc = reinterpret_cast<float*>(smem + p.stages * kStageBytes);
  int32_t* list_id = reinterpret_cast<int32_t*>(list_sc + p.kc * BM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(list_id + p.kc * BM);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform (see clm_gemm.cu)
  const int lane = threadIdx.x & 31;
  const int num_units = p.q_tiles * p.splits;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_e);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4 * kCtas);  // one arrive per epilogue warp of the pair
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (kCtas == 2) {
      tmem_alloc_2sm(tmem_slot, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == kTmaWarp) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = worker; u < num_units; u += num_workers) {
        const int split = u / p.q_tiles;
        const int q0 = (u % p.q_tiles) * TM + static_cast<int>(rank) * BM;
        const int t0 = static_cast<int>(static_cast<long long>(split) * p.tiles_n / p.splits);
        const int t1 = static_cast<int>(static_cast<long long>(split + 1) * p.tiles_n / p.splits);
        for (int t = t0; t < t1; ++t) {
          const int e0 = t * BN + static_cast<int>(rank) * (BN / kCtas);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kStageBytes;
            if (kCtas == 2) {
              // the leader's barrier collects the bytes of BOTH CTAs
              if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * kStageBytes);
              tma_load_2d_2sm_hint(sa, &map_q, &full[stage], kb * BK, q0, kEvictLast);
              tma_load_2d_2sm_hint(sa + kABytes, &map_e, &full[stage], kb * BK, e0, kEvictFirst);
            } else {
              mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(p.a_bytes + kBBytes));
              tma_load_2d_hint(sa, &map_q, &full[stage], kb * BK, q0, kEvictLast);
              tma_load_2d_hint(sa + kABytes, &map_e, &full[stage], kb * BK, e0, kEvictFirst);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer (leader CTA only) =================
    constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = worker; rank == 0 && u < num_units; u += num_workers) {
      const int split = u / p.q_tiles;
      const int t0 = static_cast<int>(static_cast<long long>(split) * p.tiles_n / p.splits);
      const int t1 = static_cast<int>(static_cast<long long>(split + 1) * p.tiles_n / p.splits);
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {  // uniform operands, one issuing lane: four UTCHMMA straight from uniform registers
            const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
            const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = umma_desc_sw128(a_addr + k * 32, 1024);
              const uint64_t db = umma_desc_sw128(b_addr + k * 32, 1024);
              if (kCtas == 2) umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (kCtas == 2) {
              umma_commit_2sm(&empty[stage]);                             // frees the slot in both CTAs
              if (kb == p.kblocks - 1) umma_commit_2sm(&tmem_full[acc]);  // scores ready (both CTAs)
            } else {
              umma_commit(&empty[stage]);
              if (kb == p.kblocks - 1) umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue: streaming top-kc per query row =================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    float* my_sc = list_sc + r;   // element j at my_sc[j * BM]  (lane-contiguous: no conflicts)
    int32_t* my_id = list_id + r;
    const int kc = p.kc;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = worker; u < num_units; u += num_workers) {
      const int split = u / p.q_tiles;
      const int q0 = (u % p.q_tiles) * TM + static_cast<int>(rank) * BM;
      const int t0 = static_cast<int>(static_cast<long long>(split) * p.tiles_n / p.splits);
      const int t1 = static_cast<int>(static_cast<long long>(split + 1) * p.tiles_n / p.splits);
      const bool warp_live = q0 + q * 32 < p.nq;  // rows past the last query: nothing to scan
      ListState ls;
      ls.cnt = 0;
      ls.runmin = INFINITY;
      // gbound is a lower bound of this query's kc-th best score over the WHOLE index: the kc-th best of
      // any subset of rows qualifies, so every unit publishes its own list minimum (atomic max in
      // global memory) and re-reads the shared bound once per tile.  Units that start after the
      // first wave inherit a nearly final bound and almost never enter the insert path.  Scores within
      // `margin` below the bound are still kept (the bf16 score of a true top-k row may sit that far below
      // the bf16 score of the k-th best row; see include/clm_b200.h).
      float* gthr = (p.thr_io != nullptr && q0 + r < p.nq) ? p.thr_io + q0 + r : nullptr;
      ls.thr = -INFINITY;
      ls.own = -INFINITY;
      ls.gbound = -INFINITY;
      ls.margin = p.margin;
      ls.hist = nullptr;
      ls.hbase = 0.f;
      if (gthr && p.hist) {
        ls.hbase = p.hist_base[q0 + r];
        if (ls.hbase > -INFINITY) ls.hist = p.hist + static_cast<size_t>(q0 + r) * kHistBins;
      }
      for (int t = t0; t < t1; ++t) {
        if (gthr) {
          ls.gbound = fmaxf(ls.gbound, __ldcg(gthr));
          if (ls.hist) {
            // highest bin edge with at least kb candidates counted at or above it (by any unit, so far)
            const uint4* h4 = reinterpret_cast<const uint4*>(ls.hist);
            unsigned int cum = 0;
            int edge = -1;
#pragma unroll
            for (int j = kHistBins / 4 - 1; j >= 0; --j) {
              const uint4 h = __ldcg(h4 + j);
              cum += h.w; if (edge < 0 && cum >= static_cast<unsigned int>(p.kb)) edge = 4 * j + 3;
              cum += h.z; if (edge < 0 && cum >= static_cast<unsigned int>(p.kb)) edge = 4 * j + 2;
              cum += h.y; if (edge < 0 && cum >= static_cast<unsigned int>(p.kb)) edge = 4 * j + 1;
              cum += h.x; if (edge < 0 && cum >= static_cast<unsigned int>(p.kb)) edge = 4 * j;
            }
            if (edge >= 0) publish_bound(ls, ls.hbase + edge * kHistStep, gthr);
          }
          ls.thr = fmaxf(ls.own, ls.gbound - ls.margin);
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * BN);
        const int row0 = t * BN;
        const bool ragged = row0 + BN > p.n;
#pragma unroll 1
        for (int c = 0; warp_live && c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
          const int base = row0 + c * 32;
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (base + i >= p.n) v[i] = 0xff800000u;  // -inf: rows past the end never win
          }
          // Maximum of each group of 8 consecutive scores (three-input FMNMX3, four independent chains), then of
          // the chunk.  Almost every chunk fails the threshold test as a whole; when a warp does enter the slow
          // path (with a 2^-7 margin at k = 50 about 40 % of the chunks have a hit in SOME lane) it rescans only
          // the groups of 8 whose maximum passed, not all 32 scores.
          float gm[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float m = fmax3(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]), __uint_as_float(v[8 * j + 2]));
            m = fmax3(m, __uint_as_float(v[8 * j + 3]), __uint_as_float(v[8 * j + 4]));
            m = fmax3(m, __uint_as_float(v[8 * j + 5]), __uint_as_float(v[8 * j + 6]));
            gm[j] = fmaxf(m, __uint_as_float(v[8 * j + 7]));
          }
          const float cmax = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
          if (cmax > ls.thr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (gm[j] > ls.thr) {
#pragma unroll
                for (int i = 8 * j; i < 8 * j + 8; ++i) {
                  const float s = __uint_as_float(v[i]);
                  if (s > ls.thr) ls = list_insert(ls, s, base + i, my_sc, my_id, kc, p.kb, gthr);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kCtas == 2) mbar_arrive_cta(&tmem_empty[acc], 0);  // the leader's MMA thread waits on it
          else mbar_arrive(&tmem_empty[acc]);
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
      // flush this unit's list: [query][split][kc], padded with (-inf, -1)
      const int qrow = q0 + r;
      if (qrow < p.nq) {
        const size_t o = (static_cast<size_t>(qrow) * p.splits + split) * kc;
        for (int j = 0; j < kc; ++j) {
          const bool ok = j < ls.cnt;
          p.cand_score[o + j] = ok ? my_sc[j * BM] : -INFINITY;
          p.cand_id[o + j] = ok ? my_id[j * BM] : -1;
        }
      }
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    if (kCtas == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int kCtas>
int launch_search(const CUtensorMap& mq, const CUtensorMap& me, const SearchParams& p, int smem, int workers,
                  cudaStream_t stream) {
  CLM_CUDA_CHECK(cudaFuncSetAttribute(search_kernel<kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * kCtas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CLM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, search_kernel<kCtas>, mq, me, p));
  return CLM_OK;
}

// -----------------------------------------------------------------------------------------
// selection helpers: k-th largest by radix select, bitonic sort of (score, id) pairs
// -----------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kMaxSel = 2048;  // candidates a query may carry into the exact re-score / final sort

// order-preserving map float -> uint32 (ascending); -inf is the smallest key that occurs
__device__ __forceinline__ uint32_t key_of(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_of(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th largest (kth >= 1) of v[0..n) -- v in shared OR global memory -- by an 8-bit-per-pass radix select over
// the order-preserving keys: 4 passes, each a block-wide histogram of the values that still match the prefix.
// hist: 256 ints of shared memory; ctl: 2 uint32 of shared memory.  Every thread of the block must call it;
// returns the same value to all.  n >= kth is required.
__device__ float block_kth_largest(const float* v, int n, int kth, int* hist, uint32_t* ctl) {
  const int tid = threadIdx.x, nt = blockDim.x;
  uint32_t prefix = 0, mask = 0;
  int remaining = kth;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      const uint32_t k = key_of(v[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns bins [8l, 8l+8); suffix sums from the top bin down locate the bin holding the kth largest
      int c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[tid * 8 + j]; sum += c[j]; }
      int incl = sum;  // inclusive suffix sum over lanes >= this one (Hillis-Steele with shfl_down)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_down_sync(0xffffffffu, incl, o);
        if (tid + o < 32) incl += x;
      }
      const int above = incl - sum;  // values in bins owned by higher lanes
      if (above < remaining && above + sum >= remaining) {
        int need = remaining - above;
        int digit = 0;
        for (int j = 7; j >= 0; --j) {
          if (c[j] >= need) { digit = j; break; }
          need -= c[j];
        }
        ctl[0] = prefix | (static_cast<uint32_t>(tid * 8 + digit) << shift);
        ctl[1] = static_cast<uint32_t>(need);
      }
    }
    __syncthreads();
    prefix = ctl[0];
    remaining = static_cast<int>(ctl[1]);
    mask |= 255u << shift;
    __syncthreads();
  }
  return float_of(prefix);
}

// (score, id) ordering of torch.topk(largest=True) with a deterministic tie rule: lower id wins
__device__ __forceinline__ bool better(float sa, long long ia, float sb, long long ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// in-place bitonic sort of n2 (a power of two) pairs in shared memory, best first
__device__ void block_sort_pairs(float* sc, long long* id, int n2) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < (n2 >> 1); i += nt) {
        const int lo = ((i / stride) * (stride << 1)) + (i % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);  // this sub-sequence is sorted best-first
        const float sa = sc[lo], sb = sc[hi];
        const long long ia = id[lo], ib = id[hi];
        const bool swap = desc ? better(sb, ib, sa, ia) : better(sa, ia, sb, ib);
        if (swap) { sc[lo] = sb; sc[hi] = sa; id[lo] = ib; id[hi] = ia; }
      }
    }
  }
  __syncthreads();
}

// -----------------------------------------------------------------------------------------
// merge / re-score
// -----------------------------------------------------------------------------------------
// One block per query.  in: [lists][kc] candidates (id < 0 = empty slot).
//   t   = k-th largest candidate score (all candidates if there are fewer than k)
//   cut = t - margin;  selected = every candidate with score >= cut  (at most kMaxSel)
//   optional exact fp32 re-score of the selected rows, sort by (score desc, id asc), emit the top k.
// overflow[q] (optional) is set when the selection may be incomplete: a list that is full (kc entries) whose
// minimum is >= cut may have dropped rows the selection should contain, or more than kMaxSel were selected.
template <typename IdT>
__global__ void __launch_bounds__(kSelThreads)
merge_kernel(const float* __restrict__ in_score, const IdT* __restrict__ in_id, int lists, int kc,
             long long q_stride, long long l_stride_s, long long l_stride_i, float margin,
             const float* __restrict__ q_f32, const float* __restrict__ index_f32, int dim, int k,
             long long id_offset, float* __restrict__ out_score, long long* __restrict__ out_id,
             int* __restrict__ overflow) {
  extern __shared__ __align__(16) uint8_t msm[];
  const int total = lists * kc;
  long long* sel_id = reinterpret_cast<long long*>(msm);                 // [kMaxSel]
  float* sel_sc = reinterpret_cast<float*>(sel_id + kMaxSel);            // [kMaxSel]
  float* sc = sel_sc + kMaxSel;                                          // [total]
  __shared__ int hist[256];
  __shared__ uint32_t ctl[2];
  __shared__ int n_valid, n_sel, ovf;

  const int qi = blockIdx.x;
  const int tid = threadIdx.x;
  // candidate (list l, slot j) of this query sits at [qi * q_stride + l * l_stride + j]; scores and ids may
  // have different list strides (the gathered per-rank chunks hold 8-byte ids and 4-byte scores)
  const float* qs = in_score + static_cast<size_t>(qi) * q_stride;
  const IdT* qid = in_id + static_cast<size_t>(qi) * q_stride;
  auto at_s = [&](int i) { return static_cast<size_t>(i / kc) * l_stride_s + (i % kc); };
  auto at_i = [&](int i) { return static_cast<size_t>(i / kc) * l_stride_i + (i % kc); };
  if (tid == 0) { n_valid = 0; n_sel = 0; ovf = 0; }
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < total; i += kSelThreads) {
    const bool ok = qid[at_i(i)] >= 0;
    sc[i] = ok ? qs[at_s(i)] : -INFINITY;
    mine += ok ? 1 : 0;
  }
  mine = static_cast<int>(warp_sum(static_cast<float>(mine)));  // exact: counts < 2^24
  if ((tid & 31) == 0 && mine) atomicAdd(&n_valid, mine);
  __syncthreads();
  const int nv = n_valid;
  float cut = -INFINITY;
  if (nv > k) cut = block_kth_largest(sc, total, k, hist, ctl) - margin;  // uniform branch

  // a full list whose minimum is still inside the cut may have dropped rows that belong to the selection
  if (overflow != nullptr) {
    for (int l = (tid >> 5); l < lists; l += kSelThreads / 32) {
      float mn = INFINITY;
      int cnt = 0;
      for (int j = (tid & 31); j < kc; j += 32) {
        const float x = sc[l * kc + j];
        if (x != -INFINITY) { ++cnt; mn = fminf(mn, x); }
      }
      cnt = static_cast<int>(warp_sum(static_cast<float>(cnt)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if ((tid & 31) == 0 && cnt == kc && mn >= cut) ovf = 1;
    }
  }
  // selection (order irrelevant: everything selected is re-scored / sorted below)
  for (int i = tid; i < total; i += kSelThreads) {
    const float x = sc[i];
    if (x != -INFINITY && x >= cut) {
      const int slot = atomicAdd(&n_sel, 1);
      if (slot < kMaxSel) {
        sel_sc[slot] = x;
        sel_id[slot] = static_cast<long long>(qid[at_i(i)]);
      }
    }
  }
  __syncthreads();
  int nsel = n_sel;
  if (nsel > kMaxSel) {
    nsel = kMaxSel;
    if (tid == 0) ovf = 1;
  }

  // exact fp32 re-score: one warp per candidate
  if (index_f32 != nullptr) {
    const float4* q4 = reinterpret_cast<const float4*>(q_f32 + static_cast<size_t>(qi) * dim);
    const int n4 = dim >> 2;
    for (int cnd = (tid >> 5); cnd < nsel; cnd += kSelThreads / 32) {
      const float4* e4 =
          reinterpret_cast<const float4*>(index_f32 + static_cast<size_t>(sel_id[cnd]) * dim);
      float s = 0.f;
      for (int i = (tid & 31); i < n4; i += 32) {
        const float4 a = q4[i];
        const float4 e = __ldg(e4 + i);
        s = fmaf(a.x, e.x, s); s = fmaf(a.y, e.y, s); s = fmaf(a.z, e.z, s); s = fmaf(a.w, e.w, s);
      }
      s = warp_sum(s);
      if ((tid & 31) == 0) sel_sc[cnd] = s;
    }
  }
  int n2 = 1;
  while (n2 < nsel) n2 <<= 1;
  for (int i = nsel + tid; i < n2; i += kSelThreads) { sel_sc[i] = -INFINITY; sel_id[i] = 0x7fffffffffffffffLL; }
  block_sort_pairs(sel_sc, sel_id, n2);  // starts with a __syncthreads()
  for (int j = tid; j < k; j += kSelThreads) {
    const bool ok = j < nsel;
    out_score[static_cast<size_t>(qi) * k + j] = ok ? sel_sc[j] : -INFINITY;
    out_id[static_cast<size_t>(qi) * k + j] = ok ? sel_id[j] + id_offset : -1;  // fewer than k candidates: pad
  }
  if (overflow != nullptr && tid == 0) overflow[qi] = ovf;
}

template <typename IdT>
int launch_merge(const float* in_score, const IdT* in_id, int nq, int lists, int kc, long long q_stride,
                 long long l_stride_s, long long l_stride_i, float margin, const float* q_f32, const float* index_f32, int dim, int k, long long id_offset,
                 float* out_score, long long* out_id, int* overflow, cudaStream_t s) {
  const int total = lists * kc;
  const size_t smem = static_cast<size_t>(kMaxSel) * 12 + static_cast<size_t>(total) * 4;
  CLM_REQUIRE(smem <= 200 * 1024, "topk merge: %d candidates per query exceed shared memory", total);
  if (smem > 48 * 1024)
    CLM_CUDA_CHECK(cudaFuncSetAttribute(merge_kernel<IdT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
  ProfScope prof(CLM_K_MERGE, 0.0, 8.0 * nq * total + (index_f32 ? 4.0 * nq * k * dim : 0.0), s);
  merge_kernel<IdT><<<nq, kSelThreads, smem, s>>>(in_score, in_id, lists, kc, q_stride, l_stride_s, l_stride_i, margin, q_f32,
                                                  index_f32, dim, k, id_offset, out_score, out_id, overflow);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

// exact fp32 scores of ONE query against every row (the reference's batch-1 matmul,
// src/embedding/search.py:96 / similarity.py:32): one warp per row, 128-bit loads, pure streaming.
__global__ void __launch_bounds__(256)
gemv_kernel(const float* __restrict__ q, const float* __restrict__ e, int n, int dim,
            float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const float4* e4 = reinterpret_cast<const float4*>(e + static_cast<size_t>(row) * dim);
  float s = 0.f;
  for (int i = (threadIdx.x & 31); i < (dim >> 2); i += 32) {
    const float4 a = __ldg(q4 + i);
    const float4 b = e4[i];
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[row] = s;
}

// kth largest value of each row of a [rows, n] fp32 matrix (n <= 16384): one block per row, the row is staged
// in shared memory and radix-selected (four histogram passes instead of kth arg-max rounds).  Used to seed the scan's per-query threshold from the
// scores of a small sample of index rows.
__global__ void __launch_bounds__(kSelThreads)
kth_largest_kernel(const float* __restrict__ x, int n, int kth, float guard, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t ksm[];
  float* stage = reinterpret_cast<float*>(ksm);
  __shared__ int hist[256];
  __shared__ uint32_t ctl[2];
  const float* row = x + static_cast<size_t>(blockIdx.x) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) stage[i] = row[i];
  __syncthreads();
  const float v = block_kth_largest(stage, n, kth, hist, ctl);
  if (threadIdx.x == 0) out[blockIdx.x] = v - guard;
}

// A LOWER BOUND of the kth largest value of each row, cheap enough for a 16 K-row sample: the row is cut into
// kSeedGroups strided groups (element i belongs to group i mod kSeedGroups), each thread keeps the maxima of its
// groups while it streams the row once (coalesced), and the kth largest of the kSeedGroups group maxima is
// selected exactly.  The kth largest group maximum has kth distinct elements at or above it, so it never
// exceeds the row's kth largest value; for kth << kSeedGroups it is close to it (the top kth elements fall into
// ~kth distinct groups: kth = 50 of 1024 groups -> it is about the 51st largest element).  The exact radix
// select over 16 K clustered scores costs 0.52 ms for 4096 queries (thousands of same-bin shared-memory atomics),
// this one 0.06 ms.
constexpr int kSeedGroups = 1024;
__global__ void __launch_bounds__(kSelThreads)
kth_lower_bound_kernel(const float* __restrict__ x, int n, int kth, float guard, float* __restrict__ out) {
  __shared__ float gmax[kSeedGroups];
  __shared__ int hist[256];
  __shared__ uint32_t ctl[2];
  constexpr int kPer = kSeedGroups / kSelThreads;  // groups per thread: thread t owns groups t + kSelThreads * j
  const float* row = x + static_cast<size_t>(blockIdx.x) * n;
  float m[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) m[j] = -INFINITY;
  // element i = t + kSelThreads * c belongs to group i mod kSeedGroups = t + kSelThreads * (c mod kPer)
  for (int c0 = 0; c0 * kSelThreads < n; c0 += kPer) {
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = (c0 + j) * kSelThreads + static_cast<int>(threadIdx.x);
      if (i < n) m[j] = fmaxf(m[j], __ldg(row + i));
    }
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) gmax[threadIdx.x + kSelThreads * j] = m[j];
  __syncthreads();
  const float v = block_kth_largest(gmax, kSeedGroups, kth, hist, ctl);
  if (threadIdx.x == 0) out[blockIdx.x] = v - guard;
}

// Exact top-k of ONE row of n fp32 scores (the reference's torch.topk(sims, k), src/embedding/search.py:99),
// for the cases the fused scan hands back (an overflowed candidate list, k larger than the lists hold).
// One block: radix select of the k-th largest value tau, gather of everything above tau plus enough entries
// equal to tau (ties are interchangeable), bitonic sort.  k <= kMaxSel.
__global__ void __launch_bounds__(1024)
topk_row_kernel(const float* __restrict__ x, int n, int k, long long id_offset, float* __restrict__ out_score,
                long long* __restrict__ out_id) {
  extern __shared__ __align__(16) uint8_t tsm[];
  long long* sel_id = reinterpret_cast<long long*>(tsm);        // [n2]
  float* sel_sc = reinterpret_cast<float*>(sel_id + kMaxSel);   // [n2]
  __shared__ int hist[256];
  __shared__ uint32_t ctl[2];
  __shared__ int n_gt, n_eq;
  const int tid = threadIdx.x;
  if (tid == 0) { n_gt = 0; n_eq = 0; }
  const float tau = block_kth_largest(x, n, k, hist, ctl);  // syncs inside
  for (int i = tid; i < n; i += blockDim.x) {
    const float v = x[i];
    if (v > tau) {
      const int slot = atomicAdd(&n_gt, 1);  // fewer than k values exceed the k-th largest
      sel_sc[slot] = v;
      sel_id[slot] = i;
    }
  }
  __syncthreads();
  const int ngt = n_gt;
  for (int i = tid; i < n; i += blockDim.x) {
    if (x[i] == tau) {
      const int slot = ngt + atomicAdd(&n_eq, 1);
      if (slot < k) { sel_sc[slot] = tau; sel_id[slot] = i; }
    }
  }
  __syncthreads();
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  for (int i = k + tid; i < n2; i += blockDim.x) { sel_sc[i] = -INFINITY; sel_id[i] = 0x7fffffffffffffffLL; }
  block_sort_pairs(sel_sc, sel_id, n2);
  for (int j = tid; j < k; j += blockDim.x) {
    out_score[j] = sel_sc[j];
    out_id[j] = sel_id[j] + id_offset;
  }
}

int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

}  // namespace

// CTA pairs (256 queries per work unit) once there are more than 128 queries; the streaming regime
// (<= 128 queries, HBM-bound) keeps the single-CTA kernel.  CLM_SEARCH_1CTA=1 forces the latter.
static int search_ctas(int num_queries) {
  static int force_1cta = -1;
  if (force_1cta < 0) {
    const char* e = getenv("CLM_SEARCH_1CTA");
    force_1cta = (e && e[0] == '1') ? 1 : 0;
  }
  return (!force_1cta && num_queries > BM) ? 2 : 1;
}

extern "C" int clm_search_num_splits(int num_queries, int num_rows) {
  if (num_queries <= 0 || num_rows <= 0) return 1;
  const int ctas = search_ctas(num_queries);
  const int q_tiles = (num_queries + BM * ctas - 1) / (BM * ctas);
  const int tiles_n = (num_rows + BN - 1) / BN;
  const int workers = clm_num_sms() / ctas;
  int splits = workers / gcd_int(q_tiles, workers);  // q_tiles*splits is a multiple of the worker count
  if (splits > tiles_n) splits = tiles_n;
  if (splits < 1) splits = 1;
  return splits;
}

constexpr int kMaxKc = 64;       // capacity of one (query, split) candidate list of the scan
constexpr int kMaxMergeK = 1024;  // largest k the merge emits (half of kMaxSel: room for the margin's extras)

extern "C" int clm_search_topk(const void* q_bf16, const void* index_bf16, int nq, int n, int dim,
                               int kc, int kb, int splits, float* thr_io, const float* hist_base,
                               uint32_t* hist, float margin, float* cand_score, int32_t* cand_id, void* stream) {
  CLM_REQUIRE(q_bf16 && index_bf16 && cand_score && cand_id, "clm_search_topk: null argument");
  CLM_REQUIRE(nq > 0 && n > 0 && dim > 0 && dim % 8 == 0, "clm_search_topk: bad shape nq=%d n=%d dim=%d",
              nq, n, dim);
  CLM_REQUIRE(kc >= 1 && kc <= kMaxKc, "clm_search_topk: kc=%d must be in [1,%d]", kc, kMaxKc);
  CLM_REQUIRE(margin >= 0.f && kb >= 1, "clm_search_topk: margin must be >= 0 and kb >= 1");
  CLM_REQUIRE(hist == nullptr || (reinterpret_cast<uintptr_t>(hist) & 15) == 0, "clm_search_topk: hist not 16-B aligned");
  // kb > kc is allowed: a list of kc < kb entries never publishes its own minimum (it proves nothing about the
  // kb-th best), the bound then comes from the caller's seed and the shared histogram only
  const int ctas = search_ctas(nq);
  const int q_tiles = (nq + BM * ctas - 1) / (BM * ctas);
  const int tiles_n = (n + BN - 1) / BN;
  CLM_REQUIRE(splits >= 1 && splits <= tiles_n, "clm_search_topk: splits=%d must be in [1,%d]", splits,
              tiles_n);
  CUtensorMap mq, me;
  int rc;
  // streaming regime with at most 64 queries: only 64 query rows are (re)loaded per k-block; the other 64
  // rows of the A tile are never written and feed TMEM lanes that nobody reads.  CLM_SEARCH_QBOX=128 disables.
  static int qbox_full = -1;
  if (qbox_full < 0) {
    const char* e = getenv("CLM_SEARCH_QBOX");
    qbox_full = (e && e[0] == '1') ? 1 : 0;
  }
  const int q_box_rows = (ctas == 1 && nq <= 64 && !qbox_full) ? 64 : BM;
  if ((rc = clm_make_tmap_bf16_2d(&mq, q_bf16, nq, dim, dim, BK, q_box_rows))) return rc;
  if ((rc = clm_make_tmap_bf16_2d(&me, index_bf16, n, dim, dim, BK, BN / ctas))) return rc;
  SearchParams p;
  p.nq = nq;
  p.n = n;
  p.kblocks = (dim + BK - 1) / BK;
  p.q_tiles = q_tiles;
  p.splits = splits;
  p.tiles_n = tiles_n;
  p.kc = kc;
  p.kb = kb;
  p.a_bytes = q_box_rows * BK * 2;
  p.thr_io = thr_io;
  p.margin = margin;
  p.hist_base = (thr_io && hist && hist_base) ? hist_base : nullptr;
  p.hist = p.hist_base ? hist : nullptr;
  p.cand_score = cand_score;
  p.cand_id = cand_id;
  const int stage_bytes = ctas == 2 ? SCfg<2>::kStageBytes : SCfg<1>::kStageBytes;
  const int fixed = kc * BM * 8 + 256 + 1024;
  int stages = (227 * 1024 - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (ctas == 1 && stages > 4) stages = 4;
  p.stages = stages;
  const int smem = stages * stage_bytes + fixed;
  const long long units = static_cast<long long>(q_tiles) * splits;
  const int max_workers = clm_num_sms() / ctas;
  const int workers = units < max_workers ? static_cast<int>(units) : max_workers;
  // algorithmic bytes: index once + queries once + candidate lists
  ProfScope prof(CLM_K_SEARCH, 2.0 * nq * static_cast<double>(n) * dim,
                 2.0 * dim * (static_cast<double>(n) + nq) + 8.0 * nq * splits * kc,
                 static_cast<cudaStream_t>(stream));
  if (ctas == 2) return launch_search<2>(mq, me, p, smem, workers, static_cast<cudaStream_t>(stream));
  return launch_search<1>(mq, me, p, smem, workers, static_cast<cudaStream_t>(stream));
}

extern "C" int clm_topk_merge(const float* cand_score, const int32_t* cand_id, int nq, int lists, int kc,
                              float margin, const float* q_f32, const float* index_f32, int dim, int k,
                              int64_t id_offset, float* out_score, int64_t* out_id, int32_t* overflow,
                              void* stream) {
  CLM_REQUIRE(cand_score && cand_id && out_score && out_id, "clm_topk_merge: null argument");
  CLM_REQUIRE(nq > 0 && lists > 0 && kc >= 1 && k >= 1 && k <= kMaxMergeK && margin >= 0.f,
              "clm_topk_merge: bad sizes nq=%d lists=%d kc=%d k=%d (k <= %d)", nq, lists, kc, k, kMaxMergeK);
  CLM_REQUIRE(index_f32 == nullptr || (q_f32 != nullptr && dim % 4 == 0),
              "clm_topk_merge: re-scoring needs fp32 queries and dim %% 4 == 0");
  return launch_merge<int32_t>(cand_score, cand_id, nq, lists, kc, static_cast<long long>(lists) * kc, kc, kc, margin,
                               q_f32, index_f32, dim, k,
                               static_cast<long long>(id_offset), out_score,
                               reinterpret_cast<long long*>(out_id), overflow, static_cast<cudaStream_t>(stream));
}

extern "C" int clm_topk_merge_sorted(const float* in_score, const int64_t* in_id, int nq, int lists,
                                     int k, float* out_score, int64_t* out_id, void* stream) {
  CLM_REQUIRE(in_score && in_id && out_score && out_id, "clm_topk_merge_sorted: null argument");
  CLM_REQUIRE(nq > 0 && lists > 0 && k >= 1 && k <= kMaxMergeK, "clm_topk_merge_sorted: bad sizes (k <= %d)",
              kMaxMergeK);
  return launch_merge<long long>(in_score, reinterpret_cast<const long long*>(in_id), nq, lists, k,
                                 static_cast<long long>(lists) * k, k, k, 0.f, nullptr, nullptr, 0, k, 0, out_score,
                                 reinterpret_cast<long long*>(out_id), nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" size_t clm_topk_gather_chunk_bytes(int nq, int k) {
  if (nq <= 0 || k <= 0) return 0;
  const size_t pairs = (static_cast<size_t>(nq) * k + 1) / 2 * 2;  // even: the next rank's ids stay 8-byte aligned
  return pairs * 12;
}

extern "C" int clm_topk_merge_gathered(const void* gathered, int world, int nq, int k, float* out_score,
                                       int64_t* out_id, void* stream) {
  CLM_REQUIRE(gathered && out_score && out_id, "clm_topk_merge_gathered: null argument");
  CLM_REQUIRE(world > 0 && nq > 0 && k >= 1 && k <= kMaxMergeK, "clm_topk_merge_gathered: bad sizes (k <= %d)",
              kMaxMergeK);
  CLM_REQUIRE((reinterpret_cast<uintptr_t>(gathered) & 7) == 0, "clm_topk_merge_gathered: buffer not 8-byte aligned");
  const size_t chunk = clm_topk_gather_chunk_bytes(nq, k);
  const size_t pairs = chunk / 12;
  const long long* ids = reinterpret_cast<const long long*>(gathered);
  const float* scores = reinterpret_cast<const float*>(static_cast<const uint8_t*>(gathered) + pairs * 8);
  // rank l's chunk starts l * chunk bytes further: chunk / 8 id elements, chunk / 4 score elements (both exact:
  // chunk = 12 * pairs with pairs even); inside a chunk query qi's k entries start at qi * k
  return launch_merge<long long>(scores, ids, nq, world, k, k, static_cast<long long>(chunk / 4),
                                 static_cast<long long>(chunk / 8), 0.f, nullptr, nullptr, 0, k, 0, out_score,
                                 reinterpret_cast<long long*>(out_id), nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int clm_topk_row(const float* scores, int n, int k, int64_t id_offset, float* out_score,
                            int64_t* out_id, void* stream) {
  CLM_REQUIRE(scores && out_score && out_id && n >= 1 && k >= 1 && k <= n && k <= kMaxSel,
              "clm_topk_row: bad argument (n=%d k=%d; k <= min(n, %d))", n, k, kMaxSel);
  ProfScope prof(CLM_K_MERGE, 0.0, 6.0 * 4.0 * n, static_cast<cudaStream_t>(stream));
  topk_row_kernel<<<1, 1024, kMaxSel * 12, static_cast<cudaStream_t>(stream)>>>(
      scores, n, k, static_cast<long long>(id_offset), out_score, reinterpret_cast<long long*>(out_id));
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_cosine_gemv(const float* q_f32, const float* index_f32, int n, int dim, float* out,
                               void* stream) {
  CLM_REQUIRE(q_f32 && index_f32 && out && n >= 0 && dim > 0 && dim % 4 == 0,
              "clm_cosine_gemv: bad argument");
  if (n == 0) return CLM_OK;
  ProfScope prof(CLM_K_SEARCH, 2.0 * n * dim, 4.0 * dim * (static_cast<double>(n) + 1) + 4.0 * n,
                 static_cast<cudaStream_t>(stream));
  gemv_kernel<<<(n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(q_f32, index_f32, n, dim, out);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_kth_lower_bound(const float* x, int rows, int n, int kth, float guard, float* out, void* stream) {
  CLM_REQUIRE(x && out && rows > 0 && n >= kSeedGroups && kth >= 1 && kth <= kSeedGroups / 4,
              "clm_kth_lower_bound: bad argument (rows=%d n=%d kth=%d; n >= %d, kth <= %d)", rows, n, kth, kSeedGroups,
              kSeedGroups / 4);
  ProfScope prof(CLM_K_MERGE, 0.0, 4.0 * rows * n, static_cast<cudaStream_t>(stream));
  kth_lower_bound_kernel<<<rows, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, n, kth, guard, out);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

extern "C" int clm_kth_largest(const float* x, int rows, int n, int kth, float guard, float* out,
                               void* stream) {
  CLM_REQUIRE(x && out && rows > 0 && n > 0 && n <= 16384 && kth >= 1 && kth <= n,
              "clm_kth_largest: bad argument (rows=%d n=%d kth=%d; n <= 16384)", rows, n, kth);
  const int smem = n * 4;
  if (smem > 48 * 1024)
    CLM_CUDA_CHECK(cudaFuncSetAttribute(kth_largest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ProfScope prof(CLM_K_MERGE, 0.0, 4.0 * rows * n, static_cast<cudaStream_t>(stream));
  kth_largest_kernel<<<rows, kSelThreads, smem, static_cast<cudaStream_t>(stream)>>>(x, n, kth, guard, out);
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}
