// clm_attention.cu — fused attention for CLIP ViT towers on tcgen05 / TMEM.
//
// CLIP sequences are short (T = 50 / 77 / 197 / 257), so the whole score row block of a
// 128-query tile fits in tensor memory and no online softmax is needed.  The kernel is
// persistent (one CTA per SM) and warp-specialised; a work item is one (batch, head):
//
//   warp 0      TMA producer: Q, K, V of the next items (128-byte-swizzled boxes straight out of
//               the fused QKV activation [B*T, 3D]) into a ring of shared-memory stages.
//               K and V are loaded ONCE per (batch, head) and shared by all its query tiles.
//   warp 1      MMA issuer.  Per 128-query tile t:  S = Q K^T  (SS form, K-major A and B) into the
//               TMEM region of parity t % 2;  later  O = P V  (TS form: P is read from TMEM, V is an
//               MN-major shared-memory operand, so neither P nor V^T is ever materialised).  Tiles of
//               the two parities are independent streams; the warp polls their barriers and issues
//               whatever is ready (the tensor pipe executes in issue order, so S(t+2) may reuse the
//               S/P columns of tile t as soon as PV(t) has been issued).
//   warps 2-17  softmax: 2 groups (tile parity) x 2 threads per query row x 4 TMEM lane quarters.
//               pass 1 row max, pass 2 p = 2^(s*c - max*c) (one FFMA + one MUFU.EX2), row sum,
//               P written back to TMEM as packed bf16 over the S columns already consumed; the two
//               threads of a row take alternate 32-column chunks and exchange max / sum through
//               shared memory.  Then each reads its half of O, scales by 1/rowsum, stores bf16.
//
// The TMA warp is one or more items ahead, so HBM latency is off the critical path.
// TMEM plan: see clm_attention_launch (S fp32 -> P bf16x2 aliases its first half; O fp32 separate,
// or inside the dead upper part of its own S region when 512 columns are not enough).
// Softmax is fp32 (modeling_clip.py:274); scale 1/sqrt(64) (modeling_clip.py:271).
#include <stdlib.h>

#include "clm_common.cuh"

namespace {

using namespace clm;

constexpr int kThreads = 640;  // TMA warp + MMA warp + 16 softmax warps + 2 extra-token warps
constexpr int kTailWarp = 18;  // warps 18, 19 compute the extra query row (T = 128k + 1) on the CUDA cores
constexpr int kXtBytes = 5504;  // extra-token scratch: partial dots [2][2][128] f32, v_x halves 16 x 64 B, 2 p rows of 288 f32
constexpr int kOutStageBytes = 8 * 4096;  // per (group, lane quarter): a 32-row x 128-byte output slab for the TMA store
constexpr int kXchBytes = 4096;   // row max / row sum exchanged between the two threads of a query row
constexpr int kHeadDim = 64;
constexpr int kMaxStages = 6;

// Optional timeline tracing (compile with -DCLM_ATTN_TRACE; tools/attn_trace.py): lane 0 of every warp of
// CTA 0 stamps clock64() at the hand-off points of its first tiles.
#ifdef CLM_ATTN_TRACE
__device__ unsigned long long* g_trace = nullptr;
constexpr int kTraceTiles = 12, kTraceEvents = 8;
#define TRACE(tile_seq, ev)                                                                        \
  do {                                                                                             \
    if (g_trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (tile_seq) < kTraceTiles)         \
      g_trace[((threadIdx.x >> 5) * kTraceTiles + (tile_seq)) * kTraceEvents + (ev)] = clock64();  \
  } while (0)
#else
#define TRACE(tile_seq, ev) do {} while (0)
#endif

struct AttnParams {
  int T, H, Tp, mtiles, num_items, stages, stage_bytes, nslots;
  int stage_out;  // output rows leave through shared memory + TMA tile stores (needs kOutStageBytes)
  int blocks, nb0, nb1;  // key blocks per tile (2 = online softmax over two blocks of nb0 / nb1 keys)
  // T = 128k + 1 (a ViT's class token on top of a 128-multiple of patches): the tensor cores see Tk = T - 1
  // queries x Tk keys; the extra KEY is added by the softmax threads (one 64-long dot product and a rank-1
  // update of O per row) and the extra QUERY row is computed by a dedicated warp, both from the staged
  // shared-memory tiles.  Otherwise Tk == Tp.
  int Tk, xt;
  // serial != 0: the exponential pass (pass 2) of tile t starts only after P(t-1) has been published by the
  // other softmax group, so the two groups take turns on the MUFU pipe instead of running in lockstep (both
  // exponentiating, then both waiting for their P V and output phases with the pipe idle)
  int serial;
  // TMEM column of the S / P region and of the O accumulator used by tiles of parity 0 / 1;
  // o_alias: O lives inside the same parity's S region (dead by then), so S(t+2) waits for the drain
  int s_col0, s_col1, o_col0, o_col1, o_alias0, o_alias1;
  __host__ __device__ int s_col(int par) const { return par ? s_col1 : s_col0; }
  __host__ __device__ int o_col(int par) const { return par ? o_col1 : o_col0; }
  __host__ __device__ int o_alias(int par) const { return par ? o_alias1 : o_alias0; }
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
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// dot product of 8 bf16 pairs, two chains
__device__ __forceinline__ void dot8(const uint4& a, const uint4& b, float& s0, float& s1) {
  s0 = fmaf(bf16_lo(a.x), bf16_lo(b.x), s0); s1 = fmaf(bf16_hi(a.x), bf16_hi(b.x), s1);
  s0 = fmaf(bf16_lo(a.y), bf16_lo(b.y), s0); s1 = fmaf(bf16_hi(a.y), bf16_hi(b.y), s1);
  s0 = fmaf(bf16_lo(a.z), bf16_lo(b.z), s0); s1 = fmaf(bf16_hi(a.z), bf16_hi(b.z), s1);
  s0 = fmaf(bf16_lo(a.w), bf16_lo(b.w), s0); s1 = fmaf(bf16_hi(a.w), bf16_hi(b.w), s1);
}

// max of 32 scores with four independent chains (the single-chain form serialises on FMNMX latency)
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    m0 = fmaxf(m0, __uint_as_float(v[i]));
    m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
    m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
    m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
__device__ __forceinline__ float chunk_max_masked(const uint32_t (&v)[32], float m, int base, int valid) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (base + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
  return m;
}

// p = 2^(s*c - max*c) for 32 scores -> 16 packed bf16x2 words; returns the chunk's sum of p
template <bool kMasked>
__device__ __forceinline__ float chunk_exp(const uint32_t (&v)[32], uint32_t (&pk)[16], float scale,
                                           float neg_mx, int base, int valid) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float e0 = fast_exp2(fmaf(__uint_as_float(v[i]), scale, neg_mx));
    float e1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), scale, neg_mx));
    float e2 = fast_exp2(fmaf(__uint_as_float(v[i + 2]), scale, neg_mx));
    float e3 = fast_exp2(fmaf(__uint_as_float(v[i + 3]), scale, neg_mx));
    if (kMasked) {
      e0 = (base + i < valid) ? e0 : 0.f;
      e1 = (base + i + 1 < valid) ? e1 : 0.f;
      e2 = (base + i + 2 < valid) ? e2 : 0.f;
      e3 = (base + i + 3 < valid) ? e3 : 0.f;
    }
    s0 += e0; s1 += e1; s2 += e2; s3 += e3;
    pk[i / 2] = pack_bf16x2(e0, e1);
    pk[i / 2 + 1] = pack_bf16x2(e2, e3);
  }
  return (s0 + s1) + (s2 + s3);
}

// The extra query row (index Tk = 32 kNF) of one (batch, head) item, from the staged Q / K / V tiles at `sb`
// (SWIZZLE_128B rows of 128 bytes).  Lane j scores keys j, j + 32, ... and, redundantly, the extra key Tk;
// the warp reduces max and sum; lane l then accumulates output dimensions 2l, 2l+1 over all keys (for one
// key the 32 lanes read the 128 contiguous bytes of its V row).  P stays fp32 on this path.
template <int kNF>
__device__ __forceinline__ void tail_row(uint32_t sb, int kv_bytes, float* prow, int lane, uint32_t* out_row) {
  constexpr int Tk = 32 * kNF;
  constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;
  const uint32_t kb = sb + static_cast<uint32_t>(kv_bytes), vb = kb + static_cast<uint32_t>(kv_bytes);
  const uint32_t lsw = static_cast<uint32_t>(lane & 7);
  float sc[kNF], sx = 0.f;
#pragma unroll
  for (int i = 0; i < kNF; ++i) sc[i] = 0.f;
  const uint32_t k_lane = kb + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
  for (int c = 0; c < 8; ++c) {  // 8 dimensions of the query at a time (row Tk starts a swizzle group: chunks in place)
    const uint4 qw = ld_shared_v4(sb + static_cast<uint32_t>(Tk * 128 + (c << 4)));
    const float q0 = bf16_lo(qw.x), q1 = bf16_hi(qw.x), q2 = bf16_lo(qw.y), q3 = bf16_hi(qw.y);
    const float q4 = bf16_lo(qw.z), q5 = bf16_hi(qw.z), q6 = bf16_lo(qw.w), q7 = bf16_hi(qw.w);
    const uint32_t koff = k_lane + ((static_cast<uint32_t>(c) ^ lsw) << 4);  // (j & 7) == (lane & 7)
    uint4 w[kNF + 1];
#pragma unroll
    for (int i = 0; i < kNF; ++i) w[i] = ld_shared_v4(koff + static_cast<uint32_t>(i) * 4096u);
    w[kNF] = ld_shared_v4(kb + static_cast<uint32_t>(Tk * 128 + (c << 4)));
#pragma unroll
    for (int i = 0; i <= kNF; ++i) {
      float a = (i < kNF) ? sc[i < kNF ? i : 0] : sx;
      a = fmaf(bf16_lo(w[i].x), q0, a); a = fmaf(bf16_hi(w[i].x), q1, a);
      a = fmaf(bf16_lo(w[i].y), q2, a); a = fmaf(bf16_hi(w[i].y), q3, a);
      a = fmaf(bf16_lo(w[i].z), q4, a); a = fmaf(bf16_hi(w[i].z), q5, a);
      a = fmaf(bf16_lo(w[i].w), q6, a); a = fmaf(bf16_hi(w[i].w), q7, a);
      if (i < kNF) sc[i < kNF ? i : 0] = a; else sx = a;
    }
  }
  float mx = sx;
#pragma unroll
  for (int i = 0; i < kNF; ++i) mx = fmaxf(mx, sc[i]);
  mx = warp_max(mx);
  const float neg_mx = -mx * kScaleLog2e;
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < kNF; ++i) {
    const float pj = fast_exp2(fmaf(sc[i], kScaleLog2e, neg_mx));
    l += pj;
    prow[lane + 32 * i] = pj;
  }
  const float p_x = fast_exp2(fmaf(sx, kScaleLog2e, neg_mx));
  l = warp_sum(l) + p_x;
  __syncwarp();
  const uint32_t pr = smem_u32(prow);
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  const uint32_t vl = vb + static_cast<uint32_t>((lane & 3) * 4);
  const uint32_t hi3 = static_cast<uint32_t>(lane >> 2);
#pragma unroll 2
  for (int j0 = 0; j0 < Tk; j0 += 8) {
    const uint4 pa = ld_shared_v4(pr + static_cast<uint32_t>(j0) * 4u);
    const uint4 pb = ld_shared_v4(pr + static_cast<uint32_t>(j0 + 4) * 4u);
    const uint32_t row = vl + static_cast<uint32_t>(j0) * 128u;
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = ld_shared_u32(row + u * 128 + ((hi3 ^ static_cast<uint32_t>(u)) << 4));
    o0 = fmaf(__uint_as_float(pa.x), bf16_lo(w[0]), o0); o1 = fmaf(__uint_as_float(pa.x), bf16_hi(w[0]), o1);
    o2 = fmaf(__uint_as_float(pa.y), bf16_lo(w[1]), o2); o3 = fmaf(__uint_as_float(pa.y), bf16_hi(w[1]), o3);
    o0 = fmaf(__uint_as_float(pa.z), bf16_lo(w[2]), o0); o1 = fmaf(__uint_as_float(pa.z), bf16_hi(w[2]), o1);
    o2 = fmaf(__uint_as_float(pa.w), bf16_lo(w[3]), o2); o3 = fmaf(__uint_as_float(pa.w), bf16_hi(w[3]), o3);
    o0 = fmaf(__uint_as_float(pb.x), bf16_lo(w[4]), o0); o1 = fmaf(__uint_as_float(pb.x), bf16_hi(w[4]), o1);
    o2 = fmaf(__uint_as_float(pb.y), bf16_lo(w[5]), o2); o3 = fmaf(__uint_as_float(pb.y), bf16_hi(w[5]), o3);
    o0 = fmaf(__uint_as_float(pb.z), bf16_lo(w[6]), o0); o1 = fmaf(__uint_as_float(pb.z), bf16_hi(w[6]), o1);
    o2 = fmaf(__uint_as_float(pb.w), bf16_lo(w[7]), o2); o3 = fmaf(__uint_as_float(pb.w), bf16_hi(w[7]), o3);
  }
  {  // the extra key itself
    const uint32_t w = ld_shared_u32(vl + static_cast<uint32_t>(Tk) * 128u + (hi3 << 4));
    o0 = fmaf(p_x, bf16_lo(w), o0); o1 = fmaf(p_x, bf16_hi(w), o1);
  }
  const float inv = 1.0f / l;
  out_row[lane] = pack_bf16x2((o0 + o2) * inv, (o1 + o3) * inv);
}

template <bool kCausal>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map16,
                 const __grid_constant__ CUtensorMap map_out, __nv_bfloat16* __restrict__ out, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ostage = smem + p.stages * p.stage_bytes;  // 1024-byte aligned (stage_bytes is a multiple of 6144)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + (p.stage_out == 1 ? kOutStageBytes : 0));
  uint64_t* stage_full = bars;                     // [kMaxStages]
  uint64_t* stage_empty = bars + kMaxStages;       // [kMaxStages]
  uint64_t* s_full = bars + 2 * kMaxStages;        // [2] S ready (MMA commit), by tile parity
  uint64_t* p_full = s_full + 2;                   // [2] P written (128 softmax threads)
  uint64_t* o_full = s_full + 4;                   // [2] O ready (MMA commit)
  uint64_t* slot_free = s_full + 6;                // [2] O drained (128 softmax threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [group][half][128]
  float* xsum = xmax + 2 * 2 * 128;
  float* xdot = xsum + 2 * 2 * 128;                         // extra-token scratch (only when p.xt)
  uint8_t* vxs = reinterpret_cast<uint8_t*>(xdot + 2 * 2 * 128);
  float* prow = reinterpret_cast<float*>(vxs + 16 * 64);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T, H = p.H, Tp = p.Tp, Tk = p.Tk, D = p.H * kHeadDim;
  const int kv_bytes = Tp * 128;  // one of Q / K / V in a stage (Tp rows are loaded; the MMAs see Tk keys)
  const int n_local = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  const int n_tiles = n_local * p.mtiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map64);
    tma_prefetch_desc(&map16);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&stage_full[s], 1);
      // last PV's commit (+ the extra-token warp that owns the item) (+ the 4 warp pairs of each of the
      // item's tiles once their output slab, staged in the tile's dead Q rows, has been read by the TMA unit)
      mbar_init(&stage_empty[s], 1 + p.xt + (p.stage_out == 2 ? 4 * p.mtiles : 0));
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 256);
      mbar_init(&o_full[s], 1);
      mbar_init(&slot_free[s], 256);
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
#ifdef CLM_ATTN_TRACE
      int tl = 0;
#endif
      for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const int b = it / H, h = it % H;
        const int row_base = b * T;
        TRACE(tl, 0);
        mbar_wait(&stage_empty[st], ph ^ 1);
        TRACE(tl, 1);
#ifdef CLM_ATTN_TRACE
        ++tl;
#endif
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
    // The tensor pipe executes in issue order, so S(t+2) may overwrite the S/P columns of tile t as
    // soon as PV(t) has been ISSUED; it only waits when O(t) is aliased into those columns.
    const uint32_t idesc_pv = umma_idesc_bf16(128, kHeadDim, 0, 1);
    // A tile goes S -> (softmax) -> PV; tiles of the two parities are independent streams.  With two S
    // regions the warp polls both streams and issues whatever is ready, so one group never waits
    // behind the other group's barrier (ViT-B/16: the odd stream must wait for its aliased O to drain).
    struct Cursor {  // position of a stream inside the CTA's tile list
      int t, mt, st;
      uint32_t ph;
      __device__ void init(int t0, const AttnParams& p) {
        t = t0; mt = t0; st = 0; ph = 0;
        while (mt >= p.mtiles) { mt -= p.mtiles; bump(p); }
      }
      __device__ void bump(const AttnParams& p) {
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      __device__ void advance(int step, const AttnParams& p) {
        t += step; mt += step;
        while (mt >= p.mtiles) { mt -= p.mtiles; bump(p); }
      }
    };
    auto do_s = [&](const Cursor& c) {
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + c.st * p.stage_bytes);
        const uint32_t k_addr = q_addr + kv_bytes;
        const uint32_t sbase = tmem + static_cast<uint32_t>(p.s_col(c.t & 1));
        for (int n0 = 0; n0 < Tk; n0 += 256) {
          const int nn = (Tk - n0) < 256 ? (Tk - n0) : 256;
          const uint32_t idesc = umma_idesc_bf16(128, nn, 0, 0);
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_bf16_ss(sbase + n0, umma_desc_sw128(q_addr + c.mt * 16384 + k * 32, 1024),
                         umma_desc_sw128(k_addr + n0 * 128 + k * 32, 1024), idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[c.t & 1]);
      }
      __syncwarp();
    };
    int pv_cnt[kMaxStages];  // PVs issued per stage: the last one of an item releases its stage
#pragma unroll
    for (int i = 0; i < kMaxStages; ++i) pv_cnt[i] = 0;
    auto do_pv = [&](const Cursor& c) {
      // O(t) = P(t) V : P from TMEM (written by the softmax group), V MN-major from smem.  p_full(t)
      // also implies that the same group has drained O(t-2), whose columns this overwrites.
      const int b = c.t & 1;
      const bool last = (++pv_cnt[c.st] == p.mtiles);
      if (last) pv_cnt[c.st] = 0;
      if (lane == 0) {
        const uint32_t v_addr = smem_u32(smem + c.st * p.stage_bytes) + 2 * kv_bytes;
        const uint32_t pbase = tmem + static_cast<uint32_t>(p.s_col(b));
        const uint32_t obase = tmem + static_cast<uint32_t>(p.o_col(b));
        const int ksteps = Tk / 16;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_bf16_ts(obase, pbase + ks * 8, umma_desc_sw128_mn(v_addr + ks * 2048), idesc_pv,
                       ks != 0 ? 1u : 0u);
        umma_commit(&o_full[b]);
        if (last) umma_commit(&stage_empty[c.st]);  // stage reusable once these MMAs retire
      }
      __syncwarp();
    };
    if (p.blocks == 2) {
      // Two key blocks per tile (T > 224: the whole score row does not fit twice in TMEM).  Per stream:
      //   S0 -> [softmax blk 0] -> PV0 (O = P0 V0), S1 -> [softmax blk 1 + rescale of O] -> PV1 (O += P1 V1)
      // Every barrier of a stream completes twice per tile, so block 0 always waits parity 0, block 1 parity 1.
      auto mma_s = [&](const Cursor& c, int key0, int nk) {
        if (lane == 0) {
          const uint32_t q_addr = smem_u32(smem + c.st * p.stage_bytes);
          const uint32_t k_addr = q_addr + kv_bytes + static_cast<uint32_t>(key0) * 128u;
          const uint32_t sbase = tmem + static_cast<uint32_t>(p.s_col(c.t & 1));
          const uint32_t idesc = umma_idesc_bf16(128, nk, 0, 0);
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_bf16_ss(sbase, umma_desc_sw128(q_addr + c.mt * 16384 + k * 32, 1024),
                         umma_desc_sw128(k_addr + k * 32, 1024), idesc, k != 0 ? 1u : 0u);
          umma_commit(&s_full[c.t & 1]);
        }
        __syncwarp();
      };
      auto mma_pv = [&](const Cursor& c, int key0, int nk, bool first, bool release) {
        if (lane == 0) {
          const int b = c.t & 1;
          const uint32_t v_addr = smem_u32(smem + c.st * p.stage_bytes) + 2 * kv_bytes + static_cast<uint32_t>(key0) * 128u;
          const uint32_t pbase = tmem + static_cast<uint32_t>(p.s_col(b));
          const uint32_t obase = tmem + static_cast<uint32_t>(p.o_col(b));
          for (int ks = 0; ks < nk / 16; ++ks)
            umma_bf16_ts(obase, pbase + ks * 8, umma_desc_sw128_mn(v_addr + ks * 2048), idesc_pv,
                         (!first || ks != 0) ? 1u : 0u);
          umma_commit(&o_full[b]);
          if (release) umma_commit(&stage_empty[c.st]);
        }
        __syncwarp();
      };
      Cursor cur[2];
      int step[2] = {0, 0};  // 0: S0 pending, 1: PV0 + S1 pending, 2: PV1 pending
      cur[0].init(0, p); cur[1].init(1, p);
      while (cur[0].t < n_tiles || cur[1].t < n_tiles) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (cur[b].t >= n_tiles) continue;
          if (step[b] == 0) {
            if (mbar_try_wait(&stage_full[cur[b].st], cur[b].ph)) {
              tc_fence_after();
              mma_s(cur[b], 0, p.nb0);
              step[b] = 1;
            }
          } else if (step[b] == 1) {
            if (mbar_try_wait(&p_full[b], 0)) {
              tc_fence_after();
              mma_pv(cur[b], 0, p.nb0, true, false);
              mma_s(cur[b], p.nb0, p.nb1);
              step[b] = 2;
            }
          } else {
            if (mbar_try_wait(&p_full[b], 1)) {
              tc_fence_after();
              const bool last = (++pv_cnt[cur[b].st] == p.mtiles);
              if (last) pv_cnt[cur[b].st] = 0;
              mma_pv(cur[b], p.nb0, p.nb1, false, last);
              cur[b].advance(2, p);
              step[b] = 0;
            }
          }
        }
      }
    } else if (p.nslots == 2) {
      Cursor sc[2], pc[2];  // next S / next PV of each stream
      sc[0].init(0, p); sc[1].init(1, p); pc[0].init(0, p); pc[1].init(1, p);
      while (pc[0].t < n_tiles || pc[1].t < n_tiles) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          // PV(t): S(t) has been issued and the group has published P(t)
          if (pc[b].t < sc[b].t && mbar_try_wait(&p_full[b], static_cast<uint32_t>((pc[b].t >> 1) & 1))) {
            tc_fence_after();
            TRACE(pc[b].t, 1);
            do_pv(pc[b]);
            pc[b].advance(2, p);
          }
          // S(t): PV(t-2) has been issued (in-order pipe: its P columns are safe); an aliased O(t-2)
          // has been drained; the item's Q/K/V have landed
          if (sc[b].t < n_tiles && sc[b].t - 2 < pc[b].t) {
            bool ok = true;
            if (p.o_alias(b) && sc[b].t >= 2)
              ok = mbar_try_wait(&slot_free[b], static_cast<uint32_t>(((sc[b].t - 2) >> 1) & 1));
            if (ok) ok = mbar_try_wait(&stage_full[sc[b].st], sc[b].ph);
            if (ok) {
              tc_fence_after();
              TRACE(sc[b].t, 0);
              do_s(sc[b]);
              sc[b].advance(2, p);
            }
          }
        }
      }
    } else {
      // one S region: strictly S(t), PV(t), S(t+1), ... (the groups alternate)
      Cursor c;
      c.init(0, p);
      if (n_tiles > 0) {
        mbar_wait(&stage_full[c.st], c.ph);
        tc_fence_after();
        do_s(c);
      }
      while (c.t < n_tiles) {
        mbar_wait(&p_full[c.t & 1], static_cast<uint32_t>((c.t >> 1) & 1));
        tc_fence_after();
        do_pv(c);
        c.advance(1, p);
        if (c.t < n_tiles) {
          mbar_wait(&stage_full[c.st], c.ph);
          tc_fence_after();
          do_s(c);
        }
      }
    }
  } else if (warp < kTailWarp) {
    // ================= softmax groups =================
    // 16 warps = 2 groups (tile parity) x 2 column halves x 4 TMEM lane quarters.  A query row is
    // shared by two threads (same lane of two warps with the same quarter): each reduces / exponentiates
    // every other 32-column chunk of the row and they exchange row max and row sum through shared memory.
    // Twice the warps per tile halves the latency of the softmax phase that the tensor pipe waits on.
    const int sel = (warp - 2) >> 2;
    const int g = sel & 1;               // group = parity of the tiles it owns
    const int hf = sel >> 1;             // which half of the row's chunks (and of O's columns)
    const int q = warp & 3;              // TMEM lane quarter of this warp
    const int r = q * 32 + lane;         // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int nchunks = (Tk + 31) / 32;
    constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t srow = tmem + lane_off + static_cast<uint32_t>(p.s_col(g));
    const uint32_t orow = tmem + lane_off + static_cast<uint32_t>(p.o_col(g)) + static_cast<uint32_t>(hf * 32);
    float* my_max = xmax + (g * 2 + hf) * 128 + r;
    float* other_max = xmax + (g * 2 + (hf ^ 1)) * 128 + r;
    float* my_sum = xsum + (g * 2 + hf) * 128 + r;
    float* other_sum = xsum + (g * 2 + (hf ^ 1)) * 128 + r;
    const int pair_bar = 1 + g * 4 + q;  // named barrier of the two warps that share these 32 rows
    int t = 0;
    if (p.blocks == 2) {
      // ---- online softmax over two key blocks (see the MMA warp).  m_run / l_part are the running row
      // max and this thread's share of the row sum (in units of 2^(-m_run*c)).
      for (int it = blockIdx.x; it < p.num_items; it += gridDim.x)
      for (int mt = 0; mt < p.mtiles; ++mt, ++t) {
        if ((t & 1) != g) continue;
        const int b = it / H, h = it - b * H;
        const int qi = mt * 128 + r;
        const bool warp_live = mt * 128 + q * 32 < T;
        int valid = T;
        if (kCausal) valid = (qi + 1 < T) ? qi + 1 : T;
        float m_run = -INFINITY, l_part = 0.f;
        uint32_t v[32];
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          const int key0 = blk ? p.nb0 : 0, nk = blk ? p.nb1 : p.nb0;
          const int nchb = (nk + 31) / 32;
          int vb = valid - key0;  // valid keys of this block for this row
          vb = vb < 0 ? 0 : (vb > nk ? nk : vb);
          const int niter = (nchb + 1) >> 1;
          mbar_wait(&s_full[g], static_cast<uint32_t>(blk));
          tc_fence_after();
          float mx = -INFINITY;
          if (warp_live) {
            for (int c = hf; c < nchb; c += 2) {
              tmem_ld_32x32b_x32(srow + c * 32, v);
              tmem_ld_wait();
              mx = (c * 32 + 32 <= vb) ? chunk_max(v, mx) : chunk_max_masked(v, mx, c * 32, vb);
            }
          }
          *my_max = mx;
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          mx = fmaxf(mx, *other_max);
          const float m_new = fmaxf(m_run, mx);
          if (blk == 1) {
            // O = P0 V0 was accumulated relative to m_run: bring it (and l) to the new maximum
            const float alpha = fast_exp2((m_run - m_new) * kScaleLog2e);
            mbar_wait(&o_full[g], 0);
            tc_fence_after();
            if (warp_live && __any_sync(0xffffffffu, alpha != 1.0f)) {
              tmem_ld_32x32b_x32(orow, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_32x32b_x32(orow, v);
              tmem_st_wait();
            }
            l_part *= alpha;
          }
          m_run = m_new;
          float sum = 0.f;
          if (warp_live) {
            const float neg_mx = -m_run * kScaleLog2e;
            uint32_t pk[16];
            for (int i = 0; i < niter; ++i) {
              const int c = 2 * i + hf;
              if (c < nchb) {
                tmem_ld_32x32b_x32(srow + c * 32, v);
                tmem_ld_wait();
                sum += (c * 32 + 32 <= vb) ? chunk_exp<false>(v, pk, kScaleLog2e, neg_mx, c * 32, vb)
                                           : chunk_exp<true>(v, pk, kScaleLog2e, neg_mx, c * 32, vb);
              }
              tc_fence_before();
              asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
              tc_fence_after();
              if (c < nchb) tmem_st_32x32b_x16(srow + c * 16, pk);
            }
            tmem_st_wait();
          }
          l_part += sum;
          if (blk == 1) *my_sum = l_part;
          tc_fence_before();
          mbar_arrive(&p_full[g]);
        }
        mbar_wait(&o_full[g], 1);
        tc_fence_after();
        const float total = l_part + *other_sum;
        if (warp_live) {
          tmem_ld_32x32b_x32(orow, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        if (warp_live && qi < T) {
          const float inv = 1.0f / total;
          uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(b * T + qi) * D + h * kHeadDim + hf * 32);
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
    } else {
    const bool xt = !kCausal && p.xt != 0;
    const uint32_t vx_w = smem_u32(vxs + (warp - 2) * 64);  // this warp's copy of its half of the extra V row
    float* my_dot = xdot + (g * 2 + hf) * 128 + r;
    int li = 0;  // items seen by this CTA: item li sits in stage li % stages
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ++li)
    for (int mt = 0; mt < p.mtiles; ++mt, ++t) {
      if ((t & 1) != g) continue;
      const int b = it / H, h = it - b * H;
      const uint32_t par = static_cast<uint32_t>((t >> 1) & 1);
      const int qi = mt * 128 + r;                     // query position in the sequence
      const bool warp_live = mt * 128 + q * 32 < T;    // does this warp own any real row?
      int valid = T < Tk ? T : Tk;
      if (kCausal) valid = (qi + 1 < T) ? qi + 1 : T;
      // ---- extra key (row Tk of K / V): this thread's half of q_row . k_x, and a private copy of its
      // half of v_x (the stage may be refilled before the output phase of the item's last tile)
      float px = 0.f;
      if (xt) {
        const int st = li % p.stages;
        mbar_wait(&stage_full[st], static_cast<uint32_t>((li / p.stages) & 1));
        TRACE(t, 7);
        const uint32_t sb = smem_u32(smem + st * p.stage_bytes);
        const uint32_t qa = sb + static_cast<uint32_t>(qi) * 128u;
        const uint32_t ka = sb + static_cast<uint32_t>(kv_bytes + Tk * 128);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t c = static_cast<uint32_t>(hf * 4 + jj);
          const uint4 a = ld_shared_v4(qa + ((c ^ static_cast<uint32_t>(r & 7)) << 4));
          const uint4 kx = ld_shared_v4(ka + (c << 4));
          dot8(a, kx, d0, d1);
        }
        *my_dot = d0 + d1;
        if (lane < 4) {
          const uint4 w = ld_shared_v4(sb + static_cast<uint32_t>(2 * kv_bytes + Tk * 128 + ((hf * 4 + lane) << 4)));
          st_shared_v4(vx_w + (lane << 4), w.x, w.y, w.z, w.w);
        }
        __syncwarp();
      }
      // chunks this warp pair has to look at: under the causal mask nothing right of its last row counts
      int nch = nchunks;
      if (kCausal) {
        const int wv = (mt * 128 + q * 32 + 32 < T) ? mt * 128 + q * 32 + 32 : T;
        nch = (wv + 31) / 32;
      }
      // the two threads of a row take alternate chunks: thread hf owns chunks 2i + hf
      const int niter = (nch + 1) >> 1;

      TRACE(t, 0);
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      TRACE(t, 1);
      uint32_t v[32];
      // ---- pass 1: max over this thread's chunks, then the row max
      float mx = -INFINITY;
      if (warp_live) {
        for (int c = hf; c < nch; c += 2) {
          tmem_ld_32x32b_x32(srow + c * 32, v);
          tmem_ld_wait();
          mx = (c * 32 + 32 <= valid) ? chunk_max(v, mx) : chunk_max_masked(v, mx, c * 32, valid);
        }
      }
      TRACE(t, 2);
      *my_max = mx;
      if (p.stage_out == 1 && hf == 0 && lane == 0) bulk_wait_read<0>();  // the previous tile's slab has left smem
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, *other_max);
      float sx = 0.f;
      if (xt) {  // both threads of the row add the two halves in the same order
        sx = xdot[(g * 2 + 0) * 128 + r] + xdot[(g * 2 + 1) * 128 + r];
        mx = fmaxf(mx, sx);
      }
      TRACE(t, 3);
      if (p.serial && t >= 1) mbar_wait(&p_full[g ^ 1], static_cast<uint32_t>(((t - 1) >> 1) & 1));
      // ---- pass 2: p = 2^(s*c - max*c), partial row sum, P (bf16x2) written over S.  P(c) lands in the
      // columns of S chunk c/2, so the pair synchronises once per iteration: by then both threads hold
      // every chunk up to 2i+1 in registers and the columns of chunk i are dead.
      float sum = 0.f;
      if (warp_live) {
        const float neg_mx = -mx * kScaleLog2e;
        if (xt) {
          px = fast_exp2(fmaf(sx, kScaleLog2e, neg_mx));  // fp32 weight of the extra key (not rounded to bf16)
          if (hf == 0) sum = px;
        }
        uint32_t pk[16];
        for (int i = 0; i < niter; ++i) {
          const int c = 2 * i + hf;
          if (c < nch) {
            tmem_ld_32x32b_x32(srow + c * 32, v);
            tmem_ld_wait();
            sum += (c * 32 + 32 <= valid) ? chunk_exp<false>(v, pk, kScaleLog2e, neg_mx, c * 32, valid)
                                          : chunk_exp<true>(v, pk, kScaleLog2e, neg_mx, c * 32, valid);
          }
          tc_fence_before();
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          tc_fence_after();
          if (c < nch) tmem_st_32x32b_x16(srow + c * 16, pk);
        }
        if (kCausal && hf == 1 && nch < nchunks) {  // P right of the causal frontier is zero (PV reads it)
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
          for (int c = nch; c < nchunks; ++c) tmem_st_32x32b_x16(srow + c * 16, pk);
        }
        tmem_st_wait();
      }
      *my_sum = sum;
      tc_fence_before();
      TRACE(t, 4);
      mbar_arrive(&p_full[g]);

      mbar_wait(&o_full[g], par);
      tc_fence_after();
      TRACE(t, 5);
      sum += *other_sum;  // published before the partner's p_full arrive, which o_full transitively follows
      if (warp_live) {
        tmem_ld_32x32b_x32(orow, v);
        tmem_ld_wait();
      }
      tc_fence_before();
      TRACE(t, 6);
      mbar_arrive(&slot_free[g]);  // this half of O(t) is in registers: its columns may be overwritten
      if (xt && warp_live) {  // O += p_x * v_x
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint4 w = ld_shared_v4(vx_w + (jj << 4));
          v[8 * jj + 0] = __float_as_uint(fmaf(px, bf16_lo(w.x), __uint_as_float(v[8 * jj + 0])));
          v[8 * jj + 1] = __float_as_uint(fmaf(px, bf16_hi(w.x), __uint_as_float(v[8 * jj + 1])));
          v[8 * jj + 2] = __float_as_uint(fmaf(px, bf16_lo(w.y), __uint_as_float(v[8 * jj + 2])));
          v[8 * jj + 3] = __float_as_uint(fmaf(px, bf16_hi(w.y), __uint_as_float(v[8 * jj + 3])));
          v[8 * jj + 4] = __float_as_uint(fmaf(px, bf16_lo(w.z), __uint_as_float(v[8 * jj + 4])));
          v[8 * jj + 5] = __float_as_uint(fmaf(px, bf16_hi(w.z), __uint_as_float(v[8 * jj + 5])));
          v[8 * jj + 6] = __float_as_uint(fmaf(px, bf16_lo(w.w), __uint_as_float(v[8 * jj + 6])));
          v[8 * jj + 7] = __float_as_uint(fmaf(px, bf16_hi(w.w), __uint_as_float(v[8 * jj + 7])));
        }
      }
      if (p.stage_out) {
        // A thread owns 64 bytes of one output row; writing them straight to global memory costs 32
        // scattered 16-byte transactions per warp instruction (measured: ~1.7k clocks per tile).  The
        // two warps of a row quarter assemble their 32 x 128-byte slab in shared memory (SWIZZLE_128B
        // pattern of the output map) and one lane hands it to the TMA unit; rows >= T are clipped.
        // stage_out == 2: no room for a staging buffer; the slab goes into this warp pair's 32 rows of the
        // tile's own Q block, which nothing reads after S has been computed (same swizzled row layout)
        const int st_cur = li % p.stages;
        const uint32_t ost = (p.stage_out == 2)
                                 ? smem_u32(smem + st_cur * p.stage_bytes) + static_cast<uint32_t>(mt * 128 + q * 32) * 128u
                                 : smem_u32(ostage + (g * 4 + q) * 4096);
        if (warp_live) {
          const float inv = 1.0f / sum;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            st_shared_v4(ost + static_cast<uint32_t>(lane) * 128u + ((static_cast<uint32_t>(hf * 4 + jj) ^ (lane & 7)) << 4),
                         pack_bf16x2(__uint_as_float(v[8 * jj + 0]) * inv, __uint_as_float(v[8 * jj + 1]) * inv),
                         pack_bf16x2(__uint_as_float(v[8 * jj + 2]) * inv, __uint_as_float(v[8 * jj + 3]) * inv),
                         pack_bf16x2(__uint_as_float(v[8 * jj + 4]) * inv, __uint_as_float(v[8 * jj + 5]) * inv),
                         pack_bf16x2(__uint_as_float(v[8 * jj + 6]) * inv, __uint_as_float(v[8 * jj + 7]) * inv));
          fence_proxy_async_smem();
        }
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        if (warp_live && hf == 0 && lane == 0) {
          tma_store_3d(&map_out, ost, h * kHeadDim, mt * 128 + q * 32, b);
          bulk_commit();
        }
        if (p.stage_out == 2 && hf == 0 && lane == 0) {
          // release this pair's share of the stage as soon as the TMA unit has read the slab: the refill of
          // the stage (two stages only) is on the critical path of the tile after next
          bulk_wait_read<0>();
          mbar_arrive(&stage_empty[st_cur]);
        }
      } else if (warp_live && qi < T) {
        const float inv = 1.0f / sum;
        uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(b * T + qi) * D + h * kHeadDim + hf * 32);
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
    }
    if (p.stage_out && hf == 0 && lane == 0) bulk_wait<0>();  // every slab of this warp pair has reached memory
  } else if (!kCausal && p.xt) {
    // ================= extra-token warps: the query row Tk of every item, on the CUDA cores =================
    // (one row per (batch, head): a third 128-row tensor-core tile would be 1/128 used).  The two warps take
    // alternate items; see tail_row.
    const int tw = warp - kTailWarp;
    float* my_prow = prow + tw * 288;
    int st = 0, tl = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ++tl) {
      if ((tl & 1) == tw) {
        const int b = it / H, h = it - b * H;
        TRACE(tl, 0);
        mbar_wait(&stage_full[st], ph);
        TRACE(tl, 1);
        const uint32_t sb = smem_u32(smem + st * p.stage_bytes);
        uint32_t* orow = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * T + Tk) * D + h * kHeadDim);
        if (Tk == 256) tail_row<8>(sb, kv_bytes, my_prow, lane, orow);
        else tail_row<4>(sb, kv_bytes, my_prow, lane, orow);
        __syncwarp();  // every lane is done with the stage and with its p row
        TRACE(tl, 3);
        if (lane == 0) mbar_arrive(&stage_empty[st]);
      }
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}


// =====================================================================================================
// attention_kernel_v2 — one softmax thread per query row, software-pipelined TMEM loads.
//
// Same work decomposition, TMEM plans, TMA producer and MMA issuer as attention_kernel (single key block,
// two S regions).  What changes is the softmax side, which was latency-bound on hand-offs between the
// two threads that shared a row (one named barrier per chunk pair, max / sum / extra-key exchange through
// shared memory) and on tcgen05.ld round trips nothing overlapped:
//   * 8 softmax warps instead of 16: group g = tile parity, one warp per TMEM lane quarter, ONE thread per
//     query row.  No pair barriers, no exchanges; the row max, row sum and the extra key's score are
//     thread-private registers.
//   * 448 threads per CTA -> 144 registers per thread: both passes keep the NEXT 32-column chunk in flight
//     (tcgen05.ld issued before the arithmetic of the current chunk, two register buffers) so the TMEM
//     latency hides behind the FFMA / MUFU work of the same warp.
//   * pass 1 uses three-input FMNMX3; a configurable share of the exponentials runs on the FMA pipe
//     (Cody-Waite split + cubic minimax polynomial, rel. error 7.5e-5 << the bf16 rounding of P) because
//     MUFU.EX2 (16 / clk / SM) is the pipe that bounds head_dim 64 attention.
//   * four extra-token warps (one per SM sub-partition) instead of two on shared sub-partitions.
// Warp roles: 0 TMA producer, 1 MMA issuer, 4..11 softmax, {2, 3, 12, 13} extra-token rows (T = 128k + 1).
// =====================================================================================================
constexpr int kThreadsV2 = 448;
constexpr int kSoftWarp0 = 4;
#ifndef CLM_ATTN_POLY
#define CLM_ATTN_POLY 0  // of every 4 exponentials, how many run on the FMA pipe (0..4)
#endif

__device__ __forceinline__ int tail_index_v2(int warp) {
  return warp == 2 ? 0 : (warp == 3 ? 1 : (warp == 12 ? 2 : (warp == 13 ? 3 : -1)));
}

// tcgen05.wait::ld that also "produces" the registers of the load it completes: nothing that reads them can
// be scheduled above it, although another load (into the other buffer) may already be in flight
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                 "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),
                 "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),
                 "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// max of 32 scores: FMNMX3, four chains
__device__ __forceinline__ float chunk_max3(const uint32_t (&v)[32], float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
    m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// 2^x for x <= 0 on the FMA / ALU pipes: x = j + f, j = round(x), f in [-0.5, 0.5]; 2^f by a cubic (minimax in
// relative error, 7.5e-5); the integer part goes straight into the exponent field.  The magic constant
// 1.5 * 2^23 leaves j in the low mantissa bits of r, so (bits(r) << 23) is j << 23 (mod 2^32).
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;
  const float f = x - (r - 12582912.0f);
  float p = fmaf(0.0551716648f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// p = 2^(s*c - max*c) for 32 scores -> 16 packed bf16x2 words; returns the chunk's sum of p.  Of every four
// elements the last kPoly use exp2_fma, the others MUFU.EX2.
template <bool kMasked, int kPoly>
__device__ __forceinline__ float chunk_exp_v2(const uint32_t (&v)[32], uint32_t (&pk)[16], float scale,
                                              float neg_mx, int base, int valid) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float x0 = fmaf(__uint_as_float(v[i]), scale, neg_mx);
    const float x1 = fmaf(__uint_as_float(v[i + 1]), scale, neg_mx);
    const float x2 = fmaf(__uint_as_float(v[i + 2]), scale, neg_mx);
    const float x3 = fmaf(__uint_as_float(v[i + 3]), scale, neg_mx);
    float e0 = (kPoly >= 4) ? exp2_fma(x0) : fast_exp2(x0);
    float e1 = (kPoly >= 3) ? exp2_fma(x1) : fast_exp2(x1);
    float e2 = (kPoly >= 2) ? exp2_fma(x2) : fast_exp2(x2);
    float e3 = (kPoly >= 1) ? exp2_fma(x3) : fast_exp2(x3);
    if (kMasked) {
      e0 = (base + i < valid) ? e0 : 0.f;
      e1 = (base + i + 1 < valid) ? e1 : 0.f;
      e2 = (base + i + 2 < valid) ? e2 : 0.f;
      e3 = (base + i + 3 < valid) ? e3 : 0.f;
    }
    s0 += e0; s1 += e1; s2 += e2; s3 += e3;
    pk[i / 2] = pack_bf16x2(e0, e1);
    pk[i / 2 + 1] = pack_bf16x2(e2, e3);
  }
  return (s0 + s1) + (s2 + s3);
}

template <bool kCausal>
__global__ void __launch_bounds__(kThreadsV2, 1)
attention_kernel_v2(const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map16,
                    const __grid_constant__ CUtensorMap map_out, __nv_bfloat16* __restrict__ out, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ostage = smem + p.stages * p.stage_bytes;  // 1024-byte aligned (stage_bytes is a multiple of 6144)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + (p.stage_out == 1 ? kOutStageBytes : 0));
  uint64_t* stage_full = bars;                     // [kMaxStages]
  uint64_t* stage_empty = bars + kMaxStages;       // [kMaxStages]
  uint64_t* s_full = bars + 2 * kMaxStages;        // [2] S ready (MMA commit), by tile parity
  uint64_t* p_full = s_full + 2;                   // [2] P written (128 softmax threads)
  uint64_t* o_full = s_full + 4;                   // [2] O ready (MMA commit)
  uint64_t* slot_free = s_full + 6;                // [2] O drained (128 softmax threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);
  // [4] "an item owned by extra-token warp j has landed": the MMA thread, which sees every stage fill in order,
  // arrives.  The extra-token warps cannot wait on stage_full themselves: each takes every fourth item, so it
  // would skip phases of a stage's barrier and a parity wait cannot tell phase k from phase k - 2.
  uint64_t* tail_go = s_full + 9;
  // extra-token scratch (only when p.xt): per softmax warp a copy of the extra V row (64 bf16), then one
  // fp32 probability row per extra-token warp
  uint8_t* vxs = reinterpret_cast<uint8_t*>(bars) + 256;      // [8 warps][128 B]
  float* prow = reinterpret_cast<float*>(vxs + 8 * 128);      // [4 warps][288]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T, H = p.H, Tp = p.Tp, Tk = p.Tk, D = p.H * kHeadDim;
  const int kv_bytes = Tp * 128;  // one of Q / K / V in a stage (Tp rows are loaded; the MMAs see Tk keys)
  const int n_local = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  const int n_tiles = n_local * p.mtiles;
  const int tail_w = tail_index_v2(warp);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map64);
    tma_prefetch_desc(&map16);
    if (p.stage_out) tma_prefetch_desc(&map_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&stage_full[s], 1);
      // last PV's commit (+ the extra-token warp that owns the item) (+ the 4 softmax warps of each of the
      // item's tiles once their output slab, staged in the tile's dead Q rows, has been read by the TMA unit)
      mbar_init(&stage_empty[s], 1 + p.xt + (p.stage_out == 2 ? 4 * p.mtiles : 0));
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&o_full[s], 1);
      mbar_init(&slot_free[s], 128);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&tail_go[s], 1);
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
#ifdef CLM_ATTN_TRACE
      int tl = 0;
#endif
      for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
        const int b = it / H, h = it % H;
        const int row_base = b * T;
        TRACE(tl, 0);
        mbar_wait(&stage_empty[st], ph ^ 1);
        TRACE(tl, 1);
#ifdef CLM_ATTN_TRACE
        ++tl;
#endif
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
    // ================= MMA issuer (two independent streams, one per tile parity) =================
    const uint32_t idesc_pv = umma_idesc_bf16(128, kHeadDim, 0, 1);
    struct Cursor {  // position of a stream inside the CTA's tile list
      int t, mt, st, li;  // tile, tile inside its item, stage, item (CTA-local sequence number)
      uint32_t ph;
      __device__ void init(int t0, const AttnParams& p) {
        t = t0; mt = t0; st = 0; ph = 0; li = 0;
        while (mt >= p.mtiles) { mt -= p.mtiles; bump(p); }
      }
      __device__ void bump(const AttnParams& p) {
        ++li;
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      __device__ void advance(int step, const AttnParams& p) {
        t += step; mt += step;
        while (mt >= p.mtiles) { mt -= p.mtiles; bump(p); }
      }
    };
    auto do_s = [&](const Cursor& c) {
      if (lane == 0) {
        const uint32_t q_addr = smem_u32(smem + c.st * p.stage_bytes);
        const uint32_t k_addr = q_addr + kv_bytes;
        const uint32_t sbase = tmem + static_cast<uint32_t>(p.s_col(c.t & 1));
        for (int n0 = 0; n0 < Tk; n0 += 256) {
          const int nn = (Tk - n0) < 256 ? (Tk - n0) : 256;
          const uint32_t idesc = umma_idesc_bf16(128, nn, 0, 0);
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_bf16_ss(sbase + n0, umma_desc_sw128(q_addr + c.mt * 16384 + k * 32, 1024),
                         umma_desc_sw128(k_addr + n0 * 128 + k * 32, 1024), idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[c.t & 1]);
        if (p.xt && c.mt == 0) mbar_arrive(&tail_go[c.li & 3]);  // the item's Q / K / V are in shared memory
      }
      __syncwarp();
    };
    int pv_cnt[kMaxStages];  // PVs issued per stage: the last one of an item releases its stage
#pragma unroll
    for (int i = 0; i < kMaxStages; ++i) pv_cnt[i] = 0;
    auto do_pv = [&](const Cursor& c) {
      // O(t) = P(t) V : P from TMEM (written by the softmax group), V MN-major from smem.  p_full(t)
      // also implies that the same group has drained O(t-2), whose columns this overwrites.
      const int b = c.t & 1;
      const bool last = (++pv_cnt[c.st] == p.mtiles);
      if (last) pv_cnt[c.st] = 0;
      if (lane == 0) {
        const uint32_t v_addr = smem_u32(smem + c.st * p.stage_bytes) + 2 * kv_bytes;
        const uint32_t pbase = tmem + static_cast<uint32_t>(p.s_col(b));
        const uint32_t obase = tmem + static_cast<uint32_t>(p.o_col(b));
        const int ksteps = Tk / 16;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_bf16_ts(obase, pbase + ks * 8, umma_desc_sw128_mn(v_addr + ks * 2048), idesc_pv,
                       ks != 0 ? 1u : 0u);
        umma_commit(&o_full[b]);
        if (last) umma_commit(&stage_empty[c.st]);  // stage reusable once these MMAs retire
      }
      __syncwarp();
    };
    Cursor sc[2], pc[2];  // next S / next PV of each stream
    sc[0].init(0, p); sc[1].init(1, p); pc[0].init(0, p); pc[1].init(1, p);
    while (pc[0].t < n_tiles || pc[1].t < n_tiles) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        // PV(t): S(t) has been issued and the group has published P(t)
        // (test_wait, not try_wait: try_wait may put the thread to sleep on one stream's barrier while the
        // other stream's becomes ready)
        if (pc[b].t < sc[b].t && mbar_test_wait(&p_full[b], static_cast<uint32_t>((pc[b].t >> 1) & 1))) {
          tc_fence_after();
          TRACE(pc[b].t, 1);
          do_pv(pc[b]);
          pc[b].advance(2, p);
        }
        // S(t): PV(t-2) has been issued (in-order pipe: its P columns are safe); an aliased O(t-2)
        // has been drained; the item's Q/K/V have landed
        if (sc[b].t < n_tiles && sc[b].t - 2 < pc[b].t) {
          bool ok = true;
          if (p.o_alias(b) && sc[b].t >= 2)
            ok = mbar_test_wait(&slot_free[b], static_cast<uint32_t>(((sc[b].t - 2) >> 1) & 1));
          if (ok) ok = mbar_test_wait(&stage_full[sc[b].st], sc[b].ph);
          if (ok) {
            tc_fence_after();
            TRACE(sc[b].t, 0);
            do_s(sc[b]);
            sc[b].advance(2, p);
          }
        }
      }
    }
  } else if (warp >= kSoftWarp0 && warp < kSoftWarp0 + 8) {
    // ================= softmax: one thread per query row =================
    const int g = (warp - kSoftWarp0) >> 2;  // group = parity of the tiles it owns
    const int q = warp & 3;                  // TMEM lane quarter of this warp (warp % 4)
    const int r = q * 32 + lane;             // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int nchunks = (Tk + 31) / 32;
    constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t srow = tmem + lane_off + static_cast<uint32_t>(p.s_col(g));
    const uint32_t orow = tmem + lane_off + static_cast<uint32_t>(p.o_col(g));
    const bool xt = !kCausal && p.xt != 0;
    const uint32_t vx_w = smem_u32(vxs + (warp - kSoftWarp0) * 128);  // this warp's copy of the extra V row
    int t = 0;
    int li = 0;  // items seen by this CTA: item li sits in stage li % stages
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ++li)
    for (int mt = 0; mt < p.mtiles; ++mt, ++t) {
      if ((t & 1) != g) continue;
      const int b = it / H, h = it - b * H;
      const uint32_t par = static_cast<uint32_t>((t >> 1) & 1);
      const int qi = mt * 128 + r;                     // query position in the sequence
      const bool warp_live = mt * 128 + q * 32 < T;    // does this warp own any real row?
      int valid = T < Tk ? T : Tk;
      if (kCausal) valid = (qi + 1 < T) ? qi + 1 : T;
      const int st_cur = li % p.stages;
      // ---- extra key (row Tk of K / V): q_row . k_x on the CUDA cores, and a private copy of v_x (the
      // stage may be refilled before the output phase of the item's last tile)
      float sx = 0.f, px = 0.f;
      if (xt) {
        mbar_wait(&stage_full[st_cur], static_cast<uint32_t>((li / p.stages) & 1));
        TRACE(t, 7);
        const uint32_t sb = smem_u32(smem + st_cur * p.stage_bytes);
        const uint32_t qa = sb + static_cast<uint32_t>(qi) * 128u;
        const uint32_t ka = sb + static_cast<uint32_t>(kv_bytes + Tk * 128);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 a = ld_shared_v4(qa + ((static_cast<uint32_t>(c) ^ static_cast<uint32_t>(r & 7)) << 4));
          const uint4 kx = ld_shared_v4(ka + (static_cast<uint32_t>(c) << 4));
          dot8(a, kx, d0, d1);
        }
        sx = d0 + d1;
        if (lane < 8) {
          const uint4 w = ld_shared_v4(sb + static_cast<uint32_t>(2 * kv_bytes + Tk * 128 + (lane << 4)));
          st_shared_v4(vx_w + (lane << 4), w.x, w.y, w.z, w.w);
        }
        __syncwarp();
      }
      // chunks this warp has to look at: under the causal mask nothing right of its last row counts
      int nch = nchunks;
      if (kCausal) {
        const int wv = (mt * 128 + q * 32 + 32 < T) ? mt * 128 + q * 32 + 32 : T;
        nch = (wv + 31) / 32;
      }

      TRACE(t, 0);
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      TRACE(t, 1);
      uint32_t va[32], vb[32];
      // ---- pass 1: row max; the load of chunk c+1 is in flight while chunk c is reduced
      float mx = -INFINITY;
      if (warp_live) {
        tmem_ld_32x32b_x32(srow, va);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait_dep(va);
          if (c + 1 < nch) tmem_ld_32x32b_x32(srow + (c + 1) * 32, vb);
          mx = (c * 32 + 32 <= valid) ? chunk_max3(va, mx) : chunk_max_masked(va, mx, c * 32, valid);
          if (c + 1 < nch) {
            tmem_ld_wait_dep(vb);
            if (c + 2 < nch) tmem_ld_32x32b_x32(srow + (c + 2) * 32, va);
            mx = (c * 32 + 64 <= valid) ? chunk_max3(vb, mx) : chunk_max_masked(vb, mx, c * 32 + 32, valid);
          }
        }
      }
      if (xt) mx = fmaxf(mx, sx);
      TRACE(t, 2);
      if (p.serial && t >= 1) mbar_wait(&p_full[g ^ 1], static_cast<uint32_t>(((t - 1) >> 1) & 1));
      TRACE(t, 3);
      // ---- pass 2: p = 2^(s*c - max*c), row sum, P (bf16x2) written over S.  P(c) lands in the columns
      // [16c, 16c + 16) of S, i.e. inside chunk c / 2, which this thread has already consumed; the chunk in
      // flight (c + 1) lies further right.
      float sum = 0.f;
      if (warp_live) {
        const float neg_mx = -mx * kScaleLog2e;
        if (xt) {
          px = fast_exp2(fmaf(sx, kScaleLog2e, neg_mx));  // fp32 weight of the extra key (not rounded to bf16)
          sum = px;
        }
        uint32_t pk[16];
        tmem_ld_32x32b_x32(srow, va);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait_dep(va);
          if (c + 1 < nch) tmem_ld_32x32b_x32(srow + (c + 1) * 32, vb);
          sum += (c * 32 + 32 <= valid)
                     ? chunk_exp_v2<false, CLM_ATTN_POLY>(va, pk, kScaleLog2e, neg_mx, c * 32, valid)
                     : chunk_exp_v2<true, 0>(va, pk, kScaleLog2e, neg_mx, c * 32, valid);
          tmem_st_32x32b_x16(srow + c * 16, pk);
          if (c + 1 < nch) {
            tmem_ld_wait_dep(vb);
            if (c + 2 < nch) tmem_ld_32x32b_x32(srow + (c + 2) * 32, va);
            sum += (c * 32 + 64 <= valid)
                       ? chunk_exp_v2<false, CLM_ATTN_POLY>(vb, pk, kScaleLog2e, neg_mx, c * 32 + 32, valid)
                       : chunk_exp_v2<true, 0>(vb, pk, kScaleLog2e, neg_mx, c * 32 + 32, valid);
            tmem_st_32x32b_x16(srow + (c + 1) * 16, pk);
          }
        }
        if (kCausal && nch < nchunks) {  // P right of the causal frontier is zero (PV reads it)
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
          for (int c = nch; c < nchunks; ++c) tmem_st_32x32b_x16(srow + c * 16, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      TRACE(t, 4);
      mbar_arrive(&p_full[g]);

      // ---- O(t): out of TMEM, + p_x v_x, / row sum, bf16, out through shared memory and a TMA tile store
      mbar_wait(&o_full[g], par);
      tc_fence_after();
      TRACE(t, 5);
      if (warp_live) {
        tmem_ld_32x32b_x32(orow, va);
        tmem_ld_32x32b_x32(orow + 32, vb);
        tmem_ld_wait_dep(va);
        tmem_ld_wait_dep(vb);
      }
      tc_fence_before();
      TRACE(t, 6);
      mbar_arrive(&slot_free[g]);  // O(t) is in registers: its columns may be overwritten
      if (xt && warp_live) {       // O += p_x * v_x
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint4 w = ld_shared_v4(vx_w + (jj << 4));
          uint32_t* o = (jj < 4) ? &va[8 * jj] : &vb[8 * (jj - 4)];
          o[0] = __float_as_uint(fmaf(px, bf16_lo(w.x), __uint_as_float(o[0])));
          o[1] = __float_as_uint(fmaf(px, bf16_hi(w.x), __uint_as_float(o[1])));
          o[2] = __float_as_uint(fmaf(px, bf16_lo(w.y), __uint_as_float(o[2])));
          o[3] = __float_as_uint(fmaf(px, bf16_hi(w.y), __uint_as_float(o[3])));
          o[4] = __float_as_uint(fmaf(px, bf16_lo(w.z), __uint_as_float(o[4])));
          o[5] = __float_as_uint(fmaf(px, bf16_hi(w.z), __uint_as_float(o[5])));
          o[6] = __float_as_uint(fmaf(px, bf16_lo(w.w), __uint_as_float(o[6])));
          o[7] = __float_as_uint(fmaf(px, bf16_hi(w.w), __uint_as_float(o[7])));
        }
      }
      const float inv = 1.0f / sum;
      if (p.stage_out) {
        // One lane owns one 128-byte output row.  The warp assembles its 32-row slab in shared memory in the
        // SWIZZLE_128B pattern of the output map and one lane hands it to the TMA unit; rows >= T are clipped.
        // stage_out == 2: no room for a staging buffer; the slab goes into this warp's 32 rows of the tile's
        // own Q block, which nothing reads after S has been computed (same swizzled row layout).
        const uint32_t ost = (p.stage_out == 2)
                                 ? smem_u32(smem + st_cur * p.stage_bytes) + static_cast<uint32_t>(mt * 128 + q * 32) * 128u
                                 : smem_u32(ostage + (g * 4 + q) * 4096);
        if (warp_live) {
          if (p.stage_out == 1) {  // the previous tile's slab has left the staging buffer
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint32_t* o = (jj < 4) ? &va[8 * jj] : &vb[8 * (jj - 4)];
            st_shared_v4(ost + static_cast<uint32_t>(lane) * 128u + ((static_cast<uint32_t>(jj) ^ (lane & 7)) << 4),
                         pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv),
                         pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv),
                         pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv),
                         pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_out, ost, h * kHeadDim, mt * 128 + q * 32, b);
            bulk_commit();
          }
        }
        if (p.stage_out == 2 && lane == 0) {
          // release this warp's share of the stage as soon as the TMA unit has read the slab: the refill of
          // the stage (two stages only) is on the critical path of the tile after next
          bulk_wait_read<0>();
          mbar_arrive(&stage_empty[st_cur]);
        }
        __syncwarp();
      } else if (warp_live && qi < T) {
        uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(b * T + qi) * D + h * kHeadDim);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t* o = (jj < 4) ? &va[8 * jj] : &vb[8 * (jj - 4)];
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          o4[jj] = w;
        }
      }
    }
    if (p.stage_out && lane == 0) bulk_wait<0>();  // every slab of this warp has reached memory
  } else if (tail_w >= 0 && !kCausal && p.xt) {
    // ================= extra-token warps: the query row Tk of every item, on the CUDA cores =================
    // (one row per (batch, head): a third 128-row tensor-core tile would be 1/128 used).  The four warps take
    // every fourth item; see tail_row.
    float* my_prow = prow + tail_w * 288;
    int st = 0, tl = 0;
    uint32_t own_ph = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ++tl) {
      if ((tl & 3) == tail_w) {
        const int b = it / H, h = it - b * H;
        TRACE(tl, 0);
        mbar_wait(&tail_go[tail_w], own_ph);
        own_ph ^= 1;
        TRACE(tl, 1);
        const uint32_t sb = smem_u32(smem + st * p.stage_bytes);
        uint32_t* orow = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * T + Tk) * D + h * kHeadDim);
        if (Tk == 256) tail_row<8>(sb, kv_bytes, my_prow, lane, orow);
        else tail_row<4>(sb, kv_bytes, my_prow, lane, orow);
        __syncwarp();  // every lane is done with the stage and with its p row
        TRACE(tl, 3);
        if (lane == 0) mbar_arrive(&stage_empty[st]);
      }
      if (++st == p.stages) st = 0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}


// =====================================================================================================
// attention_kernel_v3 — key-blocked, single-pass softmax (vision towers: 128 < keys <= 256, non-causal).
//
// attention_kernel keeps a whole score row (up to 256 columns) per 128-query tile in TMEM and reads it
// twice (row max, then exponentials); its two streams are serial chains S -> max -> exp -> P V -> drain ->
// store of ~11-13 k clocks each, so a tile leaves an SM every ~6.5 k clocks although MUFU.EX2 would allow
// one per ~2 k.  Here a tile is processed as TWO key blocks of <= 128 keys:
//   * TMEM per stream (tile parity) g: buffer A = columns [256g, 256g+128), buffer B = [256g+128, 256g+256).
//     S(block 0) -> A, S(block 1) -> B, both issued back to back; P(block) overwrites the first half of its
//     own buffer; the O accumulator lives in the upper half of A (dead once block 0's scores are in
//     registers).  512 columns = 2 streams x 2 buffers, no spare accumulator needed.
//   * softmax: one thread per query row, 4 warps per stream.  A block's 128 scores are loaded ONCE into
//     registers (four tcgen05.ld.x32 in flight together), reduced to the block max, exponentiated and written
//     back as bf16 P.  The warpgroups' register budget is raised with setmaxnreg (the TMA / MMA / extra-token
//     warps give theirs up).
//   * online softmax across the two blocks with a lazy rescale: block 1 is exponentiated relative to the
//     block-0 maximum m unless some row's block-1 maximum exceeds m by more than 8 (log2 units: P <= 256 is
//     harmless in bf16 / fp32); only then O = P0 V0 is rescaled in TMEM (tcgen05.ld -> x 2^(m - m') ->
//     tcgen05.st) before P1 V1 accumulates.  Softmax is shift invariant, so the result is the same function.
//   * P0 V0 runs on the tensor pipe while the softmax warps work on block 1, S(block 1) while they work on
//     block 0: the per-stream chain shrinks to exp(block 0) + exp(block 1) + P1 V1 + drain.
// Warp roles (16 warps): 0 TMA producer, 1 MMA issuer, 4..7 / 8..11 softmax of stream 0 / 1,
// {2, 3, 12, 13} extra-token rows (T = 128k + 1), 14, 15 idle (they complete the fourth warpgroup).
// =====================================================================================================
constexpr int kThreadsV3 = 512;
constexpr int kRegsSoftmaxV3 = 176;  // setmaxnreg targets: 2 x 128 x 176 + 2 x 128 x 80 = 65536 registers
constexpr int kRegsOtherV3 = 80;
#ifndef CLM_ATTN_POLY3
#define CLM_ATTN_POLY3 0  // of every 4 exponentials, how many run on the FMA pipe in attention_kernel_v3
#endif

template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

__global__ void __launch_bounds__(kThreadsV3, 1)
attention_kernel_v3(const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map16,
                    const __grid_constant__ CUtensorMap map_out, __nv_bfloat16* __restrict__ out, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ostage = smem + p.stages * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + (p.stage_out == 1 ? kOutStageBytes : 0));
  uint64_t* stage_full = bars;                     // [kMaxStages]
  uint64_t* stage_empty = bars + kMaxStages;       // [kMaxStages]
  uint64_t* s_full0 = bars + 2 * kMaxStages;       // [2] S(block 0) ready, by stream; one completion per tile
  uint64_t* s_full1 = s_full0 + 2;                 // [2] S(block 1) ready
  uint64_t* p_full0 = s_full0 + 4;                 // [2] P(block 0) written (128 softmax threads)
  uint64_t* p_full1 = s_full0 + 6;                 // [2] P(block 1) written
  uint64_t* pv0_done = s_full0 + 8;                // [2] O = P0 V0 complete (only waited for before a rescale)
  uint64_t* o_full = s_full0 + 10;                 // [2] O complete
  uint64_t* slot_free = s_full0 + 12;              // [2] O drained (128 softmax threads)
  uint64_t* tail_go = s_full0 + 14;                // [4] item of extra-token warp j has landed (MMA thread arrives)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full0 + 18);
  uint8_t* vxs = reinterpret_cast<uint8_t*>(bars) + 256;      // [8 warps][128 B] copies of the extra V row
  float* prow = reinterpret_cast<float*>(vxs + 8 * 128);      // [4 warps][288] extra-token probability rows

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T, H = p.H, Tp = p.Tp, Tk = p.Tk, D = p.H * kHeadDim;
  const int kv_bytes = Tp * 128;
  const int nk1 = Tk - 128;                        // keys of block 1 the MMA computes (multiple of 16)
  const int valid1 = (T < Tk ? T : Tk) - 128;      // ... of which real keys
  const int n_local = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);
  const int n_tiles = n_local * p.mtiles;
  const int tail_w = tail_index_v2(warp);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map64);
    tma_prefetch_desc(&map16);
    if (p.stage_out) tma_prefetch_desc(&map_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&stage_full[s], 1);
      mbar_init(&stage_empty[s], 1 + p.xt + (p.stage_out == 2 ? 4 * p.mtiles : 0));
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full0[s], 1);
      mbar_init(&s_full1[s], 1);
      mbar_init(&p_full0[s], 128);
      mbar_init(&p_full1[s], 128);
      mbar_init(&pv0_done[s], 1);
      mbar_init(&o_full[s], 1);
      mbar_init(&slot_free[s], 128);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&tail_go[s], 1);
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
  // register budgets: the two softmax warpgroups (warps 4..11) take what the others give up.  setmaxnreg is the
  // first statement of each warpgroup's branch so that ptxas allocates that branch against the new budget.
  if (warp < 4 || warp >= 12) {
  setmaxnreg_dec<kRegsOtherV3>();
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
    // ================= MMA issuer: two independent streams (tile parity), three steps per tile =================
    //   step 0: stage landed, O(t-2) drained      -> S0 = Q K0^T into buffer A, S1 = Q K1^T into buffer B
    //   step 1: P0 published                      -> O  = P0 V0
    //   step 2: P1 published                      -> O += P1 V1  (the last one of an item releases its stage)
    const uint32_t idesc_pv = umma_idesc_bf16(128, kHeadDim, 0, 1);
    const uint32_t idesc_s0 = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_s1 = umma_idesc_bf16(128, nk1, 0, 0);
    struct Cursor {
      int t, mt, st, li;
      uint32_t ph;   // parity of the stage's current fill
      uint32_t tph;  // parity of this stream's per-tile barriers (one completion per tile)
      __device__ void init(int t0, const AttnParams& p) {
        t = t0; mt = t0; st = 0; ph = 0; li = 0; tph = 0;
        while (mt >= p.mtiles) { mt -= p.mtiles; bump(p); }
      }
      __device__ void bump(const AttnParams& p) {
        ++li;
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      __device__ void advance(const AttnParams& p) {
        t += 2; mt += 2; tph ^= 1;
        while (mt >= p.mtiles) { mt -= p.mtiles; bump(p); }
      }
    };
    int pv_cnt[kMaxStages];
#pragma unroll
    for (int i = 0; i < kMaxStages; ++i) pv_cnt[i] = 0;
    Cursor cur[2];
    int step[2] = {0, 0};
    cur[0].init(0, p); cur[1].init(1, p);
    while (cur[0].t < n_tiles || cur[1].t < n_tiles) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        Cursor& c = cur[g];
        if (c.t >= n_tiles) continue;
        const uint32_t buf_a = tmem + static_cast<uint32_t>(g * 256);
        const uint32_t buf_b = buf_a + 128u;
        const uint32_t sbase = smem_u32(smem + c.st * p.stage_bytes);
        if (step[g] == 0) {
          bool ok = mbar_try_wait(&stage_full[c.st], c.ph);
          if (ok && c.t >= 2) ok = mbar_try_wait(&slot_free[g], c.tph ^ 1);  // O(t-2) left buffer A
          if (ok) {
            tc_fence_after();
            if (lane == 0) {
              const uint32_t q_addr = sbase + static_cast<uint32_t>(c.mt) * 16384u;
              const uint32_t k_addr = sbase + static_cast<uint32_t>(kv_bytes);
#pragma unroll
              for (int k = 0; k < kHeadDim / 16; ++k)
                umma_bf16_ss(buf_a, umma_desc_sw128(q_addr + k * 32, 1024), umma_desc_sw128(k_addr + k * 32, 1024),
                             idesc_s0, k != 0 ? 1u : 0u);
              umma_commit(&s_full0[g]);
#pragma unroll
              for (int k = 0; k < kHeadDim / 16; ++k)
                umma_bf16_ss(buf_b, umma_desc_sw128(q_addr + k * 32, 1024),
                             umma_desc_sw128(k_addr + 128 * 128 + k * 32, 1024), idesc_s1, k != 0 ? 1u : 0u);
              umma_commit(&s_full1[g]);
              if (p.xt && c.mt == 0) mbar_arrive(&tail_go[c.li & 3]);
            }
            __syncwarp();
            step[g] = 1;
          }
        } else if (step[g] == 1) {
          if (mbar_try_wait(&p_full0[g], c.tph)) {
            tc_fence_after();
            if (lane == 0) {
              const uint32_t v_addr = sbase + static_cast<uint32_t>(2 * kv_bytes);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                umma_bf16_ts(buf_a + 64u, buf_a + static_cast<uint32_t>(ks * 8), umma_desc_sw128_mn(v_addr + ks * 2048),
                             idesc_pv, ks != 0 ? 1u : 0u);
              umma_commit(&pv0_done[g]);
            }
            __syncwarp();
            step[g] = 2;
          }
        } else {
          if (mbar_try_wait(&p_full1[g], c.tph)) {
            tc_fence_after();
            const bool last = (++pv_cnt[c.st] == p.mtiles);
            if (last) pv_cnt[c.st] = 0;
            if (lane == 0) {
              const uint32_t v_addr = sbase + static_cast<uint32_t>(2 * kv_bytes) + 8u * 2048u;
              for (int ks = 0; ks < nk1 / 16; ++ks)
                umma_bf16_ts(buf_a + 64u, buf_b + static_cast<uint32_t>(ks * 8), umma_desc_sw128_mn(v_addr + ks * 2048),
                             idesc_pv, 1u);
              umma_commit(&o_full[g]);
              if (last) umma_commit(&stage_empty[c.st]);
            }
            __syncwarp();
            c.advance(p);
            step[g] = 0;
          }
        }
      }
    }
  } else if (tail_w >= 0 && p.xt) {
    // ================= extra-token warps (see attention_kernel_v2) =================
    float* my_prow = prow + tail_w * 288;
    int st = 0, tl = 0;
    uint32_t own_ph = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ++tl) {
      if ((tl & 3) == tail_w) {
        const int b = it / H, h = it - b * H;
        mbar_wait(&tail_go[tail_w], own_ph);
        own_ph ^= 1;
        const uint32_t sb = smem_u32(smem + st * p.stage_bytes);
        uint32_t* orow_g = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * T + Tk) * D + h * kHeadDim);
        if (Tk == 256) tail_row<8>(sb, kv_bytes, my_prow, lane, orow_g);
        else tail_row<4>(sb, kv_bytes, my_prow, lane, orow_g);
        __syncwarp();
        if (lane == 0) mbar_arrive(&stage_empty[st]);
      }
      if (++st == p.stages) st = 0;
    }
  }
  } else {
    setmaxnreg_inc<kRegsSoftmaxV3>();
    // ================= softmax: one thread per query row, one key block (<= 128 scores) in registers =================
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    constexpr float kScaleLog2e = 0.125f * 1.4426950408889634f;
    constexpr float kLazy = 8.0f;  // block 1 keeps the block-0 maximum unless it is exceeded by more than this (log2)
    const uint32_t buf_a = tmem + lane_off + static_cast<uint32_t>(g * 256);
    const uint32_t buf_b = buf_a + 128u;
    const uint32_t orow = buf_a + 64u;
    const bool xt = p.xt != 0;
    const uint32_t vx_w = smem_u32(vxs + (warp - 4) * 128);
    const int nch1 = (nk1 + 31) / 32;  // 32-column chunks of block 1 (3 or 4)
    uint32_t tph = 0;                  // parity of this stream's per-tile barriers
    int t = 0, li = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x, ++li)
    for (int mt = 0; mt < p.mtiles; ++mt, ++t) {
      if ((t & 1) != g) continue;
      const int b = it / H, h = it - b * H;
      const int qi = mt * 128 + r;
      const bool warp_live = mt * 128 + q * 32 < T;
      const int st_cur = li % p.stages;
      float sx = 0.f, px = 0.f;
      if (xt) {  // extra key (row Tk of K / V): q_row . k_x on the CUDA cores, private copy of v_x
        mbar_wait(&stage_full[st_cur], static_cast<uint32_t>((li / p.stages) & 1));
        const uint32_t sb = smem_u32(smem + st_cur * p.stage_bytes);
        const uint32_t qa = sb + static_cast<uint32_t>(qi) * 128u;
        const uint32_t ka = sb + static_cast<uint32_t>(kv_bytes + Tk * 128);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 a = ld_shared_v4(qa + ((static_cast<uint32_t>(c) ^ static_cast<uint32_t>(r & 7)) << 4));
          const uint4 kx = ld_shared_v4(ka + (static_cast<uint32_t>(c) << 4));
          dot8(a, kx, d0, d1);
        }
        sx = d0 + d1;
        if (lane < 8) {
          const uint4 w = ld_shared_v4(sb + static_cast<uint32_t>(2 * kv_bytes + Tk * 128 + (lane << 4)));
          st_shared_v4(vx_w + (lane << 4), w.x, w.y, w.z, w.w);
        }
        __syncwarp();
      }

      uint32_t v0[32], v1[32], v2[32], v3[32], pk[16];
      float mrow = -INFINITY, sum = 0.f;
      // ---- block 0: 128 real keys
      mbar_wait(&s_full0[g], tph);
      tc_fence_after();
      if (warp_live) {
        tmem_ld_32x32b_x32(buf_a, v0);
        tmem_ld_32x32b_x32(buf_a + 32, v1);
        tmem_ld_32x32b_x32(buf_a + 64, v2);
        tmem_ld_32x32b_x32(buf_a + 96, v3);
        tmem_ld_wait_dep(v0); tmem_ld_wait_dep(v1); tmem_ld_wait_dep(v2); tmem_ld_wait_dep(v3);
        mrow = chunk_max3(v3, chunk_max3(v2, chunk_max3(v1, chunk_max3(v0, -INFINITY))));
        if (xt) mrow = fmaxf(mrow, sx);
        const float neg_mx = -mrow * kScaleLog2e;
        if (xt) {
          px = fast_exp2(fmaf(sx, kScaleLog2e, neg_mx));
          sum = px;
        }
        sum += chunk_exp_v2<false, CLM_ATTN_POLY3>(v0, pk, kScaleLog2e, neg_mx, 0, 128);
        tmem_st_32x32b_x16(buf_a, pk);
        sum += chunk_exp_v2<false, CLM_ATTN_POLY3>(v1, pk, kScaleLog2e, neg_mx, 32, 128);
        tmem_st_32x32b_x16(buf_a + 16, pk);
        sum += chunk_exp_v2<false, CLM_ATTN_POLY3>(v2, pk, kScaleLog2e, neg_mx, 64, 128);
        tmem_st_32x32b_x16(buf_a + 32, pk);
        sum += chunk_exp_v2<false, CLM_ATTN_POLY3>(v3, pk, kScaleLog2e, neg_mx, 96, 128);
        tmem_st_32x32b_x16(buf_a + 48, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&p_full0[g]);

      // ---- block 1: nk1 computed keys, valid1 of them real
      mbar_wait(&s_full1[g], tph);
      tc_fence_after();
      if (warp_live) {
        tmem_ld_32x32b_x32(buf_b, v0);
        tmem_ld_32x32b_x32(buf_b + 32, v1);
        tmem_ld_32x32b_x32(buf_b + 64, v2);
        if (nch1 > 3) tmem_ld_32x32b_x32(buf_b + 96, v3);
        tmem_ld_wait_dep(v0); tmem_ld_wait_dep(v1); tmem_ld_wait_dep(v2); tmem_ld_wait_dep(v3);
        float m1 = -INFINITY;
        m1 = (32 <= valid1) ? chunk_max3(v0, m1) : chunk_max_masked(v0, m1, 0, valid1);
        m1 = (64 <= valid1) ? chunk_max3(v1, m1) : chunk_max_masked(v1, m1, 32, valid1);
        m1 = (96 <= valid1) ? chunk_max3(v2, m1) : chunk_max_masked(v2, m1, 64, valid1);
        if (nch1 > 3) m1 = (128 <= valid1) ? chunk_max3(v3, m1) : chunk_max_masked(v3, m1, 96, valid1);
        // lazy rescale: only when some row's maximum moved up by more than kLazy
        const bool grow = (m1 - mrow) * kScaleLog2e > kLazy;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? m1 : mrow;
          const float alpha = fast_exp2((mrow - m_new) * kScaleLog2e);  // 1 for rows that keep their maximum
          mbar_wait(&pv0_done[g], tph);
          tc_fence_after();
          uint32_t o[32];
#pragma unroll 1
          for (int hc = 0; hc < 2; ++hc) {
            tmem_ld_32x32b_x32(orow + hc * 32, o);
            tmem_ld_wait_dep(o);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(orow + hc * 32, o);
          }
          tmem_st_wait();
          sum *= alpha;
          px *= alpha;
          mrow = m_new;
        }
        const float neg_mx = -mrow * kScaleLog2e;
        sum += (32 <= valid1) ? chunk_exp_v2<false, CLM_ATTN_POLY3>(v0, pk, kScaleLog2e, neg_mx, 0, valid1)
                              : chunk_exp_v2<true, 0>(v0, pk, kScaleLog2e, neg_mx, 0, valid1);
        tmem_st_32x32b_x16(buf_b, pk);
        sum += (64 <= valid1) ? chunk_exp_v2<false, CLM_ATTN_POLY3>(v1, pk, kScaleLog2e, neg_mx, 32, valid1)
                              : chunk_exp_v2<true, 0>(v1, pk, kScaleLog2e, neg_mx, 32, valid1);
        tmem_st_32x32b_x16(buf_b + 16, pk);
        if (nk1 > 64) {  // keys 64.. of block 1 exist (the MMA reads nk1 / 2 packed columns of P)
          sum += (96 <= valid1) ? chunk_exp_v2<false, CLM_ATTN_POLY3>(v2, pk, kScaleLog2e, neg_mx, 64, valid1)
                                : chunk_exp_v2<true, 0>(v2, pk, kScaleLog2e, neg_mx, 64, valid1);
          tmem_st_32x32b_x16(buf_b + 32, pk);
        }
        if (nk1 > 96) {
          sum += (128 <= valid1) ? chunk_exp_v2<false, CLM_ATTN_POLY3>(v3, pk, kScaleLog2e, neg_mx, 96, valid1)
                                 : chunk_exp_v2<true, 0>(v3, pk, kScaleLog2e, neg_mx, 96, valid1);
          tmem_st_32x32b_x16(buf_b + 48, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&p_full1[g]);

      // ---- O: out of TMEM, + p_x v_x, / row sum, bf16, out through shared memory and a TMA tile store
      mbar_wait(&o_full[g], tph);
      tc_fence_after();
      if (warp_live) {
        tmem_ld_32x32b_x32(orow, v0);
        tmem_ld_32x32b_x32(orow + 32, v1);
        tmem_ld_wait_dep(v0);
        tmem_ld_wait_dep(v1);
      }
      tc_fence_before();
      mbar_arrive(&slot_free[g]);
      if (xt && warp_live) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint4 w = ld_shared_v4(vx_w + (jj << 4));
          uint32_t* o = (jj < 4) ? &v0[8 * jj] : &v1[8 * (jj - 4)];
          o[0] = __float_as_uint(fmaf(px, bf16_lo(w.x), __uint_as_float(o[0])));
          o[1] = __float_as_uint(fmaf(px, bf16_hi(w.x), __uint_as_float(o[1])));
          o[2] = __float_as_uint(fmaf(px, bf16_lo(w.y), __uint_as_float(o[2])));
          o[3] = __float_as_uint(fmaf(px, bf16_hi(w.y), __uint_as_float(o[3])));
          o[4] = __float_as_uint(fmaf(px, bf16_lo(w.z), __uint_as_float(o[4])));
          o[5] = __float_as_uint(fmaf(px, bf16_hi(w.z), __uint_as_float(o[5])));
          o[6] = __float_as_uint(fmaf(px, bf16_lo(w.w), __uint_as_float(o[6])));
          o[7] = __float_as_uint(fmaf(px, bf16_hi(w.w), __uint_as_float(o[7])));
        }
      }
      const float inv = 1.0f / sum;
      if (p.stage_out) {
        const uint32_t ost = (p.stage_out == 2)
                                 ? smem_u32(smem + st_cur * p.stage_bytes) + static_cast<uint32_t>(mt * 128 + q * 32) * 128u
                                 : smem_u32(ostage + (g * 4 + q) * 4096);
        if (warp_live) {
          if (p.stage_out == 1) {  // the previous tile's slab has left the staging buffer
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint32_t* o = (jj < 4) ? &v0[8 * jj] : &v1[8 * (jj - 4)];
            st_shared_v4(ost + static_cast<uint32_t>(lane) * 128u + ((static_cast<uint32_t>(jj) ^ (lane & 7)) << 4),
                         pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv),
                         pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv),
                         pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv),
                         pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_out, ost, h * kHeadDim, mt * 128 + q * 32, b);
            bulk_commit();
          }
        }
        if (p.stage_out == 2 && lane == 0) {
          bulk_wait_read<0>();
          mbar_arrive(&stage_empty[st_cur]);
        }
        __syncwarp();
      } else if (warp_live && qi < T) {
        uint4* o4 = reinterpret_cast<uint4*>(out + static_cast<size_t>(b * T + qi) * D + h * kHeadDim);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t* o = (jj < 4) ? &v0[8 * jj] : &v1[8 * (jj - 4)];
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          o4[jj] = w;
        }
      }
      tph ^= 1;
    }
    if (p.stage_out && lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

int clm_attention_launch(const void* qkv, void* out, int batch, int tokens, int heads, int causal,
                         cudaStream_t stream) {
  CLM_REQUIRE(qkv && out && batch >= 0 && tokens > 0 && heads > 0, "clm_attention: bad argument");
  CLM_REQUIRE(tokens <= 384, "clm_attention: tokens=%d > 384 unsupported (CLIP uses <= 257)", tokens);
  if (batch == 0) return CLM_OK;
  const int T = tokens;
  const int D = heads * kHeadDim;
  AttnParams p;
  p.T = T;
  p.H = heads;
  p.Tp = (T + 15) / 16 * 16;
  p.mtiles = (T + 127) / 128;
  // T = 128k + 1 (ViT-L/14: 256 patches + class token): the tensor cores take the 128k x 128k block, the
  // extra key and the extra query row are done on the CUDA cores (see AttnParams).  CLM_ATTN_XT=0 disables.
  static int xt_off = -1;
  if (xt_off < 0) {
    const char* e = getenv("CLM_ATTN_XT");
    xt_off = (e && e[0] == '0') ? 1 : 0;
  }
  p.xt = (!xt_off && !causal && T > 128 && T % 128 == 1 && T <= 257) ? 1 : 0;
  static int serial = -1;  // CLM_ATTN_SERIAL=1: the two softmax groups take turns on pass 2 (measured: no gain)
  if (serial < 0) {
    const char* e = getenv("CLM_ATTN_SERIAL");
    serial = (e && e[0] == '1') ? 1 : 0;
  }
  p.serial = serial;
  p.Tk = p.xt ? T - 1 : p.Tp;
  if (p.xt) p.mtiles = p.Tk / 128;
  const long long items = static_cast<long long>(batch) * heads;
  CLM_REQUIRE(items < 2147483647LL / 4, "clm_attention: too many (batch, head) items");
  p.num_items = static_cast<int>(items);
  // a stage holds Q, K, V ([Tp,64] bf16 each); the last query tile's 128-row UMMA window may run
  // past Q into K/V (finite values, rows never stored), so a stage is at least mtiles*16 KiB
  int stage_bytes = 3 * p.Tp * 128;
  if (stage_bytes < p.mtiles * 16384) stage_bytes = p.mtiles * 16384;
  p.stage_bytes = stage_bytes;
  // output staging for the TMA stores if two pipeline stages still fit beside it (not at T = 257)
  int smem_budget = 227 * 1024 - 1024 - 256 - kXchBytes - (p.xt ? kXtBytes : 0);
  p.stage_out = ((smem_budget - kOutStageBytes) / stage_bytes >= 2) ? 1 : 0;
  if (p.stage_out) smem_budget -= kOutStageBytes;
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  CLM_REQUIRE(stages >= 1, "clm_attention: tokens=%d needs %d bytes of shared memory per stage", T,
              stage_bytes);
  // one tile per item: each stream (tile parity) visits every other item, so with an odd ring it would skip
  // phases of a stage's barrier and a parity wait cannot tell phase k from phase k - 2
  if (p.mtiles == 1 && stages > 2 && (stages & 1)) stages -= 1;
  p.stages = stages;
  // TMEM plan (512 columns).  S needs round_up(Tp,32) fp32 columns (the softmax reads 32-column
  // chunks); P (bf16x2) reuses its first Tp/2; O needs 64.  Tiles alternate between two parities:
  //   two S regions + two O regions            (T <= 192)
  //   two S regions, O0 separate, O1 inside S1 (T <= 224: ViT-B/16's 197) -> S(t+2) of odd t waits
  //   two key blocks per tile, two S regions of one block + two O regions (T <= 384: ViT-L/14's 257)
//   [fallback: one S region, two O regions, tiles strictly in sequence]
  p.blocks = 1; p.nb0 = p.Tk; p.nb1 = 0;
  const int s_cols = (p.Tk + 31) / 32 * 32;
  const int o_in = (p.Tk / 2 + 31) / 32 * 32;  // first column past P inside an S region
  if (p.xt && p.Tk > 128) {
    // Tk = 256: two S regions fill the 512 columns; both O accumulators live in the dead upper half of
    // their own S region, so S(t+2) of either stream waits for the drain of O(t)
    CLM_REQUIRE(stages >= 2 && 2 * s_cols <= 512 && o_in + 64 <= s_cols, "clm_attention: extra-token plan does not fit");
    p.nslots = 2;
    p.s_col0 = 0; p.s_col1 = s_cols;
    p.o_col0 = o_in; p.o_col1 = s_cols + o_in;
    p.o_alias0 = 1; p.o_alias1 = 1;
    if (!p.stage_out) p.stage_out = 2;  // output slabs are staged in the tile's dead Q rows
  } else if (2 * s_cols + 128 <= 512 && stages >= 2) {
    p.nslots = 2;
    p.s_col0 = 0; p.s_col1 = s_cols;
    p.o_col0 = 2 * s_cols; p.o_col1 = 2 * s_cols + 64;
    p.o_alias0 = 0; p.o_alias1 = 0;
  } else if (2 * s_cols + 64 <= 512 && o_in + 64 <= s_cols && stages >= 2) {
    p.nslots = 2;
    p.s_col0 = 0; p.s_col1 = s_cols;
    p.o_col0 = 2 * s_cols; p.o_col1 = s_cols + o_in;
    p.o_alias0 = 0; p.o_alias1 = 1;
  } else {
    // the score row does not fit twice: two key blocks per tile with an online-softmax rescale of O,
    // so that two S regions (one per stream) fit again.  CLM_ATTN_BLOCKS=1 keeps the one-region plan.
    static int one_block = -1;
    if (one_block < 0) {
      const char* e = getenv("CLM_ATTN_BLOCKS");
      one_block = (e && e[0] == '1') ? 1 : 0;
    }
    const int nb0 = (p.Tk / 2 + 15) / 16 * 16, nb1 = p.Tk - nb0;
    const int b_cols = (nb0 + 31) / 32 * 32;
    if (!one_block && !causal && stages >= 2 && nb1 >= 16 && 2 * b_cols + 128 <= 512) {
      p.blocks = 2; p.nb0 = nb0; p.nb1 = nb1;
      p.nslots = 2;
      p.s_col0 = 0; p.s_col1 = b_cols;
      p.o_col0 = 2 * b_cols; p.o_col1 = 2 * b_cols + 64;
      p.o_alias0 = 0; p.o_alias1 = 0;
    } else {
      CLM_REQUIRE(s_cols + 128 <= 512, "clm_attention: tokens=%d needs %d TMEM columns", T, s_cols + 128);
      p.nslots = 1;
      p.s_col0 = 0; p.s_col1 = 0;
      p.o_col0 = s_cols; p.o_col1 = s_cols + 64;
      p.o_alias0 = 0; p.o_alias1 = 0;
    }
  }
  if (p.blocks == 2) p.stage_out = 0;  // the two-block path (T > 224) keeps per-thread stores
  const int smem_bytes = stages * stage_bytes + (p.stage_out == 1 ? kOutStageBytes : 0) + 256 + kXchBytes +
                         (p.xt ? kXtBytes : 0) + 1024;

  CUtensorMap map64, map16;
  int rc = clm_make_tmap_bf16_2d(&map64, qkv, static_cast<uint64_t>(batch) * T, 3ull * D, 3ull * D,
                                 kHeadDim, 64);
  if (rc) return rc;
  rc = clm_make_tmap_bf16_2d(&map16, qkv, static_cast<uint64_t>(batch) * T, 3ull * D, 3ull * D,
                             kHeadDim, 16);
  if (rc) return rc;
  CUtensorMap map_out = map64;
  if (p.stage_out) {
    rc = clm_make_tmap_bf16_3d(&map_out, out, static_cast<uint64_t>(D), static_cast<uint64_t>(T),
                               static_cast<uint64_t>(batch), 2ull * D, 2ull * D * T, kHeadDim, 32);
    if (rc) return rc;
  }
  const int grid = p.num_items < clm_num_sms() ? p.num_items : clm_num_sms();
  // algorithmic work: QK^T and PV at the true T (causal not discounted)
  ProfScope prof(CLM_K_ATTENTION, 4.0 * batch * heads * static_cast<double>(T) * T * kHeadDim,
                 2.0 * batch * T * 4.0 * D, stream);
  // CLM_ATTN_V3=1 selects the key-blocked single-pass kernel (attention_kernel_v3) for the vision shapes
  // (non-causal, 128 < keys <= 256, at least two pipeline stages).  Measured 20-25 % SLOWER than the whole-row
  // kernel (profiles/r2_attention_notes.md): with M = 128 an MMA costs >= 133 clocks whatever N is (the A tile
  // is fetched at ~32 B/clk, tools/microbench/umma_rate.cu), so two N = 128 score blocks cost more tensor time
  // than one N = 256 row, and every extra hand-off adds a ~500-clock commit -> mbarrier round trip.  Kept for
  // A/B measurements; it is exercised by the same unit tests.
  static int use_v3 = -1;
  if (use_v3 < 0) {
    const char* e = getenv("CLM_ATTN_V3");
    use_v3 = (e && e[0] == '1') ? 1 : 0;
  }
  if (use_v3 && !causal && p.Tk > 128 && p.Tk <= 256 && stages >= 2) {
    // the whole-row TMEM plans above may have switched the output staging off (two-block plan) or asked for
    // the in-place variant only for the extra-token shape: redo the choice for this kernel
    int so = ((227 * 1024 - 1024 - 256 - kXchBytes - (p.xt ? kXtBytes : 0) - kOutStageBytes) / stage_bytes >= 2) ? 1 : 0;
    if (!so && p.xt) so = 2;
    p.stage_out = so;
    const int smem_v3 = stages * stage_bytes + (so == 1 ? kOutStageBytes : 0) + 256 + kXchBytes + (p.xt ? kXtBytes : 0) + 1024;
    CUtensorMap map_out3 = map64;
    if (so) {
      rc = clm_make_tmap_bf16_3d(&map_out3, out, static_cast<uint64_t>(D), static_cast<uint64_t>(T),
                                 static_cast<uint64_t>(batch), 2ull * D, 2ull * D * T, kHeadDim, 32);
      if (rc) return rc;
    }
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_v3));
    attention_kernel_v3<<<grid, kThreadsV3, smem_v3, stream>>>(map64, map16, map_out3,
                                                                static_cast<__nv_bfloat16*>(out), p);
    CLM_CUDA_CHECK(cudaGetLastError());
    return CLM_OK;
  }
  // CLM_ATTN_V2=1 selects the one-thread-per-row kernel (attention_kernel_v2) for the plans with a single key
  // block and two S regions.  Measured (profiles/r2_attention_notes.md) it is 5-25 % SLOWER than the
  // two-threads-per-row kernel: each stream is a serial chain S -> max -> exp -> P V -> drain -> store, and
  // halving the threads per tile lengthens the chain more than the removed pair barriers shorten it.  Kept for
  // A/B measurements only.
  static int use_v2 = -1;
  if (use_v2 < 0) {
    const char* e = getenv("CLM_ATTN_V2");
    use_v2 = (e && e[0] == '1') ? 1 : 0;
  }
  if (use_v2 && p.blocks == 1 && p.nslots == 2) {
    if (causal) {
      CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel_v2<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      attention_kernel_v2<true><<<grid, kThreadsV2, smem_bytes, stream>>>(
          map64, map16, map_out, static_cast<__nv_bfloat16*>(out), p);
    } else {
      CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel_v2<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      attention_kernel_v2<false><<<grid, kThreadsV2, smem_bytes, stream>>>(
          map64, map16, map_out, static_cast<__nv_bfloat16*>(out), p);
    }
    CLM_CUDA_CHECK(cudaGetLastError());
    return CLM_OK;
  }
  if (causal) {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attention_kernel<true><<<grid, kThreads, smem_bytes, stream>>>(
        map64, map16, map_out, static_cast<__nv_bfloat16*>(out), p);
  } else {
    CLM_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attention_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(
        map64, map16, map_out, static_cast<__nv_bfloat16*>(out), p);
  }
  CLM_CUDA_CHECK(cudaGetLastError());
  return CLM_OK;
}

#ifdef CLM_ATTN_TRACE
extern "C" int clm_attention_set_trace(void* dev_buf) {
  unsigned long long* ptr = static_cast<unsigned long long*>(dev_buf);
  CLM_CUDA_CHECK(cudaMemcpyToSymbol(g_trace, &ptr, sizeof(ptr)));
  return CLM_OK;
}
#endif

extern "C" int clm_attention(const void* qkv_bf16, void* out_bf16, int batch, int tokens, int heads,
                             int causal, void* stream) {
  return clm_attention_launch(qkv_bf16, out_bf16, batch, tokens, heads, causal,
                              static_cast<cudaStream_t>(stream));
}
