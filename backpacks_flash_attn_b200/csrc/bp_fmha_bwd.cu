// FlashAttention backward for sm_100a: dQ, dK, dV of softmax(scale * Q K^T [+ causal mask]) V.
//
// Replaces the reference's FlashAttention-1 backward (csrc/flash_attn/fmha_api.cpp:338-500 mha_bwd ->
// src/fmha_dgrad_kernel_1xN_loop.h), whose single kernel walks K/V blocks in the outer loop and accumulates dQ
// through an fp32 `dq_tmp` round trip in HBM.  Written for Blackwell from scratch:
//
//   * three launches, all deterministic (no atomics, fixed accumulation order):
//       1. bwd_stats_kernel: per query row -lse * log2(e) and -delta = -sum_d dO * O into a padded workspace, blocked by 64
//          rows ([64 x lse][64 x delta] per block: one 512-byte bulk copy per step; rows beyond a sequence get
//          -lse = -inf, so their probabilities are exactly 0);
//       2. fmha_bwd_kernel<kDQ = false>: a CTA owns 128 KEYS of one (batch, head) and streams the query blocks
//          that see them; dV and dK stay in TMEM for the whole sweep;
//       3. fmha_bwd_kernel<kDQ = true>:  a CTA owns 128 QUERIES and streams the key blocks they see; dQ stays in TMEM.
//     S and dP are recomputed in both kernels (7 tile products instead of 5) in exchange for no dQ read-modify-write
//     traffic, no conversion pass and bit-wise reproducible gradients.
//   * one kernel body for both modes.  With "stationary" tiles A1, A2 (128 rows) and "streaming" tiles B1, B2
//     (64 rows per step):   T1 = A1 B1^T,  T2 = A2 B2^T   (SS tcgen05.mma, all operands K-major as TMA lands them)
//         keys own   (A1, A2, B1, B2) = (K, V, Q, dO):  T1 = S^T, T2 = dP^T;  X1 = P^T, X2 = dS^T;
//                                                        dV += X1 B2,  dK += X2 B1
//         queries own (A1, A2, B1, B2) = (Q, dO, K, V): T1 = S,   T2 = dP;    X2 = dS;   dQ += X2 B1
//     X1 / X2 are written back to TMEM as 16-bit pairs OVER the columns of T1 / T2 they were computed from (TMEM lanes
//     are private to one thread, and tcgen05.mma of one thread execute in issue order, so the next step's T products
//     overwrite them only after this step's gradient products have read them) and feed TS tcgen05.mma whose B
//     operand is the SAME streaming tile consumed MN-major -- no transposes, nothing staged in shared memory.
//   * a step runs as two half-phases so that the tensor pipe and the exponentials overlap inside a CTA:
//         softmax threads:  [T1 -> P (exp2) -> X1]  x1_ready  [T2 -> dS -> X2]  x2_ready  [next T1 ...]
//         MMA thread:        ... T2(n) | x1_ready: acc1 += X1 B2, T1(n+1) | x2_ready: acc2 += X2 B1, T2(n+1) | ...
//     i.e. the next step's S is computed while this step's dS is formed, and this step's dP / dV product while its
//     exponentials run.  (The first version ran T1,T2 -> softmax -> both products strictly in sequence: ncu showed the
//     tensor pipe 23 % and MUFU 20 % active with every warp waiting on the chain's latency.)
//   * 256 TMEM columns and ~85 KB of shared memory per CTA at head dim 64, so TWO CTAs share an SM and fill each
//     other's remaining bubbles (CTA start-up, stationary loads, epilogue stores).
//   * the softmax scale is applied once to the fp32 dK / dQ accumulators, not per element.
//
// Varlen through cu_seqlens as in the forward; head dims that are a multiple of 8 up to 128 (TMA zero-fill to 64 / 128).
#include <cstddef>
#include <cstdlib>

#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace fmha_bwd {

constexpr int BS = 128;   // stationary rows per CTA (= TMEM lanes)
constexpr int BT = 64;    // streaming rows per step
constexpr int kStatsWords = 3 * BT;   // words per 64-row block of the statistics workspace
constexpr int kItemSlots = 4;         // shared-memory ring of decoded work items (producer -> the other roles)
constexpr float kLog2e = 1.4426950408889634f;
#ifndef BP_FMHA_BWD_POLY
#define BP_FMHA_BWD_POLY 2
#endif
constexpr int kPoly = BP_FMHA_BWD_POLY;   // of every 8 exponentials, how many run on the FMA pipe (see bp_common.cuh)

template <int DP, bool kDQ>
struct Cfg {
  static constexpr int kThreads = 320;   // warps 0-7: two softmax / epilogue warpgroups (TMEM lane quadrant = warp & 3,
                                         // column half = warp >> 2), warp 8: TMA producer, warp 9: MMA issuer
  static constexpr int kStages = (DP == 64) ? 3 : 2;
  static constexpr int kPanels = DP / 64;
  static constexpr uint32_t kStatPanelBytes = BS * 128;
  static constexpr uint32_t kStrPanelBytes = BT * 128;
  static constexpr uint32_t kStatTileBytes = kPanels * kStatPanelBytes;
  static constexpr uint32_t kStrTileBytes = kPanels * kStrPanelBytes;
  static constexpr uint32_t kStatsBytes = BT * 12;  // -lse * log2e, -delta and the dropout row word of the streaming rows
  static constexpr uint32_t kStageBytes = 2 * kStrTileBytes + 1024;
  static constexpr uint32_t offStat1 = 0;
  static constexpr uint32_t offStat2 = kStatTileBytes;
  static constexpr uint32_t offStr = 2 * kStatTileBytes;
  static constexpr uint32_t offStage = offStr + kStages * kStageBytes;   // [128 rows][DP * 2 B] epilogue staging tile
  static constexpr uint32_t offItems = offStage + BS * DP * 2;          // ring of decoded work items
  static constexpr uint32_t offBar = offItems + kItemSlots * 64;
  static constexpr uint32_t kSmemBytes = offBar + 256 + 1024;
  // TMEM columns.  Keys own: T1, T2, dV, dK (256 at head dim 64: full).  Queries own: T1, two T2 buffers, dQ -- the
  // spare columns double-buffer dP, so both products of step n+1 are issued while step n's dS is being formed.
  static constexpr int kT2Bufs = kDQ ? 2 : 1;
  static constexpr uint32_t colT1 = 0, colT2 = BT;
  static constexpr uint32_t colA1 = 2 * BT;
  static constexpr uint32_t colA2 = kDQ ? 3 * BT : 2 * BT + DP;
  static constexpr uint32_t kColsUsed = colA2 + DP;
  static constexpr uint32_t kTmemCols = kColsUsed <= 256 ? 256 : 512;
  static constexpr int kCtasPerSm = (kTmemCols == 256 && kSmemBytes <= 113 * 1024) ? 2 : 1;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct Params {
  const float* stats;        // (batch * nheads, s_pad / 64, 3, 64): -lse * log2e, -delta, dropout row word per block of 64
                             // query rows
  const uint32_t* drop_c;    // (batch * nheads, s_pad_k) dropout column words (null without dropout)
  int32_t s_pad_k;
  uint32_t drop_thr24;       // round(256 p) << 24
  float drop_scale;          // 1 / (1 - p)
  void* out1;                // keys-own: dV
  void* out2;                // keys-own: dK, queries-own: dQ (both scaled by the softmax scale)
  int64_t o1_row_stride, o1_head_stride, o2_row_stride, o2_head_stride;
  const int32_t* cu_q;
  const int32_t* cu_k;
  int32_t s_pad;
  int32_t batch, nheads, headdim;
  int32_t num_tiles;         // stationary tiles per sequence
  int32_t chunk_bh;          // (batch, head) pairs per scheduling chunk
  int32_t num_tickets;       // chunks * chunk_bh * num_tiles
  unsigned int* sched;       // this launch's ticket counter (zeroed in front of the kernel)
  int32_t is_causal;
  float scale, scale_log2;
  uint64_t* trace;           // debug: per-role event timestamps of CTA 0 (null in production)
};

struct Barriers {
  uint64_t stat_full, t1_full, t2_full[2], x1_ready, x2_ready[2], acc_full, acc_free;   // t2_full / x2_ready: per T2 buffer
  uint64_t str_full[3], str_empty[3];
  uint64_t item_full[kItemSlots], item_empty[kItemSlots];
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= 256, "barrier block");
#define BBAR(field) (bars_a + static_cast<uint32_t>(offsetof(Barriers, field)))
#define BBAR_I(field, i) (bars_a + static_cast<uint32_t>(offsetof(Barriers, field)) + 8u * static_cast<uint32_t>(i))

// packed fp32x2 arithmetic with distinct operands per half (the softmax warps are bound by instruction issue)
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{\n\t.reg .b64 a, b, c, d;\n\t"
      "mov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %4};\n\tmov.b64 c, {%5, %6};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}\n"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
// (d0, d1) = (e0, e1) * ((x0, x1) + (n0, n1))
__device__ __forceinline__ void mul_add2v(float& d0, float& d1, float e0, float e1, float x0, float x1, float n0, float n1) {
  asm("{\n\t.reg .b64 e, x, n, t, d;\n\t"
      "mov.b64 e, {%2, %3};\n\tmov.b64 x, {%4, %5};\n\tmov.b64 n, {%6, %7};\n\t"
      "add.rn.f32x2 t, x, n;\n\tmul.rn.f32x2 d, e, t;\n\tmov.b64 {%0, %1}, d;\n\t}\n"
      : "=f"(d0), "=f"(d1) : "f"(e0), "f"(e1), "f"(x0), "f"(x1), "f"(n0), "f"(n1));
}

// Optional spinning wait on test_wait for the MMA issuer's two hand-over waits (-DBP_FMHA_BWD_SPIN=1).  The suspending
// try_wait of mbar_wait_a wakes up a few hundred cycles after the phase completes; spinning shortens that but takes
// issue slots from the softmax warps of the same sub-partition.  Measured at config 2: 309.5 us spinning, 306.7 us
// suspending -- the suspending wait is the default.  Bounded like mbar_wait_a (traps after 4 s).
#ifndef BP_FMHA_BWD_SPIN
#define BP_FMHA_BWD_SPIN 0
#endif
__device__ __forceinline__ void mbar_spin_a(uint32_t bar, uint32_t parity) {
  if (!BP_FMHA_BWD_SPIN) {
    mbar_wait_a(bar, parity);
    return;
  }
  uint32_t polls = 0;
  uint64_t t0 = 0;
  while (!mbar_test_a(bar, parity)) {
    if ((++polls & 0xFFFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitTimeoutNs) __trap();
    }
  }
}

__device__ __forceinline__ void bulk_load_1d_w(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n"
      ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
      : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// launch 1: per-row statistics
// ------------------------------------------------------------------------------------------------------------------
template <bool kBF16>
__global__ void __launch_bounds__(256)
bwd_stats_kernel(const void* __restrict__ dout, const void* __restrict__ out, const float* __restrict__ lse,
                 float* __restrict__ stats, const int32_t* __restrict__ cu_q, int64_t do_row, int64_t do_head,
                 int64_t o_row, int64_t o_head, int32_t lse_stride, int32_t s_pad, int32_t nheads, int32_t headdim,
                 uint64_t seed) {
  // One CTA = one block of 64 query rows of one (batch, head).  16 lanes per row (8 elements = 16 bytes each), a warp
  // covers 8 rows in 4 passes whose loads are all issued before the first use.
  const int bh = blockIdx.x;
  const int b = bh / nheads, h = bh - b * nheads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & 15;
  const int q_begin = __ldg(cu_q + b);
  const int len_q = __ldg(cu_q + b + 1) - q_begin;
  uint4 a[4], c[4];
  int rows[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = warp * 8 + it * 2 + (lane >> 4);
    const int i = blockIdx.y * 64 + r;
    rows[it] = r;
    a[it] = c[it] = make_uint4(0u, 0u, 0u, 0u);
    if (i < len_q && q * 8 < headdim) {
      a[it] = __ldg(reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint16_t*>(dout) + static_cast<int64_t>(q_begin + i) * do_row + h * do_head + q * 8));
      c[it] = __ldg(reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint16_t*>(out) + static_cast<int64_t>(q_begin + i) * o_row + h * o_head + q * 8));
    }
  }
  float* blk = stats + (static_cast<int64_t>(bh) * (s_pad / 64) + blockIdx.y) * kStatsWords;
  const uint32_t dbase = drop_base(seed, static_cast<uint32_t>(bh));
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint32_t aw[4] = {a[it].x, a[it].y, a[it].z, a[it].w}, cw[4] = {c[it].x, c[it].y, c[it].z, c[it].w};
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 fa, fc;
      if constexpr (kBF16) {
        fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[k]));
        fc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&cw[k]));
      } else {
        fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[k]));
        fc = __half22float2(*reinterpret_cast<const __half2*>(&cw[k]));
      }
      acc = fmaf(fa.x, fc.x, acc);
      acc = fmaf(fa.y, fc.y, acc);
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (q == 0) {
      const int i = blockIdx.y * 64 + rows[it];
      const bool valid = i < len_q;
      blk[rows[it]] = valid ? -__ldg(lse + static_cast<int64_t>(bh) * lse_stride + i) * kLog2e : -INFINITY;
      blk[64 + rows[it]] = valid ? -acc : 0.f;
      blk[128 + rows[it]] = __uint_as_float(drop_row_word(dbase, static_cast<uint32_t>(i)));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// launches 2 and 3
// ------------------------------------------------------------------------------------------------------------------
// One work item = one stationary tile (128 rows of one (batch, head)).  Tickets are handed out DYNAMICALLY: the items
// are numbered chunk by chunk (chunk_bh (batch, head) pairs x num_tiles tiles, sized to one wave of 2 CTAs per SM, so
// the CTAs that run together sweep the same K/V/Q/dO in L2), heaviest tiles of a chunk first; every CTA starts with
// ticket blockIdx.x and its producer warp draws the following ones from a per-launch counter (zeroed by a memset in
// front of the kernel) and publishes the decoded item to the other warp roles through a small shared-memory ring.
// (The first persistent version used a static mirrored schedule: with 10.4 chunks at config 2 the CTAs ended between
// 136 and 176 us.)  Which CTA computes a tile does not affect its result: the gradients stay bit-wise reproducible.
struct Item {
  int valid;
  int bh, batch, head, row0;
  int stat_begin, len_stat, str_begin, len_str;
  int first, n_steps;
  int end;   // sentinel: no more work for this CTA
};

__device__ __forceinline__ void put_item(uint32_t a, const Item& it) {
  sts128(a, it.valid, it.bh, it.batch, it.head);
  sts128(a + 16, it.row0, it.stat_begin, it.len_stat, it.str_begin);
  sts128(a + 32, it.len_str, it.first, it.n_steps, it.end);
}
__device__ __forceinline__ Item get_item(uint32_t a) {
  const uint4 x = lds128(a), y = lds128(a + 16), z = lds128(a + 32);
  Item it;
  it.valid = x.x; it.bh = x.y; it.batch = x.z; it.head = x.w;
  it.row0 = y.x; it.stat_begin = y.y; it.len_stat = y.z; it.str_begin = y.w;
  it.len_str = z.x; it.first = z.y; it.n_steps = z.z; it.end = z.w;
  return it;
}

template <bool kDQ>
__device__ __forceinline__ Item decode_item(const Params& p, int chunk, int pos) {
  Item it;
  it.valid = 0;
  it.end = 0;
  it.bh = it.batch = it.head = it.row0 = it.stat_begin = it.len_stat = it.str_begin = it.len_str = it.first = it.n_steps = 0;
  const int total_bh = p.batch * p.nheads;
  const int bh0 = chunk * p.chunk_bh;
  const int gc = min(p.chunk_bh, total_bh - bh0);
  int rank = pos / gc;
  if (rank >= p.num_tiles) return it;
  it.bh = bh0 + (pos - rank * gc);
  it.batch = it.bh / p.nheads;
  it.head = it.bh - it.batch * p.nheads;
  const int tile = (kDQ && p.is_causal) ? p.num_tiles - 1 - rank : rank;   // rank 0 = most steps
  it.row0 = tile * BS;
  const int q_begin = __ldg(p.cu_q + it.batch), len_q = __ldg(p.cu_q + it.batch + 1) - q_begin;
  const int k_begin = __ldg(p.cu_k + it.batch), len_k = __ldg(p.cu_k + it.batch + 1) - k_begin;
  it.stat_begin = kDQ ? q_begin : k_begin;
  it.len_stat = kDQ ? len_q : len_k;
  it.str_begin = kDQ ? k_begin : q_begin;
  it.len_str = kDQ ? len_k : len_q;
  if (it.row0 >= it.len_stat) return it;
  const int nblk = (it.len_str + BT - 1) / BT;
  int first = 0, last = nblk;
  if (p.is_causal) {
    if (kDQ) last = min(nblk, (it.row0 + BS) / BT);   // keys <= last query of the tile
    else first = it.row0 / BT;                        // queries >= first key of the tile
  }
  it.first = first;
  it.n_steps = max(0, last - first);
  it.valid = 1;
  return it;
}

template <int DP, bool kDQ, bool kBF16, bool kDrop>
__global__ void __launch_bounds__(Cfg<DP, kDQ>::kThreads, Cfg<DP, kDQ>::kCtasPerSm)
fmha_bwd_kernel(const __grid_constant__ CUtensorMap tmStat1, const __grid_constant__ CUtensorMap tmStat2,
                const __grid_constant__ CUtensorMap tmStr1, const __grid_constant__ CUtensorMap tmStr2,
                const __grid_constant__ CUtensorMap tmOut1, const __grid_constant__ CUtensorMap tmOut2, const Params p) {
  using C = Cfg<DP, kDQ>;
#ifdef BP_TRACE
  // debug: (start ns, end ns, SM id) of every CTA behind the role timelines (benchmarks/trace_kernel.py bwd_*)
  uint64_t cta_t0 = 0;
  if (p.trace && threadIdx.x == 0 && blockIdx.x < 4096) cta_t0 = global_timer_ns();
#endif
  const int per_chunk = p.chunk_bh * p.num_tiles;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_a = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars_a = smem_a + C::offBar;
  uint8_t* smem = smem_raw + (smem_a - smem_u32(smem_raw));
  Barriers& bars = *reinterpret_cast<Barriers*>(smem + C::offBar);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmStat1);
    tma_prefetch_desc(&tmStat2);
    tma_prefetch_desc(&tmStr1);
    tma_prefetch_desc(&tmStr2);
    if constexpr (!kDQ) tma_prefetch_desc(&tmOut1);
    tma_prefetch_desc(&tmOut2);
    mbar_init(&bars.stat_full, 1);
    mbar_init(&bars.t1_full, 1);
    mbar_init(&bars.t2_full[0], 1);
    mbar_init(&bars.t2_full[1], 1);
    mbar_init(&bars.x1_ready, 256);
    mbar_init(&bars.x2_ready[0], 256);
    mbar_init(&bars.x2_ready[1], 256);
    mbar_init(&bars.acc_full, 1);
    mbar_init(&bars.acc_free, 256);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars.str_full[i], 1);
      mbar_init(&bars.str_empty[i], 1);
    }
    for (int i = 0; i < kItemSlots; ++i) {
      mbar_init(&bars.item_full[i], 1);
      mbar_init(&bars.item_empty[i], 1 + 8);   // one lane of every consumer warp (issuer, 8 softmax warps)
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(&bars.tmem_base, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  // every consumer role walks the item ring with its own sequence number
  const uint32_t items_a = smem_a + C::offItems;
  uint32_t item_seq = 0;
  auto next_item = [&]() {
    const uint32_t islot = item_seq % kItemSlots;
    mbar_wait_a(BBAR_I(item_full, islot), (item_seq / kItemSlots) & 1);
    const Item it = get_item(items_a + islot * 64);
    __syncwarp();
    if (lane == 0) mbar_arrive_a(BBAR_I(item_empty, islot));
    ++item_seq;
    return it;
  };

  // Barrier phases run on across items: `ic` counts this CTA's items that have steps (stat_full, acc_full, acc_free
  // complete once per such item), `sc` counts steps (t1_full, t2_full, x1_ready, x2_ready), slot / ph walk the ring.
  if (warp == 8) {
    // ===================== TMA producer (whole warp walks the loop, an elected lane issues) =====================
    uint32_t slot = 0, ph = 1;   // ph: parity of the str_empty phase to wait for (first pass: the slots are free)
    uint32_t blk = 0;            // streaming blocks loaded so far
    uint32_t ic = 0;
    Tracer tr(p.trace, 0, blockIdx.x == 0 && lane == 0);
    uint32_t seq = 0;            // items published (valid items + the end marker)
    auto publish = [&](const Item& pit) {
      const uint32_t islot = seq % kItemSlots;
      if (seq >= kItemSlots) mbar_wait_a(BBAR_I(item_empty, islot), ((seq / kItemSlots) - 1) & 1);
      if (lane == 0) {
        put_item(items_a + islot * 64, pit);
        mbar_arrive_a(BBAR_I(item_full, islot));
      }
      ++seq;
    };
    auto draw = [&]() {   // next ticket (lane 0 draws, the warp shares it)
      unsigned int d = 0;
      if (lane == 0) d = gridDim.x + atomicAdd(p.sched, 1u);
      return static_cast<int>(__shfl_sync(0xffffffffu, d, 0));
    };
    auto fetch_valid = [&](int& w) {   // the item of ticket w, or of the next ticket that decodes to a valid tile
      while (w < p.num_tickets) {
        const Item c = decode_item<kDQ>(p, w / per_chunk, w % per_chunk);
        if (c.valid) return c;
        w = draw();
      }
      Item fin = decode_item<kDQ>(p, 0, 0);
      fin.valid = 0;
      fin.end = 1;
      return fin;
    };
    int w = blockIdx.x;   // the first ticket is static
    Item cur = fetch_valid(w);
    publish(cur);
    while (!cur.end) {
      const Item it = cur;
      int w_next = draw();   // drawn early: the counter's round trip hides behind this item's loads
      if (it.n_steps == 0) {   // nothing to load (the softmax warps write zeros)
        cur = fetch_valid(w_next);
        publish(cur);
        continue;
      }
      // the stationary tiles of the previous item are free once all of its products have completed
      if (ic > 0) mbar_wait_a(BBAR(acc_full), (ic - 1) & 1);
      mbar_arrive_expect_tx_w(BBAR(stat_full), 2 * C::kStatTileBytes);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int pn = 0; pn < C::kPanels; ++pn) {
          const uint32_t off = pn * C::kStatPanelBytes + half * (BT * 128);
          tma_load_3d_w(smem_a + C::offStat1 + off, &tmStat1, BBAR(stat_full), pn * 64, it.head,
                        it.stat_begin + it.row0 + half * BT);
          tma_load_3d_w(smem_a + C::offStat2 + off, &tmStat2, BBAR(stat_full), pn * 64, it.head,
                        it.stat_begin + it.row0 + half * BT);
        }
      for (int n = 0; n < it.n_steps; ++n, ++blk) {
        if (blk >= C::kStages) mbar_wait_a(BBAR_I(str_empty, slot), ph);
        tr.rec(1, blk);
        const uint32_t st = smem_a + C::offStr + slot * C::kStageBytes;
        const int srow = it.str_begin + (it.first + n) * BT;
        mbar_arrive_expect_tx_w(BBAR_I(str_full, slot), 2 * C::kStrTileBytes + (kDQ ? 0u : C::kStatsBytes));
#pragma unroll
        for (int pn = 0; pn < C::kPanels; ++pn) {
          tma_load_3d_w(st + pn * C::kStrPanelBytes, &tmStr1, BBAR_I(str_full, slot), pn * 64, it.head, srow);
          tma_load_3d_w(st + C::kStrTileBytes + pn * C::kStrPanelBytes, &tmStr2, BBAR_I(str_full, slot), pn * 64,
                        it.head, srow);
        }
        if constexpr (!kDQ)
          bulk_load_1d_w(st + 2 * C::kStrTileBytes,
                         p.stats + (static_cast<int64_t>(it.bh) * (p.s_pad / 64) + (it.first + n)) * kStatsWords,
                         C::kStatsBytes, BBAR_I(str_full, slot));
        if (++slot == C::kStages) {
          slot = 0;
          ph ^= 1;
        }
      }
      ++ic;
      cur = fetch_valid(w_next);
      publish(cur);
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_t = make_idesc(kBF16, BS, BT, false, false);
    constexpr uint32_t idesc_acc = make_idesc(kBF16, BS, DP, false, true);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tT1 = tm + C::colT1, tT2 = tm + C::colT2, tA1 = tm + C::colA1, tA2 = tm + C::colA2;
    const uint32_t sA1 = smem_a + C::offStat1, sA2 = smem_a + C::offStat2;
    auto issue_t1 = [&](uint32_t sB1) {
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk) {
        const uint32_t a = sA1 + (kk >> 2) * C::kStatPanelBytes + (kk & 3) * 32;
        const uint32_t b = sB1 + (kk >> 2) * C::kStrPanelBytes + (kk & 3) * 32;
        umma_ss_w(tT1, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc_t, kk > 0 ? 1u : 0u);
      }
      umma_commit_w(BBAR(t1_full));
    };
    auto issue_t2 = [&](uint32_t sB2, uint32_t buf) {
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk) {
        const uint32_t a = sA2 + (kk >> 2) * C::kStatPanelBytes + (kk & 3) * 32;
        const uint32_t b = sB2 + (kk >> 2) * C::kStrPanelBytes + (kk & 3) * 32;
        umma_ss_w(tT2 + buf * BT, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc_t,
                  kk > 0 ? 1u : 0u);
      }
      umma_commit_w(BBAR_I(t2_full, buf));
    };
    uint32_t slot = 0, ph = 0;
    uint32_t sc = 0, ic = 0;
    Tracer tr(p.trace, 1, blockIdx.x == 0 && lane == 0);
    while (true) {
      const Item it = next_item();
      if (it.end) break;
      if (it.n_steps == 0) continue;
      const int n_steps = it.n_steps;
      tr.rec(0, ic);
      mbar_wait_a(BBAR(stat_full), ic & 1);
      mbar_wait_a(BBAR_I(str_full, slot), ph);
      tc_fence_after();
      tr.rec(9, ic);
      // (T1 / T2 are free: their last readers, the previous item's final products, were issued by this thread)
      issue_t1(smem_a + C::offStr + slot * C::kStageBytes);
      issue_t2(smem_a + C::offStr + slot * C::kStageBytes + C::kStrTileBytes, kDQ ? (sc & 1u) : 0u);
      for (int n = 0; n < n_steps; ++n, ++sc) {
        const uint32_t sB1 = smem_a + C::offStr + slot * C::kStageBytes;
        const uint32_t sB2 = sB1 + C::kStrTileBytes;
        uint32_t nslot = slot + 1, nph = ph;
        if (nslot == C::kStages) {
          nslot = 0;
          nph ^= 1;
        }
        const uint32_t nB1 = smem_a + C::offStr + nslot * C::kStageBytes;
        const bool more = n + 1 < n_steps;
        // the next streaming tiles have normally landed long ago: probe now, the latency overlaps the hand-over wait
        const bool next_ready = more ? mbar_test_a(BBAR_I(str_full, nslot), nph) : true;
        // first half: X1 is in TMEM (keys own) / T1 has been read out (queries own)
        mbar_spin_a(BBAR(x1_ready), sc & 1);
        // the accumulators of the previous item must have been read out before the first products overwrite them
        if (n == 0 && ic > 0) mbar_wait_a(BBAR(acc_free), (ic - 1) & 1);
        tc_fence_after();
        tr.rec(1, sc);
        // gradient products: A = X from TMEM (8 columns per K-step of 16 streaming rows; the two column halves of X sit
        // at the start of the two halves of T, where their warpgroups wrote them), B = the streaming tile MN-major
        if constexpr (!kDQ) {
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk)
            umma_ts_w(tA1, tT1 + (kk >> 1) * 32 + (kk & 1) * 8,
                      make_smem_desc_sw128(sB2 + kk * 16 * 128, C::kStrPanelBytes, 1024), idesc_acc,
                      (n > 0 || kk > 0) ? 1u : 0u);
        }
        if (more) {
          if (!next_ready) {
            mbar_wait_a(BBAR_I(str_full, nslot), nph);
            tc_fence_after();
          }
          issue_t1(nB1);   // executes behind the product above, which is the last reader of X1 (same issuing thread)
          // queries own: dP of the next step goes to the other T2 buffer, whose last reader (the dQ product of the
          // previous step) was issued by this thread before
          if constexpr (kDQ) issue_t2(nB1 + C::kStrTileBytes, (sc + 1) & 1u);
        }
        tr.rec(2, sc);
        // second half: X2 is in TMEM
        // (queries own: one hand-over barrier per T2 buffer.  With dP of step n+1 available early, a fast softmax warp
        // can finish step n+1 before a slow one has finished step n; on a single barrier its arrival would be counted
        // towards step n and release this product too early)
        mbar_spin_a(BBAR_I(x2_ready, kDQ ? (sc & 1u) : 0u), kDQ ? ((sc >> 1) & 1u) : (sc & 1u));
        tc_fence_after();
        tr.rec(3, sc);
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_ts_w(tA2, tT2 + (kDQ ? (sc & 1u) * BT : 0u) + (kk >> 1) * 32 + (kk & 1) * 8,
                    make_smem_desc_sw128(sB1 + kk * 16 * 128, C::kStrPanelBytes, 1024), idesc_acc,
                    (n > 0 || kk > 0) ? 1u : 0u);
        umma_commit_w(BBAR_I(str_empty, slot));   // every product that reads this step's streaming tiles has been issued
        if constexpr (!kDQ) {
          if (more) issue_t2(nB1 + C::kStrTileBytes, 0u);
        }
        tr.rec(4, sc);
        slot = nslot;
        ph = nph;
      }
      umma_commit_w(BBAR(acc_full));
      ++ic;
    }
  } else {
    // ===================== softmax / gradient warps =====================
    // Two warpgroups share the 128 rows: thread (quadrant, lane) of warpgroup g owns row r and the 32 streaming columns
    // [32 g, 32 g + 32) of every step.  (One warpgroup with whole 64-column rows left 8 softmax warps per SM walking a
    // serial chain of ~2100 cycles per step; halving the per-thread work halves the chain.)
    constexpr int HC = BT / 2;
    const int g = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const int t256 = warp * 32 + lane;
    const int cbase = g * HC;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tT1 = tmem_base + lane_addr + C::colT1 + cbase, tT2 = tmem_base + lane_addr + C::colT2 + cbase;
    const float c2 = p.scale_log2;
    uint32_t slot = 0, ph = 0;
    uint32_t sc = 0, ic = 0;
    Tracer tr(p.trace, 2 + g, blockIdx.x == 0 && lane == 0 && (warp & 3) == 0);
    bool store_pending = false;   // (thread 0) a bulk store may still be reading the staging tile
    while (true) {
      const Item it = next_item();
      if (it.end) break;
      const int row0 = it.row0;
      const int row = row0 + r;
      const int len_str = it.len_str;
      const int n_steps = it.n_steps;
      float rowL = 0.f, rowD = 0.f;
      uint32_t my_word = 0;
      (void)my_word;
      if constexpr (kDQ) {
        const float* blk = p.stats + (static_cast<int64_t>(it.bh) * (p.s_pad / 64) + (row >> 6)) * kStatsWords;
        rowL = __ldg(blk + (row & 63));
        rowD = __ldg(blk + 64 + (row & 63));
        if constexpr (kDrop) my_word = __float_as_uint(__ldg(blk + 128 + (row & 63)));   // this query row's word
      } else if constexpr (kDrop) {
        my_word = __ldg(p.drop_c + static_cast<int64_t>(it.bh) * p.s_pad_k + row);       // this key's word
      }
      for (int n = 0; n < n_steps; ++n, ++sc) {
        tr.rec(0, sc);
        const int col0 = (it.first + n) * BT + cbase;   // first streaming row of this thread's columns
        const uint32_t sStats = smem_a + C::offStr + slot * C::kStageBytes + 2 * C::kStrTileBytes + cbase * 4;
        // valid columns of this row inside the thread's half block: lo <= i < hi
        int lo = 0, hi = len_str - col0;
        bool partial = col0 + HC > len_str;
        if (p.is_causal) {
          if (kDQ) {
            hi = min(hi, row + 1 - col0);
            partial = partial || (col0 + HC - 1 > row0);
          } else {
            lo = row - col0;
            partial = partial || (col0 < row0 + BS - 1);
          }
        }
        // the statistics travel with the streaming tiles, which have landed before T1 could be computed: a probe whose
        // latency overlaps the wait for T1 (a satisfied blocking wait costs ~200 cycles on its own)
        bool stats_ready = true;
        if constexpr (!kDQ) stats_ready = mbar_test_a(BBAR_I(str_full, slot), ph);

        // ---- first half: P = exp2(T1 * scale_log2 - L) ----
        mbar_wait_a(BBAR(t1_full), sc & 1);
        // dP: one barrier per T2 buffer (queries own: buffer sc & 1 completes every other step, so a late waiter can
        // never confuse the phase of step n with that of step n + 2)
        const uint32_t bar_t2 = BBAR_I(t2_full, kDQ ? (sc & 1u) : 0u);
        const uint32_t par_t2 = kDQ ? ((sc >> 1) & 1u) : (sc & 1u);
        const bool t2_ready = mbar_test_a(bar_t2, par_t2);   // consumed after the exponentials
        if constexpr (!kDQ) {
          if (!stats_ready) mbar_wait_a(BBAR_I(str_full, slot), ph);
        }
        tc_fence_after();
        tr.rec(1, sc);
        float e[HC];
        {
          uint32_t us[HC];
          tmem_ld32(tT1, us);
          tmem_ld_wait();
          tr.rec(2, sc);
          if constexpr (kDQ) {
            tc_fence_before();
            mbar_arrive_a(BBAR(x1_ready));   // T1 is in registers: the next S may overwrite it
            const float negL = rowL;   // the workspace holds -lse * log2e
#pragma unroll
            for (int i = 0; i < HC; i += 8) {
              float t8[8], e8[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) t8[k] = __uint_as_float(us[i + k]);
              exp2_scaled8<kPoly>(e8, t8, c2, negL);
#pragma unroll
              for (int k = 0; k < 8; ++k) e[i + k] = e8[k];
            }
          } else {
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
              const uint4 L = lds128(sStats + i * 4);   // -lse * log2e of four query columns (broadcast)
              float a0, a1, a2, a3;
              ffma2v(a0, a1, __uint_as_float(us[i]), __uint_as_float(us[i + 1]), c2, __uint_as_float(L.x),
                     __uint_as_float(L.y));
              ffma2v(a2, a3, __uint_as_float(us[i + 2]), __uint_as_float(us[i + 3]), c2, __uint_as_float(L.z),
                     __uint_as_float(L.w));
              a0 = fast_exp2(a0);
              a1 = fast_exp2(a1);
              if (kPoly == 4 || (kPoly == 2 && (i & 4))) {
                exp2_poly_pair(a2, a3);   // this share of the exponentials runs on the FMA pipe
              } else {
                a2 = fast_exp2(a2);
                a3 = fast_exp2(a3);
              }
              e[i] = a0;
              e[i + 1] = a1;
              e[i + 2] = a2;
              e[i + 3] = a3;
            }
          }
        }
        if (partial) {
#pragma unroll
          for (int i = 0; i < HC; ++i) e[i] = (i >= lo && i < hi) ? e[i] : 0.f;
        }
        tr.rec(3, sc);
        if constexpr (!kDQ) {
          uint32_t pk[HC / 2];
          if constexpr (kDrop) {
            // X1 = the DROPPED probabilities (dV = (D o P)^T dO / (1 - p)); e itself stays undropped for dS
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
              const uint4 R = lds128(sStats + 512 + i * 4);   // dropout words of four query columns
              pk[i / 2] = pack2<kBF16>(drop_keep(R.x, my_word, p.drop_thr24) ? e[i] : 0.f,
                                       drop_keep(R.y, my_word, p.drop_thr24) ? e[i + 1] : 0.f);
              pk[i / 2 + 1] = pack2<kBF16>(drop_keep(R.z, my_word, p.drop_thr24) ? e[i + 2] : 0.f,
                                           drop_keep(R.w, my_word, p.drop_thr24) ? e[i + 3] : 0.f);
            }
          } else {
#pragma unroll
            for (int i = 0; i < HC / 2; ++i) pk[i] = pack2<kBF16>(e[2 * i], e[2 * i + 1]);
          }
          tmem_st16(tT1, pk);   // over the first half of the columns this thread has just read
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive_a(BBAR(x1_ready));
        }
        tr.rec(4, sc);

        // ---- second half: dS = P * (T2 - delta) ----
        if (!t2_ready) mbar_wait_a(bar_t2, par_t2);
        tc_fence_after();
        tr.rec(5, sc);
        {
          uint32_t ud[HC];
          const uint32_t tT2n = tT2 + (kDQ ? (sc & 1u) * BT : 0u);
          tmem_ld32(tT2n, ud);
          tmem_ld_wait();
          if (partial) {
            // masked columns may hold products with rows of a neighbouring sequence: select, never multiply by zero
#pragma unroll
            for (int i = 0; i < HC; ++i) ud[i] = (i >= lo && i < hi) ? ud[i] : 0u;
          }
          if constexpr (kDrop) {
            // dP = D o (dO V^T) / (1 - p)
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
              uint4 W;
              if constexpr (kDQ) W = __ldg(reinterpret_cast<const uint4*>(p.drop_c + static_cast<int64_t>(it.bh) * p.s_pad_k + col0 + i));
              else W = lds128(sStats + 512 + i * 4);
              const uint32_t w[4] = {W.x, W.y, W.z, W.w};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ud[i + k] = drop_keep(kDQ ? my_word : w[k], kDQ ? w[k] : my_word, p.drop_thr24)
                                ? __float_as_uint(__uint_as_float(ud[i + k]) * p.drop_scale) : 0u;
            }
          }
          uint32_t pk[HC / 2];
          if constexpr (kDQ) {
#pragma unroll
            for (int i = 0; i < HC / 2; ++i) {
              float g0, g1;
              mul_add2v(g0, g1, e[2 * i], e[2 * i + 1], __uint_as_float(ud[2 * i]), __uint_as_float(ud[2 * i + 1]), rowD,
                        rowD);
              pk[i] = pack2<kBF16>(g0, g1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
              const uint4 D = lds128(sStats + 256 + i * 4);   // -delta of four query columns
              float g0, g1, g2, g3;
              mul_add2v(g0, g1, e[i], e[i + 1], __uint_as_float(ud[i]), __uint_as_float(ud[i + 1]),
                        __uint_as_float(D.x), __uint_as_float(D.y));
              mul_add2v(g2, g3, e[i + 2], e[i + 3], __uint_as_float(ud[i + 2]), __uint_as_float(ud[i + 3]),
                        __uint_as_float(D.z), __uint_as_float(D.w));
              pk[i / 2] = pack2<kBF16>(g0, g1);
              pk[i / 2 + 1] = pack2<kBF16>(g2, g3);
            }
          }
          tmem_st16(tT2n, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_a(BBAR_I(x2_ready, kDQ ? (sc & 1u) : 0u));
        tr.rec(7, sc);
        if (++slot == C::kStages) {
          slot = 0;
          ph ^= 1;
        }
      }

      // ---- epilogue: accumulators -> staging tile in shared memory -> global ----
      // Thread (row r, warpgroup g) owns columns [g DP/2, (g+1) DP/2) of its accumulator row; a thread-per-row store would
      // touch 32 different lines per instruction (the first version spent ~3700 cycles per tile on it).  The tile is
      // staged in the layout TMA expects ([64-column panel][128 rows][128 B], 16-byte chunks XOR-swizzled by the row) and
      // leaves as one bulk tensor store per panel; a tile that ends inside its sequence (the next sequence's rows follow
      // in memory) is written by all 256 threads with guarded 16-byte stores, every warp instruction covering whole rows.
      if (n_steps > 0) {
        mbar_wait_a(BBAR(acc_full), ic & 1);
        tc_fence_after();
      }
      tr.rec(8, ic);
      constexpr int kChunks = DP / 8;                 // 16-byte chunks per row
      const uint32_t sStage = smem_a + C::offStage;
      const bool full_tile = row0 + BS <= it.len_stat;
      auto stage_addr = [&](int rr, uint32_t chunk) {
        return sStage + (chunk >> 3) * (BS * 128) + rr * 128 + (((chunk ^ rr) & 7u) << 4);
      };
      auto store_acc = [&](uint32_t col, const CUtensorMap* tm, void* out, int64_t row_stride, int64_t head_stride,
                           float mult, bool last) {
        if (store_pending) {   // thread 0 only: the previous bulk store has read the staging tile
          tma_store_wait_read<0>();
          store_pending = false;
        }
        uint32_t o[DP / 64][32];
#pragma unroll
        for (int cc = 0; cc < DP / 64; ++cc) {
          if (n_steps > 0) {
            tmem_ld32(tmem_base + lane_addr + col + (g * (DP / 64) + cc) * 32, o[cc]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[cc][i] = 0u;
          }
        }
        if (n_steps > 0) tmem_ld_wait();
        if (last && n_steps > 0) {
          // both accumulators are in registers: the next item's products may overwrite them
          tc_fence_before();
          mbar_arrive_a(BBAR(acc_free));
        }
        named_bar_sync(1, 256);   // (thread 0 has waited for the previous store) the staging tile is free
#pragma unroll
        for (int cc = 0; cc < DP / 64; ++cc) {
          const int c = g * (DP / 64) + cc;           // 32-column group of this thread
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(stage_addr(r, static_cast<uint32_t>(c * 4 + q)),
                   pack2<kBF16>(__uint_as_float(o[cc][q * 8 + 0]) * mult, __uint_as_float(o[cc][q * 8 + 1]) * mult),
                   pack2<kBF16>(__uint_as_float(o[cc][q * 8 + 2]) * mult, __uint_as_float(o[cc][q * 8 + 3]) * mult),
                   pack2<kBF16>(__uint_as_float(o[cc][q * 8 + 4]) * mult, __uint_as_float(o[cc][q * 8 + 5]) * mult),
                   pack2<kBF16>(__uint_as_float(o[cc][q * 8 + 6]) * mult, __uint_as_float(o[cc][q * 8 + 7]) * mult));
        }
        tr.rec(10, ic);
        if (full_tile) {
          fence_proxy_async_smem();   // staged rows -> visible to the bulk store
          named_bar_sync(1, 256);
          if (t256 == 0) {
#pragma unroll
            for (int pn = 0; pn < C::kPanels; ++pn)
              tma_store_3d(tm, sStage + pn * (BS * 128), pn * 64, it.head, it.stat_begin + row0);
            tma_store_commit();
            store_pending = true;
          }
        } else {
          named_bar_sync(1, 256);
          uint8_t* obase = reinterpret_cast<uint8_t*>(out) +
                           2 * (static_cast<int64_t>(it.stat_begin + row0) * row_stride + it.head * head_stride);
#pragma unroll
          for (int k = 0; k < BS * kChunks / 256; ++k) {
            const int idx = k * 256 + t256;
            const int rr = idx / kChunks;
            const uint32_t chunk = static_cast<uint32_t>(idx % kChunks);
            if (row0 + rr < it.len_stat && static_cast<int>(chunk) * 8 < p.headdim) {
              const uint4 v = lds128(stage_addr(rr, chunk));
              *reinterpret_cast<uint4*>(obase + 2 * (static_cast<int64_t>(rr) * row_stride + chunk * 8)) = v;
            }
          }
        }
        tr.rec(11, ic);
      };
      if constexpr (!kDQ)
        store_acc(C::colA1, &tmOut1, p.out1, p.o1_row_stride, p.o1_head_stride, kDrop ? p.drop_scale : 1.f, false);
      store_acc(C::colA2, &tmOut2, p.out2, p.o2_row_stride, p.o2_head_stride, p.scale, true);
      tr.rec(9, ic);
      if (n_steps > 0) ++ic;
    }
    if (store_pending) tma_store_wait_all();   // shared memory must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, C::kTmemCols);
#ifdef BP_TRACE
  if (p.trace && threadIdx.x == 0 && blockIdx.x < 4096) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    uint64_t* rec = p.trace + 8 * kTraceRecs * 2 + 2 * 148 + 4 * blockIdx.x;
    rec[0] = cta_t0;
    rec[1] = global_timer_ns();
    rec[2] = smid;
    rec[3] = 0;
  }
#endif
}

// (batch, head) pairs per scheduling chunk: chunk_bh * num_tiles CTAs = one resident wave
inline int chunk_bh_for(int dp, int sms, int num_tiles, int total_bh) {
  const int resident = (dp == 64 ? 2 : 1) * sms;
  int c = resident / num_tiles;
  if (c < 1) c = 1;
  return c < total_bh ? c : total_bh;
}

template <int DP, bool kDQ, bool kBF16, bool kDrop>
int launch(const CUtensorMap& s1, const CUtensorMap& s2, const CUtensorMap& b1, const CUtensorMap& b2,
           const CUtensorMap& o1, const CUtensorMap& o2, const Params& p, cudaStream_t stream) {
  using C = Cfg<DP, kDQ>;
  auto kern = fmha_bwd_kernel<DP, kDQ, kBF16, kDrop>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_fmha_bwd: cudaFuncSetAttribute(%u B smem): %s", C::kSmemBytes, cudaGetErrorString(e));
  }
  // one resident wave of CTAs; each starts with ticket blockIdx.x and draws the rest from the counter
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int resident = C::kCtasPerSm * sms;
  const int grid = p.num_tickets < resident ? p.num_tickets : resident;
  e = cudaMemsetAsync(p.sched, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_fmha_bwd: cudaMemsetAsync(scheduler counter): %s", cudaGetErrorString(e));
  }
  kern<<<static_cast<unsigned>(grid), C::kThreads, C::kSmemBytes, stream>>>(s1, s2, b1, b2, o1, o2, p);
  return check_launch(kDQ ? "bp_fmha_bwd (dQ) launch" : "bp_fmha_bwd (dK, dV) launch");
}

}  // namespace fmha_bwd
}  // namespace bp

// Dropout column words C[bh][k] (bp_common.cuh) for every (batch, head) and key position below s_pad_k.
namespace bp {
__global__ void __launch_bounds__(256)
drop_col_table_kernel(uint32_t* __restrict__ table, uint64_t seed, int32_t s_pad_k) {
  const uint32_t bh = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k < s_pad_k) table[static_cast<int64_t>(bh) * s_pad_k + k] = drop_col_word(drop_base(seed, bh), static_cast<uint32_t>(k));
}

int launch_drop_col_table(uint32_t* table, uint64_t seed, int32_t bh, int32_t s_pad_k, cudaStream_t st) {
  if (bh > 65535) return fail(BP_ERR_UNSUPPORTED, "attention dropout: batch * nheads = %d exceeds 65535", bh);
  drop_col_table_kernel<<<dim3((s_pad_k + 255) / 256, bh), 256, 0, st>>>(table, seed, s_pad_k);
  return check_launch("attention dropout (column words) launch");
}

// p -> threshold of the 8-bit comparison; the effective probability is thr / 256
int drop_threshold(float p) {
  if (!(p > 0.f)) return 0;
  int thr = static_cast<int>(p * 256.f + 0.5f);
  return thr < 1 ? 1 : (thr > 255 ? 255 : thr);
}
}  // namespace bp

// workspace = [2 ticket counters, padded to 16 bytes][row statistics][dropout column words]
constexpr int64_t kSchedBytes = 16;
static int64_t stats_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_q) {
  const int64_t s_pad = (static_cast<int64_t>(max_seqlen_q) + 127) / 128 * 128;
  return kSchedBytes + static_cast<int64_t>(batch) * nheads * (s_pad / 64) * bp::fmha_bwd::kStatsWords * 4;
}

extern "C" int64_t bp_fmha_bwd_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_q) {
  if (batch <= 0 || nheads <= 0 || max_seqlen_q <= 0) return 0;
  return stats_bytes(batch, nheads, max_seqlen_q);
}

extern "C" int64_t bp_fmha_bwd_dropout_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_q,
                                                       int32_t max_seqlen_k) {
  if (batch <= 0 || nheads <= 0 || max_seqlen_q <= 0 || max_seqlen_k <= 0) return 0;
  const int64_t s_pad_k = (static_cast<int64_t>(max_seqlen_k) + 127) / 128 * 128;
  return stats_bytes(batch, nheads, max_seqlen_q) + static_cast<int64_t>(batch) * nheads * s_pad_k * 4;
}

static int fmha_bwd_impl(const void* dout, const void* q, const void* k, const void* v, const void* out,
                         const float* softmax_lse, void* dq, void* dk, void* dv, const int32_t* cu_seqlens_q,
                         const int32_t* cu_seqlens_k, int32_t batch, int32_t nheads, int32_t headdim, int32_t total_q,
                         int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k, const int64_t* strides,
                         int32_t lse_stride, float softmax_scale, int32_t is_causal, int32_t dtype, float p_dropout,
                         uint64_t seed, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace bp;
  if (!(p_dropout >= 0.f) || p_dropout >= 1.f)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: p_dropout must be in [0, 1) (got %f)", (double)p_dropout);
  const int thr = drop_threshold(p_dropout);
  const bool drop = thr > 0;
  if (!dout || !q || !k || !v || !out || !softmax_lse || !dq || !dk || !dv || !cu_seqlens_q || !cu_seqlens_k || !strides)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: null pointer argument");
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: only fp16 and bf16 are supported (dtype=%d)", dtype);
  if (batch <= 0 || nheads <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: batch and nheads must be positive");
  if (headdim <= 0 || headdim % 8 != 0 || headdim > 128)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: head dim must be a multiple of 8 and <= 128 (got %d)", headdim);
  if (total_q <= 0 || total_k <= 0 || max_seqlen_q <= 0 || max_seqlen_k <= 0)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: empty input (total_q=%d total_k=%d)", total_q, total_k);
  if (lse_stride < max_seqlen_q)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: lse_stride %d < max_seqlen_q %d", lse_stride, max_seqlen_q);
  for (int i = 0; i < 16; ++i)
    if (strides[i] <= 0 || strides[i] % 8 != 0)
      return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: row/head strides must be positive multiples of 8 elements (got %lld)",
                  (long long)strides[i]);
  const uintptr_t ptrs[] = {(uintptr_t)dout, (uintptr_t)q, (uintptr_t)k, (uintptr_t)v, (uintptr_t)out,
                            (uintptr_t)dq,   (uintptr_t)dk, (uintptr_t)dv, (uintptr_t)workspace};
  for (uintptr_t a : ptrs)
    if (a % 16 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: tensors and workspace must be 16-byte aligned");
  const int64_t need = drop ? bp_fmha_bwd_dropout_workspace_bytes(batch, nheads, max_seqlen_q, max_seqlen_k)
                            : bp_fmha_bwd_workspace_bytes(batch, nheads, max_seqlen_q);
  if (!workspace || workspace_bytes < need)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: workspace of %lld bytes needed (got %lld)", (long long)need,
                (long long)workspace_bytes);
  // strides: {row, head} of dout, q, k, v, out, dq, dk, dv
  const int64_t *s_do = strides, *s_q = strides + 2, *s_k = strides + 4, *s_v = strides + 6, *s_o = strides + 8,
                *s_dq = strides + 10, *s_dk = strides + 12, *s_dv = strides + 14;
  const int DP = headdim <= 64 ? 64 : 128;
  const int s_pad = (max_seqlen_q + 127) / 128 * 128;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bf16 = dtype == BP_DTYPE_BF16;
  unsigned int* sched = static_cast<unsigned int*>(workspace);
  float* stats = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kSchedBytes);

  {
    dim3 grid(batch * nheads, s_pad / 64);
    if (bf16)
      fmha_bwd::bwd_stats_kernel<true><<<grid, 256, 0, st>>>(dout, out, softmax_lse, stats, cu_seqlens_q, s_do[0], s_do[1],
                                                             s_o[0], s_o[1], lse_stride, s_pad, nheads, headdim, seed);
    else
      fmha_bwd::bwd_stats_kernel<false><<<grid, 256, 0, st>>>(dout, out, softmax_lse, stats, cu_seqlens_q, s_do[0], s_do[1],
                                                              s_o[0], s_o[1], lse_stride, s_pad, nheads, headdim, seed);
    if (int rc = check_launch("bp_fmha_bwd (row statistics) launch")) return rc;
  }
  const int s_pad_k = (max_seqlen_k + 127) / 128 * 128;
  uint32_t* drop_c = nullptr;
  if (drop) {
    drop_c = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + stats_bytes(batch, nheads, max_seqlen_q));
    if (int rc = launch_drop_col_table(drop_c, seed, batch * nheads, s_pad_k, st)) return rc;
  }

  CUtensorMap tmQ, tmK, tmV, tmDO, tmDQ, tmDK, tmDV;
  {
    const uint32_t box[3] = {64, 1, (uint32_t)fmha_bwd::BT};
    const uint64_t dq_[3] = {(uint64_t)headdim, (uint64_t)nheads, (uint64_t)total_q};
    const uint64_t dk_[3] = {(uint64_t)headdim, (uint64_t)nheads, (uint64_t)total_k};
    const uint64_t sq[2] = {(uint64_t)s_q[1] * 2, (uint64_t)s_q[0] * 2};
    const uint64_t sk[2] = {(uint64_t)s_k[1] * 2, (uint64_t)s_k[0] * 2};
    const uint64_t sv[2] = {(uint64_t)s_v[1] * 2, (uint64_t)s_v[0] * 2};
    const uint64_t sdo[2] = {(uint64_t)s_do[1] * 2, (uint64_t)s_do[0] * 2};
    if (int rc = encode_tensor_map(&tmQ, dtype, 3, q, dq_, sq, box, true)) return rc;
    if (int rc = encode_tensor_map(&tmK, dtype, 3, k, dk_, sk, box, true)) return rc;
    if (int rc = encode_tensor_map(&tmV, dtype, 3, v, dk_, sv, box, true)) return rc;
    if (int rc = encode_tensor_map(&tmDO, dtype, 3, dout, dq_, sdo, box, true)) return rc;
    // gradient tiles leave through bulk stores of whole stationary tiles (128 rows x one 64-column panel)
    const uint32_t obox[3] = {64, 1, (uint32_t)fmha_bwd::BS};
    const uint64_t sdq[2] = {(uint64_t)s_dq[1] * 2, (uint64_t)s_dq[0] * 2};
    const uint64_t sdk[2] = {(uint64_t)s_dk[1] * 2, (uint64_t)s_dk[0] * 2};
    const uint64_t sdv[2] = {(uint64_t)s_dv[1] * 2, (uint64_t)s_dv[0] * 2};
    if (int rc = encode_tensor_map(&tmDQ, dtype, 3, dq, dq_, sdq, obox, true)) return rc;
    if (int rc = encode_tensor_map(&tmDK, dtype, 3, dk, dk_, sdk, obox, true)) return rc;
    if (int rc = encode_tensor_map(&tmDV, dtype, 3, dv, dk_, sdv, obox, true)) return rc;
  }
  fmha_bwd::Params p;
  p.stats = stats;
  p.cu_q = cu_seqlens_q;
  p.cu_k = cu_seqlens_k;
  p.s_pad = s_pad;
  p.batch = batch;
  p.nheads = nheads;
  p.headdim = headdim;
  p.is_causal = is_causal ? 1 : 0;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * fmha_bwd::kLog2e;
  p.drop_c = drop_c;
  p.s_pad_k = s_pad_k;
  p.drop_thr24 = static_cast<uint32_t>(thr) << 24;
  p.drop_scale = drop ? 256.f / static_cast<float>(256 - thr) : 1.f;
  p.trace = nullptr;
#ifdef BP_TRACE
  const char* trace_mode = getenv("BP_TRACE_BWD");   // "dkdv" or "dq": which of the two launches writes the timeline
#endif
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  // keys own: dK, dV
  auto set_tiles = [&](int max_seqlen, unsigned int* counter) -> int {
    p.num_tiles = (max_seqlen + fmha_bwd::BS - 1) / fmha_bwd::BS;
    p.chunk_bh = fmha_bwd::chunk_bh_for(DP, sms, p.num_tiles, batch * nheads);
    const int64_t chunks = (static_cast<int64_t>(batch) * nheads + p.chunk_bh - 1) / p.chunk_bh;
    const int64_t tickets = chunks * p.chunk_bh * p.num_tiles;
    if (tickets > 0x3fffffff) return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_bwd: too many work items");
    p.num_tickets = static_cast<int32_t>(tickets);
    p.sched = counter;
    return BP_OK;
  };
  if (int rc0 = set_tiles(max_seqlen_k, sched)) return rc0;
  p.out1 = dv;
  p.o1_row_stride = s_dv[0];
  p.o1_head_stride = s_dv[1];
  p.out2 = dk;
  p.o2_row_stride = s_dk[0];
  p.o2_head_stride = s_dk[1];
#ifdef BP_TRACE
  p.trace = (trace_mode && trace_mode[0] == 'd' && trace_mode[1] == 'k') ? g_trace : nullptr;
#endif
  int rc;
  if (DP == 64)
    rc = bf16 ? (drop ? fmha_bwd::launch<64, false, true, true>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st) : fmha_bwd::launch<64, false, true, false>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st))
              : (drop ? fmha_bwd::launch<64, false, false, true>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st) : fmha_bwd::launch<64, false, false, false>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st));
  else
    rc = bf16 ? (drop ? fmha_bwd::launch<128, false, true, true>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st) : fmha_bwd::launch<128, false, true, false>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st))
              : (drop ? fmha_bwd::launch<128, false, false, true>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st) : fmha_bwd::launch<128, false, false, false>(tmK, tmV, tmQ, tmDO, tmDV, tmDK, p, st));
  if (rc) return rc;

  // queries own: dQ
  if (int rc0 = set_tiles(max_seqlen_q, sched + 1)) return rc0;
#ifdef BP_TRACE
  p.trace = (trace_mode && trace_mode[0] == 'd' && trace_mode[1] == 'q') ? g_trace : nullptr;
#endif
  p.out1 = nullptr;
  p.out2 = dq;
  p.o2_row_stride = s_dq[0];
  p.o2_head_stride = s_dq[1];
  if (DP == 64)
    rc = bf16 ? (drop ? fmha_bwd::launch<64, true, true, true>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st) : fmha_bwd::launch<64, true, true, false>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st))
              : (drop ? fmha_bwd::launch<64, true, false, true>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st) : fmha_bwd::launch<64, true, false, false>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st));
  else
    rc = bf16 ? (drop ? fmha_bwd::launch<128, true, true, true>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st) : fmha_bwd::launch<128, true, true, false>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st))
              : (drop ? fmha_bwd::launch<128, true, false, true>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st) : fmha_bwd::launch<128, true, false, false>(tmQ, tmDO, tmK, tmV, tmDQ, tmDQ, p, st));
  return rc;
}

extern "C" int bp_fmha_bwd(const void* dout, const void* q, const void* k, const void* v, const void* out,
                           const float* softmax_lse, void* dq, void* dk, void* dv, const int32_t* cu_seqlens_q,
                           const int32_t* cu_seqlens_k, int32_t batch, int32_t nheads, int32_t headdim, int32_t total_q,
                           int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k, const int64_t* strides,
                           int32_t lse_stride, float softmax_scale, int32_t is_causal, int32_t dtype, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  return fmha_bwd_impl(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k, batch, nheads, headdim,
                       total_q, total_k, max_seqlen_q, max_seqlen_k, strides, lse_stride, softmax_scale, is_causal, dtype,
                       0.f, 0, workspace, workspace_bytes, stream);
}

extern "C" int bp_fmha_bwd_dropout(const void* dout, const void* q, const void* k, const void* v, const void* out,
                                   const float* softmax_lse, void* dq, void* dk, void* dv, const int32_t* cu_seqlens_q,
                                   const int32_t* cu_seqlens_k, int32_t batch, int32_t nheads, int32_t headdim,
                                   int32_t total_q, int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k,
                                   const int64_t* strides, int32_t lse_stride, float softmax_scale, int32_t is_causal,
                                   int32_t dtype, float p_dropout, uint64_t seed, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
  return fmha_bwd_impl(dout, q, k, v, out, softmax_lse, dq, dk, dv, cu_seqlens_q, cu_seqlens_k, batch, nheads, headdim,
                       total_q, total_k, max_seqlen_q, max_seqlen_k, strides, lse_stride, softmax_scale, is_causal, dtype,
                       p_dropout, seed, workspace, workspace_bytes, stream);
}
