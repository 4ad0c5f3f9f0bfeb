// FlashAttention forward for sm_100a: softmax(scale * Q K^T [+ causal mask]) V, fp32 accumulate.
//
// Replaces the reference's FlashAttention-1 forward (csrc/flash_attn/fmha_api.cpp:189-325 ->
// src/fmha_fprop_kernel_1xN.h:199-696).  That kernel walks K/V blocks in the outer loop and bounces a
// fp32 O through HBM between blocks; this one is written for Blackwell from scratch:
//
//   * Q-outer / KV-inner: a work item is 2 x 128 query rows of one (batch, head); O stays in TMEM for the
//     whole KV sweep, so Q, K, V are read once and O written once (the algorithmic traffic).
//   * persistent with a dynamic scheduler: one CTA per SM; the producer warp draws work items from a global
//     ticket counter (heaviest causal tiles first inside chunks of 148 (batch, head) pairs, so that the K/V of
//     a chunk stay in L2 and the last tickets are the lightest ones) and publishes the decoded item through a
//     small shared-memory ring to the other roles; barrier phases run on across items, Q is double-buffered
//     per item and the K/V rings never drain, so the loads and the first S = Q K^T of the next item overlap
//     the epilogue of the current one.  The counter is per LAUNCH: a slot of a device-side ring, zeroed by a
//     memset node in front of the kernel, so concurrent launches (streams, graph replays) never share tickets
//     and an aborted launch cannot poison a later one.
//   * warp-specialised (384 threads): warp 0 = scheduler + TMA producer, warps 1 / 2 = tcgen05.mma issuers for
//     query tile 0 / 1 (two independent pipelines sharing the K/V tiles), warp 3 = TMEM allocator + O-tile store
//     issuer, then ONE softmax warpgroup per query tile: one thread per query row (TMEM lane == row, so row
//     max / row sum need no shuffles) holding the whole score row of a key block in registers.
//   * the kernel is bound by the MUFU ex2 pipe (16 exponentials per clock and SM against 128 x 128 scores per
//     512 tensor-pipe cycles at head dim 64), so everything is arranged to keep that pipe busy: each SM
//     sub-partition hosts exactly two softmax warps (one per query tile) whose exponential phases are long
//     (a full 128-key row per thread) and whose other phases (TMEM load, row max, hand-overs) overlap the
//     sibling tile's exponentials; the causal asymmetry of the two tiles keeps them out of phase.
//   * S (M=128, N=BN) lands in TMEM; the softmax threads tcgen05.ld their row and hand the S buffer back at
//     once (s_free), so S of the next key block is computed while this block's exponentials run; running
//     max with *lazy* rescaling (O is only touched when the max grew by more than 2^8); P = exp2(...) is packed to
//     bf16/f16 and written back to TMEM (tcgen05.st), and O += P V is a TS tcgen05.mma (A = P from TMEM, B = V
//     consumed MN-major straight from the TMA tile, no transpose): P never touches shared memory.
//   * chunks of 32 keys that lie above the causal diagonal for a whole warp cost no exponentials at all.
//   * a finished full O tile is staged in shared memory and leaves as ONE TMA store per 64-column panel.
//
// Varlen: sequences are addressed through cu_seqlens (rows of other sequences that fall inside a tile
// are masked / never stored), head dims that are a multiple of 8 up to 128 are handled by TMA zero-fill
// to a padded width DP of 64 or 128.
#include <atomic>
#include <cstddef>

#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace fmha {

constexpr int BM = 128;           // query rows per tile (= TMEM lanes)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kItemSlots = 4;     // shared-memory ring of decoded work items (producer -> all other roles)
constexpr int kChunkBH = 148;     // (batch, head) pairs per scheduling chunk
#ifndef BP_FMHA_POLY
#define BP_FMHA_POLY 2
#endif
// Of every 8 scores, how many take the polynomial exp2 (FMA pipe, rel. error 7.5e-5, far below the 16-bit rounding of
// P) instead of MUFU (0, 2 or 4).  Measured at config 2 (profiles/): 0 -> 103.7 us, 2 -> 99.8 us; parity-green.
constexpr int kPoly = BP_FMHA_POLY;
#ifndef BP_FMHA_STAGGER
#define BP_FMHA_STAGGER 0
#endif
constexpr int kStagger = BP_FMHA_STAGGER;   // cycles query tile 1 holds back its very first block (debug knob)

template <int DP>
struct Cfg {
  static constexpr int BN = (DP == 64) ? 128 : 64;   // keys per block
  static constexpr int kThreads = 384;               // warps 0-3 + one softmax warpgroup per query tile
  // setmaxnreg only redistributes the registers the CTA was launched with (kThreads x the compiled per-thread
  // count, 168): 128 * kRegsLow + 256 * kRegsHigh must not exceed it, or the last warps to grow block forever
  static constexpr int kRegsLow = 56;
  static constexpr int kRegsHigh = 224;
  static constexpr int kStages = 3;                  // K and V rings
  static constexpr int kQBufs = (DP == 64) ? 2 : 1;  // Q buffers across work items (smem-limited at DP = 128)
  static constexpr int kPanelsD = DP / 64;           // 64-column (128 B) panels along head dim
  static constexpr uint32_t kQTileBytes = BM * DP * 2;
  static constexpr uint32_t kKVTileBytes = BN * DP * 2;
  static constexpr uint32_t kKVPanelBytes = BN * 128;  // one 64-column panel of a K/V tile
  static constexpr uint32_t kOTileBytes = BM * DP * 2; // staging of a finished O tile (kPanelsD panels of 128 rows x 128 B)
  // shared memory map (all tile bases 1024-aligned)
  static constexpr uint32_t offQ = 0;                                  // [kQBufs item buffers][2 tiles]
  static constexpr uint32_t offK = offQ + kQBufs * 2 * kQTileBytes;
  static constexpr uint32_t offV = offK + kStages * kKVTileBytes;
  static constexpr uint32_t offO = offV + kStages * kKVTileBytes;     // [2 tiles]
  static constexpr uint32_t offItems = offO + 2 * kOTileBytes;        // ring of decoded work items
  static constexpr uint32_t offBar = offItems + kItemSlots * 64;
  static constexpr uint32_t kSmemBytes = offBar + 512 + 1024;  // + barriers + alignment slack
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  // TMEM columns: S_t (fp32, BN columns), O_t (fp32, DP columns), P_t (16-bit pairs, BN / 2 columns)
  static constexpr uint32_t colS = 0;                 // S_t at colS + t*BN
  static constexpr uint32_t colO = 2 * BN;            // O_t at colO + t*DP
  static constexpr uint32_t colP = 2 * BN + 2 * DP;   // P_t at colP + t*BN/2
  static constexpr uint32_t kTmemCols = 512;
  static_assert(colP + BN <= 512, "TMEM budget");
};

struct Params {
  void* out;
  float* lse;
  const int32_t* cu_q;
  const int32_t* cu_k;
  int64_t o_row_stride, o_head_stride;
  int32_t lse_stride;
  int32_t batch, nheads, headdim;
  int32_t num_pairs;  // ceil(max_seqlen_q / 256)
  int32_t num_items;  // num_pairs * batch * nheads
  unsigned int* sched;  // this launch's ticket counter (zeroed by a memset in front of the kernel)
  int32_t is_causal;
  int32_t out_f32;    // debug / test mode: `out` is fp32 (same element strides), written before the 16-bit rounding
  float scale;        // softmax scale
  float scale_log2;   // scale * log2(e)
  // attention dropout (bp_common.cuh: counter-based keep mask; fmha_fwd_kernel<..., kDrop = true> only)
  const uint32_t* drop_c;   // (batch * nheads, s_pad_k) column words
  uint64_t drop_seed;
  int32_t s_pad_k;
  uint32_t drop_thr24;      // round(256 p) << 24
  float drop_scale;         // 1 / (1 - p)
  uint64_t* trace;    // debug: per-role event timestamps of CTA 0 (null in production)
};

struct Barriers {
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[3], k_empty[3], v_full[3], v_empty[3];
  uint64_t s_full[2], s_free[2];           // [tile]
  uint64_t p_ready[2], pv_done[2];         // [tile]
  uint64_t item_full[kItemSlots], item_empty[kItemSlots];
  uint64_t o_staged[2], o_free[2];         // [tile]: O tile staged in smem / read out by the TMA store
  uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= 512, "barrier block");
#define BAR_I(field, i) (bars_a + static_cast<uint32_t>(offsetof(Barriers, field)) + 8u * static_cast<uint32_t>(i))

// One work item = (pair of query tiles, head, batch).  The producer warp decodes it once and publishes it
// through the shared-memory ring; every other role reads the same sequence of items from there.
struct Item {
  int head, batch, q_begin, k_begin, len_q, len_k, row0;
  int n0, n1;   // key blocks visited by query tile 0 / 1
  int end;      // sentinel: no more work for this CTA
  __device__ __forceinline__ int n_of(int t) const { return t == 0 ? n0 : n1; }
  __device__ __forceinline__ int n_max() const { return max(n0, n1); }
};

// Decoding is split so that the producer can issue the cu_seqlens loads of the NEXT item, go on issuing TMA
// loads of the current one, and only then consume them.
struct ItemPre {
  int head, batch, pq;
  int2 cq, ck;
};
__device__ __forceinline__ ItemPre decode_prefetch(const Params& p, int w) {
  // Ticket order: chunks of kChunkBH (batch, head) pairs; inside a chunk all heaviest query-tile pairs first,
  // then the next lighter class, ...  The CTAs that run together sweep the K/V of one chunk (L2 hits instead
  // of DRAM re-reads: every K/V block is needed by all pairs below it), heavy items are handed out before
  // light ones, and the last tickets of the launch are the lightest (short tail).
  ItemPre pre;
  const int total_bh = p.batch * p.nheads;
  const int per_chunk = kChunkBH * p.num_pairs;
  const int c = w / per_chunk;
  const int r = w - c * per_chunk;
  const int bh0 = c * kChunkBH;
  const int gc = min(kChunkBH, total_bh - bh0);
  pre.pq = r / gc;  // 0 = heaviest pair (most key blocks)
  const int bhi = bh0 + (r - pre.pq * gc);
  pre.batch = bhi / p.nheads;
  pre.head = bhi - pre.batch * p.nheads;
  pre.cq = make_int2(__ldg(p.cu_q + pre.batch), __ldg(p.cu_q + pre.batch + 1));
  pre.ck = make_int2(__ldg(p.cu_k + pre.batch), __ldg(p.cu_k + pre.batch + 1));
  return pre;
}
template <int BN>
__device__ __forceinline__ Item decode_finish(const Params& p, const ItemPre& pre) {
  Item it;
  it.batch = pre.batch;
  it.head = pre.head;
  it.q_begin = pre.cq.x;
  it.len_q = pre.cq.y - pre.cq.x;
  it.k_begin = pre.ck.x;
  it.len_k = pre.ck.y - pre.ck.x;
  it.row0 = (p.num_pairs - 1 - pre.pq) * 2 * BM;
  auto blocks = [&](int r0) {
    if (r0 >= it.len_q) return 0;
    int kmax = it.len_k;
    if (p.is_causal) kmax = min(kmax, r0 + BM);
    return (kmax + BN - 1) / BN;
  };
  it.n0 = blocks(it.row0);
  it.n1 = blocks(it.row0 + BM);
  it.end = 0;
  return it;
}
__device__ __forceinline__ Item end_item() {
  Item fin;
  fin.head = fin.batch = fin.q_begin = fin.k_begin = fin.len_q = fin.len_k = fin.row0 = fin.n0 = fin.n1 = 0;
  fin.end = 1;
  return fin;
}

__device__ __forceinline__ void put_item(uint32_t a, const Item& it) {
  sts128(a, it.head, it.batch, it.q_begin, it.k_begin);
  sts128(a + 16, it.len_q, it.len_k, it.row0, it.n0);
  sts128(a + 32, it.n1, it.end, 0, 0);
}
__device__ __forceinline__ Item get_item(uint32_t a) {
  const uint4 x = lds128(a), y = lds128(a + 16), z = lds128(a + 32);
  Item it;
  it.head = x.x; it.batch = x.y; it.q_begin = x.z; it.k_begin = x.w;
  it.len_q = y.x; it.len_k = y.y; it.row0 = y.z; it.n0 = y.w;
  it.n1 = z.x; it.end = z.y;
  return it;
}

template <int DP, bool kBF16, bool kDrop>
__global__ void __launch_bounds__(Cfg<DP>::kThreads, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  using C = Cfg<DP>;
  constexpr int BN = C::BN;
  constexpr int NC = BN / 32;   // 32-key chunks of a block
  extern __shared__ uint8_t smem_raw[];
  // everything below works on 32-bit shared-window addresses (see bp_common.cuh)
  const uint32_t smem_a = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars_a = smem_a + C::offBar;
  const uint32_t items_a = smem_a + C::offItems;
  uint8_t* smem = smem_raw + (smem_a - smem_u32(smem_raw));
  Barriers& bars = *reinterpret_cast<Barriers*>(smem + C::offBar);

  // Warp roles are numbered 0-3 (producer, two MMA issuers, store issuer) and 4-11 (softmax), but the service roles
  // run in the HIGHEST physical warps: the sub-partition arbiter prefers the highest warp id among eligible warps,
  // and a single-thread role that loses every issue slot to an always-eligible softmax warp issues an MMA every
  // ~130 cycles instead of every ~40 (measured with the timeline trace: 8 MMAs of a PV product took ~1000 cycles).
  // Physical warps 0-7 -> roles 4-11 (same TMEM lane quadrant, warp & 3), physical warps 8-11 -> roles 0-3.
  const int warp = role_warp<12>();
  const int lane = threadIdx.x & 31;
#ifdef BP_TRACE
  if (p.trace && threadIdx.x == 0) {
    uint64_t ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    p.trace[8 * kTraceRecs * 2 + 2 * blockIdx.x] = ns;
  }
#endif

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.q_full[i], 1);
      mbar_init(&bars.q_empty[i], 2);   // both MMA warps release a Q buffer
      mbar_init(&bars.s_full[i], 1);
      mbar_init(&bars.s_free[i], 128);
      mbar_init(&bars.p_ready[i], 128);
      mbar_init(&bars.pv_done[i], 1);
      mbar_init(&bars.o_staged[i], 128);
      mbar_init(&bars.o_free[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars.k_full[i], 1);
      mbar_init(&bars.k_empty[i], 2);   // both MMA warps release a K / V slot
      mbar_init(&bars.v_full[i], 1);
      mbar_init(&bars.v_empty[i], 2);
    }
    for (int i = 0; i < kItemSlots; ++i) {
      mbar_init(&bars.item_full[i], 1);
      mbar_init(&bars.item_empty[i], 2 + 8 + 1);   // one lane of every consumer warp (2 MMA, 8 softmax, warp 3)
    }
    fence_barrier_init();
  }
  if (warp == 3) {
    tmem_alloc(&bars.tmem_base, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  // every consumer role walks the item ring with its own sequence number
  uint32_t item_seq = 0;
  auto next_item = [&]() {
    const uint32_t slot = item_seq % kItemSlots;
    mbar_wait_a(BAR_I(item_full, slot), (item_seq / kItemSlots) & 1);
    const Item it = get_item(items_a + slot * 64);
    __syncwarp();
    if (lane == 0) mbar_arrive_a(BAR_I(item_empty, slot));
    ++item_seq;
    return it;
  };

  if (warp < 4) {
    reg_dealloc<C::kRegsLow>();
    if (warp == 0) {
      // ===================== scheduler + TMA producer (whole warp walks the loop, lane 0 issues) =====================
      uint32_t item_no = 0, blk = 0;  // running counters: Q buffer = item_no % kQBufs, K/V slot = blk % kStages
      uint32_t seq = 0;               // items published (valid items + the end marker)
      Tracer tr(p.trace, 0, blockIdx.x == 0 && lane == 0);
      auto publish = [&](const Item& it) {
        const uint32_t slot = seq % kItemSlots;
        if (seq >= kItemSlots) mbar_wait_a(BAR_I(item_empty, slot), ((seq / kItemSlots) - 1) & 1);
        if (lane == 0) {
          put_item(items_a + slot * 64, it);
          mbar_arrive_a(BAR_I(item_full, slot));
        }
        ++seq;
      };
      // lane 0 draws the next ticket; the result is only consumed (shuffled to the warp) much later
      auto draw = [&]() {
        unsigned int d = 0;
        if (lane == 0) d = gridDim.x + atomicAdd(p.sched, 1u);
        return d;
      };
      // slow path: first item, and tickets that decode to empty items (rows beyond a short sequence)
      auto fetch_valid = [&](int& w) {
        while (w < p.num_items) {
          const Item c = decode_finish<BN>(p, decode_prefetch(p, w));
          w = static_cast<int>(__shfl_sync(0xffffffffu, draw(), 0));
          if (c.n_max() > 0) return c;
        }
        return end_item();
      };
      int w = blockIdx.x;   // first ticket is static; the following ones come from the global counter
      Item cur = fetch_valid(w);
      publish(cur);
      while (!cur.end) {
        // Next item: ticket w is already known.  Start its cu_seqlens loads and the draw after it now, publish it
        // after the first loads of the current item are out, so that the other roles never wait for an item.
        const bool have_next = w < p.num_items;
        ItemPre pre;
        unsigned int drawn = 0;
        if (have_next) {
          pre = decode_prefetch(p, w);
          drawn = draw();
        }
        const Item it = cur;
        const int n_max = it.n_max();
        const uint32_t qb = item_no % C::kQBufs;
        if (item_no >= C::kQBufs) mbar_wait_a(BAR_I(q_empty, qb), ((item_no / C::kQBufs) - 1) & 1);
        {
          // ("_w" forms: the whole warp walks this code with warp-uniform values, an elected lane issues)
          const int n_q_tiles = (it.row0 + BM < it.len_q) ? 2 : 1;
          mbar_arrive_expect_tx_w(BAR_I(q_full, qb), n_q_tiles * C::kQTileBytes);
          for (int t = 0; t < n_q_tiles; ++t)
#pragma unroll
            for (int pn = 0; pn < C::kPanelsD; ++pn)
              tma_load_3d_w(smem_a + C::offQ + (qb * 2 + t) * C::kQTileBytes + pn * (BM * 128), &tmQ,
                            BAR_I(q_full, qb), pn * 64, it.head, it.q_begin + it.row0 + t * BM);
        }
        auto issue_kv = [&](int j) {
          const uint32_t slot = blk % C::kStages;
          const uint32_t ph = ((blk / C::kStages) - 1) & 1;
          const int krow = it.k_begin + j * BN;
          if (blk >= C::kStages) mbar_wait_a(BAR_I(k_empty, slot), ph);
          tr.rec(1, blk);
          mbar_arrive_expect_tx_w(BAR_I(k_full, slot), C::kKVTileBytes);
#pragma unroll
          for (int pn = 0; pn < C::kPanelsD; ++pn)
            tma_load_3d_w(smem_a + C::offK + slot * C::kKVTileBytes + pn * C::kKVPanelBytes, &tmK,
                          BAR_I(k_full, slot), pn * 64, it.head, krow);
          if (blk >= C::kStages) mbar_wait_a(BAR_I(v_empty, slot), ph);
          tr.rec(2, blk);
          mbar_arrive_expect_tx_w(BAR_I(v_full, slot), C::kKVTileBytes);
#pragma unroll
          for (int pn = 0; pn < C::kPanelsD; ++pn)
            tma_load_3d_w(smem_a + C::offV + slot * C::kKVTileBytes + pn * C::kKVPanelBytes, &tmV,
                          BAR_I(v_full, slot), pn * 64, it.head, krow);
          ++blk;
        };
        const int n_first = min(n_max, 2);
        for (int j = 0; j < n_first; ++j) issue_kv(j);
        Item nxt = end_item();
        if (have_next) {
          nxt = decode_finish<BN>(p, pre);
          w = static_cast<int>(__shfl_sync(0xffffffffu, drawn, 0));
          if (nxt.n_max() == 0) nxt = fetch_valid(w);
        }
        publish(nxt);
        for (int j = n_first; j < n_max; ++j) issue_kv(j);
        ++item_no;
        cur = nxt;
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuer of query tile t (whole warp waits, lane 0 issues) =====================
      const int t = warp - 1;
      constexpr uint32_t idesc_s = make_idesc(kBF16, BM, BN, false, false);
      constexpr uint32_t idesc_pv = make_idesc(kBF16, BM, DP, false, true);
      const uint32_t sK = smem_a + C::offK;
      const uint32_t sV = smem_a + C::offV;
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform: stays in uniform registers
      const uint32_t tS = tm + C::colS + t * BN;
      const uint32_t tO = tm + C::colO + t * DP;
      const uint32_t tP = tm + C::colP + t * (BN / 2);
      uint32_t item_no = 0, blk = 0;  // same running counters as the producer
      uint32_t s_cnt = 0;             // S tiles issued by this warp (s_free / s_full phases)
      uint32_t pv_cnt = 0;            // key blocks whose PV products were issued (p_ready / pv_done phases)
      Tracer tr(p.trace, 1 + t, blockIdx.x == 0 && lane == 0);

      while (true) {
        const Item it = next_item();
        if (it.end) break;
        const uint32_t qb = item_no % C::kQBufs;
        const uint32_t sQ = smem_a + C::offQ + (qb * 2 + t) * C::kQTileBytes;
        const int n = it.n_of(t);
        const int n_max = it.n_max();
        mbar_wait_a(BAR_I(q_full, qb), (item_no / C::kQBufs) & 1);

        // The whole warp walks these loops with warp-uniform values; every tcgen05 instruction is predicated on an
        // elected lane inside its asm block (umma_*_w), so there is no divergent region and no ELECT/R2UR waterfall
        // around each MMA (which costs more than a narrow N = 64 MMA takes to execute).  Barriers that complete
        // early (K/V tile landed, S buffer drained) are probed together ahead of their use; only p_ready -- the one
        // hand-over this warp really waits for -- is a blocking wait.
        // S(j) = Q K_j^T into this tile's S buffer, then hand the K slot back (both tiles must do so)
        auto step_s = [&](int j, uint32_t kblk, bool sfree_ready, bool k_ready) {
          const uint32_t slot = kblk % C::kStages;
          const uint32_t ph = (kblk / C::kStages) & 1;
          if (j < n) {
            if (s_cnt >= 1 && !sfree_ready) mbar_wait_a(BAR_I(s_free, t), (s_cnt - 1) & 1);   // softmax has read the previous S
            tr.rec(1, kblk);
            if (!k_ready) mbar_wait_a(BAR_I(k_full, slot), ph);
            tc_fence_after();
            tr.rec(2, kblk);
#pragma unroll
            for (int kk = 0; kk < DP / 16; ++kk) {
              const uint32_t a = sQ + (kk >> 2) * (BM * 128) + (kk & 3) * 32;
              const uint32_t b = sK + slot * C::kKVTileBytes + (kk >> 2) * C::kKVPanelBytes + (kk & 3) * 32;
              umma_ss_w(tS, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024), idesc_s,
                        kk > 0 ? 1u : 0u);
            }
            umma_commit_w(BAR_I(k_empty, slot));
            if (j == n - 1) umma_commit_w(BAR_I(q_empty, qb));   // last read of this item's Q tile
            umma_commit_w(BAR_I(s_full, t));
            ++s_cnt;
          } else {
            // Block not visited by this tile.  The slot still needs this warp's release, but only once the
            // producer has (re)filled it for THIS block: arriving earlier could complete the previous
            // phase of k_empty while the other tile still reads the previous occupant.
            if (!k_ready) mbar_wait_a(BAR_I(k_full, slot), ph);
            umma_commit_w(BAR_I(k_empty, slot));
          }
        };

        if (n == 0) umma_commit_w(BAR_I(q_empty, qb));
        step_s(0, blk, false, false);
        for (int j = 0; j < n_max; ++j, ++blk) {
          const uint32_t slot = blk % C::kStages;
          const uint32_t ph = (blk / C::kStages) & 1;
          const bool has_s = j + 1 < n_max;
          // probes issued back to back: their latencies overlap each other and the S issue below
          const bool k_ready = has_s ? mbar_test_a(BAR_I(k_full, (blk + 1) % C::kStages), ((blk + 1) / C::kStages) & 1) : true;
          const bool sfree_ready = (has_s && j + 1 < n && s_cnt >= 1) ? mbar_test_a(BAR_I(s_free, t), (s_cnt - 1) & 1) : false;
          const bool v_ready = mbar_test_a(BAR_I(v_full, slot), ph);
          if (has_s) step_s(j + 1, blk + 1, sfree_ready, k_ready);
          if (!v_ready) mbar_wait_a(BAR_I(v_full, slot), ph);
          if (j < n) {
            // O_t (+)= P_t V_j : A = P from TMEM (8 columns per K-step of 16 keys), V rows are the K dimension
            // (MN-major B operand, the TMA tile as it landed)
            mbar_wait_a(BAR_I(p_ready, t), pv_cnt & 1);
            tc_fence_after();
            tr.rec(3, blk);
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk) {
              const uint32_t b = sV + slot * C::kKVTileBytes + kk * 16 * 128;
              umma_ts_w(tO, tP + kk * 8, make_smem_desc_sw128(b, C::kKVPanelBytes, 1024), idesc_pv,
                        (j > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit_w(BAR_I(v_empty, slot));
            umma_commit_w(BAR_I(pv_done, t));
            ++pv_cnt;
          } else {
            umma_commit_w(BAR_I(v_empty, slot));   // same pacing rule as for K
          }
        }
        ++item_no;
      }
    } else {
      // ===================== warp 3: O tile stores =====================
      // The softmax warpgroups stage a finished O tile in shared memory; this otherwise idle warp issues the
      // TMA store, waits until the tile has been read out and hands the staging buffer back.
      uint32_t st_cnt[2] = {0, 0};
      while (true) {
        const Item it = next_item();
        if (it.end) break;
        if (p.out_f32) continue;   // test mode: rows are stored directly
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (it.n_of(t) == 0 || it.row0 + t * BM + BM > it.len_q) continue;   // tile absent or stored row by row
          mbar_wait_a(BAR_I(o_staged, t), st_cnt[t] & 1);
          if (lane == 0) {
            for (int pn = 0; pn < C::kPanelsD; ++pn)
              tma_store_3d(&tmO, smem_a + C::offO + t * C::kOTileBytes + pn * (BM * 128), pn * 64, it.head,
                           it.q_begin + it.row0 + t * BM);
            tma_store_commit();
            tma_store_wait_read<0>();
            mbar_arrive_a(BAR_I(o_free, t));
          }
          __syncwarp();
          ++st_cnt[t];
        }
      }
      if (lane == 0) tma_store_wait_all();   // shared memory must outlive the last store
    }
  } else {
    // ===================== softmax warpgroups: one per query tile =====================
    reg_alloc<C::kRegsHigh>();
    const int t = (warp >> 2) - 1;                 // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;          // row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + C::colS + t * BN;
    const uint32_t tO = tmem_base + lane_addr + C::colO + t * DP;
    const uint32_t tP = tmem_base + lane_addr + C::colP + t * (BN / 2);
    // this thread's row of the O staging tile: 16-byte chunk q of panel pn lives at + pn*BM*128 + ((q << 4) ^ sw)
    const uint32_t sO_row = smem_a + C::offO + t * C::kOTileBytes + r * 128;
    const uint32_t sw = (r & 7) << 4;
    bool store_pending = false;
    uint32_t st_cnt = 0;   // O tiles of this query tile handed to warp 3
    const uint32_t bar_s_full = BAR_I(s_full, t), bar_s_free = BAR_I(s_free, t);
    const uint32_t bar_p_ready = BAR_I(p_ready, t), bar_pv_done = BAR_I(pv_done, t);
    const float scale_log2 = p.scale_log2;
    uint32_t cnt = 0;  // key blocks processed by this warpgroup (phases of s_full / pv_done)
    Tracer tr(p.trace, 3 + t, blockIdx.x == 0 && r == 0);
    if constexpr (kStagger > 0) {
      if (t == 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < kStagger) {}
      }
    }
    while (true) {
      const Item it = next_item();
      if (it.end) break;
      const int n = it.n_of(t);
      if (n == 0) continue;
      const int row0_t = it.row0 + t * BM;
      const int qrow = row0_t + r;                 // query index within the sequence
      float m_used = 0.f;  // running max (raw score units) the exponentials are taken against
      float l = 0.f;
      // dropout: this query row's word in a register, the key columns' words from a small table (warp-uniform loads)
      uint32_t rq = 0;
      const uint32_t* crow = nullptr;
      if constexpr (kDrop) {
        const uint32_t bh = static_cast<uint32_t>(it.batch * p.nheads + it.head);
        rq = drop_row_word(drop_base(p.drop_seed, bh), static_cast<uint32_t>(qrow));
        crow = p.drop_c + static_cast<int64_t>(bh) * p.s_pad_k;
      }
      (void)rq;
      (void)crow;

      for (int j = 0; j < n; ++j, ++cnt) {
        tr.rec(0, cnt);
        mbar_wait_a(bar_s_full, cnt & 1);
        tc_fence_after();
        tr.rec(1, cnt);
        float s[BN];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t u[32];
          tmem_ld32(tS + c * 32, u);
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(u[i]);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_a(bar_s_free);   // S(j+1) may now overwrite the buffer while we work on registers
        tr.rec(2, cnt);

        const int col0 = j * BN;
        const bool partial = (col0 + BN > it.len_k) || (p.is_causal && (col0 + BN - 1 > row0_t));
        uint32_t dead = 0;   // bit c: every key of chunk c is masked for every row of this warp
        if (partial) {
          // visible keys of this row inside the block: local column < lim
          const int lim = (p.is_causal ? min(it.len_k, qrow + 1) : it.len_k) - col0;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (!__all_sync(0xffffffffu, lim >= (c + 1) * 32)) {
              if (__all_sync(0xffffffffu, lim <= c * 32)) {
                dead |= 1u << c;
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) s[c * 32 + i] = (c * 32 + i < lim) ? s[c * 32 + i] : -INFINITY;
              }
            }
          }
        }
        // row max over the live chunks: eight independent chains.  ptxas fuses the links into three-input FMNMX3 (66 per
        // row); forcing two-input FMNMX (alternating NaN modes) measured 6 % slower at config 2: twice the instructions.
        float mx8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) mx8[k] = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (!((dead >> c) & 1u)) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
#pragma unroll
              for (int k = 0; k < 8; ++k) mx8[k] = fmaxf(mx8[k], s[c * 32 + i + k]);
            }
          }
        }
        const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                               fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));

        float alpha = 1.f;
        bool grow = false;
        if (l == 0.f) {
          // nothing accumulated yet for this row (first block, or only fully masked keys so far)
          m_used = (mx == -INFINITY) ? 0.f : mx;
        } else {
          grow = (mx - m_used) * scale_log2 > kRescaleThreshold;
          if (grow) {
            alpha = fast_exp2((m_used - mx) * scale_log2);
            m_used = mx;
          }
        }
        tr.rec(3, cnt);
        l *= alpha;   // (alpha = 1 unless the maximum grew: the rescaling of O itself waits for the previous product, below)

        const float neg_m = -m_used * scale_log2;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        // Per 32-key chunk: all arguments (packed FMA), then all 32 exponentials, then the row sums and the packing.
        // A warp owns a quarter-rate MUFU slot every 8 cycles; with the consumer of each exponential right behind it
        // (the compiler's choice for the interleaved form) a single warp stalled on MUFU latency and reached half the
        // pipe's rate, so MUFU was only saturated while BOTH warps of a sub-partition exponentiated at the same time.
        auto exp_chunk = [&](int c, uint32_t (&pk)[16]) {
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) fma2(e[i], e[i + 1], s[c * 32 + i], s[c * 32 + i + 1], scale_log2, neg_m);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
              if (q < 8 - kPoly) {
                e[i + q] = fast_exp2(e[i + q]);
                e[i + q + 1] = fast_exp2(e[i + q + 1]);
              } else {
                exp2_poly_pair(e[i + q], e[i + q + 1]);   // this share of the exponentials runs on the FMA pipe
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            add2(sum4[0], sum4[1], e[i], e[i + 1]);
            add2(sum4[2], sum4[3], e[i + 2], e[i + 3]);
            if constexpr (kDrop) {
              // the row sum (softmax normaliser) takes every probability, the P.V product only the kept ones; the
              // 1 / (1 - p) factor is applied once, to O
              const uint4 W = __ldg(reinterpret_cast<const uint4*>(crow + col0 + c * 32 + i));
              pk[i / 2] = pack2<kBF16>(drop_keep(rq, W.x, p.drop_thr24) ? e[i] : 0.f,
                                       drop_keep(rq, W.y, p.drop_thr24) ? e[i + 1] : 0.f);
              pk[i / 2 + 1] = pack2<kBF16>(drop_keep(rq, W.z, p.drop_thr24) ? e[i + 2] : 0.f,
                                           drop_keep(rq, W.w, p.drop_thr24) ? e[i + 3] : 0.f);
            } else {
              pk[i / 2] = pack2<kBF16>(e[i], e[i + 1]);
              pk[i / 2 + 1] = pack2<kBF16>(e[i + 2], e[i + 3]);
            }
          }
        };
        // The exponentials of the whole block go to registers FIRST; only then does the thread wait for the previous
        // P.V product of its tile (P buffer and O accumulator free), rescale O in the rare case and store P.  With the
        // wait in front of the exponentials (rounds 1-2) every block idled here for the hand-over round trip
        // p_ready -> issuer wake-up -> 8 MMAs -> commit -> wake-up (~700 cycles of a ~2000-cycle block).
        uint32_t pk[NC][16];
        if (!partial) {
          // common case: one straight-line block, so the scheduler can overlap the chunks' phases
#pragma unroll
          for (int c = 0; c < NC; ++c) exp_chunk(c, pk[c]);
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if ((dead >> c) & 1u) {
              // above the causal diagonal: P = 0 without a single exponential
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[c][i] = 0u;
            } else {
              exp_chunk(c, pk[c]);
            }
          }
        }
        if (j > 0) {
          // the P buffer and the O accumulator are free once the previous PV product of this tile has completed
          mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
          tc_fence_after();
          tr.rec(4, cnt);
          if (__any_sync(0xffffffffu, grow)) {
            // rare path: 16 columns at a time keeps the register footprint small
#pragma unroll 1
            for (int c = 0; c < DP / 16; ++c) {
              uint32_t o[16];
              tmem_ld16(tO + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tO + c * 16, o);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) tmem_st16(tP + c * 16, pk[c]);
        l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        tmem_st_wait();
        tr.rec(5, cnt);
        tc_fence_before();
        mbar_arrive_a(bar_p_ready);
        tr.rec(6, cnt);
      }

      // ---- epilogue: O / l -> global, LSE ----
      tr.rec(7, cnt);
      mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
      tc_fence_after();
      tr.rec(8, cnt);
      const bool valid = qrow < it.len_q;
      const float inv_l = kDrop ? p.drop_scale / l : 1.f / l;
      // Full tiles are staged in shared memory (128B-swizzled rows, the layout the TMA store expects) and leave
      // as one bulk store per panel; a thread-per-row store would touch 32 different lines per instruction.  Tiles
      // that end inside the sequence keep the guarded per-row stores (the next sequence's rows follow in memory).
      const bool stage = !p.out_f32 && (row0_t + BM <= it.len_q);
      if (stage && store_pending) {
        mbar_wait_a(BAR_I(o_free, t), (st_cnt - 1) & 1);   // the previous tile's TMA store has read the staging buffer
        store_pending = false;
      }
      const int64_t o_elem = static_cast<int64_t>(it.q_begin + qrow) * p.o_row_stride + it.head * p.o_head_stride;
#pragma unroll
      for (int c = 0; c < DP / 32; ++c) {
        uint32_t o[32];
        float f[32];
        tmem_ld32(tO + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(o[i]) * inv_l;
        if (stage) {
          const uint32_t base = sO_row + (c >> 1) * (BM * 128);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(base + ((static_cast<uint32_t>((c & 1) * 4 + q) << 4) ^ sw), pack2<kBF16>(f[q * 8 + 0], f[q * 8 + 1]),
                   pack2<kBF16>(f[q * 8 + 2], f[q * 8 + 3]), pack2<kBF16>(f[q * 8 + 4], f[q * 8 + 5]),
                   pack2<kBF16>(f[q * 8 + 6], f[q * 8 + 7]));
        } else if (valid) {
          if (p.out_f32) {
            float* orow = reinterpret_cast<float*>(p.out) + o_elem + c * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c * 32 + q * 4 < p.headdim)
                *reinterpret_cast<float4*>(orow + q * 4) = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          } else {
            uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) + 2 * (o_elem + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (c * 32 + q * 8 < p.headdim) {
                uint4 v;
                v.x = pack2<kBF16>(f[q * 8 + 0], f[q * 8 + 1]);
                v.y = pack2<kBF16>(f[q * 8 + 2], f[q * 8 + 3]);
                v.z = pack2<kBF16>(f[q * 8 + 4], f[q * 8 + 5]);
                v.w = pack2<kBF16>(f[q * 8 + 6], f[q * 8 + 7]);
                *reinterpret_cast<uint4*>(orow + q * 16) = v;
              }
            }
          }
        }
      }
      if (valid)
        p.lse[(static_cast<int64_t>(it.batch) * p.nheads + it.head) * p.lse_stride + qrow] =
            m_used * p.scale + __logf(l);
      // The accumulator has been read (wait::ld); the next item's first PV (accumulate = 0) is ordered behind the
      // p_ready arrive of its first block, which these same threads perform.
      tr.rec(12, cnt);
      tc_fence_before();
      if (stage) {
        fence_proxy_async_smem();   // staged O rows -> visible to the TMA store
        mbar_arrive_a(BAR_I(o_staged, t));   // warp 3 stores the tile
        ++st_cnt;
        store_pending = true;
      }
      tr.rec(9, cnt);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, C::kTmemCols);
#ifdef BP_TRACE
  if (p.trace && threadIdx.x == 0) {
    uint64_t ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    p.trace[8 * kTraceRecs * 2 + 2 * blockIdx.x + 1] = ns;
  }
#endif
}

// Ticket counters of the dynamic scheduler, one per LAUNCH.  Every launch takes the next slot of a ring and zeroes
// it with a memset enqueued in front of the kernel on the same stream (a memset node when the stream is being
// captured: every replay of the graph re-arms its own counter).  Launches that may run concurrently therefore never
// share a counter, and a launch that aborted mid-way cannot leave state behind.  Eager launches recycle a ring of
// kSchedEager slots (two live launches would have to be kSchedEager launches apart to collide); launches recorded
// into CUDA graphs take slots of a second pool that is never recycled, because a graph may be replayed at any later
// time next to anything else -- when that pool is exhausted the call fails instead of aliasing.
constexpr unsigned kSchedEager = 4096, kSchedCapture = 12288;
__device__ unsigned int g_sched[kSchedEager + kSchedCapture];
static std::atomic<unsigned> g_next_eager{0}, g_next_capture{0};

static int sched_slot(cudaStream_t stream, unsigned int** out) {
  void* base = nullptr;
  if (cudaGetSymbolAddress(&base, g_sched) != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_fmha_fwd: cudaGetSymbolAddress(g_sched) failed");
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) {
    cudaGetLastError();
    cap = cudaStreamCaptureStatusNone;
  }
  unsigned idx;
  if (cap == cudaStreamCaptureStatusActive) {
    const unsigned n = g_next_capture.fetch_add(1);
    if (n >= kSchedCapture)
      return fail(BP_ERR_UNSUPPORTED,
                  "bp_fmha_fwd: more than %u attention launches have been recorded into CUDA graphs by this process; "
                  "the scheduler's per-launch counters for captured launches are exhausted", kSchedCapture);
    idx = kSchedEager + n;
  } else {
    idx = g_next_eager.fetch_add(1) % kSchedEager;
  }
  unsigned int* slot = static_cast<unsigned int*>(base) + idx;
  cudaError_t e = cudaMemsetAsync(slot, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_fmha_fwd: cudaMemsetAsync(scheduler counter): %s", cudaGetErrorString(e));
  }
  *out = slot;
  return BP_OK;
}

template <int DP, bool kBF16, bool kDrop>
int launch(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
           const Params& p, cudaStream_t stream) {
  using C = Cfg<DP>;
  auto kern = fmha_fwd_kernel<DP, kBF16, kDrop>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_fmha_fwd: cudaFuncSetAttribute(%u B smem): %s", C::kSmemBytes,
                cudaGetErrorString(e));
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.num_items < sms ? p.num_items : sms;
  Params pp = p;
  if (int rc = sched_slot(stream, &pp.sched)) return rc;
  kern<<<grid, C::kThreads, C::kSmemBytes, stream>>>(tmQ, tmK, tmV, tmO, pp);
  return check_launch("bp_fmha_fwd launch");
}

static int g_out_f32 = 0;   // bp_debug_set_fmha_out_f32

}  // namespace fmha
}  // namespace bp

// Debug hooks (not part of the public ABI).
//   bp_debug_set_trace: when set, CTA 0 of the next launches writes its timeline here (8 roles x kTraceRecs x 2
//     uint64; -DBP_TRACE builds only).  Used by benchmarks/trace_kernel.py.
//   bp_debug_set_fmha_out_f32: test mode of SURVEY.md §8c (T2): the next bp_fmha_fwd calls treat `out` as an fp32
//     tensor with the same element strides and store O before the final 16-bit rounding.
//   bp_debug_poison_fmha_sched: overwrite every scheduler counter with garbage (what an aborted launch could leave
//     behind under the old per-stream scheme); tests check that later launches are unaffected.
namespace bp { uint64_t* g_trace = nullptr; }
extern "C" void bp_debug_set_trace(void* buf) { bp::g_trace = static_cast<uint64_t*>(buf); }
extern "C" void bp_debug_set_fmha_out_f32(int on) { bp::fmha::g_out_f32 = on ? 1 : 0; }
extern "C" int bp_debug_poison_fmha_sched(void* stream) {
  void* base = nullptr;
  if (cudaGetSymbolAddress(&base, bp::fmha::g_sched) != cudaSuccess) {
    cudaGetLastError();
    return bp::fail(BP_ERR_CUDA, "bp_debug_poison_fmha_sched: cudaGetSymbolAddress failed");
  }
  cudaError_t e = cudaMemsetAsync(base, 0x5a, sizeof(unsigned int) * (bp::fmha::kSchedEager + bp::fmha::kSchedCapture),
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return bp::fail(BP_ERR_CUDA, "bp_debug_poison_fmha_sched: %s", cudaGetErrorString(e));
  }
  return BP_OK;
}

extern "C" int64_t bp_fmha_fwd_dropout_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_k);

static int fmha_fwd_impl(const void* q, const void* k, const void* v, void* out, float* softmax_lse,
                         const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int32_t batch,
                         int32_t nheads, int32_t headdim, int32_t total_q, int32_t total_k,
                         int32_t max_seqlen_q, int32_t max_seqlen_k, int64_t q_row_stride,
                         int64_t q_head_stride, int64_t k_row_stride, int64_t k_head_stride,
                         int64_t v_row_stride, int64_t v_head_stride, int64_t o_row_stride,
                         int64_t o_head_stride, int32_t lse_stride, float softmax_scale, int32_t is_causal,
                         int32_t dtype, float p_dropout, uint64_t seed, void* workspace, int64_t workspace_bytes,
                         void* stream) {
  using namespace bp;
  if (!(p_dropout >= 0.f) || p_dropout >= 1.f)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: p_dropout must be in [0, 1) (got %f)", (double)p_dropout);
  const int thr = drop_threshold(p_dropout);
  const bool drop = thr > 0;
  const int s_pad_k = (max_seqlen_k > 0 ? max_seqlen_k + 127 : 127) / 128 * 128;
  if (drop) {
    const int64_t need = bp_fmha_fwd_dropout_workspace_bytes(batch, nheads, max_seqlen_k);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) % 16) != 0 || workspace_bytes < need)
      return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: dropout needs a 16-byte aligned workspace of %lld bytes (got %lld)",
                  (long long)need, (long long)workspace_bytes);
  }
  if (!q || !k || !v || !out || !softmax_lse || !cu_seqlens_q || !cu_seqlens_k)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: null pointer argument");
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: only fp16 and bf16 are supported (dtype=%d)", dtype);
  if (batch <= 0 || nheads <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: batch and nheads must be positive");
  if (headdim <= 0 || headdim % 8 != 0 || headdim > 128)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: head dim must be a multiple of 8 and <= 128 (got %d)", headdim);
  if (total_q <= 0 || total_k <= 0 || max_seqlen_q <= 0 || max_seqlen_k <= 0)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: empty input (total_q=%d total_k=%d)", total_q, total_k);
  if (lse_stride < max_seqlen_q)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: lse_stride %d < max_seqlen_q %d", lse_stride, max_seqlen_q);
  const int64_t strides[] = {q_row_stride, q_head_stride, k_row_stride, k_head_stride,
                             v_row_stride, v_head_stride, o_row_stride, o_head_stride};
  for (int64_t s : strides)
    if (s <= 0 || s % 8 != 0)
      return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: row/head strides must be positive multiples of 8 elements (got %lld)",
                  (long long)s);
  const uintptr_t ptrs[] = {(uintptr_t)q, (uintptr_t)k, (uintptr_t)v, (uintptr_t)out};
  for (uintptr_t a : ptrs)
    if (a % 16 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: q/k/v/out must be 16-byte aligned");

  const int DP = headdim <= 64 ? 64 : 128;
  const int BN = DP == 64 ? 128 : 64;
  CUtensorMap tmQ, tmK, tmV, tmO;
  {
    const uint64_t dq[3] = {(uint64_t)headdim, (uint64_t)nheads, (uint64_t)total_q};
    const uint64_t sq[2] = {(uint64_t)q_head_stride * 2, (uint64_t)q_row_stride * 2};
    const uint32_t bq[3] = {64, 1, (uint32_t)fmha::BM};
    if (int rc = encode_tensor_map(&tmQ, dtype, 3, q, dq, sq, bq, true)) return rc;
    const uint64_t dk[3] = {(uint64_t)headdim, (uint64_t)nheads, (uint64_t)total_k};
    const uint64_t sk[2] = {(uint64_t)k_head_stride * 2, (uint64_t)k_row_stride * 2};
    const uint32_t bk[3] = {64, 1, (uint32_t)BN};
    if (int rc = encode_tensor_map(&tmK, dtype, 3, k, dk, sk, bk, true)) return rc;
    const uint64_t sv[2] = {(uint64_t)v_head_stride * 2, (uint64_t)v_row_stride * 2};
    if (int rc = encode_tensor_map(&tmV, dtype, 3, v, dk, sv, bk, true)) return rc;
    const uint64_t so[2] = {(uint64_t)o_head_stride * 2, (uint64_t)o_row_stride * 2};
    if (int rc = encode_tensor_map(&tmO, dtype, 3, out, dq, so, bq, true)) return rc;
  }
  fmha::Params p;
  p.out = out;
  p.lse = softmax_lse;
  p.cu_q = cu_seqlens_q;
  p.cu_k = cu_seqlens_k;
  p.o_row_stride = o_row_stride;
  p.o_head_stride = o_head_stride;
  p.lse_stride = lse_stride;
  p.batch = batch;
  p.nheads = nheads;
  p.headdim = headdim;
  p.num_pairs = (max_seqlen_q + 2 * fmha::BM - 1) / (2 * fmha::BM);
  if ((int64_t)p.num_pairs * batch * nheads > 0x7fffffff)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_fmha_fwd: too many work items");
  p.num_items = p.num_pairs * batch * nheads;
  p.is_causal = is_causal ? 1 : 0;
  p.out_f32 = fmha::g_out_f32;
  p.sched = nullptr;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * fmha::kLog2e;
  p.trace = g_trace;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bf16 = dtype == BP_DTYPE_BF16;
  p.drop_c = static_cast<const uint32_t*>(workspace);
  p.drop_seed = seed;
  p.s_pad_k = s_pad_k;
  p.drop_thr24 = static_cast<uint32_t>(thr) << 24;
  p.drop_scale = drop ? 256.f / static_cast<float>(256 - thr) : 1.f;
  if (drop) {
    if (int rc = launch_drop_col_table(static_cast<uint32_t*>(workspace), seed, batch * nheads, s_pad_k, st)) return rc;
    if (DP == 64)
      return bf16 ? fmha::launch<64, true, true>(tmQ, tmK, tmV, tmO, p, st)
                  : fmha::launch<64, false, true>(tmQ, tmK, tmV, tmO, p, st);
    return bf16 ? fmha::launch<128, true, true>(tmQ, tmK, tmV, tmO, p, st)
                : fmha::launch<128, false, true>(tmQ, tmK, tmV, tmO, p, st);
  }
  if (DP == 64)
    return bf16 ? fmha::launch<64, true, false>(tmQ, tmK, tmV, tmO, p, st)
                : fmha::launch<64, false, false>(tmQ, tmK, tmV, tmO, p, st);
  return bf16 ? fmha::launch<128, true, false>(tmQ, tmK, tmV, tmO, p, st)
              : fmha::launch<128, false, false>(tmQ, tmK, tmV, tmO, p, st);
}

extern "C" int64_t bp_fmha_fwd_dropout_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_k) {
  if (batch <= 0 || nheads <= 0 || max_seqlen_k <= 0) return 0;
  return static_cast<int64_t>(batch) * nheads * ((static_cast<int64_t>(max_seqlen_k) + 127) / 128 * 128) * 4;
}

extern "C" int bp_fmha_fwd(const void* q, const void* k, const void* v, void* out, float* softmax_lse,
                           const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int32_t batch,
                           int32_t nheads, int32_t headdim, int32_t total_q, int32_t total_k,
                           int32_t max_seqlen_q, int32_t max_seqlen_k, int64_t q_row_stride,
                           int64_t q_head_stride, int64_t k_row_stride, int64_t k_head_stride,
                           int64_t v_row_stride, int64_t v_head_stride, int64_t o_row_stride,
                           int64_t o_head_stride, int32_t lse_stride, float softmax_scale, int32_t is_causal,
                           int32_t dtype, void* stream) {
  return fmha_fwd_impl(q, k, v, out, softmax_lse, cu_seqlens_q, cu_seqlens_k, batch, nheads, headdim, total_q, total_k,
                       max_seqlen_q, max_seqlen_k, q_row_stride, q_head_stride, k_row_stride, k_head_stride, v_row_stride,
                       v_head_stride, o_row_stride, o_head_stride, lse_stride, softmax_scale, is_causal, dtype, 0.f, 0,
                       nullptr, 0, stream);
}

extern "C" int bp_fmha_fwd_dropout(const void* q, const void* k, const void* v, void* out, float* softmax_lse,
                                   const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int32_t batch,
                                   int32_t nheads, int32_t headdim, int32_t total_q, int32_t total_k,
                                   int32_t max_seqlen_q, int32_t max_seqlen_k, int64_t q_row_stride,
                                   int64_t q_head_stride, int64_t k_row_stride, int64_t k_head_stride,
                                   int64_t v_row_stride, int64_t v_head_stride, int64_t o_row_stride,
                                   int64_t o_head_stride, int32_t lse_stride, float softmax_scale, int32_t is_causal,
                                   int32_t dtype, float p_dropout, uint64_t seed, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  return fmha_fwd_impl(q, k, v, out, softmax_lse, cu_seqlens_q, cu_seqlens_k, batch, nheads, headdim, total_q, total_k,
                       max_seqlen_q, max_seqlen_k, q_row_stride, q_head_stride, k_row_stride, k_head_stride, v_row_stride,
                       v_head_stride, o_row_stride, o_head_stride, lse_stride, softmax_scale, is_causal, dtype, p_dropout,
                       seed, workspace, workspace_bytes, stream);
}
