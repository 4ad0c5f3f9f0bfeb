// Backpack sense-mix for sm_100a:
//     out[b,i,:] = sum_l sum_{j<=i} softmax_j(scale * q_li . k_lj) * C_l(x_j)[:]
//
// The reference runs this as un-fused PyTorch (training/src/models/backpack.py:111-122 + :313): it
// materialises alpha (b,nv,s,s) and alpha@C (b,nv,s,d) in HBM (~17.8 GB of traffic at config 3) and does
// the dense (non-causal) batched GEMM.  Here it is two tcgen05 kernels that never materialise alpha:
//
//   pass 1  sense_lse_kernel : per (b, l, i) log-sum-exp of the causal scores (S = Q_l K_l^T on the tensor
//           core, online max/sum by one thread per row).  4 B per (b,l,i) of output.
//   pass 2  sense_mix_kernel : recomputes S, forms the *normalised* probabilities P = exp2(S*c - lse) (so
//           no running max, no rescaling, and every sense can be added into ONE accumulator), and issues
//           O += P C_l over all senses l and causal key blocks j into a single TMEM accumulator.
//
// Normalising per sense needs the row statistics before the first P.C product, hence two passes; pass 1
// costs 1/17 of the MMA work.  The accumulator of a 128-row tile at d=768 is 384 KB fp32 -- more than
// the 256 KB of TMEM -- so one CTA owns a 384-column chunk of the output (O: 384 TMEM columns, S: 2 x 64).
// Layout per CTA (384 threads): warp 0 = TMA producer for C, warp 3 = TMA producer for Q_l / K_l (+ TMEM
// allocator), warp 1 = issuer of the P.C products, warp 2 = issuer of the S = Q K^T products, warpgroups 1/2 =
// softmax for even/odd steps + epilogue.  A satisfied mbarrier wait still costs ~200 cycles of latency on
// the single issuing thread, so each MMA issuer waits on exactly ONE barrier per step: the barrier that the
// TMA producer arms for the operand tile also counts the 128 softmax threads that hand over S / P
// ("s_go" = K tile landed + S buffer drained, "pv_go" = C tile landed + P stored).
// C_l(x_j) tiles are consumed as MN-major B operands exactly as TMA wrote them (no transpose).
#include <stdlib.h>

#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace sense {

#ifndef BP_SENSE_DC
#define BP_SENSE_DC 384
#endif
#ifndef BP_SENSE_P_SMEM
#define BP_SENSE_P_SMEM 0
#endif
constexpr int BM = 128;
constexpr int kThreads = 384;
constexpr float kLog2e = 1.4426950408889634f;

// =============================================================================================
// pass 1: row statistics
// =============================================================================================
constexpr int kLseSenses = 4;   // senses per CTA of pass 1 (a CTA per sense spent more time starting up than working)

template <int PK>  // 64-column panels covering dk
struct LseCfg {
  static constexpr int BN = 128;
  static constexpr int kStages = PK == 3 ? 2 : 4;
  static constexpr int QB = PK == 1 ? 2 : 1;     // Q buffers across senses
  static constexpr uint32_t kQTileBytes = BM * 128 * PK;
  static constexpr uint32_t kKTileBytes = BN * 128 * PK;
  static constexpr uint32_t offQ = 0;                                  // [QB][2 tiles]
  static constexpr uint32_t offK = offQ + QB * 2 * kQTileBytes;
  static constexpr uint32_t offBar = offK + kStages * kKTileBytes;
  static constexpr uint32_t kSmemBytes = offBar + 256 + 1024;
  static constexpr uint32_t kTmemCols = 512;  // S_t[buf] at (t*2+buf)*128
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct LseBarriers {
  uint64_t q_full[2], q_empty[2];
  uint64_t k_full[4], k_empty[4];
  uint64_t s_full[2][2], s_free[2][2];
  uint32_t tmem_base;
};

struct LseParams {
  float* lse;  // (b, nv, s)
  int32_t seqlen, nv, dk, ksteps, num_pairs;
  int32_t senses_per_cta;   // <= kLseSenses; fewer when the grid would not fill the GPU
  float scale, scale_log2;
};

// One CTA = two 128-row query tiles of one batch element x kLseSenses consecutive senses.  Barrier phases, the K
// ring and the S buffers run on across senses (cumulative counters), Q is double-buffered when it fits, so the
// next sense's loads and first S overlap the tail of the current one.
template <int PK, bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
sense_lse_kernel(const __grid_constant__ CUtensorMap tmQK, const LseParams p) {
  using C = LseCfg<PK>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  LseBarriers& bars = *reinterpret_cast<LseBarriers*>(smem + C::offBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = p.num_pairs - 1 - static_cast<int>(blockIdx.x);
  const int sense0 = blockIdx.y * p.senses_per_cta, batch = blockIdx.z;
  const int n_senses = min(p.senses_per_cta, p.nv - sense0);
  const int S = p.seqlen;
  const int row0 = pair * 2 * BM;
  int n_blk[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = row0 + t * BM;
    n_blk[t] = r0 < S ? (min(S, r0 + BM) + BN - 1) / BN : 0;
  }
  const int n_max = max(n_blk[0], n_blk[1]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    for (int i = 0; i < 2; ++i) mbar_init(&bars.q_full[i], 1), mbar_init(&bars.q_empty[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bars.k_full[i], 1), mbar_init(&bars.k_empty[i], 1);
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i < 2; ++i) mbar_init(&bars.s_full[t][i], 1), mbar_init(&bars.s_free[t][i], 128);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars.tmem_base, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  const int tok0 = batch * S;  // first row of this batch in the flattened (b*s) token dimension

  if (warp < 4) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ---- producer: per sense the Q tiles of both query tiles (coordinate 1 = which*nv + sense), K ring ----
      const int n_q_tiles = (row0 + BM < S) ? 2 : 1;
      int kb = 0;   // K tiles loaded so far (all senses)
      for (int si = 0; si < n_senses; ++si) {
        const int sense = sense0 + si;
        const int qb = si % C::QB;
        if (si >= C::QB) mbar_wait(&bars.q_empty[qb], ((si / C::QB) - 1) & 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&bars.q_full[qb], n_q_tiles * C::kQTileBytes);
          for (int t = 0; t < n_q_tiles; ++t)
            for (int pn = 0; pn < PK; ++pn)
              tma_load_3d(smem + C::offQ + (qb * 2 + t) * C::kQTileBytes + pn * (BM * 128), &tmQK, &bars.q_full[qb],
                          pn * 64, sense, tok0 + row0 + t * BM);
        }
        for (int j = 0; j < n_max; ++j, ++kb) {
          const int slot = kb % C::kStages;
          if (kb >= C::kStages) mbar_wait(&bars.k_empty[slot], ((kb / C::kStages) - 1) & 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(&bars.k_full[slot], C::kKTileBytes);
            for (int pn = 0; pn < PK; ++pn)
              tma_load_3d(smem + C::offK + slot * C::kKTileBytes + pn * (BN * 128), &tmQK, &bars.k_full[slot], pn * 64,
                          p.nv + sense, tok0 + j * BN);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ---- MMA issuer: S_t(j) into buffer (count of S tiles of tile t so far) & 1 ----
      constexpr uint32_t idesc = make_idesc(kBF16, BM, BN, false, false);
      const uint32_t sQ = smem_u32(smem + C::offQ), sK = smem_u32(smem + C::offK);
      int kb = 0;
      int sc[2] = {0, 0};   // S tiles issued per query tile (all senses)
      for (int si = 0; si < n_senses; ++si) {
        const int qb = si % C::QB;
        mbar_wait(&bars.q_full[qb], (si / C::QB) & 1);
        for (int j = 0; j < n_max; ++j, ++kb) {
          const int slot = kb % C::kStages;
          mbar_wait(&bars.k_full[slot], (kb / C::kStages) & 1);
          const int last_user = j < n_blk[1] ? 1 : 0;
          for (int t = 0; t < 2; ++t) {
            if (j >= n_blk[t]) continue;
            const int c = sc[t]++;
            if (c >= 2) mbar_wait(&bars.s_free[t][c & 1], ((c >> 1) - 1) & 1);
            tc_fence_after();
            if (lane == 0) {
              for (int kk = 0; kk < p.ksteps; ++kk) {
                const uint32_t a = sQ + (qb * 2 + t) * C::kQTileBytes + (kk >> 2) * (BM * 128) + (kk & 3) * 32;
                const uint32_t b = sK + slot * C::kKTileBytes + (kk >> 2) * (BN * 128) + (kk & 3) * 32;
                umma_ss(tmem_base + (t * 2 + (c & 1)) * BN, make_smem_desc_sw128(a, 16, 1024),
                        make_smem_desc_sw128(b, 16, 1024), idesc, kk > 0 ? 1u : 0u);
              }
              if (t == last_user) umma_commit(&bars.k_empty[slot]);
              if (j == n_max - 1 && t == last_user) umma_commit(&bars.q_empty[qb]);   // last S of this sense
              umma_commit(&bars.s_full[t][c & 1]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    reg_alloc<224>();
    const int t = (warp >> 2) - 1;
    const int r = (warp & 3) * 32 + lane;
    const int n = n_blk[t];
    if (n > 0) {
      const int qrow = row0 + t * BM + r;
      const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const float c2 = p.scale_log2;
      int c = 0;   // S tiles consumed (all senses)
      for (int si = 0; si < n_senses; ++si) {
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < n; ++j, ++c) {
          mbar_wait(&bars.s_full[t][c & 1], (c >> 1) & 1);
          tc_fence_after();
          float s[BN];
          const uint32_t tS = tmem_base + lane_addr + (t * 2 + (c & 1)) * BN;
#pragma unroll
          for (int cc = 0; cc < BN / 32; ++cc) {
            uint32_t u[32];
            tmem_ld32(tS + cc * 32, u);
#pragma unroll
            for (int i = 0; i < 32; ++i) s[cc * 32 + i] = __uint_as_float(u[i]);
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&bars.s_free[t][c & 1]);
          const int col0 = j * BN;
          if (col0 + BN - 1 > row0 + t * BM) {  // block touches the diagonal (also covers cols >= seqlen)
#pragma unroll
            for (int cc = 0; cc < BN; ++cc)
              if (col0 + cc > qrow) s[cc] = -INFINITY;
          }
          float mx = s[0];
#pragma unroll
          for (int cc = 1; cc < BN; ++cc) mx = fmaxf(mx, s[cc]);
          const float m_new = fmaxf(m, mx);  // column 0 is always visible, so m_new is finite
          const float neg = -m_new * c2;
          float sum = 0.f;
#pragma unroll
          for (int cc = 0; cc < BN; ++cc) sum += fast_exp2(fmaf(s[cc], c2, neg));
          l = l * fast_exp2((m - m_new) * c2) + sum;
          m = m_new;
        }
        if (qrow < S)
          p.lse[(static_cast<int64_t>(batch) * p.nv + sense0 + si) * S + qrow] = m * p.scale + logf(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// =============================================================================================
// pass 2: O += P_l C_l over senses and key blocks
// =============================================================================================
template <int PK>
struct MixCfg {
  static constexpr int BN = 64;    // keys per step
  // Output columns per CTA.  The fp32 accumulator of a 128-row tile at d = 768 (384 KB) exceeds TMEM, so a CTA owns
  // a column chunk and recomputes S (and the exponentials) for it.  384 columns (2 chunks) leave room for ONE S
  // buffer only; ncu shows the tensor pipe 60 % active: S(n+1) queues behind the 768-cycle PV(n-1) in the in-order
  // pipe and the softmax of a step (~1000 cycles of latency) is longer than one PV.  256 columns (-DBP_SENSE_DC=256:
  // 3 chunks, TWO S buffers, S two steps ahead of PV) was measured SLOWER (1.21 vs 0.99 ms): every chunk repeats
  // the exponentials, and with a 512-cycle PV per step the MUFU pipe becomes the bound.
  // -DBP_SENSE_P_SMEM=1 keeps P in shared memory instead (two 16 KB swizzled tiles, SS product as in the attention
  // kernel): that frees the 64 TMEM columns of the P buffers for a SECOND S buffer at 384 columns, so S runs two
  // steps ahead of PV without a third chunk.  Measured: 1.11 ms vs 1.02 ms with P in TMEM -- off by default.
  static constexpr int DC = BP_SENSE_DC;
  static constexpr bool kPSmem = BP_SENSE_P_SMEM != 0 && PK == 1;
  static constexpr int SB = (kPSmem ? DC + 2 * BN <= 512 : DC + 2 * BN + 2 * (BN / 2) <= 512) ? 2 : 1;   // S buffers
  static constexpr int QS = PK == 1 ? 2 : 1;
  static constexpr int KS = 2;
  static constexpr int CS = PK == 1 ? (DC <= 256 ? 4 : 3) : 2;
  static constexpr uint32_t kQTileBytes = BM * 128 * PK;
  static constexpr uint32_t kKTileBytes = BN * 128 * PK;
  static constexpr uint32_t kCPanelBytes = BN * 128;           // 64 keys x 64 columns
  static constexpr uint32_t kCTileBytes = (DC / 64) * kCPanelBytes;
  static constexpr uint32_t offQ = 0;
  static constexpr uint32_t offK = offQ + QS * kQTileBytes;
  static constexpr uint32_t offC = offK + KS * kKTileBytes;
  static constexpr uint32_t offP = offC + CS * kCTileBytes;            // [2] P tiles (128 rows x 64 keys), kPSmem only
  static constexpr uint32_t kPTileBytes = BM * 128;
  static constexpr uint32_t offBar = offP + (kPSmem ? 2 * kPTileBytes : 0);
  static constexpr uint32_t kSmemBytes = offBar + 256 + 1024;
  // TMEM: O accumulator, one S buffer (handed back as soon as the softmax warps hold it in registers) and
  // two P buffers: P (bf16/f16, two values per 32-bit column) is the A operand of the PV product and is
  // read by the tensor core straight from TMEM -- it never touches shared memory, whose bandwidth is
  // what bounds this kernel (every K-step already streams a 64 x 384 slice of C through it).
  static constexpr uint32_t colO = 0, colS = DC, colP = DC + SB * BN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kPSmem ? colP <= 512 : colP + 2 * (BN / 2) <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// Step order of pass 2.  A step is one (sense l, 64-key block j) pair; the n_steps = nv * nj steps of a query tile
// are walked key-GROUP outermost (kGroup blocks = 256 keys), then sense, then block within the group:
//     for g: for l: for j in group g: step(l, j)
// With the sense outermost (the first version) the query tiles of a batch element drift apart -- a tile with 2
// key blocks is done with sense l after 2 steps, the tile with 16 blocks after 16 -- so the CTAs that run together
// touch all 25 MB of that batch element's sense vectors and every C tile is re-read from DRAM by most of the 8
// query tiles that need it (measured: 4.2 GB per launch against 1.9 GB algorithmic).  Group-outermost, all query
// tiles of a batch element need the same 256-key slab (6 MB) at the same time, which stays in L2; Q_l is
// reloaded once per (group, sense) instead of once per sense (4 KB per step on average).
// Wide sense keys (dk > 64: few senses, single-buffered Q) keep the sense outermost (group = all blocks): their
// per-batch-element footprint is small and a Q reload every few steps would stall the pipeline (measured at k = 4:
// 165 -> 229 us with groups of 4).
#ifndef BP_SENSE_GROUP
#define BP_SENSE_GROUP 4
#endif
constexpr int kGroup = BP_SENSE_GROUP;
struct StepIter {
  int nj, nv, grp, g0, gsz, l, jj;
  __device__ __forceinline__ StepIter(int nj_, int nv_, int grp_)
      : nj(nj_), nv(nv_), grp(grp_), g0(0), gsz(min(grp_, nj_)), l(0), jj(0) {}
  __device__ __forceinline__ int sense() const { return l; }
  __device__ __forceinline__ int j() const { return g0 + jj; }
  __device__ __forceinline__ bool first_of_visit() const { return jj == 0; }        // first block of this (group, sense)
  __device__ __forceinline__ bool last_of_visit() const { return jj == gsz - 1; }
  __device__ __forceinline__ void next() {
    if (++jj == gsz) {
      jj = 0;
      if (++l == nv) {
        l = 0;
        g0 += gsz;
        gsz = min(grp, nj - g0);
      }
    }
  }
};

struct MixBarriers {
  uint64_t q_full[2], q_empty[2];
  uint64_t s_go[2], k_empty[2];    // s_go[n & 1]: K(n) landed (tx) + S(n-1) drained by its 128 softmax threads
  uint64_t pv_go[4], c_empty[4];   // pv_go[n % CS]: C(n) landed (tx) + P(n) stored by its 128 softmax threads
  uint64_t s_full[2], p_free[2];   // s_full[n & 1]: each warpgroup must see every phase
  uint64_t o_full;
  uint32_t tmem_base;
};

struct MixParams {
  const float* lse;  // (b, nv, s), natural log
  void* out;         // (b, s, d)
  int32_t seqlen, nv, dk, ksteps, d, num_qtiles, num_chunks;
  int32_t c_sense_inner;  // content tensor-map dims are (d, nv, s, b) instead of (d, s, nv, b)
  int32_t group;          // key blocks per group of the step order (StepIter)
  float scale_log2;
  uint64_t* trace;        // debug timeline (BP_TRACE builds), else null
};

template <int PK, bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
sense_mix_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmC, const MixParams p) {
  using C = MixCfg<PK>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  MixBarriers& bars = *reinterpret_cast<MixBarriers*>(smem + C::offBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = p.num_qtiles - 1 - static_cast<int>(blockIdx.x) / p.num_chunks;  // heaviest first
  const int chunk = static_cast<int>(blockIdx.x) % p.num_chunks;
  const int batch = blockIdx.y;
  const int S = p.seqlen;
  const int row0 = qtile * BM;
  const int nj = (min(S, row0 + BM) + BN - 1) / BN;  // causal key blocks of this query tile
  const int n_steps = p.nv * nj;                      // walked in StepIter order
  const int col_base = chunk * C::DC;
  const int ncols = min(C::DC, p.d - col_base);       // multiple of 64
  const int n1 = min(ncols, 256), n2 = ncols - n1;    // the two MMA N extents per k-step

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.q_full[i], 1), mbar_init(&bars.q_empty[i], 1);
      mbar_init(&bars.s_go[i], 129), mbar_init(&bars.k_empty[i], 1);
      mbar_init(&bars.s_full[i], 1), mbar_init(&bars.p_free[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bars.pv_go[i], 129), mbar_init(&bars.c_empty[i], 1);
    mbar_init(&bars.o_full, 1);
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
  const int tok0 = batch * S;

  if (warp < 4) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ---- producer A: content tiles C_l[j] : (ncols/64) panels of [64 keys x 64 columns] ----
      Tracer tr(p.trace, 0, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0);
      StepIter it(nj, p.nv, p.group);
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int slot = n % C::CS;
        tr.rec(0, n);
        if (n >= C::CS) mbar_wait(&bars.c_empty[slot], ((n / C::CS) - 1) & 1);
        tr.rec(1, n);
        if (lane == 0) {
          const int sense = it.sense(), j = it.j();
          mbar_arrive_expect_tx(&bars.pv_go[slot], (ncols / 64) * C::kCPanelBytes);
          for (int pn = 0; pn < ncols / 64; ++pn) {
            uint8_t* dst = smem + C::offC + slot * C::kCTileBytes + pn * C::kCPanelBytes;
            if (p.c_sense_inner)
              tma_load_4d(dst, &tmC, &bars.pv_go[slot], col_base + pn * 64, sense, j * BN, batch);
            else
              tma_load_4d(dst, &tmC, &bars.pv_go[slot], col_base + pn * 64, j * BN, sense, batch);
          }
        }
        __syncwarp();
      }
    } else if (warp == 3) {
      // ---- producer B: Q_l (once per sense) and K_l[j] ----
      StepIter it(nj, p.nv, p.group);
      int qv = 0;   // (group, sense) visits so far: Q buffer = qv % QS
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int sense = it.sense(), j = it.j();
        if (it.first_of_visit()) {
          const int qs = qv % C::QS;
          if (qv >= C::QS) mbar_wait(&bars.q_empty[qs], ((qv / C::QS) - 1) & 1);
          ++qv;
          if (lane == 0) {
            mbar_arrive_expect_tx(&bars.q_full[qs], C::kQTileBytes);
            for (int pn = 0; pn < PK; ++pn)
              tma_load_3d(smem + C::offQ + qs * C::kQTileBytes + pn * (BM * 128), &tmQ, &bars.q_full[qs], pn * 64,
                          sense, tok0 + row0);
          }
        }
        const int slot = n % C::KS;
        if (n >= C::KS) mbar_wait(&bars.k_empty[slot], ((n / C::KS) - 1) & 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&bars.s_go[slot], C::kKTileBytes);
          for (int pn = 0; pn < PK; ++pn)
            tma_load_3d(smem + C::offK + slot * C::kKTileBytes + pn * (BN * 128), &tmK, &bars.s_go[slot], pn * 64,
                        p.nv + sense, tok0 + j * BN);
        }
        __syncwarp();
      }
    } else if (warp == 2) {
      // ---- issuer of S(n) = Q_l K_l[j]^T into the single S buffer ----
      constexpr uint32_t idesc_s = make_idesc(kBF16, BM, BN, false, false);
      const uint32_t sQ = smem_u32(smem + C::offQ), sK = smem_u32(smem + C::offK);
      Tracer tr(p.trace, 4, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0);
      StepIter it(nj, p.nv, p.group);
      int qv = 0;
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int qs = qv % C::QS, ks = n & 1;
        tr.rec(0, n);
        if (it.first_of_visit()) mbar_wait(&bars.q_full[qs], (qv / C::QS) & 1);
        mbar_wait(&bars.s_go[ks], (n >> 1) & 1);
        tc_fence_after();
        tr.rec(1, n);
        if (lane == 0) {
          for (int kk = 0; kk < p.ksteps; ++kk) {
            const uint32_t a = sQ + qs * C::kQTileBytes + (kk >> 2) * (BM * 128) + (kk & 3) * 32;
            const uint32_t b = sK + ks * C::kKTileBytes + (kk >> 2) * (BN * 128) + (kk & 3) * 32;
            umma_ss(tmem_base + C::colS + (C::SB == 2 ? (n & 1) * BN : 0), make_smem_desc_sw128(a, 16, 1024),
                    make_smem_desc_sw128(b, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&bars.k_empty[ks]);
          if (it.last_of_visit()) umma_commit(&bars.q_empty[qs]);
          umma_commit(&bars.s_full[n & 1]);
        }
        if (it.last_of_visit()) ++qv;
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---- issuer of O += P(n) C_l[j]: P from TMEM (A operand), C tile MN-major from shared memory ----
      const uint32_t idesc_pv1 = make_idesc(kBF16, BM, n1, false, true);
      const uint32_t idesc_pv2 = make_idesc(kBF16, BM, n2 > 0 ? n2 : 64, false, true);
      const uint32_t sC = smem_u32(smem + C::offC);
      const uint32_t sP = smem_u32(smem + C::offP);
      Tracer tr(p.trace, 1, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0);
      bool ready = false;   // result of the early probe of pv_go for this step
      for (int n = 0; n < n_steps; ++n) {
        const int cs = n % C::CS;
        tr.rec(3, n);
        if (!ready) mbar_wait(&bars.pv_go[cs], (n / C::CS) & 1);
        tc_fence_after();
        ready = mbar_test(&bars.pv_go[(n + 1) % C::CS], ((n + 1) / C::CS) & 1);   // latency hidden by the issue below
        tr.rec(5, n);
        if (lane == 0) {
          const uint32_t a_tmem = tmem_base + C::colP + (n & 1) * (BN / 2);   // P(n): 8 columns per K-step of 16
          const uint32_t b_base = sC + cs * C::kCTileBytes;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            if constexpr (C::kPSmem) {
              const uint64_t a_desc = make_smem_desc_sw128(sP + (n & 1) * C::kPTileBytes + kk * 32, 16, 1024);
              umma_ss(tmem_base + C::colO, a_desc, make_smem_desc_sw128(b_base + kk * 2048, C::kCPanelBytes, 1024),
                      idesc_pv1, (n > 0 || kk > 0) ? 1u : 0u);
              if (n2 > 0)
                umma_ss(tmem_base + C::colO + 256, a_desc,
                        make_smem_desc_sw128(b_base + 4 * C::kCPanelBytes + kk * 2048, C::kCPanelBytes, 1024),
                        idesc_pv2, (n > 0 || kk > 0) ? 1u : 0u);
            } else {
              umma_ts(tmem_base + C::colO, a_tmem + kk * 8,
                      make_smem_desc_sw128(b_base + kk * 2048, C::kCPanelBytes, 1024), idesc_pv1,
                      (n > 0 || kk > 0) ? 1u : 0u);
              if (n2 > 0)
                umma_ts(tmem_base + C::colO + 256, a_tmem + kk * 8,
                        make_smem_desc_sw128(b_base + 4 * C::kCPanelBytes + kk * 2048, C::kCPanelBytes, 1024),
                        idesc_pv2, (n > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&bars.c_empty[cs]);
          umma_commit(&bars.p_free[n & 1]);
          if (n == n_steps - 1) umma_commit(&bars.o_full);
        }
        __syncwarp();
        // Issuing blocks while the tensor pipe's queue is full (a timeline trace shows ~850 cycles for the eight
        // MMAs of a step), so P(n+1) has usually been stored by now: probe again rather than pay the ~240-cycle
        // round trip of a blocking wait on an already completed phase with the pipe draining.
        if (!ready) ready = mbar_test(&bars.pv_go[(n + 1) % C::CS], ((n + 1) / C::CS) & 1);
        tr.rec(6, n);
      }
    }
  } else {
    // ---- softmax warpgroups: w handles steps n = 2i + w ----
    reg_alloc<224>();
    const int w = (warp >> 2) - 1;
    const int r = (warp & 3) * 32 + lane;
    const int qrow = row0 + r;
    const bool valid = qrow < S;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + C::colS + (C::SB == 2 ? w * BN : 0);
    const uint32_t tP = tmem_base + lane_addr + C::colP + w * (BN / 2);
    const float c2 = p.scale_log2;
    const float* lse_row = p.lse + static_cast<int64_t>(batch) * p.nv * S + (valid ? qrow : S - 1);
    int cur_sense = -1;
    float neg_lse2 = 0.f;
    // the row statistic of the NEXT sense is fetched one sense ahead so its global-load latency never sits
    // on the softmax path (light query tiles change sense every couple of steps)
    float lse_next = __ldg(lse_row);
    int next_id = 0;   // sense whose statistic lse_next holds
    int i = 0;
    Tracer tr(p.trace, 2 + w, blockIdx.x == 0 && blockIdx.y == 0 && r == 0);
    // "S buffer free": with two S buffers each warpgroup owns one (S(n) needs S(n-2) drained); with one buffer
    // S(0) only needs "S(-1) drained" from the warpgroup of the odd steps
    if (C::SB == 2 || w == 1) mbar_arrive(&bars.s_go[C::SB == 2 ? w : 0]);
    StepIter it(nj, p.nv, p.group);
    if (w == 1 && n_steps > 1) it.next();
    for (int n = w; n < n_steps; n += 2, ++i) {
      const int sense = it.sense(), j = it.j();
      if (n + 2 < n_steps) { it.next(); it.next(); }
      if (sense != cur_sense) {
        cur_sense = sense;
        if (next_id != sense) lse_next = __ldg(lse_row + static_cast<int64_t>(sense) * S);   // skipped a sense
        neg_lse2 = -lse_next * kLog2e;
        next_id = sense + 1 == p.nv ? 0 : sense + 1;   // senses wrap around at the end of a key group
        lse_next = __ldg(lse_row + static_cast<int64_t>(next_id) * S);
      }
      tr.rec(0, n);
      mbar_wait(&bars.s_full[w], i & 1);
      tc_fence_after();
      tr.rec(1, n);
      float s[BN];
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t u[32];
        tmem_ld32(tS + c * 32, u);
#pragma unroll
        for (int k = 0; k < 32; ++k) s[c * 32 + k] = __uint_as_float(u[k]);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bars.s_go[C::SB == 2 ? (n & 1) : ((n + 1) & 1)]);   // the S buffer is free again
      tr.rec(2, n);
      const int col0 = j * BN;
      if (col0 + BN - 1 > row0) {
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (col0 + c > qrow) s[c] = -INFINITY;
      }
      uint32_t pk[BN / 2];
#pragma unroll
      for (int c = 0; c < BN / 2; ++c)
        pk[c] = pack2<kBF16>(fast_exp2(fmaf(s[2 * c], c2, neg_lse2)), fast_exp2(fmaf(s[2 * c + 1], c2, neg_lse2)));
      tr.rec(3, n);
      if (i >= 1) {
        mbar_wait(&bars.p_free[w], (i - 1) & 1);   // the PV product of step n-2 has consumed this P buffer
        tc_fence_after();
      }
      tr.rec(4, n);
      if constexpr (C::kPSmem) {
        uint8_t* prow = smem + C::offP + w * C::kPTileBytes;   // this warpgroup's P tile, 128B-swizzled rows
#pragma unroll
        for (int c8 = 0; c8 < BN / 8; ++c8)
          *reinterpret_cast<uint4*>(prow + sw128_offset(r, c8)) =
              make_uint4(pk[c8 * 4 + 0], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
        fence_proxy_async_smem();
      } else {
        tmem_st32(tP, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&bars.pv_go[n % C::CS]);
      tr.rec(5, n);
    }
    // ---- epilogue: warpgroup w stores columns [w*ncols/2, (w+1)*ncols/2) of its row ----
    mbar_wait(&bars.o_full, 0);
    tc_fence_after();
    const int half_cols = ncols / 2;  // multiple of 32
    const uint32_t tO = tmem_base + lane_addr + C::colO + w * half_cols;
    uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) +
                    2 * ((static_cast<int64_t>(tok0) + qrow) * p.d + col_base + w * half_cols);
    for (int c = 0; c < half_cols / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack2<kBF16>(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          v.y = pack2<kBF16>(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          v.z = pack2<kBF16>(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          v.w = pack2<kBF16>(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + (c * 32 + g * 8) * 2) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, C::kTmemCols);
}

// =============================================================================================
// pass 2 on CTA pairs (cta_group::2): the default when dk <= 64 and d is a multiple of 384.
//
// The 1-CTA kernel above is bound by shared-memory bandwidth: every 64-key step streams a 48 KB C tile into shared
// memory and out again into the tensor core (~148 KB of traffic per 864 MMA cycles against 128 B/clk).  Here two
// CTAs of a cluster take two consecutive query tiles of the same (batch, column chunk) and every MMA is 256 rows
// across the two SMs: each CTA keeps its own Q tile, S, P and O (128 rows) but only HALF of the K tile (32 keys)
// and half of the C tile (192 of the 384 columns), so fill + operand reads drop to ~76 KB per step.  Only the
// leader CTA (cluster rank 0) issues MMAs; both CTAs' TMA loads credit the leader's barriers, both CTAs' softmax
// threads arrive on them, and `tcgen05.commit ... multicast::cluster` releases buffers in both CTAs.  The lighter
// query tile of a pair computes (fully masked, P = 0) the two key blocks only its sibling needs: 11 % more MMA
// work at seq 1024.
// =============================================================================================
struct PairCfg {
  static constexpr int BN = 64;
  static constexpr int DC = 384;
  static constexpr int QS = 2, KS = 2, CS = 4;
  static constexpr uint32_t kQTileBytes = BM * 128;            // own 128 rows x 64 (padded dk)
  static constexpr uint32_t kKHalfBytes = (BN / 2) * 128;      // 32 keys
  static constexpr uint32_t kCPanelBytes = BN * 128;           // 64 keys x 64 columns
  static constexpr uint32_t kCHalfBytes = 3 * kCPanelBytes;    // this CTA's 192 columns
  static constexpr uint32_t offQ = 0;
  static constexpr uint32_t offK = offQ + QS * kQTileBytes;
  static constexpr uint32_t offC = offK + KS * 8192;           // (slots kept 1024-aligned)
  static constexpr uint32_t offBar = offC + CS * kCHalfBytes;
  static constexpr uint32_t kSmemBytes = offBar + 256 + 1024;
  static constexpr uint32_t colO = 0, colS = DC, colP = DC + BN;
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct PairBarriers {
  uint64_t q_full[2], q_empty[2];
  uint64_t s_go[2], k_empty[2];
  uint64_t pv_go[4], c_empty[4];
  uint64_t s_full[2], p_free[2];
  uint64_t o_full;
  uint32_t tmem_base;
};

template <bool kBF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
sense_mix_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmC, const MixParams p) {
  using C = PairCfg;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PairBarriers& bars = *reinterpret_cast<PairBarriers*>(smem + C::offBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = static_cast<int>(blockIdx.x) >> 1;
  const int num_qpairs = (p.num_qtiles + 1) / 2;
  const int qpair = num_qpairs - 1 - cluster_id / p.num_chunks;  // heaviest first
  const int chunk = cluster_id % p.num_chunks;
  const int batch = blockIdx.y;
  const int S = p.seqlen;
  const int row0 = (qpair * 2 + rank) * BM;                      // this CTA's query tile
  const int nj = (min(S, qpair * 2 * BM + 2 * BM) + BN - 1) / BN;  // causal key blocks of the PAIR
  const int n_steps = p.nv * nj;
  const int col_base = chunk * C::DC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.q_full[i], 1), mbar_init(&bars.q_empty[i], 1);
      mbar_init(&bars.s_go[i], 9), mbar_init(&bars.k_empty[i], 1);       // producer + one lane of 4 softmax warps x 2 CTAs
      mbar_init(&bars.s_full[i], 1), mbar_init(&bars.p_free[i], 1);
    }
    for (int i = 0; i < C::CS; ++i) mbar_init(&bars.pv_go[i], 9), mbar_init(&bars.c_empty[i], 1);
    mbar_init(&bars.o_full, 1);
    fence_barrier_init();
  }
  if (warp == 3) {
    tmem_alloc_2cta(&bars.tmem_base, C::kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  const int tok0 = batch * S;

  if (warp < 4) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ---- producer A (both CTAs): this CTA's three 64-column panels of C_l[j]; bytes credited to the leader ----
      StepIter it(nj, p.nv, p.group);
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int slot = n % C::CS;
        if (n >= C::CS) mbar_wait(&bars.c_empty[slot], ((n / C::CS) - 1) & 1);
        if (lane == 0) {
          const int sense = it.sense(), j = it.j();
          if (leader) mbar_arrive_expect_tx(&bars.pv_go[slot], 2 * C::kCHalfBytes);
          for (int pn = 0; pn < 3; ++pn) {
            // panels 0,1 = this CTA's half of columns [0,256); panel 2 = its half of columns [256,384)
            const int col = col_base + (pn < 2 ? rank * 128 + pn * 64 : 256 + rank * 64);
            uint8_t* dst = smem + C::offC + slot * C::kCHalfBytes + pn * C::kCPanelBytes;
            if (p.c_sense_inner)
              tma_load_4d_pair(dst, &tmC, &bars.pv_go[slot], col, sense, j * BN, batch);
            else
              tma_load_4d_pair(dst, &tmC, &bars.pv_go[slot], col, j * BN, sense, batch);
          }
        }
        __syncwarp();
      }
    } else if (warp == 3) {
      // ---- producer B (both CTAs): own Q_l tile (once per visit) and own half (32 keys) of K_l[j] ----
      StepIter it(nj, p.nv, p.group);
      int qv = 0;
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int sense = it.sense(), j = it.j();
        if (it.first_of_visit()) {
          const int qs = qv % C::QS;
          if (qv >= C::QS) mbar_wait(&bars.q_empty[qs], ((qv / C::QS) - 1) & 1);
          ++qv;
          if (lane == 0) {
            if (leader) mbar_arrive_expect_tx(&bars.q_full[qs], 2 * C::kQTileBytes);
            tma_load_3d_pair(smem + C::offQ + qs * C::kQTileBytes, &tmQ, &bars.q_full[qs], 0, sense, tok0 + row0);
          }
        }
        const int slot = n % C::KS;
        if (n >= C::KS) mbar_wait(&bars.k_empty[slot], ((n / C::KS) - 1) & 1);
        if (lane == 0) {
          if (leader) mbar_arrive_expect_tx(&bars.s_go[slot], 2 * C::kKHalfBytes);
          tma_load_3d_pair(smem + C::offK + slot * 8192, &tmK, &bars.s_go[slot], 0, p.nv + sense,
                           tok0 + j * BN + rank * (BN / 2));
        }
        __syncwarp();
      }
    } else if (warp == 2 && leader) {
      // ---- issuer of S(n) = Q_l K_l[j]^T for both CTAs (M = 256, N = 64) ----
      constexpr uint32_t idesc_s = make_idesc(kBF16, 256, BN, false, false);
      const uint32_t sQ = smem_u32(smem + C::offQ), sK = smem_u32(smem + C::offK);
      StepIter it(nj, p.nv, p.group);
      int qv = 0;
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int qs = qv % C::QS, ks = n & 1;
        if (it.first_of_visit()) mbar_wait(&bars.q_full[qs], (qv / C::QS) & 1);
        mbar_wait(&bars.s_go[ks], (n >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          for (int kk = 0; kk < p.ksteps; ++kk) {
            const uint32_t a = sQ + qs * C::kQTileBytes + kk * 32;
            const uint32_t b = sK + ks * 8192 + kk * 32;
            umma_ss_pair(tmem_base + C::colS, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(b, 16, 1024),
                         idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit_pair(&bars.k_empty[ks]);
          if (it.last_of_visit()) umma_commit_pair(&bars.q_empty[qs]);
          umma_commit_pair(&bars.s_full[n & 1]);
        }
        if (it.last_of_visit()) ++qv;
        __syncwarp();
      }
    } else if (warp == 1 && leader) {
      // ---- issuer of O += P(n) C_l[j] for both CTAs: P from TMEM, C half-tiles MN-major from shared memory ----
      constexpr uint32_t idesc_pv1 = make_idesc(kBF16, 256, 256, false, true);
      constexpr uint32_t idesc_pv2 = make_idesc(kBF16, 256, 128, false, true);
      const uint32_t sC = smem_u32(smem + C::offC);
      bool ready = false;
      for (int n = 0; n < n_steps; ++n) {
        const int cs = n % C::CS;
        if (!ready) mbar_wait(&bars.pv_go[cs], (n / C::CS) & 1);
        tc_fence_after();
        ready = mbar_test(&bars.pv_go[(n + 1) % C::CS], ((n + 1) / C::CS) & 1);
        if (lane == 0) {
          const uint32_t a_tmem = tmem_base + C::colP + (n & 1) * (BN / 2);
          const uint32_t b_base = sC + cs * C::kCHalfBytes;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            umma_ts_pair(tmem_base + C::colO, a_tmem + kk * 8,
                         make_smem_desc_sw128(b_base + kk * 2048, C::kCPanelBytes, 1024), idesc_pv1,
                         (n > 0 || kk > 0) ? 1u : 0u);
            umma_ts_pair(tmem_base + C::colO + 256, a_tmem + kk * 8,
                         make_smem_desc_sw128(b_base + 2 * C::kCPanelBytes + kk * 2048, C::kCPanelBytes, 1024),
                         idesc_pv2, (n > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit_pair(&bars.c_empty[cs]);
          umma_commit_pair(&bars.p_free[n & 1]);
          if (n == n_steps - 1) umma_commit_pair(&bars.o_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- softmax warpgroups (both CTAs, own 128 rows): w handles steps n = 2i + w ----
    reg_alloc<224>();
    const int w = (warp >> 2) - 1;
    const int r = (warp & 3) * 32 + lane;
    const int qrow = row0 + r;
    const bool valid = qrow < S;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + C::colS;
    const uint32_t tP = tmem_base + lane_addr + C::colP + w * (BN / 2);
    const float c2 = p.scale_log2;
    const float* lse_row = p.lse + static_cast<int64_t>(batch) * p.nv * S + (valid ? qrow : S - 1);
    int cur_sense = -1;
    float neg_lse2 = 0.f;
    float lse_next = __ldg(lse_row);
    int next_id = 0;
    int i = 0;
    // Hand-overs to the leader's barriers are ONE remote arrive per warp (after a warp sync), not one per thread:
    // 256 remote arrives per barrier and step made this kernel twice as slow as the 1-CTA one.
    if (w == 1 && lane == 0) mbar_arrive_leader(&bars.s_go[0]);   // "S(-1) drained"
    StepIter it(nj, p.nv, p.group);
    if (w == 1 && n_steps > 1) it.next();
    for (int n = w; n < n_steps; n += 2, ++i) {
      const int sense = it.sense(), j = it.j();
      if (n + 2 < n_steps) { it.next(); it.next(); }
      if (sense != cur_sense) {
        cur_sense = sense;
        if (next_id != sense) lse_next = __ldg(lse_row + static_cast<int64_t>(sense) * S);
        neg_lse2 = -lse_next * kLog2e;
        next_id = sense + 1 == p.nv ? 0 : sense + 1;
        lse_next = __ldg(lse_row + static_cast<int64_t>(next_id) * S);
      }
      mbar_wait(&bars.s_full[w], i & 1);
      tc_fence_after();
      float s[BN];
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t u[32];
        tmem_ld32(tS + c * 32, u);
#pragma unroll
        for (int k = 0; k < 32; ++k) s[c * 32 + k] = __uint_as_float(u[k]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars.s_go[(n + 1) & 1]);   // S(n+1) may be computed while the exponentials run
      const int col0 = j * BN;
      if (col0 + BN - 1 > row0) {
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (col0 + c > qrow) s[c] = -INFINITY;
      }
      uint32_t pk[BN / 2];
#pragma unroll
      for (int c = 0; c < BN / 2; ++c)
        pk[c] = pack2<kBF16>(fast_exp2(fmaf(s[2 * c], c2, neg_lse2)), fast_exp2(fmaf(s[2 * c + 1], c2, neg_lse2)));
      if (i >= 1) {
        mbar_wait(&bars.p_free[w], (i - 1) & 1);   // the PV product of step n-2 has consumed this P buffer
        tc_fence_after();
      }
      tmem_st32(tP, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars.pv_go[n % C::CS]);
    }
    // ---- epilogue: warpgroup w stores columns [w*192, (w+1)*192) of its rows ----
    mbar_wait(&bars.o_full, 0);
    tc_fence_after();
    constexpr int half_cols = C::DC / 2;
    const uint32_t tO = tmem_base + lane_addr + C::colO + w * half_cols;
    uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) +
                    2 * ((static_cast<int64_t>(tok0) + qrow) * p.d + col_base + w * half_cols);
    for (int c = 0; c < half_cols / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v;
          v.x = pack2<kBF16>(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          v.y = pack2<kBF16>(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          v.z = pack2<kBF16>(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          v.w = pack2<kBF16>(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + (c * 32 + g * 8) * 2) = v;
        }
      }
    }
  }
  // no CTA of the pair may exit (or free TMEM) while its peer can still read its shared memory / signal it
  tc_fence_before();
  cluster_sync_all();
  if (warp == 3) tmem_dealloc_2cta(tmem_base, C::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int make_qk_map(CUtensorMap* tm, const void* qk, int batch, int seqlen, int nv, int dk, int dtype,
                       int box_rows) {
  // qk (b, s, 2, nv, dk) viewed as [dk, 2*nv, b*s]; panels beyond dk are zero-filled by TMA
  const uint64_t dims[3] = {(uint64_t)dk, (uint64_t)2 * nv, (uint64_t)batch * seqlen};
  const uint64_t str[2] = {(uint64_t)dk * 2, (uint64_t)2 * nv * dk * 2};
  const uint32_t box[3] = {64, 1, (uint32_t)box_rows};
  return encode_tensor_map(tm, dtype, 3, qk, dims, str, box, true);
}

static int check_common(const char* fn, int batch, int seqlen, int nv, int dk, int dtype) {
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: only fp16 and bf16 are supported", fn);
  if (batch <= 0 || seqlen <= 0 || nv <= 0 || dk <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "%s: empty input", fn);
  if (dk % 8 != 0)
    return fail(BP_ERR_UNSUPPORTED, "%s: sense key width d/nv = %d must be a multiple of 8 (TMA 16-byte stride rule)", fn, dk);
  if (dk > 192) return fail(BP_ERR_UNSUPPORTED, "%s: sense key width %d > 192 is not supported", fn, dk);
  if ((int64_t)batch * seqlen > 0x7fffffff / 2) return fail(BP_ERR_INVALID_ARGUMENT, "%s: batch*seqlen too large", fn);
  return BP_OK;
}

template <int PK, bool kBF16>
static int launch_lse(const CUtensorMap& tm, const LseParams& p, int batch, cudaStream_t st) {
  using C = LseCfg<PK>;
  auto kern = sense_lse_kernel<PK, kBF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_sense_lse_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  // several senses per CTA amortise the start-up cost, as long as the grid still covers the GPU a few times over
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  LseParams pp = p;
  pp.senses_per_cta = kLseSenses;
  while (pp.senses_per_cta > 1 &&
         static_cast<int64_t>(p.num_pairs) * ((p.nv + pp.senses_per_cta - 1) / pp.senses_per_cta) * batch < 4 * sms)
    pp.senses_per_cta >>= 1;
  kern<<<dim3(p.num_pairs, (p.nv + pp.senses_per_cta - 1) / pp.senses_per_cta, batch), kThreads, C::kSmemBytes, st>>>(tm, pp);
  return check_launch("bp_sense_lse_fwd launch");
}

template <int PK, bool kBF16>
static int launch_mix(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmC, const MixParams& p,
                      int batch, cudaStream_t st) {
  using C = MixCfg<PK>;
  auto kern = sense_mix_kernel<PK, kBF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_sense_mix_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  kern<<<dim3(p.num_qtiles * p.num_chunks, batch), kThreads, C::kSmemBytes, st>>>(tmQ, tmK, tmC, p);
  return check_launch("bp_sense_mix_fwd launch");
}

static bool use_pair_kernel() {   // BP_SENSE_PAIR=1 selects the CTA-pair kernel (A/B measurements)
  static const bool on = [] {
    const char* e = getenv("BP_SENSE_PAIR");   // measured: level with the 1-CTA kernel at 384 columns, slower than it at 256
    return e && e[0] == '1';
  }();
  return on;
}

template <bool kBF16>
static int launch_mix_pair(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmC, const MixParams& p,
                           int batch, cudaStream_t st) {
  auto kern = sense_mix_pair_kernel<kBF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "bp_sense_mix_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int num_qpairs = (p.num_qtiles + 1) / 2;
  kern<<<dim3(num_qpairs * p.num_chunks * 2, batch), kThreads, PairCfg::kSmemBytes, st>>>(tmQ, tmK, tmC, p);
  return check_launch("bp_sense_mix_fwd (pair) launch");
}

}  // namespace sense
}  // namespace bp

extern "C" int bp_sense_lse_fwd(const void* qk, float* lse, int32_t batch, int32_t seqlen, int32_t nv, int32_t dk,
                                float softmax_scale, int32_t dtype, void* stream) {
  using namespace bp;
  if (!qk || !lse) return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_lse_fwd: null pointer argument");
  if (int rc = sense::check_common("bp_sense_lse_fwd", batch, seqlen, nv, dk, dtype)) return rc;
  CUtensorMap tm;
  if (int rc = sense::make_qk_map(&tm, qk, batch, seqlen, nv, dk, dtype, 128)) return rc;
  sense::LseParams p;
  p.lse = lse;
  p.seqlen = seqlen, p.nv = nv, p.dk = dk, p.ksteps = (dk + 15) / 16;
  p.num_pairs = (seqlen + 255) / 256;
  p.scale = softmax_scale, p.scale_log2 = softmax_scale * sense::kLog2e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pk = (dk + 63) / 64;
  const bool bf = dtype == BP_DTYPE_BF16;
  switch (pk) {
    case 1: return bf ? sense::launch_lse<1, true>(tm, p, batch, st) : sense::launch_lse<1, false>(tm, p, batch, st);
    case 2: return bf ? sense::launch_lse<2, true>(tm, p, batch, st) : sense::launch_lse<2, false>(tm, p, batch, st);
    default: return bf ? sense::launch_lse<3, true>(tm, p, batch, st) : sense::launch_lse<3, false>(tm, p, batch, st);
  }
}

extern "C" int bp_sense_mix_fwd(const void* qk, const void* content, const float* lse, void* out, int32_t batch,
                                int32_t seqlen, int32_t nv, int32_t dk, int32_t d, int64_t c_batch_stride,
                                int64_t c_sense_stride, int64_t c_row_stride, float softmax_scale, int32_t dtype,
                                void* stream) {
  using namespace bp;
  if (!qk || !content || !lse || !out) return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_mix_fwd: null pointer argument");
  if (int rc = sense::check_common("bp_sense_mix_fwd", batch, seqlen, nv, dk, dtype)) return rc;
  if (d <= 0 || d % 64 != 0)
    return fail(BP_ERR_UNSUPPORTED, "bp_sense_mix_fwd: model width d = %d must be a multiple of 64", d);
  if (c_batch_stride % 8 || c_sense_stride % 8 || c_row_stride % 8 || c_row_stride < d)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_mix_fwd: content strides must be multiples of 8 elements with unit column stride");
  if ((uintptr_t)content % 16 || (uintptr_t)qk % 16 || (uintptr_t)out % 16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_mix_fwd: pointers must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmC;  // Q tiles are 128 rows, K tiles 64 rows: same tensor, two box heights
  if (int rc = sense::make_qk_map(&tmQ, qk, batch, seqlen, nv, dk, dtype, 128)) return rc;
  if (int rc = sense::make_qk_map(&tmK, qk, batch, seqlen, nv, dk, dtype, 64)) return rc;
  // The reference hands content as a transposed view of (b, s, nv, d) (backpack.py:276), i.e. the sense
  // stride is smaller than the row stride; keep the tensor-map dimensions ordered by increasing stride.
  const bool sense_inner = c_sense_stride < c_row_stride;
  if (sense_inner) {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)nv, (uint64_t)seqlen, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)c_sense_stride * 2, (uint64_t)c_row_stride * 2, (uint64_t)c_batch_stride * 2};
    const uint32_t box[4] = {64, 1, 64, 1};
    if (int rc = encode_tensor_map(&tmC, dtype, 4, content, dims, str, box, true)) return rc;
  } else {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)seqlen, (uint64_t)nv, (uint64_t)batch};
    const uint64_t str[3] = {(uint64_t)c_row_stride * 2, (uint64_t)c_sense_stride * 2, (uint64_t)c_batch_stride * 2};
    const uint32_t box[4] = {64, 64, 1, 1};
    if (int rc = encode_tensor_map(&tmC, dtype, 4, content, dims, str, box, true)) return rc;
  }
  sense::MixParams p;
  p.c_sense_inner = sense_inner ? 1 : 0;
  p.trace = g_trace;
  p.lse = lse;
  p.out = out;
  p.seqlen = seqlen, p.nv = nv, p.dk = dk, p.ksteps = (dk + 15) / 16, p.d = d;
  p.num_qtiles = (seqlen + sense::BM - 1) / sense::BM;
  p.num_chunks = (d + BP_SENSE_DC - 1) / BP_SENSE_DC;
  p.scale_log2 = softmax_scale * sense::kLog2e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pk = (dk + 63) / 64;
  p.group = pk == 1 ? sense::kGroup : 1 << 20;
  const bool bf = dtype == BP_DTYPE_BF16;
  if (pk == 1 && d % 384 == 0 && p.num_qtiles >= 2 && sense::use_pair_kernel()) {
    // CTA-pair kernel: every CTA loads 32-key halves of the K tiles
    if (int rc = sense::make_qk_map(&tmK, qk, batch, seqlen, nv, dk, dtype, 32)) return rc;
    p.num_chunks = d / 384;
    return bf ? sense::launch_mix_pair<true>(tmQ, tmK, tmC, p, batch, st) : sense::launch_mix_pair<false>(tmQ, tmK, tmC, p, batch, st);
  }
  switch (pk) {
    case 1: return bf ? sense::launch_mix<1, true>(tmQ, tmK, tmC, p, batch, st) : sense::launch_mix<1, false>(tmQ, tmK, tmC, p, batch, st);
    case 2: return bf ? sense::launch_mix<2, true>(tmQ, tmK, tmC, p, batch, st) : sense::launch_mix<2, false>(tmQ, tmK, tmC, p, batch, st);
    default: return bf ? sense::launch_mix<3, true>(tmQ, tmK, tmC, p, batch, st) : sense::launch_mix<3, false>(tmQ, tmK, tmC, p, batch, st);
  }
}
