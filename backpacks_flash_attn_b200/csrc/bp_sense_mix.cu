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
// the 256 KB of TMEM -- so one CTA owns a 384-column chunk of the output.
//
// The sense vectors C_l(x_j) come from either
//   * a (b, nv, s, d) tensor with arbitrary batch / sense / row strides (the reference's transposed view of the
//     content model's output, backpack.py:276, or an edited tensor of the intervention wrappers), or
//   * a (vocab, nv, d) TABLE of sense vectors plus the token ids: C_l(x) is a pure function of the token
//     (backpack.py:258 -- no positions, identity mixer), so for inference the content model collapses to a row
//     gather, and the gather happens INSIDE the kernel (16-byte cp.async copies straight into the swizzled operand
//     tile); no (b, s, nv, d) tensor ever exists in HBM.
#include <stdlib.h>

#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace sense {

constexpr int BM = 128;
constexpr int kThreads = 384;
constexpr float kLog2e = 1.4426950408889634f;
// Of every 8 exponentials, how many run as a polynomial on the FMA pipe instead of MUFU (0, 2 or 4; rel. error 7.5e-5).
// Pass 1 is MUFU-bound and gains 7 % from 2 of 8 (0.262 -> 0.235 ms at config 3); pass 2 is bound by the tensor pipe
// and its single issuer and measured 2-3 % SLOWER with it, so it keeps every exponential on MUFU.
#ifndef BP_SENSE_LSE_POLY
#define BP_SENSE_LSE_POLY 2
#endif
#ifndef BP_SENSE_MIX_POLY
#define BP_SENSE_MIX_POLY 0
#endif
constexpr int kPolyLse = BP_SENSE_LSE_POLY, kPolyMix = BP_SENSE_MIX_POLY;

// =============================================================================================
// pass 1: row statistics
// =============================================================================================
constexpr int kLseSenses = 4;   // senses per CTA of pass 1 (a CTA per sense spent more time starting up than working)
// Pass 1 runs FOUR softmax warpgroups: two per query tile, each owning one 64-key half of every 128-key block (the
// (max, sum) pairs of the two halves of a row are merged once per sense through shared memory).  With one warpgroup
// per tile (a whole 128-key row per thread) each SM sub-partition hosted two MUFU-bound warps walking a serial chain
// and the MUFU pipe -- the roofline of this pass -- sat at ~45 %.
constexpr int kLseThreads = 640;   // 4 service warps + 16 softmax warps

template <int PK>  // 64-column panels covering dk
struct LseCfg {
  static constexpr int BN = 128;
  static constexpr int kStages = PK == 3 ? 2 : 4;
  static constexpr int QB = PK == 1 ? 2 : 1;     // Q buffers across senses
  static constexpr uint32_t kQTileBytes = BM * 128 * PK;
  static constexpr uint32_t kKTileBytes = BN * 128 * PK;
  static constexpr uint32_t offQ = 0;                                  // [QB][2 tiles]
  static constexpr uint32_t offK = offQ + QB * 2 * kQTileBytes;
  static constexpr uint32_t offMerge = offK + kStages * kKTileBytes;   // [2 sense parities][2 tiles][2 halves][128] (m, l)
  static constexpr uint32_t offBar = offMerge + 2 * 2 * 2 * BM * 8;
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
// next sense's loads and first S overlap the tail of the current one.  The pass is bound by the MUFU pipe (one
// exponential per score, 1/17 of the operator's MMA work): like the attention kernel it keeps a whole 128-key row
// per thread, takes the row max in eight independent chains, skips 32-key chunks above the diagonal and sums in packed fp32.
template <int PK, bool kBF16>
__global__ void __launch_bounds__(kLseThreads, 1)
sense_lse_kernel(const __grid_constant__ CUtensorMap tmQK, const LseParams p) {
  using C = LseCfg<PK>;
  constexpr int BN = C::BN;
  constexpr int NC = BN / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  LseBarriers& bars = *reinterpret_cast<LseBarriers*>(smem + C::offBar);
  // roles 0-3 (producers / issuer) run in the highest physical warps, roles 4-11 (softmax) in warps 0-7: the
  // sub-partition arbiter prefers the highest eligible warp id, and the single-thread roles must not starve
  const int warp = role_warp<20>(), lane = threadIdx.x & 31;
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
      for (int i = 0; i < 2; ++i) mbar_init(&bars.s_full[t][i], 1), mbar_init(&bars.s_free[t][i], 256);
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
    // (no setmaxnreg in this kernel: 640 threads x 96 registers fit the register file as compiled)
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
      // (whole warp, warp-uniform values, elected lane inside the asm: see the issuer of pass 2)
      constexpr uint32_t idesc = make_idesc(kBF16, BM, BN, false, false);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t dQ0 = make_smem_desc_sw128(smem_u32(smem + C::offQ), 16, 1024);
      const uint64_t dK0 = make_smem_desc_sw128(smem_u32(smem + C::offK), 16, 1024);
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
            const uint64_t a0 = dQ0 + static_cast<uint32_t>(qb * 2 + t) * (C::kQTileBytes >> 4);
            const uint64_t b0 = dK0 + static_cast<uint32_t>(slot) * (C::kKTileBytes >> 4);
            for (int kk = 0; kk < p.ksteps; ++kk)
              umma_ss_w(tm + (t * 2 + (c & 1)) * BN, a0 + (kk >> 2) * ((BM * 128) >> 4) + (kk & 3) * 2,
                        b0 + (kk >> 2) * ((BN * 128) >> 4) + (kk & 3) * 2, idesc, kk > 0 ? 1u : 0u);
            if (t == last_user) umma_commit_w(smem_u32(&bars.k_empty[slot]));
            if (j == n_max - 1 && t == last_user) umma_commit_w(smem_u32(&bars.q_empty[qb]));   // last S of this sense
            umma_commit_w(smem_u32(&bars.s_full[t][c & 1]));
          }
        }
      }
    }
  } else {
    constexpr int HN = BN / 2;         // keys per block and thread
    constexpr int NCH = HN / 32;
    const int sr = warp - 4;           // softmax warp 0..15
    const int t = sr >> 3;             // query tile
    const int half = (sr >> 2) & 1;    // which 64-key half of every block
    const int r = (sr & 3) * 32 + lane;
    const int n = n_blk[t];
    if (n > 0) {
      const int row0_t = row0 + t * BM;
      const int qrow = row0_t + r;
      const uint32_t lane_addr = static_cast<uint32_t>((sr & 3) * 32) << 16;
      const float c2 = p.scale_log2;
      float2* merge = reinterpret_cast<float2*>(smem + C::offMerge);
      int c = 0;   // S tiles consumed (all senses)
      for (int si = 0; si < n_senses; ++si) {
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < n; ++j, ++c) {
          mbar_wait(&bars.s_full[t][c & 1], (c >> 1) & 1);
          tc_fence_after();
          float s[HN];
          const uint32_t tS = tmem_base + lane_addr + (t * 2 + (c & 1)) * BN + half * HN;
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            uint32_t u[32];
            tmem_ld32(tS + cc * 32, u);
#pragma unroll
            for (int i = 0; i < 32; ++i) s[cc * 32 + i] = __uint_as_float(u[i]);
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&bars.s_free[t][c & 1]);
          const int col0 = j * BN + half * HN;
          uint32_t dead = 0;   // bit cc: chunk cc is above the diagonal for every row of this warp
          if (col0 + HN - 1 > row0_t) {  // this half block touches the diagonal (also covers cols >= seqlen)
            const int lim = qrow + 1 - col0;   // visible keys of this row inside the half block
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) {
              if (!__all_sync(0xffffffffu, lim >= (cc + 1) * 32)) {
                if (__all_sync(0xffffffffu, lim <= cc * 32)) {
                  dead |= 1u << cc;
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) s[cc * 32 + i] = (cc * 32 + i < lim) ? s[cc * 32 + i] : -INFINITY;
                }
              }
            }
          }
          float mx8[8];   // eight chains of 2-input FMNMX (FMNMX3 measured ~5x slower per instruction)
#pragma unroll
          for (int k = 0; k < 8; ++k) mx8[k] = -INFINITY;
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            if (!((dead >> cc) & 1u)) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) mx8[k] = fmaxf(mx8[k], s[cc * 32 + i + k]);
              }
            }
          }
          const float mxa = fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3]));
          const float mxb = fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]));
          const float m_new = fmaxf(m, fmaxf(mxa, mxb));
          // the second half of a block may hold no visible key for a row: keep (m, l) = (-inf, 0) without a NaN
          const float m_ref = (m_new == -INFINITY) ? 0.f : m_new;
          const float neg = -m_ref * c2;
          float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            if (!((dead >> cc) & 1u)) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                float e[8];
                exp2_scaled8<kPolyLse>(e, &s[cc * 32 + i], c2, neg);
                add2(sum4[0], sum4[1], e[0], e[1]);
                add2(sum4[2], sum4[3], e[2], e[3]);
                add2(sum4[0], sum4[1], e[4], e[5]);
                add2(sum4[2], sum4[3], e[6], e[7]);
              }
            }
          }
          l = l * fast_exp2((m - m_ref) * c2) + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
          m = m_new;
        }
        // merge the two halves of the row (once per sense): the odd half publishes, the even half combines and stores
        float2* slot = merge + ((si & 1) * 4 + t * 2) * BM;
        if (half == 1) slot[BM + r] = make_float2(m, l);
        if (t == 0) named_bar_sync(1, 256);
        else named_bar_sync(2, 256);
        if (half == 0) {
          const float2 o = slot[BM + r];
          const float M = fmaxf(m, o.x);   // the even half always sees key 0: finite
          const float L = l * fast_exp2((m - M) * c2) + (o.x == -INFINITY ? 0.f : o.y * fast_exp2((o.x - M) * c2));
          if (qrow < S) p.lse[(static_cast<int64_t>(batch) * p.nv + sense0 + si) * S + qrow] = M * p.scale + logf(L);
        }
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
//
// Layout per CTA (384 threads): warp 0 = TMA producer for the sense vectors C (tiled loads from a (b,nv,s,d)
// tensor, or row gathers from the (vocab,nv,d) table), warp 3 = TMA producer for Q_l / K_l (+ TMEM allocator),
// warp 1 = the ONE issuer of all tcgen05.mma, warpgroups 1/2 = softmax for even/odd steps + epilogue.
//
// TMEM: the 384-column fp32 accumulator O plus two 64-column S/P buffers B_0, B_1.  Step n (one sense, one 64-key
// block) uses B_{n&1}: S(n) lands there, the softmax threads of warpgroup n&1 pull their row into registers and write
// P(n) (bf16/f16 pairs, 32 columns) back INTO THE SAME BUFFER, and P(n) is the A operand of the TS product
// O += P(n) C(n).  One thread issues, in this order,
//     ... PV(n), S(n+2), PV(n+1), S(n+3), ...
// and tcgen05.mma of one thread execute in issue order, so S(n+2) overwrites B_{n&1} only after PV(n) has read P(n)
// from it -- no barrier is needed for the buffer hand-back, and S runs two steps ahead of PV with only 128 TMEM
// columns (the first version had one S buffer and two P buffers; S(n+1) could not be issued before the softmax had
// drained S(n), and the tensor pipe idled 40 % of the time waiting for P).
// The exponentials of a step are handed over in two halves (keys 0-31, keys 32-63), each with its own barrier: the
// first two K-steps of PV(n) start while the second half of the exponentials is still running, which takes half of
// the MUFU time (the floor of the S -> P latency chain) off the critical path.
// A satisfied mbarrier wait still costs ~200 cycles of latency on the single issuing thread, so the barrier that
// the TMA producer arms for the C tile also counts the 128 softmax threads that hand over the first half of P.
template <int PK>
struct MixCfg {
  static constexpr int BN = 64;    // keys per step
  // Output columns per CTA.  The fp32 accumulator of a 128-row tile at d = 768 (384 KB) exceeds TMEM, so a CTA owns
  // a column chunk and recomputes S (and the exponentials) for it.  256-column chunks (three of them) were measured
  // slower in round 1: every chunk repeats the exponentials and the MUFU pipe becomes the bound.
  static constexpr int DC = 384;
  static constexpr int QS = PK == 1 ? 2 : 1;
  static constexpr int KS = 3;
  static constexpr int CS = PK == 1 ? 3 : 2;
  static constexpr uint32_t kQTileBytes = BM * 128 * PK;
  static constexpr uint32_t kKTileBytes = BN * 128 * PK;
  static constexpr uint32_t kCPanelBytes = BN * 128;           // 64 keys x 64 columns
  static constexpr uint32_t kCTileBytes = (DC / 64) * kCPanelBytes;
  static constexpr uint32_t offQ = 0;
  static constexpr uint32_t offK = offQ + QS * kQTileBytes;
  static constexpr uint32_t offC = offK + KS * kKTileBytes;
  static constexpr uint32_t offBar = offC + CS * kCTileBytes;
  static constexpr uint32_t kSmemBytes = offBar + 256 + 1024;
  static constexpr uint32_t colO = 0, colB = DC;   // B_b at colB + b * BN
  static constexpr uint32_t kTmemCols = 512;
  static_assert(colB + 2 * BN <= 512, "TMEM budget");
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
constexpr int kGroup = 4;
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
  uint64_t k_full[3], k_empty[3];
  uint64_t pa_go[3], c_empty[3];   // pa_go[n % CS]: C(n) landed (tx) + first half of P(n) stored by its 128 softmax threads
  uint64_t pb_go[2];               // pb_go[n & 1]: second half of P(n) stored
  uint64_t s_full[2];              // s_full[n & 1]: S(n) complete in B_{n&1}
  uint64_t o_full;
  uint32_t tmem_base;
};

struct MixParams {
  const float* lse;  // (b, nv, s), natural log
  void* out;         // (b, s, d)
  const int64_t* ids;     // table mode: token ids (b, s); null in tensor mode
  const void* table;      // table mode: (vocab, nv, d) sense vectors
  int32_t vocab;          // table mode: rows of ids are clamped to [0, vocab)
  int32_t seqlen, nv, dk, ksteps, d, num_qtiles, num_chunks;
  int32_t c_sense_inner;  // tensor mode: content tensor-map dims are (d, nv, s, b) instead of (d, s, nv, b)
  int32_t group;          // key blocks per group of the step order (StepIter)
  int32_t out_f32;        // debug / test mode: `out` is fp32, written before the 16-bit rounding
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
  // roles 0-3 (producers / issuer) run in the highest physical warps, roles 4-11 (softmax) in warps 0-7: the
  // sub-partition arbiter prefers the highest eligible warp id, and the single-thread roles must not starve
  const int warp = role_warp<12>(), lane = threadIdx.x & 31;
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
    if (p.ids == nullptr) tma_prefetch_desc(&tmC);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.q_full[i], 1), mbar_init(&bars.q_empty[i], 1);
      mbar_init(&bars.s_full[i], 1), mbar_init(&bars.pb_go[i], 128);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars.k_full[i], 1), mbar_init(&bars.k_empty[i], 1);
      // C(n) landed: one expect_tx arrival of the TMA producer, or one cp.async-completion arrival per gather thread
      mbar_init(&bars.pa_go[i], p.ids != nullptr ? 128 + 64 : 128 + 1), mbar_init(&bars.c_empty[i], 1);
    }
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
    reg_dealloc<104>();   // 128 x 104 + 256 x 200 = 384 x 168: the issuer keeps its iterators in registers
    if (warp == 0 && p.ids == nullptr) {
      // ---- producer A, tensor mode: content tiles C_l[j] : (ncols/64) panels of [64 keys x 64 columns] by TMA ----
      Tracer tr(p.trace, 0, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0);
      StepIter it(nj, p.nv, p.group);
      const uint32_t tile_bytes = (ncols / 64) * C::kCPanelBytes;
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int slot = n % C::CS;
        tr.rec(0, n);
        if (n >= C::CS) mbar_wait(&bars.c_empty[slot], ((n / C::CS) - 1) & 1);
        tr.rec(1, n);
        {
          const int sense = it.sense(), j = it.j();
          const uint32_t bar = smem_u32(&bars.pa_go[slot]);
          const uint32_t dst0 = smem_u32(smem + C::offC) + slot * C::kCTileBytes;
          mbar_arrive_expect_tx_w(bar, tile_bytes);
          const int c1 = p.c_sense_inner ? sense : j * BN, c2 = p.c_sense_inner ? j * BN : sense;
          for (int pn = 0; pn < ncols / 64; ++pn)
            tma_load_4d_w(dst0 + pn * C::kCPanelBytes, &tmC, bar, col_base + pn * 64, c1, c2, batch);
        }
      }
    } else if ((warp == 0 || warp == 2) && p.ids != nullptr) {
      // ---- producer A, table mode: warps 0 and 2 gather row (x_j * nv + l) of the (vocab * nv, d) table for each of
      // the 64 keys of the block with 16-byte cp.async copies, writing the 128B-swizzled panel layout the MMA expects
      // (TMA is the wrong tool here: tile::gather4 walks a descriptor per 128-byte row and measured 2.8x slower
      // than the whole tiled kernel).  Lane l copies chunk (l & 7) of rows 4i + (l >> 3): one warp instruction moves
      // four whole 128-byte lines.  Warp g takes every second 64-column panel.  Completion is signalled to the MMA
      // issuer by the copies themselves (cp.async.mbarrier.arrive.noinc): the gather warps never wait for data.
      const int gw = warp >> 1;
      const int chunk = lane & 7, r4 = lane >> 3;
      Tracer tr(p.trace, 0, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && warp == 0);
      StepIter it(nj, p.nv, p.group);
      const int64_t* ids = p.ids + static_cast<int64_t>(batch) * S;
      const uint32_t row_bytes = static_cast<uint32_t>(p.d) * 2u;
      // this lane's 16-byte chunk of the first panel of this warp, of table row 0
      const uint8_t* tb = static_cast<const uint8_t*>(p.table) + (static_cast<int64_t>(col_base) + gw * 64 + chunk * 8) * 2;
      const int npw = (ncols / 64 - gw + 1) / 2;            // panels of this warp: gw, gw + 2, ...
      // destination of row 4i + r4: (row & 7) is r4 for even i and r4 + 4 for odd i, so the swizzle has two values
      const uint32_t dst_lane = smem_u32(smem + C::offC) + gw * C::kCPanelBytes + r4 * 128;
      const uint32_t swz0 = static_cast<uint32_t>(chunk ^ r4) << 4, swz1 = static_cast<uint32_t>(chunk ^ (r4 + 4)) << 4;
      auto load_ids = [&](int j, int& lo, int& hi) {
        // keys beyond the sequence are masked (P = 0): any valid row will do; ids are clamped to the table
        const int k0 = j * BN + lane;
        lo = min(max(static_cast<int>(__ldg(ids + min(k0, S - 1))), 0), p.vocab - 1) * p.nv;
        hi = min(max(static_cast<int>(__ldg(ids + min(k0 + 32, S - 1))), 0), p.vocab - 1) * p.nv;
      };
      int lo, hi;
      load_ids(it.j(), lo, hi);
      int slot = 0;
      uint32_t ph = 0;   // parity of c_empty to wait for once the ring has wrapped: ((n / CS) - 1) & 1
      for (int n = 0; n < n_steps; ++n) {
        const int sense = it.sense();
        it.next();
        int nlo = lo, nhi = hi;
        if (n + 1 < n_steps) load_ids(it.j(), nlo, nhi);   // next step's ids travel while this step's copies are issued
        tr.rec(0, n);
        if (n >= C::CS) mbar_wait(&bars.c_empty[slot], ph);
        tr.rec(1, n);
        const uint32_t dst_slot = dst_lane + slot * C::kCTileBytes;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int rid = __shfl_sync(0xffffffffu, i < 8 ? lo : hi, ((4 * i) & 31) + r4) + sense;
          const uint8_t* src = tb + static_cast<uint64_t>(static_cast<uint32_t>(rid)) * row_bytes;   // one IMAD.WIDE.U32
          const uint32_t dst = dst_slot + i * 512 + ((i & 1) ? swz1 : swz0);
          if (npw == 3) {   // the 384-column chunk of d = 768: three panels per warp, immediates only
            cp_async16(dst, src);
            cp_async16(dst + 2 * C::kCPanelBytes, src + 256);
            cp_async16(dst + 4 * C::kCPanelBytes, src + 512);
          } else {
            for (int k = 0; k < npw; ++k) cp_async16(dst + 2 * k * C::kCPanelBytes, src + 256 * k);
          }
        }
        cp_async_arrive_noinc(smem_u32(&bars.pa_go[slot]));
        lo = nlo, hi = nhi;
        if (++slot == C::CS) {
          slot = 0;
          if (n >= C::CS) ph ^= 1;
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");   // nothing of this CTA may still be in flight at exit
    } else if (warp == 3) {
      // ---- producer B: Q_l (once per (group, sense) visit) and K_l[j] ----
      StepIter it(nj, p.nv, p.group);
      int qv = 0;   // (group, sense) visits so far: Q buffer = qv % QS
      for (int n = 0; n < n_steps; ++n, it.next()) {
        const int sense = it.sense(), j = it.j();
        if (it.first_of_visit()) {
          const int qs = qv % C::QS;
          if (qv >= C::QS) mbar_wait(&bars.q_empty[qs], ((qv / C::QS) - 1) & 1);
          ++qv;
          mbar_arrive_expect_tx_w(smem_u32(&bars.q_full[qs]), C::kQTileBytes);
#pragma unroll
          for (int pn = 0; pn < PK; ++pn)
            tma_load_3d_w(smem_u32(smem + C::offQ) + qs * C::kQTileBytes + pn * (BM * 128), &tmQ, smem_u32(&bars.q_full[qs]),
                          pn * 64, sense, tok0 + row0);
        }
        const int slot = n % C::KS;
        if (n >= C::KS) mbar_wait(&bars.k_empty[slot], ((n / C::KS) - 1) & 1);
        mbar_arrive_expect_tx_w(smem_u32(&bars.k_full[slot]), C::kKTileBytes);
#pragma unroll
        for (int pn = 0; pn < PK; ++pn)
          tma_load_3d_w(smem_u32(smem + C::offK) + slot * C::kKTileBytes + pn * (BN * 128), &tmK, smem_u32(&bars.k_full[slot]),
                        pn * 64, p.nv + sense, tok0 + j * BN);
      }
    } else if (warp == 1) {
      // ---- the issuer: PV(n), S(n+2), PV(n+1), S(n+3), ... ----
      // This thread is the kernel's critical resource.  A timeline trace plus an ncu source profile of the first
      // version of this loop showed ~1800 cycles per step against 896 cycles of tensor-pipe work, all of it the
      // issuing warp's own instruction stream: ~400 dependent scalar instructions per step (shared-memory
      // descriptors re-encoded for every MMA, % and / for the ring slots, an ELECT / vote / divergence check around
      // every tcgen05 instruction) plus four satisfied barrier waits at ~180 cycles each.  Hence:
      //   * the whole warp walks the loop with provably warp-uniform values (role index and TMEM base come out of
      //     shuffles), every tcgen05 instruction is predicated on an elected lane inside its asm block (umma_*_w);
      //   * descriptors are base descriptors plus immediates (desc_add), ring slots and phases are running counters;
      //   * barriers that complete early (K tile landed, next step's P) are PROBED ahead of time with test_wait,
      //     whose latency overlaps the MMA issue; only a failed probe falls back to the blocking wait.
      constexpr uint32_t idesc_s = make_idesc(kBF16, BM, BN, false, false);
      const uint32_t idesc_pv1 = make_idesc(kBF16, BM, n1, false, true);
      const uint32_t idesc_pv2 = make_idesc(kBF16, BM, n2 > 0 ? n2 : 64, false, true);
      const uint64_t dQ0 = make_smem_desc_sw128(smem_u32(smem + C::offQ), 16, 1024);
      const uint64_t dK0 = make_smem_desc_sw128(smem_u32(smem + C::offK), 16, 1024);
      const uint64_t dC0 = make_smem_desc_sw128(smem_u32(smem + C::offC), C::kCPanelBytes, 1024);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
      const uint32_t tO = tm + C::colO, tB = tm + C::colB;
      const uint32_t bars_a = smem_u32(&bars);
#define MIX_BAR(field, i) (bars_a + static_cast<uint32_t>(offsetof(MixBarriers, field)) + 8u * static_cast<uint32_t>(i))
      const bool table = p.ids != nullptr;
      const bool two = n2 > 0;
      const int ksteps = p.ksteps;
      Tracer tr(p.trace, 1, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0);
      StepIter its(nj, p.nv, p.group);   // walks the S products (two steps ahead of the PV products)
      // S side: step m, K ring slot / phase, Q buffer / phase
      int m = 0, s_ks = 0, s_qs = 0;
      uint32_t s_kph = 0, s_qph = 0;
      auto issue_s = [&](bool k_ready) {
        if (its.first_of_visit()) mbar_wait_a(MIX_BAR(q_full, s_qs), s_qph);
        if (!k_ready) mbar_wait_a(MIX_BAR(k_full, s_ks), s_kph);
        tc_fence_after();
        const uint64_t a0 = dQ0 + static_cast<uint32_t>(s_qs) * (C::kQTileBytes >> 4);
        const uint64_t b0 = dK0 + static_cast<uint32_t>(s_ks) * (C::kKTileBytes >> 4);
        const uint32_t d = tB + (m & 1) * BN;
        if constexpr (PK == 1) {
#pragma unroll 4
          for (int kk = 0; kk < ksteps; ++kk) umma_ss_w(d, a0 + 2u * kk, b0 + 2u * kk, idesc_s, kk > 0 ? 1u : 0u);
        } else {
          for (int kk = 0; kk < ksteps; ++kk)
            umma_ss_w(d, a0 + (kk >> 2) * ((BM * 128) >> 4) + (kk & 3) * 2, b0 + (kk >> 2) * ((BN * 128) >> 4) + (kk & 3) * 2,
                      idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit_w(MIX_BAR(k_empty, s_ks));
        if (its.last_of_visit()) {
          umma_commit_w(MIX_BAR(q_empty, s_qs));
          if (++s_qs == C::QS) s_qs = 0, s_qph ^= 1;
        }
        umma_commit_w(MIX_BAR(s_full, m & 1));
        if (++s_ks == C::KS) s_ks = 0, s_kph ^= 1;
        its.next();
        ++m;
      };
      issue_s(false);
      if (n_steps > 1) issue_s(false);
      bool pa_ready = false;   // probe of pa_go for the step at the top of the loop
      int cs = 0;              // C ring slot of step n
      uint32_t cph = 0;        // its phase
      for (int n = 0; n < n_steps; ++n) {
        const uint32_t a_tmem = tB + (n & 1) * BN;   // P(n): 8 columns per K-step of 16 keys
        const uint64_t b0 = dC0 + static_cast<uint32_t>(cs) * (C::kCTileBytes >> 4);
        const uint32_t acc0 = n > 0 ? 1u : 0u;
        tr.rec(3, n);
        if (!pa_ready) mbar_wait_a(MIX_BAR(pa_go, cs), cph);      // C(n) landed, keys 0-31 of P(n) stored
        if (table) fence_proxy_async_smem();   // the gathered C tile was written by cp.async (generic proxy)
        tc_fence_after();
        tr.rec(4, n);
        // K(n+2) landed long ago (the K ring runs three steps ahead): probe now, use after the PV products
        const bool k_ready = m < n_steps ? mbar_test_a(MIX_BAR(k_full, s_ks), s_kph) : true;
        umma_ts_w(tO, a_tmem, b0, idesc_pv1, acc0);
        if (two) umma_ts_w(tO + 256, a_tmem, desc_add(b0, 4 * C::kCPanelBytes), idesc_pv2, acc0);
        umma_ts_w(tO, a_tmem + 8, desc_add(b0, 2048), idesc_pv1, 1u);
        if (two) umma_ts_w(tO + 256, a_tmem + 8, desc_add(b0, 4 * C::kCPanelBytes + 2048), idesc_pv2, 1u);
        mbar_wait_a(MIX_BAR(pb_go, n & 1), (n >> 1) & 1);      // keys 32-63 of P(n) stored (the pipe has 2 K-steps queued)
        tc_fence_after();
        tr.rec(5, n);
        umma_ts_w(tO, a_tmem + 16, desc_add(b0, 4096), idesc_pv1, 1u);
        if (two) umma_ts_w(tO + 256, a_tmem + 16, desc_add(b0, 4 * C::kCPanelBytes + 4096), idesc_pv2, 1u);
        umma_ts_w(tO, a_tmem + 24, desc_add(b0, 6144), idesc_pv1, 1u);
        if (two) umma_ts_w(tO + 256, a_tmem + 24, desc_add(b0, 4 * C::kCPanelBytes + 6144), idesc_pv2, 1u);
        umma_commit_w(MIX_BAR(c_empty, cs));
        if (n == n_steps - 1) umma_commit_w(MIX_BAR(o_full, 0));
        if (++cs == C::CS) cs = 0, cph ^= 1;
        // P(n+1): its first half is usually stored by now (its warpgroup started when S(n+1) completed, a step ago)
        pa_ready = n + 1 < n_steps ? mbar_test_a(MIX_BAR(pa_go, cs), cph) : false;
        if (m < n_steps) issue_s(k_ready);   // S(n+2) into B_{n&1}: ordered behind PV(n) by the in-order tensor pipe
        tr.rec(6, n);
      }
#undef MIX_BAR
    }
  } else {
    // ---- softmax warpgroups: w handles steps n = 2i + w, buffer B_w ----
    reg_alloc<200>();
    const int w = (warp >> 2) - 1;
    const int r = (warp & 3) * 32 + lane;
    const int qrow = row0 + r;
    const bool valid = qrow < S;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tB = tmem_base + lane_addr + C::colB + w * BN;
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
        tmem_ld32(tB + c * 32, u);
#pragma unroll
        for (int k = 0; k < 32; ++k) s[c * 32 + k] = __uint_as_float(u[k]);
      }
      tmem_ld_wait();   // the whole row is in registers: P(n) may now overwrite the buffer
      tr.rec(2, n);
      const int col0 = j * BN;
      if (col0 + BN - 1 > row0) {
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (col0 + c > qrow) s[c] = -INFINITY;
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          float e[8];
          exp2_scaled8<kPolyMix>(e, &s[hf * 32 + c], c2, neg_lse2);
          pk[c / 2 + 0] = pack2<kBF16>(e[0], e[1]);
          pk[c / 2 + 1] = pack2<kBF16>(e[2], e[3]);
          pk[c / 2 + 2] = pack2<kBF16>(e[4], e[5]);
          pk[c / 2 + 3] = pack2<kBF16>(e[6], e[7]);
        }
        tmem_st16(tB + hf * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        if (hf == 0) mbar_arrive(&bars.pa_go[n % C::CS]); else mbar_arrive(&bars.pb_go[w]);
        tr.rec(3 + hf, n);
      }
    }
    // ---- epilogue: warpgroup w stores columns [w*ncols/2, (w+1)*ncols/2) of its row ----
    mbar_wait(&bars.o_full, 0);
    tc_fence_after();
    const int half_cols = ncols / 2;  // multiple of 32
    const uint32_t tO = tmem_base + lane_addr + C::colO + w * half_cols;
    const int64_t o_elem = (static_cast<int64_t>(tok0) + qrow) * p.d + col_base + w * half_cols;
    for (int c = 0; c < half_cols / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (valid) {
        if (p.out_f32) {
          float* orow = reinterpret_cast<float*>(p.out) + o_elem + c * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(orow + g * 4) = make_uint4(o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
        } else {
          uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) + 2 * (o_elem + c * 32);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 v;
            v.x = pack2<kBF16>(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
            v.y = pack2<kBF16>(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
            v.z = pack2<kBF16>(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
            v.w = pack2<kBF16>(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + g * 16) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, C::kTmemCols);
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
  if (batch > 65535)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: batch %d exceeds 65535 (the batch index is a grid y/z dimension)", fn, batch);
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
  kern<<<dim3(p.num_pairs, (p.nv + pp.senses_per_cta - 1) / pp.senses_per_cta, batch), kLseThreads, C::kSmemBytes, st>>>(tm, pp);
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

static int g_out_f32 = 0;   // bp_debug_set_sense_out_f32

// shared tail of the two pass-2 entry points: Q / K maps, parameters, dispatch
static int run_mix(const char* fn, const void* qk, const CUtensorMap& tmC, MixParams p, const float* lse, void* out,
                   int batch, int seqlen, int nv, int dk, int d, float softmax_scale, int dtype, void* stream) {
  CUtensorMap tmQ, tmK;  // Q tiles are 128 rows, K tiles 64 rows: same tensor, two box heights
  if (int rc = make_qk_map(&tmQ, qk, batch, seqlen, nv, dk, dtype, 128)) return rc;
  if (int rc = make_qk_map(&tmK, qk, batch, seqlen, nv, dk, dtype, 64)) return rc;
  p.trace = g_trace;
  p.lse = lse;
  p.out = out;
  p.out_f32 = g_out_f32;
  p.seqlen = seqlen, p.nv = nv, p.dk = dk, p.ksteps = (dk + 15) / 16, p.d = d;
  p.num_qtiles = (seqlen + BM - 1) / BM;
  p.num_chunks = (d + MixCfg<1>::DC - 1) / MixCfg<1>::DC;
  p.scale_log2 = softmax_scale * kLog2e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pk = (dk + 63) / 64;
  p.group = pk == 1 ? kGroup : 1 << 20;
  const bool bf = dtype == BP_DTYPE_BF16;
  (void)fn;
  switch (pk) {
    case 1: return bf ? launch_mix<1, true>(tmQ, tmK, tmC, p, batch, st) : launch_mix<1, false>(tmQ, tmK, tmC, p, batch, st);
    case 2: return bf ? launch_mix<2, true>(tmQ, tmK, tmC, p, batch, st) : launch_mix<2, false>(tmQ, tmK, tmC, p, batch, st);
    default: return bf ? launch_mix<3, true>(tmQ, tmK, tmC, p, batch, st) : launch_mix<3, false>(tmQ, tmK, tmC, p, batch, st);
  }
}

}  // namespace sense
}  // namespace bp

// debug hook (not part of the public ABI): the next bp_sense_mix*_fwd calls treat `out` as fp32 (b, s, d) and store
// the accumulator before the final 16-bit rounding (test mode T2 of SURVEY.md §8c)
extern "C" void bp_debug_set_sense_out_f32(int on) { bp::sense::g_out_f32 = on ? 1 : 0; }

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
  CUtensorMap tmC;
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
  p.ids = nullptr;
  p.table = nullptr;
  p.vocab = 0;
  p.c_sense_inner = sense_inner ? 1 : 0;
  return sense::run_mix("bp_sense_mix_fwd", qk, tmC, p, lse, out, batch, seqlen, nv, dk, d, softmax_scale, dtype, stream);
}

extern "C" int bp_sense_mix_table_fwd(const void* qk, const void* table, const int64_t* input_ids, const float* lse,
                                      void* out, int32_t batch, int32_t seqlen, int32_t nv, int32_t dk, int32_t d,
                                      int32_t vocab, float softmax_scale, int32_t dtype, void* stream) {
  using namespace bp;
  if (!qk || !table || !input_ids || !lse || !out)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_mix_table_fwd: null pointer argument");
  if (int rc = sense::check_common("bp_sense_mix_table_fwd", batch, seqlen, nv, dk, dtype)) return rc;
  if (d <= 0 || d % 64 != 0)
    return fail(BP_ERR_UNSUPPORTED, "bp_sense_mix_table_fwd: model width d = %d must be a multiple of 64", d);
  if (vocab <= 0 || (int64_t)vocab * nv > 0x7fffffff)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_mix_table_fwd: vocab * nv must fit 31 bits (vocab=%d nv=%d)", vocab, nv);
  if ((uintptr_t)table % 16 || (uintptr_t)qk % 16 || (uintptr_t)out % 16 || (uintptr_t)input_ids % 8)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_mix_table_fwd: pointers must be 16-byte aligned (ids: 8)");
  // table mode reads the (vocab, nv, d) table with plain cp.async copies; the tensor map of the C operand is unused
  // (a valid map of the table keeps the kernel signature uniform)
  CUtensorMap tmC;
  const uint64_t dims[2] = {(uint64_t)d, (uint64_t)vocab * nv};
  const uint64_t str[1] = {(uint64_t)d * 2};
  const uint32_t box[2] = {64, 1};
  if (int rc = encode_tensor_map(&tmC, dtype, 2, table, dims, str, box, true)) return rc;
  sense::MixParams p;
  p.ids = input_ids;
  p.table = table;
  p.vocab = vocab;
  p.c_sense_inner = 0;
  return sense::run_mix("bp_sense_mix_table_fwd", qk, tmC, p, lse, out, batch, seqlen, nv, dk, d, softmax_scale, dtype,
                        stream);
}
