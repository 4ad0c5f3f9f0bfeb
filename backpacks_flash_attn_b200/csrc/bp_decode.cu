// Incremental decoding (one new token per sequence) for sm_100a: the two operators that change shape when the query
// is a single position.  Both are HBM-bound row streams (every step re-reads the whole KV cache / all sense vectors of
// the context), so they are plain CUDA-core kernels built for coalesced 16-byte loads and deterministic reductions --
// there is no GEMM-shaped work to put on the tensor cores.
//
//   bp_decode_attn_fwd      softmax(scale * q K^T) V for ONE query per (batch, head) against a KV cache
//                           (flash_attn/modules/mha.py:356-380, 432-440: _update_kv_cache + inner_cross_attn with
//                           causal=False; the mask is top-left aligned, csrc/flash_attn/src/fmha/mask.h:70, so a
//                           decode step attends to every cached key).
//   bp_sense_mix_decode_fwd the Backpack sense-mix for the LAST position only:
//                           out[b,:] = sum_l sum_{j<len} softmax_j(scale q_l . k_lj) * table[ids[b,j], l, :]
//                           (training/src/models/backpack.py:116-122 + :313 restricted to row i = len - 1).  The
//                           reference has no incremental path for Backpacks: its generation loop re-runs the full forward
//                           for every token (training/src/utils/generation.py:34-44, 62-72).
#include <cooperative_groups.h>

#include <algorithm>

#include "bp_common.cuh"
#include "bp_host.h"

namespace cg = cooperative_groups;

namespace bp {
namespace decode {

template <bool kBF16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kBF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// attention of one query per (batch, head) against the KV cache
// ---------------------------------------------------------------------------------------------
// A thread-block CLUSTER of `nsplit` CTAs (1..8, chosen by the host so that small batches still fill the 148 SMs) per
// (batch, head); CTA r of the cluster takes the r-th contiguous slice of the keys.  Inside a CTA (128 threads) a key row
// (DH 16-bit values) is read by 8 lanes with one or two 16-byte loads each, so a warp streams 4 keys per instruction and
// a CTA 16; the loads of four such steps (64 keys, 16 KB for DH = 64) are issued before the first is consumed.  Every
// 8-lane group runs its own online softmax over the keys it sees (running max, sum and DH/8 accumulators per lane); the
// 16 partial results of a CTA are merged in a fixed order (shuffles inside the warp, then shared memory across warps),
// and the CTAs of the cluster are merged by rank 0 reading its peers' shared memory (DSMEM) in rank order: bitwise
// deterministic for a given nsplit, no workspace in global memory.
constexpr int kAttnUnroll = 4;

struct AttnParams {
  const void* q;        // (batch, nheads, DH)
  const void* kv;       // cache: element (b, j, which, h, :) at b*batch_stride + j*row_stride + which*which_stride + h*DH
  void* out;            // (batch, nheads, DH)
  const int32_t* lens;  // (batch) keys to attend to per sequence, or null: `len` for all
  int64_t batch_stride, row_stride, which_stride;
  int32_t batch, nheads, len, nsplit;
  float scale_log2;
};

template <int DH, bool kBF16>
__global__ void __launch_bounds__(128) decode_attn_kernel(const AttnParams p) {
  constexpr int DPL = DH / 8;          // dims per lane (8 or 16)
  constexpr int NV = DPL / 8;          // 16-byte loads per lane and row
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int bh = blockIdx.x / p.nsplit;
  const int b = bh / p.nheads, h = bh - b * p.nheads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp * 4 + (lane >> 3), slot = lane & 7;
  const int len = p.lens != nullptr ? p.lens[b] : p.len;
  const int per = ((len + p.nsplit - 1) / p.nsplit + 15) & ~15;          // keys per CTA, a multiple of 16
  const int j_begin = min(len, rank * per), j_end = min(len, j_begin + per);
  float q[DPL];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.q) + (static_cast<int64_t>(b) * p.nheads + h) * DH + slot * DPL);
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      float t[8];
      unpack8<kBF16>(__ldg(qp + c), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[c * 8 + i] = t[i] * p.scale_log2;     // scores in log2 units
    }
  }
  const uint16_t* kbase = static_cast<const uint16_t*>(p.kv) + b * p.batch_stride + static_cast<int64_t>(h) * DH + slot * DPL;
  const uint16_t* vbase = kbase + p.which_stride;
  float m = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  // uniform trip count: the shuffles below need the whole warp
  for (int j0 = j_begin; j0 < j_end; j0 += 16 * kAttnUnroll) {
    uint4 kr[kAttnUnroll][NV], vr[kAttnUnroll][NV];
#pragma unroll
    for (int u = 0; u < kAttnUnroll; ++u) {
      const int j = j0 + u * 16 + grp;
      const int64_t row = (j < j_end ? j : j_begin) * p.row_stride;       // dead lanes re-read a live row
      const uint4* kp = reinterpret_cast<const uint4*>(kbase + row);
      const uint4* vp = reinterpret_cast<const uint4*>(vbase + row);
#pragma unroll
      for (int c = 0; c < NV; ++c) kr[u][c] = __ldg(kp + c), vr[u][c] = __ldg(vp + c);
    }
#pragma unroll
    for (int u = 0; u < kAttnUnroll; ++u) {
      const bool live = j0 + u * 16 + grp < j_end;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        float t[8];
        unpack8<kBF16>(kr[u][c], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[c * 8 + i], t[i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (live) {
        const float m_new = fmaxf(m, s);
        const float alpha = fast_exp2(m - m_new), pj = fast_exp2(s - m_new);
        l = l * alpha + pj;
        m = m_new;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          float t[8];
          unpack8<kBF16>(vr[u][c], t);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[c * 8 + i] = fmaf(pj, t[i], acc[c * 8 + i] * alpha);
        }
      }
    }
  }
  // merge the four key groups of the warp (lanes with equal slot), then the four warps through shared memory
  auto merge = [&](float m2, float l2, const float* a2) {
    const float m_new = fmaxf(m, m2);
    const float w1 = m == -INFINITY ? 0.f : fast_exp2(m - m_new), w2 = m2 == -INFINITY ? 0.f : fast_exp2(m2 - m_new);
    l = l * w1 + l2 * w2;
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = acc[i] * w1 + a2[i] * w2;
    m = m_new;
  };
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off), l2 = __shfl_xor_sync(0xffffffffu, l, off);
    float a2[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) a2[i] = __shfl_xor_sync(0xffffffffu, acc[i], off);
    merge(m2, l2, a2);
  }
  __shared__ float sm[4][8][DPL + 2];      // [warp][slot][m, l, acc...]; row [0] doubles as the CTA's merged partial
  if (lane < 8 && warp != 0) {
    sm[warp][lane][0] = m, sm[warp][lane][1] = l;
#pragma unroll
    for (int i = 0; i < DPL; ++i) sm[warp][lane][2 + i] = acc[i];
  }
  __syncthreads();
  if (warp == 0 && lane < 8) {
#pragma unroll
    for (int w = 1; w < 4; ++w) merge(sm[w][lane][0], sm[w][lane][1], &sm[w][lane][2]);
    if (rank != 0) {
      sm[0][lane][0] = m, sm[0][lane][1] = l;
#pragma unroll
      for (int i = 0; i < DPL; ++i) sm[0][lane][2 + i] = acc[i];
    }
  }
  if (p.nsplit > 1) cluster.sync();        // the partials of ranks 1.. are visible cluster-wide
  if (rank == 0 && warp == 0 && lane < 8) {
    for (int r = 1; r < p.nsplit; ++r) {
      const float* peer = cluster.map_shared_rank(&sm[0][lane][0], r);
      float t[DPL + 2];
#pragma unroll
      for (int i = 0; i < DPL + 2; ++i) t[i] = peer[i];
      merge(t[0], t[1], t + 2);
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    uint16_t* op = static_cast<uint16_t*>(p.out) + (static_cast<int64_t>(b) * p.nheads + h) * DH + lane * DPL;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      uint4 o;
      o.x = pack2<kBF16>(acc[c * 8 + 0] * inv, acc[c * 8 + 1] * inv);
      o.y = pack2<kBF16>(acc[c * 8 + 2] * inv, acc[c * 8 + 3] * inv);
      o.z = pack2<kBF16>(acc[c * 8 + 4] * inv, acc[c * 8 + 5] * inv);
      o.w = pack2<kBF16>(acc[c * 8 + 6] * inv, acc[c * 8 + 7] * inv);
      reinterpret_cast<uint4*>(op)[c] = o;
    }
  }
  if (p.nsplit > 1) cluster.sync();        // peers keep their shared memory alive until rank 0 has read it
}

// ---------------------------------------------------------------------------------------------
// sense-mix of the last position
// ---------------------------------------------------------------------------------------------
// A cluster of `nsplit` CTAs (1..8) per (batch element, column chunk of the output); CTA r takes the r-th slice of the
// context.  A chunk is 8 * kLanes columns: kLanes = 32 (256 columns, a whole warp per table row) when the batch alone
// fills the GPU, kLanes = 8 (64 columns) for small batches.  The kernel is HBM-bound only if the inner loop stays under
// ~25 instructions per 512-byte row segment, so everything that is per key or per row is computed once, up front:
//   scores   The K rows of a sub-tile of keys, (key, sense)-major and contiguous in the cache, are read as one flat,
//            fully coalesced stream of 16-byte words; word g multiplies the q words (g mod nv*dk/8) staged in shared
//            memory as fp32 (pre-multiplied by scale * log2(e)) and leaves a partial dot product in shared memory; one
//            thread per row then sums the dk/8 partials into sc[sense][key].
//   pass A   per tile: scores, then the running max / sum of every sense (one warp per sense).  The CTAs of the
//            cluster exchange their (max, sum) through DSMEM and merge them in rank order: exact softmax statistics.
//   pass B   per tile: (scores again, unless the slice is a single tile and they are still in shared memory,)
//            normalise in place, per-key table offsets into shared memory, then acc += w[l][j] * table[ids[j], l, cols]:
//            kLanes lanes read one row segment with 16-byte loads, 256 / kLanes rows per step, 8 steps of loads issued
//            before the first is consumed (32 KB in flight per CTA).
//   merge    the row slots through shared memory in a fixed order, the cluster by rank 0 in rank order: deterministic
//            for a given launch configuration.
// The scores are recomputed by each of the column-chunk clusters of a batch element (for kLanes = 32 and d = 768: 3x,
// 96 bytes of K per 512 bytes of sense vector; the K rows come from L2).
constexpr int kDecThreads = 256;
constexpr int kMixUnroll = 8;
constexpr int kSubKeys = 64;     // keys per score sub-tile (partial dot products: kSubKeys * nv * dk/8 floats)

struct MixParams {
  const void* q;          // (batch, nv, dk) query of the new position
  const void* kcache;     // (batch, max_len, nv, dk)
  const int64_t* ids;     // (batch, max_len) token ids of the context incl. the new position
  const void* table;      // (vocab, nv, d)
  void* out;              // (batch, d)
  const int32_t* lens;    // (batch) or null
  int64_t k_batch_stride, ids_batch_stride;
  int32_t batch, nv, dk, d, vocab, len, nsplit, tile;   // tile: keys per score tile in shared memory (multiple of 16)
  int32_t nv_shift;       // log2(nv) when nv is a power of two (kPow2 kernels)
  float scale_log2;
};

__host__ __device__ inline size_t mix_smem_floats(int nv, int dk, int tile, int lanes) {
  const size_t partial = static_cast<size_t>(kSubKeys) * nv * (dk / 8);
  const size_t merge = static_cast<size_t>(kDecThreads / lanes) * (8 * lanes) + 8 * lanes;
  return static_cast<size_t>(nv) * dk + static_cast<size_t>(nv) * tile + tile + 4 * nv + (partial > merge ? partial : merge);
}

template <bool kBF16, int kLanes, bool kPow2>
__global__ void __launch_bounds__(kDecThreads) sense_mix_decode_kernel(const MixParams p) {
  constexpr int kCols = 8 * kLanes;                  // output columns per CTA
  constexpr int kRows = kDecThreads / kLanes;        // table rows in flight per step
  extern __shared__ float smem_f[];
  const int nv = p.nv, dk = p.dk, tile = p.tile;
  const int C = dk / 8;                              // 16-byte words per K row
  float* qs = smem_f;                                // q * scale * log2(e): word i of q (8 values) as two float4, [2][nv * dk/8][4]
  float* sc = qs + nv * dk;                          // [nv][tile] scores, then normalised weights, of the current tile
  uint32_t* keyoff = reinterpret_cast<uint32_t*>(sc + nv * tile);   // [tile] table offset of the key's token, 16-byte units
  float* stat = reinterpret_cast<float*>(keyoff + tile);            // [nv][2] row max (log2 units) and 1 / row sum
  float* lstat = stat + 2 * nv;                      // [nv][2] this CTA's running max and sum (read by the peers)
  float* scratch = lstat + 2 * nv;                   // partial dot products of a sub-tile | column partials + column sums
  float* part = scratch;                             // [kRows][kCols]
  float* colsum = part + kRows * kCols;              // [kCols] this CTA's column sums (read by rank 0)
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int b = blockIdx.y, chunk = blockIdx.x / p.nsplit;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = p.lens != nullptr ? p.lens[b] : p.len;
  const int per = ((len + p.nsplit - 1) / p.nsplit + 15) & ~15;
  const int j_begin = min(len, rank * per), j_end = min(len, j_begin + per);
  const uint16_t* kb = static_cast<const uint16_t*>(p.kcache) + b * p.k_batch_stride;
  auto split_row = [&](int row, int& jj, int& l) {   // row = jj * nv + l
    if constexpr (kPow2) {
      jj = row >> p.nv_shift, l = row & (nv - 1);
    } else {
      jj = row / nv, l = row - jj * nv;
    }
  };

  {
    const uint16_t* qb = static_cast<const uint16_t*>(p.q) + static_cast<int64_t>(b) * nv * dk;
    for (int i = tid; i < nv * C; i += kDecThreads) {
      float t[8];
      unpack8<kBF16>(__ldg(reinterpret_cast<const uint4*>(qb) + i), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) qs[(e >> 2) * nv * C * 4 + i * 4 + (e & 3)] = t[e] * p.scale_log2;   // conflict-free LDS.128
    }
    for (int i = tid; i < 2 * nv; i += kDecThreads) lstat[i] = (i & 1) ? 0.f : -INFINITY;
  }
  __syncthreads();

  // scores of the keys [j0, j0 + nkeys) into sc[l * tile + jj]; ends with the CTA in sync
  const int qwords = nv * C;                         // 16-byte words of q = words of one key's K rows
  auto compute_scores = [&](int j0, int nkeys) {
    for (int s0 = 0; s0 < nkeys; s0 += kSubKeys) {
      const int sk = min(kSubKeys, nkeys - s0);
      const int nwords = sk * qwords;
      const uint4* kt = reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(j0 + s0) * nv * dk);
      int qi = tid % qwords;
      const int qstep = kDecThreads % qwords;
      for (int g = tid; g < nwords; g += kDecThreads) {
        float k8[8];
        unpack8<kBF16>(__ldg(kt + g), k8);
        const float4 qa = *reinterpret_cast<const float4*>(qs + qi * 4);
        const float4 qc = *reinterpret_cast<const float4*>(qs + qwords * 4 + qi * 4);
        scratch[g] = qa.x * k8[0] + qa.y * k8[1] + qa.z * k8[2] + qa.w * k8[3] + qc.x * k8[4] + qc.y * k8[5] + qc.z * k8[6] + qc.w * k8[7];
        qi += qstep;
        if (qi >= qwords) qi -= qwords;
      }
      __syncthreads();
      for (int row = tid; row < sk * nv; row += kDecThreads) {
        float tot = 0.f;
        for (int c = 0; c < C; ++c) tot += scratch[row * C + c];
        int jj, l;
        split_row(row, jj, l);
        sc[l * tile + s0 + jj] = tot;
      }
      __syncthreads();
    }
  };

  // ---- pass A: running max / sum of every sense over this CTA's keys (warp w owns senses w, w + 8, ...) ----
  const bool single_tile = j_end - j_begin <= tile;
  for (int j0 = j_begin; j0 < j_end; j0 += tile) {
    const int nkeys = min(tile, j_end - j0);
    compute_scores(j0, nkeys);
    for (int l = warp; l < nv; l += kDecThreads / 32) {
      float m = -INFINITY, sum = 0.f;
      for (int jj = lane; jj < nkeys; jj += 32) {
        const float s = sc[l * tile + jj];
        const float m_new = fmaxf(m, s);
        sum = sum * fast_exp2(m - m_new) + fast_exp2(s - m_new);
        m = m_new;
      }
      const float m0 = lstat[2 * l], s0 = lstat[2 * l + 1];      // running value of the earlier tiles, merged last
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, sum, off);
        const float m_new = fmaxf(m, m2);
        sum = (m == -INFINITY ? 0.f : sum * fast_exp2(m - m_new)) + (m2 == -INFINITY ? 0.f : s2 * fast_exp2(m2 - m_new));
        m = m_new;
      }
      const float m_new = fmaxf(m, m0);
      sum = (m == -INFINITY ? 0.f : sum * fast_exp2(m - m_new)) + (m0 == -INFINITY ? 0.f : s0 * fast_exp2(m0 - m_new));
      __syncwarp();
      if (lane == 0) lstat[2 * l] = m_new, lstat[2 * l + 1] = sum;
    }
    if (!single_tile) __syncthreads();               // sc is overwritten by the next tile
  }
  if (p.nsplit > 1) cluster.sync(); else __syncthreads();
  for (int l = tid; l < nv; l += kDecThreads) {      // merge the cluster's partial statistics in rank order
    float m = -INFINITY, sum = 0.f;
    for (int r = 0; r < p.nsplit; ++r) {
      const float* peer = p.nsplit > 1 ? cluster.map_shared_rank(lstat, r) : lstat;
      const float m2 = peer[2 * l], s2 = peer[2 * l + 1];
      const float m_new = fmaxf(m, m2);
      sum = (m == -INFINITY ? 0.f : sum * fast_exp2(m - m_new)) + (m2 == -INFINITY ? 0.f : s2 * fast_exp2(m2 - m_new));
      m = m_new;
    }
    stat[2 * l] = m, stat[2 * l + 1] = sum > 0.f ? 1.f / sum : 0.f;
  }
  __syncthreads();

  // ---- pass B set-up: this thread's 8 output columns and its row slot ----
  const int slot = tid % kLanes, rslot = tid / kLanes;
  const int col0 = chunk * kCols + slot * 8;
  const bool col_ok = col0 < p.d;                     // d is a multiple of 8
  const uint4* tb = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.table) + (col_ok ? col0 : 0));
  const int64_t* ids = p.ids + b * p.ids_batch_stride;
  const uint32_t row_words = static_cast<uint32_t>(p.d / 8);          // 16-byte words per table row
  const uint32_t tok_words = row_words * static_cast<uint32_t>(nv);   // ... per token (host checks vocab * tok_words < 2^32)
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  for (int j0 = j_begin; j0 < j_end; j0 += tile) {
    const int nkeys = min(tile, j_end - j0);
    if (!single_tile) compute_scores(j0, nkeys);
    for (int l = warp; l < nv; l += kDecThreads / 32) {   // normalise in place
      const float m = stat[2 * l], inv = stat[2 * l + 1];
      for (int jj = lane; jj < nkeys; jj += 32) sc[l * tile + jj] = fast_exp2(sc[l * tile + jj] - m) * inv;
    }
    for (int jj = tid; jj < nkeys; jj += kDecThreads)
      keyoff[jj] = static_cast<uint32_t>(min(max(static_cast<int>(__ldg(ids + j0 + jj)), 0), p.vocab - 1)) * tok_words;
    __syncthreads();
    // acc += w[l][j] * table[ids[j], l, cols]; pairs (j, l) are walked key-major: the nv rows of a token are contiguous
    const int npairs = nkeys * nv;
    for (int pr0 = 0; pr0 < npairs; pr0 += kRows * kMixUnroll) {
      uint4 row[kMixUnroll];
      float w[kMixUnroll];
#pragma unroll
      for (int u = 0; u < kMixUnroll; ++u) {
        const int pr = pr0 + u * kRows + rslot;
        const bool live = pr < npairs;
        int jj, l;
        split_row(live ? pr : 0, jj, l);
        row[u] = __ldg(tb + (keyoff[jj] + static_cast<uint32_t>(l) * row_words));
        w[u] = live ? sc[l * tile + jj] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kMixUnroll; ++u) {
        float t[8];
        unpack8<kBF16>(row[u], t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w[u], t[i], acc[i]);
      }
    }
    __syncthreads();
  }
  // ---- merge the row slots (fixed order), then the cluster (rank order), and store ----
#pragma unroll
  for (int i = 0; i < 8; ++i) part[rslot * kCols + slot * 8 + i] = acc[i];
  __syncthreads();
  for (int c = tid; c < kCols; c += kDecThreads) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kRows; ++r) s += part[r * kCols + c];
    colsum[c] = s;
  }
  if (p.nsplit > 1) cluster.sync(); else __syncthreads();
  if (rank == 0) {
    for (int c = tid; c < kCols; c += kDecThreads) {
      if (chunk * kCols + c >= p.d) continue;
      float s = 0.f;
      for (int r = 0; r < p.nsplit; ++r) s += (p.nsplit > 1 ? cluster.map_shared_rank(colsum, r) : colsum)[c];
      uint16_t* op = static_cast<uint16_t*>(p.out) + static_cast<int64_t>(b) * p.d + chunk * kCols + c;
      if constexpr (kBF16) {
        *reinterpret_cast<__nv_bfloat16*>(op) = __float2bfloat16_rn(s);
      } else {
        *reinterpret_cast<__half*>(op) = __float2half_rn(s);
      }
    }
  }
  if (p.nsplit > 1) cluster.sync();        // peers keep lstat / colsum alive until everyone has read them
}

template <bool kBF16, int kLanes>
auto mix_kernel_for(bool pow2) {
  return pow2 ? sense_mix_decode_kernel<kBF16, kLanes, true> : sense_mix_decode_kernel<kBF16, kLanes, false>;
}

// number of key slices (cluster size) that brings the grid to about two CTAs per SM without slices under `min_keys`
inline int pick_nsplit(int64_t ctas, int len, int min_keys) {
  int n = static_cast<int>((2 * 148 + ctas - 1) / ctas);
  n = std::min(n, std::max(1, len / min_keys));
  return std::max(1, std::min(n, 8));
}

template <typename Kern, typename Params>
cudaError_t launch_cluster(Kern kern, dim3 grid, int threads, size_t smem, int nsplit, cudaStream_t st, const Params& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nsplit, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

}  // namespace decode
}  // namespace bp

extern "C" int bp_decode_attn_fwd(const void* q, const void* kv_cache, void* out, const int32_t* seqlens_k, int32_t batch,
                                  int32_t nheads, int32_t headdim, int32_t seqlen_k, int64_t kv_batch_stride,
                                  int64_t kv_row_stride, int64_t kv_which_stride, float softmax_scale, int32_t dtype,
                                  void* stream) {
  using namespace bp;
  const char* fn = "bp_decode_attn_fwd";
  if (!q || !kv_cache || !out) return fail(BP_ERR_INVALID_ARGUMENT, "%s: null pointer argument", fn);
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16) return fail(BP_ERR_INVALID_ARGUMENT, "%s: only fp16 and bf16 are supported", fn);
  if (batch <= 0 || nheads <= 0 || (seqlen_k <= 0 && !seqlens_k)) return fail(BP_ERR_INVALID_ARGUMENT, "%s: empty input", fn);
  if (headdim != 64 && headdim != 128)
    return fail(BP_ERR_UNSUPPORTED, "%s: head dim must be 64 or 128 (got %d)", fn, headdim);
  if (kv_batch_stride % 8 || kv_row_stride % 8 || kv_which_stride % 8 || (uintptr_t)q % 16 || (uintptr_t)kv_cache % 16 || (uintptr_t)out % 16)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: pointers must be 16-byte aligned and strides multiples of 8 elements", fn);
  if ((int64_t)batch * nheads > 0x7fffffff) return fail(BP_ERR_INVALID_ARGUMENT, "%s: batch * nheads too large", fn);
  decode::AttnParams p;
  p.q = q, p.kv = kv_cache, p.out = out, p.lens = seqlens_k;
  p.batch_stride = kv_batch_stride, p.row_stride = kv_row_stride, p.which_stride = kv_which_stride;
  p.batch = batch, p.nheads = nheads, p.len = seqlen_k;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  // with per-sequence lengths on the device the host only knows the batch: assume long contexts
  p.nsplit = decode::pick_nsplit(static_cast<int64_t>(batch) * nheads, seqlens_k ? (1 << 20) : seqlen_k, 64);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bf = dtype == BP_DTYPE_BF16;
  const dim3 grid(batch * nheads * p.nsplit);
  auto kern = headdim == 64 ? (bf ? decode::decode_attn_kernel<64, true> : decode::decode_attn_kernel<64, false>)
                            : (bf ? decode::decode_attn_kernel<128, true> : decode::decode_attn_kernel<128, false>);
  const cudaError_t e = decode::launch_cluster(kern, grid, 128, 0, p.nsplit, st, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "%s: launch failed: %s", fn, cudaGetErrorString(e));
  }
  return check_launch(fn);
}

extern "C" int bp_sense_mix_decode_fwd(const void* q, const void* k_cache, const int64_t* ids, const void* table, void* out,
                                       const int32_t* seqlens, int32_t batch, int32_t seqlen, int32_t nv, int32_t dk,
                                       int32_t d, int32_t vocab, int64_t k_batch_stride, int64_t ids_batch_stride,
                                       float softmax_scale, int32_t dtype, void* stream) {
  using namespace bp;
  const char* fn = "bp_sense_mix_decode_fwd";
  if (!q || !k_cache || !ids || !table || !out) return fail(BP_ERR_INVALID_ARGUMENT, "%s: null pointer argument", fn);
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16) return fail(BP_ERR_INVALID_ARGUMENT, "%s: only fp16 and bf16 are supported", fn);
  if (batch <= 0 || batch > 65535 || nv <= 0 || (seqlen <= 0 && !seqlens) || vocab <= 0)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: empty input or batch > 65535", fn);
  if (dk % 8 || d % 8) return fail(BP_ERR_UNSUPPORTED, "%s: dk and d must be multiples of 8 (got %d, %d)", fn, dk, d);
  if ((uintptr_t)q % 16 || (uintptr_t)k_cache % 16 || (uintptr_t)table % 16 || k_batch_stride % 8)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: pointers must be 16-byte aligned", fn);
  decode::MixParams p;
  p.q = q, p.kcache = k_cache, p.ids = ids, p.table = table, p.out = out, p.lens = seqlens;
  p.k_batch_stride = k_batch_stride, p.ids_batch_stride = ids_batch_stride;
  p.batch = batch, p.nv = nv, p.dk = dk, p.d = d, p.vocab = vocab, p.len = seqlen;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  if (static_cast<int64_t>(vocab) * nv * (d / 8) >= (int64_t{1} << 32))
    return fail(BP_ERR_UNSUPPORTED, "%s: the table must be smaller than 64 GB (32-bit offsets in 16-byte units)", fn);
  // a whole warp per table row (256-column chunks) when batch x column chunks x key slices fill the SMs
  // (counting the key slices a long context allows), 8 lanes (64 columns) otherwise
  const int len_hint = seqlens ? (1 << 20) : seqlen;
  const int lanes = static_cast<int64_t>(batch) * ((d + 255) / 256) * std::min(8, std::max(1, len_hint / 128)) >= 148 ? 32 : 8;
  const int cols = 8 * lanes, chunks = (d + cols - 1) / cols;
  p.nsplit = decode::pick_nsplit(static_cast<int64_t>(batch) * chunks, len_hint, 32);
  const int per = ((len_hint + p.nsplit - 1) / p.nsplit + 15) & ~15;
  const int tile_max = std::max(16, (64 * 1024 / (4 * nv)) & ~15);       // <= 64 KB of scores
  p.tile = std::min(per, tile_max);
  const size_t smem = sizeof(float) * decode::mix_smem_floats(nv, dk, p.tile, lanes);
  if (smem > 200 * 1024) return fail(BP_ERR_UNSUPPORTED, "%s: too many senses (%d) for the score tile", fn, nv);
  const bool bf = dtype == BP_DTYPE_BF16;
  const bool pow2 = (nv & (nv - 1)) == 0;
  p.nv_shift = 0;
  while (pow2 && (1 << p.nv_shift) < nv) ++p.nv_shift;
  auto kern = lanes == 32 ? (bf ? decode::mix_kernel_for<true, 32>(pow2) : decode::mix_kernel_for<false, 32>(pow2))
                          : (bf ? decode::mix_kernel_for<true, 8>(pow2) : decode::mix_kernel_for<false, 8>(pow2));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(BP_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(e));
    }
  }
  const cudaError_t e = decode::launch_cluster(kern, dim3(chunks * p.nsplit, batch), decode::kDecThreads, smem, p.nsplit,
                                               static_cast<cudaStream_t>(stream), p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "%s: launch failed: %s", fn, cudaGetErrorString(e));
  }
  return check_launch(fn);
}
