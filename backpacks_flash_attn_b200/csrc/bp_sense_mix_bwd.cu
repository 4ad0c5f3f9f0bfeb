// Backward of the Backpack sense-mix, the element-wise half: causal softmax + softmax backward in ONE pass, in place.
//
// The reference trains through the eager composition  alpha = softmax(q k^T scale + mask);  o = sum_l alpha_l C_l
// (training/src/models/backpack.py:116-122, 313), whose autograd graph makes seven full passes over (b, nv, s, s)
// tensors (mask add, softmax, softmax backward and their saved copies).  The backward of ops/sense_mix.py keeps the five
// GEMM-shaped products on the library's batched GEMM
//     S_l = q_l k_l^T      dA_l = dO C_l^T      dC_l = P_l^T dO      dq_l = dS_l k_l      dk_l = dS_l^T q_l
// and this kernel does everything between them, one warp per score row (row = (sense, batch, query t)), rows of up to
// 2048 keys held in registers (longer rows: a three-pass variant):
//     P    = softmax_j<=t (scale * S)                 written over S  (16-bit, zeros right of the diagonal)
//     dS'  = scale * P o (dA - sum_j P o dA)          written over dA (16-bit, zeros right of the diagonal)
// Only the causal part of a row is read (whatever S and dA hold right of the diagonal is ignored); the whole row is
// written, because the GEMMs that follow read full rows.  HBM-bound: (1/2 + 1/2 + 1 + 1) x 2 bytes per score.
#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace smb {

constexpr int kWarps = 8;
#ifndef BP_SSB_MIN_CTAS
#define BP_SSB_MIN_CTAS 3   // resident CTAs per SM asked of ptxas for rows of up to 1024 keys (80 registers, no spills)
#endif
#ifndef BP_SSB_MIN_CTAS_WIDE
#define BP_SSB_MIN_CTAS_WIDE 2   // ... and for rows of up to 2048 keys (128 registers, 92 bytes of spills: 2.9 -> 4.2 TB/s)
#endif

template <bool kBF16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
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

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV = 8-wide vectors per lane: rows of up to 256 * NV keys.
template <bool kBF16, int NV>
__global__ void __launch_bounds__(kWarps * 32, NV <= 4 ? BP_SSB_MIN_CTAS : BP_SSB_MIN_CTAS_WIDE)
sense_softmax_bwd_kernel(uint16_t* __restrict__ S, uint16_t* __restrict__ dA, int64_t rows, int seqlen, float scale,
                         float scale_log2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kWarps;
  for (int64_t r = warp0; r < rows; r += stride) {
    const int t = static_cast<int>(r % seqlen);          // causal: keys 0..t
    uint16_t* srow = S + r * seqlen;
    uint16_t* drow = dA + r * seqlen;
    float x[NV][8];
    uint4 da[NV];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = (lane + 32 * i) * 8;
      if (c0 < seqlen && c0 <= t) {
        const uint4 u = *reinterpret_cast<const uint4*>(srow + c0);
        const uint4 g4 = *reinterpret_cast<const uint4*>(drow + c0);
        uint32_t w[4] = {g4.x, g4.y, g4.z, g4.w};
        unpack8<kBF16>(u, x[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const bool vis = c0 + k <= t;
          x[i][k] = vis ? x[i][k] * scale_log2 : -INFINITY;
          if (!vis) w[k >> 1] &= (k & 1) ? 0x0000FFFFu : 0xFFFF0000u;   // right of the diagonal: ignored
          m = fmaxf(m, x[i][k]);
        }
        da[i] = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        da[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[i][k] = -INFINITY;
      }
    }
    m = warp_max(m);                                      // finite: key 0 is always visible
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        x[i][k] = exp2f(x[i][k] - m);
        l += x[i][k];
      }
    const float inv = 1.f / warp_sum(l);
    // P rounded to the storage type first: delta and dS use the values the GEMMs will see (as the eager chain does).
    // x keeps the ROUNDED probabilities (re-packing them at the store is exact), so no packed copy stays live.
    float delta = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint4 pb;
      pb.x = pack2<kBF16>(x[i][0] * inv, x[i][1] * inv);
      pb.y = pack2<kBF16>(x[i][2] * inv, x[i][3] * inv);
      pb.z = pack2<kBF16>(x[i][4] * inv, x[i][5] * inv);
      pb.w = pack2<kBF16>(x[i][6] * inv, x[i][7] * inv);
      float g[8];
      unpack8<kBF16>(pb, x[i]);
      unpack8<kBF16>(da[i], g);
#pragma unroll
      for (int k = 0; k < 8; ++k) delta = fmaf(x[i][k], g[k], delta);
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = (lane + 32 * i) * 8;
      if (c0 < seqlen) {
        float g[8];
        unpack8<kBF16>(da[i], g);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = scale * x[i][k] * (g[k] - delta);
        uint4 pb, o;
        pb.x = pack2<kBF16>(x[i][0], x[i][1]);
        pb.y = pack2<kBF16>(x[i][2], x[i][3]);
        pb.z = pack2<kBF16>(x[i][4], x[i][5]);
        pb.w = pack2<kBF16>(x[i][6], x[i][7]);
        o.x = pack2<kBF16>(g[0], g[1]);
        o.y = pack2<kBF16>(g[2], g[3]);
        o.z = pack2<kBF16>(g[4], g[5]);
        o.w = pack2<kBF16>(g[6], g[7]);
        *reinterpret_cast<uint4*>(srow + c0) = pb;
        *reinterpret_cast<uint4*>(drow + c0) = o;
      }
    }
  }
}

// Rows longer than 2048 keys do not fit in registers: three passes over the causal part of the row (running max and sum;
// delta; write), each lane re-reading its own 16-byte vectors (L1 / L2 hits: a row is read from HBM once).  In place is
// still safe: a vector is overwritten by the lane that has just read it, in the last pass only.
template <bool kBF16>
__global__ void __launch_bounds__(kWarps * 32)
sense_softmax_bwd_long_kernel(uint16_t* __restrict__ S, uint16_t* __restrict__ dA, int64_t rows, int seqlen, float scale,
                              float scale_log2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kWarps;
  for (int64_t r = warp0; r < rows; r += stride) {
    const int t = static_cast<int>(r % seqlen);
    uint16_t* srow = S + r * seqlen;
    uint16_t* drow = dA + r * seqlen;
    float m = -INFINITY, l = 0.f;
    for (int c0 = lane * 8; c0 <= t; c0 += 256) {
      float x[8];
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(srow + c0), x);
      float cm = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        x[k] = (c0 + k <= t) ? x[k] * scale_log2 : -INFINITY;
        cm = fmaxf(cm, x[k]);
      }
      const float mn = fmaxf(m, cm);          // finite: element c0 itself is visible
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += exp2f(x[k] - mn);
      l = l * exp2f(m - mn) + a;
      m = mn;
    }
    const float mw = warp_max(m);             // lanes without a visible vector hold m = -inf, l = 0
    l = warp_sum(m == -INFINITY ? 0.f : l * exp2f(m - mw));
    const float inv = 1.f / l;
    float delta = 0.f;
    for (int c0 = lane * 8; c0 <= t; c0 += 256) {
      float x[8], g[8];
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(srow + c0), x);
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(drow + c0), g);
      float pr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pr[k] = (c0 + k <= t) ? exp2f(x[k] * scale_log2 - mw) * inv : 0.f;
      uint4 pb;
      pb.x = pack2<kBF16>(pr[0], pr[1]), pb.y = pack2<kBF16>(pr[2], pr[3]);
      pb.z = pack2<kBF16>(pr[4], pr[5]), pb.w = pack2<kBF16>(pr[6], pr[7]);
      unpack8<kBF16>(pb, pr);
#pragma unroll
      for (int k = 0; k < 8; ++k) delta = (c0 + k <= t) ? fmaf(pr[k], g[k], delta) : delta;
    }
    delta = warp_sum(delta);
    for (int c0 = lane * 8; c0 < seqlen; c0 += 256) {
      uint4 pb = make_uint4(0u, 0u, 0u, 0u), o = make_uint4(0u, 0u, 0u, 0u);
      if (c0 <= t) {
        float x[8], g[8], pr[8];
        unpack8<kBF16>(*reinterpret_cast<const uint4*>(srow + c0), x);
        unpack8<kBF16>(*reinterpret_cast<const uint4*>(drow + c0), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) pr[k] = (c0 + k <= t) ? exp2f(x[k] * scale_log2 - mw) * inv : 0.f;
        pb.x = pack2<kBF16>(pr[0], pr[1]), pb.y = pack2<kBF16>(pr[2], pr[3]);
        pb.z = pack2<kBF16>(pr[4], pr[5]), pb.w = pack2<kBF16>(pr[6], pr[7]);
        unpack8<kBF16>(pb, pr);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = (c0 + k <= t) ? scale * pr[k] * (g[k] - delta) : 0.f;
        o.x = pack2<kBF16>(g[0], g[1]), o.y = pack2<kBF16>(g[2], g[3]);
        o.z = pack2<kBF16>(g[4], g[5]), o.w = pack2<kBF16>(g[6], g[7]);
      }
      *reinterpret_cast<uint4*>(srow + c0) = pb;
      *reinterpret_cast<uint4*>(drow + c0) = o;
    }
  }
}

template <bool kBF16, int NV>
int launch(void* S, void* dA, int64_t rows, int seqlen, float scale, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (rows + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(sms) * 32;
  const int grid = static_cast<int>(want < cap ? want : cap);
  sense_softmax_bwd_kernel<kBF16, NV><<<grid, kWarps * 32, 0, st>>>(
      static_cast<uint16_t*>(S), static_cast<uint16_t*>(dA), rows, seqlen, scale, scale * 1.4426950408889634f);
  return check_launch("bp_sense_softmax_bwd launch");
}

template <bool kBF16>
int dispatch(void* S, void* dA, int64_t rows, int seqlen, float scale, cudaStream_t st) {
  if (seqlen <= 256) return launch<kBF16, 1>(S, dA, rows, seqlen, scale, st);
  if (seqlen <= 512) return launch<kBF16, 2>(S, dA, rows, seqlen, scale, st);
  if (seqlen <= 1024) return launch<kBF16, 4>(S, dA, rows, seqlen, scale, st);
  if (seqlen <= 2048) return launch<kBF16, 8>(S, dA, rows, seqlen, scale, st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (rows + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(sms) * 32;
  sense_softmax_bwd_long_kernel<kBF16><<<static_cast<int>(want < cap ? want : cap), kWarps * 32, 0, st>>>(
      static_cast<uint16_t*>(S), static_cast<uint16_t*>(dA), rows, seqlen, scale, scale * 1.4426950408889634f);
  return check_launch("bp_sense_softmax_bwd launch");
}

}  // namespace smb
}  // namespace bp

extern "C" int bp_sense_softmax_bwd(void* scores_probs, void* dalpha_dscores, int64_t rows, int32_t seqlen,
                                    float softmax_scale, int32_t dtype, void* stream) {
  using namespace bp;
  if (!scores_probs || !dalpha_dscores) return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_softmax_bwd: null pointer argument");
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_softmax_bwd: only fp16 and bf16 are supported (dtype=%d)", dtype);
  if (rows <= 0 || seqlen <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_softmax_bwd: empty input");
  if (seqlen % 8 != 0 || seqlen > 8192)
    return fail(BP_ERR_UNSUPPORTED, "bp_sense_softmax_bwd: seqlen must be a multiple of 8, at most 8192 (got %d)", seqlen);
  if (rows % seqlen != 0)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_softmax_bwd: rows (%lld) must be a multiple of seqlen (%d): square causal score matrices",
                (long long)rows, seqlen);
  if (reinterpret_cast<uintptr_t>(scores_probs) % 16 || reinterpret_cast<uintptr_t>(dalpha_dscores) % 16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_sense_softmax_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == BP_DTYPE_BF16 ? smb::dispatch<true>(scores_probs, dalpha_dscores, rows, seqlen, softmax_scale, st)
                                : smb::dispatch<false>(scores_probs, dalpha_dscores, rows, seqlen, softmax_scale, st);
}
