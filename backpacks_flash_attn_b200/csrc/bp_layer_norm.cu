// (Dropout +) residual add + LayerNorm forward for sm_100a.
//
// Replaces dropout_add_ln_fwd (csrc/layer_norm/ln_api.cpp:83-251, ln_fwd_kernels.cuh:20-191) without rowscale /
// colscale / subset.  Semantics kept from the reference:
// x = dropout(x0) + x1 in fp32, x stored in the residual dtype, statistics from the fp32 sum (mean, then the
// centred second moment), z = gamma * (x - mu) * rsigma + beta rounded once to the input dtype.
// Dropout (bp_ln_residual_fwd_dropout, training): the keep mask is counter-based like the attention kernels' --
// keep(row, col) = drop_keep(R[row], C[col]) with one word per row (a register) and one per column (shared memory,
// computed once per CTA) from bp_common.cuh's hash -- so nothing is stored for the backward, which regenerates it from
// the seed.  (The reference stores a byte mask per element, ln_fwd_kernels.cuh:96-110; only the distribution is part of
// the contract.)  ops/layer_norm.layer_norm_dropout_mask restates the mask in Python for the tests.
//
// HBM-bound: one warp owns one row and keeps it in registers (cols/32 values per lane), so every byte is
// read once and written once; 16-byte vector loads/stores, warp-shuffle reductions only, no smem, no
// block barrier.  Grid is a multiple of the SM count and warps stride over rows.
#include <type_traits>

#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace ln {

constexpr int kWarpsPerCta = 8;

#ifndef BP_LN_STREAMING
#define BP_LN_STREAMING 0
#endif
// streaming (evict-first) variants of the 16-byte accesses: the residual stream is touched once per kernel
__device__ __forceinline__ uint4 ld_stream(const void* p) {
#if BP_LN_STREAMING
  return __ldcs(reinterpret_cast<const uint4*>(p));
#else
  return *reinterpret_cast<const uint4*>(p);
#endif
}
__device__ __forceinline__ void st_stream(void* p, uint4 v) {
#if BP_LN_STREAMING
  __stcs(reinterpret_cast<uint4*>(p), v);
#else
  *reinterpret_cast<uint4*>(p) = v;
#endif
}

template <typename T>
struct Vec8;  // 8 consecutive elements
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ void load_stream(const float* p) {   // global memory only
    const uint4 u = ld_stream(p), w = ld_stream(p + 4);
    a = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
    b = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
  }
  __device__ void store_stream(float* p) const {
    st_stream(p, make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)));
    st_stream(p + 4, make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)));
  }
  __device__ void to(float (&f)[8]) const {
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
  }
  __device__ void from(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  uint4 u;
  __device__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ void load_stream(const __nv_bfloat16* p) { u = ld_stream(p); }
  __device__ void store_stream(__nv_bfloat16* p) const { st_stream(p, u); }
  __device__ void to(float (&f)[8]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ void from(const float (&f)[8]) {
    u.x = pack2<true>(f[0], f[1]);
    u.y = pack2<true>(f[2], f[3]);
    u.z = pack2<true>(f[4], f[5]);
    u.w = pack2<true>(f[6], f[7]);
  }
};
template <>
struct Vec8<__half> {
  uint4 u;
  __device__ void load(const __half* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ void load_stream(const __half* p) { u = ld_stream(p); }
  __device__ void store_stream(__half* p) const { st_stream(p, u); }
  __device__ void to(float (&f)[8]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ void from(const float (&f)[8]) {
    u.x = pack2<false>(f[0], f[1]);
    u.y = pack2<false>(f[2], f[3]);
    u.z = pack2<false>(f[4], f[5]);
    u.w = pack2<false>(f[6], f[7]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV = 8-element vectors per lane; a row has cols/8 vectors, vector index = i*32 + lane.
// gamma / beta live in shared memory in their storage type (read back 16 bytes at a time, conflict-free), not
// in registers: with the row itself (NV x 8 floats) that keeps the kernel under 85 registers, so 24 warps per
// SM keep enough loads in flight to cover the HBM latency.
struct DropArgs {
  uint32_t base = 0, thr24 = 0;   // thr24 == 0: no dropout
  float scale = 1.f;
};

template <typename X, typename R, typename W, int NV, typename Z = X, bool kDrop = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
ln_residual_fwd_kernel(const X* __restrict__ x0, const R* __restrict__ x1, const W* __restrict__ gamma,
                       const W* __restrict__ beta, Z* __restrict__ z, R* __restrict__ x_out,
                       float* __restrict__ mu_out, float* __restrict__ rs_out, int64_t rows, int cols, float eps,
                       DropArgs drop) {
  extern __shared__ uint4 ln_smem[];   // [cols] gamma then [cols] beta, in W; with dropout [cols] column words
  W* s_gamma = reinterpret_cast<W*>(ln_smem);
  W* s_beta = s_gamma + cols;
  uint32_t* s_cw = reinterpret_cast<uint32_t*>(s_beta + cols);
  if constexpr (kDrop)
    for (int c = threadIdx.x; c < cols; c += blockDim.x) s_cw[c] = drop_col_word(drop.base, static_cast<uint32_t>(c));
  const int lane = threadIdx.x & 31;
  const int nvec = cols >> 3;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * kWarpsPerCta;
  const float inv_cols = 1.f / static_cast<float>(cols);

  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    Vec8<W> t;
    t.load(gamma + v * 8);
    t.store(s_gamma + v * 8);
    t.load(beta + v * 8);
    t.store(s_beta + v * 8);
  }
  __syncthreads();

  for (int64_t row = warp_global; row < rows; row += warp_stride) {
    const int64_t base = row * cols;
    float x[NV][8];
    // all loads of the row are issued before the first use
    Vec8<X> a[NV];
    Vec8<R> r[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = i * 32 + lane;
      if (v < nvec) {
        a[i].load_stream(x0 + base + v * 8);
        if (x1 != nullptr) r[i].load_stream(x1 + base + v * 8);
      }
    }
    float sum = 0.f;
    uint32_t rw = 0;
    if constexpr (kDrop) rw = drop_row_word(drop.base, static_cast<uint32_t>(row));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = i * 32 + lane;
      if (v < nvec) {
        a[i].to(x[i]);
        if constexpr (kDrop) {
          const uint4 c0 = *reinterpret_cast<const uint4*>(s_cw + v * 8), c1 = *reinterpret_cast<const uint4*>(s_cw + v * 8 + 4);
          const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) x[i][k] = drop_keep(rw, cw[k], drop.thr24) ? x[i][k] * drop.scale : 0.f;
        }
        if (x1 != nullptr) {
          float rf[8];
          r[i].to(rf);
#pragma unroll
          for (int k = 0; k < 8; ++k) x[i][k] += rf[k];
        }
        if (x_out != nullptr) {
          Vec8<R> o;
          o.from(x[i]);
          o.store_stream(x_out + base + v * 8);
          // the reference normalises the value it stored (ln_fwd_kernels.cuh keeps x in compute type;
          // with a 16-bit residual stream the stored value is the rounded one) -- keep fp32 here.
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += x[i][k];
      }
    }
    const float mu = warp_sum(sum) * inv_cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i * 32 + lane < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = x[i][k] - mu;
          sq += d * d;
        }
      }
    }
    const float rs = rsqrtf(warp_sum(sq) * inv_cols + eps);
    if (lane == 0) {
      if (mu_out) mu_out[row] = mu;
      if (rs_out) rs_out[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = i * 32 + lane;
      if (v < nvec) {
        float g[8], b[8], y[8];
        Vec8<W> t;
        t.load(s_gamma + v * 8);
        t.to(g);
        t.load(s_beta + v * 8);
        t.to(b);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = g[k] * ((x[i][k] - mu) * rs) + b[k];
        Vec8<Z> o;
        o.from(y);
        o.store(z + base + v * 8);
      }
    }
  }
}

template <typename X, typename R, typename W, int NV, typename Z, bool kDrop>
int launch_kernel(const void* x0, const void* x1, const void* gamma, const void* beta, void* z, void* x_out, float* mu,
                  float* rs, int64_t rows, int cols, float eps, DropArgs drop, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas_needed = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = static_cast<int64_t>(sms) * 3;  // 3 resident CTAs of 256 threads per SM (register-limited)
  const int grid = static_cast<int>(ctas_needed < cap ? ctas_needed : cap);
  const size_t smem = 2 * static_cast<size_t>(cols) * sizeof(W) + (kDrop ? static_cast<size_t>(cols) * 4 : 0);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ln_residual_fwd_kernel<X, R, W, NV, Z, kDrop>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(BP_ERR_CUDA, "bp_ln_residual_fwd: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    }
  }
  ln_residual_fwd_kernel<X, R, W, NV, Z, kDrop><<<grid, kWarpsPerCta * 32, smem, st>>>(
      static_cast<const X*>(x0), static_cast<const R*>(x1), static_cast<const W*>(gamma),
      static_cast<const W*>(beta), static_cast<Z*>(z), static_cast<R*>(x_out), mu, rs, rows, cols, eps, drop);
  return check_launch("bp_ln_residual_fwd launch");
}

template <typename X, typename R, typename W, int NV, typename Z = X>
int launch_nv(const void* x0, const void* x1, const void* gamma, const void* beta, void* z, void* x_out, float* mu,
              float* rs, int64_t rows, int cols, float eps, DropArgs drop, cudaStream_t st) {
  if constexpr (std::is_same<Z, X>::value) {   // dropout only exists on the residual-add entry point
    if (drop.thr24 != 0)
      return launch_kernel<X, R, W, NV, Z, true>(x0, x1, gamma, beta, z, x_out, mu, rs, rows, cols, eps, drop, st);
  }
  return launch_kernel<X, R, W, NV, Z, false>(x0, x1, gamma, beta, z, x_out, mu, rs, rows, cols, eps, drop, st);
}

template <typename X, typename R, typename W, typename Z = X>
int launch(const void* x0, const void* x1, const void* gamma, const void* beta, void* z, void* x_out, float* mu,
           float* rs, int64_t rows, int cols, float eps, cudaStream_t st, DropArgs drop = DropArgs()) {
  const int nv = (cols / 8 + 31) / 32;
#define BP_LN_CASE(N) \
  if (nv <= N) return launch_nv<X, R, W, N, Z>(x0, x1, gamma, beta, z, x_out, mu, rs, rows, cols, eps, drop, st)
  BP_LN_CASE(1);
  BP_LN_CASE(2);
  BP_LN_CASE(3);
  BP_LN_CASE(4);
  BP_LN_CASE(6);
  BP_LN_CASE(8);
  BP_LN_CASE(16);
  BP_LN_CASE(32);
#undef BP_LN_CASE
  return fail(BP_ERR_UNSUPPORTED, "bp_ln_residual_fwd: hidden size %d > 8192 is not supported", cols);
}

}  // namespace ln

int ln_drop_args(float p, uint64_t seed, int64_t rows, const char* fn, uint32_t* base, uint32_t* thr24, float* scale) {
  if (!(p >= 0.f) || p >= 1.f) return fail(BP_ERR_INVALID_ARGUMENT, "%s: dropout_p must be in [0, 1) (got %f)", fn, (double)p);
  if (rows > 0xFFFFFFFFll) return fail(BP_ERR_UNSUPPORTED, "%s: dropout supports at most 2^32 - 1 rows", fn);
  const int thr = drop_threshold(p);
  *base = drop_base(seed, 0x4C4Eu);   // "LN": keeps these masks apart from the attention kernels' (bh < 2^31 there)
  *thr24 = static_cast<uint32_t>(thr) << 24;
  *scale = 256.f / static_cast<float>(256 - thr);
  return BP_OK;
}
}  // namespace bp

namespace {
int ln_residual_fwd_impl(const void* x0, const void* x1, const void* gamma, const void* beta, void* z,
                         void* x_out, float* mu, float* rsigma, int64_t rows, int32_t cols,
                         float epsilon, int32_t x0_dtype, int32_t residual_dtype, int32_t weight_dtype,
                         bp::ln::DropArgs drop, void* stream) {
  using namespace bp;
  if (!x0 || !gamma || !beta || !z) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_fwd: null pointer argument");
  if (rows <= 0 || cols <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_fwd: empty input");
  if (cols % 8 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_fwd: hidden size must be a multiple of 8 (got %d)", cols);
  const uintptr_t ptrs[] = {(uintptr_t)x0, (uintptr_t)x1, (uintptr_t)gamma, (uintptr_t)beta, (uintptr_t)z, (uintptr_t)x_out};
  for (uintptr_t a : ptrs)
    if (a % 16 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int key = x0_dtype * 100 + residual_dtype * 10 + weight_dtype;
#define BP_LN_DISPATCH(XD, RD, WD, X, R, W) \
  if (key == XD * 100 + RD * 10 + WD)       \
  return ln::launch<X, R, W>(x0, x1, gamma, beta, z, x_out, mu, rsigma, rows, cols, epsilon, st, drop)
  using bf = __nv_bfloat16;
  using hf = __half;
  BP_LN_DISPATCH(BP_DTYPE_BF16, BP_DTYPE_F32, BP_DTYPE_BF16, bf, float, bf);
  BP_LN_DISPATCH(BP_DTYPE_F16, BP_DTYPE_F32, BP_DTYPE_F16, hf, float, hf);
  BP_LN_DISPATCH(BP_DTYPE_BF16, BP_DTYPE_BF16, BP_DTYPE_BF16, bf, bf, bf);
  BP_LN_DISPATCH(BP_DTYPE_F16, BP_DTYPE_F16, BP_DTYPE_F16, hf, hf, hf);
  BP_LN_DISPATCH(BP_DTYPE_F32, BP_DTYPE_F32, BP_DTYPE_F32, float, float, float);
  BP_LN_DISPATCH(BP_DTYPE_BF16, BP_DTYPE_F32, BP_DTYPE_F32, bf, float, float);
  BP_LN_DISPATCH(BP_DTYPE_F16, BP_DTYPE_F32, BP_DTYPE_F32, hf, float, float);
#undef BP_LN_DISPATCH
  return fail(BP_ERR_UNSUPPORTED, "bp_ln_residual_fwd: dtype combination (x0=%d, residual=%d, weight=%d) not built",
              x0_dtype, residual_dtype, weight_dtype);
}
}  // namespace

extern "C" int bp_ln_residual_fwd(const void* x0, const void* x1, const void* gamma, const void* beta, void* z,
                                  void* x_out, float* mu, float* rsigma, int64_t rows, int32_t cols,
                                  float epsilon, int32_t x0_dtype, int32_t residual_dtype, int32_t weight_dtype,
                                  void* stream) {
  return ln_residual_fwd_impl(x0, x1, gamma, beta, z, x_out, mu, rsigma, rows, cols, epsilon, x0_dtype, residual_dtype,
                              weight_dtype, bp::ln::DropArgs(), stream);
}

extern "C" int bp_ln_residual_fwd_dropout(const void* x0, const void* x1, const void* gamma, const void* beta, void* z,
                                          void* x_out, float* mu, float* rsigma, int64_t rows, int32_t cols,
                                          float epsilon, int32_t x0_dtype, int32_t residual_dtype, int32_t weight_dtype,
                                          float dropout_p, uint64_t seed, void* stream) {
  bp::ln::DropArgs drop;
  if (int rc = bp::ln_drop_args(dropout_p, seed, rows, "bp_ln_residual_fwd_dropout", &drop.base, &drop.thr24, &drop.scale))
    return rc;
  return ln_residual_fwd_impl(x0, x1, gamma, beta, z, x_out, mu, rsigma, rows, cols, epsilon, x0_dtype, residual_dtype,
                              weight_dtype, drop, stream);
}

extern "C" int bp_ln_fwd(const void* x, const void* gamma, const void* beta, void* z, float* mu, float* rsigma,
                         int64_t rows, int32_t cols, float epsilon, int32_t x_dtype, int32_t z_dtype,
                         int32_t weight_dtype, void* stream) {
  using namespace bp;
  if (!x || !gamma || !beta || !z) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_fwd: null pointer argument");
  if (rows <= 0 || cols <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_fwd: empty input");
  if (cols % 8 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_fwd: hidden size must be a multiple of 8 (got %d)", cols);
  const uintptr_t ptrs[] = {(uintptr_t)x, (uintptr_t)gamma, (uintptr_t)beta, (uintptr_t)z};
  for (uintptr_t a : ptrs)
    if (a % 16 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == z_dtype)   // same-type LayerNorm: the residual-add kernel without a residual
    return bp_ln_residual_fwd(x, nullptr, gamma, beta, z, nullptr, mu, rsigma, rows, cols, epsilon, x_dtype,
                              x_dtype == BP_DTYPE_F32 ? BP_DTYPE_F32 : x_dtype, weight_dtype, stream);
  using bf = __nv_bfloat16;
  using hf = __half;
  if (x_dtype == BP_DTYPE_F32 && z_dtype == BP_DTYPE_BF16 && weight_dtype == BP_DTYPE_BF16)
    return ln::launch<float, float, bf, bf>(x, nullptr, gamma, beta, z, nullptr, mu, rsigma, rows, cols, epsilon, st);
  if (x_dtype == BP_DTYPE_F32 && z_dtype == BP_DTYPE_F16 && weight_dtype == BP_DTYPE_F16)
    return ln::launch<float, float, hf, hf>(x, nullptr, gamma, beta, z, nullptr, mu, rsigma, rows, cols, epsilon, st);
  return fail(BP_ERR_UNSUPPORTED, "bp_ln_fwd: dtype combination (x=%d, z=%d, weight=%d) not built", x_dtype, z_dtype,
              weight_dtype);
}
