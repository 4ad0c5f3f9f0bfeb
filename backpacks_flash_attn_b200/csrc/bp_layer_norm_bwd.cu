// Residual add + LayerNorm backward for sm_100a.
//
// Replaces dropout_add_ln_bwd (csrc/layer_norm/ln_api.cpp:255-440, ln_bwd_kernels.cuh) without rowscale / colscale /
// subset.  With x = dropout(x0) + x1 the saved pre-norm sum, y = (x - mu) * rsigma, z = gamma * y + beta:
//     dy = dz * gamma;   dx = rsigma * (dy - mean(dy) - y * mean(dy * y)) + dx_residual
//     dx0 = dx (input dtype; bp_ln_residual_bwd_dropout: keep ? dx / (1 - p) : 0, the forward's mask regenerated from
//     the seed, bp_layer_norm.cu), dx1 = dx (residual dtype);   dgamma = sum_rows dz * y,  dbeta = sum_rows dz
// (`dx_residual` is the gradient that arrives through the residual output of a pre-norm block.)
//
// HBM-bound like the forward: one warp owns a row and keeps it in registers, 16-byte accesses, shuffle reductions; mu
// and rsigma come from the forward when the caller saved them (two loads instead of two of the four dependent shuffle
// reductions of a row, and exactly the forward's statistics), else they are recomputed from the row.  dgamma / dbeta are
// accumulated per lane across the rows a warp visits, reduced across the CTA's warps through shared memory in a fixed
// order, written as one fp32 partial row per CTA and summed by a second tiny kernel: deterministic, no atomics (the
// reference uses the same two-stage scheme, ln_bwd_kernels.cuh + ln_bwd_finalize_kernel).
#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace lnb {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxCtas = 1024;   // workspace rows

template <typename T>
struct V8;
template <>
struct V8<float> {
  static __device__ void load(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
  }
  static __device__ void store(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <>
struct V8<__nv_bfloat16> {
  static __device__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack2<true>(f[0], f[1]);
    u.y = pack2<true>(f[2], f[3]);
    u.z = pack2<true>(f[4], f[5]);
    u.w = pack2<true>(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct V8<__half> {
  static __device__ void load(const __half* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ void store(__half* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack2<false>(f[0], f[1]);
    u.y = pack2<false>(f[2], f[3]);
    u.z = pack2<false>(f[4], f[5]);
    u.w = pack2<false>(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// X: dtype of dz / dx0; R: dtype of the saved sum x, of dx_residual and of dx1; W: dtype of gamma.
struct DropArgs {
  uint32_t base = 0, thr24 = 0;   // thr24 == 0: no dropout
  float scale = 1.f;
};

template <typename X, typename R, typename W, int NV, bool kDrop>
__global__ void __launch_bounds__(kWarpsPerCta * 32, NV <= 3 ? 2 : 1)
ln_residual_bwd_kernel(const X* __restrict__ dz, const R* __restrict__ dxres, const R* __restrict__ x,
                       const W* __restrict__ gamma, const float* __restrict__ mu_in, const float* __restrict__ rs_in,
                       X* __restrict__ dx0, R* __restrict__ dx1, float* __restrict__ part /* [gridDim.x][2][cols] */,
                       int64_t rows, int cols, float eps, DropArgs drop) {
  extern __shared__ float lnb_smem[];   // [kWarpsPerCta][cols] reduction buffer (dgamma, then dbeta), then [cols] gamma,
                                        // with dropout [cols] column words of the keep mask
  float* s_gamma = lnb_smem + kWarpsPerCta * cols;
  uint32_t* s_cw = reinterpret_cast<uint32_t*>(s_gamma + cols);
  if constexpr (kDrop)
    for (int c = threadIdx.x; c < cols; c += blockDim.x) s_cw[c] = drop_col_word(drop.base, static_cast<uint32_t>(c));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + warp;
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * kWarpsPerCta;
  const float inv_cols = 1.f / static_cast<float>(cols);

  // gamma lives in shared memory as fp32 (keeping it in registers next to the row, dz and the two column
  // accumulators spills at hidden size 768)
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float t[8];
    V8<W>::load(gamma + v * 8, t);
    V8<float>::store(s_gamma + v * 8, t);
  }
  __syncthreads();
  float dg[NV][8], db[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dg[i][k] = db[i][k] = 0.f;
  }

  for (int64_t row = warp_global; row < rows; row += warp_stride) {
    const int64_t base = row * cols;
    float xv[NV][8], dzv[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = i * 32 + lane;
      if (v < nvec) {
        V8<R>::load(x + base + v * 8, xv[i]);
        V8<X>::load(dz + base + v * 8, dzv[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += xv[i][k];
      }
    }
    float mu, rs;
    if (mu_in != nullptr) {   // (uniform branch) the forward's statistics
      mu = __ldg(mu_in + row);
      rs = __ldg(rs_in + row);
    } else {
      mu = warp_sum(sum) * inv_cols;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i * 32 + lane < nvec) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float d = xv[i][k] - mu;
            sq += d * d;
          }
        }
      rs = rsqrtf(warp_sum(sq) * inv_cols + eps);
    }
    // y in place of x; dy in place of dz; the two row means
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i * 32 + lane < nvec) {
        float g[8];
        V8<float>::load(s_gamma + (i * 32 + lane) * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float y = (xv[i][k] - mu) * rs;
          const float d = dzv[i][k];
          dg[i][k] += d * y;
          db[i][k] += d;
          const float dy = d * g[k];
          xv[i][k] = y;
          dzv[i][k] = dy;
          s1 += dy;
          s2 += dy * y;
        }
      }
    s1 = warp_sum(s1) * inv_cols;
    s2 = warp_sum(s2) * inv_cols;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = i * 32 + lane;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rs * (dzv[i][k] - s1 - xv[i][k] * s2);
        if (dxres != nullptr) {
          float r[8];
          V8<R>::load(dxres + base + v * 8, r);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += r[k];
        }
        if constexpr (!kDrop) {
          V8<X>::store(dx0 + base + v * 8, o);
          if (dx1 != nullptr) V8<R>::store(dx1 + base + v * 8, o);
        } else {
          if (dx1 != nullptr) V8<R>::store(dx1 + base + v * 8, o);
          const uint32_t rw = drop_row_word(drop.base, static_cast<uint32_t>(row));
          const uint4 c0 = *reinterpret_cast<const uint4*>(s_cw + v * 8), c1 = *reinterpret_cast<const uint4*>(s_cw + v * 8 + 4);
          const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = drop_keep(rw, cw[k], drop.thr24) ? o[k] * drop.scale : 0.f;
          V8<X>::store(dx0 + base + v * 8, o);
        }
      }
    }
  }

  // CTA reduction of the per-lane column sums, warp by warp in a fixed order
  float* out = part + static_cast<int64_t>(blockIdx.x) * 2 * cols;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = i * 32 + lane;
      if (v < nvec) V8<float>::store(lnb_smem + warp * cols + v * 8, which == 0 ? dg[i] : db[i]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerCta; ++w) s += lnb_smem[w * cols + c];
      out[which * cols + c] = s;
    }
    __syncthreads();
  }
}

// Column sums of the per-CTA partial rows.  32 columns x 16 row lanes per CTA: every thread adds up a strided
// sixteenth of the partial rows (independent loads in flight), the sixteen lanes are combined through shared memory in
// a fixed order.  (A thread-per-column version walked ~300 dependent L2 round trips and took longer than the row pass.)
template <typename W>
__global__ void __launch_bounds__(512)
ln_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int cols, W* __restrict__ dgamma,
                       W* __restrict__ dbeta) {
  __shared__ float red[2][16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (c < cols) {
    for (int p = threadIdx.y; p < nparts; p += 16) {
      a += part[static_cast<int64_t>(p) * 2 * cols + c];
      b += part[static_cast<int64_t>(p) * 2 * cols + cols + c];
    }
  }
  red[0][threadIdx.y][threadIdx.x] = a;
  red[1][threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y < 2 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 16; ++y) s += red[threadIdx.y][y][threadIdx.x];
    W* out = threadIdx.y == 0 ? dgamma : dbeta;
    if constexpr (sizeof(W) == 4) out[c] = s;
    else out[c] = static_cast<W>(s);
  }
}

template <typename X, typename R, typename W, int NV, bool kDrop>
int launch_kernel(const void* dz, const void* dxres, const void* x, const void* gamma, const float* mu, const float* rs,
                  void* dx0, void* dx1, void* dgamma, void* dbeta, float* part, int64_t rows, int cols, float eps,
                  DropArgs drop, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas_needed = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  int64_t cap = static_cast<int64_t>(sms) * 2;
  if (cap > kMaxCtas) cap = kMaxCtas;
  const int grid = static_cast<int>(ctas_needed < cap ? ctas_needed : cap);
  const size_t smem = static_cast<size_t>(kWarpsPerCta + 1 + (kDrop ? 1 : 0)) * cols * sizeof(float);
  auto kern = ln_residual_bwd_kernel<X, R, W, NV, kDrop>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(BP_ERR_CUDA, "bp_ln_residual_bwd: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    }
  }
  kern<<<grid, kWarpsPerCta * 32, smem, st>>>(static_cast<const X*>(dz), static_cast<const R*>(dxres),
                                              static_cast<const R*>(x), static_cast<const W*>(gamma), mu, rs,
                                              static_cast<X*>(dx0), static_cast<R*>(dx1), part, rows, cols, eps, drop);
  if (int rc = check_launch("bp_ln_residual_bwd launch")) return rc;
  ln_bwd_finalize_kernel<W><<<(cols + 31) / 32, dim3(32, 16), 0, st>>>(part, grid, cols, static_cast<W*>(dgamma),
                                                                 static_cast<W*>(dbeta));
  return check_launch("bp_ln_residual_bwd (finalize) launch");
}

template <typename X, typename R, typename W, int NV>
int launch_nv(const void* dz, const void* dxres, const void* x, const void* gamma, const float* mu, const float* rs,
              void* dx0, void* dx1, void* dgamma, void* dbeta, float* part, int64_t rows, int cols, float eps,
              DropArgs drop, cudaStream_t st) {
  if (drop.thr24 != 0)
    return launch_kernel<X, R, W, NV, true>(dz, dxres, x, gamma, mu, rs, dx0, dx1, dgamma, dbeta, part, rows, cols, eps, drop, st);
  return launch_kernel<X, R, W, NV, false>(dz, dxres, x, gamma, mu, rs, dx0, dx1, dgamma, dbeta, part, rows, cols, eps, drop, st);
}

template <typename X, typename R, typename W>
int launch(const void* dz, const void* dxres, const void* x, const void* gamma, const float* mu, const float* rs,
           void* dx0, void* dx1, void* dgamma, void* dbeta, float* part, int64_t rows, int cols, float eps, DropArgs drop,
           cudaStream_t st) {
  const int nv = (cols / 8 + 31) / 32;
#define BP_LNB_CASE(N) \
  if (nv <= N) return launch_nv<X, R, W, N>(dz, dxres, x, gamma, mu, rs, dx0, dx1, dgamma, dbeta, part, rows, cols, eps, drop, st)
  BP_LNB_CASE(1);
  BP_LNB_CASE(2);
  BP_LNB_CASE(3);
  BP_LNB_CASE(4);
  BP_LNB_CASE(6);
  BP_LNB_CASE(8);
#undef BP_LNB_CASE
  return fail(BP_ERR_UNSUPPORTED, "bp_ln_residual_bwd: hidden size %d > 2048 is not supported", cols);
}

}  // namespace lnb
}  // namespace bp

extern "C" int64_t bp_ln_bwd_workspace_bytes(int32_t cols) {
  return cols > 0 ? static_cast<int64_t>(bp::lnb::kMaxCtas) * 2 * cols * 4 : 0;
}

namespace {
int ln_residual_bwd_impl(const void* dz, const void* dx_residual, const void* x, const void* gamma,
                         const float* mu, const float* rsigma, void* dx0,
                         void* dx1, void* dgamma, void* dbeta, void* workspace, int64_t workspace_bytes,
                         int64_t rows, int32_t cols, float epsilon, int32_t x0_dtype, int32_t residual_dtype,
                         int32_t weight_dtype, bp::lnb::DropArgs drop, void* stream) {
  using namespace bp;
  if (!dz || !x || !gamma || !dx0 || !dgamma || !dbeta || !workspace)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_bwd: null pointer argument");
  if ((mu == nullptr) != (rsigma == nullptr))
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_bwd: mu and rsigma must be given together (or both NULL)");
  if (rows <= 0 || cols <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_bwd: empty input");
  if (cols % 8 != 0)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_bwd: hidden size must be a multiple of 8 (got %d)", cols);
  if (workspace_bytes < bp_ln_bwd_workspace_bytes(cols))
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_bwd: workspace of %lld bytes needed (got %lld)",
                (long long)bp_ln_bwd_workspace_bytes(cols), (long long)workspace_bytes);
  const uintptr_t ptrs[] = {(uintptr_t)dz,  (uintptr_t)dx_residual, (uintptr_t)x,      (uintptr_t)gamma, (uintptr_t)dx0,
                            (uintptr_t)dx1, (uintptr_t)dgamma,      (uintptr_t)dbeta, (uintptr_t)workspace};
  for (uintptr_t a : ptrs)
    if (a % 16 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_ln_residual_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(workspace);
  const int key = x0_dtype * 100 + residual_dtype * 10 + weight_dtype;
#define BP_LNB_DISPATCH(XD, RD, WD, X, R, W) \
  if (key == XD * 100 + RD * 10 + WD)        \
  return lnb::launch<X, R, W>(dz, dx_residual, x, gamma, mu, rsigma, dx0, dx1, dgamma, dbeta, part, rows, cols, epsilon, drop, st)
  using bf = __nv_bfloat16;
  using hf = __half;
  BP_LNB_DISPATCH(BP_DTYPE_BF16, BP_DTYPE_F32, BP_DTYPE_BF16, bf, float, bf);
  BP_LNB_DISPATCH(BP_DTYPE_F16, BP_DTYPE_F32, BP_DTYPE_F16, hf, float, hf);
  BP_LNB_DISPATCH(BP_DTYPE_BF16, BP_DTYPE_BF16, BP_DTYPE_BF16, bf, bf, bf);
  BP_LNB_DISPATCH(BP_DTYPE_F16, BP_DTYPE_F16, BP_DTYPE_F16, hf, hf, hf);
  BP_LNB_DISPATCH(BP_DTYPE_F32, BP_DTYPE_F32, BP_DTYPE_F32, float, float, float);
  BP_LNB_DISPATCH(BP_DTYPE_BF16, BP_DTYPE_F32, BP_DTYPE_F32, bf, float, float);
  BP_LNB_DISPATCH(BP_DTYPE_F16, BP_DTYPE_F32, BP_DTYPE_F32, hf, float, float);
#undef BP_LNB_DISPATCH
  return fail(BP_ERR_UNSUPPORTED, "bp_ln_residual_bwd: dtype combination (x0=%d, residual=%d, weight=%d) not built",
              x0_dtype, residual_dtype, weight_dtype);
}
}  // namespace

extern "C" int bp_ln_residual_bwd(const void* dz, const void* dx_residual, const void* x, const void* gamma,
                                  const float* mu, const float* rsigma, void* dx0,
                                  void* dx1, void* dgamma, void* dbeta, void* workspace, int64_t workspace_bytes,
                                  int64_t rows, int32_t cols, float epsilon, int32_t x0_dtype, int32_t residual_dtype,
                                  int32_t weight_dtype, void* stream) {
  return ln_residual_bwd_impl(dz, dx_residual, x, gamma, mu, rsigma, dx0, dx1, dgamma, dbeta, workspace, workspace_bytes,
                              rows, cols, epsilon, x0_dtype, residual_dtype, weight_dtype, bp::lnb::DropArgs(), stream);
}

extern "C" int bp_ln_residual_bwd_dropout(const void* dz, const void* dx_residual, const void* x, const void* gamma,
                                          const float* mu, const float* rsigma, void* dx0,
                                          void* dx1, void* dgamma, void* dbeta, void* workspace, int64_t workspace_bytes,
                                          int64_t rows, int32_t cols, float epsilon, int32_t x0_dtype,
                                          int32_t residual_dtype, int32_t weight_dtype, float dropout_p, uint64_t seed,
                                          void* stream) {
  bp::lnb::DropArgs drop;
  if (int rc = bp::ln_drop_args(dropout_p, seed, rows, "bp_ln_residual_bwd_dropout", &drop.base, &drop.thr24, &drop.scale))
    return rc;
  return ln_residual_bwd_impl(dz, dx_residual, x, gamma, mu, rsigma, dx0, dx1, dgamma, dbeta, workspace, workspace_bytes,
                              rows, cols, epsilon, x0_dtype, residual_dtype, weight_dtype, drop, stream);
}
