// Softmax cross-entropy forward / backward for sm_100a (training loss over the LM-head logits).
//
// Replaces xentropy_cuda_lib.forward / .backward (csrc/xentropy/interface.cpp:57-58, xentropy_kernel.cu:430-760) as
// driven by SoftmaxCrossEntropyLossFn (flash_attn/losses/cross_entropy.py:19-109):
//     forward : lse[r] = log sum_j exp(x[r, j]);  loss[r] = (1 - eps) (lse - x[r, y_r]) + eps (lse - sum_j x[r, j] / C)
//               (loss 0 where y_r == ignore_index)
//     backward: dx[r, j] = g[r] (exp(x[r, j] - lse[r]) - (1 - eps) [j == y_r] - eps / C), optionally IN PLACE over the
//               logits (the reference's inplace_backward: the 6.6 GB logits tensor of config 3 is not duplicated)
// Both are single passes over the logits, HBM-bound: one CTA per row, 16-byte loads when the rows are 16-byte aligned
// (vocabulary padded to a multiple of 8, backpack.py:285-288; a scalar path handles e.g. the un-padded 50257 of the
// reference's own test), online max / sum-exp per thread (9 exponentials per 8 elements), a fixed-order block
// reduction.  The evaluation-only form that never materialises logits is bp_lm_head_stats_fwd.
#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace xent {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename T>
struct alignas(16) Vec {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

// fold N values into the running (max, sum-exp, sum)
template <int N>
__device__ __forceinline__ void fold(const float (&x)[N], float& m, float& s, float& sum) {
  float mx = x[0];
#pragma unroll
  for (int i = 1; i < N; ++i) mx = fmaxf(mx, x[i]);
  const float mn = fmaxf(m, mx);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    acc += __expf(x[i] - mn);
    sum += x[i];
  }
  s = s * __expf(m - mn) + acc;   // m = -inf, s = 0 on the first call: exp(-inf) = 0
  m = mn;
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kThreads)
xentropy_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ losses,
                    float* __restrict__ lse_out, int vocab, int64_t row_stride, float smoothing, int64_t ignore_index,
                    float inv_total_classes) {
  __shared__ float red[3][kThreads / 32];
  const int64_t row = blockIdx.x;
  const T* x = logits + row * row_stride;
  float m = -INFINITY, s = 0.f, sum = 0.f;
  if constexpr (kVec) {
    constexpr int N = Vec<T>::N;
    const int nvec = vocab / N;
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
      const Vec<T> t = *reinterpret_cast<const Vec<T>*>(x + static_cast<int64_t>(v) * N);
      float f[N];
#pragma unroll
      for (int i = 0; i < N; ++i) f[i] = to_f<T>(t.v[i]);
      fold<N>(f, m, s, sum);
    }
  } else {
    for (int j = threadIdx.x; j < vocab; j += kThreads) {
      const float f[1] = {to_f<T>(x[j])};
      fold<1>(f, m, s, sum);
    }
  }
  // block reduction of (m, s, sum): warp shuffles, then the 8 warp results in a fixed order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (mn == -INFINITY) ? 0.f : s * __expf(m - mn) + s2 * __expf(m2 - mn);
    m = mn;
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = m;
    red[1][warp] = s;
    red[2][warp] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = red[0][0], S = red[1][0], SUM = red[2][0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) {
      const float mn = fmaxf(M, red[0][w]);
      S = (mn == -INFINITY) ? 0.f : S * __expf(M - mn) + red[1][w] * __expf(red[0][w] - mn);
      M = mn;
      SUM += red[2][w];
    }
    const float lse = M + __logf(S);
    lse_out[row] = lse;
    const int64_t y = labels[row];
    float loss = 0.f;
    if (y != ignore_index) {
      const float xy = (y >= 0 && y < vocab) ? to_f<T>(x[y]) : 0.f;
      loss = (1.f - smoothing) * (lse - xy) + smoothing * (lse - SUM * inv_total_classes);
    }
    losses[row] = loss;
  }
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kThreads)
xentropy_bwd_kernel(const float* __restrict__ grad_losses, const T* logits, const float* __restrict__ lse,
                    const int64_t* __restrict__ labels, T* grad_logits, int vocab, int64_t row_stride,
                    int64_t grad_row_stride, float smoothing, int64_t ignore_index, float inv_total_classes) {
  // logits / grad_logits may alias (in-place backward): every element is read and written by the same thread
  const int64_t row = blockIdx.x;
  const T* x = logits + row * row_stride;
  T* dx = grad_logits + row * grad_row_stride;
  const int64_t y = labels[row];
  const float g = (y == ignore_index) ? 0.f : grad_losses[row];
  const float l = lse[row];
  const float off = smoothing * inv_total_classes;
  const float hit = 1.f - smoothing;
  if constexpr (kVec) {
    constexpr int N = Vec<T>::N;
    const int nvec = vocab / N;
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
      Vec<T> t = *reinterpret_cast<const Vec<T>*>(x + static_cast<int64_t>(v) * N);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float d = __expf(to_f<T>(t.v[i]) - l) - off;
        if (static_cast<int64_t>(v) * N + i == y) d -= hit;
        t.v[i] = from_f<T>(g * d);
      }
      *reinterpret_cast<Vec<T>*>(dx + static_cast<int64_t>(v) * N) = t;
    }
  } else {
    for (int j = threadIdx.x; j < vocab; j += kThreads) {
      float d = __expf(to_f<T>(x[j]) - l) - off;
      if (j == y) d -= hit;
      dx[j] = from_f<T>(g * d);
    }
  }
}

static bool aligned16(const void* p, int64_t row_stride, int vocab, int elem) {
  return reinterpret_cast<uintptr_t>(p) % 16 == 0 && (row_stride * elem) % 16 == 0 && (vocab * elem) % 16 == 0;
}

template <typename T>
int launch_fwd(const void* logits, const int64_t* labels, float* losses, float* lse, int64_t rows, int vocab,
               int64_t row_stride, float smoothing, int64_t ignore_index, float inv_c, cudaStream_t st) {
  const bool vec = aligned16(logits, row_stride, vocab, sizeof(T));
  if (vec)
    xentropy_fwd_kernel<T, true><<<static_cast<unsigned>(rows), kThreads, 0, st>>>(
        static_cast<const T*>(logits), labels, losses, lse, vocab, row_stride, smoothing, ignore_index, inv_c);
  else
    xentropy_fwd_kernel<T, false><<<static_cast<unsigned>(rows), kThreads, 0, st>>>(
        static_cast<const T*>(logits), labels, losses, lse, vocab, row_stride, smoothing, ignore_index, inv_c);
  return check_launch("bp_xentropy_fwd launch");
}

template <typename T>
int launch_bwd(const float* grad_losses, const void* logits, const float* lse, const int64_t* labels, void* grad_logits,
               int64_t rows, int vocab, int64_t row_stride, int64_t grad_row_stride, float smoothing, int64_t ignore_index,
               float inv_c, cudaStream_t st) {
  const bool vec = aligned16(logits, row_stride, vocab, sizeof(T)) && aligned16(grad_logits, grad_row_stride, vocab, sizeof(T));
  if (vec)
    xentropy_bwd_kernel<T, true><<<static_cast<unsigned>(rows), kThreads, 0, st>>>(
        grad_losses, static_cast<const T*>(logits), lse, labels, static_cast<T*>(grad_logits), vocab, row_stride,
        grad_row_stride, smoothing, ignore_index, inv_c);
  else
    xentropy_bwd_kernel<T, false><<<static_cast<unsigned>(rows), kThreads, 0, st>>>(
        grad_losses, static_cast<const T*>(logits), lse, labels, static_cast<T*>(grad_logits), vocab, row_stride,
        grad_row_stride, smoothing, ignore_index, inv_c);
  return check_launch("bp_xentropy_bwd launch");
}

static int check_args(const char* fn, int64_t rows, int vocab, int64_t row_stride, float smoothing, int total_classes,
                      int dtype) {
  if (rows <= 0 || vocab <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "%s: empty input", fn);
  if (rows > 0x7fffffff) return fail(BP_ERR_INVALID_ARGUMENT, "%s: more than 2^31 - 1 rows", fn);
  if (row_stride < vocab) return fail(BP_ERR_INVALID_ARGUMENT, "%s: row stride %lld < vocab %d", fn, (long long)row_stride, vocab);
  if (!(smoothing >= 0.f) || smoothing >= 1.f) return fail(BP_ERR_INVALID_ARGUMENT, "%s: smoothing must be in [0, 1)", fn);
  if (total_classes != -1 && total_classes < vocab)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: total_classes %d < vocab %d", fn, total_classes, vocab);
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16 && dtype != BP_DTYPE_F32)
    return fail(BP_ERR_INVALID_ARGUMENT, "%s: unsupported dtype %d", fn, dtype);
  return BP_OK;
}

}  // namespace xent
}  // namespace bp

extern "C" int bp_xentropy_fwd(const void* logits, const int64_t* labels, float* losses, float* lse, int64_t rows,
                               int32_t vocab, int64_t row_stride, float smoothing, int64_t ignore_index,
                               int32_t total_classes, int32_t dtype, void* stream) {
  using namespace bp;
  if (!logits || !labels || !losses || !lse) return fail(BP_ERR_INVALID_ARGUMENT, "bp_xentropy_fwd: null pointer argument");
  if (int rc = xent::check_args("bp_xentropy_fwd", rows, vocab, row_stride, smoothing, total_classes, dtype)) return rc;
  const float inv_c = 1.f / static_cast<float>(total_classes == -1 ? vocab : total_classes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == BP_DTYPE_BF16)
    return xent::launch_fwd<__nv_bfloat16>(logits, labels, losses, lse, rows, vocab, row_stride, smoothing, ignore_index, inv_c, st);
  if (dtype == BP_DTYPE_F16)
    return xent::launch_fwd<__half>(logits, labels, losses, lse, rows, vocab, row_stride, smoothing, ignore_index, inv_c, st);
  return xent::launch_fwd<float>(logits, labels, losses, lse, rows, vocab, row_stride, smoothing, ignore_index, inv_c, st);
}

extern "C" int bp_xentropy_bwd(const float* grad_losses, const void* logits, const float* lse, const int64_t* labels,
                               void* grad_logits, int64_t rows, int32_t vocab, int64_t row_stride,
                               int64_t grad_row_stride, float smoothing, int64_t ignore_index, int32_t total_classes,
                               int32_t dtype, void* stream) {
  using namespace bp;
  if (!grad_losses || !logits || !lse || !labels || !grad_logits)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_xentropy_bwd: null pointer argument");
  if (int rc = xent::check_args("bp_xentropy_bwd", rows, vocab, row_stride, smoothing, total_classes, dtype)) return rc;
  if (grad_row_stride < vocab)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_xentropy_bwd: gradient row stride %lld < vocab %d", (long long)grad_row_stride, vocab);
  const float inv_c = 1.f / static_cast<float>(total_classes == -1 ? vocab : total_classes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == BP_DTYPE_BF16)
    return xent::launch_bwd<__nv_bfloat16>(grad_losses, logits, lse, labels, grad_logits, rows, vocab, row_stride,
                                           grad_row_stride, smoothing, ignore_index, inv_c, st);
  if (dtype == BP_DTYPE_F16)
    return xent::launch_bwd<__half>(grad_losses, logits, lse, labels, grad_logits, rows, vocab, row_stride, grad_row_stride,
                                    smoothing, ignore_index, inv_c, st);
  return xent::launch_bwd<float>(grad_losses, logits, lse, labels, grad_logits, rows, vocab, row_stride, grad_row_stride,
                                 smoothing, ignore_index, inv_c, st);
}
