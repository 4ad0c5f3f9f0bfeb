// Backward of the bias + activation epilogue of the fused dense layers, for sm_100a.
//
// The reference runs the backward GEMMs of FusedDense / FusedDenseGeluDense on cuBLASLt with fused epilogues
// (csrc/fused_dense_lib/fused_dense_cuda.cu:559-787: CUBLASLT_EPILOGUE_BGRADB for the bias gradient,
// CUBLASLT_EPILOGUE_DGELU_BGRAD for dgelu + bias gradient).  Here the GEMMs stay plain GEMMs and this ONE pass does
// what those epilogues do:
//     dpre  = dact * gelu_tanh'(pre)        (16-bit, feeds the fc1 dgrad / wgrad GEMMs)
//     dbias = sum_rows dpre                 (BP_ACT_GELU_TANH)      or      sum_rows dact   (BP_ACT_NONE, no dpre)
// HBM-bound: every element is read once (16-byte loads, 512 B contiguous per warp and row) and written once; the
// column sums are kept per thread across the rows of a CTA, reduced across the CTA through shared memory in a fixed
// order, written as fp32 partials and summed by a second tiny kernel: deterministic, no atomics.
#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace bab {

constexpr int kTX = 32, kTY = 8;          // 32 x 8 threads: 256 columns x 8 rows per pass
constexpr int kMaxParts = 256;            // row slabs (grid.y) = partial rows in the workspace

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

// d/dx of 0.5 x (1 + tanh(k (x + c x^3)))  (the tanh form the forward epilogue and F.gelu(approximate='tanh') use)
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  constexpr float k = 0.7978845608028654f, c = 0.044715f;
  const float x2 = x * x;
  float t;   // hardware tanh (MUFU, rel. error 2^-11: far below the 16-bit rounding of dpre; the forward epilogue uses
             // the same instruction).  tanhf() made this pass compute-bound (0.33 ms at 65536 x 3072).
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k * x * fmaf(c, x2, 1.f)));
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * fmaf(3.f * c, x2, 1.f);
}

template <bool kBF16, bool kGelu>
__global__ void __launch_bounds__(kTX * kTY)
bias_act_bwd_kernel(const uint16_t* __restrict__ dact, const uint16_t* __restrict__ pre, uint16_t* __restrict__ dpre,
                    float* __restrict__ part, int64_t m, int n, int64_t rows_per_cta) {
  __shared__ float red[kTY][kTX * 8 + 8];
  const int col = (blockIdx.x * kTX + threadIdx.x) * 8;
  const int64_t r_begin = blockIdx.y * rows_per_cta;
  const int64_t r_end = min(m, r_begin + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < n) {
    constexpr int U = 4;   // rows in flight per thread
    for (int64_t r = r_begin + threadIdx.y; r < r_end; r += kTY * U) {
      uint4 a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + u * kTY;
        if (rr < r_end) {
          a[u] = __ldg(reinterpret_cast<const uint4*>(dact + rr * n + col));
          if constexpr (kGelu) b[u] = __ldg(reinterpret_cast<const uint4*>(pre + rr * n + col));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + u * kTY;
        if (rr < r_end) {
          float g[8];
          unpack8<kBF16>(a[u], g);
          if constexpr (kGelu) {
            float x[8];
            unpack8<kBF16>(b[u], x);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] *= gelu_tanh_grad(x[k]);
            uint4 o;
            o.x = pack2<kBF16>(g[0], g[1]);
            o.y = pack2<kBF16>(g[2], g[3]);
            o.z = pack2<kBF16>(g[4], g[5]);
            o.w = pack2<kBF16>(g[6], g[7]);
            *reinterpret_cast<uint4*>(dpre + rr * n + col) = o;
            // the bias gradient sums what the GEMMs will see: the rounded values
            unpack8<kBF16>(o, g);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += g[k];
        }
      }
    }
  }
  if (part == nullptr) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.y][threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  const int t = threadIdx.y * kTX + threadIdx.x;   // 256 threads, 256 columns
  const int c = blockIdx.x * kTX * 8 + t;
  if (c < n) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < kTY; ++y) s += red[y][t];
    part[static_cast<int64_t>(blockIdx.y) * n + c] = s;
  }
}

// Column sums of the partial rows: 32 columns x 16 row lanes per CTA, combined through shared memory in a fixed order.
template <bool kBF16>
__global__ void __launch_bounds__(512)
bias_finalize_kernel(const float* __restrict__ part, int nparts, int n, uint16_t* __restrict__ dbias) {
  __shared__ float red[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < n)
    for (int p = threadIdx.y; p < nparts; p += 16) a += part[static_cast<int64_t>(p) * n + c];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < n) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 16; ++y) s += red[y][threadIdx.x];
    if constexpr (kBF16) {
      const __nv_bfloat16 v = __float2bfloat16_rn(s);
      dbias[c] = *reinterpret_cast<const uint16_t*>(&v);
    } else {
      const __half v = __float2half_rn(s);
      dbias[c] = *reinterpret_cast<const uint16_t*>(&v);
    }
  }
}

template <bool kBF16, bool kGelu>
int launch(const void* dact, const void* pre, void* dpre, void* dbias, float* part, int64_t m, int n, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int gx = (n + kTX * 8 - 1) / (kTX * 8);
  int gy = (4 * sms + gx - 1) / gx;   // ~4 CTAs per SM in total
  if (gy > kMaxParts) gy = kMaxParts;
  const int64_t max_gy = (m + kTY - 1) / kTY;
  if (gy > max_gy) gy = static_cast<int>(max_gy);
  const int64_t rows_per_cta = (m + gy - 1) / gy;
  gy = static_cast<int>((m + rows_per_cta - 1) / rows_per_cta);
  bias_act_bwd_kernel<kBF16, kGelu><<<dim3(gx, gy), dim3(kTX, kTY), 0, st>>>(
      static_cast<const uint16_t*>(dact), static_cast<const uint16_t*>(pre), static_cast<uint16_t*>(dpre),
      dbias ? part : nullptr, m, n, rows_per_cta);
  if (int rc = check_launch("bp_bias_act_bwd launch")) return rc;
  if (dbias) {
    bias_finalize_kernel<kBF16><<<(n + 31) / 32, dim3(32, 16), 0, st>>>(part, gy, n, static_cast<uint16_t*>(dbias));
    return check_launch("bp_bias_act_bwd (finalize) launch");
  }
  return BP_OK;
}

}  // namespace bab
}  // namespace bp

extern "C" int64_t bp_bias_act_bwd_workspace_bytes(int32_t n) {
  return n > 0 ? static_cast<int64_t>(bp::bab::kMaxParts) * n * 4 : 0;
}

extern "C" int bp_bias_act_bwd(const void* dact, const void* pre, void* dpre, void* dbias, void* workspace,
                               int64_t workspace_bytes, int64_t m, int32_t n, int32_t activation, int32_t dtype,
                               void* stream) {
  using namespace bp;
  if (!dact) return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: null pointer argument");
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: only fp16 and bf16 are supported (dtype=%d)", dtype);
  if (activation != BP_ACT_NONE && activation != BP_ACT_GELU_TANH)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: unknown activation %d", activation);
  const bool gelu = activation == BP_ACT_GELU_TANH;
  if (gelu && (!pre || !dpre))
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: the GELU backward needs `pre` and `dpre`");
  if (!gelu && !dbias) return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: nothing to compute (no activation, no dbias)");
  if (m <= 0 || n <= 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: empty input");
  if (n % 8 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: n must be a multiple of 8 (got %d)", n);
  if (dbias && (!workspace || workspace_bytes < bp_bias_act_bwd_workspace_bytes(n)))
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: workspace of %lld bytes needed (got %lld)",
                (long long)bp_bias_act_bwd_workspace_bytes(n), (long long)workspace_bytes);
  const uintptr_t ptrs[] = {(uintptr_t)dact, (uintptr_t)pre, (uintptr_t)dpre, (uintptr_t)workspace};
  for (uintptr_t a : ptrs)
    if (a % 16 != 0) return fail(BP_ERR_INVALID_ARGUMENT, "bp_bias_act_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(workspace);
  const bool bf16 = dtype == BP_DTYPE_BF16;
  if (gelu)
    return bf16 ? bab::launch<true, true>(dact, pre, dpre, dbias, part, m, n, st)
                : bab::launch<false, true>(dact, pre, dpre, dbias, part, m, n, st);
  return bf16 ? bab::launch<true, false>(dact, nullptr, nullptr, dbias, part, m, n, st)
              : bab::launch<false, false>(dact, nullptr, nullptr, dbias, part, m, n, st);
}
