// In-place rotary embedding on the q and k slices of a packed qkv tensor, sm_100a.
//
// Replaces rotary_emb.apply_rotary as driven by ApplyRotaryEmbQKV_.forward
// (flash_attn/layers/rotary.py:81-105; csrc/rotary/rotary_cuda.cu:5-41): for the first rotary_dim features of
// every head, (x1, x2) = (x[:rd/2], x[rd/2:rd]) -> (x1 cos - x2 sin, x1 sin + x2 cos), fp32 math, rounded once.
// The reference launches two TensorIterator kernels (q, then k); here one launch covers both, each thread
// moving 16-byte vectors (HBM-bound element-wise work: coalescing and vector width are what matter).
#include "bp_common.cuh"
#include "bp_host.h"

namespace bp {
namespace rotary {

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
template <bool kBF16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack2<kBF16>(f[0], f[1]);
  u.y = pack2<kBF16>(f[2], f[3]);
  u.z = pack2<kBF16>(f[4], f[5]);
  u.w = pack2<kBF16>(f[6], f[7]);
  return u;
}

// VEC = elements per thread along the half-rotary dimension (8 -> 16-byte accesses, 1 -> scalar fallback)
template <bool kBF16, int VEC>
__global__ void __launch_bounds__(256)
rotary_qk_kernel(uint16_t* __restrict__ qkv, const uint16_t* __restrict__ cos_q, const uint16_t* __restrict__ sin_q,
                 const uint16_t* __restrict__ cos_k, const uint16_t* __restrict__ sin_k, int64_t total_items,
                 int seqlen, int nheads, int headdim, int half) {
  const int per_head = half / VEC;
  for (int64_t it = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; it < total_items;
       it += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(it % per_head) * VEC;
    int64_t rest = it / per_head;
    const int h = static_cast<int>(rest % nheads);
    rest /= nheads;
    const int which = static_cast<int>(rest % 2);  // 0 = q, 1 = k
    const int64_t tok = rest / 2;                  // b * seqlen + s
    const int s = static_cast<int>(tok % seqlen);
    uint16_t* x = qkv + ((tok * 3 + which) * nheads + h) * static_cast<int64_t>(headdim) + i;
    const uint16_t* c = (which ? cos_k : cos_q) + static_cast<int64_t>(s) * half + i;
    const uint16_t* sn = (which ? sin_k : sin_q) + static_cast<int64_t>(s) * half + i;
    if constexpr (VEC == 8) {
      float x1[8], x2[8], cf[8], sf[8], o1[8], o2[8];
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(x), x1);
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(x + half), x2);
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(c), cf);
      unpack8<kBF16>(*reinterpret_cast<const uint4*>(sn), sf);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o1[k] = x1[k] * cf[k] - x2[k] * sf[k];
        o2[k] = x1[k] * sf[k] + x2[k] * cf[k];
      }
      *reinterpret_cast<uint4*>(x) = pack8<kBF16>(o1);
      *reinterpret_cast<uint4*>(x + half) = pack8<kBF16>(o2);
    } else {
      auto ld = [](uint16_t v) {
        if constexpr (kBF16) return __uint_as_float(static_cast<uint32_t>(v) << 16);
        else return __half2float(__ushort_as_half(v));
      };
      auto st = [](float v) -> uint16_t {
        if constexpr (kBF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
        else return __half_as_ushort(__float2half_rn(v));
      };
      const float a = ld(x[0]), b = ld(x[half]), cc = ld(c[0]), ss = ld(sn[0]);
      x[0] = st(a * cc - b * ss);
      x[half] = st(a * ss + b * cc);
    }
  }
}

}  // namespace rotary
}  // namespace bp

extern "C" int bp_rotary_qk_inplace(void* qkv, const void* cos, const void* sin, const void* cos_k,
                                    const void* sin_k, int32_t batch, int32_t seqlen, int32_t nheads,
                                    int32_t headdim, int32_t rotary_dim, int32_t dtype, void* stream) {
  using namespace bp;
  if (!qkv || !cos || !sin) return fail(BP_ERR_INVALID_ARGUMENT, "bp_rotary_qk_inplace: null pointer argument");
  if ((cos_k == nullptr) != (sin_k == nullptr))
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_rotary_qk_inplace: cos_k and sin_k must be given together");
  if (dtype != BP_DTYPE_F16 && dtype != BP_DTYPE_BF16)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_rotary_qk_inplace: only fp16 and bf16 are supported");
  if (batch <= 0 || seqlen <= 0 || nheads <= 0 || headdim <= 0)
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_rotary_qk_inplace: empty input");
  if (rotary_dim <= 0 || rotary_dim % 2 != 0 || rotary_dim > headdim)   // rotary.py:91
    return fail(BP_ERR_INVALID_ARGUMENT, "bp_rotary_qk_inplace: rotary_dim must be even and <= headdim (got %d)", rotary_dim);
  const int half = rotary_dim / 2;
  if (!cos_k) cos_k = cos, sin_k = sin;
  const bool vec = (half % 8 == 0) && (headdim % 8 == 0) && ((uintptr_t)qkv % 16 == 0) && ((uintptr_t)cos % 16 == 0) &&
                   ((uintptr_t)sin % 16 == 0) && ((uintptr_t)cos_k % 16 == 0) && ((uintptr_t)sin_k % 16 == 0);
  const int64_t items = static_cast<int64_t>(batch) * seqlen * 2 * nheads * (vec ? half / 8 : half);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (items + 255) / 256;
  const int grid = static_cast<int>(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* x = static_cast<uint16_t*>(qkv);
  auto *c = static_cast<const uint16_t*>(cos), *s = static_cast<const uint16_t*>(sin);
  auto *ck = static_cast<const uint16_t*>(cos_k), *sk = static_cast<const uint16_t*>(sin_k);
  if (dtype == BP_DTYPE_BF16) {
    if (vec) rotary::rotary_qk_kernel<true, 8><<<grid, 256, 0, st>>>(x, c, s, ck, sk, items, seqlen, nheads, headdim, half);
    else rotary::rotary_qk_kernel<true, 1><<<grid, 256, 0, st>>>(x, c, s, ck, sk, items, seqlen, nheads, headdim, half);
  } else {
    if (vec) rotary::rotary_qk_kernel<false, 8><<<grid, 256, 0, st>>>(x, c, s, ck, sk, items, seqlen, nheads, headdim, half);
    else rotary::rotary_qk_kernel<false, 1><<<grid, 256, 0, st>>>(x, c, s, ck, sk, items, seqlen, nheads, headdim, half);
  }
  return check_launch("bp_rotary_qk_inplace launch");
}
