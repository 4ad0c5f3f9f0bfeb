// Hardware-contract probe for the descriptor conventions bp_common.cuh relies on.  Stand-alone binary
// (built by __graft_entry__.build(), run by tests/test_umma_probe.py on the GPU box).
//   case A: D[128x128] = A[128x64] * B[128x64]^T       both operands K-major, SWIZZLE_128B, TMA-loaded
//   case B: D[128x64]  = P[128x128] * V[128x64]        P written by threads with sw128_offset(),
//                                                      V TMA-loaded [key rows x 64] consumed MN-major
//   case C: like B with N = 128 (two 64-column V panels, LBO = panel stride)
//   case D: D[128x64]  = P[128x128] * V[128x64]        P packed bf16x2 into TMEM by tcgen05.st (32x32b: register c
//                                                      of lane r = elements (r, 2c) | (r, 2c+1) << 16), A operand read
//                                                      from TMEM (tcgen05.mma [d], [a_tmem], b_desc)
// Prints "caseX max_err <e>" lines and exits non-zero on mismatch.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../backpacks_flash_attn_b200/csrc/bp_common.cuh"
#include "../../backpacks_flash_attn_b200/csrc/bp_host.h"

using namespace bp;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

struct Smem {
  uint64_t full, done;
  uint32_t tmem;
};

// mode 0: case A (N=128, K=64).  mode 1/2: case B/C with NV = 64 / 128 value columns.
template <int MODE>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __nv_bfloat16* __restrict__ Pg, float* __restrict__ D) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 32 KB reserved
  uint8_t* sB = smem + 32768;    // 32 KB reserved
  Smem& sm = *reinterpret_cast<Smem*>(smem + 65536);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
  constexpr int N = MODE == 0 ? 128 : (MODE == 2 ? 128 : 64);
  if (threadIdx.x == 0) {
    mbar_init(&sm.full, 1);
    mbar_init(&sm.done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&sm.tmem, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem;

  if (MODE == 3) {
    // P row r -> 64 packed words -> TMEM columns [64, 128) of lane r
    const uint32_t* src = reinterpret_cast<const uint32_t*>(Pg + r * 128);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 64;
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      for (int i = 0; i < 32; ++i) v[i] = src[c * 32 + i];
      tmem_st32(taddr + c * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
  } else if (MODE != 0) {
    // P tile: thread r owns row r, 128 bf16 = 16 chunks of 16 B, two 64-column panels
    const uint4* src = reinterpret_cast<const uint4*>(Pg + r * 128);
    for (int c = 0; c < 16; ++c)
      *reinterpret_cast<uint4*>(sA + (c >> 3) * 16384 + sw128_offset(r, c & 7)) = src[c];
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    if (MODE == 0) {
      mbar_arrive_expect_tx(&sm.full, 2 * 16384);
      tma_load_2d(sA, &tmA, &sm.full, 0, 0);
      tma_load_2d(sB, &tmB, &sm.full, 0, 0);
    } else {
      mbar_arrive_expect_tx(&sm.full, (N / 64) * 16384);
      for (int pn = 0; pn < N / 64; ++pn) tma_load_2d(sB + pn * 16384, &tmB, &sm.full, pn * 64, 0);
    }
    mbar_wait(&sm.full, 0);
    tc_fence_after();
    if (MODE == 0) {
      constexpr uint32_t idesc = make_idesc(true, 128, 128, false, false);
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem, make_smem_desc_sw128(smem_u32(sA) + kk * 32, 16, 1024),
                make_smem_desc_sw128(smem_u32(sB) + kk * 32, 16, 1024), idesc, kk > 0);
    } else if (MODE == 3) {
      constexpr uint32_t idesc = make_idesc(true, 128, N, false, true);
      for (int kk = 0; kk < 8; ++kk)
        umma_ts(tmem, tmem + 64 + kk * 8, make_smem_desc_sw128(smem_u32(sB) + kk * 2048, 16384, 1024), idesc, kk > 0);
    } else {
      constexpr uint32_t idesc = make_idesc(true, 128, N, false, true);
      for (int kk = 0; kk < 8; ++kk)
        umma_ss(tmem, make_smem_desc_sw128(smem_u32(sA) + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                make_smem_desc_sw128(smem_u32(sB) + kk * 2048, 16384, 1024), idesc, kk > 0);
    }
    umma_commit(&sm.done);
  }
  __syncwarp();
  mbar_wait(&sm.done, 0);
  tc_fence_after();
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < N / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[r * N + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
  (void)lane;
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  if (bp_check_device() != 0) {
    printf("device check failed: %s\n", bp_last_error());
    return 3;
  }
  srand(1);
  auto rnd = [] { return bf((rand() % 2001 - 1000) / 500.0f); };
  int bad = 0;
  const int smem_bytes = 65536 + 256 + 1024;

  // ---------------- case A ----------------
  {
    std::vector<float> A(128 * 64), B(128 * 64);
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    std::vector<__nv_bfloat16> Ah(A.size()), Bh(B.size());
    for (size_t i = 0; i < A.size(); ++i) Ah[i] = __float2bfloat16(A[i]), Bh[i] = __float2bfloat16(B[i]);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, Ah.size() * 2));
    CK(cudaMalloc(&dB, Bh.size() * 2));
    CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tA, tB;
    uint64_t dims[2] = {64, 128}, str[1] = {128};
    uint32_t box[2] = {64, 128};
    if (encode_tensor_map(&tA, BP_DTYPE_BF16, 2, dA, dims, str, box, true) ||
        encode_tensor_map(&tB, BP_DTYPE_BF16, 2, dB, dims, str, box, true)) {
      printf("encode failed: %s\n", bp_last_error());
      return 4;
    }
    CK(cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    probe_kernel<0><<<1, 128, smem_bytes>>>(tA, tB, nullptr, dD);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * 128);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)A[i * 64 + k] * B[j * 64 + k];
        err = fmax(err, fabs(ref - D[i * 128 + j]));
      }
    printf("caseA max_err %.3e\n", err);
    if (!(err < 1e-3)) bad |= 1;
  }
  // ---------------- case B / C ----------------
  for (int mode = 1; mode <= 3; ++mode) {
    const int N = mode == 2 ? 128 : 64;
    std::vector<float> P(128 * 128), V(128 * N);
    for (auto& x : P) x = rnd();
    for (auto& x : V) x = rnd();
    std::vector<__nv_bfloat16> Ph(P.size()), Vh(V.size());
    for (size_t i = 0; i < P.size(); ++i) Ph[i] = __float2bfloat16(P[i]);
    for (size_t i = 0; i < V.size(); ++i) Vh[i] = __float2bfloat16(V[i]);
    __nv_bfloat16 *dP, *dV;
    float* dD;
    CK(cudaMalloc(&dP, Ph.size() * 2));
    CK(cudaMalloc(&dV, Vh.size() * 2));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dP, Ph.data(), Ph.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dV, Vh.data(), Vh.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tV;
    uint64_t dims[2] = {(uint64_t)N, 128}, str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, 128};
    if (encode_tensor_map(&tV, BP_DTYPE_BF16, 2, dV, dims, str, box, true)) {
      printf("encode failed: %s\n", bp_last_error());
      return 4;
    }
    if (mode == 1) {
      CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      probe_kernel<1><<<1, 128, smem_bytes>>>(tV, tV, dP, dD);
    } else if (mode == 2) {
      CK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      probe_kernel<2><<<1, 128, smem_bytes>>>(tV, tV, dP, dD);
    } else {
      CK(cudaFuncSetAttribute(probe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      probe_kernel<3><<<1, 128, smem_bytes>>>(tV, tV, dP, dD);
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < N; ++j) {
        double ref = 0;
        for (int k = 0; k < 128; ++k) ref += (double)P[i * 128 + k] * V[k * N + j];
        err = fmax(err, fabs(ref - D[i * N + j]));
      }
    printf("case%c max_err %.3e\n", "ABCD"[mode], err);
    if (!(err < 2e-3)) bad |= (1 << mode);
  }
  printf(bad ? "PROBE FAILED mask=%d\n" : "PROBE OK %d\n", bad);
  return bad ? 1 : 0;
}
