// Timing probe for SEQUENCES of tcgen05.mma (one CTA per SM, one issuing thread, operands resident, data = zeros):
// how many SM cycles does one repetition of a small "program" of MMAs take?  Used to pin down what bounds the
// attention and sense-mix kernels when the MMAs are narrow (N = 64 / 128) or alternate between accumulators:
//   * cost of one MMA (M = 128, K = 16) as a function of N, for A from shared memory (SS) and from tensor memory (TS);
//   * the attention kernel's per-block program (4 x SS N=BN into S, BN/16 x TS N=DP into O) for one and two tiles;
//   * the sense-mix step (8 x TS N=256/128 into O, 3 x SS N=64 into an S buffer) with the S buffer aliasing the P
//     operand of in-flight PV products (write-after-read on TMEM) and without.
//     nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 umma_pattern_probe.cu ../../backpacks_flash_attn_b200/csrc/bp_host.cu -o umma_pattern_probe.bin
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../backpacks_flash_attn_b200/csrc/bp_common.cuh"

using namespace bp;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

namespace bp { uint64_t* g_trace = nullptr; }

// One MMA with compile-time shape (descriptors fold to immediates + one add, as in the production kernels).
#ifndef WARP_ISSUE
#define WARP_ISSUE 1   // 1: whole warp walks the loop, MMA predicated on an elected lane; 0: divergent `if (thread 0)`
#endif
template <int TS, int N, int M, int MN>
__device__ __forceinline__ void mma1(uint32_t tm, uint32_t sA, uint32_t sB, int d_col, int a_col, int kk, uint32_t acc) {
  constexpr uint32_t idesc = make_idesc(true, M, N, false, MN != 0);
  const uint64_t b = MN ? make_smem_desc_sw128(sB + kk * 2048, 8192, 1024) : make_smem_desc_sw128(sB + kk * 32, 16, 1024);
#if WARP_ISSUE
  if constexpr (TS)
    umma_ts_w(tm + d_col, tm + a_col + kk * 8, b, idesc, acc);
  else
    umma_ss_w(tm + d_col, make_smem_desc_sw128(sA + kk * 32, 16, 1024), b, idesc, acc);
#else
  if constexpr (TS)
    umma_ts(tm + d_col, tm + a_col + kk * 8, b, idesc, acc);
  else
    umma_ss(tm + d_col, make_smem_desc_sw128(sA + kk * 32, 16, 1024), b, idesc, acc);
#endif
}

// programs (one repetition each); all loops unroll
template <int PROG>
__device__ __forceinline__ void program(uint32_t tm, uint32_t sA, uint32_t sB) {
  if constexpr (PROG < 100) {
    // PROG = TS*32 + MN*16 + m64*8 + log2(N/32): four K-steps of one shape into one accumulator
    constexpr int TS = (PROG >> 5) & 1, MN = (PROG >> 4) & 1, M = ((PROG >> 3) & 1) ? 64 : 128, N = 32 << (PROG & 7);
#pragma unroll
    for (int k = 0; k < 4; ++k) mma1<TS, N, M, MN>(tm, sA, sB, 0, 448, k, 1u);
  } else if constexpr (PROG == 100) {          // TS N=64, alternating accumulators every MMA
#pragma unroll
    for (int k = 0; k < 8; ++k) mma1<1, 64, 128, 1>(tm, sA, sB, (k & 1) * 64, 448, k & 3, 1u);
  } else if constexpr (PROG == 101 || PROG == 102) {   // attention d64: 1 or 2 tile-blocks
#pragma unroll
    for (int t = 0; t < (PROG == 101 ? 1 : 2); ++t) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma1<0, 128, 128, 0>(tm, sA, sB, t * 128, 0, k, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) mma1<1, 64, 128, 1>(tm, sA, sB, 256 + t * 64, 384 + t * 64 + (k >> 2) * 32, k & 3, 1u);
    }
  } else if constexpr (PROG == 103) {          // attention d128, BN = 64, two tile-blocks
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int k = 0; k < 8; ++k) mma1<0, 64, 128, 0>(tm, sA, sB, t * 64, 0, k & 3, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma1<1, 128, 128, 1>(tm, sA, sB, 128 + t * 128, 384 + t * 32, k, 1u);
    }
  } else if constexpr (PROG == 104) {          // transposed PV for both tiles: O^T (64 x 256) += V^T (M=64) P^T (N=256)
#pragma unroll
    for (int k = 0; k < 8; ++k) mma1<0, 256, 64, 0>(tm, sA, sB, 256, 0, k & 3, 1u);
  } else if constexpr (PROG == 105 || PROG == 106) {   // sense-mix, two steps: aliasing (105) / separate buffers (106)
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const int pcol = PROG == 105 ? 384 + st * 64 : 448 + st * 32;
      const int scol = PROG == 105 ? 384 + st * 64 : 384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mma1<1, 256, 128, 1>(tm, sA, sB, 0, pcol, k, 1u);
        mma1<1, 128, 128, 1>(tm, sA, sB, 256, pcol, k, 1u);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) mma1<0, 64, 128, 0>(tm, sA, sB, scol, 0, k, k > 0);
    }
  } else if constexpr (PROG == 107) {          // sense-mix PV only
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mma1<1, 256, 128, 1>(tm, sA, sB, 0, 448, k, 1u);
      mma1<1, 128, 128, 1>(tm, sA, sB, 256, 448, k, 1u);
    }
  } else if constexpr (PROG == 108) {          // attention d64 with BOTH tiles' S as one N=256 product?  (not expressible: different A)
#pragma unroll
    for (int k = 0; k < 4; ++k) mma1<0, 256, 128, 0>(tm, sA, sB, 0, 0, k, k > 0);
  } else if constexpr (PROG == 109) {          // LSE pass block: S for two tiles, N=128 each
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int k = 0; k < 3; ++k) mma1<0, 128, 128, 0>(tm, sA, sB, t * 128, 0, k, k > 0);
  }
}

struct Bars {
  uint64_t done;
  uint32_t tmem;
};

template <int PROG>
__global__ void __launch_bounds__(128, 1) pattern_kernel(int reps, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // A: 128 rows x 64 (K) bf16 (16 KB); B: 256 rows x 64 (K) K-major or 64 keys x 256 columns MN-major (32 KB)
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 16384);
  Bars& bars = *reinterpret_cast<Bars*>(smem + 16384 + 32768);
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bars.done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&bars.tmem, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = bars.tmem;
#if WARP_ISSUE
  if (threadIdx.x < 32) {
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) program<PROG>(tm, sA, sB);
    umma_commit_w(smem_u32(&bars.done));
    mbar_wait(&bars.done, 0);
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  }
#else
  if (threadIdx.x == 0) {
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) program<PROG>(tm, sA, sB);
    umma_commit(&bars.done);
    mbar_wait(&bars.done, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int PROG>
static void run(const std::string& name, double ideal, int n_mma) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int smem = 16384 + 32768 + 64 + 1024, reps = 2000;
  CK(cudaFuncSetAttribute(pattern_kernel<PROG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long* d;
  CK(cudaMalloc(&d, sms * sizeof(unsigned long long)));
  for (int rep = 0; rep < 2; ++rep) {
    pattern_kernel<PROG><<<sms, 128, smem>>>(reps, d);
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> h(sms);
  CK(cudaMemcpy(h.data(), d, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long mn = ~0ull, mx = 0;
  for (auto c : h) mn = c < mn ? c : mn, mx = c > mx ? c : mx;
  printf("%-74s cycles/rep min %8.1f max %8.1f  ideal %7.1f  eff %.2f  (%d MMAs, %.1f cyc/MMA)\n", name.c_str(),
         double(mn) / reps, double(mx) / reps, ideal, ideal / (double(mn) / reps), n_mma, double(mn) / reps / n_mma);
  CK(cudaFree(d));
}

template <int TS, int MN, int M64, int LN>
static void run_shape() {
  constexpr int N = 32 << LN;
  char nm[128];
  snprintf(nm, sizeof nm, "%s M=%-3d N=%-3d B %s-major, one accumulator, 4 K-steps", TS ? "TS" : "SS", M64 ? 64 : 128, N, MN ? "MN" : "K");
  run<TS * 32 + MN * 16 + M64 * 8 + LN>(nm, 4 * 128.0 * N / 256.0, 4);
}

int main() {
  printf("issue style: %s\n", WARP_ISSUE ? "whole warp, elected lane inside the asm" : "divergent if (thread 0)");
  // ---- 1. one MMA shape repeated (4 K-steps per repetition), same accumulator.  ideal = 128 * N / 256 cycles per MMA ----
  run_shape<0, 0, 0, 0>(); run_shape<0, 0, 0, 1>(); run_shape<0, 0, 0, 2>(); run_shape<0, 0, 0, 3>();
  run_shape<0, 1, 0, 1>(); run_shape<0, 1, 0, 2>(); run_shape<0, 1, 0, 3>();
  run_shape<1, 1, 0, 0>(); run_shape<1, 1, 0, 1>(); run_shape<1, 1, 0, 2>(); run_shape<1, 1, 0, 3>();
  run_shape<0, 0, 1, 1>(); run_shape<0, 0, 1, 2>(); run_shape<0, 0, 1, 3>();
  // ---- 2. programs ----
  run<100>("TS N=64, alternating between two accumulators every MMA (8 MMAs)", 8 * 32.0, 8);
  run<101>("attention d64: one tile-block  (4 SS N=128 -> S, 8 TS N=64 -> O)", 4 * 64.0 + 8 * 32.0, 12);
  run<102>("attention d64: two tile-blocks (both query tiles)", 2 * (4 * 64.0 + 8 * 32.0), 24);
  run<103>("attention d128 (BN=64): two tile-blocks (8 SS N=64 -> S, 4 TS N=128 -> O, each)", 2 * (8 * 32.0 + 4 * 64.0), 24);
  run<104>("attention d64: transposed PV for both tiles (8 SS M=64 N=256)", 8 * 128.0, 8);
  run<105>("sense-mix: 2 steps, S(n+2) aliases P(n) (r02a layout)", 2 * (4 * 192.0 + 3 * 32.0), 22);
  run<106>("sense-mix: 2 steps, separate S and P buffers (r01 layout)", 2 * (4 * 192.0 + 3 * 32.0), 22);
  run<107>("sense-mix: PV only (8 TS N=256/128)", 4 * 192.0, 8);
  run<109>("sense LSE pass: S for two tiles (2 x 3 SS N=128)", 6 * 64.0, 6);
  return 0;
}
