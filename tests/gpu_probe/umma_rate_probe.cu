// Throughput probe for tcgen05.mma variants (one CTA per SM, one issuing thread, operands resident):
// how many SM cycles does a "step" of the sense-mix PV pattern take when nothing else is in the way?
//   step = 4 K-steps (16 keys each) of  D[128 x N] += A[128 x 16] * B[16 x N]
// variants: A from shared memory (SS) or tensor memory (TS); B K-major or MN-major; N = 256, or 256 followed by 128
// into a second accumulator region (the 384-column split of sense_mix_kernel).
// Prints cycles per step (min over SMs of the mean over R steps) next to the ideal at 8192 dense bf16 FLOP/clk/SM.
//     nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 umma_rate_probe.cu ../../backpacks_flash_attn_b200/csrc/bp_host.cu -o umma_rate_probe.bin
#include <stdio.h>
#include <stdlib.h>

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

struct Bars {
  uint64_t done;
  uint32_t tmem;
};

// variant bits: 1 = A from TMEM, 2 = B MN-major, 4 = extra N=128 product per K-step
template <int V>
__global__ void __launch_bounds__(128, 1) rate_kernel(int steps, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // A: 128 rows x 64 (K) bf16, 128B-swizzled rows (16 KB); B: up to 384 columns x 64 keys (48 KB)
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 16384);
  Bars& bars = *reinterpret_cast<Bars*>(smem + 16384 + 49152);
  for (int i = threadIdx.x; i < (16384 + 49152) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
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
  constexpr bool kTS = V & 1, kMN = V & 2, kSplit = V & 4;
  constexpr uint32_t idesc1 = make_idesc(true, 128, 256, false, kMN);
  constexpr uint32_t idesc2 = make_idesc(true, 128, 128, false, kMN);
  if (threadIdx.x == 0) {
    const unsigned long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        // K-major B: rows = N (128 B of K each); MN-major B: 64-column panels of 64 keys x 128 B, LBO = panel stride
        const uint64_t b1 = kMN ? make_smem_desc_sw128(sB + kk * 2048, 8192, 1024) : make_smem_desc_sw128(sB + kk * 32, 16, 1024);
        const uint64_t b2 = kMN ? make_smem_desc_sw128(sB + 4 * 8192 + kk * 2048, 8192, 1024)
                                : make_smem_desc_sw128(sB + 32768 + kk * 32, 16, 1024);
        if constexpr (kTS) {
          umma_ts(tm, tm + 448 + kk * 8, b1, idesc1, 1u);
          if constexpr (kSplit) umma_ts(tm + 256, tm + 448 + kk * 8, b2, idesc2, 1u);
        } else {
          const uint64_t a = make_smem_desc_sw128(sA + kk * 32, 16, 1024);
          umma_ss(tm, a, b1, idesc1, 1u);
          if constexpr (kSplit) umma_ss(tm + 256, a, b2, idesc2, 1u);
        }
      }
    }
    umma_commit(&bars.done);
    mbar_wait(&bars.done, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int V>
static void run(const char* name, int ideal_per_step) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int smem = 16384 + 49152 + 64 + 1024, steps = 2000;
  CK(cudaFuncSetAttribute(rate_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long* d;
  CK(cudaMalloc(&d, sms * sizeof(unsigned long long)));
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<V><<<sms, 128, smem>>>(steps, d);
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> h(sms);
  CK(cudaMemcpy(h.data(), d, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long mn = ~0ull, mx = 0;
  for (auto c : h) mn = c < mn ? c : mn, mx = c > mx ? c : mx;
  printf("%-44s cycles/step min %7.1f max %7.1f   ideal %d   efficiency %.2f\n", name, double(mn) / steps, double(mx) / steps,
         ideal_per_step, ideal_per_step / (double(mn) / steps));
  CK(cudaFree(d));
}

int main() {
  run<0>("SS  N=256      B K-major", 512);
  run<2>("SS  N=256      B MN-major", 512);
  run<3>("TS  N=256      B MN-major", 512);
  run<6>("SS  N=256+128  B MN-major", 768);
  run<7>("TS  N=256+128  B MN-major (sense-mix PV)", 768);
  run<4>("SS  N=256+128  B K-major", 768);
  return 0;
}
