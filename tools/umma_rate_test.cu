// Microbenchmark: tensor-pipe time per tcgen05.mma (M=128, K=16, bf16) as a function of N on B200, operands resident in
// shared memory (no loads in the loop).  One CTA per SM on every SM.  Variants: accumulate into one / alternate between 4
// TMEM accumulators; A descriptor with the canonical SBO (1024) or a halo-style SBO (1280) and shifted start.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Iinclude -Ivinet_b200/csrc -o build/umma_rate_test tools/umma_rate_test.cu
#include <cuda.h>

#include <vector>

#include "tc_ptx.cuh"

namespace vinet {
void set_error(const char*, ...) {}
std::atomic<long long> g_launches{0};
}  // namespace vinet
using namespace vinet;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
      "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int n, uint32_t idesc, int iters, int nacc, uint32_t sbo, int shift_rows,
                                                      long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;              // 48 KB of "activations"
  uint8_t* sB = sm + 48 * 1024;  // 32 KB of "weights" (N <= 256 rows of 128 B)
  uint64_t* bars = (uint64_t*)(sB + 32 * 1024);
  uint32_t* slot = (uint32_t*)(bars + 2);
  for (int i = threadIdx.x; i < 80 * 1024 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bars), 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *(volatile uint32_t*)slot;
  if (warp == 1) {
    const uint32_t hi = (1u << 14) | (2u << 29);
    const uint32_t a_hi = (sbo >> 4) | hi, b_hi = 64u | hi;
    const uint32_t a_lo0 = (((smem_u32(sA) + shift_rows * 128) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stride = (uint32_t)((n + 31) / 32 * 32);
    long long t0 = clock64();
    uint32_t acc_i = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t td = tmem + acc_i * stride;
      const uint32_t a_lo = a_lo0 + (uint32_t)(it & 7) * 8;   // a different tap window each iteration
      if (elect_one()) {
        umma_lohi(td, a_lo, a_hi, b_lo0, b_hi, idesc, 1u);
        umma_lohi(td, a_lo + 2, a_hi, b_lo0 + 2, b_hi, idesc, 1u);
        umma_lohi(td, a_lo + 4, a_hi, b_lo0 + 4, b_hi, idesc, 1u);
        umma_lohi(td, a_lo + 6, a_hi, b_lo0 + 6, b_hi, idesc, 1u);
      }
      if (++acc_i == (uint32_t)nacc) acc_i = 0;
    }
    if (elect_one()) umma_commit(smem_u32(bars));
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(bars), 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 16));
  const size_t smem = 1024 + 80 * 1024 + 64 + 40 * 1024;   // > half of the SM: one CTA per SM
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 4096;
  printf("%4s %4s %5s %5s | %10s %10s | %8s\n", "N", "nacc", "sbo", "shift", "issue clk", "total clk", "clk/MMA");
  for (int variant = 0; variant < 6; ++variant) {
    if (variant == 3) printf("MN-major A and B (wgrad): variant 3 canonical, 4 halo SBO/shift, 5 A K-major + B MN-major\n");
    for (int n : {32, 64, 96, 128, 192, 256}) {
      int nacc = variant == 1 ? std::min(4, 512 / ((n + 31) / 32 * 32)) : 1;
      uint32_t sbo = (variant == 2 || variant == 4) ? 1280 : 1024;
      int shift = (variant == 2 || variant == 4) ? 11 : 0;
      const uint32_t idesc = variant >= 3 ? make_idesc(128, n, variant == 5 ? 0 : 1, 1) : make_idesc(128, n, 0, 0);
      rate_kernel<<<148, 128, smem>>>(n, idesc, iters, nacc, sbo, shift, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
      printf("%4d %4d %5u %5d | %10lld %10lld | %8.1f\n", n, nacc, sbo, shift, h[0], h[1], (double)h[1] / (4.0 * iters));
    }
  }
  return 0;
}
