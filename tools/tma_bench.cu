// Microbenchmark: sustained per-SM load throughput of cp.async.bulk.tensor boxes (rank 2..5) and plain bulk copies
// on B200, as used by conv_tma.cu.  One CTA per SM, one producer thread, a ring of `stages` 16 KB slots, the
// consumer only waits and re-arms.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_bench tools/tma_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6,%7}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mode 0: rank-5 box {64,32,4,1,1}; 1: rank-4 {64,32,4,1}; 2: rank-2 {64,128}; 3: bulk 16 KB; 4: rank-5 split into 2 boxes {64,32,2}
__global__ void __launch_bounds__(64, 1) bench(const __grid_constant__ CUtensorMap m5, const __grid_constant__ CUtensorMap m4,
                                                const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m5h,
                                                const uint8_t* base, int mode, int stages, int iters, int W, int H, int F,
                                                int64_t total_rows, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(sm + (size_t)stages * 16384);
  const uint32_t s0 = smem_u32(sm), b0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(b0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int tiles_w = W / 32, tiles_h = H / 4;
  // cheap incremental tile walk (no divisions in the loop): every CTA streams its own frames
  int tw = 0, th = 0, f = blockIdx.x % F;
  int row2 = (int)(((long long)blockIdx.x * 128 * 37) % (total_rows - 128));
  long long t0 = clock64();
  for (int it = 0; it < iters + stages; ++it) {
    const int s = it % stages;  // stages is small; cheap 32-bit op
    if (it >= stages) mbar_wait(b0 + 8 * s, ((it / stages) - 1) & 1);
    if (it < iters) {
      mbar_expect(b0 + 8 * s, 16384);
      const uint32_t dst = s0 + s * 16384;
      if (mode == 0) tma5(dst, &m5, b0 + 8 * s, 0, tw * 32, th * 4, f, 0);
      else if (mode == 1) tma4(dst, &m4, b0 + 8 * s, 0, tw * 32, th * 4, f);
      else if (mode == 2) tma2(dst, &m2, b0 + 8 * s, 0, row2);
      else if (mode == 3) bulk(dst, base + (size_t)row2 * 128, 16384, b0 + 8 * s);
      else { tma5(dst, &m5h, b0 + 8 * s, 0, tw * 32, th * 4, f, 0); tma5(dst + 8192, &m5h, b0 + 8 * s, 0, tw * 32, th * 4 + 2, f, 0); }
      if (++tw == tiles_w) { tw = 0; if (++th == tiles_h) { th = 0; f += gridDim.x; if (f >= F) f -= F; } }
      row2 += 128 * 148; if (row2 >= total_rows - 128) row2 -= (int)(total_rows - 128);
    }
  }
  cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int C = 64, W = 96, H = 56;
  int F = argc > 1 ? atoi(argv[1]) : 128;   // frames: 128 -> 88 MB (fits L2 126 MB), 1024 -> 704 MB (streams from HBM)
  int promo = argc > 2 ? atoi(argv[2]) : 2;
  const int64_t rows = (int64_t)F * H * W;
  uint8_t* d;
  CK(cudaMalloc(&d, rows * C * 2));
  CK(cudaMemset(d, 1, rows * C * 2));
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  Enc enc = (Enc)sym;
  CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUtensorMap m5, m4, m2, m5h;
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F, 1};
    cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)F * H * W * C * 2};
    cuuint32_t box[5] = {64, 32, 4, 1, 1}, boxh[5] = {64, 32, 2, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&m5, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r1 = enc(&m5h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, boxh, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t d2[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t s2[1] = {(cuuint64_t)C * 2};
    cuuint32_t b2[2] = {64, 128};
    CUresult r3 = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, d2, s2, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r || r1 || r2 || r3) { printf("encode failed %d %d %d %d\n", r, r1, r2, r3); return 1; }
  }
  const char* names[5] = {"rank5 {64,32,4}", "rank4 {64,32,4}", "rank2 {64,128}", "bulk 16KB", "rank5 2x{64,32,2}"};
  printf("tensor %d frames of %dx%dx%d bf16 = %.0f MB, L2 promotion %d\n", F, H, W, C, rows * C * 2 / 1e6, promo);
  for (int mode = 0; mode < 5; ++mode) {
    for (int stages : {2, 4, 8, 12}) {
      const int iters = 2000;
      size_t smem = 1024 + (size_t)stages * 16384 + 256;
      CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int rep = 0; rep < 2; ++rep) {
        bench<<<148, 64, smem>>>(m5, m4, m2, m5h, d, mode, stages, iters, W, H, F, rows, cyc);
        CK(cudaDeviceSynchronize());
      }
      long long h[148];
      CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      printf("%-20s stages %2d : %7.1f clk per 16KB box  = %6.1f B/clk/SM\n", names[mode], stages, avg / iters, 16384.0 * iters / avg);
    }
  }
  return 0;
}
