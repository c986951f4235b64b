// Experiment: can a tcgen05 K-major SWIZZLE_128B shared-memory descriptor start at an arbitrary 128-byte row of a
// TMA-written tile (not 1024-byte aligned), with an arbitrary stride between 8-row groups (SBO)?  If it can, the
// 9 taps of a 3x3 convolution can all read ONE halo tile in shared memory through shifted descriptors.
//
// For every (row offset, SBO, base-offset field) the kernel computes D[128,16] = A_view x B^T and the host works out,
// per output row m, WHICH source row r and WHICH 16-byte-chunk XOR x the hardware actually read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Iinclude -Ivinet_b200/csrc -o build/umma_shift_test tools/umma_shift_test.cu
#include <cuda.h>

#include <vector>

#include "tc_ptx.cuh"

namespace vinet {
void set_error(const char*, ...) {}
std::atomic<long long> g_launches{0};
}  // namespace vinet
using namespace vinet;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int RA = 256;  // rows of the A tile in shared memory (128 B each)
constexpr int NB = 16;

__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4}], [%2];" ::"r"(dst),
               "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

// mode 0: K-major A (rows = M).  mode 1: MN-major A (the 64 "K" columns of the tile are M, rows are the reduction dim)
__global__ void __launch_bounds__(128, 1) shift_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                                                        int off_rows, uint32_t sbo, uint32_t bo, int mode, uint32_t idesc, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;
  uint8_t* sB = sm + RA * 128;
  uint64_t* bars = (uint64_t*)(sB + 4096);
  uint32_t* slot = (uint32_t*)(bars + 2);
  const uint32_t bar0 = smem_u32(bars), bar1 = bar0 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(slot), 32);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *(volatile uint32_t*)slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar0, RA * 128 + NB * 128);
    tma2(smem_u32(sA), &mA, bar0, 0, 0);
    tma2(smem_u32(sB), &mB, bar0, 0, 0);
    mbar_wait(bar0, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + (uint32_t)off_rows * 128u;
    const uint32_t b0 = smem_u32(sB);
    if (mode == 0) {
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = (uint64_t)(((a0 + kk * 32) & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
                            ((uint64_t)(bo & 7u) << 49) | (2ull << 61);
        umma_bf16(tmem, ad, desc_kmajor_sw128(b0 + kk * 32, 0), idesc, kk != 0);
      }
    } else {
      // A MN-major: M = 64 channels of the tile (only 64 of the 128 M rows carry data; LBO points at the same block again),
      // K = 64 reduction rows starting at off_rows, 8-row atoms `sbo` bytes apart; B K-major as before but over rows.
      // (only used to learn whether shifted MN-major descriptors behave; D rows 64..127 duplicate rows 0..63)
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a = a0 + (uint32_t)kk * 2u * sbo;
        const uint64_t ad = (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)0 << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
                            ((uint64_t)(bo & 7u) << 49) | (2ull << 61);
        umma_bf16(tmem, ad, desc_kmajor_sw128(b0 + kk * 32, 0), idesc, kk != 0);
      }
    }
    umma_commit(bar1);
  }
  __syncthreads();
  mbar_wait(bar1, 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
  for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 16 + e] = __uint_as_float(r[e]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make2d(EncodeTiledFn enc, CUtensorMap* m, void* p, int rows) {
  const cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
  cudaDriverEntryPointQueryResult q;
  void* sym = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  std::vector<float> A(RA * 64), B(NB * 64);
  std::vector<__nv_bfloat16> Ab(RA * 64), Bb(NB * 64);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { A[i] = (float)(rand() % 7 - 3); Ab[i] = __float2bfloat16(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { B[i] = (float)(rand() % 7 - 3); Bb[i] = __float2bfloat16(B[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, Ab.size() * 2));
  CK(cudaMalloc(&dB, Bb.size() * 2));
  CK(cudaMalloc(&dO, 128 * 16 * 4));
  CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap mA, mB;
  make2d(enc, &mA, dA, RA);
  make2d(enc, &mB, dB, NB);
  const size_t smem = 1024 + RA * 128 + 4096 + 64 + 16384;
  CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // K-major hypotheses: D[m][n] = sum_k A[r][chunkxor(k, x)] * B[n][k]
  auto hyp = [&](int r, int x, int n) {
    float s = 0;
    for (int k = 0; k < 64; ++k) {
      const int ks = (((k >> 3) ^ x) << 3) | (k & 7);
      s += A[r * 64 + ks] * B[n * 64 + k];
    }
    return s;
  };
  const int offs[] = {0, 8, 1, 2, 3, 5, 9, 19};
  const uint32_t sbos[] = {1024, 2048, 1280, 2304};
  std::vector<float> D(128 * 16);
  for (uint32_t sbo : sbos) {
    for (int off : offs) {
      for (int bomode = 0; bomode < 2; ++bomode) {
        const uint32_t bo = bomode ? (uint32_t)(off & 7) : 0u;
        if (bomode && bo == 0) continue;
        shift_kernel<<<1, 128, smem>>>(mA, mB, off, sbo, bo, 0, make_idesc(128, NB, 0, 0), dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("off %d sbo %u bo %u: CUDA error %s\n", off, sbo, bo, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(D.data(), dO, D.size() * 4, cudaMemcpyDeviceToHost));
        int ok = 0, found = 0;
        char detail[512];
        int dl = 0;
        detail[0] = 0;
        for (int m = 0; m < 128; ++m) {
          const int want = off + (m >> 3) * (int)(sbo / 128) + (m & 7);
          bool match_want = want < RA;
          if (match_want)
            for (int n = 0; n < 16; ++n)
              if (hyp(want, 0, n) != D[m * 16 + n]) { match_want = false; break; }
          if (match_want) { ++ok; ++found; continue; }
          // search what was read instead
          int fr = -1, fx = -1;
          for (int r = 0; r < RA && fr < 0; ++r)
            for (int x = 0; x < 8 && fr < 0; ++x) {
              bool eq = true;
              for (int n = 0; n < 16; ++n)
                if (hyp(r, x, n) != D[m * 16 + n]) { eq = false; break; }
              if (eq) { fr = r; fx = x; }
            }
          if (fr >= 0) ++found;
          if (dl < 400 && m < 20) dl += snprintf(detail + dl, sizeof(detail) - dl, " m%d:want%d got r%d^%d", m, want, fr, fx);
        }
        printf("K-major off %2d sbo %4u base_offset %u: %3d/128 rows as wanted, %3d identified%s\n", off, sbo, bo, ok, found, detail);
      }
    }
  }
  // MN-major: D[c][n] = sum_{j<64} A[off + (j/8)*(sbo/128) + j%8][c] * B[n][j]
  for (uint32_t sbo : sbos) {
    for (int off : offs) {
      for (int bomode = 0; bomode < 2; ++bomode) {
        const uint32_t bo = bomode ? (uint32_t)(off & 7) : 0u;
        if (bomode && bo == 0) continue;
        shift_kernel<<<1, 128, smem>>>(mA, mB, off, sbo, bo, 1, make_idesc(128, NB, 1, 0), dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("MN off %d sbo %u bo %u: CUDA error %s\n", off, sbo, bo, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(D.data(), dO, D.size() * 4, cudaMemcpyDeviceToHost));
        int ok = 0;
        for (int c = 0; c < 64; ++c) {
          bool eq = true;
          for (int n = 0; n < 16 && eq; ++n) {
            float s = 0;
            for (int j = 0; j < 64; ++j) {
              const int r = off + (j >> 3) * (int)(sbo / 128) + (j & 7);
              if (r < RA) s += A[r * 64 + c] * B[n * 64 + j];
            }
            if (s != D[c * 16 + n]) eq = false;
          }
          ok += eq;
        }
        printf("MN-major off %2d sbo %4u base_offset %u: %2d/64 rows as wanted\n", off, sbo, bo, ok);
      }
    }
  }
  return 0;
}
