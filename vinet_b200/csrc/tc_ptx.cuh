// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers and UMMA descriptor builders shared by the
// tensor-core convolution kernels (conv_tc.cu: register-gather producers; conv_tma.cu: TMA producers).
#pragma once
#include "gather.cuh"

namespace vinet {

constexpr int TC_BM = VINET_TC_BLOCK_M;  // 128 GEMM rows per CTA (UMMA M)
constexpr int TC_BK = VINET_TC_BLOCK_K;  // 64 bf16 = one 128-byte swizzle row
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hung GPU.  The spin loop lives in
// a non-inlined function so that the common case (the barrier has already flipped) costs one try_wait in the caller.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifdef VINET_WAIT_SLEEP
    __nanosleep(VINET_WAIT_SLEEP);
#endif
    if (++spins > (1u << 26)) {
      printf("vinet_b200: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Column sums of 16 consecutive accumulator columns over the 32 rows (lanes) of a warp by a reduce-scatter butterfly (15 + 1
// shuffles instead of 16 x 5).  On return v[0] of lane L holds the 32-lane total of column
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1); lanes L and L^1 hold the same column.
__device__ __forceinline__ int colsum16_col(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
template <int W, int BIT>
__device__ __forceinline__ void colsum_step(float (&v)[16], int lane) {
  const bool up = (lane & BIT) != 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float send = up ? v[i] : v[i + W];
    const float keep = up ? v[i + W] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
  }
}
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane) {
  colsum_step<8, 16>(v, lane);
  colsum_step<4, 8>(v, lane);
  colsum_step<2, 4>(v, lane);
  colsum_step<1, 2>(v, lane);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// 128-bit fp32 reduction to global memory (sm_90+): one L2 atomic transaction for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout; version=1, SWIZZLE_128B=2).
// K-major: rows of 128 B, 8-row atoms 1024 B apart (SBO); LBO unused for swizzled K-major.
// g_tc_debug: descriptor-encoding experiments selectable at run time (vinet_debug_set), 0 in production:
//   bit0 swap LBO/SBO of MN-major descriptors, bit1 clear the version field, bit2 LBO=0 for K-major.
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr, uint32_t dbg) {
  const uint64_t lbo = (dbg & 4u) ? 0 : 1, ver = (dbg & 2u) ? 0 : 1;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | ((uint64_t)(1024 >> 4) << 32) | (ver << 46) |
         ((uint64_t)2 << 61);
}
// MN-major: 64 MN elements (128 B) contiguous, next 64-element MN block `lbo` bytes away,
// 8 K rows per atom, atoms 1024 B apart (SBO).
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo, uint32_t dbg) {
  uint64_t l = lbo >> 4, sb = 1024 >> 4;
  if (dbg & 1u) { const uint64_t t = l; l = sb; sb = t; }
  const uint64_t ver = (dbg & 2u) ? 0 : 1;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (l << 16) | (sb << 32) | (ver << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c=F32 (bit4), a=b=BF16 (bits 7,10), majors (15,16), N>>3 (17..22), M>>4 (24..28)
static inline uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


static inline uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

}  // namespace vinet
