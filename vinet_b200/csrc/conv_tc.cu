// tcgen05 implicit-GEMM convolution for sm_100a (engine VINET_ENGINE_TC).
//
//   conv_gemm_tc : D[128 rows, block_n] = gather(rows, K) x W^T      (fprop and dgrad gathers)
//       A (activations)  : gathered by 8 producer warps with the pending BN/ReLU transform applied in
//                          registers, converted to bf16 and written to shared memory in the canonical
//                          K-major SWIZZLE_128B layout (8 rows x 128 B atoms);
//       B (weights)      : pre-packed per (n_tile, k_block) in exactly that layout, fetched by one
//                          cp.async.bulk (TMA bulk copy, SASS UBLKCP) per stage, completing on the
//                          stage's "full" mbarrier;
//       D (accumulators) : fp32 in tensor memory (TMEM), issued by one elected thread
//                          (tcgen05.mma.cta_group::1.kind::f16, SASS UTCHMMA), drained by tcgen05.ld.
//   conv_wgrad_tc : D[128 (tap,c), block_n] += sum over rows  gather(row,(tap,c)) * dy[row, n]
//       both operands are MN-major (rows = reduction index are the 128-byte-strided dimension),
//       split over the reduction dimension across CTAs, fp32 red.global.add into the packed gradient.
//
// Pipeline: `stages` x {A tile, B tile, full mbarrier, empty mbarrier}; producers wait on empty,
// the MMA thread waits on full and releases a stage with tcgen05.commit -> empty.
#include "tc_ptx.cuh"

namespace vinet {

__device__ unsigned int g_tc_debug = 0;
constexpr int TC_PRODUCER_THREADS = 256;
constexpr int TC_THREADS = 320;  // 8 producer/epilogue warps + MMA warp + weight-loader warp

// raw 8-element chunk of a source row
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void to_float(float (&v)[8]) const {
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0, 0, 0, 0); }
  __device__ __forceinline__ void to_float(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// transform + convert one gathered chunk to 8 bf16 (16 bytes)
template <typename T>
__device__ __forceinline__ uint4 finish_chunk(const Raw8<T>& raw, bool valid, const vinet_src_t& s, int c) {
  if (!valid) return make_uint4(0, 0, 0, 0);
  if constexpr (sizeof(T) == 2) {
    if (s.xform == VINET_XF_IDENT) return raw.u;
  }
  float v[8];
  raw.to_float(v);
  apply_xform<8>(v, s.xform, s.scale, s.shift, c);
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

__device__ __forceinline__ void st_shared_16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct TcSmem {
  uint8_t* base;  // 1024-aligned
  __device__ __forceinline__ explicit TcSmem(uint8_t* raw) {
    base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  }
};

// =====================================================================================================
// fprop / dgrad
// =====================================================================================================
template <typename T, typename TO, bool EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ vinet_conv_t d, int stages, uint32_t tmem_cols, uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  TcSmem sm(smem_raw);
  const int BN = d.block_n;
  const uint32_t b_bytes = (uint32_t)BN * (TC_BK * 2);
  uint8_t* sA = sm.base;
  uint8_t* sB = sA + (size_t)stages * TC_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)stages * b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);
  int4* rowinfo = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + stages), accum_bar = smem_u32(bars + 2 * stages);
  const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t M = gather_rows(d.g);
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int nt = blockIdx.y;
  const int KB = d.k_blocks;

  if (tid < TC_BM) {
    RowCoord rc = decode_row(d.g, m0 + tid, M);
    rowinfo[tid] = make_int4(rc.b, rc.t, rc.h, rc.w);
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full0 + 8 * s, TC_PRODUCER_THREADS + 1);
        mbar_init(empty0 + 8 * s, 1);
      }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp < 8) {
    // ------------------------------------------------------------ A producers
    const int j = tid & 7, r0 = tid >> 3;
    RowCoord rc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int4 v = rowinfo[r0 + 32 * i];
      rc[i].b = v.x; rc[i].t = v.y; rc[i].h = v.z; rc[i].w = v.w;
    }
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % stages;
      const uint32_t ph = (uint32_t)(kb / stages) & 1u;
      const int k = kb * TC_BK + j * 8;
      const int tap = k / d.g.Cs;
      const int c = k - tap * d.g.Cs;
      Raw8<T> raw[4];
      int si[4];
      bool ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int64_t off;
        si[i] = 0;
        ok[i] = gather_locate(d.g, rc[i], tap, si[i], off);
        if (ok[i]) raw[i].load(reinterpret_cast<const T*>(d.g.src[si[i]].ptr) + off + c);
        else raw[i].zero();
      }
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      const uint32_t a_stage = sA0 + (uint32_t)s * TC_A_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = r0 + 32 * i;
        const uint4 v = finish_chunk<T>(raw[i], ok[i], d.g.src[si[i]], c);
        st_shared_16(a_stage + row * 128 + ((j ^ (row & 7)) << 4), v);
      }
      fence_proxy_async();
      mbar_arrive(full0 + 8 * s);
    }
    // ------------------------------------------------------------ epilogue: TMEM -> registers -> global
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const int4 ri = rowinfo[row];
    RowCoord orc;
    orc.b = ri.x; orc.t = ri.y; orc.h = ri.z; orc.w = ri.w;
    TO* orow = (orc.b >= 0) ? out_row_ptr<TO>(d, orc) : nullptr;
    const bool accum = (orc.b >= 0) && ((d.accumulate >> out_index(d, orc)) & 1);
    for (int g = half; g < BN / 16; g += 2) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 16), r);
      if (orow == nullptr) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = nt * BN + g * 16 + h * 8;
        if (n >= d.N) continue;
        epilogue_store8<TO, EPI>(d, orow + n, r + h * 8, n, accum);
      }
    }
    tc_fence_before();
  } else if (warp == 8) {
    // ------------------------------------------------------------ MMA issuer (one elected thread)
    const uint32_t dbg = g_tc_debug;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % stages;
      const uint32_t ph = (uint32_t)(kb / stages) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_stage = sA0 + (uint32_t)s * TC_A_BYTES;
        const uint32_t b_stage = sB0 + (uint32_t)s * b_bytes;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 16; ++kk) {
          umma_bf16(tmem_base, desc_kmajor_sw128(a_stage + kk * 32, dbg), desc_kmajor_sw128(b_stage + kk * 32, dbg), idesc,
                    (uint32_t)((kb | kk) != 0));
        }
        umma_commit(empty0 + 8 * s);
        if (kb == KB - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ weight loader (TMA bulk copy)
    if (lane == 0) {
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.w) + (size_t)nt * KB * b_bytes;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(full0 + 8 * s, b_bytes);
        bulk_copy_g2s(sB0 + (uint32_t)s * b_bytes, wp + (size_t)kb * b_bytes, b_bytes, full0 + 8 * s);
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// =====================================================================================================
// wgrad
// =====================================================================================================
template <typename T, typename TD>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ vinet_wgrad_t d, int block_n, int stages, uint32_t tmem_cols,
                     uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  TcSmem sm(smem_raw);
  const int BN = block_n;
  const int nblk = (BN + 63) / 64;                       // 64-wide MN blocks of the dy operand
  const uint32_t b_bytes = (uint32_t)nblk * (64 * 128);  // [nblk][64 rows][128 B]
  uint8_t* sA = sm.base;                                 // [2][64 rows][128 B] per stage
  uint8_t* sB = sA + (size_t)stages * TC_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)stages * b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + stages), accum_bar = smem_u32(bars + 2 * stages);
  const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t M = gather_rows(d.g);
  const int m0 = blockIdx.x * TC_BM;  // (tap,c) offset
  const int n0 = blockIdx.y * BN;
  const int64_t nchunks = cdiv(M, TC_BK);
  const int64_t per = cdiv(nchunks, d.splits);
  const int64_t c_begin = (int64_t)blockIdx.z * per;
  const int64_t c_end = min(nchunks, c_begin + per);
  const int KB = (int)max((int64_t)0, c_end - c_begin);

  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full0 + 8 * s, TC_PRODUCER_THREADS);
        mbar_init(empty0 + 8 * s, 1);
      }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (KB > 0) {
    if (warp < 8) {
      // -------------------------------------------------------- producers: A' (gather) and dy rows
      const int c16 = tid & 15, ra = tid >> 4;  // A': 16 chunks of 8 (tap,c) per row, 16 rows per pass
      const int ka = m0 + c16 * 8;
      const int tap = ka / d.g.Cs;
      const int ca = ka - tap * d.g.Cs;
      const uint32_t a_off = (uint32_t)(c16 >> 3) * 8192u;
      const int nch = BN / 8;
      const T* __restrict__ dummy = nullptr;
      (void)dummy;
      const TD* __restrict__ dy = reinterpret_cast<const TD*>(d.dy);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        const int64_t rbase = (c_begin + kb) * TC_BK;
        Raw8<T> raw[4];
        int si[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const RowCoord rc = decode_row(d.g, rbase + ra + 16 * i, M);
          int64_t off;
          si[i] = 0;
          ok[i] = gather_locate(d.g, rc, tap, si[i], off);
          if (ok[i]) raw[i].load(reinterpret_cast<const T*>(d.g.src[si[i]].ptr) + off + ca);
          else raw[i].zero();
        }
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        const uint32_t a_stage = sA0 + (uint32_t)s * TC_A_BYTES + a_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ra + 16 * i;
          const uint4 v = finish_chunk<T>(raw[i], ok[i], d.g.src[si[i]], ca);
          st_shared_16(a_stage + row * 128 + (((c16 & 7) ^ (row & 7)) << 4), v);
        }
        const uint32_t b_stage = sB0 + (uint32_t)s * b_bytes;
        for (int idx = tid; idx < 64 * nch; idx += TC_PRODUCER_THREADS) {
          const int row = idx / nch, cc = idx - row * nch;
          const int n = n0 + cc * 8;
          const int64_t grow = rbase + row;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (grow < M && n < d.N) {
            Raw8<TD> r;
            r.load(dy + grow * d.lddy + n);
            if constexpr (sizeof(TD) == 2) v = r.u;
            else {
              float f[8];
              r.to_float(f);
              v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                             pack_bf16x2(f[6], f[7]));
            }
          }
          st_shared_16(b_stage + (uint32_t)(cc >> 3) * 8192u + row * 128 + (((cc & 7) ^ (row & 7)) << 4), v);
        }
        fence_proxy_async();
        mbar_arrive(full0 + 8 * s);
      }
      // -------------------------------------------------------- epilogue: red.add into the packed gradient
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const int q = warp & 3, half = warp >> 2;
      const int m = m0 + q * 32 + lane;
      const bool mvalid = m < d.g.ntaps * d.g.Cs;
      float* drow = d.dwp + (int64_t)m * d.lddw;
      for (int g = half; g < BN / 16; g += 2) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 16), r);
        if (!mvalid) continue;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int n = n0 + g * 16 + e;
          if (n < d.N) atomicAdd(drow + n, __uint_as_float(r[e]));
        }
      }
      tc_fence_before();
    } else if (warp == 8) {
      const uint32_t dbg = g_tc_debug;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_stage = sA0 + (uint32_t)s * TC_A_BYTES;
          const uint32_t b_stage = sB0 + (uint32_t)s * b_bytes;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; ++kk) {
            // 16 reduction rows per MMA = two 8-row atoms = 2048 B
            umma_bf16(tmem_base, desc_mnmajor_sw128(a_stage + kk * 2048, 8192, dbg), desc_mnmajor_sw128(b_stage + kk * 2048, 8192, dbg),
                      idesc, (uint32_t)((kb | kk) != 0));
          }
          umma_commit(empty0 + 8 * s);
          if (kb == KB - 1) umma_commit(accum_bar);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
static int pick_stages(size_t stage_bytes, int k_blocks) {
  // prefer two CTAs per SM (epilogue of one overlaps the main loop of the other): <= ~110 KB each
  int st = (int)((110 * 1024) / stage_bytes);
  if (st < 2) st = 2;
  if (st > 6) st = 6;
  if (st > k_blocks) st = k_blocks < 1 ? 1 : k_blocks;
  return st;
}

int tc_debug_set(unsigned int v) {
  return cudaMemcpyToSymbol(g_tc_debug, &v, sizeof(v)) == cudaSuccess ? 0 : -1;
}

int conv_gemm_tc(const vinet_conv_t* d, cudaStream_t stream) {
  VINET_CHECK(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "conv_gemm_tc: bad block_n %d", d->block_n);
  VINET_CHECK(d->g.Cs % 8 == 0, "conv_gemm_tc: Cs %d must be a multiple of 8", d->g.Cs);
  VINET_CHECK(d->N % 8 == 0, "conv_gemm_tc: N %d must be a multiple of 8", d->N);
  VINET_CHECK(d->k_blocks >= 1, "conv_gemm_tc: k_blocks");
  const int64_t M = (int64_t)d->g.B * d->g.Tr * d->g.Hr * d->g.Wr;
  const size_t stage_bytes = TC_A_BYTES + (size_t)d->block_n * 128;
  const int stages = pick_stages(stage_bytes, d->k_blocks);
  const size_t smem = 1024 + stages * stage_bytes + 256 + TC_BM * sizeof(int4);
  const uint32_t idesc = make_idesc(TC_BM, d->block_n, 0, 0);
  const uint32_t cols = tmem_cols_for(d->block_n);
  dim3 grid((unsigned)cdiv(M, TC_BM), (unsigned)d->n_tiles);
#define LAUNCH_GEMM(T, TO)                                                                                     \
  do {                                                                                                         \
    auto kern = conv_has_epilogue(*d) ? conv_gemm_tc_kernel<T, TO, true> : conv_gemm_tc_kernel<T, TO, false>;  \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    kern<<<grid, TC_THREADS, smem, stream>>>(*d, stages, cols, idesc);                                         \
  } while (0)
  VINET_DISPATCH_DTYPE(d->g.dtype, T, VINET_DISPATCH_DTYPE(d->out_dtype, TO, LAUNCH_GEMM(T, TO)));
#undef LAUNCH_GEMM
  note_kernel("conv_gemm_tc_kernel");
  VINET_LAUNCH_OK("conv_gemm_tc");
  return 0;
}

int conv_wgrad_tc(const vinet_wgrad_t* d, cudaStream_t stream) {
  VINET_CHECK(d->g.Cs % 8 == 0 && d->N % 8 == 0, "conv_wgrad_tc: Cs %d / N %d must be multiples of 8", d->g.Cs, d->N);
  VINET_CHECK(d->g.mode == VINET_GATHER_FPROP, "conv_wgrad: needs an FPROP gather");
  const int mw = d->g.ntaps * d->g.Cs;
  // N tiling: one tile when it fits a single UMMA (<=256), else equal tiles
  int n16 = (int)round_up(d->N, 16);
  int n_tiles = (int)cdiv(n16, 256);
  int block_n = (int)round_up(cdiv(n16, n_tiles), 16);
  const size_t stage_bytes = TC_A_BYTES + (size_t)((block_n + 63) / 64) * 8192;
  const int64_t M = (int64_t)d->g.B * d->g.Tr * d->g.Hr * d->g.Wr;
  int stages = pick_stages(stage_bytes, (int)cdiv(cdiv(M, TC_BK), d->splits));
  const size_t smem = 1024 + stages * stage_bytes + 256;
  const uint32_t idesc = make_idesc(TC_BM, block_n, 1, 1);
  const uint32_t cols = tmem_cols_for(block_n);
  dim3 grid((unsigned)cdiv(mw, TC_BM), (unsigned)n_tiles, (unsigned)d->splits);
#define LAUNCH_WGRAD(T, TD)                                                                                    \
  do {                                                                                                         \
    auto kern = conv_wgrad_tc_kernel<T, TD>;                                                                   \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    kern<<<grid, TC_THREADS, smem, stream>>>(*d, block_n, stages, cols, idesc);                                \
  } while (0)
  VINET_DISPATCH_DTYPE(d->g.dtype, T, VINET_DISPATCH_DTYPE(d->dy_dtype, TD, LAUNCH_WGRAD(T, TD)));
#undef LAUNCH_WGRAD
  note_kernel("conv_wgrad_tc_kernel");
  VINET_LAUNCH_OK("conv_wgrad_tc");
  return 0;
}

}  // namespace vinet
