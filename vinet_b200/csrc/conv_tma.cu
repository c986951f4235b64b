// TMA-fed persistent tcgen05 implicit-GEMM convolution for sm_100a (engine VINET_ENGINE_TC, kernel VINET_KERNEL_TMA).
//
// The activation operand is never gathered by threads: every (tap, 64-channel block) of the implicit GEMM is ONE
// cp.async.bulk.tensor.5d box {64 ch, bw, bh, 1 frame, 1 clip} of the NDHWC source, addressed at the tap-shifted
// tile origin.  Hardware out-of-bound zero fill IS the convolution's zero padding (spatial and temporal), and the
// 128-byte-swizzled box is exactly the canonical UMMA shared-memory layout: rows = output positions, 128 B = 64 ch.
//
//   conv_gemm_tma_kernel  (fprop + dgrad)   D[128 positions, block_n] = sum_(tap,cblk) A_box x W_blk^T
//     persistent CTAs (one per SM), static round-robin tile schedule, 3 roles:
//       warp 0      TMA producer  : A box (tensor map) + packed weight block (cp.async.bulk) per pipeline stage
//       warp 1      MMA issuer    : tcgen05.mma kind::f16, fp32 accumulators in TMEM, 2 accumulator buffers
//       warps 2..5  epilogue      : tcgen05.ld -> scale/shift/act -> NDHWC store (overlaps the next tile's main loop)
//   conv_wgrad_tma_kernel                   D[(tap,cblk pair) 128, block_n] += sum_positions A_box^T x dY_box
//     both operands are the same boxes read MN-major; split over position chunks, fp32 red.add into the packed
//     TAP64 weight gradient.
//
// A tile of output positions is a bw x bh rectangle of ONE frame (chosen per layer to maximise 128-row use).
#include <cuda.h>

#include <algorithm>

#include "tc_ptx.cuh"

namespace vinet {

constexpr int TMA_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int TMA_WGRAD_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// pull a box into L2 only (no shared-memory destination, no barrier): hides DRAM latency behind the smem ring
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global [%0, {%1, %2, %3, %4, %5}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ bool tma_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// tcgen05.mma with both shared-memory descriptors given as (low, high) 32-bit words
__device__ __forceinline__ void tma_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

struct ConvTmaParams {
  CUtensorMap tmA[2];  // the two sources of the virtual T-concat (tmA[1] == tmA[0] without a concat)
  vinet_conv_t d;
  int32_t bw, bh, tiles_w, tiles_h, ncb, stages, num_items;
  int32_t wres;   // 1: the whole packed weight (one N tile) stays resident in shared memory for the CTA's lifetime
  int32_t halves; // position tiles per work item: 2 = "paired" 256-row item, both tiles share every weight block
  int32_t tpf, ipf;  // tiles per frame, items per frame (= ceil(tpf / halves))
  int32_t nbuf;   // TMEM accumulator sets: 2 = epilogue of item i overlaps the main loop of item i+1
  uint32_t acc_stride, tmem_cols, idesc, a_bytes, b_bytes, a_stage;
};

// One work item = `halves` position tiles of ONE frame (same taps, same source frame) x one N tile.
struct ItemCoord {
  int nt, b, t, tile0, nh;
};

__device__ __forceinline__ ItemCoord decode_item(const ConvTmaParams& p, int item) {
  ItemCoord c;
  c.nt = item % p.d.n_tiles;
  int m = item / p.d.n_tiles;
  const int ip = m % p.ipf; m /= p.ipf;
  const int tr = m % p.d.g.Tr;
  c.b = m / p.d.g.Tr;
  c.t = tr * p.d.g.row_tstep + p.d.g.row_toff;
  c.tile0 = ip * p.halves;
  c.nh = min(p.halves, p.tpf - c.tile0);
  return c;
}

// Source frame read by rows of frame t through temporal tap dt: false when it lies in the temporal padding
// (or, for the transposed convolution, off the stride lattice) - the whole (tile, tap) contributes nothing.
__device__ __forceinline__ bool tap_frame(const vinet_gather_t& g, int t, int dt, int& si, int& tl) {
  int ts;
  if (g.mode == VINET_GATHER_FPROP) {
    ts = t * g.st - g.pt + dt;
  } else {
    const int nt = t + g.pt - dt;
    if (nt < 0) return false;
    ts = nt / g.st;
    if (ts * g.st != nt) return false;
  }
  if ((unsigned)ts >= (unsigned)g.Ts) return false;
  si = (ts >= g.src[0].T) ? 1 : 0;
  tl = ts - (si ? g.src[0].T : 0);
  return true;
}

__device__ __forceinline__ bool tile_has_work(const vinet_gather_t& g, int t) {
  int si, tl;
  for (int tap = 0; tap < g.ntaps; ++tap)
    if (tap_frame(g, t, g.tap[tap][0], si, tl)) return true;
  return false;
}

template <typename TO, bool EPI>
__global__ void __launch_bounds__(TMA_THREADS, 1) conv_gemm_tma_kernel(const __grid_constant__ ConvTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  uint8_t* sA = base;
  uint8_t* sB = sA + (size_t)stages * p.a_stage;  // per-stage weight blocks, or all k_blocks of them when resident
  const int nb_slots = p.wres ? p.d.k_blocks : stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)nb_slots * p.b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 5);
  double* s_stats = reinterpret_cast<double*>(tmem_slot + 4);   // [2][256] per-CTA BatchNorm statistic partials (p.d.stats != nullptr),
                                                                // fp64: order-independent to the last fp32 bit (see conv_stream.cu)
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * stages, tfull0 = empty0 + 8 * stages, tempty0 = tfull0 + 16;
  const uint32_t wbar = tempty0 + 16;
  const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const vinet_gather_t& g = p.d.g;
  const uint32_t nbuf = (uint32_t)p.nbuf;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmA[1]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full0 + 8 * s, 1);
        mbar_init(empty0 + 8 * s, 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull0 + 8 * a, 1);
        mbar_init(tempty0 + 8 * a, 8);
      }
      mbar_init(wbar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (one elected thread)
    if (lane == 0) {
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.d.w);
      const int KB = p.d.k_blocks;
      int s = 0;
      uint32_t ph = 0;
      if (p.wres) {  // fetch the whole weight once (n_tiles == 1): the main loop then only streams activations
        mbar_arrive_expect_tx(wbar, (uint32_t)KB * p.b_bytes);
        for (int kb = 0; kb < KB; ++kb) bulk_copy_g2s(sB0 + (uint32_t)kb * p.b_bytes, wbase + (size_t)kb * p.b_bytes, p.b_bytes, wbar);
      }
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const ItemCoord ic = decode_item(p, item);
        int h0[2], w0[2];
        for (int hf = 0; hf < ic.nh; ++hf) {
          const int tif = ic.tile0 + hf;
          const int th = tif / p.tiles_w;
          h0[hf] = th * p.bh;
          w0[hf] = (tif - th * p.tiles_w) * p.bw;
        }
        const uint32_t tx = (uint32_t)ic.nh * p.a_bytes + (p.wres ? 0u : p.b_bytes);
        for (int tap = 0; tap < g.ntaps; ++tap) {
          int si, tl;
          if (!tap_frame(g, ic.t, g.tap[tap][0], si, tl)) continue;
          const int dh = g.tap[tap][1], dw = g.tap[tap][2];
          const uint8_t* wtap = wbase + ((size_t)ic.nt * KB + (size_t)tap * p.ncb) * p.b_bytes;
          for (int cb = 0; cb < p.ncb; ++cb) {
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            mbar_arrive_expect_tx(full0 + 8 * s, tx);
            for (int hf = 0; hf < ic.nh; ++hf) {
              // FPROP boxes step through the source with the conv's spatial stride (tensor-map element strides)
              const int hc = (g.mode == VINET_GATHER_FPROP) ? h0[hf] * g.sh - g.ph + dh : h0[hf] + g.ph - dh;
              const int wc = (g.mode == VINET_GATHER_FPROP) ? w0[hf] * g.sw - g.pw + dw : w0[hf] + g.pw - dw;
              tma_load_5d(sA0 + (uint32_t)s * p.a_stage + (uint32_t)hf * TC_A_BYTES, &p.tmA[si], full0 + 8 * s, cb * 64, wc, hc,
                          tl, ic.b);
            }
            if (!p.wres) bulk_copy_g2s(sB0 + (uint32_t)s * p.b_bytes, wtap + (size_t)cb * p.b_bytes, p.b_bytes, full0 + 8 * s);
            if (++s == stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: warp-uniform loop, one elected lane issues;
    // 32-bit descriptor words (the high words are loop constants) keep the issue sequence short enough for N <= 128 MMAs
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, descriptor version 1, SWIZZLE_128B
    const uint32_t sA_lo = ((sA0 & 0x3FFFFu) >> 4) | (1u << 16), sB_lo = ((sB0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_stage16 = p.a_stage >> 4, b16 = p.b_bytes >> 4, idesc = p.idesc;
    int s = 0;
    uint32_t ph = 0, lt = 0;
    if (p.wres) mbar_wait(wbar, 0);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const ItemCoord ic = decode_item(p, item);
      if (!tile_has_work(g, ic.t)) continue;
      const uint32_t as = lt % nbuf, aph = (lt / nbuf) & 1u;
      mbar_wait(tempty0 + 8 * as, aph ^ 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * (uint32_t)p.halves * p.acc_stride;
      uint32_t acc = 0;
      for (int tap = 0; tap < g.ntaps; ++tap) {
        int si, tl;
        if (!tap_frame(g, ic.t, g.tap[tap][0], si, tl)) continue;
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const int rem = g.Cs - cb * 64;
          const int nk = rem >= 64 ? 4 : (rem + 15) >> 4;
          const uint32_t b_lo = sB_lo + (uint32_t)(p.wres ? tap * p.ncb + cb : s) * b16;
          uint32_t a_lo = sA_lo + (uint32_t)s * a_stage16;
          uint32_t td = tacc;
          if (tma_elect_one()) {   // ONE elected lane issues both halves of the stage and releases it (see conv_stream.cu: the
            for (int hf = 0; hf < ic.nh; ++hf, a_lo += (TC_A_BYTES >> 4), td += p.acc_stride) {   // issue path is the critical one)
              tma_umma(td, a_lo, hi, b_lo, hi, idesc, acc);
              if (nk == 4) {
                tma_umma(td, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
                tma_umma(td, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
                tma_umma(td, a_lo + 6, hi, b_lo + 6, hi, idesc, 1u);
              } else {
                if (nk > 1) tma_umma(td, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
                if (nk > 2) tma_umma(td, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
              }
            }
            umma_commit(empty0 + 8 * s);
          }
          __syncwarp();
          acc = 1;
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
      if (tma_elect_one()) umma_commit(tfull0 + 8 * as);
      ++lt;
    }
  } else {
    // ---------------------------------------------------------------- epilogue: TMEM -> registers -> global
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;  // the two warps of a quarter take alternate 16-column chunks
    const int row = q * 32 + lane;
    const int rh = row / p.bw, rw = row - rh * p.bw;
    const int BN = p.d.block_n;
    const bool stats = p.d.stats != nullptr;
    const int etid = threadIdx.x - 64;   // 0..255 among the epilogue warps
    if (stats) {
      for (int i = etid; i < 512; i += 256) s_stats[i] = 0.0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    int stats_nt = -1;                   // N tile the partials in shared memory belong to
    auto flush_stats = [&]() {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int nl = p.d.N - stats_nt * BN;
      for (int i = etid; i < 2 * BN; i += 256) {
        const int which = i / BN, col = i - which * BN;
        if (col < nl) {
          atomicAdd(p.d.stats + (size_t)which * p.d.N + stats_nt * BN + col, s_stats[which * 256 + col]);
          s_stats[which * 256 + col] = 0.0;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    uint32_t lt = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const ItemCoord ic = decode_item(p, item);
      const bool any = tile_has_work(g, ic.t);
      const uint32_t as = lt % nbuf, aph = (lt / nbuf) & 1u;
      if (any) {
        mbar_wait(tfull0 + 8 * as, aph);
        tc_fence_after();
      }
      const int nlim = p.d.N - ic.nt * BN;  // columns of this tile that exist
      auto half_row = [&](int hf, bool& accum) -> TO* {
        const int tif = ic.tile0 + hf;
        const int th = tif / p.tiles_w;
        const int h = th * p.bh + rh, w = (tif - th * p.tiles_w) * p.bw + rw;
        const bool valid = row < p.bw * p.bh && h < g.Hr && w < g.Wr;
        if (!valid) return nullptr;
        RowCoord rc;
        rc.b = ic.b; rc.t = ic.t; rc.h = h; rc.w = w;
        accum = (p.d.accumulate >> out_index(p.d, rc)) & 1;
        return out_row_ptr<TO>(p.d, rc) + ic.nt * BN;
      };
      if (!stats) {
        for (int hf = 0; hf < ic.nh; ++hf) {
          bool accum = false;
          TO* orow = half_row(hf, accum);
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (as * (uint32_t)p.halves + (uint32_t)hf) * p.acc_stride;
          for (int gi = half; gi < BN / 16; gi += 2) {
            uint32_t r[16];
            if (any) {
              tmem_ld16(tacc + (uint32_t)(gi * 16), r);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = 0u;
            }
            if (orow == nullptr) continue;
            const int c0 = gi * 16;
            epilogue_store16<TO, EPI>(p.d, orow + c0, r, ic.nt * BN + c0, accum, nlim - c0);
          }
        }
      } else {
        // BatchNorm statistics from the fp32 accumulators (see conv_stream.cu): a CTA's items may belong to different N tiles
        // (items enumerate nt fastest), so the partials are flushed whenever the N tile changes
        if (stats_nt != ic.nt) {
          if (stats_nt >= 0) flush_stats();
          stats_nt = ic.nt;
        }
        for (int gi = half; gi < BN / 16; gi += 2) {
          const int c0 = gi * 16;
          float sv[16], sq[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) sv[e] = sq[e] = 0.f;
          for (int hf = 0; hf < ic.nh; ++hf) {
            bool accum = false;
            TO* orow = half_row(hf, accum);
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (as * (uint32_t)p.halves + (uint32_t)hf) * p.acc_stride;
            uint32_t r[16];
            if (any) {
              tmem_ld16(tacc + (uint32_t)c0, r);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) r[e] = 0u;
            }
            if (orow == nullptr) continue;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float x = __uint_as_float(r[e]);
              sv[e] += x;
              sq[e] = fmaf(x, x, sq[e]);
            }
            epilogue_store16<TO, EPI>(p.d, orow + c0, r, ic.nt * BN + c0, accum, nlim - c0);
          }
          warp_colsum16(sv, lane);
          warp_colsum16(sq, lane);
          const int col = c0 + colsum16_col(lane);
          if (col < nlim) atomicAdd(&s_stats[(lane & 1) * 256 + col], (double)((lane & 1) ? sq[0] : sv[0]));
        }
      }
      if (any) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
        ++lt;
      }
    }
    if (stats && stats_nt >= 0) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// =====================================================================================================
// wgrad
// =====================================================================================================
struct WgradTmaParams {
  CUtensorMap tmA[2];
  CUtensorMap tmDy;
  vinet_wgrad_t d;
  int32_t bw, bh, tiles_w, tiles_h, ncb, nunits, stages, block_n, nblk, R, splits, mb, transpose;
  uint32_t acc_cols, idesc, unit_bytes, stage_bytes;
};

// One CTA accumulates `mb` (1 or 2) 128-row blocks of the packed gradient = 2*mb (tap, 64-channel) units against ONE
// dY tile per position chunk (the dY box is fetched once for all of them), over its share of the position chunks.
__global__ void __launch_bounds__(TMA_WGRAD_THREADS, 1) conv_wgrad_tma_kernel(const __grid_constant__ WgradTmaParams p) {
  const vinet_gather_t& g = p.d.g;
  const int64_t nchunks = (int64_t)g.B * g.Tr * p.tiles_h * p.tiles_w;
  const int64_t per = cdiv(nchunks, p.splits);
  const int64_t c_begin = (int64_t)blockIdx.z * per;
  const int64_t c_end = min(nchunks, c_begin + per);
  if (c_end <= c_begin) return;  // uniform for the CTA
  const int KB = (int)(c_end - c_begin);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)stages * p.stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * stages, accum_bar = empty0 + 8 * stages;
  const uint32_t s0 = smem_u32(base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.block_n;
  const int n0 = blockIdx.y * BN;
  const int upc = 2 * p.mb;  // unit slots per CTA
  const int u0 = blockIdx.x * upc;
  const int nu = min(upc, p.nunits - u0);
  const int nmb = (nu + 1) >> 1;  // 128-row accumulators in use
  const int nblk_eff = min(p.nblk, (min(BN, p.d.N - n0) + 63) / 64);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmA[1]);
    tma_prefetch_desc(&p.tmDy);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full0 + 8 * s, 1);
        mbar_init(empty0 + 8 * s, 1);
      }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), p.acc_cols * p.mb);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // first chunk decoded with divisions once, then walked incrementally (tw fastest, then th, frame, clip)
      int64_t m = c_begin;
      int tw = (int)(m % p.tiles_w); m /= p.tiles_w;
      int th = (int)(m % p.tiles_h); m /= p.tiles_h;
      int tr = (int)(m % g.Tr);
      int b = (int)(m / g.Tr);
      // per-unit constants
      int u_cb[4], u_dt[4], u_dh[4], u_dw[4];
      for (int j = 0; j < nu; ++j) {
        const int u = u0 + j;
        const int tap = u / p.ncb;
        u_cb[j] = (u - tap * p.ncb) * 64;
        u_dt[j] = g.tap[tap][0];
        u_dh[j] = g.tap[tap][1] - g.ph;
        u_dw[j] = g.tap[tap][2] - g.pw;
      }
      const bool cat = g.src[1].ptr != nullptr;
      const int T0 = g.src[0].T;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (uint32_t)(nu + nblk_eff) * p.unit_bytes;
      for (int kb = 0; kb < KB; ++kb) {
        const int t = tr * g.row_tstep + g.row_toff;
        const int h0 = th * p.bh, w0 = tw * p.bw;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(full0 + 8 * s, tx);
        const uint32_t stage = s0 + (uint32_t)s * p.stage_bytes;
        for (int j = 0; j < nu; ++j) {
          const int ts = t * g.st - g.pt + u_dt[j];
          // frames outside [0,Ts) are addressed out of bounds on purpose: TMA zero-fills the temporal padding
          const int si = (cat && ts >= T0) ? 1 : 0;
          tma_load_5d(stage + (uint32_t)j * p.unit_bytes, &p.tmA[si], full0 + 8 * s, u_cb[j], w0 * g.sw + u_dw[j],
                      h0 * g.sh + u_dh[j], ts - (si ? T0 : 0), b);
        }
        for (int nb = 0; nb < nblk_eff; ++nb)
          tma_load_5d(stage + (uint32_t)(upc + nb) * p.unit_bytes, &p.tmDy, full0 + 8 * s, n0 + nb * 64, w0, h0, tr, b);
        if (++s == stages) { s = 0; ph ^= 1u; }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            if (++tr == g.Tr) { tr = 0; ++b; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, elected lane, 32-bit descriptor words (MN-major: LBO = unit_bytes between 64-wide blocks)
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lbo = (p.unit_bytes >> 4) << 16;
    const uint32_t unit16 = p.unit_bytes >> 4, stage16 = p.stage_bytes >> 4, idesc = p.idesc;
    const uint32_t s0_16 = (s0 & 0x3FFFFu) >> 4;
    const int ksteps = p.R / 16;
    int s = 0;
    uint32_t ph = 0, st16 = s0_16;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      const uint32_t b_lo0 = (st16 + (uint32_t)upc * unit16) | lbo;
      if (tma_elect_one()) {   // one elected lane issues every accumulator's MMAs of the stage and releases it
        for (int mbi = 0; mbi < nmb; ++mbi) {
          const uint32_t a_lo0 = (st16 + (uint32_t)(2 * mbi) * unit16) | lbo;
          const uint32_t tacc = tmem_base + (uint32_t)mbi * p.acc_cols;
          // 16 reduction rows (positions) per MMA = two 8-row swizzle atoms = 2048 B
          for (int kk = 0; kk < ksteps; ++kk)
            tma_umma(tacc, a_lo0 + (uint32_t)kk * 128u, hi, b_lo0 + (uint32_t)kk * 128u, hi, idesc, (uint32_t)((kb | kk) != 0));
        }
        umma_commit(empty0 + 8 * s);
        if (kb == KB - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      st16 += stage16;
      if (++s == stages) { s = 0; ph ^= 1u; st16 = s0_16; }
    }
  } else {
    mbar_wait(accum_bar, 0);  // every MMA (hence every TMA write) of this CTA has completed: the stage ring is free
    tc_fence_after();
    fence_proxy_async();
    const int q = warp & 3;
    // Transposing epilogue: a lane owns one accumulator ROW in TMEM, but rows of the packed gradient are lddw floats
    // apart, so direct red.add would touch 32 lines per instruction.  Each warp parks its 32 rows in (its own part
    // of) the free stage ring and re-reads them row by row: 32 lanes = 32 consecutive columns = one 128-byte red.
    const int P = BN + 1;  // odd pitch: conflict-free column-wise writes
    float* stile = reinterpret_cast<float*>(base) + (size_t)(q * 32) * P;
    for (int mbi = 0; mbi < nmb; ++mbi) {
      const int m0 = (u0 + 2 * mbi) * 64 + q * 32;
      for (int gi = 0; gi < BN / 16; ++gi) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mbi * p.acc_cols + (uint32_t)(gi * 16), r);
        if (p.transpose) {
#pragma unroll
          for (int e = 0; e < 16; ++e) stile[lane * P + gi * 16 + e] = __uint_as_float(r[e]);
        } else {
          const int m = m0 + lane;
          if (m < p.nunits * 64) {
            float* drow = p.d.dwp + (int64_t)m * p.d.lddw;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int n = n0 + gi * 16 + e;
              if (n < p.d.N) atomicAdd(drow + n, __uint_as_float(r[e]));
            }
          }
        }
      }
      if (p.transpose) {
        __syncwarp();
        const int ncols = min(BN, p.d.N - n0);
        const int rows = min(32, p.nunits * 64 - m0);
        for (int rr = 0; rr < rows; ++rr) {
          float* drow = p.d.dwp + (int64_t)(m0 + rr) * p.d.lddw + n0;
          const float* srow = stile + rr * P;
          for (int col = lane * 4; col < ncols; col += 128) red_add_v4(drow + col, srow[col], srow[col + 1], srow[col + 2], srow[col + 3]);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.acc_cols * p.mb);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static std::atomic<void*> cached{nullptr};
  void* f = cached.load(std::memory_order_acquire);
  if (f == nullptr) {
    cudaDriverEntryPointQueryResult q;
    void* sym = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    cached.store(sym, std::memory_order_release);
    f = sym;
  }
  return reinterpret_cast<EncodeTiledFn>(f);
}

// NDHWC bf16 view [B][T][H][W][C] (pixel stride ld, row pitch ldh) -> 5-D tiled map, box {64, bw, bh, 1, 1} taking every
// esw-th / esh-th pixel (strided convolutions), 128B swizzle.  ld < C gives overlapping sliding-window rows (WIN8).
static int make_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, int64_t ld, int64_t ldh, int bw, int bh,
                    int esw = 1, int esh = 1, int bt = 1, int64_t ldb = 0) {
  EncodeTiledFn enc = encode_fn();
  VINET_CHECK(enc != nullptr, "conv_tma: cuTensorMapEncodeTiled is not available from the driver");
  VINET_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0, "conv_tma: source must be 16-byte aligned (ld %lld)",
              (long long)ld);
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  if (ldh == 0) ldh = (int64_t)W * ld;
  VINET_CHECK(ldh % 8 == 0 && bw * esw <= 256 && bh * esh <= 256 && esw <= 8 && esh <= 8 && bt >= 1 && bt <= 256, "conv_tma: bad pitch / box");
  if (ldb == 0) ldb = (int64_t)T * H * ldh;        // dense batch; one frame = overlapping sliding windows (vinet_src_t.ldb)
  VINET_CHECK(ldb % 8 == 0, "conv_tma: batch pitch %lld", (long long)ldb);
  const cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ldh * 2, (cuuint64_t)H * ldh * 2, (cuuint64_t)ldb * 2};
  const cuuint32_t box[5] = {64, (cuuint32_t)(bw * esw), (cuuint32_t)(bh * esh), (cuuint32_t)bt, 1};
  const cuuint32_t es[5] = {1, (cuuint32_t)esw, (cuuint32_t)esh, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VINET_CHECK(r == CUDA_SUCCESS, "conv_tma: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d,%d] ld %lld box %dx%d", (int)r,
              B, T, H, W, C, (long long)ld, bw, bh);
  return 0;
}

// The 4-channel clip [B][T][H][Wp][4] (bf16, Wp even) as pixel PAIRS: 5-D map {8 elements, Wp/2, H, T, B}, box {8, bw pairs, bh
// rows, 1, 1}, no swizzle: a compact row-major patch the un-swizzled UMMA descriptors of conv_gemm_stream_win4 read in place.
int make_tma_map_pairs(CUtensorMap* m, const void* ptr, int Wp, int H, int T, int B, int bw, int bh, int64_t ldb) {
  EncodeTiledFn enc = encode_fn();
  VINET_CHECK(enc != nullptr, "conv_tma: cuTensorMapEncodeTiled is not available from the driver");
  VINET_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && Wp % 2 == 0 && bw >= 1 && bw <= 256 && bh >= 1 && bh <= 256,
              "conv_tma: bad 4-channel clip view (Wp %d box %dx%d)", Wp, bw, bh);
  const int64_t ldh = (int64_t)Wp * 4;
  if (ldb == 0) ldb = (int64_t)T * H * ldh;
  VINET_CHECK(ldb % 8 == 0, "conv_tma: batch pitch %lld", (long long)ldb);
  const cuuint64_t dims[5] = {8, (cuuint64_t)(Wp / 2), (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t strides[4] = {16, (cuuint64_t)ldh * 2, (cuuint64_t)H * ldh * 2, (cuuint64_t)ldb * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VINET_CHECK(r == CUDA_SUCCESS, "conv_tma: cuTensorMapEncodeTiled failed (%d) for the 4-channel clip [%d,%d,%d,%d] box %dx%d", (int)r,
              B, T, H, Wp, bw, bh);
  return 0;
}

// bw x bh rectangle (<= max_rows positions, rows a multiple of `mult`) maximising the fraction of useful rows;
// ties go to more rows per box, then to wider boxes (longer contiguous runs for the TMA engine).
static void pick_box(int H, int W, int max_rows, int mult, bool full_tile_cost, int* bw_out, int* bh_out) {
  double best = -1.0;
  int bbw = 1, bbh = 1;
  for (int bw = 1; bw <= W && bw <= max_rows; ++bw) {
    for (int bh = 1; bw * bh <= max_rows && bh <= 256; ++bh) {
      if ((bw * bh) % mult != 0) continue;
      if (bh > H && ((bh - 1) >= H && (bw * (bh - 1)) % mult == 0)) break;  // taller only adds zero-filled rows
      const double tiles = (double)cdiv(H, bh) * (double)cdiv(W, bw);
      const double cost = tiles * (full_tile_cost ? (double)max_rows : (double)(bw * bh));
      const double util = (double)H * W / cost;
      const double score = util + 1e-6 * (bw * bh) + 1e-9 * bw;
      if (score > best) { best = score; bbw = bw; bbh = bh; }
    }
  }
  *bw_out = bbw;
  *bh_out = bbh;
}

static int sm_count() {
  static std::atomic<int> cached{0};
  int n = cached.load();
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cached.store(n);
  }
  return n;
}

// shared with conv_stream.cu
int make_tma_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, int64_t ld, int64_t ldh, int bw, int bh,
                 int esw, int esh, int bt, int64_t ldb) {
  return make_map(m, ptr, C, W, H, T, B, ld, ldh, bw, bh, esw, esh, bt, ldb);
}
void pick_tma_box(int H, int W, int max_rows, int mult, bool full_tile_cost, int* bw_out, int* bh_out) {
  pick_box(H, W, max_rows, mult, full_tile_cost, bw_out, bh_out);
}
int tma_sm_count() { return sm_count(); }
int conv_gemm_stream(const vinet_conv_t* d, cudaStream_t stream);
int conv_wgrad_halo(const vinet_wgrad_t* d, cudaStream_t stream, bool dry_run = false);

// development switch (vinet_debug_set key 1): 0 disables the paired 256-row work items
int g_tma_pair = 1;
int tma_pair_set(int v) { g_tma_pair = v; return 0; }

static int check_tma_gather(const vinet_gather_t& g, const char* what) {
  VINET_CHECK(g.dtype == VINET_BF16, "%s: the TMA kernel needs bf16 sources", what);
  VINET_CHECK((g.sh == 1 && g.sw == 1) || g.mode == VINET_GATHER_FPROP,
              "%s: the TMA kernel handles spatial strides (%d,%d) only for FPROP gathers", what, g.sh, g.sw);
  VINET_CHECK(g.Cs % 8 == 0, "%s: Cs %d must be a multiple of 8", what, g.Cs);
  // source 0 may be a low-res tensor read through the fused relu + 2x up-sampling (FPROP gathers; interpolating producer warps)
  const bool up2 = (g.src[0].xform & ~VINET_XF_RELU) == VINET_XF_UP2 && g.mode == VINET_GATHER_FPROP;
  VINET_CHECK((g.src[0].xform == VINET_XF_IDENT || up2) && (g.src[1].ptr == nullptr || g.src[1].xform == VINET_XF_IDENT),
              "%s: the TMA kernel cannot apply pending source transforms", what);
  VINET_CHECK(g.src[0].T + (g.src[1].ptr ? g.src[1].T : 0) == g.Ts, "%s: Ts %d != sum of source frames", what, g.Ts);
  return 0;
}

int conv_gemm_tma(const vinet_conv_t* d, cudaStream_t stream) {
  const vinet_gather_t& g = d->g;
  if (check_tma_gather(g, "conv_gemm_tma")) return -1;
  VINET_CHECK(d->stats == nullptr || (!conv_has_epilogue(*d) && d->accumulate == 0 && d->block_n <= 256),
              "conv_gemm_tma: epilogue statistics need a raw (no epilogue, non-accumulating) output");
  {  // convolutions with tap re-use (spatial halo / temporal frame sharing) go to the streaming kernel, conv_stream.cu
    const int r = conv_gemm_stream(d, stream);
    if (r != 0) return r < 0 ? r : 0;
  }
  VINET_CHECK(g.src[0].xform == VINET_XF_IDENT, "conv_gemm_tma: no fused up-sampling input stage for this geometry "
              "(vinet_conv_up2_fused tells; materialise with vinet_upsample_fwd)");
  VINET_CHECK(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "conv_gemm_tma: bad block_n %d", d->block_n);
  VINET_CHECK(d->N % 8 == 0, "conv_gemm_tma: N %d must be a multiple of 8", d->N);
  ConvTmaParams p;
  p.d = *d;
  p.ncb = (g.Cs + 63) / 64;
  VINET_CHECK(d->k_blocks == g.ntaps * p.ncb, "conv_gemm_tma: k_blocks %d != ntaps*ceil(Cs/64) = %d (TAP64 weights expected)",
              d->k_blocks, g.ntaps * p.ncb);
  pick_box(g.Hr, g.Wr, TC_BM, 1, true, &p.bw, &p.bh);
  p.tiles_w = (int)cdiv(g.Wr, p.bw);
  p.tiles_h = (int)cdiv(g.Hr, p.bh);
  p.tpf = p.tiles_w * p.tiles_h;
  p.a_bytes = (uint32_t)(p.bw * p.bh * 128);
  p.b_bytes = (uint32_t)d->block_n * 128u;
  p.idesc = make_idesc(TC_BM, d->block_n, 0, 0);
  p.acc_stride = (uint32_t)round_up(d->block_n, 32);
  // These kernels run at the L2->SM bandwidth roof (~9.4 TB/s chip-wide), so what matters is bytes per FLOP.
  // Pairing two position tiles of a frame per work item re-uses every weight block for 256 rows.
  const int64_t frames = (int64_t)g.B * g.Tr;
  const int64_t single_items = frames * p.tpf * d->n_tiles;
  // Measured on B200 (tools/diag_tma.py): pairing wins whenever the four accumulators (2 tiles x 2 buffers) still fit
  // the 512 TMEM columns (block_n <= 128); above that the lost epilogue overlap and the shallower ring cost more.
  p.halves = (p.tpf >= 2 && g_tma_pair && 4u * p.acc_stride <= 512u && frames * cdiv(p.tpf, 2) * d->n_tiles >= 64) ? 2 : 1;
  (void)single_items;
  p.ipf = (int)cdiv(p.tpf, p.halves);
  const int64_t items = frames * p.ipf * d->n_tiles;
  VINET_CHECK(items < (1ll << 31), "conv_gemm_tma: too many tiles");
  p.num_items = (int)items;
  p.nbuf = (2u * p.halves * p.acc_stride <= 512u) ? 2 : 1;
  const uint32_t need_cols = (uint32_t)p.nbuf * p.halves * p.acc_stride;
  p.tmem_cols = tmem_cols_for((int)need_cols);
  p.a_stage = (uint32_t)p.halves * TC_A_BYTES;
  const size_t wbytes = (size_t)d->k_blocks * p.b_bytes;
  p.wres = (d->n_tiles == 1 && wbytes <= 96 * 1024 && items > 2 * (int64_t)sm_count()) ? 1 : 0;
  const size_t stage_bytes = p.wres ? p.a_stage : p.a_stage + p.b_bytes;
  int stages = (int)((200 * 1024 - (p.wres ? wbytes : 0)) / stage_bytes);
  if (stages > 12) stages = 12;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + (p.wres ? wbytes : 0) + 8 * (2 * stages + 5) + 64 + 4096;
  VINET_CHECK(smem <= 227 * 1024, "conv_gemm_tma: %zu bytes of shared memory", smem);
  for (int i = 0; i < 2; ++i) {
    const vinet_src_t& s = g.src[(i == 1 && g.src[1].ptr == nullptr) ? 0 : i];
    if (make_map(&p.tmA[i], s.ptr, g.Cs, g.Ws, g.Hs, s.T, g.B, s.ld, s.ldh, p.bw, p.bh, g.sw, g.sh, 1, s.ldb)) return -1;
  }
  const unsigned grid = (unsigned)std::min<int64_t>(items, sm_count());
#define LAUNCH_TMA(TO)                                                                                  \
  do {                                                                                                  \
    auto kern = conv_has_epilogue(*d) ? conv_gemm_tma_kernel<TO, true> : conv_gemm_tma_kernel<TO, false>; \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
    kern<<<grid, TMA_THREADS, smem, stream>>>(p);                                                       \
  } while (0)
  VINET_DISPATCH_DTYPE(d->out_dtype, TO, LAUNCH_TMA(TO));
#undef LAUNCH_TMA
  note_kernel("conv_gemm_tma_kernel");
  VINET_LAUNCH_OK("conv_gemm_tma");
  return 0;
}

int conv_wgrad_tma(const vinet_wgrad_t* d, cudaStream_t stream) {
  const vinet_gather_t& g = d->g;
  if (check_tma_gather(g, "conv_wgrad_tma")) return -1;
  VINET_CHECK(g.mode == VINET_GATHER_FPROP, "conv_wgrad_tma: needs an FPROP gather");
  {  // spatial convolutions: all kh*kw taps of a channel block share one halo tile per chunk (conv_wgrad_halo.cu)
    const int r = conv_wgrad_halo(d, stream);
    if (r == 0 && g.src[0].xform != VINET_XF_IDENT) {
      set_error("conv_wgrad_tma: no fused up-sampling input stage for this geometry (vinet_conv_up2_fused tells)");
      return -1;
    }
    if (r != 0) return r < 0 ? r : 0;
  }
  VINET_CHECK(d->dy_dtype == VINET_BF16 && d->N % 8 == 0 && d->lddy % 8 == 0, "conv_wgrad_tma: dy must be bf16, N %d lddy %lld",
              d->N, (long long)d->lddy);
  WgradTmaParams p;
  p.d = *d;
  p.ncb = (g.Cs + 63) / 64;
  p.nunits = g.ntaps * p.ncb;
  const int n16 = (int)round_up(d->N, 16);
  const int n_tiles = (int)cdiv(n16, 256);
  p.block_n = (int)round_up(cdiv(n16, n_tiles), 16);
  p.nblk = (p.block_n + 63) / 64;
  p.acc_cols = tmem_cols_for(p.block_n);
  const int mblocks = (int)cdiv(p.nunits, 2);
  p.mb = mblocks >= 2 ? 2 : 1;  // two accumulators share every dY tile: halves the dY traffic per FLOP
  // position chunk: as many rows as leave >= 3 pipeline stages in shared memory
  const int slots = 2 * p.mb + p.nblk;
  const int max_rows = (slots * 128 * 128 * 3 <= 200 * 1024) ? 128 : 64;
  pick_box(g.Hr, g.Wr, max_rows, 16, false, &p.bw, &p.bh);
  p.R = p.bw * p.bh;
  p.tiles_w = (int)cdiv(g.Wr, p.bw);
  p.tiles_h = (int)cdiv(g.Hr, p.bh);
  p.unit_bytes = (uint32_t)p.R * 128u;
  p.stage_bytes = (uint32_t)slots * p.unit_bytes;
  p.idesc = make_idesc(TC_BM, p.block_n, 1, 1);
  const int64_t nchunks = (int64_t)g.B * g.Tr * p.tiles_h * p.tiles_w;
  const int gx = (int)cdiv(mblocks, p.mb);
  // split the position chunks so that the CTA count fills whole waves of the machine (1 CTA per SM)
  const int64_t base_ctas = (int64_t)gx * n_tiles;
  const int sms = sm_count();
  const int64_t smax = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nchunks / 4, 128), 65535));
  int64_t splits = 1;
  double best = 1e30;
  for (int64_t sp = 1; sp <= smax; ++sp) {
    // time in chunk units: waves x (main loop + ~3 chunks of prologue / transposing red.add epilogue)
    const double t = (double)cdiv(base_ctas * sp, sms) * ((double)cdiv(nchunks, sp) + 3.0);
    if (t < best * 0.999) { best = t; splits = sp; }
  }
  p.splits = (int)splits;
  int stages = (int)((200 * 1024) / p.stage_bytes);
  stages = std::max(2, std::min(stages, 8));
  stages = (int)std::max<int64_t>(1, std::min<int64_t>(stages, cdiv(nchunks, splits)));
  p.stages = stages;
  p.transpose = ((size_t)stages * p.stage_bytes >= (size_t)128 * (p.block_n + 1) * sizeof(float)) ? 1 : 0;
  const size_t smem = 1024 + (size_t)stages * p.stage_bytes + 8 * (2 * stages + 1) + 64;
  for (int i = 0; i < 2; ++i) {
    const vinet_src_t& s = g.src[(i == 1 && g.src[1].ptr == nullptr) ? 0 : i];
    if (make_map(&p.tmA[i], s.ptr, g.Cs, g.Ws, g.Hs, s.T, g.B, s.ld, s.ldh, p.bw, p.bh, g.sw, g.sh, 1, s.ldb)) return -1;
  }
  if (make_map(&p.tmDy, d->dy, d->N, g.Wr, g.Hr, g.Tr, g.B, d->lddy, 0, p.bw, p.bh)) return -1;
  dim3 grid((unsigned)gx, (unsigned)n_tiles, (unsigned)splits);
  cudaFuncSetAttribute(conv_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  conv_wgrad_tma_kernel<<<grid, TMA_WGRAD_THREADS, smem, stream>>>(p);
  note_kernel("conv_wgrad_tma_kernel");
  VINET_LAUNCH_OK("conv_wgrad_tma");
  return 0;
}

}  // namespace vinet
