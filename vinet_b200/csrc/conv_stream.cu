// Streaming tcgen05 implicit-GEMM convolution for sm_100a: every source tile is fetched ONCE and re-used by all the
// taps that touch it (engine VINET_ENGINE_TC, kernel VINET_KERNEL_TMA; chosen by conv_gemm_tma when there is re-use).
//
// conv_tma.cu fetches one TMA box per (tap, 64-channel block): a 3x3 conv pulls every activation through L2->SM nine
// times and those kernels sit on the L2->SM bandwidth roof (~33 B/clk/SM), not on the tensor pipe.  Here
//   * spatial taps share ONE halo box {64 ch, 8*nsub + kw-1, 16 + kh-1} per (source frame, channel block).  A UMMA
//     SWIZZLE_128B K-major descriptor may start at any 128-byte row of a TMA-written tile with any stride between
//     8-row groups (the swizzle is a function of the absolute shared-memory address; profiles/r1_umma_shift_test.txt),
//     so tap (dh,dw) of sub-tile s is the descriptor {start = box + ((dh*PW + dw + 8s) * 128), SBO = PW*128}: an
//     8-wide x 16-high window of the halo.  nsub side-by-side sub-tiles also share every weight block;
//   * temporal taps share the source FRAME: a CTA walks the source frames of a run of output frames in order and
//     issues, for each frame, the MMAs of every (tap, output frame) pair it feeds into a ring of TMEM accumulators;
//     an accumulator is committed to the epilogue warps when its last source frame has been consumed;
//   * weights stay resident in shared memory when one N tile of them fits, else stream through their own ring;
//   * row-strided convs over sliding-window rows (the stem conv_s) fetch one box per source-row lattice (parity) with a row
//     element-stride, so a kernel row is again a plain row offset inside its lattice's box (conv_gemm_stream_strided).
//
// Roles (448 threads): warp 0 activation producer (TMA), warp 1 TMEM allocator + weight producer (bulk copies), warps 2..9
// epilogue (two per TMEM lane quarter), warps 10..13 MMA issuers (highest ids = highest issue priority).  grid = (CTAs per N tile, N tiles).
// UP variant (576 threads; source 0 has VINET_XF_UP2, i.e. the decoder's relu -> 2x bilinear up-sampling, model.py:254, sits in
// front of this convolution): four epilogue warps (2..5) instead of eight, warps 6..13 are a second activation producer, 14..17 issue.  For every halo stage that belongs to source 0 they
// read the LOW-RES tensor with 128-bit loads, apply ReLU + the bilinear blend and write the bf16 result into the stage in the
// SWIZZLE_128B layout a TMA box load of the up-sampled tensor would have produced (up2.cuh), fence the generic-proxy writes
// towards the async proxy and arrive on the stage's full barrier.  The up-sampled tensor never exists in memory; stages of
// source 1 (the skip tensor of the T-concat) still arrive by TMA from warp 0, in the same ring.
// Several issuing warps because ONE warp cannot feed the tensor pipe with small-N MMAs: a 128xNx16 MMA lasts max(N/2, 32+N/4)
// clocks (tools/umma_rate_test.cu) while its issue sequence costs a single warp ~80-100 clocks.  Issuer k owns the sub-tiles
// k, k+ni, ... of every work item (its own TMEM accumulators), waits on the same full barriers and commits to the same empty
// barriers (arrival count ni).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "up2.cuh"

namespace vinet {

constexpr int ST_THREADS = 448;
// VINET_XF_UP2 sources: the UP variant trades four of the eight epilogue warps for EIGHT interpolating warps (576 threads).  Its
// decoder convolutions have long K loops (18 .. 45 taps x up to 832 channels), so the epilogue is nowhere near critical, while one
// interpolating warp per scheduler could not fill a halo stage in the time the tensor core needs to consume one.
constexpr int ST_UP_EPI_WARPS = 4;
constexpr int ST_UP_THREADS = 256;
constexpr int ST_UP_TOTAL = 32 * (2 + 4 + ST_UP_EPI_WARPS) + ST_UP_THREADS;   // 576
constexpr int ST_MAX_ISSUERS = 4;
constexpr int ST_MAX_TG = 8;
// per-CTA BatchNorm statistic partials: [sum | sum of squares][column of the N tile], kept in fp64: the epilogue warps of a CTA add
// their lane-reduced sums in whatever order they finish, and an fp32 accumulator made the batch statistics (hence the whole
// train-mode forward, through ~60 normalisations of tiny batches) differ in the last bit from run to run
constexpr int ST_STATS_FLOATS = 2 * 2 * 256;   // in units of 4 bytes

struct StreamParams {
  CUtensorMap tmA[2];
  vinet_conv_t d;
  int32_t halo, nsub, tw, th;      // sub-tile = tw x th output positions (halo: 8 x 16; flat: the TMA box)
  int32_t items_w, items_h;        // work items per frame (flat: items_w groups of nsub consecutive tiles, items_h = 1)
  int32_t tiles_w, tpf;            // flat: tiles per row / per frame
  int32_t PW, PH, ew0, eh0;        // halo box extent and origin offset
  int32_t ncb, run, nruns, S, ntg;
  int32_t e_min, e_max;
  // temporal tap group g (ascending source offset e_g = tg_eq[g]*S + tg_er[g], 0 <= tg_er < S): source frame f = q*S + r
  // feeds output frame i = q - tg_eq[g] when r == tg_er[g]
  int32_t tg_er[ST_MAX_TG], tg_eq[ST_MAX_TG];
  int32_t tg_first[ST_MAX_TG + 1]; // spatial taps of group g: sp_*[tg_first[g] .. tg_first[g+1])
  int32_t sp_aoff[VINET_MAX_TAPS]; // byte offset of the tap's window inside an activation stage
  int32_t sp_kb[VINET_MAX_TAPS];   // first weight k-block of the tap (tap * ncb)
  int32_t sp_aoff16[VINET_MAX_TAPS], sp_boff16[VINET_MAX_TAPS];  // sp_aoff and sp_kb * b_bytes in 16-byte descriptor units
  int32_t nacc, a_stages, b_slots, wres, items_per_nt, ni;
  int32_t tt, pos, tstep, toff, walk_Ts, walk_Tr;  // temporal-halo tiles: tt output frames x pos positions per sub-tile (tt = 1: off)
  int32_t par_h0[2], par_off[2], par_sh;           // halo == 2 (row-strided conv): one box per source-row parity, see conv_gemm_stream_strided
  uint32_t acc_stride, tmem_cols, idesc, a_stage_bytes, a_tx_sub, sub_stride, sbo, b_bytes;
};

struct StItem {
  int b, i0, i1, tx, ty;
};

__device__ __forceinline__ StItem st_decode(const StreamParams& p, int item) {
  StItem c;
  int m = item;
  c.tx = m % p.items_w; m /= p.items_w;
  c.ty = m % p.items_h; m /= p.items_h;
  const int r = m % p.nruns;
  c.b = m / p.nruns;
  c.i0 = r * p.run;
  c.i1 = min(p.walk_Tr, c.i0 + p.run);
  return c;
}

// The work items of one CTA (item = blockIdx.x, += gridDim.x) decoded INCREMENTALLY: the stride is split into its (tx, ty, run,
// clip) digits once, every further item costs a few adds and compares instead of four divisions (the item prologue is on the MMA
// issuers' critical path: the layers with few taps spend as long there as in their MMAs).
struct StItemIter {
  int tx, ty, r, b, dtx, dty, dr, db;
  __device__ __forceinline__ void init(const StreamParams& p, int first, int stride) {
    int m = first;
    tx = m % p.items_w; m /= p.items_w;
    ty = m % p.items_h; m /= p.items_h;
    r = m % p.nruns;
    b = m / p.nruns;
    m = stride;
    dtx = m % p.items_w; m /= p.items_w;
    dty = m % p.items_h; m /= p.items_h;
    dr = m % p.nruns;
    db = m / p.nruns;
  }
  __device__ __forceinline__ StItem get(const StreamParams& p) const {
    StItem c;
    c.tx = tx; c.ty = ty; c.b = b;
    c.i0 = r * p.run;
    c.i1 = min(p.walk_Tr, c.i0 + p.run);
    return c;
  }
  __device__ __forceinline__ void next(const StreamParams& p) {
    tx += dtx;
    int carry = tx >= p.items_w;
    tx -= carry ? p.items_w : 0;
    ty += dty + carry;
    carry = ty >= p.items_h;
    ty -= carry ? p.items_h : 0;
    r += dr + carry;
    carry = r >= p.nruns;
    r -= carry ? p.nruns : 0;
    b += db + carry;
  }
};

// Walk of the real source frames [max(first,0), min(last,Ts-1)] of one work item, carrying f = q*S + r without divisions.
struct StWalk {
  int f, f_end, q, r, i0, n;  // n = i1 - i0
  __device__ __forceinline__ void init(const StreamParams& p, const StItem& c) {
    const int f_lo = c.i0 * p.S + p.e_min;
    f = max(f_lo, 0);
    f_end = min((c.i1 - 1) * p.S + p.e_max, p.walk_Ts - 1);
    q = f / p.S;
    r = f - q * p.S;
    i0 = c.i0;
    n = c.i1 - c.i0;
  }
  __device__ __forceinline__ void next(const StreamParams& p) {
    ++f;
    if (++r == p.S) { r = 0; ++q; }
  }
  // output frame fed through tap group g, or -1
  __device__ __forceinline__ int out_of(const StreamParams& p, int g) const {
    const int i = q - p.tg_eq[g];
    return (r == p.tg_er[g] && (unsigned)(i - i0) < (unsigned)n) ? i : -1;
  }
  __device__ __forceinline__ bool used(const StreamParams& p) const {
    for (int g = 0; g < p.ntg; ++g)
      if (out_of(p, g) >= 0) return true;
    return false;
  }
};

// sub-tiles of this item that hold any output position
__device__ __forceinline__ int st_nsub_eff(const StreamParams& p, const StItem& c) {
  if (p.halo) return min(p.nsub, (p.d.g.Wr - c.tx * 8 * p.nsub + 7) >> 3);
  return min(p.nsub, p.tpf - c.tx * p.nsub);
}

__device__ __forceinline__ void st_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                               int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ bool st_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// tcgen05.mma with the two shared-memory descriptors given as (low, high) words: the high words (SBO, version, swizzle) are
// loop constants and the low words (start address >> 4) advance by plain 32-bit adds, which keeps the issue loop short.
__device__ __forceinline__ void st_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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

// The MMAs of taps [j0, j1) of one (stage, accumulator) for ONE sub-tile, issued by the calling thread alone (the elected lane of
// an issuing warp).  Resident weights: the tap's block sits at b_cb + its table offset.  Streamed weights: the taps arrive through
// the ring (sb, phb, b_lo_slot are the ring position at entry; every lane of the warp advances its copy by j1 - j0 afterwards).
__device__ __forceinline__ void st_issue_taps(const int2* s_tap, int j0, int j1, int nk, uint32_t td, uint32_t a_base, uint32_t a_hi,
                                              uint32_t b_cb, uint32_t b_hi, uint32_t idesc, uint32_t keep, bool wres, uint32_t full_b,
                                              uint32_t empty_b, int sb, uint32_t phb, uint32_t b_lo_slot, uint32_t sB_lo, uint32_t b16,
                                              int BS) {
  if (nk == 4) {
#pragma unroll 3
    for (int j = j0; j < j1; ++j) {
      const int2 tj = s_tap[j];
      uint32_t b_lo;
      if (wres) {
        b_lo = b_cb + (uint32_t)tj.y;
      } else {
        mbar_wait(full_b + 8 * sb, phb);
        tc_fence_after();
        b_lo = b_lo_slot;
      }
      const uint32_t a_lo = a_base + (uint32_t)tj.x;
      st_umma(td, a_lo, a_hi, b_lo, b_hi, idesc, keep);
      st_umma(td, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
      st_umma(td, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
      st_umma(td, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
      keep = 1u;
      if (!wres) {
        umma_commit(empty_b + 8 * sb);
        b_lo_slot += b16;
        if (++sb == BS) { sb = 0; phb ^= 1u; b_lo_slot = sB_lo; }
      }
    }
  } else {
    for (int j = j0; j < j1; ++j) {
      const int2 tj = s_tap[j];
      uint32_t b_lo;
      if (wres) {
        b_lo = b_cb + (uint32_t)tj.y;
      } else {
        mbar_wait(full_b + 8 * sb, phb);
        tc_fence_after();
        b_lo = b_lo_slot;
      }
      const uint32_t a_lo = a_base + (uint32_t)tj.x;
      st_umma(td, a_lo, a_hi, b_lo, b_hi, idesc, keep);
      if (nk > 1) st_umma(td, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
      if (nk > 2) st_umma(td, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
      keep = 1u;
      if (!wres) {
        umma_commit(empty_b + 8 * sb);
        b_lo_slot += b16;
        if (++sb == BS) { sb = 0; phb ^= 1u; b_lo_slot = sB_lo; }
      }
    }
  }
}

// -DVINET_ST_PROF (development builds only): every role accumulates the clocks it spends in its mbarrier waits (the issuers also
// the clocks inside their tcgen05.mma groups) and CTA (0,0) prints them at the end.
#ifdef VINET_ST_PROF
#define ST_PROF_WAIT(acc, ...)            \
  do {                                    \
    const long long t0__ = clock64();     \
    __VA_ARGS__;                          \
    (acc) += clock64() - t0__;            \
  } while (0)
#else
#define ST_PROF_WAIT(acc, ...) \
  do {                         \
    __VA_ARGS__;               \
  } while (0)
#endif

template <typename TO, bool EPI, bool UP>
__global__ void __launch_bounds__(UP ? ST_UP_TOTAL : ST_THREADS, 1) conv_stream_kernel(const __grid_constant__ StreamParams p) {
  [[maybe_unused]] long long prof_w0 = 0, prof_w1 = 0, prof_w2 = 0;
  [[maybe_unused]] const long long prof_t0 = clock64();
  constexpr int EW = UP ? ST_UP_EPI_WARPS : 8;          // epilogue warps (EW / 4 per TMEM lane quarter)
  constexpr int ET = 32 * EW;                           // epilogue threads
  // Warp order = scheduling priority (the SM's arbiter prefers the highest warp id of a scheduler, B300_MICROARCH.md): the MMA issuers
  // take the HIGHEST ids.  With the issuers at warps 2..5, the epilogue warps above them - spinning on their accumulator barriers
  // most of the time - took the issue slots first, and an issuer needed ~900 clocks for the ~45 scalar instructions around
  // four tcgen05.mma (measured with -DVINET_ST_PROF: 8 % of an issuer's time was spent inside its MMA groups).
  constexpr int FIRST_EPI_WARP = 2;
  constexpr int FIRST_UP_WARP = FIRST_EPI_WARP + EW;                       // UP only: ST_UP_THREADS / 32 interpolating warps
  constexpr int FIRST_MMA_WARP = FIRST_UP_WARP + (UP ? ST_UP_THREADS / 32 : 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int AS = p.a_stages, BS = p.b_slots;
  const int KB = p.d.k_blocks;
  uint8_t* sA = base;
  uint8_t* sB = sA + (size_t)AS * p.a_stage_bytes;
  const int nb_slots = p.wres ? KB : BS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)nb_slots * p.b_bytes);
  // barrier layout: full_a[AS] empty_a[AS] full_b[BS] empty_b[BS] full_acc[nacc] empty_acc[nacc] wbar
  const uint32_t full_a = smem_u32(bars), empty_a = full_a + 8 * AS, full_b = empty_a + 8 * AS, empty_b = full_b + 8 * BS;
  const uint32_t full_acc = empty_b + 8 * BS, empty_acc = full_acc + 8 * p.nacc, wbar = empty_acc + 8 * p.nacc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * AS + 2 * BS + 2 * p.nacc + 1);
  double* s_stats = reinterpret_cast<double*>(tmem_slot + 4);   // [2][256] doubles, used when p.d.stats != nullptr
  // per-tap descriptor offsets (activation window, weight block) in shared memory: the issue loop reads them once per tap, and an
  // indexed load from the parameter bank (c[0][R + ...]) costs it several times an LDS
  int2* s_tap = reinterpret_cast<int2*>(s_stats + 2 * 256);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + VINET_MAX_TAPS)
    s_tap[threadIdx.x - 64] = make_int2(p.sp_aoff16[threadIdx.x - 64], p.sp_boff16[threadIdx.x - 64]);
  const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const vinet_gather_t& g = p.d.g;
  const int nt = blockIdx.y;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA[1])) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < AS; ++s) {
        mbar_init(full_a + 8 * s, UP ? 2 : 1);   // UP: the TMA thread and the interpolating group both arrive on every stage
        mbar_init(empty_a + 8 * s, p.ni);
      }
      for (int s = 0; s < BS; ++s) {
        mbar_init(full_b + 8 * s, 1);
        mbar_init(empty_b + 8 * s, p.ni);
      }
      for (int a = 0; a < p.nacc; ++a) {
        mbar_init(full_acc + 8 * a, p.ni);
        mbar_init(empty_acc + 8 * a, EW);
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
    // ---------------------------------------------------------------- activation producer (one thread)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      StItemIter it;
      it.init(p, blockIdx.x, gridDim.x);
      for (int item = blockIdx.x; item < p.items_per_nt; item += gridDim.x, it.next(p)) {
        const StItem c = it.get(p);
        const int ns = st_nsub_eff(p, c);
        StWalk wk;
        for (wk.init(p, c); wk.f <= wk.f_end; wk.next(p)) {
          if (!wk.used(p)) continue;
          const int fr = wk.f * p.tstep + p.toff;   // first source frame of the box (may lie in the temporal padding)
          const int si = (fr >= g.src[0].T && g.src[1].ptr != nullptr) ? 1 : 0;
          const int tl = fr - (si ? g.src[0].T : 0);
          for (int cb = 0; cb < p.ncb; ++cb) {
            // Two producers share the ring (UP): BOTH wait for every stage to drain and BOTH arrive on its full barrier (count 2),
            // also for the stages they do not fill.  A producer that merely skipped a stage could run (or fall) more than one ring
            // revolution away from the consumers, where the parity of an mbarrier wait aliases.
            ST_PROF_WAIT(prof_w0, mbar_wait(empty_a + 8 * s, ph ^ 1u));
            if (UP && si == 0) {   // filled by the interpolating warps
              mbar_arrive(full_a + 8 * s);
              if (++s == AS) { s = 0; ph ^= 1u; }
              continue;
            }
            const uint32_t dst = sA0 + (uint32_t)s * p.a_stage_bytes;
            if (p.halo == 3) {   // compact patch of the 4-channel clip: {2 pixels, PW pairs, PH rows}, rows 2*h0 - ph ..
              mbar_arrive_expect_tx(full_a + 8 * s, p.a_tx_sub);
              st_tma_load_5d(dst, &p.tmA[0], full_a + 8 * s, 0, c.tx * 8 * p.nsub, c.ty * p.th * 2 + p.par_h0[0], tl, c.b);
            } else if (p.halo == 2) {   // spatially strided conv: the rows a tile needs form par_sh lattices, one box per lattice
              mbar_arrive_expect_tx(full_a + 8 * s, (uint32_t)p.par_sh * p.a_tx_sub);
              for (int par = 0; par < p.par_sh; ++par)
                st_tma_load_5d(dst + (uint32_t)p.par_off[par], &p.tmA[0], full_a + 8 * s, cb * 64, c.tx * 8 * p.nsub + p.ew0,
                               c.ty * p.th * p.par_sh + p.par_h0[par], tl, c.b);
            } else if (p.halo) {
              mbar_arrive_expect_tx(full_a + 8 * s, p.a_tx_sub);
              st_tma_load_5d(dst, &p.tmA[si], full_a + 8 * s, cb * 64, c.tx * 8 * p.nsub + p.ew0, c.ty * p.th + p.eh0, tl, c.b);
            } else {
              mbar_arrive_expect_tx(full_a + 8 * s, (uint32_t)ns * p.a_tx_sub);
              for (int sub = 0; sub < ns; ++sub) {
                const int tif = c.tx * p.nsub + sub;
                const int ty = tif / p.tiles_w, tx = tif - ty * p.tiles_w;
                st_tma_load_5d(dst + (uint32_t)sub * p.sub_stride, &p.tmA[si], full_a + 8 * s, cb * 64, tx * p.tw, ty * p.th, tl, c.b);
              }
            }
            if (++s == AS) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- weight producer (one thread)
    if (lane == 0) {
      const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.d.w) + (size_t)nt * KB * p.b_bytes;
      if (p.wres) {
        mbar_arrive_expect_tx(wbar, (uint32_t)KB * p.b_bytes);
        for (int kb = 0; kb < KB; ++kb) bulk_copy_g2s(sB0 + (uint32_t)kb * p.b_bytes, wbase + (size_t)kb * p.b_bytes, p.b_bytes, wbar);
      } else {
        int s = 0;
        uint32_t ph = 0;
        StItemIter it;
        it.init(p, blockIdx.x, gridDim.x);
        for (int item = blockIdx.x; item < p.items_per_nt; item += gridDim.x, it.next(p)) {
          const StItem c = it.get(p);
          StWalk wk;
          for (wk.init(p, c); wk.f <= wk.f_end; wk.next(p)) {
            for (int cb = 0; cb < p.ncb; ++cb) {
              for (int tg = p.ntg - 1; tg >= 0; --tg) {
                if (wk.out_of(p, tg) < 0) continue;
                for (int j = p.tg_first[tg]; j < p.tg_first[tg + 1]; ++j) {
                  mbar_wait(empty_b + 8 * s, ph ^ 1u);
                  mbar_arrive_expect_tx(full_b + 8 * s, p.b_bytes);
                  bulk_copy_g2s(sB0 + (uint32_t)s * p.b_bytes, wbase + (size_t)(p.sp_kb[j] + cb) * p.b_bytes, p.b_bytes, full_b + 8 * s);
                  if (++s == BS) { s = 0; ph ^= 1u; }
                }
              }
            }
          }
        }
      }
    }
  } else if (warp >= FIRST_MMA_WARP) {
    if (warp - FIRST_MMA_WARP < p.ni) {
    // ---------------------------------------------------------------- MMA issuer: the warp runs the loop uniformly (so the
    // descriptor arithmetic can live in uniform registers) and one elected lane issues each tcgen05 instruction.  The issuing
    // warp is the critical resource (a 128xNx16 MMA only lasts max(N/2, 32+N/4) clocks), so this loop is kept as short as
    // possible: 32-bit descriptor words, per-tap constants straight from the parameter bank, no divisions.
    const uint32_t hi_common = (1u << 14) | (2u << 29);  // descriptor version 1, SWIZZLE_128B
    // halo == 3: the activation operand is read straight out of a compact, UN-swizzled patch (see conv_gemm_stream_win4)
    const uint32_t a_hi = (p.sbo >> 4) | (p.halo == 3 ? (1u << 14) : hi_common), b_hi = (1024u >> 4) | hi_common;
    const uint32_t sA_lo = ((sA0 & 0x3FFFFu) >> 4) | (1u << 16), sB_lo = ((sB0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_stage16 = p.a_stage_bytes >> 4, b16 = p.b_bytes >> 4, sub16 = p.sub_stride >> 4;
    const uint32_t nacc = (uint32_t)p.nacc, acc_set = (uint32_t)p.nsub * p.acc_stride;
    const uint32_t idesc = p.idesc, acc_stride = p.acc_stride;
    const int tg_last = p.ntg - 1;
    const bool wres = p.wres != 0;
    const int sub0 = warp - FIRST_MMA_WARP, ni = p.ni;
    const uint32_t sub16_0 = (uint32_t)sub0 * sub16, td_0 = (uint32_t)sub0 * acc_stride;
    const uint32_t sub16_step = (uint32_t)ni * sub16, td_step = (uint32_t)ni * acc_stride;
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    uint32_t a_lo_stage = sA_lo, b_lo_slot = sB_lo;
    uint32_t slot_done = 0, slot_start = 0, ph_start = 0, fresh = 0;
    if (wres) mbar_wait(wbar, 0);
    StItemIter it;
    it.init(p, blockIdx.x, gridDim.x);
    // One tap group at offset 0 (every (1,k,k) convolution and every temporal-halo tile: most launches of the model) with resident
    // weights and one sub-tile per issuer: output i reads exactly stage i, so the frame walk, the tap-group search and the slot
    // arithmetic of the general loop below collapse - and those ~150 scalar instructions per item were as long as the MMAs of a
    // 7-tap item.  Same barrier sequence as the general loop (the producers and the epilogue do not change).
    if (p.ntg == 1 && p.S == 1 && p.e_min == 0 && p.e_max == 0 && p.walk_Ts >= p.walk_Tr && p.nsub <= p.ni) {
      const int j0 = p.tg_first[0], j1 = p.tg_first[1];
      const uint32_t a_sub = sub16_0, td_sub = td_0;
      uint32_t slot = 0, ph_acc = 0;
      for (int item = blockIdx.x; item < p.items_per_nt; item += gridDim.x, it.next(p)) {
        const StItem c = it.get(p);
        const bool mine = sub0 < st_nsub_eff(p, c);
        for (int i = c.i0; i < c.i1; ++i) {
          ST_PROF_WAIT(prof_w1, mbar_wait(empty_acc + 8 * slot, ph_acc ^ 1u));
          tc_fence_after();
          const uint32_t td = tmem_base + slot * acc_set + td_sub;
          uint32_t keep = 0u;
          for (int cb = 0; cb < p.ncb; ++cb) {
            ST_PROF_WAIT(prof_w0, mbar_wait(full_a + 8 * sa, pha));
            tc_fence_after();
            const int rem = g.Cs - cb * 64;
            const int nk = rem >= 64 ? 4 : (rem + 15) >> 4;
            const uint32_t b_cb = sB_lo + (uint32_t)cb * b16;
            if (st_elect_one()) {
              if (mine) {
                st_issue_taps(s_tap, j0, j1, nk, td, a_lo_stage + a_sub, a_hi, b_cb, b_hi, idesc, keep, wres, full_b, empty_b, sb, phb,
                              b_lo_slot, sB_lo, b16, BS);
              } else if (!wres) {   // a sub-tile past the right edge: nothing to add, the weight ring still counts this issuer
                int s2 = sb;
                uint32_t ph2 = phb;
                for (int j = j0; j < j1; ++j) {
                  mbar_wait(full_b + 8 * s2, ph2);
                  umma_commit(empty_b + 8 * s2);
                  if (++s2 == BS) { s2 = 0; ph2 ^= 1u; }
                }
              }
              umma_commit(empty_a + 8 * sa);
              if (cb == p.ncb - 1) umma_commit(full_acc + 8 * slot);
            }
            if (!wres) {      // every lane advances its copy of the weight ring position
              sb += j1 - j0;
              while (sb >= BS) { sb -= BS; phb ^= 1u; }
              b_lo_slot = sB_lo + (uint32_t)sb * b16;
            }
            __syncwarp();
            keep = 1u;
            a_lo_stage += a_stage16;
            if (++sa == AS) { sa = 0; pha ^= 1u; a_lo_stage = sA_lo; }
          }
          if (++slot == nacc) { slot = 0; ph_acc ^= 1u; }
        }
      }
    } else
    for (int item = blockIdx.x; item < p.items_per_nt; item += gridDim.x, it.next(p)) {
      const StItem c = it.get(p);
      const int ns = st_nsub_eff(p, c);
      int i_done = c.i0, i_start = c.i0;  // next output frame to commit / to start (slot_start == slot_done here)
      StWalk wk;
      for (wk.init(p, c); wk.f <= wk.f_end; wk.next(p)) {
        if (wk.used(p)) {
          for (int cb = 0; cb < p.ncb; ++cb) {
            ST_PROF_WAIT(prof_w0, mbar_wait(full_a + 8 * sa, pha));
            tc_fence_after();
            const int rem = g.Cs - cb * 64;
            const int nk = rem >= 64 ? 4 : (rem + 15) >> 4;
            const uint32_t b_cb = sB_lo + (uint32_t)cb * b16;
            for (int tg = tg_last; tg >= 0; --tg) {  // descending source offset = ascending output frame
              const int i = wk.out_of(p, tg);
              if (i < 0) continue;
              while (i_start <= i) {  // first touch of an output frame: its accumulator slot must have been drained
                ST_PROF_WAIT(prof_w1, mbar_wait(empty_acc + 8 * slot_start, ph_start ^ 1u));
                tc_fence_after();
                fresh |= 1u << slot_start;
                ++i_start;
                if (++slot_start == nacc) { slot_start = 0; ph_start ^= 1u; }
              }
              uint32_t slot = slot_done + (uint32_t)(i - i_done);
              if (slot >= nacc) slot -= nacc;
              const uint32_t tacc = tmem_base + slot * acc_set;
              uint32_t keep = ((fresh >> slot) & 1u) ^ 1u;
              fresh &= ~(1u << slot);
              const int j_end = p.tg_first[tg + 1];
              if (sub0 + ni >= ns) {
                // lean path (one sub-tile per issuer): ONE elected lane issues the MMAs of all taps of the group back to back -
                // per tap one LDS, two adds and four tcgen05.mma (+ the weight ring's wait / commit when the weights stream)
                const int j_first = p.tg_first[tg];
                if (st_elect_one()) {
                  if (sub0 < ns)
                    st_issue_taps(s_tap, j_first, j_end, nk, tacc + td_0, a_lo_stage + sub16_0, a_hi, b_cb, b_hi, idesc, keep, wres, full_b,
                                  empty_b, sb, phb, b_lo_slot, sB_lo, b16, BS);
                  else if (!wres) {  // a sub-tile past the right edge: nothing to add, but the weight ring counts this issuer (it
                    int s2 = sb;     // waits like the others: an arrival ahead of the ring would land in the wrong phase)
                    uint32_t ph2 = phb;
                    for (int j = j_first; j < j_end; ++j) {
                      mbar_wait(full_b + 8 * s2, ph2);
                      umma_commit(empty_b + 8 * s2);
                      if (++s2 == BS) { s2 = 0; ph2 ^= 1u; }
                    }
                  }
                }
                __syncwarp();
                if (!wres) {      // every lane advances its copy of the ring position
                  const int n = j_end - j_first;
                  sb += n;
                  while (sb >= BS) { sb -= BS; phb ^= 1u; }
                  b_lo_slot = sB_lo + (uint32_t)sb * b16;
                }
                continue;
              }
              for (int j = p.tg_first[tg]; j < j_end; ++j) {
                const int2 tj = s_tap[j];
                uint32_t b_lo;
                if (wres) {
                  b_lo = b_cb + (uint32_t)tj.y;
                } else {
                  mbar_wait(full_b + 8 * sb, phb);
                  tc_fence_after();
                  b_lo = b_lo_slot;
                }
                uint32_t a_lo = a_lo_stage + (uint32_t)tj.x + sub16_0;
                uint32_t td = tacc + td_0;
                for (int sub = sub0; sub < ns; sub += ni, a_lo += sub16_step, td += td_step) {
                  if (st_elect_one()) {
                    st_umma(td, a_lo, a_hi, b_lo, b_hi, idesc, keep);
                    if (nk == 4) {
                      st_umma(td, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                      st_umma(td, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                      st_umma(td, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                    } else {
                      if (nk > 1) st_umma(td, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                      if (nk > 2) st_umma(td, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                    }
                  }
                }
                keep = 1u;
                if (!wres) {
                  if (st_elect_one()) umma_commit(empty_b + 8 * sb);
                  b_lo_slot += b16;
                  if (++sb == BS) { sb = 0; phb ^= 1u; b_lo_slot = sB_lo; }
                }
              }
            }
            if (st_elect_one()) umma_commit(empty_a + 8 * sa);
            a_lo_stage += a_stage16;
            if (++sa == AS) { sa = 0; pha ^= 1u; a_lo_stage = sA_lo; }
          }
        }
        // the output frame whose LAST source frame is f is complete (outputs complete in order)
        if (wk.out_of(p, tg_last) >= 0) {
          if (st_elect_one()) umma_commit(full_acc + 8 * slot_done);
          ++i_done;
          if (++slot_done == nacc) slot_done = 0;
        }
      }
      while (i_done < c.i1) {  // outputs whose last source frames lie beyond the end of the clip (temporal padding)
        if (st_elect_one()) umma_commit(full_acc + 8 * slot_done);
        ++i_done;
        if (++slot_done == nacc) slot_done = 0;
      }
    }
    }
  } else if (UP && warp >= FIRST_UP_WARP && warp < FIRST_MMA_WARP) {
    // ---------------------------------------------------------------- interpolating activation producer (128 threads, halo mode)
    const int itid = threadIdx.x - 32 * FIRST_UP_WARP;
    const vinet_src_t& s0 = g.src[0];
    const int lh = g.Hs >> 1, lw = g.Ws >> 1;
    const bool relu = (s0.xform & VINET_XF_RELU) != 0;
    const int64_t frame_elems = (int64_t)lh * lw * s0.ld;
    const int64_t clip_elems = s0.ldb ? s0.ldb : (int64_t)s0.T * frame_elems;
    const __nv_bfloat16* src0 = reinterpret_cast<const __nv_bfloat16*>(s0.ptr);
    int s = 0;
    uint32_t ph = 0;
    StItemIter it;
    it.init(p, blockIdx.x, gridDim.x);
    for (int item = blockIdx.x; item < p.items_per_nt; item += gridDim.x, it.next(p)) {
      const StItem c = it.get(p);
      StWalk wk;
      for (wk.init(p, c); wk.f <= wk.f_end; wk.next(p)) {
        if (!wk.used(p)) continue;
        const int fr = wk.f * p.tstep + p.toff;
        const int si = (fr >= s0.T && g.src[1].ptr != nullptr) ? 1 : 0;
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(empty_a + 8 * s, ph ^ 1u);   // every stage, see the TMA producer
          if (si == 0) {
            const int rem = g.Cs - cb * 64;
            const int nk = rem >= 64 ? 4 : (rem + 15) >> 4;
            up2_fill_box(sA0 + (uint32_t)s * p.a_stage_bytes, src0 + (int64_t)c.b * clip_elems + (int64_t)fr * frame_elems + cb * 64,
                         fr >= 0 && fr < s0.T, lh, lw, s0.ld, c.tx * 8 * p.nsub + p.ew0, c.ty * p.th + p.eh0, p.PW, p.PH, 2 * nk,
                         min(8, rem >> 3), relu, itid, ST_UP_THREADS);
            fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
          }
          // the group moves through the ring in lockstep (every stage), its leader arrives for it
          asm volatile("bar.sync 2, %0;" ::"n"(ST_UP_THREADS) : "memory");
          if (itid == 0) mbar_arrive(full_a + 8 * s);
          if (++s == AS) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: TMEM -> registers -> global
    const int q = warp & 3;
    const int half = (warp - FIRST_EPI_WARP) >> 2;   // 0 .. EW/4 - 1: column groups are dealt round-robin to the warps of a quarter
    const int row = q * 32 + lane;
    const int rt = row / p.pos, rrem = row - rt * p.pos;   // temporal-halo tiles stack tt frames of pos positions
    const int rh = rrem / p.tw, rw = rrem - rh * p.tw;
    const int BN = p.d.block_n;
    const int nlim = p.d.N - nt * BN;
    const uint32_t nacc = (uint32_t)p.nacc;
    const bool stats = p.d.stats != nullptr;
    const int etid = threadIdx.x - 32 * FIRST_EPI_WARP;   // 0..255 among the epilogue warps
    if (stats) {
      for (int i = etid; i < 2 * 256; i += ET) s_stats[i] = 0.0;
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
    }
    uint32_t slot = 0, ph = 0;
    StItemIter it;
    it.init(p, blockIdx.x, gridDim.x);
    for (int item = blockIdx.x; item < p.items_per_nt; item += gridDim.x, it.next(p)) {
      const StItem c = it.get(p);
      const int ns = st_nsub_eff(p, c);
      for (int i = c.i0; i < c.i1; ++i) {
        ST_PROF_WAIT(prof_w0, mbar_wait(full_acc + 8 * slot, ph));
        tc_fence_after();
        const int ti = i * p.tt + rt;
        const int t = ti * g.row_tstep + g.row_toff;
        // output row of this lane in sub-tile `sub` (nullptr: the row lies outside the problem)
        auto sub_row = [&](int sub, bool& accum) -> TO* {
          int h, w;
          if (p.halo) {
            h = c.ty * p.th + rh;
            w = (c.tx * p.nsub + sub) * 8 + rw;
          } else {
            const int tif = c.tx * p.nsub + sub;
            const int ty = tif / p.tiles_w;
            h = ty * p.th + rh;
            w = (tif - ty * p.tiles_w) * p.tw + rw;
          }
          const bool valid = row < p.pos * p.tt && ti < g.Tr && h < g.Hr && w < g.Wr;
          if (!valid) return nullptr;
          RowCoord rc;
          rc.b = c.b; rc.t = t; rc.h = h; rc.w = w;
          accum = (p.d.accumulate >> out_index(p.d, rc)) & 1;
          return out_row_ptr<TO>(p.d, rc) + nt * BN;
        };
        if (!stats) {
          for (int sub = 0; sub < ns; ++sub) {
            bool accum = false;
            TO* orow = sub_row(sub, accum);
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (slot * (uint32_t)p.nsub + (uint32_t)sub) * p.acc_stride;
            for (int gi = half; gi < BN / 16; gi += EW / 4) {
              uint32_t r[16];
              tmem_ld16(tacc + (uint32_t)(gi * 16), r);
              if (orow == nullptr) continue;
              const int c0 = gi * 16;
              epilogue_store16<TO, EPI>(p.d, orow + c0, r, nt * BN + c0, accum, nlim - c0);
            }
          }
        } else {
          // BatchNorm statistics of the raw output from the fp32 accumulators: column-group outermost, so that the lane-local
          // partial sums run over every sub-tile of the item before ONE cross-lane reduction per 16 columns
          for (int gi = half; gi < BN / 16; gi += EW / 4) {
            const int c0 = gi * 16;
            float sv[16], sq[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) sv[e] = sq[e] = 0.f;
            for (int sub = 0; sub < ns; ++sub) {
              bool accum = false;
              TO* orow = sub_row(sub, accum);
              const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (slot * (uint32_t)p.nsub + (uint32_t)sub) * p.acc_stride;
              uint32_t r[16];
              tmem_ld16(tacc + (uint32_t)c0, r);
              if (orow == nullptr) continue;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float x = __uint_as_float(r[e]);
                sv[e] += x;
                sq[e] = fmaf(x, x, sq[e]);
              }
              epilogue_store16<TO, EPI>(p.d, orow + c0, r, nt * BN + c0, accum, nlim - c0);
            }
            warp_colsum16(sv, lane);
            warp_colsum16(sq, lane);
            const int col = c0 + colsum16_col(lane);
            if (col < nlim) atomicAdd(&s_stats[(lane & 1) * 256 + col], (double)((lane & 1) ? sq[0] : sv[0]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_acc + 8 * slot);
        if (++slot == nacc) { slot = 0; ph ^= 1u; }
      }
    }
    if (stats) {   // per-CTA partials -> fp64 atomics on the layer's statistics
      asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
      for (int i = etid; i < 2 * BN; i += ET) {
        const int which = i / BN, col = i - which * BN;
        if (col < nlim) atomicAdd(p.d.stats + (size_t)which * p.d.N + nt * BN + col, s_stats[which * 256 + col]);
      }
    }
  }
#ifdef VINET_ST_PROF
  {
    const long long w2 = __shfl_sync(0xffffffffu, prof_w2, 0) + __shfl_xor_sync(0xffffffffu, prof_w2, 16);   // (whichever lane was elected)
    long long w2s = prof_w2;
    for (int o = 16; o > 0; o >>= 1) w2s += __shfl_xor_sync(0xffffffffu, w2s, o);
    (void)w2;
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == FIRST_MMA_WARP || warp == FIRST_MMA_WARP + 1 || warp == FIRST_EPI_WARP))
      printf("st_prof warp %d (%s): total %lld clk, wait0 %lld, wait1 %lld, in-mma-groups %lld | items/cta %d ncb %d nsub %d nacc %d a_stages %d wres %d run %d BN %d\n",
             warp, warp == 0 ? "tma: wait0 = free stage" : warp >= FIRST_MMA_WARP ? "mma: wait0 = stage data, wait1 = free accumulator" : "epilogue: wait0 = finished accumulator",
             clock64() - prof_t0, prof_w0, prof_w1, w2s, (p.items_per_nt + (int)gridDim.x - 1) / (int)gridDim.x, p.ncb, p.nsub, p.nacc, p.a_stages,
             p.wres, p.run, p.d.block_n);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
int make_tma_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, int64_t ld, int64_t ldh, int bw, int bh,
                 int esw, int esh, int bt, int64_t ldb = 0);
void pick_tma_box(int H, int W, int max_rows, int mult, bool full_tile_cost, int* bw_out, int* bh_out);
int tma_sm_count();

int g_stream_enable = 1;
int stream_enable_set(int v) { g_stream_enable = v; return 0; }
int g_st_issue_clk = 150;   // development knob (vinet_debug_set key 4): clocks one issuing warp needs per MMA in the cost model
int stream_issue_clk_set(int v) { g_st_issue_clk = v; return 0; }

namespace {

constexpr double ST_LOAD_BPC = 30.0;      // sustained L2->SM bytes per clock per SM with every SM pulling (tools/tma_bench.cu)
constexpr size_t ST_SMEM_BUDGET = 220 * 1024;   // + barriers + the 2 KB of per-CTA BatchNorm statistic partials <= 227 KB

// 128 x n x 16 MMA: tensor pipe vs shared-memory operand reads vs what one issuing warp sustains (~90 clk per MMA, measured)
double mma_clk16(int n, int issuers) { return std::max(std::max(n / 2.0, 32.0 + n / 4.0), (double)g_st_issue_clk / issuers); }

struct TapMap {
  int ntg, S, e_min, e_max, L;
  int tg_e[ST_MAX_TG];
  int tg_first[ST_MAX_TG + 1];
  int sp_tap[VINET_MAX_TAPS], sp_eh[VINET_MAX_TAPS], sp_ew[VINET_MAX_TAPS];
  int nsp;
  int eh0, eh1, ew0, ew1;
};

// temporal / spatial offsets of every tap: source = (i*S + e, h + eh, w + ew)
bool build_tap_map(const vinet_gather_t& g, TapMap* m) {
  int e[VINET_MAX_TAPS], eh[VINET_MAX_TAPS], ew[VINET_MAX_TAPS];
  bool use[VINET_MAX_TAPS];
  if (g.sh != 1 || g.sw != 1) return false;
  if (g.mode == VINET_GATHER_FPROP) {
    m->S = g.row_tstep * g.st;
  } else {
    if (g.row_tstep % g.st != 0) return false;
    m->S = g.row_tstep / g.st;
  }
  for (int t = 0; t < g.ntaps; ++t) {
    const int dt = g.tap[t][0], dh = g.tap[t][1], dw = g.tap[t][2];
    use[t] = true;
    if (g.mode == VINET_GATHER_FPROP) {
      e[t] = g.row_toff * g.st - g.pt + dt;
      eh[t] = dh - g.ph;
      ew[t] = dw - g.pw;
    } else {
      const int num = g.row_toff + g.pt - dt;
      if (((num % g.st) + g.st) % g.st != 0) { use[t] = false; continue; }
      e[t] = (num >= 0 ? num / g.st : -((-num) / g.st));
      eh[t] = g.ph - dh;
      ew[t] = g.pw - dw;
    }
  }
  // distinct temporal offsets, ascending
  m->ntg = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    if (!use[t]) continue;
    bool seen = false;
    for (int k = 0; k < m->ntg; ++k) seen |= (m->tg_e[k] == e[t]);
    if (seen) continue;
    if (m->ntg == ST_MAX_TG) return false;
    m->tg_e[m->ntg++] = e[t];
  }
  if (m->ntg == 0) return false;
  std::sort(m->tg_e, m->tg_e + m->ntg);
  m->e_min = m->tg_e[0];
  m->e_max = m->tg_e[m->ntg - 1];
  m->L = (m->e_max - m->e_min) / m->S + 1;
  m->nsp = 0;
  m->eh0 = m->ew0 = 1 << 20;
  m->eh1 = m->ew1 = -(1 << 20);
  for (int k = 0; k < m->ntg; ++k) {
    m->tg_first[k] = m->nsp;
    for (int t = 0; t < g.ntaps; ++t) {
      if (!use[t] || e[t] != m->tg_e[k]) continue;
      m->sp_tap[m->nsp] = t;
      m->sp_eh[m->nsp] = eh[t];
      m->sp_ew[m->nsp] = ew[t];
      ++m->nsp;
      m->eh0 = std::min(m->eh0, eh[t]); m->eh1 = std::max(m->eh1, eh[t]);
      m->ew0 = std::min(m->ew0, ew[t]); m->ew1 = std::max(m->ew1, ew[t]);
    }
  }
  m->tg_first[m->ntg] = m->nsp;
  // every output frame must read at least one real source frame (true for every layer of this model)
  for (int i = 0; i < g.Tr; ++i) {
    bool any = false;
    for (int k = 0; k < m->ntg; ++k) {
      const int f = i * m->S + m->tg_e[k];
      any |= (f >= 0 && f < g.Ts);
    }
    if (!any) return false;
  }
  return true;
}

struct Plan {
  bool ok = false;
  double cost = 1e300;
  int block_n = 0, n_tiles = 0, nsub = 0, nacc = 0, run = 0, a_stages = 0, b_slots = 0, wres = 0;
  int halo = 0, tw = 0, th = 0, items_w = 0, items_h = 0, tiles_w = 0, tpf = 0, PW = 0, PH = 0, tt = 1, PT = 1;
  uint32_t a_stage_bytes = 0, a_tx_sub = 0, sub_stride = 0, sbo = 0;
};

// Enumerate (N tiling, sub-tiles per item, output-frame run) and keep the cheapest under a max(tensor, L2->SM, epilogue) model.
Plan plan_stream(const vinet_conv_t& d, const TapMap& tm, int fixed_block_n, int fixed_n_tiles, int sms) {
  const vinet_gather_t& g = d.g;
  Plan best;
  const bool spatial = tm.eh0 != 0 || tm.eh1 != 0 || tm.ew0 != 0 || tm.ew1 != 0;
  if (!spatial && tm.ntg <= 1) return best;   // a 1x1x1 conv has nothing to re-use: conv_tma.cu's kernel is the right one
  if (spatial && g.Hr < 10) return best;      // 16-row halo tiles waste most of a 7-row map
  const int ncb = (g.Cs + 63) / 64;
  int nk_sum = 0;
  for (int cb = 0; cb < ncb; ++cb) {
    const int rem = g.Cs - cb * 64;
    nk_sum += rem >= 64 ? 4 : (rem + 15) >> 4;
  }
  const int n16 = (int)round_up(d.N, 16);
  int fbw = 0, fbh = 0;
  if (!spatial) pick_tma_box(g.Hr, g.Wr, TC_BM, 1, true, &fbw, &fbh);
  const int KB = g.ntaps * ncb;
  int last_bn = -1;
  for (int n_tiles = (int)cdiv(n16, 256); n_tiles <= (int)cdiv(n16, 16); ++n_tiles) {
    const int block_n = (int)round_up(cdiv(n16, n_tiles), 16);
    if (block_n == last_bn) continue;
    last_bn = block_n;
    if ((int)cdiv(n16, block_n) != n_tiles) continue;
    if (fixed_block_n && (block_n != fixed_block_n || n_tiles != fixed_n_tiles)) continue;
    if (n_tiles > sms) continue;
    const int acc_stride = (int)round_up(block_n, 32);
    const uint32_t b_bytes = (uint32_t)block_n * 128u;
    const int ctas = std::max(1, sms / n_tiles);
    // temporal-halo tiles (tt > 1): stride-1 temporal taps of a single source read ONE box of tt + kt - 1 frames
    // (temporally strided convs, S > 1: 8 positions per frame so that every 8-row group is one source frame, SBO = S frames)
    const bool thalo_ok = !spatial && tm.ntg > 1 && g.src[1].ptr == nullptr;
    for (int tt = 1; tt <= (thalo_ok ? 16 : 1); tt *= 2) {
      if (tt > 1 && tt / 2 >= g.Tr) break;
      if (!spatial && tm.S > 1 && tt != 16) continue;   // strided frame walks measured slower than per-tap boxes: temporal halo or nothing
      const int L = tt > 1 ? 1 : tm.L;
      const int walk_Tr = tt > 1 ? (int)cdiv(g.Tr, tt) : g.Tr;
      int tbw = fbw, tbh = fbh;
      if (tt > 1) pick_tma_box(g.Hr, g.Wr, TC_BM / tt, 8, true, &tbw, &tbh);
      const int pos = tbw * tbh;
      if (tt > 1 && pos * tt != TC_BM) continue;   // only full 128-row tiles keep the 8-row groups contiguous
      for (int nsub = 1; nsub <= 8; ++nsub) {
        const int slots = 512 / (nsub * acc_stride);
        if (slots < L) break;
        const int nacc = std::min(L + 1, slots);
        Plan c;
        c.block_n = block_n; c.n_tiles = n_tiles; c.nsub = nsub; c.nacc = nacc; c.tt = tt;
        double util;
        if (spatial) {
          c.halo = 1; c.tw = 8; c.th = 16;
          c.PW = 8 * nsub + (tm.ew1 - tm.ew0);
          c.PH = 16 + (tm.eh1 - tm.eh0);
          if (c.PW > 256 || c.PH > 256) break;
          c.items_w = (int)cdiv(g.Wr, 8 * nsub);
          c.items_h = (int)cdiv(g.Hr, 16);
          c.a_tx_sub = (uint32_t)(c.PW * c.PH * 128);
          c.a_stage_bytes = (uint32_t)round_up(c.a_tx_sub, 1024);
          c.sub_stride = 8 * 128;
          c.sbo = (uint32_t)c.PW * 128u;
          util = (double)cdiv(g.Wr, 8) / (double)(c.items_w * nsub);   // sub-tiles past the right edge are skipped
        } else {
          c.halo = 0; c.tw = tbw; c.th = tbh;
          c.PT = tt > 1 ? (tt - 1) * tm.S + (tm.e_max - tm.e_min) + 1 : 1;
          if (c.PT > 256) continue;
          c.tiles_w = (int)cdiv(g.Wr, tbw);
          c.tpf = c.tiles_w * (int)cdiv(g.Hr, tbh);
          if (nsub > c.tpf) break;
          c.items_w = (int)cdiv(c.tpf, nsub);
          c.items_h = 1;
          c.a_tx_sub = (uint32_t)(pos * c.PT * 128);
          c.sub_stride = (uint32_t)round_up(std::max<uint32_t>(c.a_tx_sub, TC_A_BYTES), 1024);
          c.a_stage_bytes = (uint32_t)nsub * c.sub_stride;
          c.sbo = (tt > 1 && tm.S > 1) ? (uint32_t)(tm.S * pos * 128) : 1024u;
          util = (double)c.tpf / (double)(c.items_w * nsub);
        }
        // shared memory: resident weights when they fit beside two activation stages, else a ring of weight blocks
        const size_t bar_bytes = 1024 + 1536;
        const size_t wbytes = (size_t)KB * b_bytes;
        if (wbytes + 2 * (size_t)c.a_stage_bytes + bar_bytes <= ST_SMEM_BUDGET) {
          c.wres = 1;
          c.b_slots = 1;
          c.a_stages = (int)std::min<size_t>(6, (ST_SMEM_BUDGET - bar_bytes - wbytes) / c.a_stage_bytes);
        } else {
          c.wres = 0;
          if (2 * (size_t)c.a_stage_bytes + 3 * (size_t)b_bytes + bar_bytes > ST_SMEM_BUDGET) continue;
          c.a_stages = 2;
          size_t left = ST_SMEM_BUDGET - bar_bytes - 2 * (size_t)c.a_stage_bytes;
          c.b_slots = (int)std::min<size_t>(12, left / b_bytes);
          // spend what is left beyond 6 weight slots on a third activation stage
          if (c.b_slots > 6 && left - 6 * (size_t)b_bytes >= c.a_stage_bytes) {
            c.a_stages = 3;
            c.b_slots = (int)std::min<size_t>(12, (left - c.a_stage_bytes) / b_bytes);
          }
        }
        const int issuers = std::min(ST_MAX_ISSUERS, nsub);
        for (int run = 1; run <= walk_Tr; run = (run * 2 > walk_Tr && run != walk_Tr) ? walk_Tr : run * 2) {
          const int nruns = (int)cdiv(walk_Tr, run);
          // source stages walked per item and taps issued per output (a temporal-halo stage holds every tap)
          const double frames = tt > 1 ? run : (double)(run - 1) * tm.S + (tm.e_max - tm.e_min) + 1;
          const double used = tt > 1 ? run : std::min(frames, (double)run * tm.ntg);
          const double taps_per_out = (double)tm.nsp;
          const double mma = run * taps_per_out * nk_sum * nsub * util * mma_clk16(block_n, issuers);
          const double bytes = used * ncb * c.a_tx_sub * (c.halo ? 1.0 : nsub * util) +
                               (c.wres ? 0.0 : run * taps_per_out * ncb * (double)b_bytes);
          const double load = bytes / ST_LOAD_BPC;
          const double epi = run * nsub * util * 128.0 * block_n * (d.accumulate ? 6.0 : 2.0) / 32.0 + 300.0 * run;
          const double t_item = (nacc > L ? std::max(std::max(mma, load), epi) : std::max(mma, load) + epi) + 1500.0;
          const int64_t items = (int64_t)g.B * nruns * c.items_w * c.items_h;
          const double total = (double)cdiv(items, ctas) * t_item + (c.wres ? wbytes / ST_LOAD_BPC : 0.0);
          if (total < best.cost) {
            best = c;
            best.ok = true;
            best.cost = total;
            best.run = run;
          }
          if (run == walk_Tr) break;
        }
      }
    }
  }
  return best;
}

bool stream_up2(const vinet_gather_t& g) {
  return g.src[0].ptr != nullptr && (g.src[0].xform & ~VINET_XF_RELU) == VINET_XF_UP2;
}

bool stream_eligible(const vinet_conv_t& d) {
  const vinet_gather_t& g = d.g;
  if (!g_stream_enable) return false;
  if (g.dtype != VINET_BF16 || g.Cs % 8 != 0 || d.N % 8 != 0) return false;
  for (int i = 0; i < 2; ++i) {
    const vinet_src_t& s = g.src[i];
    if (s.ptr == nullptr) continue;
    if (i == 0 && stream_up2(g)) {   // low-res source read through the fused 2x up-sampling: dense low-res rows only
      if (g.mode != VINET_GATHER_FPROP || (g.Hs & 1) || (g.Ws & 1) || s.ldh != 0 || s.ld < g.Cs) return false;
      continue;
    }
    if (s.xform != VINET_XF_IDENT) return false;
    if (s.ldh != 0 && s.ldh != (int64_t)g.Ws * s.ld) return false;   // sliding-window (WIN8) sources stay with conv_tma.cu
    if (s.ld < g.Cs) return false;
  }
  return true;
}

// the interpolating producer works on 2x2 quads: the halo box must start at odd hi-res coordinates and have even extents
// (true for the pad-1 3x3 halos of this model's decoder; anything else is materialised by the caller)
bool up2_plan_ok(const Plan& pl, const TapMap& tm) {
  return pl.ok && pl.halo == 1 && (tm.ew0 & 1) && (tm.eh0 & 1) && (pl.PW % 2 == 0) && (pl.PH % 2 == 0) && (pl.th % 2 == 0);
}

}  // namespace

// N tiling the streaming kernel wants for this convolution (0 = not handled by it)
int conv_stream_tiling(const vinet_conv_t* d, int* block_n, int* n_tiles) {
  *block_n = *n_tiles = 0;
  if (!stream_eligible(*d)) return 0;
  TapMap tm;
  if (!build_tap_map(d->g, &tm)) return 0;
  const Plan pl = plan_stream(*d, tm, 0, 0, tma_sm_count());
  if (getenv("VINET_STREAM_DEBUG"))
    fprintf(stderr, "stream plan: mode %d rows %dx%dx%dx%d Cs %d N %d taps %d S %d L %d -> ok %d block_n %d x %d nsub %d nacc %d run %d "
            "halo %d tt %d box %dx%d a_stages %d b_slots %d wres %d est %.0f kclk\n", d->g.mode, d->g.B, d->g.Tr, d->g.Hr, d->g.Wr, d->g.Cs,
            d->N, d->g.ntaps, tm.S, tm.L, (int)pl.ok, pl.block_n, pl.n_tiles, pl.nsub, pl.nacc, pl.run, pl.halo, pl.tt,
            pl.halo ? pl.PW : pl.tw, pl.halo ? pl.PH : pl.th, pl.a_stages, pl.b_slots, pl.wres, pl.cost / 1e3);
  if (!pl.ok) return 0;
  if (stream_up2(d->g) && !up2_plan_ok(pl, tm)) return 0;
  *block_n = pl.block_n;
  *n_tiles = pl.n_tiles;
  return 1;
}

// 1 when conv_gemm_stream serves this FPROP convolution with a VINET_XF_UP2 source 0 (host only)
int conv_stream_up2_ok(const vinet_conv_t* d) {
  int bn = 0, nt = 0;
  return stream_up2(d->g) && conv_stream_tiling(d, &bn, &nt);
}

static int stream_launch(StreamParams& p, const vinet_conv_t* d, int sms, cudaStream_t stream, bool up = false) {
  const int nb_slots = p.wres ? d->k_blocks : p.b_slots;
  size_t smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)nb_slots * p.b_bytes +
                8 * (size_t)(2 * p.a_stages + 2 * p.b_slots + 2 * p.nacc + 1) + 64 + 8 * VINET_MAX_TAPS + 4 * ST_STATS_FLOATS;
  if (smem > 227 * 1024) {
    set_error("conv_gemm_stream: %zu bytes of shared memory", smem);
    return -1;
  }
  smem = std::max<size_t>(smem, 120 * 1024);   // one CTA per SM: two co-resident CTAs would fight over the 512 TMEM columns
  const int ctas = (int)std::min<int64_t>(p.items_per_nt, std::max(1, sms / d->n_tiles));
  dim3 grid((unsigned)ctas, (unsigned)d->n_tiles);
#define LAUNCH_STREAM(TO)                                                                                              \
  do {                                                                                                                 \
    auto kern = up ? (conv_has_epilogue(*d) ? conv_stream_kernel<TO, true, true> : conv_stream_kernel<TO, false, true>)   \
                   : (conv_has_epilogue(*d) ? conv_stream_kernel<TO, true, false> : conv_stream_kernel<TO, false, false>); \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                \
    kern<<<grid, up ? ST_UP_TOTAL : ST_THREADS, smem, stream>>>(p);                                                    \
  } while (0)
  VINET_DISPATCH_DTYPE(d->out_dtype, TO, LAUNCH_STREAM(TO));
#undef LAUNCH_STREAM
  note_kernel("conv_stream_kernel");
  if (up) g_up2_launches.fetch_add(1, std::memory_order_relaxed);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("conv_gemm_stream: launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 1;
}

// Row-strided convolution over sliding-window rows (the stem conv_s: (1,7,7)/(1,2,2) on the WIN8 input, one K block per kernel
// row dh).  Output row h reads source rows h*sh - ph + dh: for a 16-row tile these form `sh` lattices (row parities for sh = 2),
// each fetched as ONE box with a row element-stride of sh; tap dh is then a plain row offset inside its lattice's box, i.e. a
// canonical (SBO = row pitch) descriptor.  7 boxes of 16 rows per tile become 2 boxes of 19: 3x less L2->SM traffic.
static int conv_gemm_stream_strided(const vinet_conv_t* d, cudaStream_t stream) {
  const vinet_gather_t& g = d->g;
  if (!g_stream_enable || g.mode != VINET_GATHER_FPROP || g.dtype != VINET_BF16) return 0;
  if (g.src[1].ptr != nullptr || g.src[0].xform != VINET_XF_IDENT) return 0;
  if (g.Cs != 64 || g.sw != 1 || g.sh < 2 || g.sh > 2 || g.st != 1 || g.row_tstep != 1 || g.row_toff != 0 || g.pt != 0) return 0;
  if (d->n_tiles != 1 || d->block_n > 128 || d->N % 8 != 0 || g.ntaps > 16 || d->k_blocks != g.ntaps) return 0;
  if (g.Hr < 10) return 0;
  int q[VINET_MAX_TAPS], par[VINET_MAX_TAPS], qmin[2] = {1 << 20, 1 << 20}, qmax[2] = {-(1 << 20), -(1 << 20)};
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.tap[t][0] != 0 || g.tap[t][2] != 0) return 0;
    const int o = g.tap[t][1] - g.ph;
    par[t] = ((o % 2) + 2) % 2;
    q[t] = (o - par[t]) / 2;
    qmin[par[t]] = std::min(qmin[par[t]], q[t]);
    qmax[par[t]] = std::max(qmax[par[t]], q[t]);
  }
  if (qmax[0] < qmin[0] || qmax[1] < qmin[1]) return 0;
  const int PHm = 16 + std::max(qmax[0] - qmin[0], qmax[1] - qmin[1]);
  const int sms = tma_sm_count();
  StreamParams p;
  p.d = *d;
  p.ncb = 1;
  p.acc_stride = (uint32_t)round_up(d->block_n, 32);
  p.b_bytes = (uint32_t)d->block_n * 128u;
  const size_t wbytes = (size_t)d->k_blocks * p.b_bytes;
  int nsub = 0;
  for (int ns = 4; ns >= 1; --ns) {   // widest item whose two stages fit beside the resident weights and whose accumulators double-buffer
    const size_t stage = (size_t)2 * PHm * 8 * ns * 128;
    if (2 * stage + wbytes + 4096 <= ST_SMEM_BUDGET && 2 * ns * (int)p.acc_stride <= 512) { nsub = ns; break; }
  }
  if (nsub == 0) return 0;
  p.halo = 2; p.nsub = nsub; p.tw = 8; p.th = 16;
  p.items_w = (int)cdiv(g.Wr, 8 * nsub); p.items_h = (int)cdiv(g.Hr, 16); p.tiles_w = 0; p.tpf = 0;
  p.PW = 8 * nsub; p.PH = PHm; p.ew0 = 0; p.eh0 = 0;
  p.par_sh = 2;
  p.a_tx_sub = (uint32_t)(PHm * 8 * nsub * 128);
  for (int k = 0; k < 2; ++k) {
    p.par_h0[k] = 2 * qmin[k] + k;               // source row of the box's first row, relative to 2 * (tile's first output row)
    p.par_off[k] = k * (int)p.a_tx_sub;
  }
  p.a_stage_bytes = 2 * p.a_tx_sub;
  p.tt = 1; p.pos = 128; p.tstep = 1; p.toff = 0; p.walk_Tr = g.Tr; p.walk_Ts = g.Ts;
  p.run = 1; p.nruns = g.Tr; p.S = 1; p.ntg = 1; p.e_min = 0; p.e_max = 0;
  for (int k = 0; k < ST_MAX_TG; ++k) { p.tg_er[k] = 0; p.tg_eq[k] = 0; }
  for (int k = 0; k <= ST_MAX_TG; ++k) p.tg_first[k] = k == 0 ? 0 : g.ntaps;
  for (int j = 0; j < VINET_MAX_TAPS; ++j) {
    p.sp_aoff[j] = j < g.ntaps ? p.par_off[par[j]] + (q[j] - qmin[par[j]]) * p.PW * 128 : 0;
    p.sp_kb[j] = j < g.ntaps ? j : 0;
    p.sp_aoff16[j] = p.sp_aoff[j] >> 4;
    p.sp_boff16[j] = (int32_t)(((int64_t)p.sp_kb[j] * d->block_n * 128) >> 4);
  }
  p.nacc = 2; p.a_stages = 2; p.b_slots = 1; p.wres = 1; p.ni = std::min(ST_MAX_ISSUERS, nsub);
  const int64_t items = (int64_t)g.B * g.Tr * p.items_w * p.items_h;
  if (items >= (1ll << 31)) return 0;
  p.items_per_nt = (int)items;
  p.tmem_cols = tmem_cols_for(p.nacc * nsub * (int)p.acc_stride);
  p.idesc = make_idesc(TC_BM, d->block_n, 0, 0);
  p.sub_stride = 8 * 128;
  p.sbo = (uint32_t)p.PW * 128u;
  const vinet_src_t& s = g.src[0];
  if (make_tma_map(&p.tmA[0], s.ptr, g.Cs, g.Ws, g.Hs, s.T, g.B, s.ld, s.ldh, p.PW, PHm, 1, 2, 1, s.ldb)) return -1;
  p.tmA[1] = p.tmA[0];
  return stream_launch(p, d, sms, stream);
}

// Stem conv_s on the 4-CHANNEL clip (VINET_KLAYOUT_WIN4: (1,kh,kw<=7)/(1,2,2), Cin <= 4): no im2col, no window expansion at all.
// A pixel is 8 bytes, so the kw-wide windows of consecutive output pixels (stride 2) start 16 bytes apart - exactly the row
// pitch of an UN-swizzled K-major core matrix (8 rows x 16 bytes, rows 16 bytes apart).  The tensor core therefore reads the
// A operand of kernel row dh straight out of a COMPACT patch of the clip {PH = 30 + kh source rows, 8*nsub + 3 pixel pairs}:
//   descriptor start = patch + dh*pitch + sub*128 + kstep*32, LBO (next 8 K elements = next 2 pixels) = 16 bytes: the core
//   matrices of one MMA overlap in shared memory; SBO (next 8 output pixels = next output row) = 2*pitch.
// K per kernel row = 8 pixels x 4 channels = 32 (two K=16 MMAs; the 8th pixel and the 4th channel meet zero weights): 14 MMAs
// per 128 outputs instead of 28, and 21 KB of L2->SM traffic per 512 outputs instead of 155 KB of 4x-overlapping window rows
// (the halo == 2 mode above), which is what bound this layer.
int make_tma_map_pairs(CUtensorMap* m, const void* ptr, int Wp, int H, int T, int B, int bw, int bh, int64_t ldb);

static bool win4_geometry(const vinet_conv_t* d, int* nsub_out, int* PH_out) {
  const vinet_gather_t& g = d->g;
  if (!g_stream_enable || g.mode != VINET_GATHER_FPROP || g.dtype != VINET_BF16) return false;
  const vinet_src_t& s = g.src[0];
  if (g.src[1].ptr != nullptr || s.xform != VINET_XF_IDENT || s.ptr == nullptr) return false;
  if (g.Cs != 32 || s.ld != 8 || g.sw != 1 || g.sh != 2 || g.st != 1 || g.row_tstep != 1 || g.row_toff != 0 || g.pt != 0) return false;
  if (s.ldh <= 0 || s.ldh % 8 != 0 || (s.ldh / 4) < 2 * (int64_t)(g.Wr - 1) + 8) return false;   // even padded width, windows inside the row
  if (d->N % 8 != 0 || d->N > 128 || g.ntaps > 16 || g.Hr < 10) return false;
  for (int t = 0; t < g.ntaps; ++t)
    if (g.tap[t][0] != 0 || g.tap[t][1] != t || g.tap[t][2] != 0) return false;
  const int PH = 30 + g.ntaps;
  const int acc = (int)round_up(round_up(d->N, 16), 32);
  const size_t wbytes = (size_t)g.ntaps * round_up(d->N, 16) * 128;
  int nsub = 0;
  for (int ns = 4; ns >= 1; --ns) {
    const size_t stage = round_up((size_t)PH * (8 * ns + 3) * 16, 1024);
    if (2 * ns * acc <= 512 && 3 * stage + wbytes + 4096 <= ST_SMEM_BUDGET) { nsub = ns; break; }
  }
  if (nsub == 0) return false;
  *nsub_out = nsub;
  *PH_out = PH;
  return true;
}

int conv_stream_win4_ok(const vinet_conv_t* d) {
  int nsub, PH;
  return win4_geometry(d, &nsub, &PH) ? 1 : 0;
}

static int conv_gemm_stream_win4(const vinet_conv_t* d, cudaStream_t stream) {
  const vinet_gather_t& g = d->g;
  int nsub, PH;
  if (!win4_geometry(d, &nsub, &PH)) return 0;
  if (d->n_tiles != 1 || d->block_n != (int)round_up(d->N, 16) || d->k_blocks != g.ntaps) return 0;
  const int sms = tma_sm_count();
  StreamParams p;
  p.d = *d;
  p.ncb = 1;
  p.acc_stride = (uint32_t)round_up(d->block_n, 32);
  p.b_bytes = (uint32_t)d->block_n * 128u;
  const size_t wbytes = (size_t)d->k_blocks * p.b_bytes;
  const int PWp = 8 * nsub + 3;                       // pixel pairs per patch row
  const uint32_t pitch = (uint32_t)PWp * 16u;
  p.halo = 3; p.nsub = nsub; p.tw = 8; p.th = 16;
  p.items_w = (int)cdiv(g.Wr, 8 * nsub); p.items_h = (int)cdiv(g.Hr, 16); p.tiles_w = 0; p.tpf = 0;
  p.PW = PWp; p.PH = PH; p.ew0 = 0; p.eh0 = 0;
  p.par_sh = 1; p.par_h0[0] = -g.ph; p.par_h0[1] = 0; p.par_off[0] = p.par_off[1] = 0;
  p.a_tx_sub = (uint32_t)PH * pitch;
  p.a_stage_bytes = (uint32_t)round_up(p.a_tx_sub, 1024);   // the SWIZZLE_128B weight blocks behind the ring need 1024-byte alignment
  p.tt = 1; p.pos = 128; p.tstep = 1; p.toff = 0; p.walk_Tr = g.Tr; p.walk_Ts = g.Ts;
  p.run = 1; p.nruns = g.Tr; p.S = 1; p.ntg = 1; p.e_min = 0; p.e_max = 0;
  for (int k = 0; k < ST_MAX_TG; ++k) { p.tg_er[k] = 0; p.tg_eq[k] = 0; }
  for (int k = 0; k <= ST_MAX_TG; ++k) p.tg_first[k] = k == 0 ? 0 : g.ntaps;
  for (int j = 0; j < VINET_MAX_TAPS; ++j) {
    p.sp_aoff[j] = j < g.ntaps ? j * (int)pitch : 0;      // kernel row dh = patch row dh (+ 2 per output row: SBO)
    p.sp_kb[j] = j < g.ntaps ? j : 0;
    p.sp_aoff16[j] = p.sp_aoff[j] >> 4;
    p.sp_boff16[j] = (int32_t)(((int64_t)p.sp_kb[j] * d->block_n * 128) >> 4);
  }
  p.nacc = 2; p.b_slots = 1; p.wres = 1; p.ni = std::min(ST_MAX_ISSUERS, nsub);
  p.a_stages = (int)std::min<size_t>(6, (ST_SMEM_BUDGET - 4096 - wbytes) / p.a_stage_bytes);
  const int64_t items = (int64_t)g.B * g.Tr * p.items_w * p.items_h;
  if (items >= (1ll << 31)) return 0;
  p.items_per_nt = (int)items;
  p.tmem_cols = tmem_cols_for(p.nacc * nsub * (int)p.acc_stride);
  p.idesc = make_idesc(TC_BM, d->block_n, 0, 0);
  p.sub_stride = 8 * 16;                              // 8 output pixels = 16 source pixels of 8 bytes
  p.sbo = 2u * pitch;                                 // next output row = two source rows down
  const vinet_src_t& s = g.src[0];
  if (make_tma_map_pairs(&p.tmA[0], s.ptr, (int)(s.ldh / 4), g.Hs, s.T, g.B, PWp, PH, s.ldb)) return -1;
  p.tmA[1] = p.tmA[0];
  return stream_launch(p, d, sms, stream);
}

// returns 1 when the launch was handled here, 0 when the caller should use its own kernel, <0 on error
int conv_gemm_stream(const vinet_conv_t* d, cudaStream_t stream) {
  const vinet_gather_t& g = d->g;
  if (g.sh != 1 && g.src[0].ptr != nullptr && g.src[0].ld < g.Cs)
    return g.Cs == 32 ? conv_gemm_stream_win4(d, stream) : conv_gemm_stream_strided(d, stream);
  if (!stream_eligible(*d)) return 0;
  TapMap tm;
  if (!build_tap_map(g, &tm)) return 0;
  const int sms = tma_sm_count();
  const Plan pl = plan_stream(*d, tm, d->block_n, d->n_tiles, sms);
  if (!pl.ok) return 0;
  const bool up = stream_up2(g);
  if (up && !up2_plan_ok(pl, tm)) return 0;
  StreamParams p;
  p.d = *d;
  p.ncb = (g.Cs + 63) / 64;
  if (d->k_blocks != g.ntaps * p.ncb) {
    set_error("conv_gemm_stream: k_blocks %d != ntaps*ceil(Cs/64) = %d (TAP64 weights expected)", d->k_blocks, g.ntaps * p.ncb);
    return -1;
  }
  p.halo = pl.halo; p.nsub = pl.nsub; p.tw = pl.tw; p.th = pl.th;
  p.items_w = pl.items_w; p.items_h = pl.items_h; p.tiles_w = pl.tiles_w; p.tpf = pl.tpf;
  p.PW = pl.PW; p.PH = pl.PH; p.ew0 = tm.ew0; p.eh0 = tm.eh0;
  const bool thalo = pl.tt > 1;
  p.tt = pl.tt; p.pos = pl.tw * pl.th;
  p.tstep = thalo ? pl.tt * tm.S : 1; p.toff = thalo ? tm.e_min : 0;
  p.walk_Tr = thalo ? (int)cdiv(g.Tr, pl.tt) : g.Tr;
  p.walk_Ts = thalo ? p.walk_Tr : g.Ts;
  p.run = pl.run; p.nruns = (int)cdiv(p.walk_Tr, pl.run);
  // temporal-halo tiles: every tap reads the same box, i.e. ONE tap group at offset 0 in units of whole tiles
  p.S = thalo ? 1 : tm.S; p.ntg = thalo ? 1 : tm.ntg;
  p.e_min = thalo ? 0 : tm.e_min; p.e_max = thalo ? 0 : tm.e_max;
  for (int k = 0; k < ST_MAX_TG; ++k) {
    const int e = (k < tm.ntg && !thalo) ? tm.tg_e[k] : 0;
    p.tg_er[k] = ((e % p.S) + p.S) % p.S;
    p.tg_eq[k] = (e - p.tg_er[k]) / p.S;
  }
  for (int k = 0; k <= ST_MAX_TG; ++k) p.tg_first[k] = thalo ? (k == 0 ? 0 : tm.nsp) : (k <= tm.ntg ? tm.tg_first[k] : tm.nsp);
  for (int j = 0; j < VINET_MAX_TAPS; ++j) {
    if (j < tm.nsp) {
      int grp = 0;
      while (j >= tm.tg_first[grp + 1]) ++grp;
      p.sp_aoff[j] = pl.halo ? ((tm.sp_eh[j] - tm.eh0) * pl.PW + (tm.sp_ew[j] - tm.ew0)) * 128
                             : (thalo ? (tm.tg_e[grp] - tm.e_min) * p.pos * 128 : 0);
      p.sp_kb[j] = tm.sp_tap[j] * p.ncb;
    } else {
      p.sp_aoff[j] = p.sp_kb[j] = 0;
    }
    p.sp_aoff16[j] = p.sp_aoff[j] >> 4;
    p.sp_boff16[j] = (int32_t)(((int64_t)p.sp_kb[j] * d->block_n * 128) >> 4);
  }
  p.nacc = pl.nacc; p.a_stages = pl.a_stages; p.b_slots = pl.b_slots; p.wres = pl.wres;
  p.ni = std::min(ST_MAX_ISSUERS, pl.nsub);
  const int64_t items = (int64_t)g.B * p.nruns * p.items_w * p.items_h;
  if (items >= (1ll << 31)) return 0;
  p.items_per_nt = (int)items;
  p.acc_stride = (uint32_t)round_up(d->block_n, 32);
  p.tmem_cols = tmem_cols_for(pl.nacc * pl.nsub * (int)p.acc_stride);
  p.idesc = make_idesc(TC_BM, d->block_n, 0, 0);
  p.a_stage_bytes = pl.a_stage_bytes; p.a_tx_sub = pl.a_tx_sub; p.sub_stride = pl.sub_stride; p.sbo = pl.sbo;
  p.b_bytes = (uint32_t)d->block_n * 128u;
  for (int i = 0; i < 2; ++i) {
    const vinet_src_t& s = g.src[(i == 1 && g.src[1].ptr == nullptr) ? 0 : i];
    const int bw = pl.halo ? pl.PW : pl.tw, bh = pl.halo ? pl.PH : pl.th;
    // an up-sampled source is never fetched by TMA: its map describes the low-res tensor (valid, unused)
    const int dv = (up && &s == &g.src[0]) ? 2 : 1;
    if (make_tma_map(&p.tmA[i], s.ptr, g.Cs, g.Ws / dv, g.Hs / dv, s.T, g.B, s.ld, s.ldh, bw, bh, 1, 1, pl.PT, s.ldb)) return -1;
  }
  p.par_sh = 1; p.par_h0[0] = p.par_h0[1] = 0; p.par_off[0] = p.par_off[1] = 0;
  return stream_launch(p, d, sms, stream, up);
}

}  // namespace vinet
