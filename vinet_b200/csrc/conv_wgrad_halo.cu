// Weight gradient of spatial (kh x kw > 1, stride 1) convolutions with ONE halo tile per position chunk (sm_100a).
//
//   dW[(dt,dh,dw,c), n] += sum over positions  A[b, t*st-pt+dt, h-ph+dh, w-pw+dw, c] * dY[b,t,h,w,n]
//
// conv_tma.cu's wgrad kernel fetches one activation box per (tap, 64-channel block) unit and keeps at most 4 units per CTA,
// so a 3x3 conv pulls every activation through L2->SM nine times and every dY tile once per 4 units: those launches sit on
// the L2->SM bandwidth roof (convtsp4.0: 360 TFLOP/s).  Here a CTA owns ALL kh*kw spatial taps of one (temporal tap,
// 64-channel block) group against one N block of dY:
//   * per chunk of 8 x 16 output positions it fetches ONE halo box {64 ch, 8+kw-1, 16+kh-1} and the dY box(es);
//   * the A operand of tap (dh,dw) is the MN-major SWIZZLE_128B descriptor {start = box + (dh*PW + dw)*128, SBO = PW*128}
//     (8-position atoms are rows of the tile); two taps are stacked into one M=128 accumulator through LBO = the byte
//     distance between their windows (profiles/r1_umma_shift_test.txt: shifted MN-major descriptors read what they should);
//   * ceil(kh*kw/2) fp32 accumulators of block_n <= 96 columns live in TMEM for the whole CTA lifetime (split over position
//     chunks across CTAs, fp32 red.add into the packed TAP64 gradient at the end);
//   * up to 4 warps issue the MMAs (one accumulator each; a single warp cannot feed the tensor pipe with N <= 96 MMAs).
// Roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator, warps 2..5 MMA issuers, warps 6..9 epilogue.
// UP variant (576 threads; source 0 has VINET_XF_UP2: the activation operand is the decoder's relu -> 2x bilinear up-sampling of a
// low-res tensor, model.py:254): warps 10..17 interpolate the halo box of every chunk that reads source 0 straight into the
// stage (up2.cuh, same layout and values as a TMA load of the materialised tensor); dY and source-1 boxes still arrive by TMA.
// The stage's full barrier then counts two arrivals: the TMA thread's expect_tx and the interpolating group's.
#include <cuda.h>

#include <algorithm>

#include "tc_ptx.cuh"
#include "up2.cuh"

namespace vinet {

constexpr int WH_THREADS = 320;
constexpr int WH_UP_THREADS = 256;
constexpr int WH_MAX_ISSUERS = 4;
constexpr int WH_MAX_SP = 16;       // spatial taps per group (<= 8 accumulators)
constexpr uint32_t WH_UNIT = 16384;  // one 8 x 16 x 64-channel box

int make_tma_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int B, int64_t ld, int64_t ldh, int bw, int bh,
                 int esw, int esh, int bt, int64_t ldb = 0);
int tma_sm_count();
extern int g_stream_enable;

struct WgHaloParams {
  CUtensorMap tmA[2];
  CUtensorMap tmDy;
  vinet_wgrad_t d;
  int32_t ncb, nsp, naccs, nblk, block_n, splits, stages, ni, tiles_w, tiles_h, nT, temporal;
  int32_t cw, ch, aw0, ah0, at_step, at0, dyt_step;   // chunk (tw, th, tr) -> TMA coordinates of the activation / dY boxes
  int32_t nbox, ah_mul, box_h0[2], box_off[2];        // row-strided convs: one activation box per source-row lattice
  int32_t abw, abh;                                   // extent of the activation box (positions)
  int32_t uoff16[WH_MAX_SP];                          // window of unit j inside the activation box, in 16-byte units
  uint32_t a_sbo, kstep_a16;                          // stride between 8-position atoms; 16 positions in 16-byte units
  uint32_t acc_cols, tmem_cols, idesc, a_bytes, a_tx, dy_unit, stage_bytes;
};

__device__ __forceinline__ void wh_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                               int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ bool wh_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wh_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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

#ifdef VINET_ST_PROF
#define WH_PROF_WAIT(acc, ...)            \
  do {                                    \
    const long long t0__ = clock64();     \
    __VA_ARGS__;                          \
    (acc) += clock64() - t0__;            \
  } while (0)
#else
#define WH_PROF_WAIT(acc, ...) \
  do {                         \
    __VA_ARGS__;               \
  } while (0)
#endif

template <bool UP>
__global__ void __launch_bounds__(UP ? WH_THREADS + WH_UP_THREADS : WH_THREADS, 1) conv_wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  [[maybe_unused]] long long prof_w = 0, prof_e = 0;
  [[maybe_unused]] const long long prof_t0 = clock64();
  const vinet_gather_t& g = p.d.g;
  const int64_t nchunks = (int64_t)g.B * p.nT * p.tiles_h * p.tiles_w;
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
  // this CTA's group: (temporal tap, channel block) with all spatial taps, or (channel block) with all temporal taps
  const int dt = p.temporal ? 0 : blockIdx.x / p.ncb, cb = blockIdx.x - dt * p.ncb;
  const int nblk_eff = min(p.nblk, (min(BN, p.d.N - n0) + 63) / 64);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA[1])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmDy)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(full0 + 8 * s, UP ? 2 : 1);
        mbar_init(empty0 + 8 * s, p.ni);
      }
      mbar_init(accum_bar, p.ni);
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
    // ---------------------------------------------------------------- producer: one halo box + the dY box(es) per chunk
    if (lane == 0) {
      int64_t m = c_begin;
      int tw = (int)(m % p.tiles_w); m /= p.tiles_w;
      int th = (int)(m % p.tiles_h); m /= p.tiles_h;
      int tr = (int)(m % p.nT);
      int b = (int)(m / p.nT);
      const bool cat = g.src[1].ptr != nullptr;
      const int T0 = g.src[0].T;
      const int at0 = p.at0 + (p.temporal ? 0 : g.tap[dt * p.nsp][0]);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < KB; ++kb) {
        WH_PROF_WAIT(prof_w, mbar_wait(empty0 + 8 * s, ph ^ 1u));
        const uint32_t stage = s0 + (uint32_t)s * p.stage_bytes;
        const int ts = tr * p.at_step + at0;   // out-of-range frames are addressed on purpose: TMA zero-fills the temporal padding
        const int si = (cat && ts >= T0) ? 1 : 0;
        const bool interp = UP && si == 0;     // the activation box of this chunk is built by the interpolating warps
        mbar_arrive_expect_tx(full0 + 8 * s, (interp ? 0u : p.a_tx) + (uint32_t)nblk_eff * p.dy_unit);
        for (int k = 0; k < p.nbox && !interp; ++k)
          wh_tma_load_5d(stage + (uint32_t)p.box_off[k], &p.tmA[si], full0 + 8 * s, cb * 64, tw * p.cw + p.aw0,
                         th * p.ch * p.ah_mul + p.ah0 + p.box_h0[k], ts - (si ? T0 : 0), b);
        for (int nb = 0; nb < nblk_eff; ++nb)
          wh_tma_load_5d(stage + p.a_bytes + (uint32_t)nb * p.dy_unit, &p.tmDy, full0 + 8 * s, n0 + nb * 64, tw * p.cw, th * p.ch,
                         tr * p.dyt_step, b);
        if (++s == stages) { s = 0; ph ^= 1u; }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            if (++tr == p.nT) { tr = 0; ++b; }
          }
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + WH_MAX_ISSUERS) {
    // ---------------------------------------------------------------- MMA issuers: issuer k owns accumulators k, k+ni, ...
    const int k = warp - 2;
    if (k < p.ni) {
      const uint32_t hi_common = (1u << 14) | (2u << 29);
      const uint32_t a_hi = (p.a_sbo >> 4) | hi_common;   // SBO = distance between 8-position atoms (a halo tile row / S frames)
      const uint32_t b_hi = (1024u >> 4) | hi_common;
      const uint32_t b_lbo = (p.dy_unit >> 4) << 16;
      const uint32_t kstep_a = p.kstep_a16;               // 16 positions = two atoms
      const uint32_t stage16 = p.stage_bytes >> 4;
      const uint32_t s0_16 = (s0 & 0x3FFFFu) >> 4;
      int s = 0;
      uint32_t ph = 0, st16 = s0_16;
      for (int kb = 0; kb < KB; ++kb) {
        WH_PROF_WAIT(prof_w, mbar_wait(full0 + 8 * s, ph));
        tc_fence_after();
        const uint32_t b_lo0 = (st16 + (p.a_bytes >> 4)) | b_lbo;
        if (wh_elect_one()) {   // one elected lane issues this issuer's accumulators of the chunk and releases the stage
          for (int a = k; a < p.naccs; a += p.ni) {
            const int j0 = 2 * a, j1 = min(2 * a + 1, p.nsp - 1);
            const int o0 = p.uoff16[j0], o1 = p.uoff16[j1];
            const uint32_t a_lo0 = (st16 + (uint32_t)o0) | ((uint32_t)(o1 - o0) << 16);   // LBO = distance between the two taps
            const uint32_t tacc = tmem_base + (uint32_t)a * p.acc_cols;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              wh_umma(tacc, a_lo0 + (uint32_t)kk * kstep_a, a_hi, b_lo0 + (uint32_t)kk * 128u, b_hi, p.idesc, (uint32_t)((kb | kk) != 0));
          }
          umma_commit(empty0 + 8 * s);
          if (kb == KB - 1) umma_commit(accum_bar);
        }
        __syncwarp();
        st16 += stage16;
        if (++s == stages) { s = 0; ph ^= 1u; st16 = s0_16; }
      }
    }
  } else if (UP && warp >= WH_THREADS / 32) {
    // ---------------------------------------------------------------- interpolating producer (128 threads): activation boxes of
    // source 0 = relu + 2x bilinear of the low-res tensor, written in the TMA box layout
    const int itid = threadIdx.x - WH_THREADS;
    const vinet_src_t& src = g.src[0];
    const int lh = g.Hs >> 1, lw = g.Ws >> 1;
    const bool relu = (src.xform & VINET_XF_RELU) != 0;
    const int64_t frame_elems = (int64_t)lh * lw * src.ld;
    const int64_t clip_elems = src.ldb ? src.ldb : (int64_t)src.T * frame_elems;
    const __nv_bfloat16* src0 = reinterpret_cast<const __nv_bfloat16*>(src.ptr);
    int64_t m = c_begin;
    int tw = (int)(m % p.tiles_w); m /= p.tiles_w;
    int th = (int)(m % p.tiles_h); m /= p.tiles_h;
    int tr = (int)(m % p.nT);
    int b = (int)(m / p.nT);
    const bool cat = g.src[1].ptr != nullptr;
    const int T0 = src.T;
    const int at0 = p.at0 + (p.temporal ? 0 : g.tap[dt * p.nsp][0]);
    const int rem = g.Cs - cb * 64;
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      const int ts = tr * p.at_step + at0;
      if (!(cat && ts >= T0)) {
        up2_fill_box(s0 + (uint32_t)s * p.stage_bytes, src0 + (int64_t)b * clip_elems + (int64_t)ts * frame_elems + cb * 64,
                     ts >= 0 && ts < T0, lh, lw, src.ld, tw * p.cw + p.aw0, th * p.ch + p.ah0, p.abw, p.abh, 8, min(8, rem >> 3), relu,
                     itid, WH_UP_THREADS);
        fence_proxy_async();
      }
      asm volatile("bar.sync 2, %0;" ::"n"(WH_UP_THREADS) : "memory");
      if (itid == 0) mbar_arrive(full0 + 8 * s);
      if (++s == stages) { s = 0; ph ^= 1u; }
      if (++tw == p.tiles_w) {
        tw = 0;
        if (++th == p.tiles_h) {
          th = 0;
          if (++tr == p.nT) { tr = 0; ++b; }
        }
      }
    }
  } else if (warp >= 2 + WH_MAX_ISSUERS && warp < WH_THREADS / 32) {
    // ---------------------------------------------------------------- epilogue: TMEM -> smem transpose -> coalesced red.add
    WH_PROF_WAIT(prof_w, mbar_wait(accum_bar, 0));  // every MMA (hence every TMA write) of this CTA has completed: the stage ring is free
    tc_fence_after();
    fence_proxy_async();
    const int q = warp & 3;
    const int P = BN + 1;  // odd pitch: conflict-free column-wise writes
    float* stile = reinterpret_cast<float*>(base) + (size_t)(q * 32) * P;
    const int ncols = min(BN, p.d.N - n0);
    for (int a = 0; a < p.naccs; ++a) {
      // rows 32q..32q+31 of accumulator a belong to spatial tap 2a + (q >> 1), channels (q & 1)*32 ..
      const int j = 2 * a + (q >> 1);
      for (int gi = 0; gi < BN / 16; ++gi) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * p.acc_cols + (uint32_t)(gi * 16), r);
#pragma unroll
        for (int e = 0; e < 16; ++e) stile[lane * P + gi * 16 + e] = __uint_as_float(r[e]);
      }
      __syncwarp();
      if (j < p.nsp) {
        const int unit = (dt * p.nsp + j) * p.ncb + cb;
        const int c0 = (q & 1) * 32;
        const int rows = min(32, g.Cs - cb * 64 - c0);   // channels that exist
        for (int rr = 0; rr < rows; ++rr) {
          float* drow = p.d.dwp + ((int64_t)unit * 64 + c0 + rr) * p.d.lddw + n0;
          const float* srow = stile + rr * P;
          for (int col = lane * 4; col < ncols; col += 128) red_add_v4(drow + col, srow[col], srow[col + 1], srow[col + 2], srow[col + 3]);
        }
      }
      __syncwarp();
    }
  }
#ifdef VINET_ST_PROF
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp == 0 || warp == 2 || warp == 6))
    printf("wh_prof warp %d (%s): total %lld clk, waiting %lld | chunks/cta %d naccs %d ni %d block_n %d stages %d splits %d groups %d\n", warp,
           warp == 0 ? "tma: free stage" : warp == 2 ? "mma: stage data" : "epilogue: all MMAs done; rest = transpose + red.add",
           clock64() - prof_t0, prof_w, KB, p.naccs, p.ni, p.block_n, p.stages, p.splits, (int)gridDim.x);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

void pick_tma_box(int H, int W, int max_rows, int mult, bool full_tile_cost, int* bw_out, int* bh_out);

// returns 1 when the launch was handled here, 0 when the caller should use its own kernel, <0 on error
int conv_wgrad_halo(const vinet_wgrad_t* d, cudaStream_t stream, bool dry_run) {
  const vinet_gather_t& g = d->g;
  if (!g_stream_enable) return 0;
  if (g.mode != VINET_GATHER_FPROP || g.dtype != VINET_BF16 || d->dy_dtype != VINET_BF16) return 0;
  if (g.sw != 1 || g.Cs % 8 != 0 || d->N % 8 != 0 || d->lddy % 8 != 0) return 0;
  // row-strided conv over sliding-window rows (stem conv_s on the WIN8 input): see conv_gemm_stream_strided
  const bool rowstride = g.sh == 2 && g.src[1].ptr == nullptr && g.src[0].ld < g.Cs && g.Cs == 64 && g.st == 1 && g.pt == 0 &&
                         g.row_tstep == 1 && g.row_toff == 0 && g.src[0].xform == VINET_XF_IDENT;
  if (g.sh != 1 && !rowstride) return 0;
  // source 0 read through the fused relu + 2x up-sampling (interpolating warps; spatial mode with quad-aligned halo boxes only)
  const bool up = !rowstride && (g.src[0].xform & ~VINET_XF_RELU) == VINET_XF_UP2;
  for (int i = 0; i < 2 && !rowstride; ++i) {
    const vinet_src_t& s = g.src[i];
    if (s.ptr == nullptr) continue;
    if (i == 0 && up) {
      if ((g.Hs & 1) || (g.Ws & 1) || s.ldh != 0 || s.ld < g.Cs) return 0;
      continue;
    }
    if (s.xform != VINET_XF_IDENT || (s.ldh != 0 && s.ldh != (int64_t)g.Ws * s.ld) || s.ld < g.Cs) return 0;
  }
  // taps must be the natural (dt, dh, dw) enumeration of a kt x kh x kw kernel
  int kt = 0, kh = 0, kw = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    kt = std::max(kt, g.tap[t][0] + 1);
    kh = std::max(kh, g.tap[t][1] + 1);
    kw = std::max(kw, g.tap[t][2] + 1);
  }
  const int nspat = kh * kw;
  if (kt * nspat != g.ntaps) return 0;
  for (int t = 0; t < g.ntaps; ++t)
    if (g.tap[t][0] != t / nspat || g.tap[t][1] != (t / kw) % kh || g.tap[t][2] != t % kw) return 0;
  // spatial mode: kh*kw > 1 taps share a spatial halo box.  temporal mode: kh = kw = 1, the kt temporal taps of a single
  // source share a box of (tt-1)*st + kt frames x pos positions (tt*pos = 128 output positions per chunk).
  const bool temporal = !rowstride && nspat == 1 && kt > 1 && g.src[1].ptr == nullptr && g.row_tstep == 1 && g.row_toff == 0;
  if (rowstride && (kt != 1 || kw != 1 || kh < 2)) return 0;
  if (!temporal && (nspat < 2 || g.Hr < 10)) return 0;
  if (up && (temporal || !(g.pw & 1) || !(g.ph & 1) || !(kw & 1) || !(kh & 1))) return 0;
  WgHaloParams p;
  p.d = *d;
  p.temporal = temporal ? 1 : 0;
  p.nsp = temporal ? kt : nspat;
  if (p.nsp > WH_MAX_SP) return 0;
  p.ncb = (g.Cs + 63) / 64;
  p.naccs = (p.nsp + 1) / 2;
  const int n16 = (int)round_up(d->N, 16);
  const int bn_max = std::min(256, (512 / p.naccs) / 16 * 16);
  const int n_tiles = (int)cdiv(n16, bn_max);
  p.block_n = (int)round_up(cdiv(n16, n_tiles), 16);
  p.acc_cols = (uint32_t)round_up(p.block_n, 32);
  if ((int)p.acc_cols * p.naccs > 512) p.acc_cols = (uint32_t)p.block_n;   // e.g. 5 x 96
  if ((int)p.acc_cols * p.naccs > 512) return 0;
  p.tmem_cols = tmem_cols_for(p.naccs * (int)p.acc_cols);
  p.nblk = (p.block_n + 63) / 64;
  p.dy_unit = WH_UNIT;
  int abw, abh, abt = 1, dbw, dbh, dbt = 1, aesh = 1;   // activation / dY box extents, row element stride of the activation box
  p.nbox = 1; p.ah_mul = 1; p.box_h0[0] = p.box_h0[1] = 0; p.box_off[0] = p.box_off[1] = 0;
  if (rowstride) {
    // tap dh reads source row 2h + (dh - ph): lattice par = (dh - ph) mod 2, row q = (dh - ph - par) / 2 inside its box
    int q[WH_MAX_SP], par[WH_MAX_SP], qmin[2] = {1 << 20, 1 << 20}, qmax[2] = {-(1 << 20), -(1 << 20)};
    for (int j = 0; j < kh; ++j) {
      const int o = j - g.ph;
      par[j] = ((o % 2) + 2) % 2;
      q[j] = (o - par[j]) / 2;
      qmin[par[j]] = std::min(qmin[par[j]], q[j]);
      qmax[par[j]] = std::max(qmax[par[j]], q[j]);
    }
    if (qmax[0] < qmin[0] || qmax[1] < qmin[1]) return 0;
    const int PHm = 16 + std::max(qmax[0] - qmin[0], qmax[1] - qmin[1]);
    abw = 8; abh = PHm; dbw = 8; dbh = 16; aesh = 2;
    p.cw = 8; p.ch = 16; p.aw0 = 0; p.ah0 = 0; p.ah_mul = 2;
    p.at_step = 1; p.at0 = 0; p.dyt_step = 1;
    p.nT = g.Tr;
    p.tiles_w = (int)cdiv(g.Wr, 8);
    p.tiles_h = (int)cdiv(g.Hr, 16);
    p.nbox = 2;
    const int first = par[0];                    // the lattice of tap 0 goes first so that stacked pairs have LBO >= 0
    for (int k = 0; k < 2; ++k) {
      p.box_h0[k] = 2 * qmin[k] + k;
      p.box_off[k] = (k == first ? 0 : PHm * 8 * 128);
    }
    for (int j = 0; j < WH_MAX_SP; ++j) p.uoff16[j] = j < kh ? (p.box_off[par[j]] + (q[j] - qmin[par[j]]) * 8 * 128) >> 4 : 0;
    for (int a2 = 0; 2 * a2 + 1 < kh; ++a2)
      if (p.uoff16[2 * a2 + 1] < p.uoff16[2 * a2]) return 0;
    p.a_sbo = 1024u;
    p.kstep_a16 = 2048u >> 4;
  } else if (!temporal) {
    const int PW = 8 + kw - 1, PH = 16 + kh - 1;
    abw = PW; abh = PH; dbw = 8; dbh = 16;
    p.cw = 8; p.ch = 16; p.aw0 = -g.pw; p.ah0 = -g.ph;
    p.at_step = g.row_tstep * g.st; p.at0 = g.row_toff * g.st - g.pt; p.dyt_step = 1;
    p.nT = g.Tr;
    p.tiles_w = (int)cdiv(g.Wr, 8);
    p.tiles_h = (int)cdiv(g.Hr, 16);
    for (int j = 0; j < WH_MAX_SP; ++j) p.uoff16[j] = j < nspat ? ((j / kw) * PW + (j % kw)) * 8 : 0;
    p.a_sbo = (uint32_t)PW * 128u;
    p.kstep_a16 = (uint32_t)(2 * PW * 128) >> 4;
  } else {
    // tt output frames x pos positions per chunk; a temporally strided conv needs pos = 8 (one atom = one source frame)
    int tt = g.st > 1 ? 16 : 8;
    while (tt > 1 && tt / 2 >= g.Tr) tt /= 2;
    if (g.st > 1 && tt != 16) return 0;
    const int pos = TC_BM / tt;
    pick_tma_box(g.Hr, g.Wr, pos, 8, true, &dbw, &dbh);
    if (dbw * dbh != pos) return 0;
    abw = dbw; abh = dbh;
    abt = (tt - 1) * g.st + kt;
    dbt = tt;
    if (abt > 256) return 0;
    p.cw = dbw; p.ch = dbh; p.aw0 = 0; p.ah0 = 0;
    p.at_step = tt * g.st; p.at0 = -g.pt; p.dyt_step = tt;
    p.nT = (int)cdiv(g.Tr, tt);
    p.tiles_w = (int)cdiv(g.Wr, dbw);
    p.tiles_h = (int)cdiv(g.Hr, dbh);
    for (int j = 0; j < WH_MAX_SP; ++j) p.uoff16[j] = j < kt ? j * pos * 8 : 0;   // tap dt = the box shifted by dt frames
    p.a_sbo = g.st > 1 ? (uint32_t)(g.st * pos * 128) : 1024u;
    p.kstep_a16 = (2u * p.a_sbo) >> 4;
  }
  p.abw = abw; p.abh = abh;
  p.a_tx = (uint32_t)(abw * abh * abt * 128) * (uint32_t)p.nbox;
  p.a_bytes = (uint32_t)round_up(p.a_tx, 1024);
  p.stage_bytes = p.a_bytes + (uint32_t)p.nblk * p.dy_unit;
  p.idesc = make_idesc(TC_BM, p.block_n, 1, 1);
  p.ni = std::min(WH_MAX_ISSUERS, p.naccs);
  const int64_t nchunks = (int64_t)g.B * p.nT * p.tiles_h * p.tiles_w;
  const int groups = (temporal ? 1 : kt) * p.ncb;
  const int64_t base_ctas = (int64_t)groups * n_tiles;
  const int sms = tma_sm_count();
  // split the position chunks so that the CTA count fills whole waves of the machine (1 CTA per SM)
  const int64_t smax = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nchunks / 2, 256), 65535));
  int64_t splits = 1;
  double best = 1e30;
  for (int64_t sp = 1; sp <= smax; ++sp) {
    // time in chunk units: waves x (main loop + prologue / transposing red.add epilogue, ~2.5 chunks per 96-column accumulator)
    const double t = (double)cdiv(base_ctas * sp, sms) * ((double)cdiv(nchunks, sp) + 2.0 + 2.5 * p.naccs * p.block_n / 96.0);
    if (t < best * 0.999) { best = t; splits = sp; }
  }
  p.splits = (int)splits;
  int stages = (int)((200 * 1024) / p.stage_bytes);
  stages = std::max(2, std::min(stages, 6));
  p.stages = stages;
  const size_t ring = (size_t)stages * p.stage_bytes;
  if (ring < (size_t)128 * (p.block_n + 1) * sizeof(float)) return 0;   // the epilogue transposes through the ring
  const size_t smem = std::max<size_t>(1024 + ring + 8 * (2 * stages + 1) + 64, 120 * 1024);
  if (smem > 227 * 1024) return 0;
  if (dry_run) return 1;      // host-only eligibility query (vinet_conv_up2_fused): no tensor maps, no launch
  for (int i = 0; i < 2; ++i) {
    const vinet_src_t& s = g.src[(i == 1 && g.src[1].ptr == nullptr) ? 0 : i];
    const int dv = (up && &s == &g.src[0]) ? 2 : 1;   // an up-sampled source is never fetched by TMA: valid map of the low-res tensor, unused
    if (make_tma_map(&p.tmA[i], s.ptr, g.Cs, g.Ws / dv, g.Hs / dv, s.T, g.B, s.ld, s.ldh, abw, abh, 1, aesh, abt, s.ldb)) return -1;
  }
  if (make_tma_map(&p.tmDy, d->dy, d->N, g.Wr, g.Hr, g.Tr, g.B, d->lddy, 0, dbw, dbh, 1, 1, dbt)) return -1;
  dim3 grid((unsigned)groups, (unsigned)n_tiles, (unsigned)splits);
  if (up) {
    cudaFuncSetAttribute(conv_wgrad_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_wgrad_halo_kernel<true><<<grid, WH_THREADS + WH_UP_THREADS, smem, stream>>>(p);
  } else {
    cudaFuncSetAttribute(conv_wgrad_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_wgrad_halo_kernel<false><<<grid, WH_THREADS, smem, stream>>>(p);
  }
  note_kernel("conv_wgrad_halo_kernel");
  if (up) g_up2_launches.fetch_add(1, std::memory_order_relaxed);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("conv_wgrad_halo: launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 1;
}

}  // namespace vinet
