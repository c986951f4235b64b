// nn.MaxPool3d on NDHWC views (model.py:696-714 stage pools, model_utils.py:178 3x3x3 stride-1 branch pools,
// model.py:229 AV pool).  One thread = one output position x 8 channels (one 128-bit vector per tap).
// Forward optionally records, per output element, the index of the winning tap (first maximum in (t,h,w) scan
// order, like ATen) as one byte; backward then scatters the gradient without re-scanning the window.  Without a
// recorded index (inference-only forward / legacy callers) backward recomputes the arg-max.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vinet {

struct PoolItem {
  int c, wo, ho, to, b;
};

__device__ __forceinline__ PoolItem pool_decode(const vinet_pool_t& d, int64_t i, int G) {
  PoolItem it;
  // items < 2^31 for every tensor of this workload; keep the divisions 32-bit when they are
  if (i < (int64_t)0x7fffffff) {
    unsigned r = (unsigned)i;
    it.c = (int)(r % (unsigned)G) * 8; r /= (unsigned)G;
    it.wo = (int)(r % (unsigned)d.Wo); r /= (unsigned)d.Wo;
    it.ho = (int)(r % (unsigned)d.Ho); r /= (unsigned)d.Ho;
    it.to = (int)(r % (unsigned)d.To);
    it.b = (int)(r / (unsigned)d.To);
  } else {
    int64_t r = i;
    it.c = (int)(r % G) * 8; r /= G;
    it.wo = (int)(r % d.Wo); r /= d.Wo;
    it.ho = (int)(r % d.Ho); r /= d.Ho;
    it.to = (int)(r % d.To);
    it.b = (int)(r / d.To);
  }
  return it;
}

// running maximum + winning tap for 8 channels
template <typename T>
struct Max8;

template <>
struct Max8<float> {
  float best[8];
  unsigned idx[8];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; idx[e] = 0xffu; }
  }
  __device__ __forceinline__ void update(const float* p, const vinet_pool_t& d, int c, unsigned tap) {
    float v[8];
    load8(p, v);
    apply_xform<8>(v, d.xform, d.scale, d.shift, c);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (v[e] > best[e]) { best[e] = v[e]; idx[e] = tap; }
  }
  __device__ __forceinline__ void values(float (&v)[8]) const {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = best[e];
  }
  __device__ __forceinline__ unsigned tap_of(int e) const { return idx[e]; }
};

template <>
struct Max8<__nv_bfloat16> {
  // packed bf16x2 compare / max; tap indices ride in the two 16-bit lanes of a 32-bit word
  __nv_bfloat162 best[4];
  unsigned idx[4];
  float fbest[8];  // only used when the source carries a pending transform (values are then compared in fp32)
  unsigned fidx[8];
  bool plain;
  __device__ __forceinline__ void init() {
    const __nv_bfloat162 ninf = __halves2bfloat162(__ushort_as_bfloat16(0xff80), __ushort_as_bfloat16(0xff80));
#pragma unroll
    for (int q = 0; q < 4; ++q) { best[q] = ninf; idx[q] = 0x00ff00ffu; }
#pragma unroll
    for (int e = 0; e < 8; ++e) { fbest[e] = -INFINITY; fidx[e] = 0xffu; }
    plain = true;
  }
  __device__ __forceinline__ void update(const __nv_bfloat16* p, const vinet_pool_t& d, int c, unsigned tap) {
    if (d.xform == VINET_XF_IDENT) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
      const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&u);
      const unsigned tap2 = tap | (tap << 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned m = __hgt2_mask(v[q], best[q]);
        idx[q] = (idx[q] & ~m) | (tap2 & m);
        best[q] = __hmax2(best[q], v[q]);
      }
    } else {
      plain = false;
      float v[8];
      load8(p, v);
      apply_xform<8>(v, d.xform, d.scale, d.shift, c);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (v[e] > fbest[e]) { fbest[e] = v[e]; fidx[e] = tap; }
    }
  }
  __device__ __forceinline__ void values(float (&v)[8]) const {
    if (plain) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[2 * q] = __low2float(best[q]);
        v[2 * q + 1] = __high2float(best[q]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fbest[e];
    }
  }
  __device__ __forceinline__ unsigned tap_of(int e) const {
    return plain ? ((idx[e >> 1] >> ((e & 1) * 16)) & 0xffu) : fidx[e];
  }
};

// scan the window of one output position; returns the source offset (elements, without channel) of its origin
template <typename T>
__device__ __forceinline__ void pool_scan(const vinet_pool_t& d, const T* __restrict__ x, const PoolItem& it, Max8<T>& acc) {
  const int t0 = it.to * d.st - d.pt, h0 = it.ho * d.sh - d.ph, w0 = it.wo * d.sw - d.pw;
  const int dt_lo = max(0, -t0), dt_hi = min(d.kt, d.Ti - t0);
  const int dh_lo = max(0, -h0), dh_hi = min(d.kh, d.Hi - h0);
  const int dw_lo = max(0, -w0), dw_hi = min(d.kw, d.Wi - w0);
  acc.init();
  for (int dt = dt_lo; dt < dt_hi; ++dt) {
    for (int dh = dh_lo; dh < dh_hi; ++dh) {
      const int64_t row = (((int64_t)it.b * d.Ti + (t0 + dt)) * d.Hi + (h0 + dh)) * d.Wi + w0;
      const T* p = x + (row + dw_lo) * d.ldx + it.c;
      unsigned tap = (unsigned)((dt * d.kh + dh) * d.kw + dw_lo);
      for (int dw = dw_lo; dw < dw_hi; ++dw, ++tap, p += d.ldx) acc.update(p, d, it.c, tap);
    }
  }
}

template <typename T, typename TO, bool IDX>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __grid_constant__ vinet_pool_t d) {
  const T* __restrict__ x = reinterpret_cast<const T*>(d.x);
  const int G = d.C / 8;
  const int64_t total = (int64_t)d.B * d.To * d.Ho * d.Wo * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const PoolItem it = pool_decode(d, i, G);
    Max8<T> acc;
    pool_scan<T>(d, x, it, acc);
    const int64_t opos = (((int64_t)it.b * d.To + it.to) * d.Ho + it.ho) * d.Wo + it.wo;
    float best[8];
    acc.values(best);
    store8(reinterpret_cast<TO*>(d.out) + opos * d.ldo + it.c, best);
    if constexpr (IDX) {
      uint2 pk;
      pk.x = acc.tap_of(0) | (acc.tap_of(1) << 8) | (acc.tap_of(2) << 16) | (acc.tap_of(3) << 24);
      pk.y = acc.tap_of(4) | (acc.tap_of(5) << 8) | (acc.tap_of(6) << 16) | (acc.tap_of(7) << 24);
      *reinterpret_cast<uint2*>(d.idx + opos * d.C + it.c) = pk;
    }
  }
}

// ---- 3x3x3 / stride 1 / pad 1 (the nine Mixed_* branch pools, model_utils.py:178...): one thread owns an (h, w, 8-channel)
// column and walks the frames, keeping the 3x3 spatial maximum (value + winning tap) of the last three frames in registers:
// 9 loads and ~11 packed compare steps per output instead of 27 + 27.  Ties resolve exactly like the scan-order kernel:
// within a frame taps are visited in (dh, dw) order with a strict >, and frames are combined in dt order with a strict >.
struct FrameMax {
  __nv_bfloat162 v[4];
  unsigned idx[4];  // two 16-bit lanes holding dh*3+dw
};

__device__ __forceinline__ void frame_max(const vinet_pool_t& d, const __nv_bfloat16* __restrict__ x, int b, int t, int h, int w, int c,
                                          FrameMax& f) {
  const __nv_bfloat162 ninf = __halves2bfloat162(__ushort_as_bfloat16(0xff80), __ushort_as_bfloat16(0xff80));
#pragma unroll
  for (int q = 0; q < 4; ++q) { f.v[q] = ninf; f.idx[q] = 0x00ff00ffu; }
  const int dh_lo = h == 0 ? 1 : 0, dh_hi = h == d.Hi - 1 ? 2 : 3;
  const int dw_lo = w == 0 ? 1 : 0, dw_hi = w == d.Wi - 1 ? 2 : 3;
  for (int dh = dh_lo; dh < dh_hi; ++dh) {
    const int64_t row = (((int64_t)b * d.Ti + t) * d.Hi + (h - 1 + dh)) * d.Wi + (w - 1);
    for (int dw = dw_lo; dw < dw_hi; ++dw) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (row + dw) * d.ldx + c));
      const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&u);
      const unsigned tap = (unsigned)(dh * 3 + dw), tap2 = tap | (tap << 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned m = __hgt2_mask(v[q], f.v[q]);
        f.idx[q] = (f.idx[q] & ~m) | (tap2 & m);
        f.v[q] = __hmax2(f.v[q], v[q]);
      }
    }
  }
}

// A thread owns a (h, w, 8-channel) column and a run of `tc` output frames (blockIdx.y): the deep stages have few positions and
// few frames, and one thread per column walking ALL frames left most of the machine idle (0.7 TB/s); a run re-computes the
// 3x3 maxima of its two boundary frames.
template <bool IDX>
__global__ void __launch_bounds__(256) maxpool333_fwd_kernel(const __grid_constant__ vinet_pool_t d, int tc) {
  const __nv_bfloat16* __restrict__ x = reinterpret_cast<const __nv_bfloat16*>(d.x);
  __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(d.out);
  const int G = d.C / 8;
  const int64_t total = (int64_t)d.B * d.Hi * d.Wi * G;
  const int t_begin = blockIdx.y * tc, t_end = min(d.Ti, t_begin + tc);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned r = (unsigned)i;
    const int c = (int)(r % (unsigned)G) * 8; r /= (unsigned)G;
    const int w = (int)(r % (unsigned)d.Wi); r /= (unsigned)d.Wi;
    const int h = (int)(r % (unsigned)d.Hi);
    const int b = (int)(r / (unsigned)d.Hi);
    FrameMax prev, cur, next;
    if (t_begin > 0) frame_max(d, x, b, t_begin - 1, h, w, c, prev);
    frame_max(d, x, b, t_begin, h, w, c, cur);
    if (t_begin + 1 < d.Ti) frame_max(d, x, b, t_begin + 1, h, w, c, next);
    for (int t = t_begin; t < t_end; ++t) {
      __nv_bfloat162 bv[4];
      unsigned bi[4];
      if (t > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { bv[q] = prev.v[q]; bi[q] = prev.idx[q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned m = __hgt2_mask(cur.v[q], bv[q]);
          bi[q] = (bi[q] & ~m) | ((cur.idx[q] + 0x00090009u) & m);
          bv[q] = __hmax2(bv[q], cur.v[q]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { bv[q] = cur.v[q]; bi[q] = cur.idx[q] + 0x00090009u; }
      }
      if (t + 1 < d.Ti) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned m = __hgt2_mask(next.v[q], bv[q]);
          bi[q] = (bi[q] & ~m) | ((next.idx[q] + 0x00120012u) & m);
          bv[q] = __hmax2(bv[q], next.v[q]);
        }
      }
      const int64_t opos = (((int64_t)b * d.To + t) * d.Ho + h) * d.Wo + w;
      uint4 o;
      o.x = *reinterpret_cast<unsigned*>(&bv[0]); o.y = *reinterpret_cast<unsigned*>(&bv[1]);
      o.z = *reinterpret_cast<unsigned*>(&bv[2]); o.w = *reinterpret_cast<unsigned*>(&bv[3]);
      *reinterpret_cast<uint4*>(out + opos * d.ldo + c) = o;
      if constexpr (IDX) {
        uint2 pk;
        pk.x = (bi[0] & 0xffu) | ((bi[0] >> 8) & 0xff00u) | ((bi[1] & 0xffu) << 16) | ((bi[1] >> 16) << 24);
        pk.y = (bi[2] & 0xffu) | ((bi[2] >> 8) & 0xff00u) | ((bi[3] & 0xffu) << 16) | ((bi[3] >> 16) << 24);
        *reinterpret_cast<uint2*>(d.idx + opos * d.C + c) = pk;
      }
      prev = cur;
      cur = next;
      if (t + 1 < t_end && t + 2 < d.Ti) frame_max(d, x, b, t + 2, h, w, c, next);
    }
  }
}

template <typename TGI>
__device__ __forceinline__ void grad_atomic_add(TGI* gin, int e, float g) {
  if constexpr (sizeof(TGI) == 4) {
    atomicAdd(gin + e, g);
  } else {
    // bf16 gradients: native red.add.bf16x2 on the aligned pair holding channel e
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f), v = __float2bfloat16_rn(g);
    const __nv_bfloat162 pair = (e & 1) ? __halves2bfloat162(z, v) : __halves2bfloat162(v, z);
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(gin + (e & ~1)), pair);
  }
}

// T: activation type (arg-max recompute when no index was recorded), TGO / TGI: gradient types
template <typename T, typename TGO, typename TGI, bool IDX>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __grid_constant__ vinet_pool_t d) {
  const T* __restrict__ x = reinterpret_cast<const T*>(d.x);
  const int G = d.C / 8;
  const int khw = d.kh * d.kw;
  const int64_t total = (int64_t)d.B * d.To * d.Ho * d.Wo * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const PoolItem it = pool_decode(d, i, G);
    const int64_t opos = (((int64_t)it.b * d.To + it.to) * d.Ho + it.ho) * d.Wo + it.wo;
    unsigned taps[8];
    if constexpr (IDX) {
      const uint2 pk = __ldg(reinterpret_cast<const uint2*>(d.idx + opos * d.C + it.c));
#pragma unroll
      for (int e = 0; e < 4; ++e) { taps[e] = (pk.x >> (8 * e)) & 0xffu; taps[4 + e] = (pk.y >> (8 * e)) & 0xffu; }
    } else {
      Max8<T> acc;
      pool_scan<T>(d, x, it, acc);
#pragma unroll
      for (int e = 0; e < 8; ++e) taps[e] = acc.tap_of(e);
    }
    float g[8];
    load8(reinterpret_cast<const TGO*>(d.gout) + opos * d.ldgo + it.c, g);
    const int t0 = it.to * d.st - d.pt, h0 = it.ho * d.sh - d.ph, w0 = it.wo * d.sw - d.pw;
    TGI* gin = reinterpret_cast<TGI*>(d.gin);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned tap = taps[e];
      if (tap == 0xffu) continue;
      const int dt = (int)tap / khw, r = (int)tap - dt * khw;
      const int dh = r / d.kw, dw = r - dh * d.kw;
      const int64_t pos = (((int64_t)it.b * d.Ti + (t0 + dt)) * d.Hi + (h0 + dh)) * d.Wi + (w0 + dw);
      grad_atomic_add<TGI>(gin + pos * d.ldgi + it.c, e, g[e]);
    }
  }
}

// acc (4 x bf16x2) += g (8 packed bf16) where the recorded tap bytes equal tap4: branch-free (a warp would otherwise run the
// per-channel test-and-add body for nearly every candidate window, because SOME lane always matches).  The few terms per element
// are summed in bf16, like the red.add.bf16x2 of the scatter kernel.
__device__ __forceinline__ void masked_add_bf16x8(__nv_bfloat162 (&acc)[4], const uint4& g, const uint2& pk, unsigned tap4) {
  const unsigned m0 = __vcmpeq4(pk.x, tap4), m1 = __vcmpeq4(pk.y, tap4);
  const unsigned w0 = g.x & __byte_perm(m0, 0, 0x1100), w1 = g.y & __byte_perm(m0, 0, 0x3322);
  const unsigned w2 = g.z & __byte_perm(m1, 0, 0x1100), w3 = g.w & __byte_perm(m1, 0, 0x3322);
  acc[0] = __hadd2(acc[0], *reinterpret_cast<const __nv_bfloat162*>(&w0));
  acc[1] = __hadd2(acc[1], *reinterpret_cast<const __nv_bfloat162*>(&w1));
  acc[2] = __hadd2(acc[2], *reinterpret_cast<const __nv_bfloat162*>(&w2));
  acc[3] = __hadd2(acc[3], *reinterpret_cast<const __nv_bfloat162*>(&w3));
}

// Backward as a GATHER over the recorded arg-max taps: one thread = one INPUT position x 8 channels.  It visits the (at most
// ceil(k/s)^3) windows that contain the position, compares their recorded tap bytes with the tap this position would be in
// that window, and sums the matching output gradients.  No atomics (deterministic), no zero-initialised gradient buffer.
template <typename TGO, typename TGI>
__global__ void __launch_bounds__(256) maxpool_bwd_gather_kernel(const __grid_constant__ vinet_pool_t d) {
  const int G = d.C / 8;
  const int64_t total = (int64_t)d.B * d.Ti * d.Hi * d.Wi * G;
  const TGO* __restrict__ gout = reinterpret_cast<const TGO*>(d.gout);
  TGI* __restrict__ gin = reinterpret_cast<TGI*>(d.gin);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c = (int)(r % G) * 8; r /= G;
    const int w = (int)(r % d.Wi); r /= d.Wi;
    const int h = (int)(r % d.Hi); r /= d.Hi;
    const int t = (int)(r % d.Ti);
    const int b = (int)(r / d.Ti);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // windows (to, ho, wo) containing (t, h, w): to*st - pt <= t < to*st - pt + kt
    const int to_hi = min(d.To - 1, (t + d.pt) / d.st), to_lo = max(0, (t + d.pt - d.kt + d.st) / d.st);
    const int ho_hi = min(d.Ho - 1, (h + d.ph) / d.sh), ho_lo = max(0, (h + d.ph - d.kh + d.sh) / d.sh);
    const int wo_hi = min(d.Wo - 1, (w + d.pw) / d.sw), wo_lo = max(0, (w + d.pw - d.kw + d.sw) / d.sw);
    for (int to = to_lo; to <= to_hi; ++to) {
      const int dt = t + d.pt - to * d.st;
      for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        const int dh = h + d.ph - ho * d.sh;
        const int64_t orow = (((int64_t)b * d.To + to) * d.Ho + ho) * d.Wo;
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          const int dw = w + d.pw - wo * d.sw;
          const unsigned tap = (unsigned)((dt * d.kh + dh) * d.kw + dw);
          const unsigned tap4 = tap * 0x01010101u;
          const uint2 pk = __ldg(reinterpret_cast<const uint2*>(d.idx + (orow + wo) * d.C + c));
          const unsigned m0 = __vcmpeq4(pk.x, tap4), m1 = __vcmpeq4(pk.y, tap4);
          if ((m0 | m1) == 0u) continue;
          float g[8];
          load8(gout + (orow + wo) * d.ldgo + c, g);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if ((m0 >> (8 * e)) & 1u) acc[e] += g[e];
            if ((m1 >> (8 * e)) & 1u) acc[4 + e] += g[4 + e];
          }
        }
      }
    }
    TGI* dst = gin + ((((int64_t)b * d.Ti + t) * d.Hi + h) * d.Wi + w) * d.ldgi + c;
    if (!d.gin_overwrite) {
      float o[8];
      load8(dst, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += o[e];
    }
    store8(dst, acc);
  }
}

// The same gather with the window geometry as compile-time constants (the four pools of this model): one block walks one input
// row (b, t, h), a thread owns (w, 8 channels); the candidate windows are enumerated by tap with the stride-lattice tests
// folded at compile time, all index arithmetic is 32-bit.  The generic kernel above spends ~450 instructions per item on 64-bit
// divisions and run-time loop bounds (0.38 ms for the stem pool); this one is bound by its 16-byte loads and stores.
template <int KT, int KH, int KW, int ST, int SH, int SW, int PT, int PH, int PW, typename TGO, typename TGI>
__global__ void __launch_bounds__(256) maxpool_bwd_gather_fixed_kernel(const __grid_constant__ vinet_pool_t d) {
  // candidate windows per dimension: to = floor((t + PT) / ST) - j, j < ceil(KT / ST)
  constexpr int NT = (KT + ST - 1) / ST, NH = (KH + SH - 1) / SH, NW = (KW + SW - 1) / SW, NC = NT * NH * NW;
  constexpr bool EAGER = NC <= 8;   // few candidates: fetch their gradients unconditionally, all loads in flight at once
  const int G = d.C / 8;
  const TGO* __restrict__ gout = reinterpret_cast<const TGO*>(d.gout);
  TGI* __restrict__ gin = reinterpret_cast<TGI*>(d.gin);
  const int nrows = d.B * d.Ti * d.Hi;
  const int items = d.Wi * G;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int h = row % d.Hi;
    const int bt = row / d.Hi;
    const int t = bt % d.Ti, b = bt / d.Ti;
    const int th = (t + PT) / ST, hh = (h + PH) / SH;
    for (int i = threadIdx.x; i < items; i += 256) {
      const int w = i / G;
      const int c = (i - w * G) * 8;
      const int wh = (w + PW) / SW;
      // phase 1: every candidate's recorded tap bytes (and, when few, its gradient) - independent loads
      uint2 pk[NC];
      int off[NC];          // output position, or -1
      unsigned tap4[NC];
      uint4 gq[EAGER ? NC : 1];
#pragma unroll
      for (int jt = 0; jt < NT; ++jt) {
        const int to = th - jt, dt = t + PT - to * ST;
        const bool vt = to >= 0 && to < d.To && dt < KT;
#pragma unroll
        for (int jh = 0; jh < NH; ++jh) {
          const int ho = hh - jh, dh = h + PH - ho * SH;
          const bool vh = vt && ho >= 0 && ho < d.Ho && dh < KH;
#pragma unroll
          for (int jw = 0; jw < NW; ++jw) {
            const int wo = wh - jw, dw = w + PW - wo * SW;
            const bool v = vh && wo >= 0 && wo < d.Wo && dw < KW;
            const int k = (jt * NH + jh) * NW + jw;
            const int o = ((b * d.To + to) * d.Ho + ho) * d.Wo + wo;
            off[k] = v ? o : -1;
            tap4[k] = (unsigned)((dt * KH + dh) * KW + dw) * 0x01010101u;
            pk[k] = v ? __ldg(reinterpret_cast<const uint2*>(d.idx + (int64_t)o * d.C + c)) : make_uint2(0xffffffffu, 0xffffffffu);
            if constexpr (EAGER) {
              if constexpr (sizeof(TGO) == 2) {
                gq[k] = v ? __ldg(reinterpret_cast<const uint4*>(gout + (int64_t)o * d.ldgo + c)) : make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        }
      }
      // phase 2: sum the gradients of the windows whose arg-max is this element
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      if constexpr (EAGER && sizeof(TGO) == 2) {      // branch-free packed accumulation (invalid candidates carry 0xff taps)
        __nv_bfloat162 acc2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc2[q] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NC; ++k) masked_add_bf16x8(acc2, gq[k], pk[k], tap4[k]);
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[2 * q] = __low2float(acc2[q]); acc[2 * q + 1] = __high2float(acc2[q]); }
      }
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if constexpr (EAGER && sizeof(TGO) == 2) break;
        if (off[k] < 0) continue;
        const unsigned m0 = __vcmpeq4(pk[k].x, tap4[k]), m1 = __vcmpeq4(pk[k].y, tap4[k]);
        if ((m0 | m1) == 0u) continue;
        float g[8];
        if constexpr (EAGER && sizeof(TGO) == 2) {
          g[0] = bf16_lo(gq[k].x); g[1] = bf16_hi(gq[k].x); g[2] = bf16_lo(gq[k].y); g[3] = bf16_hi(gq[k].y);
          g[4] = bf16_lo(gq[k].z); g[5] = bf16_hi(gq[k].z); g[6] = bf16_lo(gq[k].w); g[7] = bf16_hi(gq[k].w);
        } else {
          load8(gout + (int64_t)off[k] * d.ldgo + c, g);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if ((m0 >> (8 * e)) & 1u) acc[e] += g[e];
          if ((m1 >> (8 * e)) & 1u) acc[4 + e] += g[4 + e];
        }
      }
      TGI* dst = gin + ((int64_t)row * d.Wi + w) * d.ldgi + c;
      if (!d.gin_overwrite) {
        float o[8];
        load8(dst, o);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += o[e];
      }
      store8(dst, acc);
    }
  }
}

}  // namespace vinet
using namespace vinet;

namespace vinet {

// ---- backward of the 3x3x3 / stride 1 / pad 1 pools (the nine Mixed_* branch pools) without atomics ---------------------------------
// Every input element belongs to 27 windows.  A block owns an 8 x 8 spatial tile x 32 channels and walks the frames: it stages
// the recorded arg-max bytes and the gradients of the 10 x 10 windows around its tile for three consecutive output frames in
// shared memory (ring of 3), and each thread then gathers for ONE input element (8 channels) from its 27 windows with LDS only:
// no global atomics (the scatter kernel is bound by ~8 L2 red.add per window, 325 us for Mixed_3c at batch 8), no zero-filled
// gradient buffer, deterministic summation order.
constexpr int PT_TS = 8, PT_HS = PT_TS + 2, PT_NP = PT_HS * PT_HS, PT_CG = 4;



template <typename TGO, typename TGI>
__global__ void __launch_bounds__(256) maxpool333_bwd_tile_kernel(const __grid_constant__ vinet_pool_t d, int tiles_w, int tiles_h, int nchunk) {
  constexpr int GW = sizeof(TGO) / 2;                      // uint4 words per 8-channel gradient vector
  __shared__ uint2 s_idx[3][PT_NP][PT_CG];
  __shared__ uint4 s_g[3][PT_NP][PT_CG * GW];
  int bid = blockIdx.x;
  const int chunk = bid % nchunk; bid /= nchunk;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h;
  const int b = bid / tiles_h;
  const int h0 = th * PT_TS, w0 = tw * PT_TS, c0 = chunk * (8 * PT_CG);
  const int tid = threadIdx.x, grp = tid & (PT_CG - 1), pos = tid >> 2;
  const int ih = pos >> 3, iw = pos & 7;
  const int h = h0 + ih, w = w0 + iw;
  const bool mine = h < d.Hi && w < d.Wi && c0 + grp * 8 < d.C;
  const TGO* __restrict__ gout = reinterpret_cast<const TGO*>(d.gout);
  TGI* __restrict__ gin = reinterpret_cast<TGI*>(d.gin);

  auto load_frame = [&](int t) {
    const int slot = (t + 3) % 3;
    for (int e = tid; e < PT_NP * PT_CG; e += 256) {
      const int p = e >> 2, g = e & (PT_CG - 1);
      const int oh = h0 - 1 + p / PT_HS, ow = w0 - 1 + p % PT_HS;
      const bool v = t >= 0 && t < d.To && oh >= 0 && oh < d.Ho && ow >= 0 && ow < d.Wo && c0 + g * 8 < d.C;
      uint2 ix = make_uint2(0xffffffffu, 0xffffffffu);
      uint4 gv[GW];
#pragma unroll
      for (int q = 0; q < GW; ++q) gv[q] = make_uint4(0u, 0u, 0u, 0u);
      if (v) {
        const int64_t o = (((int64_t)b * d.To + t) * d.Ho + oh) * d.Wo + ow;
        ix = __ldg(reinterpret_cast<const uint2*>(d.idx + o * d.C + c0 + g * 8));
        const uint4* gp = reinterpret_cast<const uint4*>(gout + o * d.ldgo + c0 + g * 8);
#pragma unroll
        for (int q = 0; q < GW; ++q) gv[q] = __ldg(gp + q);
      }
      s_idx[slot][p][g] = ix;
#pragma unroll
      for (int q = 0; q < GW; ++q) s_g[slot][p][g * GW + q] = gv[q];
    }
  };

  load_frame(-1);
  load_frame(0);
  for (int t = 0; t < d.Ti; ++t) {
    load_frame(t + 1);
    __syncthreads();
    if (mine) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      __nv_bfloat162 acc2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc2[q] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        const int slot = (t + 1 - dt + 3) % 3;        // window frame to = t + 1 - dt reads input frame t through tap dt
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const int pl = (ih + 2 - dh) * PT_HS + (iw + 2 - dw);
            const unsigned tap4 = (unsigned)((dt * 3 + dh) * 3 + dw) * 0x01010101u;
            const uint2 pk = s_idx[slot][pl][grp];
            if constexpr (GW == 1) {      // bf16 gradients: branch-free packed accumulation
              masked_add_bf16x8(acc2, s_g[slot][pl][grp], pk, tap4);
              continue;
            }
            const unsigned m0 = __vcmpeq4(pk.x, tap4), m1 = __vcmpeq4(pk.y, tap4);
            if ((m0 | m1) == 0u) continue;
            float g[8];
            if constexpr (GW == 1) {
              const uint4 u = s_g[slot][pl][grp];
              g[0] = bf16_lo(u.x); g[1] = bf16_hi(u.x); g[2] = bf16_lo(u.y); g[3] = bf16_hi(u.y);
              g[4] = bf16_lo(u.z); g[5] = bf16_hi(u.z); g[6] = bf16_lo(u.w); g[7] = bf16_hi(u.w);
            } else {
              const uint4 u0 = s_g[slot][pl][grp * 2], u1 = s_g[slot][pl][grp * 2 + 1];
              g[0] = __uint_as_float(u0.x); g[1] = __uint_as_float(u0.y); g[2] = __uint_as_float(u0.z); g[3] = __uint_as_float(u0.w);
              g[4] = __uint_as_float(u1.x); g[5] = __uint_as_float(u1.y); g[6] = __uint_as_float(u1.z); g[7] = __uint_as_float(u1.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if ((m0 >> (8 * e)) & 1u) acc[e] += g[e];
              if ((m1 >> (8 * e)) & 1u) acc[4 + e] += g[4 + e];
            }
          }
        }
      }
      if constexpr (GW == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[2 * q] = __low2float(acc2[q]); acc[2 * q + 1] = __high2float(acc2[q]); }
      }
      TGI* dst = gin + ((((int64_t)b * d.Ti + t) * d.Hi + h) * d.Wi + w) * d.ldgi + c0 + grp * 8;
      if (!d.gin_overwrite) {
        float o[8];
        load8(dst, o);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += o[e];
      }
      store8(dst, acc);
    }
    __syncthreads();
  }
}

static bool pool_bwd_tile333(const vinet_pool_t* d, cudaStream_t stream) {
  if (!d->idx || d->kt != 3 || d->kh != 3 || d->kw != 3 || d->st != 1 || d->sh != 1 || d->sw != 1 || d->pt != 1 || d->ph != 1 || d->pw != 1)
    return false;
  const int tiles_w = (int)cdiv(d->Wi, PT_TS), tiles_h = (int)cdiv(d->Hi, PT_TS), nchunk = (int)cdiv(d->C, 8 * PT_CG);
  const int64_t blocks = (int64_t)d->B * tiles_h * tiles_w * nchunk;
  if (blocks >= (int64_t)0x7fffffff) return false;
  VINET_DISPATCH_DTYPE(d->gout_dtype, TGO, VINET_DISPATCH_DTYPE(d->gin_dtype, TGI,
      (maxpool333_bwd_tile_kernel<TGO, TGI><<<(unsigned)blocks, 256, 0, stream>>>(*d, tiles_w, tiles_h, nchunk))));
  return true;
}
}  // namespace vinet

namespace vinet {
int g_pool_fast = 1;   // vinet_debug_set key 3: bit 0 = frame-walking 3x3x3 forward (production), bit 1 = force the generic gather
                       // backward, bit 2 = disable the compile-time-specialised gather backward (atomic scatter instead),
                       // bit 3 = use the specialised gather for every pool geometry (tests / A-B timing)
int pool_fast_set(int v) { g_pool_fast = v; return 0; }
}  // namespace vinet

template <int KT, int KH, int KW, int ST, int SH, int SW, int PT, int PH, int PW>
static bool pool_cfg_is(const vinet_pool_t* d) {
  return d->kt == KT && d->kh == KH && d->kw == KW && d->st == ST && d->sh == SH && d->sw == SW && d->pt == PT && d->ph == PH && d->pw == PW;
}

// returns true when one of the compile-time-specialised gather kernels took the launch
static bool pool_bwd_gather_fixed(const vinet_pool_t* d, cudaStream_t stream) {
  if (!d->idx || (int64_t)d->B * d->To * d->Ho * d->Wo >= (int64_t)0x7fffffff || (int64_t)d->B * d->Ti * d->Hi >= (int64_t)0x7fffffff) return false;
  const unsigned nb = (unsigned)std::min<int64_t>((int64_t)d->B * d->Ti * d->Hi, 148 * 32);
#define POOL_TRY(KT, KH, KW, ST, SH, SW, PT, PH, PW)                                                                          \
  if (pool_cfg_is<KT, KH, KW, ST, SH, SW, PT, PH, PW>(d)) {                                                                   \
    VINET_DISPATCH_DTYPE(d->gout_dtype, TGO, VINET_DISPATCH_DTYPE(d->gin_dtype, TGI,                                          \
        (maxpool_bwd_gather_fixed_kernel<KT, KH, KW, ST, SH, SW, PT, PH, PW, TGO, TGI><<<nb, 256, 0, stream>>>(*d))));        \
    return true;                                                                                                              \
  }
  // measured on B200 (tools/pool_bench.py, B=8 workload), gather vs atomic scatter: base1.1 256 vs 374 us, maxp2 225 vs 246 us;
  // the gather LOSES where an element has 8 or 27 candidate windows (maxp3 233 vs 112 us, Mixed_3c pool 857 vs 325 us), so only
  // the two (1,3,3)/(1,2,2) stage pools and the non-overlapping (2,2,2) pool take it.  g_pool_fast bit 3 forces it everywhere.
  POOL_TRY(1, 3, 3, 1, 2, 2, 0, 1, 1)   // base1.1, maxp2      (model.py:696,700)
  POOL_TRY(2, 2, 2, 2, 2, 2, 0, 0, 0)   // maxt4 + maxp4       (model.py:713-714)
  if (g_pool_fast & 8) {
    POOL_TRY(3, 3, 3, 2, 2, 2, 1, 1, 1)   // maxp3               (model.py:705)
    POOL_TRY(3, 3, 3, 1, 1, 1, 1, 1, 1)   // Mixed_* branch3     (model_utils.py:178)
  }
#undef POOL_TRY
  return false;
}

static unsigned pool_grid(const vinet_pool_t* d) {
  int64_t total = (int64_t)d->B * d->To * d->Ho * d->Wo * (d->C / 8);
  int64_t nb = cdiv(total, 256);
  if (nb > 148 * 64) nb = 148 * 64;
  return (unsigned)(nb < 1 ? 1 : nb);
}

static int pool_check(const vinet_pool_t* d) {
  VINET_CHECK(d->C % 8 == 0, "maxpool: C %d", d->C);
  VINET_CHECK(d->kt * d->kh * d->kw < 255, "maxpool: window of %d taps does not fit the byte index", d->kt * d->kh * d->kw);
  return 0;
}


static bool pool_is_333(const vinet_pool_t* d) {
  return (g_pool_fast & 1) && d->dtype == VINET_BF16 && d->out_dtype == VINET_BF16 && d->xform == VINET_XF_IDENT && d->kt == 3 && d->kh == 3 &&
         d->kw == 3 && d->st == 1 && d->sh == 1 && d->sw == 1 && d->pt == 1 && d->ph == 1 && d->pw == 1 && d->Hi >= 2 && d->Wi >= 2 &&
         (int64_t)d->B * d->Hi * d->Wi * (d->C / 8) < (int64_t)0x7fffffff;
}

extern "C" int vinet_maxpool_fwd(const vinet_pool_t* d, vinet_stream_t stream) {
  if (pool_check(d)) return -1;
  if (pool_is_333(d)) {  // frame-walking kernel: every 3x3 spatial maximum is computed once and used by three outputs
    const int64_t total = (int64_t)d->B * d->Hi * d->Wi * (d->C / 8);
    const unsigned nb = (unsigned)std::min<int64_t>(cdiv(total, 256), 148 * 64);
    // frames per thread: as long as possible (fewer re-computed boundary frames) while the launch still fills the machine twice
    int tc = d->Ti;
    while (tc > 2 && (int64_t)cdiv(d->Ti, tc) * total < (int64_t)2 * 148 * 2048) tc = (tc + 1) / 2;
    const dim3 grid(nb, (unsigned)cdiv(d->Ti, tc));
    if (d->idx) maxpool333_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(*d, tc);
    else maxpool333_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(*d, tc);
    VINET_LAUNCH_OK("maxpool333_fwd");
    return 0;
  }
  if (d->idx) {
    VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->out_dtype, TO,
        (maxpool_fwd_kernel<T, TO, true><<<pool_grid(d), 256, 0, (cudaStream_t)stream>>>(*d))));
  } else {
    VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->out_dtype, TO,
        (maxpool_fwd_kernel<T, TO, false><<<pool_grid(d), 256, 0, (cudaStream_t)stream>>>(*d))));
  }
  VINET_LAUNCH_OK("maxpool_fwd");
  return 0;
}

extern "C" int vinet_maxpool_bwd(const vinet_pool_t* d, vinet_stream_t stream) {
  if (pool_check(d)) return -1;
  // compile-time-specialised gathers for the pools of this model (key 3 bit 2 switches them off for A/B runs)
  if (!(g_pool_fast & 4) && !(g_pool_fast & 2) && pool_bwd_tile333(d, (cudaStream_t)stream)) {
    VINET_LAUNCH_OK("maxpool333_bwd_tile");
    return 0;
  }
  if (!(g_pool_fast & 4) && !(g_pool_fast & 2) && pool_bwd_gather_fixed(d, (cudaStream_t)stream)) {
    VINET_LAUNCH_OK("maxpool_bwd_gather_fixed");
    return 0;
  }
  // windows that can contain one input element: the gather visits all of them, so it only pays for strided pools
  const int cand = (int)(cdiv(d->kt, d->st) * cdiv(d->kh, d->sh) * cdiv(d->kw, d->sw));
  static const int gather_max = getenv("VINET_POOL_GATHER_MAX") ? atoi(getenv("VINET_POOL_GATHER_MAX")) : 0;
  if (d->idx && ((g_pool_fast & 2) || cand <= gather_max)) {   // gather over the recorded taps: no atomics, writes or accumulates every input element once
    const int64_t total = (int64_t)d->B * d->Ti * d->Hi * d->Wi * (d->C / 8);
    const unsigned nb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256), 148 * 64));
    VINET_DISPATCH_DTYPE(d->gout_dtype, TGO, VINET_DISPATCH_DTYPE(d->gin_dtype, TGI,
        (maxpool_bwd_gather_kernel<TGO, TGI><<<nb, 256, 0, (cudaStream_t)stream>>>(*d))));
    VINET_LAUNCH_OK("maxpool_bwd_gather");
    return 0;
  }
  if (d->gin_overwrite) {        // the scatter kernels accumulate: start from zero
    const size_t esz = d->gin_dtype == VINET_BF16 ? 2 : 4;
    VINET_CHECK(d->ldgi == d->C, "maxpool_bwd: gin_overwrite needs a dense gradient buffer for the scatter kernels");
    cudaMemsetAsync(d->gin, 0, (size_t)d->B * d->Ti * d->Hi * d->Wi * d->C * esz, (cudaStream_t)stream);
  }
  if (d->idx) {
    VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->gout_dtype, TGO, VINET_DISPATCH_DTYPE(d->gin_dtype, TGI,
        (maxpool_bwd_kernel<T, TGO, TGI, true><<<pool_grid(d), 256, 0, (cudaStream_t)stream>>>(*d)))));
  } else {
    VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->gout_dtype, TGO, VINET_DISPATCH_DTYPE(d->gin_dtype, TGI,
        (maxpool_bwd_kernel<T, TGO, TGI, false><<<pool_grid(d), 256, 0, (cudaStream_t)stream>>>(*d)))));
  }
  VINET_LAUNCH_OK("maxpool_bwd");
  return 0;
}
