// Layout conversion kernels: clip -> NDHWC, PyTorch conv weights -> GEMM B operand, packed wgrad -> PyTorch.
#include <algorithm>

#include "common.cuh"
#include "up2.cuh"

namespace vinet {

// ------------------------------------------------------------------ input clip
// One thread per (b,t,h,w): reads C strided fp32 values, writes cpad contiguous channels.
template <typename TO>
__global__ void pack_input_kernel(const __grid_constant__ vinet_pack_input_t d) {
  const int Wp = d.Wp > 0 ? d.Wp : d.W;
  const int64_t total = (int64_t)d.B * d.T * d.H * Wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int w = (int)(r % Wp) - d.wl; r /= Wp;
    const int h = (int)(r % d.H); r /= d.H;
    const int t = (int)(r % d.T);
    const int b = (int)(r / d.T);
    const bool in = (unsigned)w < (unsigned)d.W;
    const float* src = d.x + b * d.sb + t * d.st + h * d.sh + (in ? w : 0) * d.sw;
    TO* dst = reinterpret_cast<TO*>(d.out) + i * d.cpad;
    for (int c0 = 0; c0 < d.cpad && d.out != nullptr; c0 += 8) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (in && c0 + e < d.C) ? __ldg(src + (c0 + e) * d.sc) : 0.f;
      store8(dst + c0, v);
    }
    if (d.out4 != nullptr) {   // (bf16 only, C <= 4: checked by the host)
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (in && e < d.C) ? __ldg(src + e * d.sc) : 0.f;
      reinterpret_cast<uint2*>(d.out4)[i] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
  }
}

// Fast path of the common case (C = 3 -> 8 channels, unit stride along W, W % 4 == 0, bf16 output): one thread converts FOUR
// consecutive pixels - three 128-bit loads (one per colour plane) and four 128-bit stores - instead of 3 scalar loads per pixel;
// the first / last group of a row also writes the zero columns on its side.
__global__ void __launch_bounds__(256) pack_input_vec4_kernel(const __grid_constant__ vinet_pack_input_t d) {
  const int Wp = d.Wp > 0 ? d.Wp : d.W;
  const int G = d.W / 4;
  const int64_t total = (int64_t)d.B * d.T * d.H * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int gq = (int)(r % G); r /= G;
    const int h = (int)(r % d.H); r /= d.H;
    const int t = (int)(r % d.T);
    const int b = (int)(r / d.T);
    const float* src = d.x + b * d.sb + t * d.st + h * d.sh + gq * 4;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(src + d.sc));
    const float4 c2 = __ldg(reinterpret_cast<const float4*>(src + 2 * d.sc));
    if (d.out != nullptr) {
      uint4* row = reinterpret_cast<uint4*>(d.out) + (((int64_t)b * d.T + t) * d.H + h) * Wp;   // one uint4 = one 8-channel pixel
      uint4* dst = row + d.wl + gq * 4;
      dst[0] = make_uint4(pack_bf16x2(c0.x, c1.x), pack_bf16x2(c2.x, 0.f), 0u, 0u);
      dst[1] = make_uint4(pack_bf16x2(c0.y, c1.y), pack_bf16x2(c2.y, 0.f), 0u, 0u);
      dst[2] = make_uint4(pack_bf16x2(c0.z, c1.z), pack_bf16x2(c2.z, 0.f), 0u, 0u);
      dst[3] = make_uint4(pack_bf16x2(c0.w, c1.w), pack_bf16x2(c2.w, 0.f), 0u, 0u);
      if (gq == 0)
        for (int k = 0; k < d.wl; ++k) row[k] = make_uint4(0u, 0u, 0u, 0u);
      if (gq == G - 1)
        for (int k = d.wl + d.W; k < Wp; ++k) row[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (d.out4 != nullptr) {   // the 4-channel copy: one uint2 per pixel
      uint2* row4 = reinterpret_cast<uint2*>(d.out4) + (((int64_t)b * d.T + t) * d.H + h) * Wp;
      uint2* dst4 = row4 + d.wl + gq * 4;
      dst4[0] = make_uint2(pack_bf16x2(c0.x, c1.x), pack_bf16x2(c2.x, 0.f));
      dst4[1] = make_uint2(pack_bf16x2(c0.y, c1.y), pack_bf16x2(c2.y, 0.f));
      dst4[2] = make_uint2(pack_bf16x2(c0.z, c1.z), pack_bf16x2(c2.z, 0.f));
      dst4[3] = make_uint2(pack_bf16x2(c0.w, c1.w), pack_bf16x2(c2.w, 0.f));
      if (gq == 0)
        for (int k = 0; k < d.wl; ++k) row4[k] = make_uint2(0u, 0u);
      if (gq == G - 1)
        for (int k = d.wl + d.W; k < Wp; ++k) row4[k] = make_uint2(0u, 0u);
    }
  }
}

// ------------------------------------------------------------------ weights
// element (n, k) of the GEMM B operand; DENSE: k = tap_idx*cs + c, TAP64: k = tap_idx*round_up(cs,64) + c
__device__ __forceinline__ float weight_elem(const vinet_pack_t& d, int n, int k) {
  const int ldc = d.ld_cin > 0 ? d.ld_cin : d.Cin;   // channel extent of the underlying tensor (Cin may name a slice of it)
  if (d.layout == VINET_KLAYOUT_WIN8) {  // k = dh_tap*64 + dw*8 + ci; FPROP only
    const int tap = k >> 6, dw = (k >> 3) & 7, ci = k & 7;
    if (tap >= d.ntaps || dw >= d.kw || ci >= d.Cin || n >= d.Cout) return 0.f;
    const int dt = d.tap[tap][0], dh = d.tap[tap][1];
    return __ldg(d.w + ((((int64_t)n * ldc + ci) * d.kt + dt) * d.kh + dh) * d.kw + dw);
  }
  if (d.layout == VINET_KLAYOUT_WIN4) {  // k = dh_tap*64 + dw*4 + ci in the lower half of each 64-wide block; FPROP only
    const int tap = k >> 6, dw = (k >> 2) & 15, ci = k & 3;
    if (tap >= d.ntaps || dw >= d.kw || ci >= d.Cin || n >= d.Cout) return 0.f;
    const int dt = d.tap[tap][0], dh = d.tap[tap][1];
    return __ldg(d.w + ((((int64_t)n * ldc + ci) * d.kt + dt) * d.kh + dh) * d.kw + dw);
  }
  const int csk = (d.layout == VINET_KLAYOUT_TAP64) ? ((d.cs + 63) / 64) * 64 : d.cs;
  const int tap = k / csk, c = k - tap * csk;
  if (tap >= d.ntaps || c >= d.cs) return 0.f;
  int co, ci;
  if (d.mode == VINET_GATHER_FPROP) { co = n; ci = c; } else { co = c; ci = n; }
  if (co >= d.Cout || ci >= d.Cin) return 0.f;
  const int dt = d.tap[tap][0], dh = d.tap[tap][1], dw = d.tap[tap][2];
  return __ldg(d.w + ((((int64_t)co * ldc + ci) * d.kt + dt) * d.kh + dh) * d.kw + dw);
}

// term `part` of the bf16 expansion w ~= p0 + p1 + p2 (split-precision parity mode): the value left after removing the leading terms
__device__ __forceinline__ float weight_part(float v, int part) {
  for (int q = 0; q < part; ++q) v -= __bfloat162float(__float2bfloat16_rn(v));
  return v;
}

// Which (row nl, K block kb, N tile nt) the r-th 128-byte row of a pack is.  Tap-aligned layouts enumerate the TAP innermost:
// the threads of a warp / block that handle the taps of one (row, 64-channel block) read the same few sectors of the PyTorch
// weight (elements of one (co, ci) pair are its kt*kh*kw consecutive floats), so they are fetched once instead of once per tap
// by some other SM (measured: the every-step re-pack of all weights 0.35 ms -> see DESIGN.md).
__device__ __forceinline__ void pack_chunk_coords(const vinet_pack_t& d, int64_t r, int& nl, int& kb, int& nt) {
  if (d.layout != VINET_KLAYOUT_DENSE && d.ntaps > 1 && d.k_blocks % d.ntaps == 0) {
    const int ncb = d.k_blocks / d.ntaps;
    const int tap = (int)(r % d.ntaps); r /= d.ntaps;
    nl = (int)(r % d.block_n); r /= d.block_n;
    const int cb = (int)(r % ncb);
    nt = (int)(r / ncb);
    kb = tap * ncb + cb;
  } else {
    nl = (int)(r % d.block_n); r /= d.block_n;
    kb = (int)(r % d.k_blocks);
    nt = (int)(r / d.k_blocks);
  }
}

// The 8 bf16 of chunk j of row n of K block kb.  TAP64 (every pack but the stem's): the 8 elements share their tap and are 8
// consecutive channels, so ONE address computation and a constant stride replace eight general weight_elem() index chains
// (the every-step re-pack of all weights was instruction-bound: ~500 instructions per 16 output bytes).
__device__ __forceinline__ uint4 pack_chunk(const vinet_pack_t& d, int n, int kb, int j) {
  float v[8];
  if (d.layout == VINET_KLAYOUT_TAP64) {
    const int ncb = (d.cs + 63) >> 6;
    const int tap = kb / ncb, c0 = (kb - tap * ncb) * 64 + j * 8;
    const int ldc = d.ld_cin > 0 ? d.ld_cin : d.Cin;
    const int KT = d.kt * d.kh * d.kw;
    const int toff = (d.tap[tap][0] * d.kh + d.tap[tap][1]) * d.kw + d.tap[tap][2];
    const bool fprop = d.mode == VINET_GATHER_FPROP;
    const int nmax = fprop ? d.Cout : d.Cin, cmax = min(d.cs, fprop ? d.Cin : d.Cout);
    // FPROP: (co, ci) = (n, c); DGRAD: (co, ci) = (c, n)
    const float* p = d.w + (fprop ? ((int64_t)n * ldc + c0) : ((int64_t)c0 * ldc + n)) * KT + toff;
    const int64_t step = fprop ? KT : (int64_t)ldc * KT;
    const bool row_ok = tap < d.ntaps && n < nmax;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (row_ok && c0 + e < cmax) ? weight_part(__ldg(p + e * step), d.part) : 0.f;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = weight_part(weight_elem(d, n, kb * 64 + j * 8 + e), d.part);
  }
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// TC: one thread per 16-byte chunk (8 consecutive k) of [n_tiles][k_blocks][block_n][64], 128B-swizzled.
__global__ void pack_weights_tc_kernel(const __grid_constant__ vinet_pack_t d) {
  const int64_t chunks = (int64_t)d.n_tiles * d.k_blocks * d.block_n * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    int nl, kb, nt;
    pack_chunk_coords(d, i >> 3, nl, kb, nt);
    const int n = nt * d.block_n + nl;
    uint8_t* tile = reinterpret_cast<uint8_t*>(d.out) + ((int64_t)nt * d.k_blocks + kb) * d.block_n * 128;
    *reinterpret_cast<uint4*>(tile + nl * 128 + ((j ^ (nl & 7)) << 4)) = pack_chunk(d, n, kb, j);
  }
}

// Every cached packed weight of a model in ONE launch: weights change at every optimizer step, and ~150 separate pack
// launches of a few microseconds each would cost more than the bytes they move.  tab / begin live in device memory;
// begin[e] is the first 16-byte chunk of entry e in the concatenated chunk space.
__global__ void pack_weights_tc_multi_kernel(const vinet_pack_t* __restrict__ tab, const int64_t* __restrict__ begin, int n,
                                             int64_t total) {
  const int lane = threadIdx.x & 31;
  // (warp-uniform loop bound: the shuffles below need the whole warp; lanes past the end idle inside)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i - lane < total; i += (int64_t)gridDim.x * blockDim.x) {
    // entry of this chunk: the warp's first lane searches, the others only when they lie past that entry's end
    const int64_t i0 = i - lane;
    int lo = 0, hi = n - 1;
    if (lane == 0) {
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(begin + mid) <= i0) lo = mid; else hi = mid - 1;
      }
    }
    lo = __shfl_sync(0xffffffffu, lo, 0);
    if (i >= total) continue;
    if (lo + 1 < n && __ldg(begin + lo + 1) <= i) {
      hi = n - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(begin + mid) <= i) lo = mid; else hi = mid - 1;
      }
    }
    const vinet_pack_t& d = tab[lo];
    const int64_t li = i - __ldg(begin + lo);
    const int j = (int)(li & 7);
    int nl, kb, nt;
    pack_chunk_coords(d, li >> 3, nl, kb, nt);
    const int nn = nt * d.block_n + nl;
    uint8_t* tile = reinterpret_cast<uint8_t*>(d.out) + ((int64_t)nt * d.k_blocks + kb) * d.block_n * 128;
    *reinterpret_cast<uint4*>(tile + nl * 128 + ((j ^ (nl & 7)) << 4)) = pack_chunk(d, nn, kb, j);
  }
}

// SIMT: fp32 [k_blocks*64][npad]
__global__ void pack_weights_simt_kernel(const __grid_constant__ vinet_pack_t d, int npad) {
  const int64_t total = (int64_t)d.k_blocks * 64 * npad;
  float* out = reinterpret_cast<float*>(d.out);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % npad);
    const int k = (int)(i / npad);
    out[i] = weight_elem(d, n, k);
  }
}

__global__ void unpack_wgrad_kernel(float* __restrict__ dwp, int lddw, int cs, float* __restrict__ grad, int Cout,
                                    int Cin, int ntaps) {
  const int64_t total = (int64_t)Cout * Cin * ntaps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % ntaps);
    int64_t r = i / ntaps;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    grad[i] = dwp[((int64_t)tap * cs + ci) * lddw + co];
  }
}

__global__ void unpack_wgrad_win8_kernel(float* __restrict__ dwp, int lddw, float* __restrict__ grad, int Cout, int Cin,
                                         int kh, int kw) {
  const int64_t total = (int64_t)Cout * Cin * kh * kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int dw = (int)(i % kw);
    int64_t r = i / kw;
    const int dh = (int)(r % kh); r /= kh;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    grad[i] = dwp[((int64_t)dh * 64 + dw * 8 + ci) * lddw + co];
  }
}

// ------------------------------------------------------------------ split-precision operands (parity mode)
// One thread per 8 channels of one row: x -> (pending transform) -> bf16 expansion planes.
template <typename T>
__global__ void __launch_bounds__(256) split_bf16_kernel(const __grid_constant__ vinet_split_t d) {
  const T* __restrict__ x = reinterpret_cast<const T*>(d.x);
  const int G = d.C / 8;
  const int64_t total = d.rows * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / G;
    const int c = (int)(i - r * G) * 8;
    float v[8];
    if (d.xform & VINET_XF_UP2) {   // r enumerates hi-res pixels (n, Y, X) of low-res frames [up_h, up_w, ld]
      const int W2 = 2 * d.up_w, H2 = 2 * d.up_h;
      const int X = (int)(r % W2);
      const int64_t q = r / W2;
      const int Y = (int)(q % H2);
      const int64_t n = q / H2;
      up2_load8(x + n * d.up_h * d.up_w * d.ld + c, d.up_h, d.up_w, d.ld, Y, X, (d.xform & 1) != 0, v);
    } else {
      load8(x + r * d.ld + c, v);
      apply_xform<8>(v, d.xform, d.scale, d.shift, c);
    }
    for (int p = 0; p < d.nparts; ++p) {
      float h[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
        v[e] -= h[e];
      }
      store8(reinterpret_cast<__nv_bfloat16*>(d.part[p]) + r * d.ldo + c, h);
    }
  }
}

struct UnpackMulti {
  vinet_unpack_t e[VINET_UNPACK_MAX];
  int32_t n, total;
};

__global__ void __launch_bounds__(256) unpack_wgrad_multi_kernel(const __grid_constant__ UnpackMulti p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = p.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (p.e[mid].begin <= i) lo = mid; else hi = mid - 1;
    }
    const vinet_unpack_t& d = p.e[lo];
    const int li = i - d.begin;
    if (d.win8_kh == 0) {
      const int tap = li % d.ntaps;
      const int r = li / d.ntaps;
      const int ci = r % d.Cin, co = r / d.Cin;
      d.grad[li] = d.dwp[((int64_t)tap * d.cs + ci) * d.lddw + co];
    } else {
      const int dw = li % d.win8_kw;
      int r = li / d.win8_kw;
      const int dh = r % d.win8_kh; r /= d.win8_kh;
      const int ci = r % d.Cin, co = r / d.Cin;
      const int cpp = d.cs > 0 ? d.cs : 8;     // channels per pixel of the window layout: WIN8 (default) or WIN4
      d.grad[li] = d.dwp[((int64_t)dh * 64 + dw * cpp + ci) * d.lddw + co];
    }
  }
}

static inline unsigned grid_for(int64_t n, int block) {
  int64_t g = cdiv(n, block);
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace vinet

using namespace vinet;

extern "C" int vinet_pack_input(const vinet_pack_input_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->cpad % 8 == 0 && d->cpad >= d->C, "pack_input: cpad %d", d->cpad);
  VINET_CHECK(d->Wp == 0 || d->Wp >= d->wl + d->W, "pack_input: Wp %d < wl %d + W %d", d->Wp, d->wl, d->W);
  VINET_CHECK(d->out != nullptr || d->out4 != nullptr, "pack_input: no output");
  VINET_CHECK(d->out4 == nullptr || (d->out_dtype == VINET_BF16 && d->C <= 4 && (d->Wp > 0 ? d->Wp : d->W) % 2 == 0 &&
                                     (reinterpret_cast<uintptr_t>(d->out4) & 15) == 0),
              "pack_input: the 4-channel copy needs bf16 output, C <= 4 and an even row width");
  const int64_t total = (int64_t)d->B * d->T * d->H * (d->Wp > 0 ? d->Wp : d->W);
  const bool vec4 = d->out_dtype == VINET_BF16 && d->C == 3 && d->cpad == 8 && d->sw == 1 && d->W % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && d->sb % 4 == 0 && d->sc % 4 == 0 && d->st % 4 == 0 && d->sh % 4 == 0;
  if (vec4) {
    pack_input_vec4_kernel<<<grid_for((int64_t)d->B * d->T * d->H * (d->W / 4), 256), 256, 0, (cudaStream_t)stream>>>(*d);
    VINET_LAUNCH_OK("pack_input_vec4");
    return 0;
  }
  VINET_DISPATCH_DTYPE(d->out_dtype, TO, (pack_input_kernel<TO><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*d)));
  VINET_LAUNCH_OK("pack_input");
  return 0;
}

extern "C" int vinet_split_bf16(const vinet_split_t* d, vinet_stream_t stream) {
  VINET_CHECK(d && d->C % 8 == 0 && d->C >= 8 && d->rows >= 1, "split_bf16: C %d rows %lld", d ? d->C : -1, d ? (long long)d->rows : -1ll);
  VINET_CHECK(d->nparts >= 1 && d->nparts <= 3 && d->ldo >= d->C && d->ldo % 8 == 0, "split_bf16: nparts %d ldo %lld", d->nparts,
              (long long)d->ldo);
  for (int p = 0; p < d->nparts; ++p) VINET_CHECK(d->part[p] != nullptr, "split_bf16: part %d is null", p);
  if (d->xform & VINET_XF_UP2)
    VINET_CHECK(!(d->xform & 2) && d->up_h >= 1 && d->up_w >= 1 && d->rows % (4ll * d->up_h * d->up_w) == 0,
                "split_bf16: up-sampled source needs up_h/up_w with rows a multiple of 4*up_h*up_w, and no affine transform");
  const int64_t total = d->rows * (d->C / 8);
  VINET_DISPATCH_DTYPE(d->dtype, T, (split_bf16_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*d)));
  VINET_LAUNCH_OK("split_bf16");
  return 0;
}

extern "C" size_t vinet_packed_weight_bytes(int32_t engine, int32_t N, int32_t block_n, int32_t n_tiles, int32_t k_blocks) {
  if (engine == VINET_ENGINE_TC) return (size_t)n_tiles * k_blocks * block_n * 128;
  return (size_t)k_blocks * 64 * round_up(N, 64) * sizeof(float);
}

extern "C" int vinet_pack_weights(const vinet_pack_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->ntaps <= VINET_MAX_TAPS && d->cs % 8 == 0, "pack_weights: ntaps %d cs %d", d->ntaps, d->cs);
  VINET_CHECK(d->layout >= VINET_KLAYOUT_DENSE && d->layout <= VINET_KLAYOUT_WIN4, "pack_weights: layout %d", d->layout);
  VINET_CHECK(d->layout != VINET_KLAYOUT_WIN4 || (d->mode == VINET_GATHER_FPROP && d->Cin <= 4 && d->kw <= 8 && (d->cs == 32 || d->cs == 64)),
              "pack_weights: WIN4 needs an FPROP pack with Cin <= 4, kw <= 8, cs 32 (8-pixel windows) or 64 (16-pixel windows)");
  VINET_CHECK(d->part >= 0 && d->part <= 2 && (d->part == 0 || d->engine == VINET_ENGINE_TC), "pack_weights: part %d", d->part);
  VINET_CHECK(d->ld_cin == 0 || d->ld_cin >= d->Cin, "pack_weights: ld_cin %d < Cin %d", d->ld_cin, d->Cin);
  VINET_CHECK(d->layout != VINET_KLAYOUT_WIN8 || (d->mode == VINET_GATHER_FPROP && d->Cin <= 8 && d->kw <= 8 && d->cs == 64),
              "pack_weights: WIN8 needs an FPROP pack with Cin <= 8, kw <= 8, cs == 64");
  const int64_t csk = d->layout == VINET_KLAYOUT_DENSE ? d->cs : round_up(d->cs, 64);
  VINET_CHECK((int64_t)d->k_blocks * 64 >= (int64_t)d->ntaps * csk, "pack_weights: k_blocks too small");
  if (d->engine == VINET_ENGINE_TC) {
    const int64_t chunks = (int64_t)d->n_tiles * d->k_blocks * d->block_n * 8;
    pack_weights_tc_kernel<<<grid_for(chunks, 256), 256, 0, (cudaStream_t)stream>>>(*d);
  } else {
    const int N = (d->mode == VINET_GATHER_FPROP) ? d->Cout : d->Cin;
    const int npad = (int)round_up(N, 64);
    pack_weights_simt_kernel<<<grid_for((int64_t)d->k_blocks * 64 * npad, 256), 256, 0, (cudaStream_t)stream>>>(*d, npad);
  }
  VINET_LAUNCH_OK("pack_weights");
  return 0;
}

extern "C" int vinet_pack_weights_multi(const vinet_pack_t* table_dev, const int64_t* chunk_begin_dev, int32_t n, int64_t total_chunks,
                                        vinet_stream_t stream) {
  VINET_CHECK(table_dev && chunk_begin_dev && n >= 1 && total_chunks >= 1, "pack_weights_multi: empty table");
  unsigned nb = (unsigned)std::min<int64_t>(cdiv(total_chunks, 256), 148 * 16);
  pack_weights_tc_multi_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(table_dev, chunk_begin_dev, n, total_chunks);
  VINET_LAUNCH_OK("pack_weights_multi");
  return 0;
}

extern "C" int vinet_unpack_wgrad_win8(float* dwp, int32_t lddw, float* grad, int32_t Cout, int32_t Cin, int32_t kh,
                                       int32_t kw, vinet_stream_t stream) {
  VINET_CHECK(Cin <= 8 && kw <= 8, "unpack_wgrad_win8: Cin %d kw %d", Cin, kw);
  const int64_t total = (int64_t)Cout * Cin * kh * kw;
  unpack_wgrad_win8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dwp, lddw, grad, Cout, Cin, kh, kw);
  VINET_LAUNCH_OK("unpack_wgrad_win8");
  return 0;
}

extern "C" int vinet_unpack_wgrad_multi(const vinet_unpack_t* d, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(d && n >= 1 && n <= VINET_UNPACK_MAX, "unpack_wgrad_multi: %d entries", n);
  UnpackMulti p;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    p.e[i] = d[i];
    p.e[i].begin = (int32_t)total;
    total += (int64_t)d[i].Cout * d[i].Cin * (d[i].win8_kh ? d[i].win8_kh * d[i].win8_kw : d[i].ntaps);
    VINET_CHECK(total < (int64_t)0x7fffffff, "unpack_wgrad_multi: too many elements");
  }
  p.n = n;
  p.total = (int32_t)total;
  unpack_wgrad_multi_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
  VINET_LAUNCH_OK("unpack_wgrad_multi");
  return 0;
}

extern "C" int vinet_unpack_wgrad(float* dwp, int32_t lddw, int32_t cs, float* grad, int32_t Cout, int32_t Cin,
                                  int32_t ntaps, vinet_stream_t stream) {
  const int64_t total = (int64_t)Cout * Cin * ntaps;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dwp, lddw, cs, grad, Cout, Cin, ntaps);
  VINET_LAUNCH_OK("unpack_wgrad");
  return 0;
}
