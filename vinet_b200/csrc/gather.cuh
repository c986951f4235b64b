// Implicit-GEMM operand gather shared by the SIMT and tcgen05 conv kernels.
// Semantics are defined in include/vinet_b200.h (vinet_gather_t).
#pragma once
#include "common.cuh"
#include "up2.cuh"

namespace vinet {

struct RowCoord {
  int b, t, h, w;  // b < 0: row beyond the problem (reads as zero)
};

__device__ __forceinline__ int64_t gather_rows(const vinet_gather_t& g) {
  return (int64_t)g.B * g.Tr * g.Hr * g.Wr;
}

__device__ __forceinline__ RowCoord decode_row(const vinet_gather_t& g, int64_t row, int64_t M) {
  RowCoord rc;
  if (row >= M) {
    rc.b = -1; rc.t = rc.h = rc.w = 0;
    return rc;
  }
  rc.w = (int)(row % g.Wr); row /= g.Wr;
  rc.h = (int)(row % g.Hr); row /= g.Hr;
  int tr = (int)(row % g.Tr);
  rc.b = (int)(row / g.Tr);
  rc.t = tr * g.row_tstep + g.row_toff;
  return rc;
}

// Locate the source position of (row, tap): source index si, its local frame ts and the position (hs, ws) inside the
// frame. Returns false for zero fill.
__device__ __forceinline__ bool gather_coords(const vinet_gather_t& g, const RowCoord& rc, int tap, int& si, int& ts, int& hs,
                                              int& ws) {
  if (rc.b < 0 || tap >= g.ntaps) return false;
  const int dt = g.tap[tap][0], dh = g.tap[tap][1], dw = g.tap[tap][2];
  if (g.mode == VINET_GATHER_FPROP) {
    ts = rc.t * g.st - g.pt + dt;
    hs = rc.h * g.sh - g.ph + dh;
    ws = rc.w * g.sw - g.pw + dw;
  } else {
    int nt = rc.t + g.pt - dt, nh = rc.h + g.ph - dh, nw = rc.w + g.pw - dw;
    if (nt < 0 || nh < 0 || nw < 0) return false;
    ts = nt / g.st; hs = nh / g.sh; ws = nw / g.sw;
    if (ts * g.st != nt || hs * g.sh != nh || ws * g.sw != nw) return false;
  }
  if ((unsigned)ts >= (unsigned)g.Ts || (unsigned)hs >= (unsigned)g.Hs || (unsigned)ws >= (unsigned)g.Ws)
    return false;
  si = (ts >= g.src[0].T) ? 1 : 0;
  if (si) ts -= g.src[0].T;
  return true;
}

// Locate the source element (b, tap-shifted position, channel 0). Returns false for zero fill.
__device__ __forceinline__ bool gather_locate(const vinet_gather_t& g, const RowCoord& rc, int tap, int& si,
                                              int64_t& off) {
  int ts, hs, ws;
  if (!gather_coords(g, rc, tap, si, ts, hs, ws)) return false;
  off = ((((int64_t)rc.b * g.src[si].T + ts) * g.Hs + hs) * g.Ws + ws) * g.src[si].ld;
  return true;
}

// Gather V (4 or 8) consecutive K elements starting at flattened index k = tap*Cs + c (k % V == 0).
template <typename T, int V>
__device__ __forceinline__ void gather_vec(const vinet_gather_t& g, const RowCoord& rc, int k, float (&v)[V]) {
  const int tap = k / g.Cs;
  const int c = k - tap * g.Cs;
  int si, ts, hs, ws;
  if (!gather_coords(g, rc, tap, si, ts, hs, ws)) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
    return;
  }
  const vinet_src_t& s = g.src[si];
  if (s.xform & VINET_XF_UP2) {
    // the source is stored at half resolution and read through the 2x bilinear up-sampling (after its ReLU)
    const int h = g.Hs >> 1, w = g.Ws >> 1;
    const T* frame = reinterpret_cast<const T*>(s.ptr) + ((int64_t)rc.b * s.T + ts) * h * w * s.ld + c;
    if constexpr (V == 8) up2_load8(frame, h, w, s.ld, hs, ws, (s.xform & 1) != 0, v);
    else up2_load4(frame, h, w, s.ld, hs, ws, (s.xform & 1) != 0, v);
    if constexpr (sizeof(T) == 2) {   // bf16 storage: round where the materialised up-sampled tensor would have been rounded
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
    }
    return;
  }
  const int64_t off = ((((int64_t)rc.b * s.T + ts) * g.Hs + hs) * g.Ws + ws) * s.ld;
  const T* p = reinterpret_cast<const T*>(s.ptr) + off + c;
  if constexpr (V == 8) load8(p, v); else load4(p, v);
  apply_xform<V>(v, s.xform, s.scale, s.shift, c);
  if constexpr (sizeof(T) == 2) {
    // bf16 storage: the tensor-core engine feeds bf16 operands, so the fp32-FFMA cross-check engine rounds the
    // transformed value at the same point (otherwise ReLU masks downstream differ by more than accumulation order)
    if (s.xform & 2) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
    }
  }
}

// output row -> destination pointer (two destinations split on the frame index)
__device__ __forceinline__ int out_index(const vinet_conv_t& d, const RowCoord& rc) {
  return (rc.t >= d.out_T[0]) ? 1 : 0;
}
template <typename TO>
__device__ __forceinline__ TO* out_row_ptr(const vinet_conv_t& d, const RowCoord& rc) {
  int i = (rc.t >= d.out_T[0]) ? 1 : 0;
  int t = rc.t - (i ? d.out_T[0] : 0);
  int64_t pos = (((int64_t)rc.b * d.out_T[i] + t) * d.g.Hr + rc.h) * d.g.Wr + rc.w;
  return reinterpret_cast<TO*>(d.out[i]) + pos * d.ldo[i];
}

// Epilogue of 8 consecutive output channels [n, n+8) of one row: optional per-channel scale/shift/activation
// (EPI; vector loads, branches hoisted out of the element loop), optional read-modify-write, one 16/32-byte store.
template <typename TO, bool EPI>
__device__ __forceinline__ void epilogue_store8(const vinet_conv_t& d, TO* p, const uint32_t* r, int n, bool accum) {
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]);
  if constexpr (EPI) {
    if (d.ep_scale) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(d.ep_scale + n));
      const float4 b = __ldg(reinterpret_cast<const float4*>(d.ep_scale + n) + 1);
      v[0] *= a.x; v[1] *= a.y; v[2] *= a.z; v[3] *= a.w; v[4] *= b.x; v[5] *= b.y; v[6] *= b.z; v[7] *= b.w;
    }
    if (d.ep_shift) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(d.ep_shift + n));
      const float4 b = __ldg(reinterpret_cast<const float4*>(d.ep_shift + n) + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (d.ep_act == VINET_ACT_RELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    } else if (d.ep_act == VINET_ACT_SIGMOID) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 1.f / (1.f + __expf(-v[e]));
    }
  }
  if (accum) {
    float o[8];
    load8(p, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += o[e];
  }
  store8(p, v);
}
// 16 consecutive output channels [n, n+16) of one row, nvalid = channels that exist (columns up to the N tile's edge).
// bf16 rows whose 16-column group is 32-byte aligned go out as ONE 256-bit store (sm_100 STG.256): an epilogue lane owns a row, so
// every lane of a store instruction hits a different 128-byte line - two 16-byte stores per group made every 32-byte sector
// arrive at the L2 as two partial writes, and the L1/L2 write path (not HBM) bound the 64-channel layers (ncu, stem conv_s:
// 44 M sector writes for 22 M sectors).  The read-modify-write path loads the same way.
template <typename TO, bool EPI>
__device__ __forceinline__ void epilogue_store16(const vinet_conv_t& d, TO* p, const uint32_t* r, int n, bool accum, int nvalid) {
  if constexpr (sizeof(TO) == 2) {
    if (nvalid >= 16 && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
      if constexpr (EPI) {     // per-channel scale / shift as 128-bit loads (n is a multiple of 16), activation on registers
        if (d.ep_scale) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(d.ep_scale + n) + q4);
            v[4 * q4] *= a.x; v[4 * q4 + 1] *= a.y; v[4 * q4 + 2] *= a.z; v[4 * q4 + 3] *= a.w;
          }
        }
        if (d.ep_shift) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(d.ep_shift + n) + q4);
            v[4 * q4] += a.x; v[4 * q4 + 1] += a.y; v[4 * q4 + 2] += a.z; v[4 * q4 + 3] += a.w;
          }
        }
        if (d.ep_act == VINET_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
        } else if (d.ep_act == VINET_ACT_SIGMOID) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 1.f / (1.f + __expf(-v[e]));
        }
      }
      if (accum) {
        uint32_t o[8];
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
                     : "l"(p)
                     : "memory");
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[2 * e] += bf16_lo(o[e]); v[2 * e + 1] += bf16_hi(o[e]); }
      }
      uint32_t u[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) u[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
                   "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                   : "memory");
      return;
    }
  }
  if (nvalid > 0) epilogue_store8<TO, EPI>(d, p, r, n, accum);
  if (nvalid > 8) epilogue_store8<TO, EPI>(d, p + 8, r + 8, n + 8, accum);
}

__host__ __device__ static inline bool conv_has_epilogue(const vinet_conv_t& d) {
  return d.ep_scale != nullptr || d.ep_shift != nullptr || d.ep_act != VINET_ACT_NONE;
}

__device__ __forceinline__ float epilogue_value(const vinet_conv_t& d, float acc, int n) {
  if (d.ep_scale) acc *= __ldg(d.ep_scale + n);
  if (d.ep_shift) acc += __ldg(d.ep_shift + n);
  if (d.ep_act == VINET_ACT_RELU) acc = fmaxf(acc, 0.f);
  else if (d.ep_act == VINET_ACT_SIGMOID) acc = 1.f / (1.f + __expf(-acc));
  return acc;
}

}  // namespace vinet
