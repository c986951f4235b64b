// nn.Upsample(scale_factor=(1,2,2), mode='trilinear', align_corners=False) (model.py:254) as a READ transform
// (VINET_XF_UP2): per-frame 2x bilinear with taps {.25,.75} and index clamp at the borders.
// ONE definition of the arithmetic, shared by the standalone kernel (upsample.cu), the FFMA gather (gather.cuh), the
// interpolating producer warps of the tcgen05 kernels (conv_stream.cu, conv_wgrad_halo.cu), the operand split (pack.cu)
// and the head (head.cu): every consumer sees bit-identical values, so fused and materialised plans agree exactly.
#pragma once
#include "common.cuh"

namespace vinet {

// source taps of output coordinate Y on an axis of length n: indices i0,i1 and the weight of i1
__device__ __forceinline__ void up_taps(int Y, int n, int& i0, int& i1, float& l1) {
  float src = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f);
  i0 = (int)src;
  i1 = min(i0 + 1, n - 1);
  l1 = src - (float)i0;
}

// blend of the four neighbours (a: y0x0, b: y0x1, e0: y1x0, e1: y1x1), in the order nn.Upsample's kernel uses
// (explicit mul / fma intrinsics: every user of this header must round identically, whatever contraction the compiler prefers)
__device__ __forceinline__ float up_lerp(float a, float b, float l) { return __fmaf_rn(l, b, __fmul_rn(1.f - l, a)); }
__device__ __forceinline__ float up_blend(float a, float b, float e0, float e1, float lx, float ly) {
  return up_lerp(up_lerp(a, b, lx), up_lerp(e0, e1, lx), ly);
}

// 8 consecutive channels of hi-res pixel (Y, X) of one low-res frame [h, w, ld] (frame = pointer to its first element + channel)
template <typename T>
__device__ __forceinline__ void up2_load8(const T* __restrict__ frame, int h, int w, int64_t ld, int Y, int X, bool relu,
                                          float (&o)[8]) {
  int y0, y1, x0, x1;
  float ly, lx;
  up_taps(Y, h, y0, y1, ly);
  up_taps(X, w, x0, x1, lx);
  float a[8], b[8], e0[8], e1[8];
  load8(frame + ((int64_t)y0 * w + x0) * ld, a);
  load8(frame + ((int64_t)y0 * w + x1) * ld, b);
  load8(frame + ((int64_t)y1 * w + x0) * ld, e0);
  load8(frame + ((int64_t)y1 * w + x1) * ld, e1);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (relu) { a[e] = fmaxf(a[e], 0.f); b[e] = fmaxf(b[e], 0.f); e0[e] = fmaxf(e0[e], 0.f); e1[e] = fmaxf(e1[e], 0.f); }
    o[e] = up_blend(a[e], b[e], e0[e], e1[e], lx, ly);
  }
}
template <typename T>
__device__ __forceinline__ void up2_load4(const T* __restrict__ frame, int h, int w, int64_t ld, int Y, int X, bool relu,
                                          float (&o)[4]) {
  int y0, y1, x0, x1;
  float ly, lx;
  up_taps(Y, h, y0, y1, ly);
  up_taps(X, w, x0, x1, lx);
  float a[4], b[4], e0[4], e1[4];
  load4(frame + ((int64_t)y0 * w + x0) * ld, a);
  load4(frame + ((int64_t)y0 * w + x1) * ld, b);
  load4(frame + ((int64_t)y1 * w + x0) * ld, e0);
  load4(frame + ((int64_t)y1 * w + x1) * ld, e1);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (relu) { a[e] = fmaxf(a[e], 0.f); b[e] = fmaxf(b[e], 0.f); e0[e] = fmaxf(e0[e], 0.f); e1[e] = fmaxf(e1[e], 0.f); }
    o[e] = up_blend(a[e], b[e], e0[e], e1[e], lx, ly);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Interpolating producer of the TMA-fed tensor-core kernels: builds, in shared memory, exactly the tile a
// cp.async.bulk.tensor box load of the (never materialised) up-sampled bf16 tensor would have written:
//   box = {64 channels, PW, PH} with hi-res origin (X0, Y0); pixel (px, py) is the 128-byte row r = py*PW + px, whose
//   16-byte chunk j sits at  tile + r*128 + ((j ^ (r & 7)) << 4)   (SWIZZLE_128B; tile is 1024-byte aligned);
//   pixels outside the hi-res image are zero (the convolution's zero padding), as TMA's out-of-bounds fill does.
// X0, Y0 odd and PW, PH even (true for pad-1 3x3 halos of 8- and 16-aligned tiles): the box splits into 2x2 quads
// {2k+1, 2k+2} x {2m+1, 2m+2} whose four pixels blend the SAME four low-res neighbours (k, k+1) x (m, m+1): four 16-byte
// loads feed four 16-byte shared-memory stores.  Values are computed by up_taps / up_blend, i.e. bit-identical to upsample.cu.
// `frame` points at channel block start of low-res frame [h, w, ld]; nchunk = 16-byte chunks to fill (2 * k-steps the MMA reads),
// cvalid = chunks that hold real channels (the rest are zero-filled); tid / nthr enumerate the cooperating threads.
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- packed fp32 pairs (sm_100 FMUL2 / FFMA2: two IEEE round-to-nearest operations per lane and instruction, so the blend of a
// 16-byte chunk costs half the issue slots and still rounds exactly like up_lerp) ----
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t up_lerp2(f32x2_t a, f32x2_t b, float l) {   // == (up_lerp(a.lo, b.lo, l), up_lerp(a.hi, b.hi, l))
  const f32x2_t l2 = f2_pack(l, l), m2 = f2_pack(1.f - l, 1.f - l);
  f32x2_t t, r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(a), "l"(m2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(l2), "l"(t));
  return r;
}
// relu on packed bf16 pairs, then widen: the 8 channels of one 16-byte chunk as 4 fp32 pairs
__device__ __forceinline__ void up2_unpack8(uint4 u, bool relu, f32x2_t (&v)[4]) {
  if (relu) {
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
  }
  v[0] = f2_pack(bf16_lo(u.x), bf16_hi(u.x)); v[1] = f2_pack(bf16_lo(u.y), bf16_hi(u.y));
  v[2] = f2_pack(bf16_lo(u.z), bf16_hi(u.z)); v[3] = f2_pack(bf16_lo(u.w), bf16_hi(u.w));
}
__device__ __forceinline__ uint4 up2_pack8(const f32x2_t (&v)[4]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    f2_unpack(v[i], lo, hi);
    w[i] = pack_bf16x2(lo, hi);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct Up2Quad {      // one task: chunk j of quad (qx, qy)
  int qx, qy;
  bool any;
  uint4 raw[4];       // low-res neighbours (ya,xa) (ya,xb) (yb,xa) (yb,xb)
};

__device__ __forceinline__ void up2_fill_box(uint32_t tile, const __nv_bfloat16* __restrict__ frame, bool frame_valid, int h, int w,
                                             int64_t ld, int X0, int Y0, int PW, int PH, int nchunk, int cvalid, bool relu, int tid,
                                             int nthr) {
  const int QW = PW >> 1, nquad = QW * (PH >> 1);
  const int j = tid & 7;                                  // this thread's 16-byte chunk (nthr is a multiple of 8)
  if (j >= nchunk) return;
  const bool chan_ok = frame_valid && j < cvalid;
  const int kx0 = (X0 - 1) >> 1, ky0 = (Y0 - 1) >> 1;     // low-res column / row left of / above the box's first quad (may be -1)
  const int step = nthr >> 3, step_y = step / QW, step_x = step - step_y * QW;
  const __nv_bfloat16* f = frame + j * 8;
  const uint32_t jb = (uint32_t)j;
  int qi = tid >> 3;
  int qy = qi / QW, qx = qi - qy * QW;
  constexpr int NP = 2;   // quads in flight per thread: their 4 + 4 loads are issued before either is consumed (L2 latency)
  while (qi < nquad) {
    Up2Quad t[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) {
      t[u].qx = qx; t[u].qy = qy;
      const int ky = ky0 + qy, kx = kx0 + qx;
      const bool live = qi < nquad;
      t[u].any = live && chan_ok && ky >= -1 && ky < h && kx >= -1 && kx < w;
      if (!live) t[u].qx = -1;
      if (t[u].any) {
        // low-res neighbours shared by the whole quad, clamped to the image (nn.Upsample's index clamp)
        const int ya = max(ky, 0), yb = min(ky + 1, h - 1), xa = max(kx, 0), xb = min(kx + 1, w - 1);
        const __nv_bfloat16* ra = f + (int64_t)ya * w * ld;
        const __nv_bfloat16* rb = f + (int64_t)yb * w * ld;
        t[u].raw[0] = __ldg(reinterpret_cast<const uint4*>(ra + xa * ld));
        t[u].raw[1] = __ldg(reinterpret_cast<const uint4*>(ra + xb * ld));
        t[u].raw[2] = __ldg(reinterpret_cast<const uint4*>(rb + xa * ld));
        t[u].raw[3] = __ldg(reinterpret_cast<const uint4*>(rb + xb * ld));
      }
      qi += step; qx += step_x; qy += step_y;
      if (qx >= QW) { qx -= QW; ++qy; }
    }
#pragma unroll
    for (int u = 0; u < NP; ++u) {
      if (t[u].qx < 0) continue;
      const int px = 2 * t[u].qx, py = 2 * t[u].qy;       // first pixel of the quad inside the box
      uint4 out[2][2];
      if (t[u].any) {
        const int ky = ky0 + t[u].qy, kx = kx0 + t[u].qx;
        // Hi-res rows 2ky+1 / 2ky+2 blend low rows (ky, ky+1) with weights .25 / .75 of the second one (up_taps); on the clamped
        // borders the two candidates coincide, and row 0 (ky = -1) takes weight 0 as up_taps does: values are bit-identical to
        // up2_load8 / upsample.cu.  Rows / columns outside the hi-res image are the convolution's zero padding.
        const float ly1 = ky < 0 ? 0.f : 0.75f, lx1 = kx < 0 ? 0.f : 0.75f;
        f32x2_t a[4], b[4], tx0[4], tx1[4], bx0[4], bx1[4], o[4];
        up2_unpack8(t[u].raw[0], relu, a);
        up2_unpack8(t[u].raw[1], relu, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) { tx0[e] = up_lerp2(a[e], b[e], 0.25f); tx1[e] = up_lerp2(a[e], b[e], lx1); }
        up2_unpack8(t[u].raw[2], relu, a);
        up2_unpack8(t[u].raw[3], relu, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) { bx0[e] = up_lerp2(a[e], b[e], 0.25f); bx1[e] = up_lerp2(a[e], b[e], lx1); }
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = up_lerp2(tx0[e], bx0[e], 0.25f);
        out[0][0] = up2_pack8(o);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = up_lerp2(tx1[e], bx1[e], 0.25f);
        out[0][1] = up2_pack8(o);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = up_lerp2(tx0[e], bx0[e], ly1);
        out[1][0] = up2_pack8(o);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = up_lerp2(tx1[e], bx1[e], ly1);
        out[1][1] = up2_pack8(o);
        const uint4 zero = make_uint4(0, 0, 0, 0);
        if (ky < 0 || kx < 0 || ky == h - 1 || kx == w - 1) {   // the quad straddles the image border: its outer row / column is padding
          if (ky < 0) out[0][0] = out[0][1] = zero;
          if (ky == h - 1) out[1][0] = out[1][1] = zero;
          if (kx < 0) out[0][0] = out[1][0] = zero;
          if (kx == w - 1) out[0][1] = out[1][1] = zero;
        }
      } else {
        out[0][0] = out[0][1] = out[1][0] = out[1][1] = make_uint4(0, 0, 0, 0);
      }
      // rows r = py*PW + px (+1, +PW, +PW+1); px, PW even: r and r + PW share r & 7 parity bits only through PW
      const uint32_t r0 = (uint32_t)(py * PW + px), r1 = r0 + (uint32_t)PW;
      st_shared_u4(tile + r0 * 128u + ((jb ^ (r0 & 7u)) << 4), out[0][0]);
      st_shared_u4(tile + (r0 + 1u) * 128u + ((jb ^ ((r0 + 1u) & 7u)) << 4), out[0][1]);
      st_shared_u4(tile + r1 * 128u + ((jb ^ (r1 & 7u)) << 4), out[1][0]);
      st_shared_u4(tile + (r1 + 1u) * 128u + ((jb ^ ((r1 + 1u) & 7u)) << 4), out[1][1]);
    }
  }
}

}  // namespace vinet
