// relu? + nn.Upsample(scale_factor=(1,2,2), mode='trilinear', align_corners=False) (model.py:254):
// per-frame 2x bilinear with taps {.25,.75} and index clamp at the borders.
#include "common.cuh"

namespace vinet {

// source taps of output coordinate Y on an axis of length n: indices i0,i1 and the weight of i1
__device__ __forceinline__ void up_taps(int Y, int n, int& i0, int& i1, float& l1) {
  float src = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f);
  i0 = (int)src;
  i1 = min(i0 + 1, n - 1);
  l1 = src - (float)i0;
}

template <typename T, typename TO>
__global__ void upsample_fwd_kernel(const __grid_constant__ vinet_upsample_t d) {
  const T* __restrict__ z = reinterpret_cast<const T*>(d.z);
  TO* __restrict__ u = reinterpret_cast<TO*>(d.u);
  const int G = d.C / 8, H2 = 2 * d.h, W2 = 2 * d.w;
  const int64_t total = (int64_t)d.B * d.T * H2 * W2 * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c = (int)(r % G) * 8; r /= G;
    const int X = (int)(r % W2); r /= W2;
    const int Y = (int)(r % H2); r /= H2;  // r = b*T + t
    int y0, y1, x0, x1;
    float ly, lx;
    up_taps(Y, d.h, y0, y1, ly);
    up_taps(X, d.w, x0, x1, lx);
    const int64_t base = r * d.h;
    float a[8], b[8], e0[8], e1[8], o[8];
    load8(z + ((base + y0) * d.w + x0) * d.ldz + c, a);
    load8(z + ((base + y0) * d.w + x1) * d.ldz + c, b);
    load8(z + ((base + y1) * d.w + x0) * d.ldz + c, e0);
    load8(z + ((base + y1) * d.w + x1) * d.ldz + c, e1);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (d.relu) { a[e] = fmaxf(a[e], 0.f); b[e] = fmaxf(b[e], 0.f); e0[e] = fmaxf(e0[e], 0.f); e1[e] = fmaxf(e1[e], 0.f); }
      const float top = (1.f - lx) * a[e] + lx * b[e];
      const float bot = (1.f - lx) * e0[e] + lx * e1[e];
      o[e] = (1.f - ly) * top + ly * bot;
    }
    store8(u + ((r * H2 + Y) * W2 + X) * d.ldu + c, o);
  }
}

// weight with which low-res index i contributes to output coordinate Y
__device__ __forceinline__ float up_weight(int Y, int n, int i) {
  int i0, i1;
  float l1;
  up_taps(Y, n, i0, i1, l1);
  float w = 0.f;
  if (i0 == i) w += 1.f - l1;
  if (i1 == i) w += l1;
  return w;
}

template <typename T, typename TD, typename TG>
__global__ void upsample_bwd_kernel(const __grid_constant__ vinet_upsample_t d) {
  const T* __restrict__ z = reinterpret_cast<const T*>(d.z);
  const TG* __restrict__ gu = reinterpret_cast<const TG*>(d.gu);
  TD* __restrict__ dz = reinterpret_cast<TD*>(d.dz);
  const int G = d.C / 8, H2 = 2 * d.h, W2 = 2 * d.w;
  const int64_t total = (int64_t)d.B * d.T * d.h * d.w * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c = (int)(r % G) * 8; r /= G;
    const int x = (int)(r % d.w); r /= d.w;
    const int y = (int)(r % d.h); r /= d.h;  // r = b*T + t
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int Y = max(0, 2 * y - 2); Y <= min(H2 - 1, 2 * y + 3); ++Y) {
      const float wy = up_weight(Y, d.h, y);
      if (wy == 0.f) continue;
      for (int X = max(0, 2 * x - 2); X <= min(W2 - 1, 2 * x + 3); ++X) {
        const float wx = up_weight(X, d.w, x);
        if (wx == 0.f) continue;
        float g[8];
        load8(gu + ((r * H2 + Y) * W2 + X) * d.ldgu + c, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wy * wx, g[e], acc[e]);
      }
    }
    if (d.relu) {
      float v[8];
      load8(z + ((r * d.h + y) * d.w + x) * d.ldz + c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!(v[e] > 0.f)) acc[e] = 0.f;
    }
    store8(dz + ((r * d.h + y) * d.w + x) * d.lddz + c, acc);
  }
}

}  // namespace vinet
using namespace vinet;

static unsigned up_grid(int64_t total) {
  int64_t nb = cdiv(total, 256);
  if (nb > 148 * 32) nb = 148 * 32;
  return (unsigned)(nb < 1 ? 1 : nb);
}

extern "C" int vinet_upsample_fwd(const vinet_upsample_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0, "upsample: C %d", d->C);
  const int64_t total = (int64_t)d->B * d->T * 4 * d->h * d->w * (d->C / 8);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->u_dtype, TO,
      (upsample_fwd_kernel<T, TO><<<up_grid(total), 256, 0, (cudaStream_t)stream>>>(*d))));
  VINET_LAUNCH_OK("upsample_fwd");
  return 0;
}

extern "C" int vinet_upsample_bwd(const vinet_upsample_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0, "upsample: C %d", d->C);
  const int64_t total = (int64_t)d->B * d->T * d->h * d->w * (d->C / 8);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->dz_dtype, TD, VINET_DISPATCH_DTYPE(d->gu_dtype, TG,
      (upsample_bwd_kernel<T, TD, TG><<<up_grid(total), 256, 0, (cudaStream_t)stream>>>(*d)))));
  VINET_LAUNCH_OK("upsample_bwd");
  return 0;
}
