// relu? + nn.Upsample(scale_factor=(1,2,2), mode='trilinear', align_corners=False) (model.py:254):
// per-frame 2x bilinear with taps {.25,.75} and index clamp at the borders.
// The forward arithmetic lives in up2.cuh, shared with the consumers that read a low-res tensor THROUGH the up-sampling
// (VINET_XF_UP2); this standalone kernel is the fallback for consumers without a fused input stage.
#include "up2.cuh"

namespace vinet {

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
    float o[8];
    up2_load8(z + r * d.h * d.w * d.ldz + c, d.h, d.w, d.ldz, Y, X, d.relu != 0, o);
    store8(u + ((r * H2 + Y) * W2 + X) * d.ldu + c, o);
  }
}

// Transposed interpolation: low-res index i receives from the hi-res coordinates 2i-1, 2i, 2i+1, 2i+2 with weights .25, .75, .75,
// .25 (up_taps: Y = 2k+1 -> (k: .75, k+1: .25), Y = 2k+2 -> (k: .25, k+1: .75)); at the borders the clamped tap folds its weight
// onto the edge pixel (2i -> 1 for i = 0, 2i+1 -> 1 for i = n-1) and the coordinates outside the image do not exist.
__device__ __forceinline__ void up_bwd_taps(int i, int n, float (&w)[4]) {
  w[0] = i > 0 ? 0.25f : 0.f;
  w[1] = i > 0 ? 0.75f : 1.f;
  w[2] = i < n - 1 ? 0.75f : 1.f;
  w[3] = i < n - 1 ? 0.25f : 0.f;
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
    float wy[4], wx[4];
    up_bwd_taps(y, d.h, wy);
    up_bwd_taps(x, d.w, wx);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const TG* frame = gu + r * H2 * W2 * d.ldgu + c;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.f) continue;
      const TG* row = frame + (int64_t)(2 * y - 1 + a) * W2 * d.ldgu;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (wx[b] == 0.f) continue;
        float g[8];
        load8(row + (int64_t)(2 * x - 1 + b) * d.ldgu, g);
        const float wgt = wy[a] * wx[b];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, g[e], acc[e]);
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

// dz = g where z > 0 else 0
template <typename TG, typename TZ, typename TD>
__global__ void relu_bwd_kernel(const TG* __restrict__ g, int64_t ldg, const TZ* __restrict__ z, int64_t ldz, int64_t rows, int C,
                                TD* __restrict__ dz, int64_t lddz) {
  const int G = C / 8;
  const int64_t total = rows * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / G;
    const int c = (int)(i - r * G) * 8;
    float gv[8], zv[8];
    load8(g + r * ldg + c, gv);
    load8(z + r * ldz + c, zv);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (!(zv[e] > 0.f)) gv[e] = 0.f;
    store8(dz + r * lddz + c, gv);
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

extern "C" int vinet_relu_bwd(const void* g, int64_t ldg, int32_t g_dtype, const void* z, int64_t ldz, int32_t z_dtype, int64_t rows,
                              int32_t C, void* dz, int64_t lddz, int32_t dz_dtype, vinet_stream_t stream) {
  VINET_CHECK(C % 8 == 0 && rows >= 1 && ldg % 8 == 0 && ldz % 8 == 0 && lddz % 8 == 0, "relu_bwd: C %d rows %lld", C, (long long)rows);
  const int64_t total = rows * (C / 8);
  VINET_DISPATCH_DTYPE(g_dtype, TG, VINET_DISPATCH_DTYPE(z_dtype, TZ, VINET_DISPATCH_DTYPE(dz_dtype, TD,
      (relu_bwd_kernel<TG, TZ, TD><<<up_grid(total), 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const TG*>(g), ldg, reinterpret_cast<const TZ*>(z), ldz, rows, C, reinterpret_cast<TD*>(dz), lddz)))));
  VINET_LAUNCH_OK("relu_bwd");
  return 0;
}
