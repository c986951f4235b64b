// Device post-processing of saliency maps for the sliding-window inference driver (generate_result.py:100-104, utils.py:61-78):
// bilinear resize to the source resolution (cv2.resize INTER_LINEAR, float path), 11x11 Gaussian blur (cv2.GaussianBlur with
// sigma 0 -> sigma = 0.3*((k-1)/2 - 1) + 0.8 = 2, BORDER_REFLECT_101), per-map min-max normalisation
// (torchvision make_grid(normalize=True): (x - min) / (max - min + 1e-5)) and rounding to 8 bits (round(255 x + 0.5)).
// All passes are HBM-bound elementwise / 11-tap kernels on (N, oh, ow) fp32 maps.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace vinet {

constexpr int PP_K = 11, PP_R = 5;
__constant__ float c_gauss[PP_K];

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

__global__ void pp_init_kernel(float* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = __uint_as_float(0x7f800000u);   // +inf
    mm[2 * i + 1] = 0.f;
  }
}

__global__ void __launch_bounds__(256) pp_resize_kernel(const __grid_constant__ vinet_postproc_t d) {
  const double sx = (double)d.W / (double)d.ow, sy = (double)d.H / (double)d.oh;   // cv2 computes source coordinates in double
  const int64_t total = (int64_t)d.N * d.oh * d.ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % d.ow);
    const int y = (int)((i / d.ow) % d.oh);
    const int n = (int)(i / ((int64_t)d.ow * d.oh));
    const double dx = ((double)x + 0.5) * sx - 0.5, dy = ((double)y + 0.5) * sy - 0.5;
    int x0 = (int)floor(dx), y0 = (int)floor(dy);
    float fx = (float)(dx - (double)x0), fy = (float)(dy - (double)y0);
    if (x0 < 0) { x0 = 0; fx = 0.f; }
    if (x0 >= d.W - 1) { x0 = d.W - 1; fx = 0.f; }
    if (y0 < 0) { y0 = 0; fy = 0.f; }
    if (y0 >= d.H - 1) { y0 = d.H - 1; fy = 0.f; }
    const int x1 = min(x0 + 1, d.W - 1), y1 = min(y0 + 1, d.H - 1);
    const float* p = d.x + (int64_t)n * d.H * d.W;
    const float top = p[y0 * d.W + x0] * (1.f - fx) + p[y0 * d.W + x1] * fx;
    const float bot = p[y1 * d.W + x0] * (1.f - fx) + p[y1 * d.W + x1] * fx;
    d.ws0[i] = top * (1.f - fy) + bot * fy;
  }
}

// one 11-tap pass along x (DIR 0) or y (DIR 1); the y pass also folds the per-map min / max
template <int DIR>
__global__ void __launch_bounds__(256) pp_blur_kernel(const float* __restrict__ src, float* __restrict__ dst, float* mm, int N, int oh, int ow,
                                                      int blur) {
  const int64_t total = (int64_t)N * oh * ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % ow);
    const int y = (int)((i / ow) % oh);
    const int n = (int)(i / ((int64_t)ow * oh));
    const float* p = src + (int64_t)n * oh * ow;
    float acc;
    if (blur) {
      acc = 0.f;
#pragma unroll
      for (int k = 0; k < PP_K; ++k) {
        const float v = DIR == 0 ? p[y * ow + reflect101(x + k - PP_R, ow)] : p[reflect101(y + k - PP_R, oh) * ow + x];
        acc = fmaf(c_gauss[k], v, acc);
      }
    } else {
      acc = p[y * ow + x];
    }
    dst[i] = acc;
    if (DIR == 1) {   // maps are positive (sigmoid outputs, positive weights): float order == unsigned order of the bit patterns
      atomicMin(reinterpret_cast<unsigned*>(mm + 2 * n), __float_as_uint(acc));
      atomicMax(reinterpret_cast<unsigned*>(mm + 2 * n + 1), __float_as_uint(acc));
    }
  }
}

__global__ void __launch_bounds__(256) pp_quantise_kernel(const float* __restrict__ src, const float* __restrict__ mm, uint8_t* __restrict__ out,
                                                          int N, int oh, int ow) {
  const int64_t total = (int64_t)N * oh * ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / ((int64_t)ow * oh));
    const float lo = mm[2 * n], hi = mm[2 * n + 1];
    float v = (fminf(fmaxf(src[i], lo), hi) - lo) / (hi - lo + 1e-5f);
    v = fminf(fmaxf(v * 255.f + 0.5f, 0.f), 255.f);
    out[i] = (uint8_t)rintf(v);
  }
}

}  // namespace vinet
using namespace vinet;

extern "C" int vinet_saliency_postprocess(const vinet_postproc_t* d, vinet_stream_t stream) {
  VINET_CHECK(d && d->x && d->ws0 && d->ws1 && d->minmax && d->out, "saliency_postprocess: null pointer");
  VINET_CHECK(d->N >= 1 && d->H >= 1 && d->W >= 1 && d->oh >= 1 && d->ow >= 1, "saliency_postprocess: empty shape");
  static bool table_ready[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (dev < 64 && !table_ready[dev]) {        // cv2.getGaussianKernel(11, sigma = 2): exp(-x^2 / (2 sigma^2)), normalised
    float k[PP_K];
    double sum = 0.0;
    for (int i = 0; i < PP_K; ++i) { const double x = i - PP_R; k[i] = (float)exp(-x * x / 8.0); sum += k[i]; }
    for (int i = 0; i < PP_K; ++i) k[i] = (float)(k[i] / sum);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_gauss, k, sizeof(k), 0, cudaMemcpyHostToDevice, st);
    VINET_CHECK(e == cudaSuccess, "saliency_postprocess: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(st);   // k lives on this stack frame
    table_ready[dev] = true;
  }
  const int64_t total = (int64_t)d->N * d->oh * d->ow;
  const unsigned nb = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256), 148 * 16));
  pp_init_kernel<<<(unsigned)cdiv(d->N, 128), 128, 0, st>>>(d->minmax, d->N);
  VINET_LAUNCH_OK("pp_init");
  pp_resize_kernel<<<nb, 256, 0, st>>>(*d);
  VINET_LAUNCH_OK("pp_resize");
  pp_blur_kernel<0><<<nb, 256, 0, st>>>(d->ws0, d->ws1, d->minmax, d->N, d->oh, d->ow, d->blur);
  VINET_LAUNCH_OK("pp_blur_x");
  pp_blur_kernel<1><<<nb, 256, 0, st>>>(d->ws1, d->ws0, d->minmax, d->N, d->oh, d->ow, d->blur);
  VINET_LAUNCH_OK("pp_blur_y");
  pp_quantise_kernel<<<nb, 256, 0, st>>>(d->ws0, d->minmax, d->out, d->N, d->oh, d->ow);
  VINET_LAUNCH_OK("pp_quantise");
  return 0;
}
