// Input pipeline on the device: the reference's per-frame transform (dataloader.py:242-249,
// generate_result.py:77-89) - transforms.Resize((224,384)) on a PIL image, ToTensor, ImageNet Normalize - for a batch of
// decoded uint8 RGB frames, and the audio excerpt windowing (dataloader.py:113-118).
//
// The resize reproduces Pillow's bilinear resampling bit for bit (src/libImaging/Resample.c, 8bpc path): an antialiased
// separable convolution in 22-bit fixed point, horizontal pass first, 8-bit intermediate image.  The host computes the
// per-output bounds / coefficient tables exactly as precompute_coeffs + normalize_coeffs_8bpc do (vinet_b200/preprocess.py).
// Both passes are HBM-bound byte kernels: pass 1 reads the source frames once and writes the [N,h,W,3] intermediate, pass 2
// reads it once and writes the fp32 (N,3,H,W) planes the model's packing kernel consumes.
#include <algorithm>

#include "common.cuh"

namespace vinet {

constexpr int PRE_BITS = 22;

__device__ __forceinline__ int clip8(int v) { return min(max(v >> PRE_BITS, 0), 255); }

// pass 1: one thread per (n, y, X): 3 channels of one intermediate pixel
__global__ void __launch_bounds__(256) preproc_h_kernel(const __grid_constant__ vinet_preproc_t d) {
  const int64_t total = (int64_t)d.N * d.h * d.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % d.W);
    const int64_t row = i / d.W;   // n*h + y
    const int xmin = __ldg(d.xb + 2 * X), xn = __ldg(d.xb + 2 * X + 1);
    const int32_t* __restrict__ k = d.xk + (int64_t)X * d.xks;
    const uint8_t* __restrict__ p = d.frames + (row * d.w + xmin) * 3;
    int s0 = 1 << (PRE_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xn; ++x) {
      const int kk = __ldg(k + x);
      s0 += (int)p[3 * x] * kk;
      s1 += (int)p[3 * x + 1] * kk;
      s2 += (int)p[3 * x + 2] * kk;
    }
    uint8_t* o = d.tmp + i * 3;
    o[0] = (uint8_t)clip8(s0);
    o[1] = (uint8_t)clip8(s1);
    o[2] = (uint8_t)clip8(s2);
  }
}

// pass 2: one thread per (n, Y, X): vertical convolution, then ToTensor (/255) and Normalize ((v - mean) / std) per channel
__global__ void __launch_bounds__(256) preproc_v_kernel(const __grid_constant__ vinet_preproc_t d) {
  const int64_t total = (int64_t)d.N * d.H * d.W;
  const int64_t plane = (int64_t)d.H * d.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % d.W);
    const int Y = (int)((i / d.W) % d.H);
    const int n = (int)(i / plane);
    const int ymin = __ldg(d.yb + 2 * Y), yn = __ldg(d.yb + 2 * Y + 1);
    const int32_t* __restrict__ k = d.yk + (int64_t)Y * d.yks;
    const uint8_t* __restrict__ p = d.tmp + (((int64_t)n * d.h + ymin) * d.W + X) * 3;
    int s0 = 1 << (PRE_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < yn; ++y) {
      const int kk = __ldg(k + y);
      const uint8_t* q = p + (int64_t)y * d.W * 3;
      s0 += (int)q[0] * kk;
      s1 += (int)q[1] * kk;
      s2 += (int)q[2] * kk;
    }
    const int v[3] = {clip8(s0), clip8(s1), clip8(s2)};
    float* o = d.out + (int64_t)n * 3 * plane + (int64_t)Y * d.W + X;
#pragma unroll
    for (int c = 0; c < 3; ++c)   // the exact float operations of ToTensor + Normalize: no contraction, IEEE division
      o[c * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[c], 255.f), d.mean[c]), d.std[c]);
  }
}

// out[total]: zeros with excerpt[n] * hanning(n) centred (np.hanning: 0.5 - 0.5 cos(2 pi i / (n - 1)), computed in double)
__global__ void audio_window_kernel(const float* __restrict__ ex, int n, float* __restrict__ out, int total, int B) {
  const int lo = total / 2 - n / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)B * total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % total) - lo;
    const int b = (int)(i / total);
    float v = 0.f;
    if (j >= 0 && j < n) {
      const double w = n > 1 ? 0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * (double)j / (double)(n - 1)) : 1.0;
      v = (float)w * ex[(int64_t)b * n + j];
    }
    out[i] = v;
  }
}

}  // namespace vinet
using namespace vinet;

extern "C" int vinet_preprocess_frames(const vinet_preproc_t* d, vinet_stream_t stream) {
  VINET_CHECK(d && d->frames && d->tmp && d->out && d->xb && d->xk && d->yb && d->yk, "preprocess_frames: null pointer");
  VINET_CHECK(d->N >= 1 && d->h >= 1 && d->w >= 1 && d->H >= 1 && d->W >= 1 && d->xks >= 1 && d->yks >= 1, "preprocess_frames: empty shape");
  const int64_t t1 = (int64_t)d->N * d->h * d->W, t2 = (int64_t)d->N * d->H * d->W;
  preproc_h_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(t1, 256), 148 * 32)), 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("preproc_h");
  preproc_v_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(t2, 256), 148 * 32)), 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("preproc_v");
  return 0;
}

extern "C" int vinet_audio_window(const float* excerpt, int32_t B, int32_t n, float* out, int32_t total, vinet_stream_t stream) {
  VINET_CHECK(excerpt && out && B >= 1 && n >= 0 && n <= total, "audio_window: n %d total %d", n, total);
  const int64_t t = (int64_t)B * total;
  audio_window_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(t, 256), 148 * 8)), 256, 0, (cudaStream_t)stream>>>(excerpt, n, out, total, B);
  VINET_LAUNCH_OK("audio_window");
  return 0;
}
