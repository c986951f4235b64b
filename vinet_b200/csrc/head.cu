// Decoder head: relu? -> Conv3d(C,1,kernel 1,bias) -> Sigmoid on an NDHWC view (model.py:280-283).
// Memory bound (AI ~ 1): one thread per pixel, weights in shared memory.
// up2: the head reads a low-res tensor through the decoder's last 2x bilinear up-sampling (up2.cuh) - the hi-res input of the
// head (the largest activation of the decoder) is never written.
#include "up2.cuh"

namespace vinet {

constexpr int HEAD_MAXC = 64;

// 8 channels of the head's input at row r (hi-res pixel when up2), BEFORE the post-interpolation ReLU
template <typename T>
__device__ __forceinline__ void head_load8(const vinet_head_t& d, const T* __restrict__ x, int64_t r, int c, float (&v)[8]) {
  if (d.up2) {
    const int W2 = 2 * d.up_w, H2 = 2 * d.up_h;
    const int X = (int)(r % W2);
    const int64_t q = r / W2;
    const int Y = (int)(q % H2);
    const int64_t n = q / H2;
    up2_load8(x + n * d.up_h * d.up_w * d.ldx + c, d.up_h, d.up_w, d.ldx, Y, X, d.relu_pre != 0, v);
    if constexpr (sizeof(T) == 2) {   // bf16 storage: round where the materialised up-sampled tensor would have been rounded
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
    }
  } else {
    load8(x + r * d.ldx + c, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const __grid_constant__ vinet_head_t d) {
  __shared__ float w[HEAD_MAXC];
  if (threadIdx.x < d.C) w[threadIdx.x] = d.w[threadIdx.x];
  __syncthreads();
  const float bias = d.b ? d.b[0] : 0.f;
  const T* __restrict__ x = reinterpret_cast<const T*>(d.x);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < d.rows; r += (int64_t)gridDim.x * blockDim.x) {
    float acc = bias;
    for (int c = 0; c < d.C; c += 8) {
      float v[8];
      head_load8(d, x, r, c, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a = d.relu ? fmaxf(v[e], 0.f) : v[e];
        acc = fmaf(a, w[c + e], acc);
      }
    }
    d.out[r] = 1.f / (1.f + expf(-acc));
  }
}

// dlogit = gout * o * (1 - o); dx = dlogit * w (ReLU-masked); dw += dlogit * act(x); db += dlogit
template <typename T, typename TD>
__global__ void __launch_bounds__(256) head_bwd_kernel(const __grid_constant__ vinet_head_t d) {
  __shared__ float w[HEAD_MAXC];
  __shared__ float red[8][HEAD_MAXC + 1];
  if (threadIdx.x < d.C) w[threadIdx.x] = d.w[threadIdx.x];
  __syncthreads();
  const T* __restrict__ x = reinterpret_cast<const T*>(d.x);
  TD* __restrict__ dx = reinterpret_cast<TD*>(d.dx);
  float dwacc[HEAD_MAXC];
#pragma unroll
  for (int c = 0; c < HEAD_MAXC; ++c) dwacc[c] = 0.f;
  float dbacc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < d.rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float o = d.out[r];
    const float dl = d.gout[r] * o * (1.f - o);
    dbacc += dl;
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; c += 8) {
      if (c < d.C) {
        float v[8], g[8];
        head_load8(d, x, r, c, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool on = !d.relu || v[e] > 0.f;
          const float a = on ? v[e] : 0.f;
          dwacc[c + e] = fmaf(dl, a, dwacc[c + e]);
          g[e] = on ? dl * w[c + e] : 0.f;
        }
        store8(dx + r * d.lddx + c, g);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < HEAD_MAXC; ++c) {
    const float s = warp_sum(dwacc[c]);
    if (lane == 0) red[warp][c] = s;
  }
  {
    const float s = warp_sum(dbacc);
    if (lane == 0) red[warp][HEAD_MAXC] = s;
  }
  __syncthreads();
  if (threadIdx.x <= HEAD_MAXC) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][c];
    if (c < d.C) atomicAdd(d.dw + c, s);
    else if (c == HEAD_MAXC) atomicAdd(d.db, s);
  }
}

}  // namespace vinet
using namespace vinet;

static bool head_up2_ok(const vinet_head_t* d) {
  return !d->up2 || (d->up_h >= 1 && d->up_w >= 1 && d->rows % (4ll * d->up_h * d->up_w) == 0);
}

extern "C" int vinet_head_fwd(const vinet_head_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= HEAD_MAXC, "head: C %d", d->C);
  VINET_CHECK(head_up2_ok(d), "head: up2 needs rows %lld = frames * 4 * up_h %d * up_w %d", (long long)d->rows, d->up_h, d->up_w);
  int64_t nb = cdiv(d->rows, 256);
  if (nb > 148 * 16) nb = 148 * 16;
  VINET_DISPATCH_DTYPE(d->dtype, T, (head_fwd_kernel<T><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(*d)));
  VINET_LAUNCH_OK("head_fwd");
  return 0;
}

extern "C" int vinet_head_bwd(const vinet_head_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= HEAD_MAXC, "head: C %d", d->C);
  VINET_CHECK(head_up2_ok(d), "head: up2 needs rows %lld = frames * 4 * up_h %d * up_w %d", (long long)d->rows, d->up_h, d->up_w);
  int64_t nb = cdiv(d->rows, 256 * 8);
  if (nb > 148 * 4) nb = 148 * 4;
  if (nb < 1) nb = 1;
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->dx_dtype, TD,
      (head_bwd_kernel<T, TD><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(*d))));
  VINET_LAUNCH_OK("head_bwd");
  return 0;
}
