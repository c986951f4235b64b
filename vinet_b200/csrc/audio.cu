// AViNet audio branch: SoundNet 1-D convolutions (model.py:750-786, nn.Conv2d with (k,1) kernels on a
// (B,1,L,1) waveform), BatchNorm2d+ReLU+MaxPool, and the audio-visual bilinear fusion (model.py:229-237).
// 0.19 GFLOP per clip: warp-per-output kernels with shuffle reductions, fp32 throughout.
#include "common.cuh"

namespace vinet {

// ------------------------------------------------------------------ conv1d
// warp per output (b,co,l); lanes stride over the (ci,k) reduction (contiguous in both x and w)
__global__ void conv1d_fwd_kernel(const __grid_constant__ vinet_conv1d_t d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t total = (int64_t)d.B * d.Cout * d.Lout;
  if (warp >= total) return;
  const int l = (int)(warp % d.Lout);
  const int co = (int)((warp / d.Lout) % d.Cout);
  const int b = (int)(warp / ((int64_t)d.Lout * d.Cout));
  const int R = d.Cin * d.k;
  const float* __restrict__ w = d.w + (int64_t)co * R;
  const float* __restrict__ x = d.x + (int64_t)b * d.Cin * d.Lin;
  const int i0 = l * d.stride - d.pad;
  float acc = 0.f;
  for (int r = lane; r < R; r += 32) {
    const int ci = r / d.k, kk = r - ci * d.k;
    const int i = i0 + kk;
    if ((unsigned)i < (unsigned)d.Lin) acc = fmaf(w[r], x[(int64_t)ci * d.Lin + i], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) d.y[warp] = acc + (d.bias ? d.bias[co] : 0.f);
}

// warp per input element (b,ci,i); lanes stride over (co,k)
__global__ void conv1d_dgrad_kernel(const __grid_constant__ vinet_conv1d_t d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t total = (int64_t)d.B * d.Cin * d.Lin;
  if (warp >= total) return;
  const int i = (int)(warp % d.Lin);
  const int ci = (int)((warp / d.Lin) % d.Cin);
  const int b = (int)(warp / ((int64_t)d.Lin * d.Cin));
  const int R = d.Cout * d.k;
  float acc = 0.f;
  for (int r = lane; r < R; r += 32) {
    const int co = r / d.k, kk = r - co * d.k;
    const int num = i + d.pad - kk;
    if (num < 0 || num % d.stride) continue;
    const int l = num / d.stride;
    if (l >= d.Lout) continue;
    acc = fmaf(d.w[((int64_t)co * d.Cin + ci) * d.k + kk], d.dy[((int64_t)b * d.Cout + co) * d.Lout + l], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) d.dx[warp] = acc;
}

// warp per weight (co,ci,k) (+ one warp per bias); lanes stride over (b,l)
__global__ void conv1d_wgrad_kernel(const __grid_constant__ vinet_conv1d_t d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t)d.Cout * d.Cin * d.k;
  if (warp >= nw + d.Cout) return;
  const int64_t BL = (int64_t)d.B * d.Lout;
  float acc = 0.f;
  if (warp < nw) {
    const int kk = (int)(warp % d.k);
    const int ci = (int)((warp / d.k) % d.Cin);
    const int co = (int)(warp / ((int64_t)d.k * d.Cin));
    for (int64_t r = lane; r < BL; r += 32) {
      const int b = (int)(r / d.Lout), l = (int)(r - (int64_t)b * d.Lout);
      const int i = l * d.stride - d.pad + kk;
      if ((unsigned)i < (unsigned)d.Lin)
        acc = fmaf(d.dy[((int64_t)b * d.Cout + co) * d.Lout + l], d.x[((int64_t)b * d.Cin + ci) * d.Lin + i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) d.dw[warp] = acc;
  } else {
    const int co = (int)(warp - nw);
    for (int64_t r = lane; r < BL; r += 32) {
      const int b = (int)(r / d.Lout), l = (int)(r - (int64_t)b * d.Lout);
      acc += d.dy[((int64_t)b * d.Cout + co) * d.Lout + l];
    }
    acc = warp_sum(acc);
    if (lane == 0 && d.dbias) d.dbias[co] = acc;
  }
}

// ------------------------------------------------------------------ BatchNorm2d + ReLU + MaxPool((p,1))
__device__ __forceinline__ double block_sum256(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < 8; ++i) s += sh[i];
  return s;
}

// block per channel
__global__ void __launch_bounds__(256) bn1d_fwd_kernel(const __grid_constant__ vinet_bn1d_t d) {
  __shared__ double sh[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int64_t N = (int64_t)d.B * d.L;
  float mean, invstd;
  if (d.training) {
    double s = 0.0, ss = 0.0;
    for (int64_t r = tid; r < N; r += 256) {
      const int b = (int)(r / d.L), l = (int)(r - (int64_t)b * d.L);
      const double v = d.y[((int64_t)b * d.C + c) * d.L + l];
      s += v; ss += v * v;
    }
    s = block_sum256(s, sh); ss = block_sum256(ss, sh);
    const double m = s / (double)N;
    double var = ss / (double)N - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)d.eps));
    if (tid == 0 && d.running_mean) {
      const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
      d.running_mean[c] = (float)((1.0 - d.momentum) * d.running_mean[c] + d.momentum * m);
      d.running_var[c] = (float)((1.0 - d.momentum) * d.running_var[c] + d.momentum * unb);
    }
  } else {
    mean = d.running_mean[c];
    invstd = 1.0f / sqrtf(d.running_var[c] + d.eps);
  }
  if (tid == 0) { d.mean[c] = mean; d.invstd[c] = invstd; }
  const float sc = d.gamma[c] * invstd, shf = d.beta[c] - mean * sc;
  const int Lo = d.L / d.pool;
  for (int64_t r = tid; r < (int64_t)d.B * Lo; r += 256) {
    const int b = (int)(r / Lo), lo = (int)(r - (int64_t)b * Lo);
    const float* y = d.y + ((int64_t)b * d.C + c) * d.L + (int64_t)lo * d.pool;
    float best = -INFINITY;
    for (int p = 0; p < d.pool; ++p) best = fmaxf(best, fmaxf(fmaf(y[p], sc, shf), 0.f));
    d.out[((int64_t)b * d.C + c) * Lo + lo] = best;
  }
}

__global__ void __launch_bounds__(256) bn1d_bwd_kernel(const __grid_constant__ vinet_bn1d_t d) {
  __shared__ double sh[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int64_t N = (int64_t)d.B * d.L;
  const float mean = d.mean[c], invstd = d.invstd[c];
  const float sc = d.gamma[c] * invstd, shf = d.beta[c] - mean * sc;
  const int Lo = d.L / d.pool;
  // gradient w.r.t. the activated, pre-pool value of element (b,l): routed to the first window maximum
  auto g_act = [&](int b, int l) -> float {
    const int lo = l / d.pool;
    if (lo >= Lo) return 0.f;
    const float* y = d.y + ((int64_t)b * d.C + c) * d.L + (int64_t)lo * d.pool;
    float best = -INFINITY;
    int arg = 0;
    for (int p = 0; p < d.pool; ++p) {
      const float v = fmaxf(fmaf(y[p], sc, shf), 0.f);
      if (v > best) { best = v; arg = p; }
    }
    if (lo * d.pool + arg != l) return 0.f;
    if (!(fmaf(y[arg], sc, shf) > 0.f)) return 0.f;
    return d.gout[((int64_t)b * d.C + c) * Lo + lo];
  };
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = tid; r < N; r += 256) {
    const int b = (int)(r / d.L), l = (int)(r - (int64_t)b * d.L);
    const float g = g_act(b, l);
    const float yn = (d.y[((int64_t)b * d.C + c) * d.L + l] - mean) * invstd;
    s1 += g; s2 += (double)g * yn;
  }
  s1 = block_sum256(s1, sh); s2 = block_sum256(s2, sh);
  if (tid == 0) { d.dbeta[c] = (float)s1; d.dgamma[c] = (float)s2; }
  const float k1 = d.training ? (float)(s1 / (double)N) : 0.f, k2 = d.training ? (float)(s2 / (double)N) : 0.f;
  for (int64_t r = tid; r < N; r += 256) {
    const int b = (int)(r / d.L), l = (int)(r - (int64_t)b * d.L);
    const float g = g_act(b, l);
    const float yn = (d.y[((int64_t)b * d.C + c) * d.L + l] - mean) * invstd;
    d.dy[((int64_t)b * d.C + c) * d.L + l] = sc * (g - k1 - yn * k2);
  }
}

// ------------------------------------------------------------------ audio-visual bilinear fusion
constexpr int AV_I = 42, AV_J = 3, AV_O = 336, AV_IJ = AV_I * AV_J;

// v[b,c,i] = max over the 4 frames of act(y0[b,t,h,2j,c]), i = h*6 + j   (MaxPool3d((4,1,1),stride=(2,1,2)))
template <typename T>
__global__ void avfuse_pool_kernel(const __grid_constant__ vinet_avfuse_t d) {
  const T* __restrict__ y0 = reinterpret_cast<const T*>(d.y0);
  const int64_t total = (int64_t)d.B * AV_I * d.C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % d.C);
    const int i = (int)((idx / d.C) % AV_I);
    const int b = (int)(idx / ((int64_t)d.C * AV_I));
    const int h = i / 6, w = (i % 6) * 2;
    float best = -INFINITY;
    for (int t = 0; t < 4; ++t) {
      float v = load1(y0 + ((((int64_t)b * 4 + t) * 7 + h) * 12 + w) * d.ld + c);
      if (d.xform & 2) v = fmaf(v, d.scale[c], d.shift[c]);
      if (d.xform & 1) v = fmaxf(v, 0.f);
      best = fmaxf(best, v);
    }
    d.vbuf[((int64_t)b * d.C + c) * AV_I + i] = best;
  }
}

// out[b,o,c] = sum_ij v[b,c,i] a[b,c,j] W[o,i,j] + bias[o];  block = 32 channels x 8 outputs
template <typename TO>
__global__ void __launch_bounds__(256) avfuse_fwd_kernel(const __grid_constant__ vinet_avfuse_t d) {
  __shared__ float P[32][AV_IJ + 1];
  const int cl = threadIdx.x & 31, ol = threadIdx.x >> 5;
  const int nct = d.C / 32;
  const int b = blockIdx.x / nct, c0 = (blockIdx.x % nct) * 32;
  for (int e = threadIdx.x; e < 32 * AV_IJ; e += 256) {
    const int ch = e / AV_IJ, ij = e - ch * AV_IJ;
    const int64_t bc = (int64_t)b * d.C + c0 + ch;
    P[ch][ij] = d.vbuf[bc * AV_I + ij / AV_J] * d.audio[bc * AV_J + ij % AV_J];
  }
  __syncthreads();
  TO* __restrict__ out = reinterpret_cast<TO*>(d.out);
  for (int o = blockIdx.y * 8 + ol; o < AV_O; o += gridDim.y * 8) {
    const float* __restrict__ w = d.w + (int64_t)o * AV_IJ;
    float acc = d.bias[o];
#pragma unroll 6
    for (int ij = 0; ij < AV_IJ; ++ij) acc = fmaf(P[cl][ij], __ldg(w + ij), acc);
    store1(out + ((int64_t)b * AV_O + o) * d.ldo + c0 + cl, acc);
  }
}

// dp[b,c,ij] = sum_o gout[b,o,c] W[o,ij]; dv[i] = sum_j dp[ij] a[j] -> scattered to the arg-max frame; da[j] = sum_i dp[ij] v[i]
template <typename T>
__global__ void __launch_bounds__(256) avfuse_bwd_data_kernel(const __grid_constant__ vinet_avfuse_t d) {
  __shared__ float DP[32][AV_IJ + 1];
  const int nct = d.C / 32;
  const int b = blockIdx.x / nct, c0 = (blockIdx.x % nct) * 32;
  for (int e = threadIdx.x; e < 32 * AV_IJ; e += 256) {
    const int ch = e & 31, ij = e >> 5;
    float acc = 0.f;
    for (int o = 0; o < AV_O; ++o)
      acc = fmaf(__ldg(d.gout + ((int64_t)b * AV_O + o) * d.ldgo + c0 + ch), __ldg(d.w + (int64_t)o * AV_IJ + ij), acc);
    DP[ch][ij] = acc;
  }
  __syncthreads();
  const T* __restrict__ y0 = reinterpret_cast<const T*>(d.y0);
  for (int e = threadIdx.x; e < 32 * (AV_I + AV_J); e += 256) {
    const int ch = e & 31, q = e >> 5;
    const int c = c0 + ch;
    const int64_t bc = (int64_t)b * d.C + c;
    if (q < AV_I) {
      float dv = 0.f;
      for (int j = 0; j < AV_J; ++j) dv = fmaf(DP[ch][q * AV_J + j], d.audio[bc * AV_J + j], dv);
      const int h = q / 6, w = (q % 6) * 2;
      float best = -INFINITY;
      int arg = 0;
      for (int t = 0; t < 4; ++t) {
        float v = load1(y0 + ((((int64_t)b * 4 + t) * 7 + h) * 12 + w) * d.ld + c);
        if (d.xform & 2) v = fmaf(v, d.scale[c], d.shift[c]);
        if (d.xform & 1) v = fmaxf(v, 0.f);
        if (v > best) { best = v; arg = t; }
      }
      atomicAdd(d.gy0 + ((((int64_t)b * 4 + arg) * 7 + h) * 12 + w) * d.ldgy0 + c, dv);
    } else {
      const int j = q - AV_I;
      float da = 0.f;
      for (int i = 0; i < AV_I; ++i) da = fmaf(DP[ch][i * AV_J + j], d.vbuf[bc * AV_I + i], da);
      d.gaudio[bc * AV_J + j] = da;
    }
  }
}

// dW[o,ij] = sum_{b,c} gout[b,o,c] v[b,c,i] a[b,c,j]; dbias[o] = sum_{b,c} gout: warp per (o,ij) / per o
__global__ void avfuse_bwd_weight_kernel(const __grid_constant__ vinet_avfuse_t d) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t)AV_O * AV_IJ;
  if (warp >= nw + AV_O) return;
  const int64_t BC = (int64_t)d.B * d.C;
  float acc = 0.f;
  if (warp < nw) {
    const int o = (int)(warp / AV_IJ), ij = (int)(warp % AV_IJ);
    const int i = ij / AV_J, j = ij % AV_J;
    for (int64_t r = lane; r < BC; r += 32) {
      const int b = (int)(r / d.C), c = (int)(r - (int64_t)b * d.C);
      acc = fmaf(d.gout[((int64_t)b * AV_O + o) * d.ldgo + c], d.vbuf[r * AV_I + i] * d.audio[r * AV_J + j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) d.dw[warp] = acc;
  } else {
    const int o = (int)(warp - nw);
    for (int64_t r = lane; r < BC; r += 32) {
      const int b = (int)(r / d.C), c = (int)(r - (int64_t)b * d.C);
      acc += d.gout[((int64_t)b * AV_O + o) * d.ldgo + c];
    }
    acc = warp_sum(acc);
    if (lane == 0) d.dbias[o] = acc;
  }
}

static unsigned warps_grid(int64_t warps) { return (unsigned)cdiv(warps * 32, 256); }

}  // namespace vinet
using namespace vinet;

extern "C" int vinet_conv1d_fwd(const vinet_conv1d_t* d, vinet_stream_t stream) {
  conv1d_fwd_kernel<<<warps_grid((int64_t)d->B * d->Cout * d->Lout), 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("conv1d_fwd");
  return 0;
}

extern "C" int vinet_conv1d_bwd(const vinet_conv1d_t* d, vinet_stream_t stream) {
  if (d->dx) {
    conv1d_dgrad_kernel<<<warps_grid((int64_t)d->B * d->Cin * d->Lin), 256, 0, (cudaStream_t)stream>>>(*d);
    VINET_LAUNCH_OK("conv1d_dgrad");
  }
  conv1d_wgrad_kernel<<<warps_grid((int64_t)d->Cout * d->Cin * d->k + d->Cout), 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("conv1d_wgrad");
  return 0;
}

extern "C" int vinet_bn1d_fwd(const vinet_bn1d_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->pool >= 1, "bn1d: pool");
  bn1d_fwd_kernel<<<d->C, 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("bn1d_fwd");
  return 0;
}

extern "C" int vinet_bn1d_bwd(const vinet_bn1d_t* d, vinet_stream_t stream) {
  bn1d_bwd_kernel<<<d->C, 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("bn1d_bwd");
  return 0;
}

extern "C" int vinet_avfuse_fwd(const vinet_avfuse_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 32 == 0, "avfuse: C %d", d->C);
  const int64_t total = (int64_t)d->B * AV_I * d->C;
  VINET_DISPATCH_DTYPE(d->dtype, T, (avfuse_pool_kernel<T><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(*d)));
  VINET_LAUNCH_OK("avfuse_pool");
  dim3 grid((unsigned)(d->B * (d->C / 32)), 6);
  VINET_DISPATCH_DTYPE(d->out_dtype, TO, (avfuse_fwd_kernel<TO><<<grid, 256, 0, (cudaStream_t)stream>>>(*d)));
  VINET_LAUNCH_OK("avfuse_fwd");
  return 0;
}

extern "C" int vinet_avfuse_bwd(const vinet_avfuse_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 32 == 0, "avfuse: C %d", d->C);
  VINET_DISPATCH_DTYPE(d->dtype, T,
                       (avfuse_bwd_data_kernel<T><<<(unsigned)(d->B * (d->C / 32)), 256, 0, (cudaStream_t)stream>>>(*d)));
  VINET_LAUNCH_OK("avfuse_bwd_data");
  avfuse_bwd_weight_kernel<<<warps_grid((int64_t)AV_O * AV_IJ + AV_O), 256, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("avfuse_bwd_weight");
  return 0;
}
