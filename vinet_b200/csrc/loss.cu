// Saliency losses and their input gradients (loss.py:13-120): one CTA per sample, warp-shuffle +
// shared-memory reductions in a fixed order (deterministic), mean over the batch by the last CTA.
//   kldiv      loss.py:13-38     cc   loss.py:80-99
//   similarity loss.py:53-78     nss  loss.py:101-120 (same-size branch)
#include "common.cuh"

namespace vinet {

constexpr int LOSS_THREADS = 512;
constexpr float LOSS_EPS = 2.2204e-16f;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < LOSS_THREADS / 32; ++i) s += sh[i];
  return s;
}

// minimum and the first index attaining it
__device__ __forceinline__ void block_argmin(float& v, int& idx, float* shv, int* shi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if (lane == 0) { shv[warp] = v; shi[warp] = idx; }
  __syncthreads();
  v = shv[0]; idx = shi[0];
  for (int i = 1; i < LOSS_THREADS / 32; ++i)
    if (shv[i] < v || (shv[i] == v && shi[i] < idx)) { v = shv[i]; idx = shi[i]; }
}

__device__ __forceinline__ float block_max(float v, float* shv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) shv[warp] = v;
  __syncthreads();
  v = shv[0];
  for (int i = 1; i < LOSS_THREADS / 32; ++i) v = fmaxf(v, shv[i]);
  return v;
}

// per_sample[b][0] = value; [1..7] = reductions reused by the backward
__global__ void __launch_bounds__(LOSS_THREADS) loss_fwd_kernel(const __grid_constant__ vinet_loss_t d) {
  __shared__ double sh[LOSS_THREADS / 32];
  __shared__ float shv[LOSS_THREADS / 32];
  __shared__ int shi[LOSS_THREADS / 32];
  __shared__ bool is_last;
  const int b = blockIdx.x, n = d.n, tid = threadIdx.x;
  const float* __restrict__ s = d.s + (int64_t)b * n;
  const float* __restrict__ g = d.g + (int64_t)b * n;
  float* ps = d.per_sample + b * 8;
  float value = 0.f;
  if (d.kind == VINET_LOSS_KLDIV) {
    float ls = 0.f, lg = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) { ls += s[i]; lg += g[i]; }
    const float S = (float)block_sum((double)ls, sh);
    const float G = (float)block_sum((double)lg, sh);
    float acc = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float sn = s[i] / S, gn = g[i] / G;
      acc += gn * logf(LOSS_EPS + gn / (sn + LOSS_EPS));
    }
    value = (float)block_sum((double)acc, sh);
    if (tid == 0) { ps[1] = S; ps[2] = G; }
  } else if (d.kind == VINET_LOSS_CC || d.kind == VINET_LOSS_NSS) {
    float ls = 0.f, lg = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) { ls += s[i]; lg += g[i]; }
    const double sum_s = block_sum((double)ls, sh), sum_g = block_sum((double)lg, sh);
    const float ms = (float)(sum_s / n), mg = (float)(sum_g / n);
    float ss = 0.f, gg = 0.f, sg = 0.f, sf = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float a = s[i] - ms, c = g[i] - mg;
      ss = fmaf(a, a, ss); gg = fmaf(c, c, gg); sg = fmaf(a, c, sg);
      sf = fmaf(a, g[i], sf);
    }
    const double Sss = block_sum((double)ss, sh), Sgg = block_sum((double)gg, sh), Ssg = block_sum((double)sg, sh);
    const double P = block_sum((double)sf, sh);
    if (d.kind == VINET_LOSS_CC) {
      value = (float)(Ssg / sqrt(Sss * Sgg));
      if (tid == 0) { ps[1] = ms; ps[2] = mg; ps[3] = (float)Sss; ps[4] = (float)Sgg; ps[5] = (float)Ssg; }
    } else {
      const float sigma = (float)sqrt(Sss / (double)(n - 1));
      value = (float)(P / ((double)(sigma + LOSS_EPS) * sum_g));
      if (tid == 0) { ps[1] = ms; ps[2] = sigma; ps[3] = (float)P; ps[4] = (float)sum_g; }
    }
  } else {  // similarity
    float mns = INFINITY, mng = INFINITY, mxs = -INFINITY, mxg = -INFINITY;
    int is = 0x7fffffff, ig = 0x7fffffff;
    float ls = 0.f, lg = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float a = s[i], c = g[i];
      if (a < mns) { mns = a; is = i; }
      if (c < mng) { mng = c; ig = i; }
      mxs = fmaxf(mxs, a); mxg = fmaxf(mxg, c);
      ls += a; lg += c;
    }
    block_argmin(mns, is, shv, shi);
    block_argmin(mng, ig, shv, shi);
    mxs = block_max(mxs, shv);
    mxg = block_max(mxg, shv);
    // normalize_map then divide by the sum, as the reference does (loss.py:62-74)
    const float ds = mxs - mns, dg = mxg - mng;
    float l1 = 0.f, l2 = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) { l1 += (s[i] - mns) / ds; l2 += (g[i] - mng) / dg; }
    const float Us = (float)block_sum((double)l1, sh), Ug = (float)block_sum((double)l2, sh);
    float acc = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) acc += fminf(((s[i] - mns) / ds) / Us, ((g[i] - mng) / dg) / Ug);
    value = (float)block_sum((double)acc, sh);
    if (tid == 0) {
      ps[1] = mns; ps[2] = ds; ps[3] = Us; ps[4] = mng; ps[5] = dg; ps[6] = Ug; ps[7] = __int_as_float(is);
    }
  }
  if (tid == 0) {
    ps[0] = value;
    __threadfence();
    const int done = atomicAdd(d.counter, 1);
    is_last = (done == d.B - 1);
  }
  __syncthreads();
  if (is_last && tid == 0) {
    __threadfence();
    float tot = 0.f;
    for (int i = 0; i < d.B; ++i) tot += reinterpret_cast<volatile float*>(d.per_sample)[i * 8];
    d.out[0] = tot / (float)d.B;
    *d.counter = 0;
  }
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(const __grid_constant__ vinet_loss_t d) {
  __shared__ double sh[LOSS_THREADS / 32];
  const int b = blockIdx.x, n = d.n, tid = threadIdx.x;
  const float* __restrict__ s = d.s + (int64_t)b * n;
  const float* __restrict__ g = d.g + (int64_t)b * n;
  float* __restrict__ gs = d.grad_s + (int64_t)b * n;
  const float* ps = d.per_sample + b * 8;
  const float up = d.gout[0] / (float)d.B;
  if (d.kind == VINET_LOSS_KLDIV) {
    const float S = ps[1], G = ps[2];
    // q_i = dL/ds_n,i ; dL/ds_i = (q_i - sum_j q_j s_n,j) / S
    float acc = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float sn = s[i] / S, gn = g[i] / G;
      const float den = sn + LOSS_EPS;
      const float q = -gn * gn / ((LOSS_EPS + gn / den) * den * den);
      acc = fmaf(q, sn, acc);
    }
    const float Q = (float)block_sum((double)acc, sh);
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float sn = s[i] / S, gn = g[i] / G;
      const float den = sn + LOSS_EPS;
      const float q = -gn * gn / ((LOSS_EPS + gn / den) * den * den);
      gs[i] = up * (q - Q) / S;
    }
  } else if (d.kind == VINET_LOSS_CC) {
    const float ms = ps[1], mg = ps[2], Sss = ps[3], Sgg = ps[4], Ssg = ps[5];
    const float inv = 1.f / sqrtf(Sss * Sgg);
    const float r = Ssg * inv;
    const float k = r * sqrtf(Sgg / Sss);
    for (int i = tid; i < n; i += LOSS_THREADS) gs[i] = up * ((g[i] - mg) - k * (s[i] - ms)) * inv;
  } else if (d.kind == VINET_LOSS_NSS) {
    const float ms = ps[1], sigma = ps[2], P = ps[3], F = ps[4];
    const float se = sigma + LOSS_EPS;
    const float k1 = 1.f / (se * F), k2 = P / (F * se * se * (float)(n - 1) * sigma);
    const float fbar = F / (float)n;
    for (int i = tid; i < n; i += LOSS_THREADS) gs[i] = up * ((g[i] - fbar) * k1 - k2 * (s[i] - ms));
  } else {
    const float mns = ps[1], ds = ps[2], Us = ps[3], mng = ps[4], dg = ps[5], Ug = ps[6];
    const int istar = __float_as_int(ps[7]);
    // s''_i = u_i / U with u_i = s_i - min, U = sum_j u_j  (the max cancels)
    const float U = Us * ds;
    float a1 = 0.f, a2 = 0.f;
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float sv = ((s[i] - mns) / ds) / Us, gv = ((g[i] - mng) / dg) / Ug;
      if (sv < gv) { a1 += 1.f; a2 += (s[i] - mns); }
    }
    const float A = (float)block_sum((double)a1, sh), Bq = (float)block_sum((double)a2, sh);
    for (int i = tid; i < n; i += LOSS_THREADS) {
      const float sv = ((s[i] - mns) / ds) / Us, gv = ((g[i] - mng) / dg) / Ug;
      float v = ((sv < gv) ? 1.f / U : 0.f) - Bq / (U * U);
      if (i == istar) v += -A / U + (float)n * Bq / (U * U);
      gs[i] = up * v;
    }
  }
}

}  // namespace vinet
using namespace vinet;

extern "C" int vinet_loss_fwd(const vinet_loss_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->kind >= 0 && d->kind <= 3 && d->B >= 1 && d->n >= 2, "loss: bad arguments");
  loss_fwd_kernel<<<d->B, LOSS_THREADS, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("loss_fwd");
  return 0;
}

extern "C" int vinet_loss_bwd(const vinet_loss_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->kind >= 0 && d->kind <= 3 && d->B >= 1 && d->n >= 2, "loss: bad arguments");
  loss_bwd_kernel<<<d->B, LOSS_THREADS, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("loss_bwd");
  return 0;
}
