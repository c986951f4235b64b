// BatchNorm3d statistics / finalize / backward on NDHWC views (model_utils.py:132,145,149).
// The normalisation itself is never a kernel: consumers apply scale/shift(+ReLU) when they read.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vinet {

// Column reduction skeleton: blockDim = (G = C/8 channel groups, Ry rows in flight).
// Each thread accumulates NV fp32 partial vectors of 8 channels over its rows, the block combines
// them in shared memory and issues one double atomicAdd per channel.
// Returns true in every thread of the LAST block to finish (ticket counter at sums[NV*C], left at zero): that block may read
// the completed sums and run the layer's finalisation, which saves a dependent tiny launch per BatchNorm layer.
template <int NV, bool TICKET, typename F>
__device__ __forceinline__ bool column_reduce(int64_t rows, int C, int64_t rows_per_block, double* sums, F&& body) {
  extern __shared__ float red[];  // [Ry][NV][C]
  const int gx = threadIdx.x, ry = threadIdx.y, Ry = blockDim.y;
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
#pragma unroll 4
  for (int64_t r = r_begin + ry; r < r_end; r += Ry) body(r, gx * 8, acc);
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[((size_t)ry * NV + v) * C + gx * 8 + e] = acc[v][e];
  __syncthreads();
  const int tid = ry * blockDim.x + gx, nthr = blockDim.x * blockDim.y;
  for (int i = tid; i < NV * C; i += nthr) {
    double s = 0.0;
    for (int y = 0; y < Ry; ++y) s += (double)red[(size_t)y * NV * C + i];
    atomicAdd(sums + i, s);
  }
  if constexpr (!TICKET) return false;
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(sums + NV * C);
    const unsigned long long t = atomicAdd(ticket, 1ull);
    is_last = (t == (unsigned long long)gridDim.x - 1);
    if (is_last) *ticket = 0ull;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last != 0;
}

template <typename T>
__global__ void bn_stats_kernel(const __grid_constant__ vinet_bn_stats_t d, int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  column_reduce<2, false>(d.rows, d.C, rows_per_block, d.sums, [&](int64_t r, int c, float (&acc)[2][8]) {
    float v[8];
    load8(y + r * d.ld + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += v[e];
      acc[1][e] = fmaf(v[e], v[e], acc[1][e]);
    }
  });
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int C, double* sums,
                              int64_t rows_per_block) {
  column_reduce<1, false>(rows, C, rows_per_block, sums, [&](int64_t r, int c, float (&acc)[1][8]) {
    float v[8];
    load8(x + r * ld + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] += v[e];
  });
}

__global__ void colsum_finish_kernel(const double* sums, float* out, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)sums[c];
}

__device__ __forceinline__ void bn_finalize_channel(const vinet_bn_finalize_t& d, int c) {
  float mean, invstd;
  if (d.training) {
    const double n = (double)d.rows;
    volatile double* vs = d.sums;   // written by other blocks' atomics
    const double m = vs[c] / n;
    double var = vs[d.C + c] / n - m * m;
    d.sums[c] = 0.0;  // leave the accumulators clean for the next step
    d.sums[d.C + c] = 0.0;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)d.eps));
    if (d.running_mean) {
      const double unbiased = (d.rows > 1) ? var * n / (n - 1.0) : var;
      d.running_mean[c] = (float)((1.0 - d.momentum) * d.running_mean[c] + d.momentum * m);
      d.running_var[c] = (float)((1.0 - d.momentum) * d.running_var[c] + d.momentum * unbiased);
    }
  } else {
    mean = d.running_mean[c];
    invstd = 1.0f / sqrtf(d.running_var[c] + d.eps);
  }
  const float sc = d.gamma[c] * invstd;
  d.scale[c] = sc;
  d.shift[c] = d.beta[c] - mean * sc;
  d.mean[c] = mean;
  d.invstd[c] = invstd;
}

__global__ void bn_finalize_kernel(const __grid_constant__ vinet_bn_finalize_t d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d.C) bn_finalize_channel(d, c);
}

// statistics + finalisation in one launch: the last block to finish turns the sums into scale/shift/mean/invstd
template <typename T>
__global__ void bn_stats_finalize_kernel(const __grid_constant__ vinet_bn_stats_t d, const __grid_constant__ vinet_bn_finalize_t f,
                                         int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const bool last = column_reduce<2, true>(d.rows, d.C, rows_per_block, d.sums, [&](int64_t r, int c, float (&acc)[2][8]) {
    float v[8];
    load8(y + r * d.ld + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += v[e];
      acc[1][e] = fmaf(v[e], v[e], acc[1][e]);
    }
  });
  if (last)
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < f.C; c += blockDim.x * blockDim.y) bn_finalize_channel(f, c);
}

template <typename T, typename TG>
__global__ void bn_bwd_reduce_kernel(const __grid_constant__ vinet_bn_bwd_t d, int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const TG* __restrict__ gp = reinterpret_cast<const TG*>(d.g);
  float sc[8], sh[8], mu[8], is[8];  // this thread's 8 channels never change: keep their constants in registers
  {
    const int c = threadIdx.x * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e);
      mu[e] = __ldg(d.mean + c + e); is[e] = __ldg(d.invstd + c + e);
    }
  }
  const bool relu = d.relu != 0;
  const bool last = column_reduce<2, true>(d.rows, d.C, rows_per_block, d.sums, [&](int64_t r, int c, float (&acc)[2][8]) {
    float v[8], g[8];
    load8(y + r * d.ldy + c, v);
    load8(gp + r * d.ldg + c, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yh = fmaf(v[e], sc[e], sh[e]);
      const float gm = (relu && !(yh > 0.f)) ? 0.f : g[e];
      const float yn = (v[e] - mu[e]) * is[e];
      acc[0][e] += gm;
      acc[1][e] = fmaf(gm, yn, acc[1][e]);
    }
  });
  if (last) {  // the last block to finish publishes dbeta / dgamma and clears the accumulators for the next step
    volatile double* vs = d.sums;
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < d.C; c += blockDim.x * blockDim.y) {
      d.dbeta[c] = (float)vs[c];
      d.dgamma[c] = (float)vs[d.C + c];
      d.sums[c] = 0.0;
      d.sums[d.C + c] = 0.0;
    }
  }
}

// ---- whole-layer kernels (cooperative launch: every block is resident) ------------------------------------------------------------
// A BatchNorm layer is reduce -> (tiny) finalise -> element-wise apply.  Most of the 77 layers of this model are small (a few
// MB), so three dependent launches cost more in latency than in bandwidth.  Here one launch does all three: every block reduces
// its rows, the last block to arrive finalises and raises a flag, every block waits for the flag and applies to the SAME rows
// (which it has just pulled through L2).  Flag protocol after the sums: [ticket][flag][departures], all zero between launches.
__device__ __forceinline__ void grid_wait_flag(double* sums, int nv_c, bool last) {
  unsigned long long* flag = reinterpret_cast<unsigned long long*>(sums + nv_c + 1);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (last) {
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(flag, 1ull);
  }
  if (tid == 0) {
    unsigned ns = 32;
    while (atomicAdd(flag, 0ull) == 0ull) {
      __nanosleep(ns);
      if (ns < 1024) ns *= 2;
    }
  }
  __syncthreads();
  __threadfence();
}
__device__ __forceinline__ void grid_depart(double* sums, int nv_c) {
  unsigned long long* flag = reinterpret_cast<unsigned long long*>(sums + nv_c + 1);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  __syncthreads();
  if (tid == 0) {
    const unsigned long long t = atomicAdd(flag + 1, 1ull);
    if (t == (unsigned long long)gridDim.x - 1) {   // everyone has seen the flag: leave the words clean for the next launch
      flag[1] = 0ull;
      __threadfence();
      atomicExch(flag, 0ull);
    }
  }
}

template <typename T, typename TO>
__global__ void bn_fwd_fused_kernel(const __grid_constant__ vinet_bn_stats_t d, const __grid_constant__ vinet_bn_finalize_t f,
                                    const __grid_constant__ vinet_bn_apply_t a, int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const bool last = column_reduce<2, true>(d.rows, d.C, rows_per_block, d.sums, [&](int64_t r, int c, float (&acc)[2][8]) {
    float v[8];
    load8(y + r * d.ld + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += v[e];
      acc[1][e] = fmaf(v[e], v[e], acc[1][e]);
    }
  });
  if (last)
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < f.C; c += blockDim.x * blockDim.y) bn_finalize_channel(f, c);
  grid_wait_flag(d.sums, 2 * d.C, last);
  // ---- apply to this block's rows
  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
  const int c = threadIdx.x * 8, Ry = blockDim.y;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = __ldcg(a.scale + c + e); sh[e] = __ldcg(a.shift + c + e); }
  const bool relu = a.relu != 0;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(a.rows, r_begin + rows_per_block);
#pragma unroll 4
  for (int64_t r = r_begin + threadIdx.y; r < r_end; r += Ry) {
    float v[8];
    load8(y + r * a.ldy + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = fmaf(v[e], sc[e], sh[e]);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(out + r * a.ldo + c, v);
  }
  grid_depart(d.sums, 2 * d.C);
}

template <typename T, typename TD, typename TG>
__global__ void bn_bwd_fused_kernel(const __grid_constant__ vinet_bn_bwd_t d, int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const TG* __restrict__ gp = reinterpret_cast<const TG*>(d.g);
  TD* __restrict__ dy = reinterpret_cast<TD*>(d.dy);
  const int c = threadIdx.x * 8, Ry = blockDim.y;
  float sc[8], sh[8], mu[8], is[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e);
    mu[e] = __ldg(d.mean + c + e); is[e] = __ldg(d.invstd + c + e);
  }
  const bool relu = d.relu != 0;
  const bool last = column_reduce<2, true>(d.rows, d.C, rows_per_block, d.sums, [&](int64_t r, int cc, float (&acc)[2][8]) {
    float v[8], g[8];
    load8(y + r * d.ldy + cc, v);
    load8(gp + r * d.ldg + cc, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yh = fmaf(v[e], sc[e], sh[e]);
      const float gm = (relu && !(yh > 0.f)) ? 0.f : g[e];
      const float yn = (v[e] - mu[e]) * is[e];
      acc[0][e] += gm;
      acc[1][e] = fmaf(gm, yn, acc[1][e]);
    }
  });
  if (last) {
    volatile double* vs = d.sums;
    for (int ch = threadIdx.y * blockDim.x + threadIdx.x; ch < d.C; ch += blockDim.x * blockDim.y) {
      d.dbeta[ch] = (float)vs[ch];
      d.dgamma[ch] = (float)vs[d.C + ch];
      d.sums[ch] = 0.0;
      d.sums[d.C + ch] = 0.0;
    }
  }
  grid_wait_flag(d.sums, 2 * d.C, last);
  const float inv_n = 1.0f / (float)d.rows;
  float k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    k1[e] = d.training ? __ldcg(d.dbeta + c + e) * inv_n : 0.f;
    k2[e] = d.training ? __ldcg(d.dgamma + c + e) * inv_n : 0.f;
  }
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(d.rows, r_begin + rows_per_block);
  // second pass over g and y: walk the block's rows BACKWARDS - the reduction pass above read them front to back, so the tail of
  // the chunk is what the 126 MB L2 still holds (every block is resident: cooperative launch)
  const int64_t n_it = r_end > r_begin + threadIdx.y ? (r_end - r_begin - threadIdx.y + Ry - 1) / Ry : 0;
#pragma unroll 4
  for (int64_t it = n_it - 1; it >= 0; --it) {
    const int64_t r = r_begin + threadIdx.y + it * Ry;
    float v[8], g[8], o[8];
    load8(y + r * d.ldy + c, v);
    load8(gp + r * d.ldg + c, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yh = fmaf(v[e], sc[e], sh[e]);
      const float gm = (relu && !(yh > 0.f)) ? 0.f : g[e];
      const float yn = (v[e] - mu[e]) * is[e];
      o[e] = sc[e] * (gm - k1[e] - yn * k2[e]);
    }
    store8(dy + r * d.lddy + c, o);
  }
  grid_depart(d.sums, 2 * d.C);
}

// Elementwise passes use the same (channel group, row) thread layout as the reductions: a thread owns 8 channels,
// keeps their per-channel constants in registers and walks rows with several 128-bit loads in flight.
template <typename T, typename TD, typename TG>
__global__ void bn_bwd_apply_kernel(const __grid_constant__ vinet_bn_bwd_t d, int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const TG* __restrict__ gp = reinterpret_cast<const TG*>(d.g);
  TD* __restrict__ dy = reinterpret_cast<TD*>(d.dy);
  const int c = threadIdx.x * 8, Ry = blockDim.y;
  const float inv_n = 1.0f / (float)d.rows;
  float sc[8], sh[8], mu[8], is[8], k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e);
    mu[e] = __ldg(d.mean + c + e); is[e] = __ldg(d.invstd + c + e);
    k1[e] = d.training ? __ldg(d.dbeta + c + e) * inv_n : 0.f;
    k2[e] = d.training ? __ldg(d.dgamma + c + e) * inv_n : 0.f;
  }
  const bool relu = d.relu != 0;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(d.rows, r_begin + rows_per_block);
#pragma unroll 4
  for (int64_t r = r_begin + threadIdx.y; r < r_end; r += Ry) {
    float v[8], g[8], o[8];
    load8(y + r * d.ldy + c, v);
    load8(gp + r * d.ldg + c, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yh = fmaf(v[e], sc[e], sh[e]);
      const float gm = (relu && !(yh > 0.f)) ? 0.f : g[e];
      const float yn = (v[e] - mu[e]) * is[e];
      o[e] = sc[e] * (gm - k1[e] - yn * k2[e]);
    }
    store8(dy + r * d.lddy + c, o);
  }
}

template <typename T, typename TO>
__global__ void bn_apply_kernel(const __grid_constant__ vinet_bn_apply_t d, int64_t rows_per_block) {
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  TO* __restrict__ out = reinterpret_cast<TO*>(d.out);
  const int c = threadIdx.x * 8, Ry = blockDim.y;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e); }
  const bool relu = d.relu != 0;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(d.rows, r_begin + rows_per_block);
#pragma unroll 4
  for (int64_t r = r_begin + threadIdx.y; r < r_end; r += Ry) {
    float v[8];
    load8(y + r * d.ldy + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = fmaf(v[e], sc[e], sh[e]);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(out + r * d.ldo + c, v);
  }
}

// ---- multi-layer launches ---------------------------------------------------------------------------------------------------------
// Most BatchNorm layers of this model are small (a Mixed block has seven, a few MB each) and cost ~6 us per launch whatever
// they do.  Layers that are ready at the same time (the three fused 1x1 branches + the pool branch, the two conv_s, the two
// conv_t of a Mixed block) are therefore processed by ONE launch: blockIdx.y selects the layer ("segment"), every segment
// brings its own descriptor, block count and sums / ticket buffer.  Blocks are 1-D (256 threads) and derive the
// (channel group, row lane) mapping from their segment's channel count.
constexpr int BN_MAX_SEG = 4;
constexpr int BN_MT = 256;

struct SegMap {
  int gx, ry, Ry;
  bool active;
};
__device__ __forceinline__ SegMap seg_map(int C) {
  SegMap m;
  const int G = C / 8;
  m.Ry = BN_MT / G;
  m.gx = threadIdx.x % G;
  m.ry = threadIdx.x / G;
  m.active = m.ry < m.Ry;
  return m;
}

template <int NV, typename F>
__device__ __forceinline__ bool column_reduce_1d(const SegMap& m, int64_t rows, int C, int64_t rows_per_block, int nblocks, double* sums,
                                                 F&& body) {
  extern __shared__ float red[];  // [Ry][NV][C]
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  if (m.active) {
#pragma unroll 4
    for (int64_t r = r_begin + m.ry; r < r_end; r += m.Ry) body(r, m.gx * 8, acc);
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[((size_t)m.ry * NV + v) * C + m.gx * 8 + e] = acc[v][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV * C; i += BN_MT) {
    double s = 0.0;
    for (int y = 0; y < m.Ry; ++y) s += (double)red[(size_t)y * NV * C + i];
    atomicAdd(sums + i, s);
  }
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(sums + NV * C);
    const unsigned long long t = atomicAdd(ticket, 1ull);
    is_last = (t == (unsigned long long)nblocks - 1);
    if (is_last) *ticket = 0ull;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last != 0;
}

struct BnFwdMulti {
  vinet_bn_stats_t s[BN_MAX_SEG];
  vinet_bn_finalize_t f[BN_MAX_SEG];
  int64_t rpb[BN_MAX_SEG];
  int32_t nb[BN_MAX_SEG];
};
struct BnApplyMulti {
  vinet_bn_apply_t a[BN_MAX_SEG];
  int64_t rpb[BN_MAX_SEG];
  int32_t nb[BN_MAX_SEG];
};
struct BnBwdMulti {
  vinet_bn_bwd_t b[BN_MAX_SEG];
  int64_t rpb[BN_MAX_SEG];
  int32_t nb[BN_MAX_SEG];
};

template <typename T>
__global__ void __launch_bounds__(BN_MT) bn_stats_finalize_multi_kernel(const __grid_constant__ BnFwdMulti p) {
  const int seg = blockIdx.y;
  if ((int)blockIdx.x >= p.nb[seg]) return;
  const vinet_bn_stats_t& d = p.s[seg];
  const vinet_bn_finalize_t& f = p.f[seg];
  const SegMap m = seg_map(d.C);
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const bool last = column_reduce_1d<2>(m, d.rows, d.C, p.rpb[seg], p.nb[seg], d.sums, [&](int64_t r, int c, float (&acc)[2][8]) {
    float v[8];
    load8(y + r * d.ld + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += v[e];
      acc[1][e] = fmaf(v[e], v[e], acc[1][e]);
    }
  });
  if (last)
    for (int c = threadIdx.x; c < f.C; c += BN_MT) bn_finalize_channel(f, c);
}

template <typename T, typename TO>
__global__ void __launch_bounds__(BN_MT) bn_apply_multi_kernel(const __grid_constant__ BnApplyMulti p) {
  const int seg = blockIdx.y;
  if ((int)blockIdx.x >= p.nb[seg]) return;
  const vinet_bn_apply_t& d = p.a[seg];
  const SegMap m = seg_map(d.C);
  if (!m.active) return;
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  TO* __restrict__ out = reinterpret_cast<TO*>(d.out);
  const int c = m.gx * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e); }
  const bool relu = d.relu != 0;
  // the convolution wrote the raw output front to back: read it back to front (blocks in reverse order, rows downwards), so that
  // the part the L2 still holds is consumed before this kernel's own traffic evicts it
  const int64_t r_begin = (int64_t)(p.nb[seg] - 1 - (int)blockIdx.x) * p.rpb[seg];
  const int64_t r_end = min(d.rows, r_begin + p.rpb[seg]);
  const int64_t n_it = r_end > r_begin + m.ry ? (r_end - r_begin - m.ry + m.Ry - 1) / m.Ry : 0;
#pragma unroll 4
  for (int64_t it = n_it - 1; it >= 0; --it) {
    const int64_t r = r_begin + m.ry + it * m.Ry;
    float v[8];
    load8(y + r * d.ldy + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = fmaf(v[e], sc[e], sh[e]);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(out + r * d.ldo + c, v);
  }
}

// Finalisation + materialisation for layers whose statistics were accumulated by the convolution epilogue (vinet_conv_t.stats):
// every thread derives scale / shift of its 8 channels from the fp64 sums (same arithmetic as bn_finalize_channel); the first
// block of a layer publishes scale / shift / mean / invstd for the backward pass and updates the running statistics.
struct BnApplyStatsMulti {
  vinet_bn_finalize_t f[BN_MAX_SEG];
  vinet_bn_apply_t a[BN_MAX_SEG];
  int64_t rpb[BN_MAX_SEG];
  double inv_n[BN_MAX_SEG];   // 1 / rows, computed on the host (an fp64 division costs microseconds of latency on the device)
  int32_t nb[BN_MAX_SEG];
};

template <typename T, typename TO>
__global__ void __launch_bounds__(BN_MT) bn_apply_stats_multi_kernel(const __grid_constant__ BnApplyStatsMulti p) {
  const int seg = blockIdx.y;
  if ((int)blockIdx.x >= p.nb[seg]) return;
  const vinet_bn_apply_t& d = p.a[seg];
  const vinet_bn_finalize_t& f = p.f[seg];
  const SegMap m = seg_map(d.C);
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  TO* __restrict__ out = reinterpret_cast<TO*>(d.out);
  const int c = m.gx * 8;
  const int64_t sqs = f.sq_stride > 0 ? f.sq_stride : f.C;
  // B200 has few fp64 lanes: ONE thread per channel group (row lane 0) turns the fp64 sums into fp32 scale / shift and shares
  // them through shared memory (measured: with every thread doing it a 33 MB layer took 64 us instead of 10)
  __shared__ float s_sc[1024], s_sh[1024];
  if (m.ry == 0) {
    const bool publish = blockIdx.x == 0;
    const double inv_n = p.inv_n[seg];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const double mu = __ldcg(f.sums + c + e) * inv_n;
      const double var_d = fma(-mu, mu, __ldcg(f.sums + sqs + c + e) * inv_n);
      const float var = fmaxf((float)var_d, 0.f);
      const float mean = (float)mu;
      const float invstd = rsqrtf(var + f.eps);
      const float scv = __ldg(f.gamma + c + e) * invstd;
      const float shv = __ldg(f.beta + c + e) - mean * scv;
      s_sc[c + e] = scv;
      s_sh[c + e] = shv;
      if (publish) {
        f.scale[c + e] = scv;
        f.shift[c + e] = shv;
        f.mean[c + e] = mean;
        f.invstd[c + e] = invstd;
        if (f.running_mean) {
          const float n = (float)f.rows;
          const float unbiased = (f.rows > 1) ? var * (n / (n - 1.f)) : var;
          f.running_mean[c + e] = (1.f - f.momentum) * f.running_mean[c + e] + f.momentum * mean;
          f.running_var[c + e] = (1.f - f.momentum) * f.running_var[c + e] + f.momentum * unbiased;
        }
      }
    }
  }
  __syncthreads();
  if (!m.active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = s_sc[c + e]; sh[e] = s_sh[c + e]; }
  const bool relu = d.relu != 0;
  // the convolution wrote the raw output front to back: read it back to front (blocks in reverse order, rows downwards), so that
  // the part the L2 still holds is consumed before this kernel's own traffic evicts it
  const int64_t r_begin = (int64_t)(p.nb[seg] - 1 - (int)blockIdx.x) * p.rpb[seg];
  const int64_t r_end = min(d.rows, r_begin + p.rpb[seg]);
  const int64_t n_it = r_end > r_begin + m.ry ? (r_end - r_begin - m.ry + m.Ry - 1) / m.Ry : 0;
#pragma unroll 4
  for (int64_t it = n_it - 1; it >= 0; --it) {
    const int64_t r = r_begin + m.ry + it * m.Ry;
    float v[8];
    load8(y + r * d.ldy + c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = fmaf(v[e], sc[e], sh[e]);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(out + r * d.ldo + c, v);
  }
}

template <typename T, typename TG>
__global__ void __launch_bounds__(BN_MT) bn_bwd_reduce_multi_kernel(const __grid_constant__ BnBwdMulti p) {
  const int seg = blockIdx.y;
  if ((int)blockIdx.x >= p.nb[seg]) return;
  const vinet_bn_bwd_t& d = p.b[seg];
  const SegMap m = seg_map(d.C);
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const TG* __restrict__ gp = reinterpret_cast<const TG*>(d.g);
  float sc[8], sh[8], mu[8], is[8];
  {
    const int c = m.gx * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e);
      mu[e] = __ldg(d.mean + c + e); is[e] = __ldg(d.invstd + c + e);
    }
  }
  const bool relu = d.relu != 0;
  const bool last = column_reduce_1d<2>(m, d.rows, d.C, p.rpb[seg], p.nb[seg], d.sums, [&](int64_t r, int c, float (&acc)[2][8]) {
    float v[8], g[8];
    load8(y + r * d.ldy + c, v);
    load8(gp + r * d.ldg + c, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yh = fmaf(v[e], sc[e], sh[e]);
      const float gm = (relu && !(yh > 0.f)) ? 0.f : g[e];
      const float yn = (v[e] - mu[e]) * is[e];
      acc[0][e] += gm;
      acc[1][e] = fmaf(gm, yn, acc[1][e]);
    }
  });
  if (last) {
    volatile double* vs = d.sums;
    for (int c = threadIdx.x; c < d.C; c += BN_MT) {
      d.dbeta[c] = (float)vs[c];
      d.dgamma[c] = (float)vs[d.C + c];
      d.sums[c] = 0.0;
      d.sums[d.C + c] = 0.0;
    }
  }
}

template <typename T, typename TD, typename TG>
__global__ void __launch_bounds__(BN_MT) bn_bwd_apply_multi_kernel(const __grid_constant__ BnBwdMulti p) {
  const int seg = blockIdx.y;
  if ((int)blockIdx.x >= p.nb[seg]) return;
  const vinet_bn_bwd_t& d = p.b[seg];
  const SegMap m = seg_map(d.C);
  if (!m.active) return;
  const T* __restrict__ y = reinterpret_cast<const T*>(d.y);
  const TG* __restrict__ gp = reinterpret_cast<const TG*>(d.g);
  TD* __restrict__ dy = reinterpret_cast<TD*>(d.dy);
  const int c = m.gx * 8;
  const float inv_n = 1.0f / (float)d.rows;
  float sc[8], sh[8], mu[8], is[8], k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = __ldg(d.scale + c + e); sh[e] = __ldg(d.shift + c + e);
    mu[e] = __ldg(d.mean + c + e); is[e] = __ldg(d.invstd + c + e);
    k1[e] = d.training ? __ldg(d.dbeta + c + e) * inv_n : 0.f;
    k2[e] = d.training ? __ldg(d.dgamma + c + e) * inv_n : 0.f;
  }
  const bool relu = d.relu != 0;
  const int64_t r_begin = (int64_t)blockIdx.x * p.rpb[seg];
  const int64_t r_end = min(d.rows, r_begin + p.rpb[seg]);
#pragma unroll 4
  for (int64_t r = r_begin + m.ry; r < r_end; r += m.Ry) {
    float v[8], g[8], o[8];
    load8(y + r * d.ldy + c, v);
    load8(gp + r * d.ldg + c, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float yh = fmaf(v[e], sc[e], sh[e]);
      const float gm = (relu && !(yh > 0.f)) ? 0.f : g[e];
      const float yn = (v[e] - mu[e]) * is[e];
      o[e] = sc[e] * (gm - k1[e] - yn * k2[e]);
    }
    store8(dy + r * d.lddy + c, o);
  }
}

struct ColGrid {
  dim3 block;
  unsigned grid;
  int64_t rows_per_block;
  size_t smem;
};
// reduce: column reductions pay 2C fp64 atomics per block, so they stop at two blocks per SM; element-wise passes take as
// many blocks as leave every thread at least two rows.  Either way small tensors (most of the 77 BatchNorm layers of this
// model) get enough blocks to cover the machine instead of a few blocks walking 16+ rows per thread.
static ColGrid col_grid(int64_t rows, int C, int nv, bool reduce) {
  ColGrid g;
  const int G = C / 8;
  int Ry = 256 / G;
  if (Ry < 1) Ry = 1;
  if (Ry > 128) Ry = 128;
  g.block = dim3(G, Ry);
  // 8 blocks per SM keep enough loads in flight to saturate HBM; a reduction also pays 2C fp64 atomics per block, so wide
  // layers stop earlier.  Every thread keeps >= 8 rows: its per-channel constants (up to 48 loads) are set up once.
  static const int env_maxb = getenv("VINET_BN_MAXB") ? atoi(getenv("VINET_BN_MAXB")) : 0;        // tuning knobs (tools/bn_bench.py)
  static const int env_minrpt = getenv("VINET_BN_MINRPT") ? atoi(getenv("VINET_BN_MINRPT")) : 0;
  static const int env_atom = getenv("VINET_BN_ATOM") ? atoi(getenv("VINET_BN_ATOM")) : 160000;
  int64_t max_blocks = env_maxb ? env_maxb : 148 * 8;
  if (reduce) max_blocks = std::min<int64_t>(max_blocks, std::max<int64_t>(148, env_atom / (2 * C)));
  int64_t rpt = cdiv(rows, (int64_t)Ry * max_blocks);   // rows per thread
  const int64_t min_rpt = env_minrpt ? env_minrpt : 8;
  if (rpt < min_rpt) rpt = min_rpt;
  int64_t rpb = (int64_t)Ry * rpt;
  int64_t nb = cdiv(rows, rpb);
  g.grid = (unsigned)(nb < 1 ? 1 : nb);
  g.rows_per_block = rpb;
  g.smem = (size_t)Ry * nv * C * sizeof(float);
  return g;
}

}  // namespace vinet

using namespace vinet;

extern "C" int vinet_bn_stats(const vinet_bn_stats_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= 1024, "bn_stats: C %d", d->C);
  ColGrid g = col_grid(d->rows, d->C, 2, true);
  VINET_DISPATCH_DTYPE(d->dtype, T, {
    if (g.smem > 48 * 1024) cudaFuncSetAttribute(bn_stats_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    bn_stats_kernel<T><<<g.grid, g.block, g.smem, (cudaStream_t)stream>>>(*d, g.rows_per_block);
  });
  VINET_LAUNCH_OK("bn_stats");
  return 0;
}

extern "C" int vinet_bn_stats_finalize(const vinet_bn_stats_t* d, const vinet_bn_finalize_t* f, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= 1024 && f->C == d->C && f->sums == d->sums && f->training, "bn_stats_finalize: C %d / %d", d->C, f->C);
  ColGrid g = col_grid(d->rows, d->C, 2, true);
  VINET_DISPATCH_DTYPE(d->dtype, T, {
    if (g.smem > 48 * 1024)
      cudaFuncSetAttribute(bn_stats_finalize_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    bn_stats_finalize_kernel<T><<<g.grid, g.block, g.smem, (cudaStream_t)stream>>>(*d, *f, g.rows_per_block);
  });
  VINET_LAUNCH_OK("bn_stats_finalize");
  return 0;
}

// grid of a cooperative whole-layer kernel: col_grid capped at what can be resident at once
template <typename K>
static int coop_grid(K kernel, ColGrid* g, int64_t rows, const char* what) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (g->smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)(g->block.x * g->block.y), g->smem);
  VINET_CHECK(e == cudaSuccess && per_sm >= 1 && sms >= 1, "%s: occupancy query failed", what);
  const int64_t cap = (int64_t)per_sm * sms;
  if ((int64_t)g->grid > cap) {
    const int64_t Ry = g->block.y;
    g->rows_per_block = round_up(cdiv(rows, cap), Ry);
    g->grid = (unsigned)cdiv(rows, g->rows_per_block);
  }
  return 0;
}

extern "C" int vinet_bn_fwd_fused(const vinet_bn_stats_t* d, const vinet_bn_finalize_t* f, const vinet_bn_apply_t* a,
                                  vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= 1024 && f->C == d->C && a->C == d->C && f->sums == d->sums && f->training &&
              a->y == d->y && a->ldy == d->ld && a->rows == d->rows && a->dtype == d->dtype, "bn_fwd_fused: inconsistent descriptors");
  ColGrid g = col_grid(d->rows, d->C, 2, true);
  int64_t rpb = 0;
#define LAUNCH_FWD(T, TO)                                                                             \
  do {                                                                                                \
    if (coop_grid(bn_fwd_fused_kernel<T, TO>, &g, d->rows, "bn_fwd_fused")) return -1;                \
    rpb = g.rows_per_block;                                                                           \
    void* args[] = {(void*)d, (void*)f, (void*)a, (void*)&rpb};                                       \
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bn_fwd_fused_kernel<T, TO>, dim3(g.grid), g.block, args, g.smem, \
                                                (cudaStream_t)stream);                                \
    VINET_CHECK(e == cudaSuccess, "bn_fwd_fused: cooperative launch failed: %s", cudaGetErrorString(e)); \
  } while (0)
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(a->out_dtype, TO, LAUNCH_FWD(T, TO)));
#undef LAUNCH_FWD
  VINET_LAUNCH_OK("bn_fwd_fused");
  return 0;
}

extern "C" int vinet_bn_bwd_fused(const vinet_bn_bwd_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= 1024, "bn_bwd_fused: C %d", d->C);
  ColGrid g = col_grid(d->rows, d->C, 2, true);
  int64_t rpb = 0;
#define LAUNCH_BWD(T, TD, TG)                                                                         \
  do {                                                                                                \
    if (coop_grid(bn_bwd_fused_kernel<T, TD, TG>, &g, d->rows, "bn_bwd_fused")) return -1;            \
    rpb = g.rows_per_block;                                                                           \
    void* args[] = {(void*)d, (void*)&rpb};                                                           \
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bn_bwd_fused_kernel<T, TD, TG>, dim3(g.grid), g.block, args, g.smem, \
                                                (cudaStream_t)stream);                                \
    VINET_CHECK(e == cudaSuccess, "bn_bwd_fused: cooperative launch failed: %s", cudaGetErrorString(e)); \
  } while (0)
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->dy_dtype, TD, VINET_DISPATCH_DTYPE(d->g_dtype, TG, LAUNCH_BWD(T, TD, TG))));
#undef LAUNCH_BWD
  VINET_LAUNCH_OK("bn_bwd_fused");
  return 0;
}

// per-segment block count / rows per block of a multi-layer launch (1-D blocks of BN_MT threads)
static void seg_grid(int64_t rows, int C, bool reduce, int64_t* rpb, int32_t* nb, size_t* smem, int nv) {
  const int G = C / 8, Ry = BN_MT / G;
  int64_t max_blocks = 148 * 8;
  if (reduce) max_blocks = std::min<int64_t>(max_blocks, std::max<int64_t>(148, 160000 / (2 * C)));
  int64_t rpt = cdiv(rows, (int64_t)Ry * max_blocks);
  if (rpt < 8) rpt = 8;
  *rpb = (int64_t)Ry * rpt;
  *nb = (int32_t)std::max<int64_t>(1, cdiv(rows, *rpb));
  *smem = std::max(*smem, (size_t)Ry * nv * C * sizeof(float));
}

extern "C" int vinet_bn_stats_finalize_multi(const vinet_bn_stats_t* d, const vinet_bn_finalize_t* f, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(n >= 1 && n <= BN_MAX_SEG, "bn_stats_finalize_multi: %d segments", n);
  BnFwdMulti p;
  size_t smem = 0;
  int32_t gx = 1;
  for (int i = 0; i < n; ++i) {
    VINET_CHECK(d[i].C % 8 == 0 && d[i].C <= 1024 && f[i].C == d[i].C && f[i].sums == d[i].sums && f[i].training &&
                d[i].dtype == d[0].dtype, "bn_stats_finalize_multi: segment %d", i);
    p.s[i] = d[i];
    p.f[i] = f[i];
    seg_grid(d[i].rows, d[i].C, true, &p.rpb[i], &p.nb[i], &smem, 2);
    gx = std::max(gx, p.nb[i]);
  }
  for (int i = n; i < BN_MAX_SEG; ++i) p.nb[i] = 0;
  VINET_DISPATCH_DTYPE(d[0].dtype, T, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(bn_stats_finalize_multi_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bn_stats_finalize_multi_kernel<T><<<dim3((unsigned)gx, (unsigned)n), BN_MT, smem, (cudaStream_t)stream>>>(p);
  });
  VINET_LAUNCH_OK("bn_stats_finalize_multi");
  return 0;
}

extern "C" int vinet_bn_apply_multi(const vinet_bn_apply_t* a, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(n >= 1 && n <= BN_MAX_SEG, "bn_apply_multi: %d segments", n);
  BnApplyMulti p;
  size_t smem = 0;
  int32_t gx = 1;
  for (int i = 0; i < n; ++i) {
    VINET_CHECK(a[i].C % 8 == 0 && a[i].C <= 1024 && a[i].dtype == a[0].dtype && a[i].out_dtype == a[0].out_dtype, "bn_apply_multi: segment %d", i);
    p.a[i] = a[i];
    seg_grid(a[i].rows, a[i].C, false, &p.rpb[i], &p.nb[i], &smem, 0);
    gx = std::max(gx, p.nb[i]);
  }
  for (int i = n; i < BN_MAX_SEG; ++i) p.nb[i] = 0;
  VINET_DISPATCH_DTYPE(a[0].dtype, T, VINET_DISPATCH_DTYPE(a[0].out_dtype, TO,
      (bn_apply_multi_kernel<T, TO><<<dim3((unsigned)gx, (unsigned)n), BN_MT, 0, (cudaStream_t)stream>>>(p))));
  VINET_LAUNCH_OK("bn_apply_multi");
  return 0;
}

extern "C" int vinet_bn_apply_stats_multi(const vinet_bn_finalize_t* f, const vinet_bn_apply_t* a, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(n >= 1 && n <= BN_MAX_SEG, "bn_apply_stats_multi: %d segments", n);
  BnApplyStatsMulti p;
  size_t smem = 0;
  int32_t gx = 1;
  for (int i = 0; i < n; ++i) {
    VINET_CHECK(a[i].C % 8 == 0 && a[i].C <= 1024 && f[i].C == a[i].C && f[i].rows == a[i].rows && f[i].training && f[i].sums != nullptr &&
                a[i].dtype == a[0].dtype && a[i].out_dtype == a[0].out_dtype, "bn_apply_stats_multi: segment %d", i);
    p.f[i] = f[i];
    p.a[i] = a[i];
    p.inv_n[i] = 1.0 / (double)f[i].rows;
    seg_grid(a[i].rows, a[i].C, false, &p.rpb[i], &p.nb[i], &smem, 0);
    gx = std::max(gx, p.nb[i]);
  }
  for (int i = n; i < BN_MAX_SEG; ++i) p.nb[i] = 0;
  VINET_DISPATCH_DTYPE(a[0].dtype, T, VINET_DISPATCH_DTYPE(a[0].out_dtype, TO,
      (bn_apply_stats_multi_kernel<T, TO><<<dim3((unsigned)gx, (unsigned)n), BN_MT, 0, (cudaStream_t)stream>>>(p))));
  VINET_LAUNCH_OK("bn_apply_stats_multi");
  return 0;
}

extern "C" int vinet_bn_bwd_multi(const vinet_bn_bwd_t* b, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(n >= 1 && n <= BN_MAX_SEG, "bn_bwd_multi: %d segments", n);
  BnBwdMulti pr, pa;
  size_t smem = 0, none = 0;
  int32_t gr = 1, ga = 1;
  for (int i = 0; i < n; ++i) {
    VINET_CHECK(b[i].C % 8 == 0 && b[i].C <= 1024 && b[i].dtype == b[0].dtype && b[i].g_dtype == b[0].g_dtype &&
                b[i].dy_dtype == b[0].dy_dtype, "bn_bwd_multi: segment %d", i);
    pr.b[i] = b[i];
    pa.b[i] = b[i];
    seg_grid(b[i].rows, b[i].C, true, &pr.rpb[i], &pr.nb[i], &smem, 2);
    seg_grid(b[i].rows, b[i].C, false, &pa.rpb[i], &pa.nb[i], &none, 0);
    gr = std::max(gr, pr.nb[i]);
    ga = std::max(ga, pa.nb[i]);
  }
  for (int i = n; i < BN_MAX_SEG; ++i) pr.nb[i] = pa.nb[i] = 0;
  VINET_DISPATCH_DTYPE(b[0].dtype, T, VINET_DISPATCH_DTYPE(b[0].g_dtype, TG, {
    if (smem > 48 * 1024) cudaFuncSetAttribute(bn_bwd_reduce_multi_kernel<T, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bn_bwd_reduce_multi_kernel<T, TG><<<dim3((unsigned)gr, (unsigned)n), BN_MT, smem, (cudaStream_t)stream>>>(pr);
  }));
  VINET_LAUNCH_OK("bn_bwd_reduce_multi");
  VINET_DISPATCH_DTYPE(b[0].dtype, T, VINET_DISPATCH_DTYPE(b[0].dy_dtype, TD, VINET_DISPATCH_DTYPE(b[0].g_dtype, TG,
      (bn_bwd_apply_multi_kernel<T, TD, TG><<<dim3((unsigned)ga, (unsigned)n), BN_MT, 0, (cudaStream_t)stream>>>(pa)))));
  VINET_LAUNCH_OK("bn_bwd_apply_multi");
  return 0;
}

extern "C" int vinet_colsum(const void* x, int64_t ld, int32_t dtype, int64_t rows, int32_t C, double* ws, float* out,
                            vinet_stream_t stream) {
  VINET_CHECK(C % 8 == 0 && C <= 1024, "colsum: C %d", C);
  cudaMemsetAsync(ws, 0, sizeof(double) * C, (cudaStream_t)stream);
  ColGrid g = col_grid(rows, C, 1, true);
  VINET_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<g.grid, g.block, g.smem, (cudaStream_t)stream>>>(
                                     reinterpret_cast<const T*>(x), ld, rows, C, ws, g.rows_per_block)));
  VINET_LAUNCH_OK("colsum");
  colsum_finish_kernel<<<(unsigned)cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(ws, out, C);
  VINET_LAUNCH_OK("colsum_finish");
  return 0;
}

extern "C" int vinet_bn_finalize(const vinet_bn_finalize_t* d, vinet_stream_t stream) {
  bn_finalize_kernel<<<(unsigned)cdiv(d->C, 128), 128, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("bn_finalize");
  return 0;
}

extern "C" int vinet_bn_apply(const vinet_bn_apply_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0, "bn_apply: C %d", d->C);
  VINET_CHECK(d->C <= 1024, "bn_apply: C %d", d->C);
  ColGrid g = col_grid(d->rows, d->C, 0, false);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->out_dtype, TO,
      (bn_apply_kernel<T, TO><<<g.grid, g.block, 0, (cudaStream_t)stream>>>(*d, g.rows_per_block))));
  VINET_LAUNCH_OK("bn_apply");
  return 0;
}

extern "C" int vinet_bn_bwd_reduce(const vinet_bn_bwd_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= 1024, "bn_bwd: C %d", d->C);
  ColGrid g = col_grid(d->rows, d->C, 2, true);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->g_dtype, TG, {
    if (g.smem > 48 * 1024)
      cudaFuncSetAttribute(bn_bwd_reduce_kernel<T, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    bn_bwd_reduce_kernel<T, TG><<<g.grid, g.block, g.smem, (cudaStream_t)stream>>>(*d, g.rows_per_block);
  }));
  VINET_LAUNCH_OK("bn_bwd_reduce");   // the last block to finish writes dgamma / dbeta and clears the sums (no finish launch)
  return 0;
}

extern "C" int vinet_bn_bwd_apply(const vinet_bn_bwd_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0 && d->C <= 1024, "bn_bwd_apply: C %d", d->C);
  ColGrid g = col_grid(d->rows, d->C, 0, false);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->dy_dtype, TD, VINET_DISPATCH_DTYPE(d->g_dtype, TG,
      (bn_bwd_apply_kernel<T, TD, TG><<<g.grid, g.block, 0, (cudaStream_t)stream>>>(*d, g.rows_per_block)))));
  VINET_LAUNCH_OK("bn_bwd_apply");
  return 0;
}
