// Transformer fusion variants of AViNet (reference model.py:8-69, 116-189, 211-221, 239-247): 32-339 tokens of 336-512 features,
// three post-norm encoder layers.  Under 1 % of the step's FLOPs, so everything here is plain fp32 FFMA work, written for few
// launches and no layout copies: ONE strided batched GEMM covers the linear layers, both attention products, the 1x1 convolutions
// with bias, every permute / flatten / cat / mean / repeat of the reference's forward, and all of their gradients.
#include <algorithm>

#include "common.cuh"

namespace vinet {

// ------------------------------------------------------------------ strided batched GEMM
// These products are small (<= 4 GFLOP) and latency-bound, not FLOP-bound: what matters is loads in flight and CTAs in flight.
// A k step covers 32 reduction elements (16 independent global loads per thread), the loads of step i + 1 are issued into
// registers before step i is computed, and outputs with few tiles but a long reduction (weight and bias gradients: K = all token
// rows) are split along K over up to two waves of CTAs that combine with fp32 atomics (gradients only, never a forward result).
constexpr int BG_T = 64;   // C tile (BG_T x BG_T), 256 threads, 4 x 4 outputs per thread
constexpr int BG_K = 32;

__device__ __forceinline__ float bg_load(const void* p, int64_t off, int dtype) {
  return dtype == VINET_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[off]) : __ldg(reinterpret_cast<const float*>(p) + off);
}

// splits > 1: blockIdx.z is a K split of the single batch (k range [z * ksplit_len, ...)), results are added atomically
__global__ void __launch_bounds__(256) bgemm_kernel(const vinet_bgemm_t d, int ksplit_len, int splits) {
  __shared__ float As[BG_K][BG_T + 4];
  __shared__ float Bs[BG_K][BG_T + 4];
  const int tid = threadIdx.x;
  const int bz = splits > 1 ? 0 : (int)blockIdx.z;
  const int b1 = bz / d.nb2, b2 = bz % d.nb2;
  const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
  const int k_begin = splits > 1 ? (int)blockIdx.z * ksplit_len : 0;
  const int k_end = splits > 1 ? min(d.K, k_begin + ksplit_len) : d.K;
  const int64_t offA = b1 * d.sAb1 + b2 * d.sAb2, offB = b1 * d.sBb1 + b2 * d.sBb2;
  // the fastest-running thread index follows the unit-stride axis of each operand
  const bool a_kfast = d.sAk == 1, b_kfast = d.sBk == 1;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      {
        const int kk = a_kfast ? (idx & 31) : (idx >> 6), mm = a_kfast ? (idx >> 5) : (idx & 63);
        const int m = m0 + mm, k = k0 + kk;
        float v = 0.f;
        if (m < d.M && k < k_end) {
          v = bg_load(d.A, offA + m * d.sAm + k * d.sAk, d.a_dtype);
          if (d.a_scale) v = fmaf(v, __ldg(d.a_scale + (d.a_xf_on_m ? m : k)), __ldg(d.a_shift + (d.a_xf_on_m ? m : k)));
          if (d.a_relu) v = fmaxf(v, 0.f);
        }
        ra[i] = v;
      }
      {
        const int kk = b_kfast ? (idx & 31) : (idx >> 6), nn = b_kfast ? (idx >> 5) : (idx & 63);
        const int n = n0 + nn, k = k0 + kk;
        rb[i] = (n < d.N && k < k_end) ? bg_load(d.B, offB + n * d.sBn + k * d.sBk, d.b_dtype) : 0.f;
      }
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += BG_K) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      As[a_kfast ? (idx & 31) : (idx >> 6)][a_kfast ? (idx >> 5) : (idx & 63)] = ra[i];
      Bs[b_kfast ? (idx & 31) : (idx >> 6)][b_kfast ? (idx >> 5) : (idx & 63)] = rb[i];
    }
    __syncthreads();
    if (k0 + BG_K < k_end) fetch(k0 + BG_K);        // in flight while this step is computed
#pragma unroll
    for (int kk = 0; kk < BG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int64_t offC = b1 * d.sCb1 + b2 * d.sCb2, off2 = b1 * d.s2b1 + b2 * d.s2b2;
  // Epilogue through shared memory (As is free now) in four slabs of 16 rows, so that consecutive threads store consecutive
  // addresses whichever of C's two axes has the smaller stride.
  const bool c_nfast = d.sCn <= d.sCm;
  const bool lead = splits <= 1 || blockIdx.z == 0;     // the bias terms are added once
  float* Ct = &As[0][0];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) Ct[ty * (BG_T + 4) + tx + 16 * j] = acc[i][j];     // slab rows m = ty + 16 i, all 64 columns
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // slab element (r, c): r = 0..15 (m = m0 + r + 16 i), c = 0..63
      const int e = tid + j * 256;
      const int r = c_nfast ? (e >> 6) : (e & 15), c = c_nfast ? (e & 63) : (e >> 4);
      const int m = m0 + r + 16 * i, n = n0 + c;
      if (m >= d.M || n >= d.N) continue;
      float v = d.alpha * Ct[r * (BG_T + 4) + c];
      if (lead && d.bias1) v += __ldg(d.bias1 + m * d.s1m + n * d.s1n);
      if (lead && d.bias2) v += __ldg(d.bias2 + off2 + m * d.s2m + n * d.s2n);
      if (d.relu) v = fmaxf(v, 0.f);
      const int64_t o = offC + m * d.sCm + n * d.sCn;
      if (d.c_dtype == VINET_BF16) {
        reinterpret_cast<__nv_bfloat16*>(d.C)[o] = __float2bfloat16_rn(v);
      } else {
        float* cp = reinterpret_cast<float*>(d.C) + o;
        if (splits > 1) atomicAdd(cp, v);
        else *cp = (d.accumulate & 1) ? *cp + v : v;
      }
    }
  }
}

// ------------------------------------------------------------------ softmax rows (one warp per row)
__global__ void __launch_bounds__(128) softmax_fwd_kernel(float* __restrict__ s, int64_t rows, int n) {
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = s + row * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, p[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float e = expf(p[j] - mx);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < n; j += 32) p[j] *= inv;
}

__global__ void __launch_bounds__(128) softmax_bwd_kernel(const float* __restrict__ p, float* __restrict__ dp, int64_t rows, int n) {
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* pr = p + row * n;
  float* gr = dp + row * n;
  float dot = 0.f;
  for (int j = lane; j < n; j += 32) dot = fmaf(gr[j], pr[j], dot);
  dot = warp_sum(dot);
  for (int j = lane; j < n; j += 32) gr[j] = pr[j] * (gr[j] - dot);
}

// ------------------------------------------------------------------ dropout
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
// keep(i): 24 uniform bits from two rounds of an integer hash keyed by (seed, step counter, salt); oracle/kernel_spec.py mirrors it
__device__ __forceinline__ bool drop_keep(int64_t i, uint32_t key, float p) {
  const uint32_t h = mix32(mix32((uint32_t)i ^ key) + (uint32_t)((uint64_t)i >> 32) + (key << 7 | key >> 25));
  return (float)(h >> 8) * (1.f / 16777216.f) >= p;
}

__global__ void dropout_fwd_kernel(const float* x, float* y, uint8_t* __restrict__ mask, int64_t n, float p,
                                   const int64_t* __restrict__ rng, uint32_t salt) {
  const uint32_t key = (uint32_t)rng[0] ^ ((uint32_t)rng[1] * 0x85EBCA6Bu) ^ (salt * 0xC2B2AE35u);
  const float inv = 1.f / (1.f - p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool k = drop_keep(i, key, p);
    mask[i] = k;
    y[i] = k ? x[i] * inv : 0.f;
  }
}

__global__ void dropout_bwd_kernel(const float* g, float* out, const uint8_t* __restrict__ mask,
                                   const float* __restrict__ relu_ref, int64_t n, float inv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i];
    if (mask) v = mask[i] ? v * inv : 0.f;
    if (relu_ref && !(relu_ref[i] > 0.f)) v = 0.f;
    out[i] = v;
  }
}

__global__ void rng_advance_kernel(int64_t* rng) { rng[1] += 1; }

// ------------------------------------------------------------------ residual add + LayerNorm
constexpr int LN_MAXC = 16;    // columns per lane: n <= 512
constexpr int LN_ROWS = 32;    // rows per block in the backward (4 warps x 8 rows): one atomic per column and block

__global__ void __launch_bounds__(128) addln_fwd_kernel(const vinet_addln_t d) {
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= d.rows) return;
  const float* x = d.x + row * d.n;
  const float* y = d.y + row * d.n;
  float v[LN_MAXC];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c) {
    const int j = lane + 32 * c;
    v[c] = j < d.n ? x[j] + y[j] : 0.f;
    sum += v[c];
  }
  const float mean = warp_sum(sum) / (float)d.n;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c) {
    const int j = lane + 32 * c;
    const float t = j < d.n ? v[c] - mean : 0.f;
    sq = fmaf(t, t, sq);
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)d.n + d.eps);
  if (lane == 0) {
    d.stat[row * 2] = mean;
    d.stat[row * 2 + 1] = rstd;
  }
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c) {
    const int j = lane + 32 * c;
    if (j < d.n) {
      d.z[row * d.n + j] = v[c];
      d.out[row * d.n + j] = fmaf((v[c] - mean) * rstd, __ldg(d.gamma + j), __ldg(d.beta + j));
    }
  }
}

__global__ void __launch_bounds__(128) addln_bwd_kernel(const vinet_addln_t d) {
  __shared__ float red[2][4][32 * LN_MAXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dg[LN_MAXC], db[LN_MAXC], gam[LN_MAXC];
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c) {
    const int j = lane + 32 * c;
    dg[c] = db[c] = 0.f;
    gam[c] = j < d.n ? __ldg(d.gamma + j) : 0.f;
  }
  const int64_t r0 = (int64_t)blockIdx.x * LN_ROWS;
  for (int rr = warp; rr < LN_ROWS; rr += 4) {
    const int64_t row = r0 + rr;
    if (row >= d.rows) break;
    const float mean = d.stat[row * 2], rstd = d.stat[row * 2 + 1];
    float xh[LN_MAXC], gy[LN_MAXC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int j = lane + 32 * c;
      const float g = j < d.n ? d.gout[row * d.n + j] : 0.f;
      xh[c] = j < d.n ? (d.z[row * d.n + j] - mean) * rstd : 0.f;
      gy[c] = g * gam[c];
      s1 += gy[c];
      s2 = fmaf(gy[c], xh[c], s2);
      dg[c] = fmaf(g, xh[c], dg[c]);
      db[c] += g;
    }
    const float c1 = warp_sum(s1) / (float)d.n, c2 = warp_sum(s2) / (float)d.n;
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int j = lane + 32 * c;
      if (j < d.n) d.dz[row * d.n + j] = rstd * (gy[c] - c1 - xh[c] * c2);
    }
  }
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c) {
    red[0][warp][lane + 32 * c] = dg[c];
    red[1][warp][lane + 32 * c] = db[c];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d.n; j += 128) {
    atomicAdd(d.dgamma + j, (red[0][0][j] + red[0][1][j]) + (red[0][2][j] + red[0][3][j]));
    atomicAdd(d.dbeta + j, (red[1][0][j] + red[1][1][j]) + (red[1][2][j] + red[1][3][j]));
  }
}

static unsigned ew_grid(int64_t n) { return (unsigned)std::min<int64_t>(cdiv(n, 256), 148 * 8); }

}  // namespace vinet

using namespace vinet;

extern "C" int vinet_bgemm(const vinet_bgemm_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->M >= 1 && d->N >= 1 && d->K >= 0 && d->nb1 >= 1 && d->nb2 >= 1, "bgemm: bad shape %d x %d x %d (%d, %d batches)", d->M, d->N, d->K,
              d->nb1, d->nb2);
  VINET_CHECK(d->C && (d->K == 0 || (d->A && d->B)), "bgemm: null operand");
  VINET_CHECK(!((d->accumulate & 1) && d->c_dtype != VINET_F32), "bgemm: accumulation needs an fp32 C");
  VINET_CHECK(!d->a_scale == !d->a_shift, "bgemm: a_scale and a_shift come together");
  VINET_CHECK((int64_t)d->nb1 * d->nb2 <= 65535, "bgemm: too many batches");
  const int64_t tiles = cdiv(d->N, BG_T) * cdiv(d->M, BG_T);
  // split K when one wave of CTAs is far from full and the reduction is long - only where the caller allows an order-free
  // reduction (accumulate bit 1: gradients): fp32 C without ReLU, a single batch, and (unless the call accumulates anyway) a DENSE
  // C block that one memset can clear
  const bool dense = (d->sCn == 1 && d->sCm == d->N) || (d->sCm == 1 && d->sCn == d->M) || (d->N == 1 && d->sCm == 1) || (d->M == 1 && d->sCn == 1);
  int splits = 1, len = d->K;
  if ((d->accumulate & 2) && d->nb1 * d->nb2 == 1 && d->c_dtype == VINET_F32 && !d->relu && d->K >= 8 * BG_K && tiles <= 74 &&
      ((d->accumulate & 1) || dense)) {
    splits = (int)std::min<int64_t>(cdiv(2 * 148, tiles), d->K / (4 * BG_K));
    len = (int)round_up(cdiv(d->K, splits), BG_K);
    splits = (int)cdiv(d->K, len);
  }
  if (splits > 1 && !(d->accumulate & 1)) cudaMemsetAsync(d->C, 0, (size_t)d->M * d->N * sizeof(float), (cudaStream_t)stream);
  dim3 grid((unsigned)cdiv(d->N, BG_T), (unsigned)cdiv(d->M, BG_T), (unsigned)(splits > 1 ? splits : d->nb1 * d->nb2));
  bgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d, len, splits);
  VINET_LAUNCH_OK("bgemm");
  return 0;
}

extern "C" int vinet_softmax_fwd(float* s, int64_t rows, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(rows >= 1 && n >= 1, "softmax: bad shape");
  softmax_fwd_kernel<<<(unsigned)cdiv(rows, 4), 128, 0, (cudaStream_t)stream>>>(s, rows, n);
  VINET_LAUNCH_OK("softmax_fwd");
  return 0;
}

extern "C" int vinet_softmax_bwd(const float* p, float* dp, int64_t rows, int32_t n, vinet_stream_t stream) {
  VINET_CHECK(rows >= 1 && n >= 1, "softmax: bad shape");
  softmax_bwd_kernel<<<(unsigned)cdiv(rows, 4), 128, 0, (cudaStream_t)stream>>>(p, dp, rows, n);
  VINET_LAUNCH_OK("softmax_bwd");
  return 0;
}

extern "C" int vinet_dropout_fwd(const float* x, float* y, uint8_t* mask, int64_t n, float p, const int64_t* rng, uint32_t salt,
                                 vinet_stream_t stream) {
  VINET_CHECK(n >= 1 && p >= 0.f && p < 1.f && mask && rng, "dropout_fwd: bad arguments");
  dropout_fwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, mask, n, p, rng, salt);
  VINET_LAUNCH_OK("dropout_fwd");
  return 0;
}

extern "C" int vinet_dropout_bwd(const float* g, float* out, const uint8_t* mask, const float* relu_ref, int64_t n, float p,
                                 vinet_stream_t stream) {
  VINET_CHECK(n >= 1 && p >= 0.f && p < 1.f, "dropout_bwd: bad arguments");
  dropout_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(g, out, mask, relu_ref, n, 1.f / (1.f - p));
  VINET_LAUNCH_OK("dropout_bwd");
  return 0;
}

extern "C" int vinet_rng_advance(int64_t* rng, vinet_stream_t stream) {
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng);
  VINET_LAUNCH_OK("rng_advance");
  return 0;
}

extern "C" int vinet_add_layernorm_fwd(const vinet_addln_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->n >= 1 && d->n <= 32 * LN_MAXC && d->rows >= 1, "add_layernorm: n %d (max %d)", d->n, 32 * LN_MAXC);
  addln_fwd_kernel<<<(unsigned)cdiv(d->rows, 4), 128, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("add_layernorm_fwd");
  return 0;
}

extern "C" int vinet_add_layernorm_bwd(const vinet_addln_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->n >= 1 && d->n <= 32 * LN_MAXC && d->rows >= 1, "add_layernorm: n %d (max %d)", d->n, 32 * LN_MAXC);
  addln_bwd_kernel<<<(unsigned)cdiv(d->rows, LN_ROWS), 128, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("add_layernorm_bwd");
  return 0;
}
