// AViNet audio branch: SoundNet 1-D convolutions (model.py:750-786, nn.Conv2d with (k,1) kernels on a
// (B,1,L,1) waveform), BatchNorm2d+ReLU+MaxPool, and the audio-visual bilinear fusion (model.py:229-237).
// 0.19 GFLOP per clip: warp-per-output kernels with shuffle reductions, fp32 throughout.
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace vinet {

// ------------------------------------------------------------------ conv1d as tiled fp32 GEMMs
// The SoundNet convolutions are small GEMMs (0.19 GFLOP per clip in total) with awkward shapes: 141k x 64 x 16 for the first layer,
// 12 x 2048 x 1024 for the last.  One 64 x 64 x 16 tiled FFMA kernel serves forward, data gradient and weight gradient through
// three operand loaders; the reduction dimension is split across CTAs (fp32 atomics into a zeroed output) whenever the tile
// grid alone would leave most of the 148 SMs idle.  (Round 1 ran one warp per output element: 7.5 ms per step for AViNet.)
enum { A1D_FWD = 0, A1D_DGRAD = 1, A1D_WGRAD = 2 };
constexpr int A1D_BM = 64, A1D_BN = 64, A1D_BK = 16;

template <int MODE>
struct A1dOps {
  const vinet_conv1d_t& d;
  __device__ __forceinline__ A1dOps(const vinet_conv1d_t& d_) : d(d_) {}
  __device__ __forceinline__ int M() const { return MODE == A1D_FWD ? d.B * d.Lout : (MODE == A1D_DGRAD ? d.B * d.Lin : d.Cout); }
  __device__ __forceinline__ int N() const { return MODE == A1D_FWD ? d.Cout : (MODE == A1D_DGRAD ? d.Cin : d.Cin * d.k); }
  __device__ __forceinline__ int K() const { return MODE == A1D_FWD ? d.Cin * d.k : (MODE == A1D_DGRAD ? d.Cout * d.k : d.B * d.Lout); }
  __device__ __forceinline__ float a(int m, int kk) const {
    if (MODE == A1D_FWD) {            // x[b, ci, l*stride - pad + t]
      const int b = m / d.Lout, l = m - b * d.Lout;
      const int ci = kk / d.k, t = kk - ci * d.k;
      const int i = l * d.stride - d.pad + t;
      return (unsigned)i < (unsigned)d.Lin ? __ldg(d.x + ((int64_t)b * d.Cin + ci) * d.Lin + i) : 0.f;
    } else if (MODE == A1D_DGRAD) {   // dy[b, co, (i + pad - t) / stride]
      const int b = m / d.Lin, i = m - b * d.Lin;
      const int co = kk / d.k, t = kk - co * d.k;
      const int num = i + d.pad - t;
      if (num < 0 || num % d.stride) return 0.f;
      const int l = num / d.stride;
      return l < d.Lout ? __ldg(d.dy + ((int64_t)b * d.Cout + co) * d.Lout + l) : 0.f;
    } else {                          // dy[b, co = m, l]
      const int b = kk / d.Lout, l = kk - b * d.Lout;
      return __ldg(d.dy + ((int64_t)b * d.Cout + m) * d.Lout + l);
    }
  }
  __device__ __forceinline__ float b(int kk, int n) const {
    if (MODE == A1D_FWD) return __ldg(d.w + (int64_t)n * (d.Cin * d.k) + kk);      // w[co = n][ci*k + t]
    if (MODE == A1D_DGRAD) {                                                        // w[co][ci = n][t]
      const int co = kk / d.k, t = kk - co * d.k;
      return __ldg(d.w + ((int64_t)co * d.Cin + n) * d.k + t);
    }
    const int bb = kk / d.Lout, l = kk - bb * d.Lout;                               // x[b, ci, l*stride - pad + t], n = ci*k + t
    const int ci = n / d.k, t = n - ci * d.k;
    const int i = l * d.stride - d.pad + t;
    return (unsigned)i < (unsigned)d.Lin ? __ldg(d.x + ((int64_t)bb * d.Cin + ci) * d.Lin + i) : 0.f;
  }
  __device__ __forceinline__ float* out(int m, int n) const {
    if (MODE == A1D_FWD) { const int b = m / d.Lout, l = m - b * d.Lout; return d.y + ((int64_t)b * d.Cout + n) * d.Lout + l; }
    if (MODE == A1D_DGRAD) { const int b = m / d.Lin, i = m - b * d.Lin; return d.dx + ((int64_t)b * d.Cin + n) * d.Lin + i; }
    return d.dw + (int64_t)m * (d.Cin * d.k) + n;
  }
};

// grid = (tiles over N, tiles over M, K splits); 256 threads, each a 4 x 4 micro-tile.  MODE FWD/DGRAD map M to long contiguous
// runs in memory, so the m index of the A tile is the fast thread index there; WGRAD reads dy / x along K.
// Split-K partial sums are combined by atomics (backward: order-dependent in the last bit, like every other gradient reduction
// here) or, for the FORWARD pass, in a fixed order: the CTAs of one output tile take turns by split index (a per-tile turn
// counter; CTAs are dispatched in block-id order, so a waiting CTA only ever waits for CTAs dispatched before it), split 0 stores,
// the others add, the last one resets the counter.  The forward output is then bit-reproducible run to run, which the bf16
// network behind it turns from a 1e-7 into a 1e-2 matter (tools/av_determinism.py).
constexpr int A1D_TURN_SLOTS = 64, A1D_TURN_TILES = 512;
__device__ int g_a1d_turn[A1D_TURN_SLOTS * A1D_TURN_TILES];

template <int MODE>
__global__ void __launch_bounds__(256) audio_gemm_kernel(const __grid_constant__ vinet_conv1d_t d, int ksplit_len, int atomic, int* turn) {
  __shared__ float As[A1D_BK][A1D_BM + 4], Bs[A1D_BK][A1D_BN + 4];
  const A1dOps<MODE> op(d);
  const int M = op.M(), N = op.N(), K = op.K();
  const int m0 = blockIdx.y * A1D_BM, n0 = blockIdx.x * A1D_BN;
  const int k_begin = blockIdx.z * ksplit_len, k_end = min(K, k_begin + ksplit_len);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += A1D_BK) {
    // 64 x 16 elements of each operand, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (MODE == A1D_WGRAD) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < k_end) ? op.a(m, k) : 0.f;
      int nn, kb;
      if (MODE == A1D_FWD) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
      const int n = n0 + nn, k2 = k0 + kb;
      Bs[kb][nn] = (n < N && k2 < k_end) ? op.b(k2, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < A1D_BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool ordered = turn != nullptr;
  int* my_turn = ordered ? turn + blockIdx.y * gridDim.x + blockIdx.x : nullptr;
  if (ordered && blockIdx.z > 0) {
    if (tid == 0) {
      while (atomicAdd(my_turn, 0) != (int)blockIdx.z) __nanosleep(32);
      __threadfence();
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (MODE == A1D_FWD && d.bias && blockIdx.z == 0) v += __ldg(d.bias + n);
      if (ordered) {
        float* o = op.out(m, n);
        __stcg(o, blockIdx.z == 0 ? v : __ldcg(o) + v);      // L2 is the point of coherence between the CTAs of a tile
      } else if (atomic) {
        atomicAdd(op.out(m, n), v);
      } else {
        *op.out(m, n) = v;
      }
    }
  }
  if (ordered) {
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(my_turn, blockIdx.z + 1 == gridDim.z ? 0 : (int)blockIdx.z + 1);
  }
}

// dbias[co] = sum over (b, l) of dy: one block per output channel
__global__ void __launch_bounds__(256) conv1d_dbias_kernel(const __grid_constant__ vinet_conv1d_t d) {
  __shared__ float sh[8];
  const int co = blockIdx.x;
  float acc = 0.f;
  for (int64_t r = threadIdx.x; r < (int64_t)d.B * d.Lout; r += 256) {
    const int b = (int)(r / d.Lout), l = (int)(r - (int64_t)b * d.Lout);
    acc += d.dy[((int64_t)b * d.Cout + co) * d.Lout + l];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sh[i];
    d.dbias[co] = s;
  }
}

template <int MODE>
static int audio_gemm_launch(const vinet_conv1d_t* d, int M, int N, int K, float* out, size_t out_elems, cudaStream_t stream, const char* what) {
  const int tm = (int)cdiv(M, A1D_BM), tn = (int)cdiv(N, A1D_BN);
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(2 * 148, (int64_t)tm * tn), cdiv(K, 4 * A1D_BK)));
  // forward: ordered (bit-reproducible) combination of the K splits; VINET_AUDIO_FWD_ATOMIC=1 restores the atomics for A/B runs
  static const bool fwd_atomic = getenv("VINET_AUDIO_FWD_ATOMIC") != nullptr;
  static std::atomic<unsigned> next_slot{0};
  int len = (int)round_up(cdiv(K, splits), A1D_BK);
  splits = (int)cdiv(K, len);
  int* turn = nullptr;
  if (MODE == A1D_FWD && splits > 1 && !fwd_atomic && (int64_t)tm * tn <= A1D_TURN_TILES) {
    static int* bases[64] = {nullptr};            // per device; resolved once, outside any stream capture (warm-up steps)
    int dev = 0;
    cudaGetDevice(&dev);
    int* base = dev < 64 ? bases[dev] : nullptr;
    if (!base) {
      if (cudaGetSymbolAddress((void**)&base, g_a1d_turn) != cudaSuccess) {
        set_error("%s: cudaGetSymbolAddress failed", what);
        return -1;
      }
      if (dev < 64) bases[dev] = base;
    }
    turn = base + (next_slot.fetch_add(1, std::memory_order_relaxed) % A1D_TURN_SLOTS) * A1D_TURN_TILES;
  }
  if (splits > 1 && !turn) cudaMemsetAsync(out, 0, out_elems * sizeof(float), stream);
  audio_gemm_kernel<MODE><<<dim3((unsigned)tn, (unsigned)tm, (unsigned)splits), 256, 0, stream>>>(*d, len, splits > 1 ? 1 : 0, turn);
  VINET_LAUNCH_OK(what);
  return 0;
}

// ------------------------------------------------------------------ BatchNorm2d + ReLU + MaxPool((p,1))
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < nw; ++i) s += sh[i];
  return s;
}

constexpr int BN1D_T = 1024;

// block per channel (1024 threads: the first layers have 16 / 32 channels and 10^5 elements each)
__global__ void __launch_bounds__(BN1D_T) bn1d_fwd_kernel(const __grid_constant__ vinet_bn1d_t d) {
  __shared__ double sh[32];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int64_t N = (int64_t)d.B * d.L;
  float mean, invstd;
  if (d.training) {
    double s = 0.0, ss = 0.0;
    for (int64_t r = tid; r < N; r += BN1D_T) {
      const int b = (int)(r / d.L), l = (int)(r - (int64_t)b * d.L);
      const double v = d.y[((int64_t)b * d.C + c) * d.L + l];
      s += v; ss += v * v;
    }
    s = block_sum(s, sh); ss = block_sum(ss, sh);
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
  for (int64_t r = tid; r < (int64_t)d.B * Lo; r += BN1D_T) {
    const int b = (int)(r / Lo), lo = (int)(r - (int64_t)b * Lo);
    const float* y = d.y + ((int64_t)b * d.C + c) * d.L + (int64_t)lo * d.pool;
    float best = -INFINITY;
    for (int p = 0; p < d.pool; ++p) best = fmaxf(best, fmaxf(fmaf(y[p], sc, shf), 0.f));
    d.out[((int64_t)b * d.C + c) * Lo + lo] = best;
  }
}

// Backward walks pooling WINDOWS: one scan finds the (first) maximum, which alone receives the window's gradient.
__global__ void __launch_bounds__(BN1D_T) bn1d_bwd_kernel(const __grid_constant__ vinet_bn1d_t d) {
  __shared__ double sh[32];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int64_t N = (int64_t)d.B * d.L;
  const float mean = d.mean[c], invstd = d.invstd[c];
  const float sc = d.gamma[c] * invstd, shf = d.beta[c] - mean * sc;
  const int Lo = d.L / d.pool;
  const int64_t nwin = (int64_t)d.B * Lo;
  // argmax of window r and its masked gradient
  auto window = [&](int64_t r, int& arg, float& g) {
    const int b = (int)(r / Lo), lo = (int)(r - (int64_t)b * Lo);
    const float* y = d.y + ((int64_t)b * d.C + c) * d.L + (int64_t)lo * d.pool;
    float best = -INFINITY;
    arg = 0;
    for (int p = 0; p < d.pool; ++p) {
      const float v = fmaxf(fmaf(y[p], sc, shf), 0.f);
      if (v > best) { best = v; arg = p; }
    }
    g = (fmaf(y[arg], sc, shf) > 0.f) ? d.gout[((int64_t)b * d.C + c) * Lo + lo] : 0.f;
  };
  double s1 = 0.0, s2 = 0.0;
  for (int64_t r = tid; r < nwin; r += BN1D_T) {
    int arg;
    float g;
    window(r, arg, g);
    const int b = (int)(r / Lo), lo = (int)(r - (int64_t)b * Lo);
    const float yn = (d.y[((int64_t)b * d.C + c) * d.L + (int64_t)lo * d.pool + arg] - mean) * invstd;
    s1 += g; s2 += (double)g * yn;
  }
  s1 = block_sum(s1, sh); s2 = block_sum(s2, sh);
  if (tid == 0) { d.dbeta[c] = (float)s1; d.dgamma[c] = (float)s2; }
  const float k1 = d.training ? (float)(s1 / (double)N) : 0.f, k2 = d.training ? (float)(s2 / (double)N) : 0.f;
  for (int64_t r = tid; r < nwin; r += BN1D_T) {
    int arg;
    float g;
    window(r, arg, g);
    const int b = (int)(r / Lo), lo = (int)(r - (int64_t)b * Lo);
    const int64_t base = ((int64_t)b * d.C + c) * d.L + (int64_t)lo * d.pool;
    for (int p = 0; p < d.pool; ++p) {
      const float yn = (d.y[base + p] - mean) * invstd;
      d.dy[base + p] = sc * ((p == arg ? g : 0.f) - k1 - yn * k2);
    }
  }
  // elements past the last full window (L % pool) receive no gradient through the pool
  const int tail = d.L - Lo * d.pool;
  for (int64_t r = tid; r < (int64_t)d.B * tail; r += BN1D_T) {
    const int b = (int)(r / tail), l = Lo * d.pool + (int)(r - (int64_t)b * tail);
    const int64_t at = ((int64_t)b * d.C + c) * d.L + l;
    const float yn = (d.y[at] - mean) * invstd;
    d.dy[at] = sc * (0.f - k1 - yn * k2);
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
  VINET_CHECK(d->B >= 1 && d->Cin >= 1 && d->Cout >= 1 && d->k >= 1 && d->stride >= 1 && d->Lout >= 1, "conv1d_fwd: bad shape");
  VINET_CHECK((int64_t)d->B * std::max(d->Lin, d->Lout) * std::max(d->Cin, d->Cout) < (int64_t)0x7fffffff, "conv1d: too large");
  return audio_gemm_launch<A1D_FWD>(d, d->B * d->Lout, d->Cout, d->Cin * d->k, d->y, (size_t)d->B * d->Cout * d->Lout, (cudaStream_t)stream,
                                    "conv1d_fwd");
}

extern "C" int vinet_conv1d_bwd(const vinet_conv1d_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->B >= 1 && d->Cin >= 1 && d->Cout >= 1 && d->k >= 1 && d->stride >= 1 && d->Lout >= 1, "conv1d_bwd: bad shape");
  if (d->dx) {
    if (audio_gemm_launch<A1D_DGRAD>(d, d->B * d->Lin, d->Cin, d->Cout * d->k, d->dx, (size_t)d->B * d->Cin * d->Lin, (cudaStream_t)stream,
                                     "conv1d_dgrad")) return -1;
  }
  if (audio_gemm_launch<A1D_WGRAD>(d, d->Cout, d->Cin * d->k, d->B * d->Lout, d->dw, (size_t)d->Cout * d->Cin * d->k, (cudaStream_t)stream,
                                   "conv1d_wgrad")) return -1;
  if (d->dbias) {
    conv1d_dbias_kernel<<<(unsigned)d->Cout, 256, 0, (cudaStream_t)stream>>>(*d);
    VINET_LAUNCH_OK("conv1d_dbias");
  }
  return 0;
}

extern "C" int vinet_bn1d_fwd(const vinet_bn1d_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->pool >= 1, "bn1d: pool");
  bn1d_fwd_kernel<<<d->C, BN1D_T, 0, (cudaStream_t)stream>>>(*d);
  VINET_LAUNCH_OK("bn1d_fwd");
  return 0;
}

extern "C" int vinet_bn1d_bwd(const vinet_bn1d_t* d, vinet_stream_t stream) {
  bn1d_bwd_kernel<<<d->C, BN1D_T, 0, (cudaStream_t)stream>>>(*d);
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
