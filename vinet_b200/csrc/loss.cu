// Saliency losses and their input gradients (loss.py:13-120): one thread-block CLUSTER of 8 CTAs per sample (a batch of 8 maps
// would otherwise keep 8 of 148 SMs busy through five dependent passes), warp-shuffle + shared-memory reductions inside a CTA and
// a fixed-order sum of the CTAs' partials over distributed shared memory (deterministic), mean over the batch by the last cluster.
//   kldiv      loss.py:13-38     cc   loss.py:80-99
//   similarity loss.py:53-78     nss  loss.py:101-120 (same-size branch)
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vinet {

constexpr int LOSS_THREADS = 512;
constexpr int LOSS_CLUSTER = 8;      // CTAs per sample (portable cluster size)
constexpr float LOSS_EPS = 2.2204e-16f;

// Shared-memory scratch of one CTA; the xch* slots are read by the other CTAs of the cluster (DSMEM).  Two slots alternate
// between consecutive reductions: a CTA can only reach the write of reduction k+2 after every CTA has passed the cluster barrier
// of reduction k+1, i.e. has finished reading slot k & 1.
struct LossSmem {
  double warp[LOSS_THREADS / 32];
  float warpv[LOSS_THREADS / 32];
  int warpi[LOSS_THREADS / 32];
  double xch[2];
  float xchv[2];
  int xchi[2];
  int turn;
};

// sum over all threads of the cluster, identical in every thread (fixed order: warps of a CTA, then CTAs by rank)
__device__ __forceinline__ double block_sum(double v, LossSmem& sm) {
  cg::cluster_group cl = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm.warp[warp] = v;
  __syncthreads();
  const int slot = sm.turn & 1;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < LOSS_THREADS / 32; ++i) s += sm.warp[i];
    sm.xch[slot] = s;
  }
  cl.sync();
  double tot = 0.0;
  for (unsigned r = 0; r < cl.num_blocks(); ++r) tot += *cl.map_shared_rank(&sm.xch[slot], r);
  __syncthreads();
  if (threadIdx.x == 0) sm.turn = slot ^ 1;
  return tot;
}

// minimum and the first index attaining it
__device__ __forceinline__ void block_argmin(float& v, int& idx, LossSmem& sm) {
  cg::cluster_group cl = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if (lane == 0) { sm.warpv[warp] = v; sm.warpi[warp] = idx; }
  __syncthreads();
  const int slot = sm.turn & 1;
  if (threadIdx.x == 0) {
    float bv = sm.warpv[0];
    int bi = sm.warpi[0];
    for (int i = 1; i < LOSS_THREADS / 32; ++i)
      if (sm.warpv[i] < bv || (sm.warpv[i] == bv && sm.warpi[i] < bi)) { bv = sm.warpv[i]; bi = sm.warpi[i]; }
    sm.xchv[slot] = bv; sm.xchi[slot] = bi;
  }
  cl.sync();
  v = *cl.map_shared_rank(&sm.xchv[slot], 0); idx = *cl.map_shared_rank(&sm.xchi[slot], 0);
  for (unsigned r = 1; r < cl.num_blocks(); ++r) {
    const float ov = *cl.map_shared_rank(&sm.xchv[slot], r);
    const int oi = *cl.map_shared_rank(&sm.xchi[slot], r);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) sm.turn = slot ^ 1;
}

__device__ __forceinline__ float block_max(float v, LossSmem& sm) {
  cg::cluster_group cl = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sm.warpv[warp] = v;
  __syncthreads();
  const int slot = sm.turn & 1;
  if (threadIdx.x == 0) {
    float m = sm.warpv[0];
    for (int i = 1; i < LOSS_THREADS / 32; ++i) m = fmaxf(m, sm.warpv[i]);
    sm.xchv[slot] = m;
  }
  cl.sync();
  v = *cl.map_shared_rank(&sm.xchv[slot], 0);
  for (unsigned r = 1; r < cl.num_blocks(); ++r) v = fmaxf(v, *cl.map_shared_rank(&sm.xchv[slot], r));
  __syncthreads();
  if (threadIdx.x == 0) sm.turn = slot ^ 1;
  return v;
}

// per_sample[b][0] = value; [1..7] = reductions reused by the backward
__global__ void __launch_bounds__(LOSS_THREADS) loss_fwd_kernel(const __grid_constant__ vinet_loss_t d) {
  __shared__ LossSmem sh;
  __shared__ bool is_last;
  cg::cluster_group cl = cg::this_cluster();
  const int CL = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const int b = blockIdx.x / CL, n = d.n;
  const int tid = rank * LOSS_THREADS + threadIdx.x, STEP = CL * LOSS_THREADS;   // element walk of the whole cluster
  if (threadIdx.x == 0) sh.turn = 0;
  __syncthreads();
  const float* __restrict__ s = d.s + (int64_t)b * n;
  const float* __restrict__ g = d.g + (int64_t)b * n;
  float* ps = d.per_sample + b * 8;
  float value = 0.f;
  if (d.kind == VINET_LOSS_KLDIV) {
    float ls = 0.f, lg = 0.f;
    for (int i = tid; i < n; i += STEP) { ls += s[i]; lg += g[i]; }
    const float S = (float)block_sum((double)ls, sh);
    const float G = (float)block_sum((double)lg, sh);
    float acc = 0.f;
    for (int i = tid; i < n; i += STEP) {
      const float sn = s[i] / S, gn = g[i] / G;
      acc += gn * logf(LOSS_EPS + gn / (sn + LOSS_EPS));
    }
    value = (float)block_sum((double)acc, sh);
    if (tid == 0) { ps[1] = S; ps[2] = G; }
  } else if (d.kind == VINET_LOSS_CC || d.kind == VINET_LOSS_NSS) {
    float ls = 0.f, lg = 0.f;
    for (int i = tid; i < n; i += STEP) { ls += s[i]; lg += g[i]; }
    const double sum_s = block_sum((double)ls, sh), sum_g = block_sum((double)lg, sh);
    const float ms = (float)(sum_s / n), mg = (float)(sum_g / n);
    float ss = 0.f, gg = 0.f, sg = 0.f, sf = 0.f;
    for (int i = tid; i < n; i += STEP) {
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
    for (int i = tid; i < n; i += STEP) {
      const float a = s[i], c = g[i];
      if (a < mns) { mns = a; is = i; }
      if (c < mng) { mng = c; ig = i; }
      mxs = fmaxf(mxs, a); mxg = fmaxf(mxg, c);
      ls += a; lg += c;
    }
    block_argmin(mns, is, sh);
    block_argmin(mng, ig, sh);
    mxs = block_max(mxs, sh);
    mxg = block_max(mxg, sh);
    // normalize_map then divide by the sum, as the reference does (loss.py:62-74)
    const float ds = mxs - mns, dg = mxg - mng;
    float l1 = 0.f, l2 = 0.f;
    for (int i = tid; i < n; i += STEP) { l1 += (s[i] - mns) / ds; l2 += (g[i] - mng) / dg; }
    const float Us = (float)block_sum((double)l1, sh), Ug = (float)block_sum((double)l2, sh);
    float acc = 0.f;
    for (int i = tid; i < n; i += STEP) acc += fminf(((s[i] - mns) / ds) / Us, ((g[i] - mng) / dg) / Ug);
    value = (float)block_sum((double)acc, sh);
    if (tid == 0) {
      ps[1] = mns; ps[2] = ds; ps[3] = Us; ps[4] = mng; ps[5] = dg; ps[6] = Ug; ps[7] = __int_as_float(is);
    }
  }
  if (tid == 0) {      // thread 0 of the cluster's rank-0 CTA
    ps[0] = value;
    __threadfence();
    const int done = atomicAdd(d.counter, 1);
    is_last = (done == d.B - 1);
  }
  __syncthreads();
  if (tid == 0 && is_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = 0; i < d.B; ++i) tot += reinterpret_cast<volatile float*>(d.per_sample)[i * 8];
    d.out[0] = tot / (float)d.B;
    *d.counter = 0;
  }
  cl.sync();           // no CTA may leave while others can still read its shared memory
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(const __grid_constant__ vinet_loss_t d) {
  __shared__ LossSmem sh;
  cg::cluster_group cl = cg::this_cluster();
  const int CL = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const int b = blockIdx.x / CL, n = d.n;
  const int tid = rank * LOSS_THREADS + threadIdx.x, STEP = CL * LOSS_THREADS;
  if (threadIdx.x == 0) sh.turn = 0;
  __syncthreads();
  const float* __restrict__ s = d.s + (int64_t)b * n;
  const float* __restrict__ g = d.g + (int64_t)b * n;
  float* __restrict__ gs = d.grad_s + (int64_t)b * n;
  const float* ps = d.per_sample + b * 8;
  const float up = d.gout[0] / (float)d.B;
  if (d.kind == VINET_LOSS_KLDIV) {
    const float S = ps[1], G = ps[2];
    // q_i = dL/ds_n,i ; dL/ds_i = (q_i - sum_j q_j s_n,j) / S
    float acc = 0.f;
    for (int i = tid; i < n; i += STEP) {
      const float sn = s[i] / S, gn = g[i] / G;
      const float den = sn + LOSS_EPS;
      const float q = -gn * gn / ((LOSS_EPS + gn / den) * den * den);
      acc = fmaf(q, sn, acc);
    }
    const float Q = (float)block_sum((double)acc, sh);
    for (int i = tid; i < n; i += STEP) {
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
    for (int i = tid; i < n; i += STEP) gs[i] = up * ((g[i] - mg) - k * (s[i] - ms)) * inv;
  } else if (d.kind == VINET_LOSS_NSS) {
    const float ms = ps[1], sigma = ps[2], P = ps[3], F = ps[4];
    const float se = sigma + LOSS_EPS;
    const float k1 = 1.f / (se * F), k2 = P / (F * se * se * (float)(n - 1) * sigma);
    const float fbar = F / (float)n;
    for (int i = tid; i < n; i += STEP) gs[i] = up * ((g[i] - fbar) * k1 - k2 * (s[i] - ms));
  } else {
    const float mns = ps[1], ds = ps[2], Us = ps[3], mng = ps[4], dg = ps[5], Ug = ps[6];
    const int istar = __float_as_int(ps[7]);
    // s''_i = u_i / U with u_i = s_i - min, U = sum_j u_j  (the max cancels)
    const float U = Us * ds;
    float a1 = 0.f, a2 = 0.f;
    for (int i = tid; i < n; i += STEP) {
      const float sv = ((s[i] - mns) / ds) / Us, gv = ((g[i] - mng) / dg) / Ug;
      if (sv < gv) { a1 += 1.f; a2 += (s[i] - mns); }
    }
    const float A = (float)block_sum((double)a1, sh), Bq = (float)block_sum((double)a2, sh);
    for (int i = tid; i < n; i += STEP) {
      const float sv = ((s[i] - mns) / ds) / Us, gv = ((g[i] - mng) / dg) / Ug;
      float v = ((sv < gv) ? 1.f / U : 0.f) - Bq / (U * U);
      if (i == istar) v += -A / U + (float)n * Bq / (U * U);
      gs[i] = up * v;
    }
  }
  cl.sync();           // no CTA may leave while others can still read its shared memory
}

}  // namespace vinet
using namespace vinet;

// one cluster of LOSS_CLUSTER CTAs per sample (fewer for tiny maps: every CTA should have a few elements per thread)
static int launch_clustered(void (*kern)(const vinet_loss_t), const vinet_loss_t* d, cudaStream_t stream) {
  int cl = LOSS_CLUSTER;
  while (cl > 1 && (int64_t)d->n < (int64_t)cl * LOSS_THREADS * 2) cl >>= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(d->B * cl));
  cfg.blockDim = dim3(LOSS_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *d);
  if (e != cudaSuccess) {
    set_error("loss: cluster launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

extern "C" int vinet_loss_fwd(const vinet_loss_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->kind >= 0 && d->kind <= 3 && d->B >= 1 && d->n >= 2, "loss: bad arguments");
  if (launch_clustered(loss_fwd_kernel, d, (cudaStream_t)stream)) return -2;
  VINET_LAUNCH_OK("loss_fwd");
  return 0;
}

extern "C" int vinet_loss_bwd(const vinet_loss_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->kind >= 0 && d->kind <= 3 && d->B >= 1 && d->n >= 2, "loss: bad arguments");
  if (launch_clustered(loss_bwd_kernel, d, (cudaStream_t)stream)) return -2;
  VINET_LAUNCH_OK("loss_bwd");
  return 0;
}
