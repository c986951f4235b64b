// fp32 FFMA implicit-GEMM convolution (engine VINET_ENGINE_SIMT): the parity-mode path and the
// on-device cross-check for the tcgen05 kernels.  Same descriptors, same gather, fp32 math.
//   conv_gemm : out[rows,N] (+)= act(gather(rows,K) x W[K,N])      rows = B*Tr*Hr*Wr
//   conv_wgrad: dwp[(tap,c),n] += sum_rows gather(row,(tap,c)) * dy[row,n]
#include "gather.cuh"

namespace vinet {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16;

template <typename T, typename TO>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(const __grid_constant__ vinet_conv_t d, int npad) {
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ __align__(16) float Bs[SM_BK][SM_BN];
  const int tid = threadIdx.x;
  const int64_t M = gather_rows(d.g);
  const int64_t m0 = (int64_t)blockIdx.x * SM_BM;
  const int n0 = blockIdx.y * SM_BN;
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const RowCoord rc = decode_row(d.g, m0 + a_row, M);
  const float* __restrict__ W = reinterpret_cast<const float*>(d.w);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int K = d.k_blocks * VINET_TC_BLOCK_K;
  for (int k0 = 0; k0 < K; k0 += SM_BK) {
    float av[4];
    gather_vec<T, 4>(d.g, rc, k0 + a_k, av);
    const float4 bv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + b_k) * npad + n0 + b_n));
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = av[i];
    *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SM_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= M) continue;
    const RowCoord orc = decode_row(d.g, row, M);
    TO* p = out_row_ptr<TO>(d, orc);
    const bool accum = (d.accumulate >> out_index(d, orc)) & 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.N) continue;
      float v = epilogue_value(d, acc[i][j], n);
      if (accum) v += load1(p + n);
      store1(p + n, v);
    }
  }
}

template <typename T, typename TD>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const __grid_constant__ vinet_wgrad_t d) {
  __shared__ __align__(16) float As[SM_BK][SM_BM];
  __shared__ __align__(16) float Bs[SM_BK][SM_BN];
  const int tid = threadIdx.x;
  const int64_t M = gather_rows(d.g);
  const int m0 = blockIdx.x * SM_BM;
  const int n0 = blockIdx.y * SM_BN;
  const int64_t chunk = cdiv(cdiv(M, SM_BK), d.splits) * SM_BK;
  const int64_t r_begin = (int64_t)blockIdx.z * chunk;
  const int64_t r_end = min(M, r_begin + chunk);
  const int p = tid >> 4, q = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  const TD* __restrict__ dy = reinterpret_cast<const TD*>(d.dy);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += SM_BK) {
    const int64_t row = r0 + p;
    const RowCoord rc = decode_row(d.g, row, r_end);
    float av[4], bv[4];
    gather_vec<T, 4>(d.g, rc, m0 + q, av);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + q + j;
      bv[j] = (row < r_end && n < d.N) ? load1(dy + row * d.lddy + n) : 0.f;
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[p][q]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[p][q]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SM_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const int mw = d.g.ntaps * d.g.Cs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= mw) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d.N) atomicAdd(d.dwp + (int64_t)m * d.lddw + n, acc[i][j]);
    }
  }
}

int conv_gemm_simt(const vinet_conv_t* d, cudaStream_t stream) {
  const int64_t M = (int64_t)d->g.B * d->g.Tr * d->g.Hr * d->g.Wr;
  const int npad = (int)round_up(d->N, SM_BN);
  dim3 grid((unsigned)cdiv(M, SM_BM), (unsigned)(npad / SM_BN));
  VINET_DISPATCH_DTYPE(d->g.dtype, T, VINET_DISPATCH_DTYPE(d->out_dtype, TO,
      (conv_gemm_simt_kernel<T, TO><<<grid, 256, 0, stream>>>(*d, npad))));
  note_kernel("conv_gemm_simt_kernel");
  VINET_LAUNCH_OK("conv_gemm_simt");
  return 0;
}

int conv_wgrad_simt(const vinet_wgrad_t* d, cudaStream_t stream) {
  const int mw = d->g.ntaps * d->g.Cs;
  dim3 grid((unsigned)cdiv(mw, SM_BM), (unsigned)cdiv(d->N, SM_BN), (unsigned)d->splits);
  VINET_DISPATCH_DTYPE(d->g.dtype, T, VINET_DISPATCH_DTYPE(d->dy_dtype, TD,
      (conv_wgrad_simt_kernel<T, TD><<<grid, 256, 0, stream>>>(*d))));
  note_kernel("conv_wgrad_simt_kernel");
  VINET_LAUNCH_OK("conv_wgrad_simt");
  return 0;
}

}  // namespace vinet
