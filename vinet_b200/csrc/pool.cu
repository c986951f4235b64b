// nn.MaxPool3d on NDHWC views with the producer's pending transform applied on read
// (model.py:696-714 stage pools, model_utils.py:178 3x3x3 stride-1 branch pools, model.py:229 AV pool).
// Backward recomputes the arg-max (first maximum in (t,h,w) scan order, like ATen) instead of storing indices.
#include "common.cuh"

namespace vinet {

// TO: forward output type / backward gout type; TGI: backward gin type
template <typename T, typename TO, typename TGI, bool BWD>
__global__ void maxpool_kernel(const __grid_constant__ vinet_pool_t d) {
  const T* __restrict__ x = reinterpret_cast<const T*>(d.x);
  const int G = d.C / 8;
  const int64_t total = (int64_t)d.B * d.To * d.Ho * d.Wo * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c = (int)(r % G) * 8; r /= G;
    const int wo = (int)(r % d.Wo); r /= d.Wo;
    const int ho = (int)(r % d.Ho); r /= d.Ho;
    const int to = (int)(r % d.To);
    const int b = (int)(r / d.To);
    float best[8];
    int64_t arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = -1; }
    for (int dt = 0; dt < d.kt; ++dt) {
      const int t = to * d.st - d.pt + dt;
      if ((unsigned)t >= (unsigned)d.Ti) continue;
      for (int dh = 0; dh < d.kh; ++dh) {
        const int h = ho * d.sh - d.ph + dh;
        if ((unsigned)h >= (unsigned)d.Hi) continue;
        for (int dw = 0; dw < d.kw; ++dw) {
          const int w = wo * d.sw - d.pw + dw;
          if ((unsigned)w >= (unsigned)d.Wi) continue;
          const int64_t pos = (((int64_t)b * d.Ti + t) * d.Hi + h) * d.Wi + w;
          float v[8];
          load8(x + pos * d.ldx + c, v);
          apply_xform<8>(v, d.xform, d.scale, d.shift, c);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (v[e] > best[e]) { best[e] = v[e]; arg[e] = pos; }
        }
      }
    }
    const int64_t opos = (((int64_t)b * d.To + to) * d.Ho + ho) * d.Wo + wo;
    if constexpr (!BWD) {
      store8(reinterpret_cast<TO*>(d.out) + opos * d.ldo + c, best);
    } else {
      float g[8];
      load8(reinterpret_cast<const TO*>(d.gout) + opos * d.ldgo + c, g);
      TGI* gin = reinterpret_cast<TGI*>(d.gin);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (arg[e] < 0) continue;
        if constexpr (sizeof(TGI) == 4) {
          atomicAdd(gin + arg[e] * d.ldgi + c + e, g[e]);
        } else {
          // bf16 gradients: native red.add.bf16x2 on the aligned pair holding channel c+e
          const __nv_bfloat16 z = __float2bfloat16_rn(0.f), v = __float2bfloat16_rn(g[e]);
          __nv_bfloat162 pair = (e & 1) ? __halves2bfloat162(z, v) : __halves2bfloat162(v, z);
          atomicAdd(reinterpret_cast<__nv_bfloat162*>(gin + arg[e] * d.ldgi + c + (e & ~1)), pair);
        }
      }
    }
  }
}

}  // namespace vinet
using namespace vinet;

static unsigned pool_grid(const vinet_pool_t* d) {
  int64_t total = (int64_t)d->B * d->To * d->Ho * d->Wo * (d->C / 8);
  int64_t nb = cdiv(total, 256);
  if (nb > 148 * 32) nb = 148 * 32;
  return (unsigned)(nb < 1 ? 1 : nb);
}

extern "C" int vinet_maxpool_fwd(const vinet_pool_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0, "maxpool: C %d", d->C);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->out_dtype, TO,
      (maxpool_kernel<T, TO, float, false><<<pool_grid(d), 256, 0, (cudaStream_t)stream>>>(*d))));
  VINET_LAUNCH_OK("maxpool_fwd");
  return 0;
}

extern "C" int vinet_maxpool_bwd(const vinet_pool_t* d, vinet_stream_t stream) {
  VINET_CHECK(d->C % 8 == 0, "maxpool: C %d", d->C);
  VINET_DISPATCH_DTYPE(d->dtype, T, VINET_DISPATCH_DTYPE(d->gout_dtype, TGO, VINET_DISPATCH_DTYPE(d->gin_dtype, TGI,
      (maxpool_kernel<T, TGO, TGI, true><<<pool_grid(d), 256, 0, (cudaStream_t)stream>>>(*d)))));
  VINET_LAUNCH_OK("maxpool_bwd");
  return 0;
}
