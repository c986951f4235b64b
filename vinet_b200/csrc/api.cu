// C-ABI glue: error string, engine dispatch, small utilities.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vinet {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<long long> g_up2_launches{0};
static thread_local const char* g_last_kernel = "";
void note_kernel(const char* name) { g_last_kernel = name; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int conv_gemm_simt(const vinet_conv_t* d, cudaStream_t stream);
int conv_wgrad_simt(const vinet_wgrad_t* d, cudaStream_t stream);
int conv_gemm_tc(const vinet_conv_t* d, cudaStream_t stream);
int conv_wgrad_tc(const vinet_wgrad_t* d, cudaStream_t stream);
int conv_gemm_tma(const vinet_conv_t* d, cudaStream_t stream);
int conv_wgrad_tma(const vinet_wgrad_t* d, cudaStream_t stream);
int tc_debug_set(unsigned int v);
int tma_pair_set(int v);
int stream_enable_set(int v);
int pool_fast_set(int v);
int stream_issue_clk_set(int v);
int conv_stream_tiling(const vinet_conv_t* d, int* block_n, int* n_tiles);
int conv_stream_up2_ok(const vinet_conv_t* d);
int conv_stream_win4_ok(const vinet_conv_t* d);
int conv_wgrad_halo(const vinet_wgrad_t* d, cudaStream_t stream, bool dry_run);

// the SIMT and register-gather kernels address sources densely: h pitch == Ws*ld and non-overlapping pixels
static bool dense_sources(const vinet_gather_t& g) {
  for (int i = 0; i < 2; ++i) {
    const vinet_src_t& s = g.src[i];
    if (s.ptr == nullptr) continue;
    if (s.ldh != 0 && s.ldh != (int64_t)g.Ws * s.ld) return false;
    if (s.ldb != 0 && s.ldb != (int64_t)s.T * g.Hs * g.Ws * s.ld) return false;
    if (s.ld < g.Cs) return false;
  }
  return true;
}

__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = accumulate ? dst[i] + src[i] : src[i];
}

}  // namespace vinet
using namespace vinet;

extern "C" int vinet_conv_gemm(const vinet_conv_t* d, int32_t engine, vinet_stream_t stream) {
  VINET_CHECK(d && d->g.ntaps >= 1 && d->g.ntaps <= VINET_MAX_TAPS, "conv_gemm: bad tap count");
  VINET_CHECK(d->g.B > 0 && d->g.Tr > 0 && d->g.Hr > 0 && d->g.Wr > 0, "conv_gemm: empty row space");
  VINET_CHECK(d->g.st > 0 && d->g.sh > 0 && d->g.sw > 0 && d->g.row_tstep > 0, "conv_gemm: bad strides");
  if (engine == VINET_ENGINE_TC && d->kernel == VINET_KERNEL_TMA) return conv_gemm_tma(d, (cudaStream_t)stream);
  VINET_CHECK(d->stats == nullptr, "conv_gemm: epilogue BatchNorm statistics are a feature of the TMA-fed tensor-core kernels");
  VINET_CHECK(dense_sources(d->g), "conv_gemm: only the TMA kernel reads pitched / sliding-window sources");
  if (engine == VINET_ENGINE_TC) {
    VINET_CHECK(!((d->g.src[0].xform | (d->g.src[1].ptr ? d->g.src[1].xform : 0)) & VINET_XF_UP2),
                "conv_gemm: the register-gather tensor-core kernel has no up-sampling input stage");
    return conv_gemm_tc(d, (cudaStream_t)stream);
  }
  VINET_CHECK(!(d->g.src[1].ptr && (d->g.src[1].xform & VINET_XF_UP2)) && (!(d->g.src[0].xform & VINET_XF_UP2) ||
              (d->g.mode == VINET_GATHER_FPROP && !(d->g.src[0].xform & 2) && d->g.Hs % 2 == 0 && d->g.Ws % 2 == 0)),
              "conv_gemm: VINET_XF_UP2 is for source 0 of an FPROP gather with even extents and no affine transform");
  if (engine == VINET_ENGINE_SIMT) return conv_gemm_simt(d, (cudaStream_t)stream);
  set_error("conv_gemm: unknown engine %d", engine);
  return -1;
}

extern "C" int vinet_conv_tiling(const vinet_conv_t* d, int32_t engine, int32_t* block_n, int32_t* n_tiles) {
  VINET_CHECK(d && block_n && n_tiles, "conv_tiling: null argument");
  VINET_CHECK(d->N >= 1, "conv_tiling: N %d", d->N);
  int bn = 0, nt = 0;
  if (engine == VINET_ENGINE_TC && d->kernel == VINET_KERNEL_TMA && d->g.ntaps >= 1 && d->g.ntaps <= VINET_MAX_TAPS &&
      conv_stream_tiling(d, &bn, &nt)) {
    *block_n = bn;
    *n_tiles = nt;
    return 0;
  }
  const int n16 = (int)round_up(d->N, 16);   // default: the fewest tiles of at most 256 columns
  *n_tiles = (int)cdiv(n16, 256);
  *block_n = (int)round_up(cdiv(n16, *n_tiles), 16);
  return 0;
}

extern "C" int vinet_conv_up2_fused(const vinet_gather_t* g, int32_t N, int32_t engine, int32_t kernel) {
  if (!g || g->mode != VINET_GATHER_FPROP || g->src[0].ptr == nullptr) return 0;
  if ((g->src[0].xform & ~VINET_XF_RELU) != VINET_XF_UP2 || (g->Hs & 1) || (g->Ws & 1)) return 0;
  if (g->src[1].ptr != nullptr && g->src[1].xform != VINET_XF_IDENT && engine == VINET_ENGINE_TC) return 0;
  if (engine == VINET_ENGINE_SIMT) return dense_sources(*g) ? 1 : 0;     // the FFMA gather interpolates on read (gather.cuh)
  if (engine != VINET_ENGINE_TC || kernel != VINET_KERNEL_TMA || g->dtype != VINET_BF16) return 0;
  // tensor-core engine: both the streaming fprop kernel and the halo weight-gradient kernel need their interpolating producer
  vinet_conv_t c;
  memset(&c, 0, sizeof(c));
  c.g = *g;
  c.N = N;
  c.kernel = kernel;
  c.out_dtype = VINET_BF16;
  if (!conv_stream_up2_ok(&c)) return 0;
  vinet_wgrad_t w;
  memset(&w, 0, sizeof(w));
  w.g = *g;
  w.N = N;
  w.dy_dtype = VINET_BF16;
  w.lddy = (N + 7) / 8 * 8;
  w.lddw = (N + 63) / 64 * 64;
  w.splits = 1;
  w.kernel = kernel;
  return conv_wgrad_halo(&w, nullptr, true) == 1 ? 1 : 0;
}

extern "C" int vinet_conv_win4_fused(const vinet_gather_t* g, int32_t N) {
  if (!g) return 0;
  vinet_conv_t c;
  memset(&c, 0, sizeof(c));
  c.g = *g;
  c.N = N;
  c.kernel = VINET_KERNEL_TMA;
  c.out_dtype = VINET_BF16;
  return conv_stream_win4_ok(&c);
}

extern "C" int vinet_conv_wgrad(const vinet_wgrad_t* d, int32_t engine, vinet_stream_t stream) {
  VINET_CHECK(d && d->g.ntaps >= 1 && d->g.ntaps <= VINET_MAX_TAPS, "conv_wgrad: bad tap count");
  VINET_CHECK(d->splits >= 1, "conv_wgrad: splits");
  VINET_CHECK(d->lddw >= d->N, "conv_wgrad: lddw");
  if (engine == VINET_ENGINE_TC && d->kernel == VINET_KERNEL_TMA) return conv_wgrad_tma(d, (cudaStream_t)stream);
  VINET_CHECK(dense_sources(d->g), "conv_wgrad: only the TMA kernel reads pitched / sliding-window sources");
  if (engine == VINET_ENGINE_TC) {
    VINET_CHECK(!((d->g.src[0].xform | (d->g.src[1].ptr ? d->g.src[1].xform : 0)) & VINET_XF_UP2),
                "conv_wgrad: the register-gather tensor-core kernel has no up-sampling input stage");
    return conv_wgrad_tc(d, (cudaStream_t)stream);
  }
  VINET_CHECK(!(d->g.src[1].ptr && (d->g.src[1].xform & VINET_XF_UP2)) && (!(d->g.src[0].xform & VINET_XF_UP2) ||
              (d->g.mode == VINET_GATHER_FPROP && !(d->g.src[0].xform & 2) && d->g.Hs % 2 == 0 && d->g.Ws % 2 == 0)),
              "conv_wgrad: VINET_XF_UP2 is for source 0 of an FPROP gather with even extents and no affine transform");
  if (engine == VINET_ENGINE_SIMT) return conv_wgrad_simt(d, (cudaStream_t)stream);
  set_error("conv_wgrad: unknown engine %d", engine);
  return -1;
}

extern "C" int vinet_memset_async(void* ptr, int value, size_t bytes, vinet_stream_t stream) {
  cudaError_t e = cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream);
  VINET_CHECK(e == cudaSuccess, "memset: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int vinet_axpy_f32(float* dst, const float* src, int64_t n, int32_t accumulate, vinet_stream_t stream) {
  int64_t nb = cdiv(n, 256);
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  axpy_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(dst, src, n, accumulate);
  VINET_LAUNCH_OK("axpy");
  return 0;
}

extern "C" const char* vinet_last_error(void) { return g_err; }
extern "C" const char* vinet_last_kernel(void) { return g_last_kernel; }
extern "C" const char* vinet_version(void) { return "vinet_b200 0.3 (sm_100a; streaming halo-tile tcgen05+TMEM conv with multi-warp MMA issue, per-tap TMA conv, fp32 SIMT parity engine)"; }
extern "C" int64_t vinet_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int64_t vinet_up2_launch_count(void) { return (int64_t)g_up2_launches.load(); }

extern "C" int vinet_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  VINET_CHECK(e == cudaSuccess, "device_info: %s", cudaGetErrorString(e));
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  VINET_CHECK(e == cudaSuccess, "device_info: %s", cudaGetErrorString(e));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

extern "C" int vinet_abi_sizes(int64_t* out, int32_t n) {
  const int64_t sizes[] = {sizeof(vinet_src_t),      sizeof(vinet_gather_t),     sizeof(vinet_conv_t),     sizeof(vinet_wgrad_t),
                           sizeof(vinet_pack_t),     sizeof(vinet_pack_input_t), sizeof(vinet_bn_stats_t), sizeof(vinet_bn_finalize_t),
                           sizeof(vinet_bn_apply_t), sizeof(vinet_bn_bwd_t),   sizeof(vinet_pool_t),       sizeof(vinet_upsample_t), sizeof(vinet_head_t),
                           sizeof(vinet_loss_t),     sizeof(vinet_conv1d_t),     sizeof(vinet_bn1d_t),     sizeof(vinet_avfuse_t),
                           sizeof(vinet_split_t),    sizeof(vinet_postproc_t), sizeof(vinet_preproc_t),
                           sizeof(vinet_unpack_t),   sizeof(vinet_bgemm_t),    sizeof(vinet_addln_t)};
  const int32_t m = (int32_t)(sizeof(sizes) / sizeof(sizes[0]));
  for (int32_t i = 0; i < n && i < m; ++i) out[i] = sizes[i];
  return m;
}

extern "C" int vinet_debug_set(int32_t key, int32_t value) {
  if (key == 1) return tma_pair_set(value);
  if (key == 2) return stream_enable_set(value);
  if (key == 3) return pool_fast_set(value);
  if (key == 4) return stream_issue_clk_set(value);
  VINET_CHECK(key == 0, "debug_set: unknown key %d", key);
  VINET_CHECK(tc_debug_set((unsigned int)value) == 0, "debug_set: cudaMemcpyToSymbol failed");
  return 0;
}
