/*
 * vinet_b200.h — C-ABI of the B200-native ViNet/AViNet hot path.
 *
 * The reference (samyak0210/ViNet) has no FFI for this path: every device op is a torch.nn call made
 * from model.py / model_utils.py / loss.py.  This header is the boundary a maintainer binds instead
 * (ctypes stub in INTEGRATION.md; `vinet_b200/lib.py` is that binding).  Each entry point names the
 * reference call site it replaces.  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - activations are NDHWC ("channels last"): element (b,t,h,w,c) of a view lives at
 *     ptr[(((b*T + t)*H + h)*W + w)*ld + c]; `ptr` already includes the view's channel offset and
 *     `ld` is the channel count of the underlying buffer, so channel slices of a concat buffer are views;
 *   - every function only enqueues work on `stream` (a cudaStream_t) and returns 0, or a negative
 *     error code with a message retrievable through vinet_last_error(); nothing allocates;
 *   - entry points are re-entrant (no global mutable state besides the thread-local error string).
 */
#ifndef VINET_B200_H
#define VINET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vinet_stream_t; /* cudaStream_t */

enum { VINET_BF16 = 0, VINET_F32 = 1 };
/* transform applied to a source element when it is read ("pending" BN/ReLU of the producer layer) */
enum { VINET_XF_IDENT = 0, VINET_XF_RELU = 1, VINET_XF_AFFINE = 2, VINET_XF_AFFINE_RELU = 3,
       /* bit 2: the source is STORED AT HALF RESOLUTION, [B,T,Hs/2,Ws/2,ld], and read through nn.Upsample(scale_factor=(1,2,2),
        * mode='trilinear') (model.py:254; applied after the ReLU of bit 0, not combinable with AFFINE): the up-sampled tensor of
        * the decoder (model.py:286-311) is never materialised.  Hs, Ws of the gather stay the (even) hi-res extents.  Accepted by the
        * FFMA engine and, for FPROP gathers, by the TMA-fed tensor-core kernels whose producer warps interpolate straight into
        * the swizzled shared-memory tile (vinet_conv_up2_fused tells); everything else reports an error. */
       VINET_XF_UP2 = 4, VINET_XF_RELU_UP2 = 5 };
enum { VINET_GATHER_FPROP = 0, VINET_GATHER_DGRAD = 1 };
/* VINET_ENGINE_TC: bf16 tcgen05.mma with fp32 TMEM accumulators; VINET_ENGINE_SIMT: fp32 FFMA (parity mode) */
enum { VINET_ENGINE_TC = 0, VINET_ENGINE_SIMT = 1 };
enum { VINET_ACT_NONE = 0, VINET_ACT_RELU = 1, VINET_ACT_SIGMOID = 2 };
enum { VINET_LOSS_KLDIV = 0, VINET_LOSS_CC = 1, VINET_LOSS_SIM = 2, VINET_LOSS_NSS = 3 };
/* TC engine operand feed. GATHER: producer warps gather A through registers (any stride, pending transforms).
 * TMA: A/dy tiles fetched by cp.async.bulk.tensor (5-D tiled tensor maps, hardware zero fill = conv padding),
 * persistent CTAs, double-buffered TMEM accumulators; needs bf16 sources with VINET_XF_IDENT and spatial stride 1. */
enum { VINET_KERNEL_GATHER = 0, VINET_KERNEL_TMA = 1 };
/* K index of packed weights / packed weight gradients. DENSE: k = tap*cs + c.  TAP64: every tap is padded to a
 * multiple of 64 channels, k = (tap*ceil(cs/64) + c/64)*64 + c%64 (one 64-wide K block per TMA box).
 * WIN8 (stem, Cin <= 8, kw <= 8): the gather's taps enumerate dh only and one 64-wide K block holds the whole
 * (dw, c) window of a padded NDHWC8 row: k = dh*64 + dw*8 + c.  The matching source view has Cs = 64, ld = 8*sw.
 * WIN4 (stem forward, Cin <= 4, kw <= 8): the same idea on the 4-channel copy of the clip (vinet_pack_input_t.out4):
 * k = dh*64 + dw*4 + c (dw < 16; a forward gather with Cs = 32 reads the 8 pixels of the lower half, the weight gradient and
 * the fallback kernels take 16-pixel windows, Cs = 64); the source view has ld = 4*sw.  With sw = 2
 * consecutive windows start 16 bytes apart, which lets the streaming kernel read them in place from a compact patch through
 * un-swizzled UMMA descriptors (csrc/conv_stream.cu, conv_gemm_stream_win4; vinet_conv_win4_fused tells). */
enum { VINET_KLAYOUT_DENSE = 0, VINET_KLAYOUT_TAP64 = 1, VINET_KLAYOUT_WIN8 = 2, VINET_KLAYOUT_WIN4 = 3 };

#define VINET_MAX_TAPS 64
#define VINET_TC_BLOCK_M 128
#define VINET_TC_BLOCK_K 64

/* One source of a virtual concat along T (torch.cat(..., 2), model.py:290,296,302). */
typedef struct vinet_src {
  const void* ptr;
  const float* scale; /* per-channel, for VINET_XF_AFFINE*; may be NULL otherwise */
  const float* shift;
  int64_t ld;
  int32_t T;     /* frames held by this source */
  int32_t xform; /* VINET_XF_* */
  int64_t ldh;   /* elements between consecutive h rows; 0 = dense (Ws*ld).  Only the TMA kernels accept a
                    non-dense pitch or ld < Cs (overlapping sliding-window rows, see vinet_pack_input_t) */
  int64_t ldb;   /* elements between consecutive clips; 0 = dense (T*Hs*row pitch).  ldb = ONE frame makes the batch a set of
                    overlapping sliding windows over a frame sequence (window b = frames b .. b+T-1): the inference driver
                    computes the per-frame stem conv once per frame and feeds the windows from that (TMA kernels only) */
} vinet_src_t;

/*
 * Implicit-GEMM operand gather.  GEMM rows enumerate (b, tr, h, w) over [B,Tr,Hr,Wr]; the row's
 * frame is t = tr*row_tstep + row_toff.  The K index enumerates (tap, c) with c < Cs, tap < ntaps;
 * tap[i] = (dt,dh,dw) is a kernel index of the forward convolution.
 *   FPROP: source position = (t*st - pt + dt, h*sh - ph + dh, w*sw - pw + dw)
 *   DGRAD: source position = ((t + pt - dt)/st, ...) when divisible, i.e. the transposed convolution;
 * out-of-range positions read as 0 (zero padding is applied AFTER the source transform).
 */
typedef struct vinet_gather {
  int32_t mode;  /* VINET_GATHER_* */
  int32_t dtype; /* storage type of both sources */
  int32_t B, Tr, Hr, Wr;
  int32_t row_tstep, row_toff;
  int32_t Ts, Hs, Ws; /* source extent; Ts = src[0].T + src[1].T */
  int32_t Cs;         /* channels per tap, multiple of 8 */
  int32_t ntaps;
  int32_t st, sh, sw, pt, ph, pw;
  int8_t tap[VINET_MAX_TAPS][4];
  vinet_src_t src[2]; /* src[1].ptr == NULL when there is no concat */
} vinet_gather_t;

/*
 * out[rows, N] (+)= act( gather(rows, K) x W[K, N] * ep_scale + ep_shift )
 * Replaces nn.Conv3d forward (model_utils.py:131,144,148; model.py:256-281) in FPROP mode and its
 * autograd data-gradient in DGRAD mode.  Rows whose frame t < out_T[0] are written to out[0] (frame t),
 * the others to out[1] (frame t - out_T[0]): the data-gradient of a T-concat lands in both producers.
 */
typedef struct vinet_conv {
  vinet_gather_t g;
  const void* w;    /* from vinet_pack_weights with the same engine/block_n/n_tiles */
  int32_t N;        /* real output channels */
  int32_t block_n;  /* TC engine: UMMA N per tile (multiple of 16, <= 256) */
  int32_t n_tiles;  /* TC engine: tiles along N */
  int32_t k_blocks; /* DENSE: ceil(ntaps*Cs / 64); TAP64: ntaps*ceil(Cs/64) */
  void* out[2];
  int64_t ldo[2];
  int32_t out_T[2];
  int32_t out_dtype;
  int32_t accumulate; /* bit i set: out[i] += result (read-modify-write); clear: out[i] = result */
  const float* ep_scale; /* per-output-channel, may be NULL */
  const float* ep_shift; /* per-output-channel (bias), may be NULL */
  int32_t ep_act;        /* VINET_ACT_* */
  int32_t kernel;        /* VINET_KERNEL_* (TC engine); TMA expects w packed with VINET_KLAYOUT_TAP64 */
  double* stats;         /* optional (TC engine + TMA kernels, no epilogue, accumulate == 0): BatchNorm batch statistics of the raw
                            output taken from the fp32 accumulators in the epilogue: stats[n] += sum over rows, stats[N + n] += sum
                            of squares (per-CTA partials, fp64 atomics across CTAs).  Caller zeroes; vinet_bn_apply_stats consumes.
                            Replaces the read pass of vinet_bn_stats (model_utils.py:132-138,145-150). */
} vinet_conv_t;
int vinet_conv_gemm(const vinet_conv_t* d, int32_t engine, vinet_stream_t stream);
/* N tiling (block_n, n_tiles) this library wants for the convolution described by d (every field but w / block_n /
 * n_tiles filled in): the caller packs the weights with it (vinet_pack_weights) and passes it back in d.  Host only. */
int vinet_conv_tiling(const vinet_conv_t* d, int32_t engine, int32_t* block_n, int32_t* n_tiles);

/*
 * dwp[(tap,c), n] += sum_rows gather(row,(tap,c)) * dy[row, n]   (fp32 atomics; caller zeroes dwp)
 * Replaces the autograd weight-gradient of nn.Conv3d.  g must be an FPROP gather whose rows are the
 * conv's output positions; dy is the materialised gradient w.r.t. the raw conv output.
 */
typedef struct vinet_wgrad {
  vinet_gather_t g;
  const void* dy;
  int64_t lddy;
  int32_t dy_dtype;
  int32_t N;
  float* dwp;    /* [round_up(K,128)][lddw], K = ntaps*Cs (GATHER) or ntaps*round_up(Cs,64) (TMA: TAP64 rows) */
  int32_t lddw;  /* >= round_up(N,64) */
  int32_t splits; /* CTAs along the row (reduction) dimension */
  int32_t kernel; /* VINET_KERNEL_* (TC engine) */
} vinet_wgrad_t;
int vinet_conv_wgrad(const vinet_wgrad_t* d, int32_t engine, vinet_stream_t stream);
/* 1 when BOTH vinet_conv_gemm and vinet_conv_wgrad serve the FPROP gather g (N output channels, VINET_KERNEL_* kernel) with a
 * VINET_XF_UP2 source 0 through their fused interpolating input stage, 0 when the caller has to materialise the up-sampled
 * tensor (vinet_upsample_fwd) first.  Host only. */
int vinet_conv_up2_fused(const vinet_gather_t* g, int32_t N, int32_t engine, int32_t kernel);
/* 1 when vinet_conv_gemm serves the FPROP gather g over the 4-channel clip (Cs = 32, ld = 8, VINET_KLAYOUT_WIN4 weights) with the
 * compact-patch streaming kernel; 0: use the 8-channel window view (VINET_KLAYOUT_WIN8).  Host only. */
int vinet_conv_win4_fused(const vinet_gather_t* g, int32_t N);

/* Weight (re)packing: PyTorch (Cout,Cin,kt,kh,kw) fp32 -> GEMM B operand. */
typedef struct vinet_pack {
  const float* w;
  int32_t Cout, Cin, kt, kh, kw;
  int32_t cs;    /* channels per tap in the packed K index (>= Cin for FPROP when the input is channel-padded) */
  int32_t mode;  /* FPROP: n=cout,k=(tap,cin); DGRAD: n=cin,k=(tap,cout) */
  int32_t ntaps;
  int8_t tap[VINET_MAX_TAPS][4];
  int32_t engine, block_n, n_tiles, k_blocks;
  void* out; /* TC: bf16 [n_tiles][k_blocks][block_n][64] 128B-swizzled; SIMT: fp32 [k_blocks*64][round_up(N,64)] */
  int32_t layout; /* VINET_KLAYOUT_* */
  int32_t part;   /* TC engine: which bf16 term of the weight is packed: 0 = bf16(w), 1 = bf16(w - part0), 2 = bf16(w - part0 - part1)
                     (split-precision "bf16x3" parity mode, see vinet_split_bf16); 0 everywhere else */
  int32_t ld_cin; /* input-channel extent of the tensor `w` points into when Cin names a channel SLICE of it (w already offset
                     to the slice's first channel); 0 = Cin.  The parity mode packs K chunks of a convolution separately. */
} vinet_pack_t;
int vinet_pack_weights(const vinet_pack_t* d, vinet_stream_t stream);
/* n TC-engine packs in ONE launch.  table_dev: n descriptors in DEVICE memory (each as for vinet_pack_weights, already
 * validated once through it); chunk_begin_dev[e] = sum over entries < e of n_tiles*k_blocks*block_n*8 (16-byte chunks). */
int vinet_pack_weights_multi(const vinet_pack_t* table_dev, const int64_t* chunk_begin_dev, int32_t n, int64_t total_chunks,
                             vinet_stream_t stream);
size_t vinet_packed_weight_bytes(int32_t engine, int32_t N, int32_t block_n, int32_t n_tiles, int32_t k_blocks);

/* WIN8 variant: grad[co][ci][0][dh][dw] = dwp[(dh*64 + dw*8 + ci)*lddw + co] */
int vinet_unpack_wgrad_win8(float* dwp, int32_t lddw, float* grad, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw,
                            vinet_stream_t stream);
/* grad[co][ci][tap] = dwp[(tap*cs + ci)*lddw + co]  (taps in natural (dt,dh,dw) order) */
int vinet_unpack_wgrad(float* dwp, int32_t lddw, int32_t cs, float* grad, int32_t Cout, int32_t Cin,
                       int32_t ntaps, vinet_stream_t stream);

/* n <= VINET_UNPACK_MAX unpacks in ONE launch (descriptors by value: no device table, CUDA-graph safe): a backward pass of this
 * model ends with ~80 of these 4-us kernels. */
#define VINET_UNPACK_MAX 48
typedef struct vinet_unpack {
  const float* dwp;
  float* grad;
  int32_t lddw, cs, Cout, Cin, ntaps;
  int32_t win8_kh, win8_kw; /* both 0: vinet_unpack_wgrad layout; else the window layout of vinet_unpack_wgrad_win8, with `cs` =
                               channels per pixel of the window (0 or 8: WIN8; 4: WIN4, dwp row = dh*64 + dw*4 + ci) */
  int32_t begin;            /* first flat output element of this entry in the launch's concatenated element space */
} vinet_unpack_t;
int vinet_unpack_wgrad_multi(const vinet_unpack_t* d, int32_t n, vinet_stream_t stream);

/* (B,C,T,H,W) strided fp32 clip (train.py:205 hands a permuted view) -> NDHWC with C padded to cpad. */
typedef struct vinet_pack_input {
  const float* x;
  int64_t sb, sc, st, sh, sw; /* element strides */
  int32_t B, C, T, H, W;
  int32_t cpad;
  void* out;
  int32_t out_dtype;
  int32_t wl, Wp; /* rows are written Wp >= wl + W pixels wide: wl zero pixels, the W pixels, zeros (0,0 = dense).
                     With explicit zero columns the stem conv can address a row as overlapping windows (WIN8). */
  void* out4;     /* optional (bf16 output, C <= 4): the clip as [B,T,H,Wp,4], FOUR channels per pixel (same wl / Wp, Wp even): the
                     forward stem convolution reads it in place, its weight gradient as 16-pixel windows (VINET_KLAYOUT_WIN4);
                     NULL = not written.  `out` may be NULL when out4 is given (the bf16 tensor-core mode needs only this copy) */
} vinet_pack_input_t;
int vinet_pack_input(const vinet_pack_input_t* d, vinet_stream_t stream);

/*
 * Split-precision operands for the tensor-core PARITY mode ("bf16x3"): an fp32 view x[rows, C] (after its pending
 * transform) is written as nparts bf16 planes with x ~= part[0] + part[1] (+ part[2]): part[0] = bf16(x),
 * part[1] = bf16(x - part[0]), part[2] = bf16(x - part[0] - part[1]).  A convolution is then the sum of the tcgen05 launches
 * A_i x W_j with i + j < nparts (3 launches for 2 parts: the dropped A_1 x W_1 term is 2^-16 relative), accumulated in fp32
 * through vinet_conv_t.accumulate / the wgrad atomics: the same kernels as the bf16 throughput mode reach fp32-class
 * accuracy, which is how the north-star tolerances (maps 1e-3, losses 1e-5) are gated on the measured tcgen05 path.
 */
typedef struct vinet_split {
  const void* x;
  int64_t ld;
  int32_t dtype; /* storage type of x (VINET_F32 in the parity mode) */
  int64_t rows;
  int32_t C;     /* multiple of 8 */
  const float* scale; /* pending transform of x, as for vinet_src_t */
  const float* shift;
  int32_t xform;
  int32_t nparts; /* 2 or 3 */
  void* part[3];  /* bf16 [rows, ldo] each */
  int64_t ldo;
  int32_t up_h, up_w; /* xform & VINET_XF_UP2: x holds rows / (4*up_h*up_w) low-res frames [up_h, up_w, ld]; `rows` counts the
                         hi-res rows the planes hold (the interpolation happens while splitting) */
} vinet_split_t;
int vinet_split_bf16(const vinet_split_t* d, vinet_stream_t stream);

/* ---- BatchNorm3d / BatchNorm2d (model_utils.py:132,145,149; model.py:752...) ---- */
typedef struct vinet_bn_stats {
  const void* y;
  int64_t ld;
  int32_t dtype;
  int64_t rows;
  int32_t C;
  double* sums; /* [2][C]: sum, sum of squares; must be zero on entry (vinet_bn_finalize re-zeroes it after reading) */
} vinet_bn_stats_t;
int vinet_bn_stats(const vinet_bn_stats_t* d, vinet_stream_t stream);

typedef struct vinet_bn_finalize {
  double* sums; /* training: consumed AND cleared, so the next step's vinet_bn_stats needs no memset */
  int64_t rows;
  int32_t C;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean; /* training: updated in place (unbiased variance) */
  float* running_var;
  int32_t training; /* 0: scale/shift from the running statistics */
  float* scale;     /* gamma * invstd */
  float* shift;     /* beta - mean * scale */
  float* mean;
  float* invstd;
  int64_t sq_stride; /* doubles between sum[c] and sum-of-squares[c] in `sums`; 0 = C.  A layer that is a channel slice of a fused
                        convolution (Mixed_* 1x1 group) reads its slice of the group's statistics with sq_stride = the group's N. */
} vinet_bn_finalize_t;
int vinet_bn_finalize(const vinet_bn_finalize_t* d, vinet_stream_t stream);
/* vinet_bn_stats + vinet_bn_finalize (training) in ONE launch: the last block to finish finalises.  Both descriptors must
 * name the same sums buffer, which here holds [2][C] doubles FOLLOWED BY one 8-byte ticket (zero on entry, left zero). */
int vinet_bn_stats_finalize(const vinet_bn_stats_t* d, const vinet_bn_finalize_t* f, vinet_stream_t stream);

/* out = relu?(scale*y + shift): materialises BatchNorm(+ReLU) of a raw conv output (possibly into a channel
 * slice of a concat buffer, model_utils.py:187), so that consumers read it untransformed (TMA-fed kernels). */
typedef struct vinet_bn_apply {
  const void* y;
  int64_t ldy;
  int32_t dtype;
  int64_t rows;
  int32_t C;
  int32_t relu;
  const float* scale;
  const float* shift;
  void* out;
  int64_t ldo;
  int32_t out_dtype;
} vinet_bn_apply_t;
int vinet_bn_apply(const vinet_bn_apply_t* d, vinet_stream_t stream);
/* Training-mode finalisation + materialisation in ONE launch for n <= 4 layers whose statistics were accumulated by the
 * convolution epilogue (vinet_conv_t.stats): every thread derives scale / shift of its 8 channels from f[i].sums, the first block
 * of a layer also publishes scale / shift / mean / invstd and updates the running statistics (momentum, unbiased variance);
 * then out = relu?(scale*y + shift).  f[i].sums is only read (the caller zeroes its statistics arena once per step). */
int vinet_bn_apply_stats_multi(const vinet_bn_finalize_t* f, const vinet_bn_apply_t* a, int32_t n, vinet_stream_t stream);

/* Backward of y_hat = relu?(scale*y + shift) w.r.t. the raw conv output y. */
typedef struct vinet_bn_bwd {
  const void* g; /* grad w.r.t. the activated output (g_dtype) */
  int64_t ldg;
  const void* y;
  int64_t ldy;
  int32_t dtype;
  int64_t rows;
  int32_t C;
  int32_t relu;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  const float* gamma;
  double* sums; /* [2][C] sum(g*m), sum(g*m*y_norm) followed by one 8-byte ticket; zero on entry, cleared again by
                   vinet_bn_bwd_reduce on exit (its last block also writes dgamma / dbeta) */
  float* dgamma;
  float* dbeta;
  void* dy;
  int64_t lddy;
  int32_t dy_dtype;
  int32_t training; /* 0: eval-mode BN (running stats): dy = g*m*scale, no mean/projection terms */
  int32_t g_dtype;
} vinet_bn_bwd_t;
int vinet_bn_bwd_reduce(const vinet_bn_bwd_t* d, vinet_stream_t stream);
int vinet_bn_bwd_apply(const vinet_bn_bwd_t* d, vinet_stream_t stream);
/* Whole-layer variants (one cooperative launch each): training forward = stats + finalize + apply, backward = reduce + apply.
 * The blocks reduce, the last one finalises and raises a flag, all of them then apply to the rows they have just read.
 * The sums buffers hold [2][C] doubles followed by THREE 8-byte words (ticket, flag, departures), zero between launches. */
int vinet_bn_fwd_fused(const vinet_bn_stats_t* d, const vinet_bn_finalize_t* f, const vinet_bn_apply_t* a, vinet_stream_t stream);
/* Multi-layer variants: n <= 4 independent BatchNorm layers (arrays of n descriptors, same storage types) in ONE launch per
 * pass - small layers cost ~6 us per launch whatever they do.  stats_finalize_multi = batch statistics + finalisation
 * (sums buffers as for vinet_bn_stats_finalize), apply_multi = materialisation, bwd_multi = reduce launch + apply launch. */
int vinet_bn_stats_finalize_multi(const vinet_bn_stats_t* d, const vinet_bn_finalize_t* f, int32_t n, vinet_stream_t stream);
int vinet_bn_apply_multi(const vinet_bn_apply_t* a, int32_t n, vinet_stream_t stream);
int vinet_bn_bwd_multi(const vinet_bn_bwd_t* b, int32_t n, vinet_stream_t stream);
int vinet_bn_bwd_fused(const vinet_bn_bwd_t* d, vinet_stream_t stream);

/* ---- nn.MaxPool3d (model.py:696-714, model_utils.py:178...; model.py:229) ---- */
typedef struct vinet_pool {
  const void* x;
  int64_t ldx;
  int32_t dtype;
  const float* scale;
  const float* shift;
  int32_t xform;
  int32_t B, Ti, Hi, Wi, C;
  int32_t kt, kh, kw, st, sh, sw, pt, ph, pw;
  int32_t To, Ho, Wo;
  void* out;
  int64_t ldo;
  int32_t out_dtype;
  const void* gout; /* backward: grad w.r.t. out (gout_dtype) */
  int64_t ldgo;
  void* gin; /* backward: grad w.r.t. the activated input (gin_dtype), accumulated with atomics (caller initialises) */
  int64_t ldgi;
  int32_t gout_dtype, gin_dtype;
  uint8_t* idx; /* optional [B,To,Ho,Wo,C] bytes: forward records the winning tap ((dt*kh+dh)*kw+dw, first maximum in
                   scan order) and backward scatters with it (or gathers per input element, vinet_debug_set key 3);
                   NULL: backward recomputes the arg-max */
  int32_t gin_overwrite; /* backward: 1 = gin is written (the pool is the first writer of that gradient), 0 = gin += */
} vinet_pool_t;
int vinet_maxpool_fwd(const vinet_pool_t* d, vinet_stream_t stream);
int vinet_maxpool_bwd(const vinet_pool_t* d, vinet_stream_t stream);

/* ---- relu? + nn.Upsample((1,2,2),'trilinear') (model.py:254): per-frame 2x bilinear ---- */
typedef struct vinet_upsample {
  const void* z;
  int64_t ldz;
  int32_t dtype;
  int32_t relu;
  int32_t B, T, h, w, C;
  void* u; /* [B,T,2h,2w,C] */
  int64_t ldu;
  int32_t u_dtype;
  const void* gu; /* backward in: grad w.r.t. u (gu_dtype) */
  int64_t ldgu;
  void* dz; /* backward out: grad w.r.t. raw z (ReLU-masked) */
  int64_t lddz;
  int32_t dz_dtype;
  int32_t gu_dtype;
} vinet_upsample_t;
int vinet_upsample_fwd(const vinet_upsample_t* d, vinet_stream_t stream);
int vinet_upsample_bwd(const vinet_upsample_t* d, vinet_stream_t stream);

/* dz = g where z > 0 else 0: backward of a ReLU whose input (or output: same mask) was kept (decoder convs, model.py:256-281) */
int vinet_relu_bwd(const void* g, int64_t ldg, int32_t g_dtype, const void* z, int64_t ldz, int32_t z_dtype, int64_t rows, int32_t C,
                   void* dz, int64_t lddz, int32_t dz_dtype, vinet_stream_t stream);

/* ---- decoder head: relu? -> Conv3d(C,1,1x1x1,bias) -> Sigmoid (model.py:280-283) ----
 * up2 != 0: the 2x bilinear up-sampling in front of the head (model.py:278, 340, 402, 464) is fused into its input stage: x holds
 * rows / (4*up_h*up_w) low-res frames [up_h, up_w, ldx], `rows` counts hi-res pixels, relu_pre applies a ReLU to x BEFORE the
 * interpolation (conv -> ReLU -> up -> head, T = 8 / 16) and relu AFTER it (conv -> ReLU -> up -> (kt,1,1) conv -> ReLU -> head, T = 32 / 48, where
 * the pointwise-in-space conv has been commuted in front of the up-sampling); dx is the hi-res gradient w.r.t. the interpolated
 * value (vinet_upsample_bwd turns it into the gradient w.r.t. x). */
typedef struct vinet_head {
  const void* x;
  int64_t ldx;
  int32_t dtype;
  int32_t relu;
  int64_t rows;
  int32_t C; /* <= 64 */
  const float* w;
  const float* b;
  float* out;        /* [rows] */
  const float* gout; /* backward */
  void* dx;          /* grad w.r.t. x (ReLU-masked) */
  int64_t lddx;
  int32_t dx_dtype;
  float* dw; /* [C], atomics; caller zeroes */
  float* db; /* [1] */
  int32_t up2, up_h, up_w, relu_pre;
} vinet_head_t;
int vinet_head_fwd(const vinet_head_t* d, vinet_stream_t stream);
int vinet_head_bwd(const vinet_head_t* d, vinet_stream_t stream);

/* ---- losses (loss.py:13-120): mean over the batch of a per-sample reduction ---- */
typedef struct vinet_loss {
  int32_t kind; /* VINET_LOSS_* */
  const float* s; /* [B,n] prediction */
  const float* g; /* [B,n] ground truth / fixation map */
  int32_t B;
  int32_t n;
  float* per_sample; /* [B][8] workspace: value + saved reductions for the backward */
  float* out;        /* [1] */
  int32_t* counter;  /* [1] zero-initialised once; left at zero */
  const float* gout; /* backward: device scalar */
  float* grad_s;     /* [B,n] */
} vinet_loss_t;
int vinet_loss_fwd(const vinet_loss_t* d, vinet_stream_t stream);
int vinet_loss_bwd(const vinet_loss_t* d, vinet_stream_t stream);

/* ---- SoundNet 1-D convs (model.py:750-786) and the audio-visual bilinear fusion (model.py:229-237) ---- */
typedef struct vinet_conv1d {
  const float* x; /* [B,Cin,Lin] fp32 */
  const float* w; /* [Cout,Cin,k] */
  const float* bias;
  int32_t B, Cin, Lin, Cout, Lout, k, stride, pad;
  float* y; /* [B,Cout,Lout] raw conv output (+bias) */
  const float* dy; /* backward */
  float* dx;       /* [B,Cin,Lin] (may be NULL) */
  float* dw;       /* [Cout,Cin,k] */
  float* dbias;    /* [Cout] */
} vinet_conv1d_t;
int vinet_conv1d_fwd(const vinet_conv1d_t* d, vinet_stream_t stream);
int vinet_conv1d_bwd(const vinet_conv1d_t* d, vinet_stream_t stream);

/* BatchNorm2d(+ReLU)(+MaxPool (p,1)) on [B,C,L] fp32, materialised (the audio branch is 0.19 GFLOP). */
typedef struct vinet_bn1d {
  const float* y; /* raw conv output [B,C,L] */
  int32_t B, C, L, pool;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  int32_t training;
  float* mean;   /* [C] saved */
  float* invstd; /* [C] saved */
  float* out;    /* [B,C,L/pool] */
  const float* gout; /* backward */
  float* dy;         /* [B,C,L] */
  float* dgamma;
  float* dbeta;
} vinet_bn1d_t;
int vinet_bn1d_fwd(const vinet_bn1d_t* d, vinet_stream_t stream);
int vinet_bn1d_bwd(const vinet_bn1d_t* d, vinet_stream_t stream);

/* fused[b,c,o] = sum_ij v[b,c,i] W[o,i,j] a[b,c,j] + bias[o]; v = max over 4 frames / every 2nd column of y0. */
typedef struct vinet_avfuse {
  const void* y0; /* [B,4,7,12,C] NDHWC view with pending transform */
  int64_t ld;
  int32_t dtype;
  const float* scale;
  const float* shift;
  int32_t xform;
  const float* audio; /* [B,C,3] */
  const float* w;     /* [336,42,3] */
  const float* bias;  /* [336] */
  int32_t B, C;
  float* vbuf; /* [B,C,42] pooled visual features (written by fwd, read by bwd) */
  void* out; /* [B,4,7,12,C] NDHWC, identity transform */
  int64_t ldo;
  int32_t out_dtype;
  const float* gout; /* backward: fp32 grad w.r.t. out [B,4,7,12,C] */
  int64_t ldgo;
  float* gy0; /* fp32 grad w.r.t. activated y0 (atomics) */
  int64_t ldgy0;
  float* gaudio; /* [B,C,3] */
  float* dw;     /* [336,42,3] atomics; caller zeroes */
  float* dbias;  /* [336] */
} vinet_avfuse_t;
int vinet_avfuse_fwd(const vinet_avfuse_t* d, vinet_stream_t stream);
int vinet_avfuse_bwd(const vinet_avfuse_t* d, vinet_stream_t stream);

/* ---- transformer fusion variants (model.py:8-69 PositionalEncoding / Transformer; model.py:211-221,239-247
 * VideoAudioSaliencyModel(use_transformer=True); model.py:116-189 VideoAudioSaliencyFusionModel).  The path is tiny (32-339 tokens of
 * 336-512 features): fp32 FFMA kernels; every matrix product, layout shuffle and reduction below is ONE strided batched GEMM. ---- */
/* C[b1,b2](m,n) = alpha * sum_k A[b1,b2](m,k) * B[b1,b2](n,k) + bias1(m,n) + bias2[b1,b2](m,n)  (then ReLU, then += C if accumulate).
 * Element (i,j) of batch (b1,b2) of an operand X lives at X[i*sXi + j*sXj + b1*sXb1 + b2*sXb2] (elements, any sign-free stride incl. 0),
 * so transposes, head slices of a packed QKV matrix, NDHWC rows and broadcast vectors are all views.  K = 0 is a strided copy of
 * bias2.  Replaces nn.Linear / F.linear of nn.MultiheadAttention + nn.TransformerEncoderLayer, the two bmm of the attention, the
 * Conv3d/Conv2d 1x1 with bias (model.py:135,148,213-214), permute/flatten/view/cat/mean/repeat (model.py:151-183,240-247). */
typedef struct vinet_bgemm {
  const void* A; /* (M,K) */
  int64_t sAm, sAk, sAb1, sAb2;
  int32_t a_dtype;
  int32_t a_relu; /* 1: max(.,0) on read, after the affine below */
  const float* a_scale; /* optional pending BatchNorm transform of A (NULL: none), per k - or per m when a_xf_on_m */
  const float* a_shift;
  const void* B; /* (N,K) */
  int64_t sBn, sBk, sBb1, sBb2;
  int32_t b_dtype;
  int32_t c_dtype;
  void* C; /* (M,N) */
  int64_t sCm, sCn, sCb1, sCb2;
  int32_t M, N, K, nb1, nb2;
  float alpha;
  int32_t relu;       /* ReLU epilogue */
  const float* bias1; /* optional, shared by all batches */
  int64_t s1m, s1n;
  const float* bias2; /* optional, per batch */
  int64_t s2m, s2n, s2b1, s2b2;
  int32_t accumulate; /* bit 0: C += result (fp32 C only); bit 1: the reduction order is free (gradients): a long K over few
                       * output tiles may then be split across CTAs that combine with fp32 atomics */
  int32_t a_xf_on_m;  /* a_scale / a_shift are indexed by A's row m instead of the reduction index k */
} vinet_bgemm_t;
int vinet_bgemm(const vinet_bgemm_t* d, vinet_stream_t stream);
/* rows of n contiguous fp32: p = softmax(s) in place (F.softmax(attn, dim=-1)); backward in place on dp: ds = p * (dp - sum(dp * p)) */
int vinet_softmax_fwd(float* s, int64_t rows, int32_t n, vinet_stream_t stream);
int vinet_softmax_bwd(const float* p, float* dp, int64_t rows, int32_t n, vinet_stream_t stream);
/* nn.Dropout (train mode): y[i] = keep(i) ? x[i] / (1 - p) : 0, mask[i] = keep(i); keep(i) is a counter-based hash of
 * (rng[0] = seed, rng[1] = step counter, salt, i), read from DEVICE memory so that a replayed CUDA graph draws fresh masks
 * (vinet_rng_advance bumps the counter once per forward).  y may alias x.  Not bit-compatible with torch's Philox stream. */
int vinet_dropout_fwd(const float* x, float* y, uint8_t* mask, int64_t n, float p, const int64_t* rng, uint32_t salt,
                      vinet_stream_t stream);
/* out[i] = g[i] * (mask ? mask[i] / (1 - p) : 1) * (relu_ref ? relu_ref[i] > 0 : 1); out may alias g */
int vinet_dropout_bwd(const float* g, float* out, const uint8_t* mask, const float* relu_ref, int64_t n, float p,
                      vinet_stream_t stream);
int vinet_rng_advance(int64_t* rng, vinet_stream_t stream);
/* out = LayerNorm(x + y) over rows of n <= 512 contiguous fp32 (post-norm residual of nn.TransformerEncoderLayer, eps 1e-5).
 * Backward: dz = d(x + y), dgamma / dbeta ACCUMULATED (caller zeroes). */
typedef struct vinet_addln {
  const float* x; /* residual input [rows, n] */
  const float* y; /* sub-layer output [rows, n] */
  float* z;       /* x + y, saved for the backward */
  float* stat;    /* [rows, 2] mean, rstd */
  const float* gamma;
  const float* beta;
  float eps;
  int32_t n;
  int64_t rows;
  float* out;
  const float* gout; /* backward */
  float* dz;
  float* dgamma;
  float* dbeta;
} vinet_addln_t;
int vinet_add_layernorm_fwd(const vinet_addln_t* d, vinet_stream_t stream);
int vinet_add_layernorm_bwd(const vinet_addln_t* d, vinet_stream_t stream);

/* ---- sliding-window inference post-processing (generate_result.py:100-104, utils.py:61-78) ---- */
typedef struct vinet_postproc {
  const float* x; /* [N, H, W] saliency maps */
  int32_t N, H, W;
  int32_t oh, ow;  /* source-video resolution the maps are resized to (cv2.resize, bilinear) */
  int32_t blur;    /* 1: 11x11 Gaussian blur (sigma 2, reflect-101 border) after the resize */
  float* ws0;      /* [N, oh, ow] workspace */
  float* ws1;      /* [N, oh, ow] workspace */
  float* minmax;   /* [N, 2] workspace: per-map min / max after the blur */
  uint8_t* out;    /* [N, oh, ow]: round(255 * (v - min) / (max - min + 1e-5) + 0.5) */
} vinet_postproc_t;
int vinet_saliency_postprocess(const vinet_postproc_t* d, vinet_stream_t stream);

/* ---- input pipeline (dataloader.py:242-249, generate_result.py:77-89; dataloader.py:113-118) ---- */
typedef struct vinet_preproc {
  const uint8_t* frames; /* [N, h, w, 3] decoded RGB frames */
  int32_t N, h, w;
  int32_t H, W;          /* output size (224, 384 in the reference) */
  const int32_t* xb;     /* [W][2] first source column / tap count per output column  (Pillow Resample.c precompute_coeffs) */
  const int32_t* xk;     /* [W][xks] 22-bit fixed-point coefficients                   (normalize_coeffs_8bpc) */
  int32_t xks;
  const int32_t* yb;     /* [H][2], [H][yks]: the same for the vertical pass */
  const int32_t* yk;
  int32_t yks;
  uint8_t* tmp;          /* [N, h, W, 3] workspace: the 8-bit image after the horizontal pass */
  float mean[3], std[3]; /* transforms.Normalize */
  float* out;            /* [N, 3, H, W] fp32: ((resized / 255) - mean) / std, bit-identical to the PIL + torchvision pipeline */
} vinet_preproc_t;
int vinet_preprocess_frames(const vinet_preproc_t* d, vinet_stream_t stream);
/* out[b, :] = zeros(total) with excerpt[b, :n] * np.hanning(n) centred (dataloader.py:113-118) */
int vinet_audio_window(const float* excerpt, int32_t B, int32_t n, float* out, int32_t total, vinet_stream_t stream);

/* ---- misc ---- */
int vinet_memset_async(void* ptr, int value, size_t bytes, vinet_stream_t stream);
/* dst[i] = src[i] (+ dst[i] if accumulate); fp32 */
int vinet_axpy_f32(float* dst, const float* src, int64_t n, int32_t accumulate, vinet_stream_t stream);
/* column sums of a [rows, C] view -> out[C] fp32 (bias gradients) */
int vinet_colsum(const void* x, int64_t ld, int32_t dtype, int64_t rows, int32_t C, double* ws, float* out,
                 vinet_stream_t stream);
const char* vinet_last_error(void);
/* name of the CUDA kernel the calling thread's last vinet_conv_gemm / vinet_conv_wgrad launched (measurement aid) */
const char* vinet_last_kernel(void);
const char* vinet_version(void);
int vinet_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);
/* sizeof() of every struct above, in declaration order (vinet_bgemm_t and vinet_addln_t last) (host-only; lets a binding check its layout) */
int vinet_abi_sizes(int64_t* out, int32_t n);
/* development switches (key 0: tcgen05 descriptor-encoding experiments, csrc/conv_tc.cu, 0 in production;
 * key 1: paired 256-row work items of the TMA conv kernel, 1 in production;
 * key 2: the streaming (halo / frame re-use) conv kernels of csrc/conv_stream.cu and conv_wgrad_halo.cu, 1 in production;
 * key 3: max-pool kernels, bit 0 = frame-walking 3x3x3 forward, bit 1 = gather (atomic-free) backward; 1 in production) */
int vinet_debug_set(int32_t key, int32_t value);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t vinet_launch_count(void);
/* how many of them were tensor-core launches whose producer warps interpolated a VINET_XF_UP2 source (the fused up-sampling) */
int64_t vinet_up2_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VINET_B200_H */
