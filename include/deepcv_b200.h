/*
 * deepcv_b200.h — C ABI of the B200-native DeepcvModule conv / BatchNorm / augment hot path.
 *
 * One shared library (deepcv_b200/libdeepcv_b200.so), plain pointers and sizes only: every pointer below is a DEVICE
 * pointer unless its name ends in `_host`; `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default
 * stream). No call allocates, synchronises or copies to the host: outputs are caller-allocated (the Python host side
 * hands out torch caching-allocator memory), so every entry point is CUDA-graph capturable.
 *
 * Return value: 0 on success, non-zero on failure; `dcv_last_error()` then returns a thread-local message. There is no
 * CPU implementation behind any entry point: a missing / non-sm_100 GPU is an error, not a fallback.
 *
 * Activation tensors are NHWC ("channels last"), `dtype` = DCV_F32 or DCV_BF16. Convolution weights are
 * [K][R][S][C] (out-channel, kernel-row, kernel-col, in-channel), i.e. the memory of a torch OIHW tensor held in
 * channels_last format. Each entry point cites the reference call site it stands in for (paths relative to
 * /root/reference/src/deepcv/).
 */
#ifndef DEEPCV_B200_H_
#define DEEPCV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCV_ABI_VERSION 4

enum dcv_dtype { DCV_F32 = 0, DCV_BF16 = 1 };
enum dcv_act { DCV_ACT_NONE = 0, DCV_ACT_RELU = 1, DCV_ACT_LEAKY_RELU = 2, DCV_ACT_SIGMOID = 3 };
enum dcv_conv_algo { DCV_ALGO_AUTO = 0, DCV_ALGO_DIRECT = 1, DCV_ALGO_TCGEN05 = 2 };

/* Geometry of one 2-D convolution: x[n][h][w][c] (*) w[k][r][s][c] -> y[n][p][q][k]. */
typedef struct dcv_conv_shape {
  int32_t n, h, w, c;
  int32_t k, r, s;
  int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int32_t p, q;
} dcv_conv_shape;

/* ---- library ------------------------------------------------------------------------------------------------- */
int dcv_abi_version(void);
const char* dcv_last_error(void);
/* 0 iff the current CUDA device exists and is compute capability 10.x (B200). */
int dcv_device_check(void);
/* Accumulator outputs (statistics, sums and gradients that the kernels fill with atomics: stats_nc, s_nc, dbias_c, dw, dw_col, db) are zeroed by the
 * entry point that fills them unless the caller passes `acc_prezeroed` != 0: it then guarantees that it has zeroed them itself on the same stream (one
 * memset per step instead of one per kernel; see deepcv_b200/ops.py AccumulatorArena). The flag travels with each call: there is no library-wide state
 * besides the launch counter and the thread-local error string. */
/* Number of kernel launches issued through this library since load (for bench.py's `gpu_launches`). */
uint64_t dcv_launch_count(void);

/* ---- preprocess / augmentation ---------------------------------------------------------------------------------
 * Replaces torchvision ToTensor + Normalize applied per sample in PreprocessedDataset.__getitem__
 * (meta/data/preprocess.py:44-57, recipe conf/base/parameters.yml:197-210) and the flip / crop the augmentation recipe
 * names (meta/data/augmentation.py:39-44, parameters.yml:152,157): crop (zero pad `pad`, offsets crop_yx[n] = {top,left})
 * -> horizontal flip (flip[n] != 0) on the uint8 image, then (u8/255 - mean[c]) / std[c] in fp32.
 * src: uint8 [n][h][w][c]; dst: [n][out_h][out_w][c_out] (NHWC, channels >= c are written as 0) or [n][c][out_h][out_w]
 * when `nchw_out` != 0 (then c_out must equal c). flip / crop_yx may be NULL (no flip / centred crop at (pad,pad)). */
int dcv_preprocess_u8(const uint8_t* src, void* dst, int n, int h, int w, int c, int out_h, int out_w, int pad,
                      const float* mean, const float* std, const uint8_t* flip, const int32_t* crop_yx,
                      int out_dtype, int c_out, int nchw_out, void* stream);

/* ---- layout / dtype --------------------------------------------------------------------------------------------- */
/* [n][c][h][w] -> [n][h][w][c] and back, converting dtype on the way (torch.nn.Flatten of parameters.yml:87 is the
 * NHWC->NCHW direction with h*w*c flattened). */
int dcv_nchw_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype, int n, int c, int h, int w, void* stream);
int dcv_nhwc_to_nchw(const void* src, int src_dtype, void* dst, int dst_dtype, int n, int c, int h, int w, void* stream);
int dcv_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t count, void* stream);
/* fp32 [K][R][S][C] weights -> `dst_dtype` copy; with `transpose_flip` != 0 writes [C][R-1-r][S-1-s][K] (the operand of
 * the data-gradient convolution). */
int dcv_pack_conv_weight(const float* w_krsc, void* dst, int dst_dtype, int k, int r, int s, int c, int transpose_flip, void* stream);
/* The same transposed + flipped copy for MANY layers in one launch (once per training step): entry e reads the [K][R][S][C] fp32 weight at
 * flat_params + src_off and writes its [C][R-1-r][S-1-s][K] copy at dst + dst_off (elements of dst_dtype). unit0 = number of work units of the entries
 * before e, one unit = (tap, 8-channel block, 32-filter block): units(e) = r * s * ceil(c / 8) * ceil(k / 32); total_units = their sum. <= 64 entries
 * (device array). */
typedef struct dcv_pack_entry { uint64_t src_off, dst_off; long long unit0; int32_t k, r, s, c; } dcv_pack_entry;
int dcv_pack_conv_weights_batched(const float* flat_params, void* dst, int dst_dtype, const dcv_pack_entry* entries_dev, int n_entries, long long total_units, void* stream);

/* Explicit im2col for convolutions the implicit-GEMM tensor-core kernel cannot address (few input channels, strides — the 7x7/stride-2 stem):
 * col[n][p][q][kpad] holds the (r, s, c) receptive field of every output pixel in [K][R][S][C] weight order, zero padded to kpad; the convolution is
 * then the 1x1 convolution of `col` with the weights padded to [K][kpad]. */
int dcv_im2col(const dcv_conv_shape* shape, const void* x, void* col, int kpad, int dtype, void* stream);
int dcv_fill_zero(void* dst, size_t bytes, void* stream);
/* Batch assembly of the device-resident input pipeline (replaces the DataLoader + collate of meta/ignite_training.py:211-218 and the prefetcher of
 * meta/data/datasets.py:76-115): dst row j = src row idx[j] for j < n; rows of `row_bytes` (a multiple of 4; 16-byte multiples take the vector path) bytes. An index outside [0, n_src) copies
 * nothing and sets *err_flag (device int, may be NULL) to 1. */
int dcv_gather_rows(const void* src, const int64_t* idx, void* dst, int n, long long n_src, size_t row_bytes, int* err_flag, void* stream);

/* ---- convolution (torch.nn.Conv2d built at meta/submodule_creators.py:251, run at meta/nn.py:553) ---------------- */
/* 1 iff DCV_ALGO_AUTO would run `op` (0 = forward, 1 = data gradient, 2 = weight gradient) of this shape / dtype on the tcgen05 kernels
 * (bf16, stride 1, dilation 1, c and k multiples of 64); lets the caller skip preparing the `wt` operand otherwise. */
int dcv_conv2d_tc_supported(const dcv_conv_shape* shape, int dtype, int op);
/* Convolutions the TMA-fed kernels cannot address (few input channels, strides: the 3 -> 64, 7x7 / stride-2 stem of the ImageNet-shaped nets) without
 * materialising col[n][p][q][kpad]: producer warps build the im2col tile in shared memory (software gather into the tcgen05 operand layout).
 * `w_col` = [K][kpad] bf16 from dcv_gather_pack_weight; kpad a multiple of 64, <= 256; K = 64 or 128; dilation_w = 1.
 * dcv_conv2d_gather_supported: 1 iff this shape / pointer alignment is served (else use dcv_im2col + the 1x1 GEMM). */
int dcv_conv2d_gather_supported(const dcv_conv_shape* shape, const void* x, int kpad, int dtype);
/* K order of the gather kernels ("row padded"): column r*RP + x of w_col / dw_col = w[k][r][x] for x < S*C, RP = S*C rounded up to 8; kpad >= R*RP.
 * dcv_gather_pack_weight: [K][R][S*C] (dtype) -> w_col [K][kpad] (same dtype, zeros elsewhere); dcv_gather_unpack_wgrad: dw_col [K][kpad] -> [K][R][S*C], fp32. */
int dcv_gather_pack_weight(const void* w_krsc, void* w_col, int k, int r, int sc, int kpad, int dtype, void* stream);
int dcv_gather_unpack_wgrad(const float* dw_col, float* dw_krsc, int k, int r, int sc, int kpad, void* stream);
int dcv_conv2d_fwd_gather(const dcv_conv_shape* shape, const void* x, const void* w_col, int kpad, const float* bias, void* y, float* stats_nc,
                          int act, float slope, int acc_prezeroed, void* stream);
/* dw_col[K][kpad] (fp32, gather K order, overwritten) = sum over pixels of dy * im2col(x); unpack with dcv_gather_unpack_wgrad. */
int dcv_conv2d_wgrad_gather(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw_col, int kpad, int acc_prezeroed, void* stream);
/* Stride-2 few-channel convolutions (stride_w = 2, C <= 4, S + (pad_w & 1) <= 8, R <= 8, K = 64 or 128: the 3 -> 64, 7x7 / stride-2 stem) on the "pixel pair"
 * kernels: the input rows of an output row are staged one 128-byte line per pixel PAIR and the tensor core reads overlapping K-major (forward) / MN-major
 * (weight gradient) tiles straight out of those lines — no im2col tile is built. w_col / dw_col: [K][256] in the K order
 * column pp*64 + r*8 + px*C + c = w[k][r][2*pp + px - (pad_w & 1)][c] (dcv_pairs_pack_weight / dcv_pairs_unpack_wgrad; dw_col fp32, overwritten).
 * Arguments otherwise as the gather entry points above. */
int dcv_conv2d_pairs_supported(const dcv_conv_shape* shape, const void* x, int dtype);
int dcv_pairs_pack_weight(const void* w_krsc, void* w_col, const dcv_conv_shape* shape, int dtype, void* stream);
int dcv_pairs_unpack_wgrad(const float* dw_col, float* dw_krsc, const dcv_conv_shape* shape, void* stream);
int dcv_conv2d_fwd_pairs(const dcv_conv_shape* shape, const void* x, const void* w_col, const float* bias, void* y, float* stats_nc,
                         int act, float slope, int acc_prezeroed, void* stream);
int dcv_conv2d_wgrad_pairs(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw_col, int acc_prezeroed, void* stream);
/* Flags of the `acc_prezeroed` argument of dcv_conv2d_fwd / dcv_conv2d_fwd_gather / dcv_norm_stats / dcv_norm_bwd_reduce (other entry points: 0 / 1). */
#define DCV_ACC_PREZEROED 1          /* the caller has zeroed every accumulator this call adds into */
#define DCV_STATS_CHANNEL_TOTALS 2   /* only the per-CHANNEL totals of the [n][c][..] sums will be used (a BatchNorm-only block: no GroupNorm / InstanceNorm):
                                      * they are credited to image 0 (rows of the other images stay zero) and the batch is reduced as one image */
#define DCV_STATS_IN_EPILOGUE 4      /* with DCV_STATS_CHANNEL_TOTALS: the tcgen05 convolution kernels produce the statistics in their epilogue (opt-in) */
/* y = act(conv(x, w) + bias); if stats_nc != NULL also accumulates per-(image, channel) sum(y) and sum(y*y) of the
 * values written to y into stats_nc[n][k][2] (fp32, overwritten). bias may be NULL. */
int dcv_conv2d_fwd(const dcv_conv_shape* shape, const void* x, const void* w, const float* bias, void* y, float* stats_nc,
                   int act, float slope, int dtype, int algo, int acc_prezeroed, void* stream);
/* dx[n][h][w][c] = sum_{k,r,s} dy[n][p][q][k] * w[k][r][s][c]; `w` as in fwd, `wt` the transpose_flip packing (may be
 * NULL for the direct algorithm). */
int dcv_conv2d_dgrad(const dcv_conv_shape* shape, const void* dy, const void* w, const void* wt, void* dx, int dtype, int algo, void* stream);
/* dw[k][r][s][c] (fp32) = sum_{n,p,q} dy[n][p][q][k] * x[n][..][..][c]. dw is overwritten. `workspace` (fp32, at least
 * dcv_conv2d_wgrad_workspace(shape) bytes, may be NULL when that is 0) is scratch for split accumulation. */
size_t dcv_conv2d_wgrad_workspace(const dcv_conv_shape* shape, int dtype, int algo);
int dcv_conv2d_wgrad(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw, void* workspace, int dtype, int algo, int acc_prezeroed, void* stream);

/* ---- few-channel convolution blocks with the normalisation folded into their neighbours (deepcv_b200/csrc/conv_small.cu) ----------------------------
 * The default `image_classifier` (conf/base/parameters.yml:8-19,79-88: 3/4/16 channels, 5x5 / 3x3 filters, 32x32 / 16x16 maps; block order
 * Conv2d -> act -> BatchNorm2d -> GroupNorm, meta/nn.py:553) in bf16: per block ONE forward, ONE weight-gradient and ONE data-gradient launch of
 * warp-level tensor-core (mma.sync) implicit-GEMM kernels. A block's BatchNorm o GroupNorm is never a pass of its own: the block's kernel leaves its RAW
 * output y plus raw sums, and whoever consumes y applies z = A[n][c]*y + B[n][c] while loading it, deriving A, B per image from the sums (`dcv_sc_norm`
 * with enabled != 0 = "this tensor is a raw output with this pending normalisation"). Backward mirrors it: the consumer's data-gradient kernel writes
 * dz = gradient w.r.t. z and the sums the BatchNorm / GroupNorm adjoint needs; the block's own backward kernels apply dy = act'(y)*(P*dz + Q*y + R) while
 * loading. All buffers of a dcv_sc_norm are caller-allocated fp32; bn_sums and u_sums are ACCUMULATED into with atomics and must be zero before the
 * first kernel of the step that touches them (dcv_sc_norm_floats gives the sizes). */
typedef struct dcv_sc_norm {
  int32_t enabled;                  /* 0: the tensor is plain, nothing below is read */
  int32_t n, c, hw;                 /* the normalised tensor: n images of hw pixels of c channels (c even, <= 32) */
  int32_t use_bn, bn_training;      /* as dcv_norm_params */
  float bn_eps, bn_momentum;        /* momentum < 0: cumulative moving average */
  const float* bn_weight; const float* bn_bias; float* bn_running_mean; float* bn_running_var; int64_t* bn_num_batches_tracked;
  int32_t use_gn, gn_groups;
  float gn_eps;
  const float* gn_weight; const float* gn_bias;
  float* stats_nc;                  /* [n][c][2]   sum y, sum y*y per (image, channel): written by the kernel that produces y */
  float* bn_sums;                   /* [16][c][2]  the same summed over the batch, in 16 shards (image %% 16): accumulated by that kernel; training-mode BatchNorm only */
  float* s_nc;                      /* [n][c][2]   backward: sum dz, sum dz*y: written by the kernel that produces dz (diagnostic; may be NULL) */
  float* u_sums;                    /* [16][c][4]  backward: BatchNorm adjoint sums and GroupNorm parameter gradients, sharded: accumulated by that kernel */
  float* coef_nc;                   /* [n][c][8]   A, B, BatchNorm alpha / beta / mean / rstd, GroupNorm mean / rstd per (image, channel): written by the FIRST
                                     *             forward consumer of y, loaded by the backward kernels; NULL when no backward will follow */
  float* d_nc;                      /* [n][c][4]   backward: GroupNorm adjoint coefficients D1, D2, D3 per (image, channel): written by the kernel that produces dz */
} dcv_sc_norm;
/* 1 iff the shape is served: bf16, square 3x3 or 5x5 filter, stride 1, "same" padding, w in {16, 32, 64}, c in {1..4, 16}, k even in {2, 4, 16}
 * (5x5: c <= 4 and k <= 4). Everything else stays on dcv_conv2d_* + dcv_norm_*. */
int dcv_sc_conv_supported(const dcv_conv_shape* shape, int dtype);
/* floats of one buffer of a dcv_sc_norm: which = 0 stats_nc, 1 bn_sums, 2 s_nc, 3 u_sums, 4 coef_nc, 5 d_nc */
size_t dcv_sc_norm_floats(int n, int c, int which);
/* y = act(conv(z, w) + bias) with z = x (x_norm NULL / disabled) or the normalised raw output x (x_norm enabled: applied on load; when `update_running`
 * != 0 this launch also performs the running-statistics update of x_norm's BatchNorm — exactly one consumer launch per step must). y_norm enabled:
 * the sums of y are produced for its consumers. w: [K][R][S][C] bf16. */
int dcv_sc_conv_fwd(const dcv_conv_shape* shape, const void* x, const dcv_sc_norm* x_norm, int update_running, const void* w, const float* bias, int act, float slope,
                    void* y, const dcv_sc_norm* y_norm, void* stream);
/* dw[k][r][s][c] += sum dy * z, dbias[k] += sum dy (fp32, accumulated: zero them first), with dy = act'(y)*(P*dz + Q*y + R) (y_norm enabled; its s_nc / u_sums
 * must be complete) or act'(y)*dz. Also writes the BatchNorm / GroupNorm parameter gradients of y_norm (any may be NULL; overwritten). */
int dcv_sc_conv_wgrad(const dcv_conv_shape* shape, const void* x, const dcv_sc_norm* x_norm, const void* dz, const void* y, const dcv_sc_norm* y_norm, int act, float slope,
                      float* dw, float* dbias, float* d_bn_weight, float* d_bn_bias, float* d_gn_weight, float* d_gn_bias, void* stream);
/* dx = data gradient (bf16). x_norm enabled: dx is the gradient w.r.t. the NORMALISED input and x_norm's s_nc / u_sums are produced from it and the
 * producer's raw output `x_raw`. */
int dcv_sc_conv_dgrad(const dcv_conv_shape* shape, const void* dz, const void* y, const dcv_sc_norm* y_norm, int act, float slope, const void* w, void* dx,
                      const void* x_raw, const dcv_sc_norm* x_norm, void* stream);
/* Both gradients of a block in ONE launch (the dy = act'(y)*(P*dz + Q*y + R) tile is staged once and feeds the data-gradient and the weight-gradient
 * GEMMs): arguments as dcv_sc_conv_wgrad + dcv_sc_conv_dgrad (x is the layer input — raw when x_norm is enabled; dw / dbias accumulated, dx written).
 * Served (dcv_sc_conv_bwd_supported): 4 -> 4 channels 5x5, 4 -> 16 and 16 -> 16 channels 3x3 with whole-vector channel counts. */
int dcv_sc_conv_bwd_supported(const dcv_conv_shape* shape, int dtype);
int dcv_sc_conv_bwd(const dcv_conv_shape* shape, const void* x, const dcv_sc_norm* x_norm, const void* dz, const void* y, const dcv_sc_norm* y_norm, int act, float slope, const void* w,
                    void* dx, float* dw, float* dbias, float* d_bn_weight, float* d_bn_bias, float* d_gn_weight, float* d_gn_bias, void* stream);
/* z[n][h/pool][w/pool][c] = A*avgpool(y) + B: materialises a pending normalisation (pool = 1) or fuses it with the pool x pool average pooling that follows
 * (meta/submodule_creators.py:163-176); and its backward: dz (full resolution, written) + norm's s_nc / u_sums from dzp. */
int dcv_sc_affine_pool_fwd(const void* y, const dcv_sc_norm* norm, int update_running, void* z, int n, int h, int w, int c, int pool, void* stream);
int dcv_sc_affine_pool_bwd(const void* dzp, const void* y, const dcv_sc_norm* norm, void* dz, int n, int h, int w, int c, int pool, void* stream);

/* ---- dropout and stand-alone activations (csrc/elementwise.cu) ---
 * Reference block orders (meta/nn.py:553): post-activation `[Dropout] -> op -> act -> norms`, pre-activation `[Dropout] -> norms -> act -> op`.
 * y = x * keep / (1 - p) with keep ~ Bernoulli(1 - p) per element (meta/nn.py:535-541 -> torch.nn.Dropout). keep is a counter-based function
 * (Philox4x32-10) of (seed, *call_dev, element index): nothing is stored, the backward pass calls the same entry point on dy with call_dev = the value
 * the forward call saved in *saved_call_dev (may be NULL). Advance *call_dev with dcv_counter_add after each forward call (graph-replay safe).
 * mask_out (may be NULL): keep as one uint8 per element, for tests that hand the same mask to the CPU oracle. x, y 16-byte aligned; y may alias x. */
int dcv_dropout(const void* x, void* y, uint8_t* mask_out, size_t count, float p, uint64_t seed, const int32_t* call_dev, int32_t* saved_call_dev, int dtype, void* stream);
/* y = act(x); dx = dy * act'(.) with the derivative taken from the activation OUTPUT act_out (DCV_ACT_*). Buffers 16-byte aligned; in place allowed. */
int dcv_activation_fwd(const void* x, void* y, size_t count, int act, float slope, int dtype, void* stream);
int dcv_activation_bwd(const void* dy, const void* act_out, void* dx, size_t count, int act, float slope, int dtype, void* stream);

/* ---- normalisation (BatchNorm2d / GroupNorm after the activation: meta/nn.py:448-516,553; parameters.yml:10,82) ---
 * Any {BatchNorm, GroupNorm} stack applied to y collapses to one affine per (image, channel): z = A[n][c]*y + B[n][c],
 * with A, B closed-form in sum(y), sum(y*y) per (image, channel). */
typedef struct dcv_norm_params {
  int32_t n, c, hw;
  int32_t use_bn, bn_training;       /* bn_training: batch statistics (and running-stat update); else running stats */
  float bn_eps, bn_momentum;         /* momentum < 0: cumulative moving average using num_batches_tracked */
  const float* bn_weight; const float* bn_bias;      /* [c] or NULL (affine=False) */
  float* bn_running_mean; float* bn_running_var;     /* [c] or NULL (track_running_stats=False) */
  int64_t* bn_num_batches_tracked;                   /* device scalar or NULL */
  int32_t use_gn, gn_groups;
  float gn_eps;
  const float* gn_weight; const float* gn_bias;      /* [c] or NULL */
} dcv_norm_params;

/* sum(y), sum(y*y) per (image, channel) -> stats_nc[n][c][2] (overwritten). For tensors whose producer did not fuse it. */
int dcv_norm_stats(const void* y, float* stats_nc, int n, int hw, int c, int dtype, int acc_prezeroed, void* stream);
/* stats_nc -> ab_nc[n][c][2] (A, B) and `saved` (fp32, dcv_norm_saved_floats(n,c,groups) floats: BN mean/rstd [c][2],
 * BN alpha/beta [c][2], GN mean/rstd [n][groups][2]); updates the running statistics when bn_training. */
size_t dcv_norm_saved_floats(int n, int c, int groups);
int dcv_norm_fwd_finalize(const dcv_norm_params* prm, const float* stats_nc, float* ab_nc, float* saved, void* stream);
/* z = A*y + B. */
int dcv_norm_apply_fwd(const void* y, const float* ab_nc, void* z, int n, int hw, int c, int dtype, void* stream);
/* The same pass fused with what follows the block in the architecture (meta/base_module.py decides):
 *   _add:  z = A*y + B + other — a `residual_link` (reduction 'sum', reference meta/submodule_creators.py:272-332) whose first operand is the block output;
 *   _pool: zp[n][h/2][w/2][c] = A*avgpool2x2(y) + B — an `avg_pooling` with kernel = stride = 2 (reference :163-176); the full-resolution normalised
 *          tensor is never written. Backward of both needs no kernel of its own: d/dy is what the block's backward already computes from dz
 *          (dz = the link's incoming gradient; dz = dcv_avgpool2d_bwd of the pooled gradient). */
int dcv_norm_apply_add_fwd(const void* y, const float* ab_nc, const void* other, void* z, int n, int hw, int c, int dtype, void* stream);
/* BatchNorm-only block, batch handed over as ONE image of rows x w pixels (rows = N*H): dcv_norm_fwd_finalize + one of the three apply passes in one launch.
 * `stats_c` = [c][2] channel totals {sum y, sum y^2} (DCV_STATS_CHANNEL_TOTALS); the kernel computes alpha / beta itself, writes `saved` (4*c floats: what
 * the backward reads) and updates the running statistics (momentum >= 0; momentum = None goes through dcv_norm_fwd_finalize). other != NULL: + residual
 * sum; pool != 0: 2x2 / stride-2 average pooling of the result. c a whole number of 16-byte vectors. */
int dcv_bn_apply_fold_fwd(const void* y, const void* other, int pool, void* out, const float* stats_c, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float eps, float momentum, int bn_training, float* saved, int rows, int w, int c, int dtype, void* stream);
int dcv_norm_apply_pool_fwd(const void* y, const float* ab_nc, void* zp, int n, int h, int w, int c, int dtype, void* stream);
/* s_nc[n][c][2] = { sum(dz), sum(dz*y) } (overwritten). */
int dcv_norm_bwd_reduce(const void* dz, const void* y, float* s_nc, int n, int hw, int c, int dtype, int acc_prezeroed, void* stream);
/* -> pqr_nc[n][c][3] with dy_pre_activation = P*dz + Q*y + R, plus parameter gradients (any may be NULL; overwritten).
 * `saved` is the forward's buffer; its trailing scratch region is written. */
int dcv_norm_bwd_finalize(const dcv_norm_params* prm, const float* stats_nc, const float* s_nc, float* saved, float* pqr_nc,
                          float* d_bn_weight, float* d_bn_bias, float* d_gn_weight, float* d_gn_bias, void* stream);
/* dy = act'(y) * (P*dz + Q*y + R) (pqr_nc == NULL: P=1,Q=R=0); if dbias_c != NULL accumulates sum over (n,hw) of dy into
 * dbias_c[c] (fp32, overwritten). `y` is the activation OUTPUT (ReLU / LeakyReLU / Sigmoid / none are recoverable
 * from it). */
int dcv_act_norm_bwd_apply(const void* dz, const void* y, const float* pqr_nc, void* dy, float* dbias_c, int act, float slope,
                           int n, int hw, int c, int dtype, int acc_prezeroed, void* stream);
/* The two backward passes of a block whose normalised output went straight into a 2x2 / stride-2 average pooling (dcv_norm_apply_pool_fwd): `dzp` is the
 * POOLED gradient [n][h/2][w/2][c]; the gradient of a pixel is a quarter of its pooled pixel's, read in place — the full-resolution dz is never
 * written. c a whole number of 16-byte vectors, h and w even (dcv_norm_bwd_pooled_supported). Bit-identical to dcv_avgpool2d_bwd followed by the
 * plain passes. */
/* BatchNorm-only block, batch handed over as ONE image of rows x w pixels (rows = N*H): dcv_norm_bwd_finalize + dcv_act_norm_bwd_apply[_pooled] in one
 * launch. `s_c` = [c][2] channel totals {sum dz, sum dz*y} of dcv_norm_bwd_reduce[_pooled] (DCV_STATS_CHANNEL_TOTALS), `saved` = what dcv_norm_fwd_finalize
 * wrote for n = 1; P, Q, R are computed in the kernel's prologue, d_bn_weight / d_bn_bias (may be NULL) written by it. `pooled`: dz is the pooled gradient
 * [rows/2][w/2][c]. c a whole number of 16-byte vectors. */
int dcv_act_bn_bwd_apply_fold(const void* dz, int pooled, const void* y, const float* s_c, const float* saved, int bn_training, float* d_bn_weight, float* d_bn_bias,
                              void* dy, float* dbias_c, int act, float slope, int rows, int w, int c, int dtype, int acc_prezeroed, void* stream);
int dcv_norm_bwd_pooled_supported(int n, int h, int w, int c, int dtype);
int dcv_norm_bwd_reduce_pooled(const void* dzp, const void* y, float* s_nc, int n, int h, int w, int c, int dtype, int acc_prezeroed, void* stream);
int dcv_act_norm_bwd_apply_pooled(const void* dzp, const void* y, const float* pqr_nc, void* dy, float* dbias_c, int act, float slope,
                                  int n, int h, int w, int c, int dtype, int acc_prezeroed, void* stream);

/* ---- pooling / links (meta/submodule_creators.py:163-176, 272-332; meta/nn.py:416, 665-676) --------------------- */
/* AvgPool2d(kernel, stride), no padding, floor output size. */
int dcv_avgpool2d_fwd(const void* x, void* y, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int dtype, void* stream);
int dcv_avgpool2d_bwd(const void* dy, void* dx, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int dtype, void* stream);
/* out = alpha*a + beta*b (b may be NULL), elementwise over `count` values: residual sum / mean links. */
int dcv_axpby(const void* a, const void* b, void* out, float alpha, float beta, size_t count, int dtype, void* stream);
/* dst[pix][c_off : c_off+c_src] = src[pix][:] (concat = write into a channel slice) and the reverse slice read. */
int dcv_copy_channels_in(const void* src, void* dst, size_t pixels, int c_src, int c_dst, int c_off, int dtype, void* stream);
int dcv_copy_channels_out(const void* src, void* dst, size_t pixels, int c_src, int c_off, int c_dst, int dtype, void* stream);
/* `dense_link` in one launch (reference meta/submodule_creators.py:272-332 with reduction 'concat'): out[n][h][w][:] = the sources' channels one after the
 * other. A source with pool == 2 is an N x 2h x 2w x channels tensor and contributes its 2x2 averages — the reference's `F.interpolate(mode='bilinear',
 * align_corners=False)` at an exact 2x reduction (meta/nn.py:665-676). `_bwd` scatters dout back: grads[i].ptr receives source i's gradient (NULL: not
 * needed), the pooled ones dout/4 at each of the four positions. At most DCV_LINK_MAX_SOURCES sources. */
#define DCV_LINK_MAX_SOURCES 8
typedef struct dcv_link_source {
  void* ptr;
  int32_t channels;
  int32_t pool;
} dcv_link_source;
int dcv_link_concat_fwd(const dcv_link_source* sources, int count, void* out, int n, int h, int w, int dtype, void* stream);
int dcv_link_concat_bwd(const void* dout, const dcv_link_source* grads, int count, int n, int h, int w, int dtype, void* stream);
/* F.interpolate(mode='bilinear', align_corners) forward, and its adjoint written to fp32 dx (overwritten). */
int dcv_bilinear_fwd(const void* x, void* y, int n, int h, int w, int c, int oh, int ow, int align_corners, int dtype, void* stream);
int dcv_bilinear_bwd(const void* dy, float* dx_f32, int n, int h, int w, int c, int oh, int ow, int align_corners, int dtype, void* stream);

/* ---- fully connected head (torch.nn.Linear built at meta/submodule_creators.py:268-269) -------------------------- */
/* y[m][n] = act(sum_k x[m][k]*w[n][k] + bias[n]); w, bias fp32.
 * x_nhwc_channels > 0: `torch.nn.Flatten` (conf/base/parameters.yml:87) is fused in — x is the NHWC image tensor [m][k / C pixels][C channels] itself and
 * weight column f = c * (k / C) + p (the (C, H, W) feature order of the logical N x C x H x W tensor) multiplies element p * C + c of the row; dx comes
 * back in the same NHWC order. Served for the skinny head only: dcv_linear_flatten_fused(n, k) != 0 (n <= 32); pass 0 otherwise. */
int dcv_linear_flatten_fused(int n, int k);
int dcv_linear_fwd(const void* x, const float* w, const float* bias, void* y, int m, int n, int k, int act, float slope,
                   int x_dtype, int y_dtype, int x_nhwc_channels, void* stream);
/* dpre = act'(y)*dy; dx = dpre @ w (x_dtype, may be NULL); dw = dpre^T @ x (may be NULL), db = sum_m dpre (fp32, overwritten).
 * dpre_ws: fp32 [m][n] scratch. */
int dcv_linear_bwd(const void* x, const float* w, const void* y, const void* dy, void* dx, float* dw, float* db, float* dpre_ws,
                   int m, int n, int k, int act, float slope, int x_dtype, int y_dtype, int acc_prezeroed, int x_nhwc_channels, void* stream);

/* ---- data-parallel gradient exchange over peer memory (meta/ignite_training.py:373-390: DistributedDataParallel's gradient averaging) ----
 * One-shot SUM all-reduce of the float slice [offset, offset + count) of `grads` (this rank's flat gradient buffer, ordinary device memory), in place, for
 * small gradient buckets (latency-bound in a collective library): every rank pushes its slice into every peer's receive area, then adds the W slices in rank
 * order (bit-identical results on all ranks). peer_recv_dev / peer_flags_dev: DEVICE arrays of `world` pointers — entry r addresses rank r's receive area
 * (float [2][world][recv_stride], element i of the flat buffer at index i) / flag area (dcv_peer_flag_words() zero-initialised uint32), both in symmetric
 * memory mapped into this process. Every rank calls it with the same offset / count / slot in the same order (slot < DCV_PEER_MAX_SLOTS: the flag words
 * and epoch counter of one bucket). state_dev: 2 * DCV_PEER_MAX_SLOTS zero-initialised uint32 of this rank. count <= dcv_peer_max_floats() (the launch
 * must be co-resident: its CTAs wait for the peers). CUDA-graph replayable. */
#define DCV_PEER_MAX_WORLD 8
#define DCV_PEER_MAX_CTAS 32
#define DCV_PEER_MAX_SLOTS 16
size_t dcv_peer_flag_words(void);
size_t dcv_peer_max_floats(void);
int dcv_peer_allreduce_sum(float* grads, float* const* peer_recv_dev, uint32_t* const* peer_flags_dev, int rank, int world, size_t recv_stride, size_t offset, size_t count, int slot,
                           uint32_t* state_dev, void* stream);

/* ---- loss / optimiser (classification/image.py:70-71) ------------------------------------------------------------ */
/* CrossEntropyLoss(reduction='mean') over logits[m][n] (fp32) and int64 targets: loss[0] and dlogits = (softmax - onehot)/m_valid. Rows whose target is
 * -100 (torch's default ignore_index) are skipped and the mean runs over the others; any other target outside [0, n) makes the loss NaN. */
int dcv_softmax_ce(const float* logits, const int64_t* target, float* loss, float* dlogits, int m, int n, void* stream);
/* Evaluation metrics of a classifier (the `create_supervised_evaluator` metrics of meta/ignite_training.py:288-307: loss + accuracy), ACCUMULATED into
 * acc3 (device float[3], zeroed by the caller before the first batch): acc3[0] += sum of per-row cross entropies, acc3[1] += rows whose argmax (lowest
 * index on ties) equals the target, acc3[2] += rows counted (targets equal to -100 are skipped). */
int dcv_classification_metrics(const float* logits, const int64_t* target, float* acc3, int m, int n, void* stream);
/* dst[i] = src[i] * (*scale_dev): chain rule through the loss with the upstream gradient left on the device (no host sync). */
int dcv_scale_by_device_scalar(const float* src, const float* scale_dev, float* dst, size_t count, void* stream);
/* AdamW over a flat fp32 buffer (torch.optim.AdamW semantics, amsgrad=False). `step_dev` is a device int32 holding the
 * 1-based step number (advance it with dcv_counter_add before the call: graph-replay safe); grads are multiplied by
 * grad_scale first (1/world_size after a sum all-reduce). lr is read from the device scalar `lr_dev`. */
int dcv_counter_add(int32_t* counter_dev, int32_t delta, void* stream);
int dcv_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t count, const float* lr_dev,
                   float beta1, float beta2, float eps, float weight_decay, float grad_scale, const int32_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPCV_B200_H_ */
