/*
 * b200_unet.h -- C ABI of the B200-native adaptive-depth U-Net hot path.
 *
 * The reference (KunalNN/Adaptive-Depth-U-Net-...) has no FFI of its own: its
 * "operator API" for this path is the Keras 3 layer API as used by
 *   Super_resolution/code/train_adaptive_unet.py:200-287 (conv_block, builder),
 *   shared/custom_layers.py:85-139 (ResizeByScale, ResizeToMatch, ClippedResidualAdd),
 *   Segmenation/code/train_adaptive_unet.py:258-362, Segmenation/code/unet_vinillia.py:42-99.
 * Each entry point below replaces the TF/Keras kernel(s) behind one of those call
 * sites (cited per function).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - Plain C.  Every pointer inside a b200_tensor / b200_filter and every
 *     pointer argument not marked "host" is a DEVICE pointer owned by the caller.
 *     The library never allocates device memory and keeps no per-tensor state.
 *   - Activations are NHWC with channel stride 1; strides are in ELEMENTS, so a
 *     channel slice of a wider buffer (concat written in place) is a valid tensor.
 *   - Every function returns 0 on success or a negative b200_status; a message is
 *     available from b200_last_error() (thread local).  Unsupported shapes fail
 *     loudly -- there is no CPU fallback.
 *   - All launches are asynchronous on `stream` (a cudaStream_t passed as void*).
 */
#ifndef B200_UNET_H_
#define B200_UNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b200_status {
  B200_OK = 0,
  B200_ERR_BAD_ARG = -1,
  B200_ERR_UNSUPPORTED = -2,
  B200_ERR_LAUNCH = -3
} b200_status;

typedef enum b200_dtype { B200_F32 = 0, B200_BF16 = 1, B200_U8 = 2 /* images of the patch pipeline only */ } b200_dtype;

/* Epilogue / activation selector. */
typedef enum b200_act {
  B200_ACT_NONE = 0,
  B200_ACT_RELU = 1,
  B200_ACT_SIGMOID = 2
} b200_act;

/* Kernel selection for the convolutions. AUTO picks tcgen05 when the shape is
 * supported and falls back to the SIMT kernel for the rest (Cin = 3, fp32). */
typedef enum b200_algo {
  B200_ALGO_AUTO = 0, B200_ALGO_SIMT = 1, B200_ALGO_TCGEN05 = 2,
  B200_ALGO_TCGEN05_1CTA = 3   /* tcgen05 without CTA pairs (cta_group::1 only): the reference point of the pair kernels */
} b200_algo;

typedef enum b200_sr_loss_kind { B200_LOSS_CHARBONNIER = 0, B200_LOSS_L1 = 1, B200_LOSS_MSE = 2 } b200_sr_loss_kind;

typedef struct b200_tensor {
  void* data;                 /* element (n=0,h=0,w=0,c=0) */
  int32_t n, h, w, c;
  int64_t stride_n, stride_h, stride_w; /* elements; channel stride is 1 */
  int32_t dtype;              /* b200_dtype */
  int32_t reserved;
} b200_tensor;

/* A convolution kernel.
 *   hwio : Keras layout [kh][kw][cin][cout].  Every kernel consumes it as is: the tcgen05 fprop reads
 *          it as an MN-major B operand, the dgrad as a K-major one with the taps reversed.
 *   ohwi : optional repacked copy [kh][kw][cout][cin] (b200_filter_pack); not needed by any kernel,
 *          kept for callers that want the transposed layout.  May be NULL. */
typedef struct b200_filter {
  const void* hwio;
  const void* ohwi;
  int32_t kh, kw, cin, cout;
  int32_t dtype;              /* b200_dtype of both copies */
  int32_t reserved;
} b200_filter;

/* Caller-owned device scratch handed to ONE call (16-byte aligned).  The library keeps no pointer to it after
 * the launches of that call are enqueued; calls that share a buffer must be ordered on one stream.  NULL (or
 * {NULL, 0}) where a call needs none. */
typedef struct b200_scratch {
  void* ptr;
  size_t bytes;
} b200_scratch;

/* ---- library ---------------------------------------------------------- */
const char* b200_version(void);
const char* b200_last_error(void);
/* host out-params: SM count, compute capability. */
int b200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched in the calling process (optionally reset). */
long long b200_launch_count(int reset);

/* ---- convolution (keras Conv2D, stride 1, padding "same") ---------------
 * Replaces L.Conv2D at train_adaptive_unet.py:202,207,259,267; seg :326,329,361;
 * unet_vinillia.py:44,49,90.  y = act(conv(x, f) + bias); bias is fp32 [cout] or NULL. */
int b200_conv2d_fprop(const b200_tensor* x, const b200_filter* f, const float* bias,
                      const b200_tensor* y, int act, int algo, const b200_scratch* ws, void* stream);
/* Conv2D -> LayerNormalization(axis=-1) [-> ReLU] in one call (conv_block, train_adaptive_unet.py:202-204).
 * z = conv(x)+bias in the storage dtype (kept for the backward pass; z->data may be NULL for inference),
 * y = act(LN(z)); mean/rstd fp32 [n*h*w].  Fused into the tcgen05 epilogue when Cout is 64 or 128,
 * otherwise executed as convolution + b200_layernorm_fwd inside the library. */
int b200_conv2d_ln_fprop(const b200_tensor* x, const b200_filter* f, const float* bias, const float* gamma,
                         const float* beta, float eps, int relu, const b200_tensor* z, const b200_tensor* y,
                         float* mean, float* rstd, int algo, const b200_scratch* ws, void* stream);
/* dx (+)= conv_transpose(dy, f)  -- autodiff of the above w.r.t. its input. */
int b200_conv2d_dgrad(const b200_tensor* dy, const b200_filter* f, const b200_tensor* dx,
                      int accumulate, int algo, const b200_scratch* ws, void* stream);
/* dgrad with the backward pass of the LayerNormalization(+ReLU) that PRODUCED the convolution's input fused into its
 * epilogue (autodiff of conv_block's Conv2D <- ReLU <- LayerNormalization chain, train_adaptive_unet.py:202-209): the
 * accumulator tile is dy of the LayerNorm output; with that layer's saved pre-norm activations z, its per-pixel mean /
 * rstd and gamma / beta the kernel writes
 *     dz = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),   g = dy * [relu: y > 0] * gamma,  xhat = (z - mean) * rstd
 * and ADDS sum_pixels(dy_masked * xhat) to dgamma, sum_pixels(dy_masked) to dbeta and sum_pixels(dz) to dbias (the bias
 * gradient of the convolution that produced z); any of the three may be NULL.  dy of the LayerNorm output is never
 * written to memory.  b200_conv2d_dgrad_ln_bwd_supported tells whether the shapes take this kernel (3x3 bf16 filter, 64
 * input channels, <= 128 output channels, plain 16x8 tiles); otherwise call b200_conv2d_dgrad + b200_layernorm_bwd. */
int b200_conv2d_dgrad_ln_bwd_supported(const b200_tensor* dy, const b200_filter* f, const b200_tensor* dz);
int b200_conv2d_dgrad_ln_bwd(const b200_tensor* dy, const b200_filter* f, const b200_tensor* z, const float* mean,
                             const float* rstd, const float* gamma, const float* beta, int relu, const b200_tensor* dz,
                             float* dgamma, float* dbeta, float* dbias, void* stream);
/* Scratch of the split-K path that serves small-spatial layers (images <= 4x4 pixels, the deep U-Net
 * levels): b200_conv2d_workspace returns the bytes fprop (dgrad = 0, x = the input) or dgrad (dgrad = 1,
 * x = dy) of this layer needs in `ws` (0 = the layer does not take that path; `ws` may then be NULL).
 * Which path a layer takes is a function of its shapes only -- a call that passes too little scratch fails
 * with B200_ERR_BAD_ARG, it never runs another algorithm.  There is no process-wide registration: the
 * caller (one plan) passes its buffer with every call. */
size_t b200_conv2d_workspace(const b200_tensor* x, const b200_filter* f, int dgrad);

/* dw_hwio[kh][kw][cin][cout] (fp32) = sum_pixels x (*) dy.  Overwrites dw.
 * `workspace` must hold b200_conv2d_wgrad_workspace() bytes (may be 0). */
size_t b200_conv2d_wgrad_workspace(const b200_tensor* x, const b200_tensor* dy, int kh, int kw, int algo);
int b200_conv2d_wgrad(const b200_tensor* x, const b200_tensor* dy, int kh, int kw, float* dw_hwio,
                      void* workspace, size_t workspace_bytes, int algo, void* stream);
/* dw_hwio += sum_pixels x (*) dy without a workspace: every CTA adds its partial sums straight into dw with
 * vector atomics (fp32 `red.global.add.v4`), which removes the partial slabs and the reduce launch.  dw must hold
 * zeros (or a partial gradient to add to) on entry; the summation order, hence the last bits, vary run to run.
 * tcgen05 shapes only (bf16, Cin and Cout multiples of 64, 3x3 or 1x1); B200_ERR_UNSUPPORTED otherwise. */
int b200_conv2d_wgrad_atomic(const b200_tensor* x, const b200_tensor* dy, int kh, int kw, float* dw_hwio, void* stream);
/* hwio -> ohwi repack (same dtype). */
int b200_filter_pack(const void* hwio, void* ohwi, int kh, int kw, int cin, int cout, int dtype, void* stream);
/* im2col of a narrow 3x3 "same" convolution input (Cin*9 <= 64: the RGB stem, train_adaptive_unet.py:202
 * with Cin=3): xcol[n,h,w, (kh*3+kw)*Cin + c] = x[n,h+kh-1,w+kw-1,c], zero padded to 64 bf16 channels.
 * The stem then runs as a 1x1 convolution of xcol on the tcgen05 kernels: the conv2d fprop (plain or with LayerNorm) and
 * wgrad entry points with kh = kw = 1 and the HWIO kernel zero-padded to [64][Cout]. */
int b200_im2col3x3(const b200_tensor* x, const b200_tensor* xcol, void* stream);

/* ---- transposed convolution (keras Conv2DTranspose(nf, 2, strides=2)) ----
 * unet_vinillia.py:67.  kernel layout [2][2][cout][cin] in `dtype`, fp32 bias. */
int b200_convT2x2_fprop(const b200_tensor* x, const void* kernel, const float* bias, int cout,
                        const b200_tensor* y, const b200_scratch* ws, void* stream);
int b200_convT2x2_dgrad(const b200_tensor* dy, const void* kernel, int cout, const b200_tensor* dx,
                        const b200_scratch* ws, void* stream);
/* scratch bytes of the two calls above (x = the input for fprop, = dy for dgrad; 0 for most shapes). */
size_t b200_convT2x2_workspace(const b200_tensor* x, int cin, int cout, int dgrad);
int b200_convT2x2_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dkernel, float* dbias, void* stream);

/* ---- bias/activation backward -------------------------------------------
 * dz = dy * act'(y); dbias[c] += sum dz (fp32 atomics; caller zeroes; may be NULL). */
int b200_bias_act_bwd(const b200_tensor* dy, const b200_tensor* y, int act, const b200_tensor* dz,
                      float* dbias, void* stream);

/* ---- keras LayerNormalization(axis=-1) [+ Activation("relu")] ------------
 * train_adaptive_unet.py:203-204,208-209.  mean/rstd: fp32 [n*h*w] saved for backward. */
int b200_layernorm_fwd(const b200_tensor* z, const float* gamma, const float* beta, float eps, int relu,
                       const b200_tensor* y, float* mean, float* rstd, void* stream);
/* dz from dy; dgamma/dbeta/dbias (fp32 [c]) are accumulated with atomics (caller
 * zeroes).  dbias is the gradient of the bias of the conv that produced z. */
int b200_layernorm_bwd(const b200_tensor* dy, const b200_tensor* z, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, int relu, const b200_tensor* dz,
                       float* dgamma, float* dbeta, float* dbias, void* stream);

/* ---- keras BatchNormalization() [+ relu], training and inference ---------
 * Segmenation/code/train_adaptive_unet.py:327-328,330-331.
 * stats_ws: fp64 [2*c] scratch, zeroed by the call.  save_mean/save_rstd fp32 [c]. */
int b200_batchnorm_fwd_train(const b200_tensor* z, const float* gamma, const float* beta, float eps,
                             float momentum, int relu, const b200_tensor* y, float* save_mean,
                             float* save_rstd, float* moving_mean, float* moving_var, double* stats_ws,
                             void* stream);
int b200_batchnorm_fwd_infer(const b200_tensor* z, const float* gamma, const float* beta, float eps, int relu,
                             const float* moving_mean, const float* moving_var, const b200_tensor* y,
                             void* stream);
int b200_batchnorm_bwd(const b200_tensor* dy, const b200_tensor* z, const float* save_mean,
                       const float* save_rstd, const float* gamma, const float* beta, int relu,
                       const b200_tensor* dz, float* dgamma, float* dbeta, float* dbias, double* stats_ws,
                       void* stream);
/* Synchronised BatchNorm under data parallelism (BatchNorm models are not sample-independent) -- the three
 * calls above in phases, so the caller can sum the statistics over the ranks in between:
 *   forward : b200_batchnorm_stats(z, NULL, ...)  -> all-reduce stats_ws (2*C doubles: sum z, sum z^2)
 *             -> b200_batchnorm_fwd_apply(..., stats_ws, GLOBAL pixel count)
 *   backward: b200_batchnorm_stats(z, dy, ...)    (also adds the LOCAL sums to dgamma / dbeta, which then join the
 *             ordinary gradient exchange) -> all-reduce stats_ws (sum g, sum g*xhat) -> b200_batchnorm_bwd_apply(...,
 *             GLOBAL pixel count).  With one rank the phases equal b200_batchnorm_fwd_train / b200_batchnorm_bwd. */
int b200_batchnorm_stats(const b200_tensor* z, const b200_tensor* dy, const float* save_mean, const float* save_rstd,
                         const float* gamma, const float* beta, int relu, double* stats_ws, float* dgamma, float* dbeta,
                         void* stream);
int b200_batchnorm_fwd_apply(const b200_tensor* z, const float* gamma, const float* beta, float eps, float momentum,
                             int relu, const b200_tensor* y, float* save_mean, float* save_rstd, float* moving_mean,
                             float* moving_var, const double* stats_ws, double count, void* stream);
int b200_batchnorm_bwd_apply(const b200_tensor* dy, const b200_tensor* z, const float* save_mean, const float* save_rstd,
                             const float* gamma, const float* beta, int relu, const b200_tensor* dz,
                             const double* stats_ws, double count, void* stream);

/* ---- separable linear resampling ----------------------------------------
 * tf.image.resize(bilinear, antialias) of ResizeByScale / ResizeToMatch
 * (shared/custom_layers.py:102,124) and UpSampling2D bilinear (seg :357), forward
 * and exact-transpose backward, as one gather kernel over span tables.
 * Host helpers fill the tables; the caller uploads them.
 *   starts [out]        : first source index of each output's span
 *   weights[out * taps] : span weights, zero padded */
int b200_resize_extent(int extent, float scale);                    /* max(1, ceil(f32(extent)*scale)) */
int b200_resample_taps(int in_size, int out_size, int antialias);    /* span width (host) */
int b200_resample_plan(int in_size, int out_size, int antialias, int32_t* starts /*host*/,
                       float* weights /*host*/, int taps);
/* transpose of a plan: tables indexed by SOURCE index; returns taps needed when
 * t_starts == NULL. */
int b200_resample_plan_transpose(int in_size, int out_size, int taps, const int32_t* starts /*host*/,
                                 const float* weights /*host*/, int32_t* t_starts /*host*/,
                                 float* t_weights /*host*/, int t_taps);
/* y[n,oh,ow,:] (+)= sum_i sum_j wh[oh][i] * ww[ow][j] * x[n, sh[oh]+i, sw[ow]+j, :] */
int b200_resample2d(const b200_tensor* x, const b200_tensor* y, const int32_t* h_starts,
                    const float* h_weights, int h_taps, const int32_t* w_starts, const float* w_weights,
                    int w_taps, int accumulate, void* stream);
/* Optional table preparation for the row-marching kernels (host, in place): shift leading zero weights
 * out of every row and return the effective tap count (rows keep their pitch `taps`; re-pack to that
 * count before uploading); then ask which kernel may walk the ROW axis of the compacted table:
 * 0 = gather only, 1 = "up" (taps <= 3), 2 / 3 = "down" with 4 / 6 accumulator slots. */
int b200_resample_compact(int n_out, int taps, int32_t* starts /*host*/, float* weights /*host*/);
int b200_resample_mode(int n_out, int taps, const int32_t* starts /*host*/);
/* b200_resample2d with the row-axis kernel chosen by `h_mode` (from b200_resample_mode on h_starts);
 * bf16 tensors with C % 8 == 0 take the marching kernels, everything else the gather kernel. */
int b200_resample2d_ex(const b200_tensor* x, const b200_tensor* y, const int32_t* h_starts,
                       const float* h_weights, int h_taps, const int32_t* w_starts, const float* w_weights,
                       int w_taps, int accumulate, int h_mode, void* stream);

/* ---- keras MaxPooling2D(2) -- seg :351, unet_vinillia.py:62 --------------- */
int b200_maxpool2_fwd(const b200_tensor* x, const b200_tensor* y, void* stream);
int b200_maxpool2_bwd(const b200_tensor* x, const b200_tensor* y, const b200_tensor* dy,
                      const b200_tensor* dx, int accumulate, void* stream);

/* ---- ClippedResidualAdd -- shared/custom_layers.py:136-139 ----------------- */
int b200_clipadd_fwd(const b200_tensor* inp, const b200_tensor* res, const b200_tensor* y, void* stream);
int b200_clipadd_bwd(const b200_tensor* inp, const b200_tensor* res, const b200_tensor* dy,
                     const b200_tensor* dres, void* stream);

/* ---- SR losses + PSNR metric -- train_adaptive_unet.py:308-348 -------------
 * out[0] = loss, out[1] = psnr metric (fp32, device).  dpred may have data==NULL
 * (evaluation).  grad_scale multiplies dpred (1/world_size under data parallel).
 * ws: fp32 [2 + n] scratch, zeroed by the call. */
int b200_sr_loss(const b200_tensor* pred, const b200_tensor* target, int kind, float eps, float grad_scale,
                 float* out, const b200_tensor* dpred, float* ws, void* stream);

/* ---- BCE + Dice (+IoU) on probabilities -- seg :258-304 ---------------------
 * out (fp32 [5]): [0]=loss, [1]=bce, [2]=dice (per-sample ratios averaged, seg :258-265), [3]=iou (:272-280),
 * [4]=dice as ONE ratio over the whole batch (unet_vinillia.py:94-99).  ws: fp32 [1 + 3*n], zeroed by the call. */
int b200_bce_dice_loss(const b200_tensor* pred, const b200_tensor* target, float bce_weight,
                       float dice_weight, float grad_scale, float* out, const b200_tensor* dpred, float* ws,
                       void* stream);

/* ---- softmax head + categorical cross-entropy (config C4; extrapolated) ----
 * unet_vinillia.py:89-90 softmax head; keras CategoricalCrossentropy semantics. */
/* keras BinaryAccuracy / Precision / Recall (unet_vinillia.py:266-270) as counters: counts (device fp32[4]) =
 * {true positives, false positives, false negatives, correct} over all elements, prediction positive = pred > threshold,
 * label positive = target != 0.  accuracy = correct / elements, precision = tp / (tp + fp), recall = tp / (tp + fn). */
int b200_binary_confusion(const b200_tensor* pred, const b200_tensor* target, float threshold, float* counts, void* stream);
int b200_softmax_fwd(const b200_tensor* z, const b200_tensor* p, void* stream);
int b200_softmax_ce_loss(const b200_tensor* prob, const int32_t* labels, float grad_scale, float* out,
                         const b200_tensor* dlogits, float* ws, void* stream);

/* ---- Adam -- train_adaptive_unet.py:490 -----------------------------------
 * hyper (device fp32[6]) = {lr, beta1, beta2, eps, 1-beta1, 1-beta2} (the last two evaluated in
 * double on the host and then rounded, as keras does); step (device int32[1]) is the
 * 1-based step count, incremented by b200_adam_advance.  Updates p/m/v in place and
 * writes the compute-dtype shadow copy of p (bf16 or NULL).
 * loss_scale (device fp32[4], or NULL) = {scale, finite steps in a row, found_inf flag, skipped steps}: the
 * state of the keras LossScaleOptimizer (the wrapper `mixed_float16` puts around Adam, train_adaptive_unet.py:471-477).
 * With it, g is divided by `scale`, and a step whose found_inf flag is set changes nothing (p, m, v, step). */
int b200_adam_advance(int32_t* step, const float* loss_scale, void* stream);
int b200_adam_step(float* p, const float* g, float* m, float* v, size_t count, const float* hyper,
                   const int32_t* step, void* shadow_bf16, const float* loss_scale, void* stream);
/* Dynamic loss scaling around a training step, all state on the device (capturable in a CUDA graph):
 *   b200_loss_scale_apply  : data *= scale (the loss gradient, right after the loss kernel)
 *   b200_loss_scale_check  : found_inf |= any non-finite value in g (the flat gradient buffer, after backward)
 *   b200_loss_scale_update : after the optimizer -- found_inf ? (scale = max(scale/2, 1), streak = 0, clear the flag)
 *                            : (streak += 1; streak == growth_interval ? scale *= 2, streak = 0). */
int b200_loss_scale_apply(void* data, int dtype, size_t count, const float* loss_scale, void* stream);
int b200_loss_scale_check(const float* g, size_t count, float* loss_scale, void* stream);
int b200_loss_scale_update(float* loss_scale, float growth_interval, void* stream);

/* ---- patch pipeline on the device -- shared/pipeline.py:79-136, 177-246 -------------
 * random_patches / grid_patches + degrade_image (cv2 INTER_AREA shrink to round(P*scale), INTER_CUBIC
 * enlargement back to P, input clipped to [0,1], output not clipped) on a decoded image that is
 * already resident in HBM.  OpenCV's resize is a separable tap table per axis; the host helpers build
 * the tables OpenCV builds (explicit source indices, border taps clamped), b200_gather2d applies them
 * horizontal-first in fp32.  Patch origins are drawn on the host (numpy Generator, the reference's
 * seed discipline) and uploaded. */
typedef enum b200_cv_interp { B200_CV_INTER_AREA = 0, B200_CV_INTER_CUBIC = 1 } b200_cv_interp;
int b200_cv_resize_taps(int in_size, int out_size, int interp);          /* taps per output index (host) */
int b200_cv_resize_plan(int in_size, int out_size, int interp, int32_t* idx /*host [out*taps]*/,
                        float* weights /*host [out*taps]*/, int taps);
/* hr[i] = image[top_i : top_i+P, left_i : left_i+P, :] as fp32 (uint8 images are scaled by 1/255, as
 * load_rgb_image_full :70-76 does).  image: device, HxWx3 dense, B200_U8 or B200_F32.  origins: device
 * int32 [n][2] = (top, left), clamped into the image by the kernel.  hr: fp32 [n,P,P,3], dense rows. */
int b200_patch_extract(const void* image, int image_dtype, int img_h, int img_w, const int32_t* origins,
                       const b200_tensor* hr, void* stream);
/* y[n,oy,ox,:] = sum_j hw[oy][j] * sum_k ww[ox][k] * f(x[n, hi[oy][j], wi[ox][k], :]); f = clip to [0,1]
 * when clip01, identity otherwise.  fp32, 1 or 3 channels.  Tables: device. */
int b200_gather2d(const b200_tensor* x, const b200_tensor* y, const int32_t* h_idx, const float* h_w, int h_taps,
                  const int32_t* w_idx, const float* w_w, int w_taps, int clip01, void* stream);
/* dst[dst_rows ? dst_rows[r] : r][:] = src[src_rows ? src_rows[r] : r][:] for r < n_rows (fp32 rows of
 * row_elems elements): the shuffle buffer of make_training_patch_dataset :214-246 kept in HBM. */
int b200_copy_rows(const float* src, const int32_t* src_rows, float* dst, const int32_t* dst_rows, int n_rows,
                   long long row_elems, void* stream);

/* ---- evaluation metrics -- train_adaptive_unet.py:144-157, 673-721; evaluate_model.py:94-163 ---
 * b200_luma_pair: prediction clipped to [0,1]; BT.601 luma (65.481 R + 128.553 G + 24.966 B + 16)/255 of
 * prediction (f32/bf16 [n,h,w,3]) and target (f32 [n,h,w,3]), clipped to [0,1]; `shave` border pixels
 * dropped; writes the two fp32 planes [n][h-2s][w-2s] and the per-image sum of squared luma differences
 * sse[n] (zeroed by the call) -> MSE(Y), PSNR(Y). */
int b200_luma_pair(const b200_tensor* pred_rgb, const b200_tensor* hr_rgb, int shave, float* pred_y, float* hr_y,
                   float* sse, void* stream);
/* tf.image.ssim on fp32 images [n][h][w][channels] (h, w >= 11; channels = 1 for the luma planes, 3 for the RGB
 * evaluation of u-net-vinillia.py:222-230), every channel filtered on its own as TF's depthwise filter does:
 * out[n*channels + c][0] = sum of the SSIM map, [1] = sum of its contrast-structure factor over the
 * (h-10)(w-10) valid window positions (zeroed by the call; divide by that count for the per-channel means,
 * average those over the channels for tf.image.ssim). */
int b200_ssim_planes(const float* a, const float* b, int n, int h, int w, int channels, float max_val, float* out,
                     void* stream);
/* 2x2 average pooling between MS-SSIM scales; odd extents padded symmetrically: y is [n][(h+1)/2][(w+1)/2][channels]. */
int b200_avgpool2_planes(const float* x, int n, int h, int w, int channels, float* y, void* stream);

/* ---- utilities -------------------------------------------------------------- */
int b200_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t count, void* stream);
int b200_copy_tensor(const b200_tensor* src, const b200_tensor* dst, void* stream); /* strided, converting */
int b200_scale_inplace(float* p, size_t count, float s, void* stream);

/* ---- debug --------------------------------------------------------------------
 * Single-CTA tcgen05 descriptor probe used by tests/test_umma_probe.py to pin the UMMA
 * shared-memory addressing model the convolution kernels rely on (shifted start addresses
 * and non-1024 stride-byte-offsets under the 128-byte swizzle).
 * a: [a_rows][64] bf16, b: [64][64] bf16, out: [128][64] fp32. */
int b200_debug_umma_probe(const void* a, int a_rows, const void* b, int start_bytes, int sbo_bytes,
                          int lbo_bytes, int mn_major, float* out, void* stream);

/* MMA issue-rate probe: `grid` CTAs each issue iters x 4 tcgen05.mma (M=128, N=n, K=16) on fixed smem
 * operands; cycles[grid] (device int64) receives the elapsed SM cycles per CTA. */
int b200_debug_umma_rate(int n, int iters, int a_stride_bytes, long long* cycles, int grid, void* stream);

/* ---- data-parallel exchange over NVLink peer memory (one node) -----------------------
 * The reference trains on one GPU (Super_resolution/code/train_adaptive_unet.py:622-632 is the step that is sharded);
 * this is the exchange of the batch-sharded step: see csrc/peer.cu for the protocol.
 *   b200_peer_alloc / _free     : zero-filled device memory that can be exported to the other ranks of the node
 *   b200_peer_export / _open / _close : cudaIpc handle (B200_PEER_HANDLE_BYTES bytes) of such a buffer / a peer's mapping
 *   b200_peer_signal            : counter += 1, visible to the peers once everything before it on `stream` is
 *   b200_peer_wait              : stream waits until each of the n_peers (<= 32) counters has reached *own
 *                                 (peer_counters = DEVICE array of pointers); traps after timeout_s (<= 0: 30 s)
 *   b200_peer_pull              : n copy-engine copies peer -> local (host arrays of pointers / devices / sizes)
 *   b200_peer_gather_sum        : staging[j] <- peer_src[j] (count floats each, copy engines), then
 *                                 acc[i] += sum_j staging[j][i] in one kernel (reduce-scatter of my shard) */
#define B200_PEER_HANDLE_BYTES 64
int b200_peer_alloc(void** out, size_t bytes);
int b200_peer_free(void* p);
int b200_peer_export(const void* p, void* handle64);
int b200_peer_open(const void* handle64, void** out);
int b200_peer_close(void* p);
int b200_peer_signal(unsigned long long* counter, void* stream);
int b200_peer_wait(const unsigned long long* const* peer_counters, int n_peers, const unsigned long long* own,
                   double timeout_s, void* stream);
int b200_peer_pull(void* const* dst, const void* const* src, const int* src_device, const size_t* bytes, int n,
                   int my_device, void* stream);
int b200_peer_gather_sum(float* acc, float* staging, const void* const* peer_src, const int* peer_device, int n_peers,
                         size_t count, int my_device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_UNET_H_ */
