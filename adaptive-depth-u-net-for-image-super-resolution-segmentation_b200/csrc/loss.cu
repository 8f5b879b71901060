// Head and loss kernels: ClippedResidualAdd, SR losses (+PSNR metric), BCE+Dice,
// softmax + categorical cross-entropy.  All HBM-bound; reductions are warp
// shuffles -> one shared atomic per warp -> one global atomic per block.
//
// Replaces shared/custom_layers.py:136-139 (ClippedResidualAdd),
// Super_resolution/code/train_adaptive_unet.py:308-348 (charbonnier / l1 / mse, psnr),
// Segmenation/code/train_adaptive_unet.py:258-304 (dice, iou, BCE hybrids) and the
// softmax head of Segmenation/code/unet_vinillia.py:89-90.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ void pix_decode(long long p, int H, int W, int& n, int& h, int& w) {
  w = (int)(p % W); p /= W;
  h = (int)(p % H);
  n = (int)(p / H);
}

__device__ __forceinline__ void block_atomic_add(float v, float* dst) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
}

inline int grid_for(long long items) {
  long long b = (items + NT - 1) / NT;
  long long cap = 8LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- ClippedResidualAdd ---------------------------------------------------------
template <typename TI, typename TR, typename TY>
__global__ void __launch_bounds__(NT)
clipadd_fwd_kernel(TView inp, TView res, TView y, long long total) {
  pdl_sync();
  const TI* ip = reinterpret_cast<const TI*>(inp.data);
  const TR* rp = reinterpret_cast<const TR*>(res.data);
  TY* yp = reinterpret_cast<TY*>(y.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % y.c), n, h, w;
    pix_decode(i / y.c, y.h, y.w, n, h, w);
    float s = ldf(ip + pix_offset(inp, n, h, w) + c) + ldf(rp + pix_offset(res, n, h, w) + c);
    stf(yp + pix_offset(y, n, h, w) + c, fminf(fmaxf(s, 0.f), 1.f));
  }
}

template <typename TI, typename TR>
__global__ void __launch_bounds__(NT)
clipadd_bwd_kernel(TView inp, TView res, TView dy, TView dres, long long total) {
  pdl_sync();
  const TI* ip = reinterpret_cast<const TI*>(inp.data);
  const TR* rp = reinterpret_cast<const TR*>(res.data);
  const TR* dp = reinterpret_cast<const TR*>(dy.data);
  TR* op = reinterpret_cast<TR*>(dres.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % dy.c), n, h, w;
    pix_decode(i / dy.c, dy.h, dy.w, n, h, w);
    float s = ldf(ip + pix_offset(inp, n, h, w) + c) + ldf(rp + pix_offset(res, n, h, w) + c);
    float g = (s >= 0.f && s <= 1.f) ? ldf(dp + pix_offset(dy, n, h, w) + c) : 0.f;  // clip_by_value: closed interval
    stf(op + pix_offset(dres, n, h, w) + c, g);
  }
}

// ---- SR losses ---------------------------------------------------------------------
// grid = (blocks per image, n).  ws[0] += sum of per-element loss, ws[2+n] += squared error of clip(pred)
template <typename TP, typename TT>
__global__ void __launch_bounds__(NT)
sr_loss_kernel(TView pred, TView tgt, int kind, float eps, float gscale, float* __restrict__ ws, TView dpred,
               long long per_img) {
  pdl_sync();
  const int n = blockIdx.y;
  const TP* pp = reinterpret_cast<const TP*>(pred.data);
  const TT* tp = reinterpret_cast<const TT*>(tgt.data);
  TP* gp = reinterpret_cast<TP*>(dpred.data);
  float lsum = 0.f, sq = 0.f;
  const float eps2 = eps * eps;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < per_img; i += (long long)gridDim.x * NT) {
    int c = (int)(i % pred.c);
    long long p = i / pred.c;
    int w = (int)(p % pred.w), h = (int)(p / pred.w);
    float pv = ldf(pp + pix_offset(pred, n, h, w) + c);
    float tv = ldf(tp + pix_offset(tgt, n, h, w) + c);
    float d = tv - pv, g;
    if (kind == B200_LOSS_CHARBONNIER) {
      float r = sqrtf(d * d + eps2);
      lsum += r;
      g = -d / r;
    } else if (kind == B200_LOSS_L1) {
      lsum += fabsf(d);
      g = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
    } else {
      lsum += d * d;
      g = -2.f * d;
    }
    float pc = fminf(fmaxf(pv, 0.f), 1.f);
    sq += (tv - pc) * (tv - pc);
    if (gp) stf(gp + pix_offset(dpred, n, h, w) + c, g * gscale);
  }
  __shared__ float s[2];
  if (threadIdx.x < 2) s[threadIdx.x] = 0.f;
  __syncthreads();
  block_atomic_add(lsum, &s[0]);
  block_atomic_add(sq, &s[1]);
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(ws, s[0]); atomicAdd(ws + 2 + n, s[1]); }
}

__global__ void sr_loss_finalize_kernel(const float* __restrict__ ws, int n, float total, float per_img, float* out) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  out[0] = ws[0] / total;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += 10.f * log10f(1.f / (ws[2 + i] / per_img));
  out[1] = acc / (float)n;
}

// ---- BCE + Dice ----------------------------------------------------------------------
// ws layout: [0] bce sum, [1+3n] inter, [2+3n] union(sum y+p), [3+3n] unused
template <typename TP, typename TT>
__global__ void __launch_bounds__(NT)
bce_dice_reduce_kernel(TView pred, TView tgt, float* __restrict__ ws, long long per_img) {
  pdl_sync();
  const int n = blockIdx.y;
  const TP* pp = reinterpret_cast<const TP*>(pred.data);
  const TT* tp = reinterpret_cast<const TT*>(tgt.data);
  float bce = 0.f, inter = 0.f, uni = 0.f;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < per_img; i += (long long)gridDim.x * NT) {
    int c = (int)(i % pred.c);
    long long p = i / pred.c;
    int w = (int)(p % pred.w), h = (int)(p / pred.w);
    float pv = ldf(pp + pix_offset(pred, n, h, w) + c);
    float y = ldf(tp + pix_offset(tgt, n, h, w) + c);
    float pc = fminf(fmaxf(pv, 1e-7f), 1.f - 1e-7f);
    bce += -(y * logf(pc) + (1.f - y) * logf(1.f - pc));
    inter += y * pc;
    uni += y + pc;
  }
  __shared__ float s[3];
  if (threadIdx.x < 3) s[threadIdx.x] = 0.f;
  __syncthreads();
  block_atomic_add(bce, &s[0]);
  block_atomic_add(inter, &s[1]);
  block_atomic_add(uni, &s[2]);
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(ws, s[0]); atomicAdd(ws + 1 + 3 * n, s[1]); atomicAdd(ws + 2 + 3 * n, s[2]); }
}

__global__ void bce_dice_finalize_kernel(const float* __restrict__ ws, int n, float total, float bw, float dw,
                                         float* out) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float smooth = 1e-6f;
  float dice = 0.f, iou = 0.f, isum = 0.f, usum = 0.f;
  for (int i = 0; i < n; ++i) {
    float I = ws[1 + 3 * i], U = ws[2 + 3 * i];
    dice += (2.f * I + smooth) / (U + smooth);
    iou += (I + smooth) / (U - I + smooth);
    isum += I; usum += U;
  }
  dice /= (float)n; iou /= (float)n;
  float bce = ws[0] / total;
  out[0] = bw * bce + dw * (1.f - dice);
  out[1] = bce; out[2] = dice; out[3] = iou;
  out[4] = (2.f * isum + smooth) / (usum + smooth);   // unet_vinillia.py:94-99: ONE ratio over the whole batch
}

template <typename TP, typename TT>
__global__ void __launch_bounds__(NT)
bce_dice_grad_kernel(TView pred, TView tgt, const float* __restrict__ ws, float bw, float dw, float gscale,
                     float total, TView dpred, long long per_img) {
  pdl_sync();
  const int n = blockIdx.y;
  const TP* pp = reinterpret_cast<const TP*>(pred.data);
  const TT* tp = reinterpret_cast<const TT*>(tgt.data);
  TP* gp = reinterpret_cast<TP*>(dpred.data);
  const float smooth = 1e-6f;
  const float I = ws[1 + 3 * n], U = ws[2 + 3 * n];
  const float den = U + smooth, num = 2.f * I + smooth;
  const float nimg = (float)gridDim.y;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < per_img; i += (long long)gridDim.x * NT) {
    int c = (int)(i % pred.c);
    long long p = i / pred.c;
    int w = (int)(p % pred.w), h = (int)(p / pred.w);
    float pv = ldf(pp + pix_offset(pred, n, h, w) + c);
    float y = ldf(tp + pix_offset(tgt, n, h, w) + c);
    float g = 0.f;
    if (pv >= 1e-7f && pv <= 1.f - 1e-7f) {
      float dbce = -(y / pv - (1.f - y) / (1.f - pv)) / total;
      float ddice = (2.f * y * den - num) / (den * den) / nimg;
      g = bw * dbce - dw * ddice;
    }
    stf(gp + pix_offset(dpred, n, h, w) + c, g * gscale);
  }
}

// ---- BinaryAccuracy / Precision / Recall counters (keras metrics of the baseline trainer, unet_vinillia.py:266-270) ----
// counts += {true positives, false positives, false negatives, correct}: prediction positive = pred > threshold,
// label positive = target != 0 (keras casts y_true to bool); "correct" compares the thresholded prediction with the target.
template <typename TP, typename TT>
__global__ void __launch_bounds__(NT)
binary_confusion_kernel(TView pred, TView tgt, float threshold, float* __restrict__ counts, long long total) {
  pdl_sync();
  const TP* pp = reinterpret_cast<const TP*>(pred.data);
  const TT* tp = reinterpret_cast<const TT*>(tgt.data);
  float c_tp = 0.f, c_fp = 0.f, c_fn = 0.f, c_ok = 0.f;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % pred.c), n, h, w;
    pix_decode(i / pred.c, pred.h, pred.w, n, h, w);
    const float pv = ldf(pp + pix_offset(pred, n, h, w) + c);
    const float y = ldf(tp + pix_offset(tgt, n, h, w) + c);
    const bool pos = pv > threshold, lab = y != 0.f;
    c_tp += (pos && lab) ? 1.f : 0.f;
    c_fp += (pos && !lab) ? 1.f : 0.f;
    c_fn += (!pos && lab) ? 1.f : 0.f;
    c_ok += ((pos ? 1.f : 0.f) == y) ? 1.f : 0.f;
  }
  __shared__ float s[4];
  if (threadIdx.x < 4) s[threadIdx.x] = 0.f;
  __syncthreads();
  block_atomic_add(c_tp, &s[0]); block_atomic_add(c_fp, &s[1]); block_atomic_add(c_fn, &s[2]); block_atomic_add(c_ok, &s[3]);
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(counts + threadIdx.x, s[threadIdx.x]);   // integer-valued sums: exact up to 2^24 per block
}

// ---- softmax / categorical cross-entropy -----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
softmax_fwd_kernel(TView z, TView p, long long npix) {
  pdl_sync();
  const T* zp = reinterpret_cast<const T*>(z.data);
  T* pp = reinterpret_cast<T*>(p.data);
  const int C = z.c;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < npix; i += (long long)gridDim.x * NT) {
    int n, h, w;
    pix_decode(i, z.h, z.w, n, h, w);
    const T* s = zp + pix_offset(z, n, h, w);
    T* d = pp + pix_offset(p, n, h, w);
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, ldf(s + c));
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += __expf(ldf(s + c) - m);
    float inv = 1.f / sum;
    for (int c = 0; c < C; ++c) stf(d + c, __expf(ldf(s + c) - m) * inv);
  }
}

// loss = mean_pixels -log(clip(p_y / sum p)); dlogits = (p - onehot) / npix where p_y inside the clip range
template <typename T>
__global__ void __launch_bounds__(NT)
softmax_ce_kernel(TView prob, const int* __restrict__ labels, float gscale, float* __restrict__ ws, TView dz,
                  long long npix) {
  pdl_sync();
  const T* pp = reinterpret_cast<const T*>(prob.data);
  T* gp = reinterpret_cast<T*>(dz.data);
  const int C = prob.c;
  float lsum = 0.f;
  const float inv_n = gscale / (float)npix;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < npix; i += (long long)gridDim.x * NT) {
    int n, h, w;
    pix_decode(i, prob.h, prob.w, n, h, w);
    const T* s = pp + pix_offset(prob, n, h, w);
    const int y = labels[i];
    // a label outside [0, C) (e.g. a 255 "ignore" id) has no one-hot row: the pixel contributes no loss and no gradient
    // (it still counts in the mean's denominator, as an all-zero one-hot target does in keras) -- never an out-of-bounds read
    const bool known = y >= 0 && y < C;
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += ldf(s + c);
    float q = known ? ldf(s + y) / sum : 1.f;
    float qc = fminf(fmaxf(q, 1e-7f), 1.f - 1e-7f);
    if (known) lsum += -logf(qc);
    if (gp) {
      const float pass = (known && q >= 1e-7f && q <= 1.f - 1e-7f) ? inv_n : 0.f;
      T* d = gp + pix_offset(dz, n, h, w);
      for (int c = 0; c < C; ++c) stf(d + c, pass * (ldf(s + c) / sum - (c == y ? 1.f : 0.f)));
    }
  }
  __shared__ float sh;
  if (threadIdx.x == 0) sh = 0.f;
  __syncthreads();
  block_atomic_add(lsum, &sh);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(ws, sh);
}

__global__ void mean_finalize_kernel(const float* ws, float total, float* out) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = ws[0] / total;
}

}  // namespace

#define B200_DISPATCH_2(dta, dtb, TA, TB, ...)                                        \
  do {                                                                                \
    if ((dta) == B200_BF16 && (dtb) == B200_BF16) { using TA = __nv_bfloat16; using TB = __nv_bfloat16; __VA_ARGS__ } \
    else if ((dta) == B200_BF16) { using TA = __nv_bfloat16; using TB = float; __VA_ARGS__ }  \
    else if ((dtb) == B200_BF16) { using TA = float; using TB = __nv_bfloat16; __VA_ARGS__ }  \
    else { using TA = float; using TB = float; __VA_ARGS__ }                          \
  } while (0)

int clipadd_fwd(const b200_tensor* inp, const b200_tensor* res, const b200_tensor* y, cudaStream_t st) {
  B200_REQUIRE(same_shape(inp, res) && same_shape(inp, y), B200_ERR_BAD_ARG, "clipadd_fwd: shape mismatch");
  B200_REQUIRE(res->dtype == y->dtype, B200_ERR_BAD_ARG, "clipadd_fwd: residual and output dtypes differ");
  long long total = (long long)y->n * y->h * y->w * y->c;
  TView iv = view_of(inp), rv = view_of(res), yv = view_of(y);
  B200_DISPATCH_2(inp->dtype, res->dtype, TI, TR, {
    launch_pdl(clipadd_fwd_kernel<TI, TR, TR>, grid_for(total), NT, 0, st, iv, rv, yv, total);
  });
  return check_launch("clipadd_fwd_kernel");
}

int clipadd_bwd(const b200_tensor* inp, const b200_tensor* res, const b200_tensor* dy, const b200_tensor* dres,
                cudaStream_t st) {
  B200_REQUIRE(same_shape(inp, res) && same_shape(inp, dy) && same_shape(inp, dres), B200_ERR_BAD_ARG,
               "clipadd_bwd: shape mismatch");
  B200_REQUIRE(res->dtype == dy->dtype && res->dtype == dres->dtype, B200_ERR_BAD_ARG, "clipadd_bwd: dtype mismatch");
  long long total = (long long)dy->n * dy->h * dy->w * dy->c;
  TView iv = view_of(inp), rv = view_of(res), dv = view_of(dy), ov = view_of(dres);
  B200_DISPATCH_2(inp->dtype, res->dtype, TI, TR, {
    launch_pdl(clipadd_bwd_kernel<TI, TR>, grid_for(total), NT, 0, st, iv, rv, dv, ov, total);
  });
  return check_launch("clipadd_bwd_kernel");
}

int sr_loss(const b200_tensor* pred, const b200_tensor* target, int kind, float eps, float grad_scale, float* out,
            const b200_tensor* dpred, float* ws, cudaStream_t st) {
  B200_REQUIRE(same_shape(pred, target), B200_ERR_BAD_ARG, "sr_loss: shape mismatch");
  B200_REQUIRE(kind >= 0 && kind <= 2, B200_ERR_BAD_ARG, "sr_loss: unknown loss kind %d", kind);
  TView pv = view_of(pred), tv = view_of(target), gv = pv;
  gv.data = nullptr;
  if (dpred && dpred->data) {
    B200_REQUIRE(same_shape(pred, dpred) && dpred->dtype == pred->dtype, B200_ERR_BAD_ARG, "sr_loss: dpred mismatch");
    gv = view_of(dpred);
  }
  const long long per_img = (long long)pred->h * pred->w * pred->c;
  const float total = (float)per_img * (float)pred->n;
  cudaMemsetAsync(ws, 0, sizeof(float) * (2 + pred->n), st);
  long long bx = (per_img + NT * 4 - 1) / (NT * 4);
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, pred->n);
  B200_DISPATCH_2(pred->dtype, target->dtype, TP, TT, {
    launch_pdl(sr_loss_kernel<TP, TT>, grid, NT, 0, st, pv, tv, kind, eps, grad_scale / total, ws, gv, per_img);
  });
  launch_pdl(sr_loss_finalize_kernel, 1, 32, 0, st, ws, pred->n, total, (float)per_img, out);
  count_launches(1);
  return check_launch("sr_loss_kernel");
}

int bce_dice_loss(const b200_tensor* pred, const b200_tensor* target, float bw, float dw, float grad_scale,
                  float* out, const b200_tensor* dpred, float* ws, cudaStream_t st) {
  B200_REQUIRE(same_shape(pred, target), B200_ERR_BAD_ARG, "bce_dice_loss: shape mismatch");
  TView pv = view_of(pred), tv = view_of(target);
  const long long per_img = (long long)pred->h * pred->w * pred->c;
  const float total = (float)per_img * (float)pred->n;
  cudaMemsetAsync(ws, 0, sizeof(float) * (1 + 3 * pred->n), st);
  long long bx = (per_img + NT * 4 - 1) / (NT * 4);
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, pred->n);
  B200_DISPATCH_2(pred->dtype, target->dtype, TP, TT, {
    launch_pdl(bce_dice_reduce_kernel<TP, TT>, grid, NT, 0, st, pv, tv, ws, per_img);
    launch_pdl(bce_dice_finalize_kernel, 1, 32, 0, st, ws, pred->n, total, bw, dw, out);
    if (dpred && dpred->data) {
      TView gv = view_of(dpred);
      launch_pdl(bce_dice_grad_kernel<TP, TT>, grid, NT, 0, st, pv, tv, ws, bw, dw, grad_scale, total, gv, per_img);
    }
  });
  count_launches((dpred && dpred->data) ? 2 : 1);
  return check_launch("bce_dice_loss");
}

int binary_confusion(const b200_tensor* pred, const b200_tensor* target, float threshold, float* counts, cudaStream_t st) {
  B200_REQUIRE(same_shape(pred, target), B200_ERR_BAD_ARG, "binary_confusion: shape mismatch");
  const long long total = (long long)pred->n * pred->h * pred->w * pred->c;
  cudaMemsetAsync(counts, 0, sizeof(float) * 4, st);
  TView pv = view_of(pred), tv = view_of(target);
  B200_DISPATCH_2(pred->dtype, target->dtype, TP, TT, {
    launch_pdl(binary_confusion_kernel<TP, TT>, grid_for(total), NT, 0, st, pv, tv, threshold, counts, total);
  });
  return check_launch("binary_confusion_kernel");
}

int softmax_fwd(const b200_tensor* z, const b200_tensor* p, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, p) && z->dtype == p->dtype, B200_ERR_BAD_ARG, "softmax_fwd: shape/dtype mismatch");
  const long long npix = (long long)z->n * z->h * z->w;
  TView zv = view_of(z), pv = view_of(p);
  B200_DISPATCH_DTYPE(z->dtype, T, { launch_pdl(softmax_fwd_kernel<T>, grid_for(npix), NT, 0, st, zv, pv, npix); });
  return check_launch("softmax_fwd_kernel");
}

int softmax_ce_loss(const b200_tensor* prob, const int32_t* labels, float grad_scale, float* out,
                    const b200_tensor* dlogits, float* ws, cudaStream_t st) {
  const long long npix = (long long)prob->n * prob->h * prob->w;
  TView pv = view_of(prob), gv = pv;
  gv.data = nullptr;
  if (dlogits && dlogits->data) {
    B200_REQUIRE(same_shape(prob, dlogits) && prob->dtype == dlogits->dtype, B200_ERR_BAD_ARG,
                 "softmax_ce_loss: dlogits mismatch");
    gv = view_of(dlogits);
  }
  cudaMemsetAsync(ws, 0, sizeof(float), st);
  B200_DISPATCH_DTYPE(prob->dtype, T, {
    launch_pdl(softmax_ce_kernel<T>, grid_for(npix), NT, 0, st, pv, labels, grad_scale, ws, gv, npix);
  });
  launch_pdl(mean_finalize_kernel, 1, 32, 0, st, ws, (float)npix, out);
  count_launches(1);
  return check_launch("softmax_ce_kernel");
}

}  // namespace b200
