// Separable linear resampling (antialiased bilinear resize, bilinear x2 upsample and
// their exact-transpose backward passes) and 2x2 max pooling -- HBM-bound kernels.
//
// Replaces tf.image.resize(..., "bilinear", antialias=True) behind ResizeByScale /
// ResizeToMatch (shared/custom_layers.py:102,124), keras UpSampling2D(2,"bilinear")
// (Segmenation/code/train_adaptive_unet.py:357) and MaxPooling2D(2) (:351;
// unet_vinillia.py:62).  The span rule is TensorFlow's ScaleAndTranslate
// (triangle kernel, float32 arithmetic); it is computed on the host into small
// tables and the device kernel is a pure gather, so forward and backward are the
// same kernel over a table and its transpose.
#include <math.h>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int NT = 256;

// grid = (ceil(OW * C/8 / 256), OH, N): one block per output-row segment, one thread per
// (output pixel, 8-channel chunk).  No per-element divisions by H/W; the row's h-taps are block-uniform.
template <typename T>
__global__ void __launch_bounds__(NT)
resample_vec_kernel(TView x, TView y, const int* __restrict__ hs, const float* __restrict__ hw, int ht,
                    const int* __restrict__ ws, const float* __restrict__ ww, int wt, int accumulate) {
  const int chunks = y.c >> 3;
  const int item = blockIdx.x * NT + threadIdx.x;
  const int ow = item / chunks, j = item - ow * chunks;
  if (ow >= y.w) return;
  const int oh = blockIdx.y, n = blockIdx.z;
  const T* xp = reinterpret_cast<const T*>(x.data) + (long long)n * x.sn + j * 8;
  T* dst = reinterpret_cast<T*>(y.data) + pix_offset(y, n, oh, ow) + j * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int h0 = hs[oh], w0 = ws[ow];
  const float* wrow = ww + ow * wt;
  for (int a = 0; a < ht; ++a) {
    const float wa = hw[oh * ht + a];
    if (wa == 0.f) continue;
    const T* row = xp + (long long)min(h0 + a, x.h - 1) * x.sh;
    for (int b = 0; b < wt; ++b) {
      const float wb = wrow[b];
      if (wb == 0.f) continue;
      float v[8];
      Vec8<T>::load(row + (long long)min(w0 + b, x.w - 1) * x.sw, v);
      const float wgt = wa * wb;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += wgt * v[k];
    }
  }
  if (accumulate) {
    float o[8];
    Vec8<T>::load(dst, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += o[k];
  }
  Vec8<T>::store(dst, acc);
}

template <typename T>
__global__ void __launch_bounds__(NT)
resample_scalar_kernel(TView x, TView y, const int* __restrict__ hs, const float* __restrict__ hw, int ht,
                       const int* __restrict__ ws, const float* __restrict__ ww, int wt, int accumulate,
                       long long total) {
  const T* xp = reinterpret_cast<const T*>(x.data);
  T* yp = reinterpret_cast<T*>(y.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % y.c);
    long long p = i / y.c;
    const int ow = (int)(p % y.w); p /= y.w;
    const int oh = (int)(p % y.h);
    const int n = (int)(p / y.h);
    float acc = 0.f;
    const int h0 = hs[oh], w0 = ws[ow];
    for (int a = 0; a < ht; ++a) {
      const float wa = hw[oh * ht + a];
      if (wa == 0.f) continue;
      const int ih = min(h0 + a, x.h - 1);
      for (int b = 0; b < wt; ++b) {
        const float wb = ww[ow * wt + b];
        if (wb == 0.f) continue;
        const int iw = min(w0 + b, x.w - 1);
        acc += wa * wb * ldf(xp + pix_offset(x, n, ih, iw) + c);
      }
    }
    T* dst = yp + pix_offset(y, n, oh, ow) + c;
    if (accumulate) acc += ldf(dst);
    stf(dst, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
maxpool2_fwd_kernel(TView x, TView y, long long total) {
  const T* xp = reinterpret_cast<const T*>(x.data);
  T* yp = reinterpret_cast<T*>(y.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % y.c);
    long long p = i / y.c;
    const int ow = (int)(p % y.w); p /= y.w;
    const int oh = (int)(p % y.h);
    const int n = (int)(p / y.h);
    const T* s = xp + pix_offset(x, n, 2 * oh, 2 * ow) + c;
    float m = fmaxf(fmaxf(ldf(s), ldf(s + x.sw)), fmaxf(ldf(s + x.sh), ldf(s + x.sh + x.sw)));
    stf(yp + pix_offset(y, n, oh, ow) + c, m);
  }
}

// gradient goes to the first maximum of each window (row-major), as TF's MaxPoolGrad
template <typename T>
__global__ void __launch_bounds__(NT)
maxpool2_bwd_kernel(TView x, TView y, TView dy, TView dx, int accumulate, long long total) {
  const T* xp = reinterpret_cast<const T*>(x.data);
  const T* yp = reinterpret_cast<const T*>(y.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  T* dxp = reinterpret_cast<T*>(dx.data);
  // one thread per INPUT element (covers odd trailing rows/cols, which get zero)
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % x.c);
    long long p = i / x.c;
    const int iw = (int)(p % x.w); p /= x.w;
    const int ih = (int)(p % x.h);
    const int n = (int)(p / x.h);
    const int oh = ih / 2, ow = iw / 2;
    float g = 0.f;
    if (oh < y.h && ow < y.w) {
      const float m = ldf(yp + pix_offset(y, n, oh, ow) + c);
      const T* s = xp + pix_offset(x, n, 2 * oh, 2 * ow) + c;
      int first = 3;
      if (ldf(s) == m) first = 0;
      else if (ldf(s + x.sw) == m) first = 1;
      else if (ldf(s + x.sh) == m) first = 2;
      if (first == (ih - 2 * oh) * 2 + (iw - 2 * ow)) g = ldf(dyp + pix_offset(dy, n, oh, ow) + c);
    }
    T* dst = dxp + pix_offset(dx, n, ih, iw) + c;
    if (accumulate) g += ldf(dst);
    stf(dst, g);
  }
}

inline int grid_for(long long items) {
  long long b = (items + NT - 1) / NT;
  long long cap = 16LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

// ---- host-side span tables (ScaleAndTranslate, triangle kernel) ----------------
int resize_extent(int extent, float scale) {
  float prod = (float)extent * scale;
  int v = (int)ceilf(prod);
  return v < 1 ? 1 : v;
}

int resample_taps(int in_size, int out_size, int antialias) {
  if (in_size == out_size) return 1;
  float scale = (float)out_size / (float)in_size;
  float inv = 1.0f / scale;
  float ks = antialias ? fmaxf(inv, 1.0f) : 1.0f;
  int span = 2 * (int)ceilf(ks) + 1;
  return span < in_size ? span : in_size;
}

int resample_plan(int in_size, int out_size, int antialias, int32_t* starts, float* weights, int taps) {
  if (in_size <= 0 || out_size <= 0) return fail(B200_ERR_BAD_ARG, "resample_plan: bad sizes %d -> %d", in_size, out_size);
  const int need = resample_taps(in_size, out_size, antialias);
  if (taps < need) return fail(B200_ERR_BAD_ARG, "resample_plan: taps %d < required %d", taps, need);
  for (int i = 0; i < out_size * taps; ++i) weights[i] = 0.f;
  if (in_size == out_size) {  // tf.image.resize returns the input when the size is unchanged
    for (int x = 0; x < out_size; ++x) { starts[x] = x; weights[x * taps] = 1.f; }
    return B200_OK;
  }
  const float scale = (float)out_size / (float)in_size;
  const float inv_scale = 1.0f / scale;
  const float ks = antialias ? fmaxf(inv_scale, 1.0f) : 1.0f;
  const float inv_ks = 1.0f / ks;
  std::vector<float> tmp;
  for (int x = 0; x < out_size; ++x) {
    volatile float col = (float)x + 0.5f;
    volatile float sample = col * inv_scale;
    starts[x] = 0;
    if (sample < 0.f || sample > (float)in_size) continue;
    volatile float lo_f = sample - ks;
    lo_f = lo_f - 0.5f;
    volatile float hi_f = sample + ks;
    hi_f = hi_f - 0.5f;
    long lo = (long)ceilf(lo_f), hi = (long)floorf(hi_f);
    lo = lo < 0 ? 0 : (lo > in_size - 1 ? in_size - 1 : lo);
    hi = (hi < 0 ? 0 : (hi > in_size - 1 ? in_size - 1 : hi)) + 1;
    volatile float total = 0.f;
    tmp.clear();
    for (long s = lo; s < hi; ++s) {
      volatile float pos = (float)s + 0.5f;
      pos = pos - sample;
      volatile float r = pos * inv_ks;
      volatile float wgt = 1.0f - fabsf(r);
      if (wgt < 0.f) wgt = 0.f;
      total = total + wgt;
      tmp.push_back((float)wgt);
    }
    if (fabsf(total) >= 1000.0f * 1.17549435e-38f) {
      const float inv_total = 1.0f / total;
      for (size_t k = 0; k < tmp.size() && (int)k < taps; ++k) weights[x * taps + k] = tmp[k] * inv_total;
    }
    starts[x] = (int32_t)lo;
  }
  return B200_OK;
}

int resample_plan_transpose(int in_size, int out_size, int taps, const int32_t* starts, const float* weights,
                            int32_t* t_starts, float* t_weights, int t_taps) {
  // for every source index i: the contiguous range of outputs whose span contains i
  std::vector<int> lo(in_size, out_size), hi(in_size, -1);
  for (int o = 0; o < out_size; ++o)
    for (int k = 0; k < taps; ++k) {
      int i = starts[o] + k;
      if (i >= in_size || weights[o * taps + k] == 0.f) continue;
      if (o < lo[i]) lo[i] = o;
      if (o > hi[i]) hi[i] = o;
    }
  int need = 1;
  for (int i = 0; i < in_size; ++i)
    if (hi[i] >= lo[i] && hi[i] - lo[i] + 1 > need) need = hi[i] - lo[i] + 1;
  if (!t_starts) return need;
  if (t_taps < need) return fail(B200_ERR_BAD_ARG, "resample_plan_transpose: taps %d < required %d", t_taps, need);
  for (int i = 0; i < in_size * t_taps; ++i) t_weights[i] = 0.f;
  for (int i = 0; i < in_size; ++i) {
    t_starts[i] = hi[i] >= lo[i] ? lo[i] : 0;
    for (int o = lo[i]; o <= hi[i]; ++o) {
      int k = i - starts[o];
      if (k >= 0 && k < taps) t_weights[i * t_taps + (o - lo[i])] = weights[o * taps + k];
    }
  }
  return need;
}

int resample2d(const b200_tensor* x, const b200_tensor* y, const int32_t* hs, const float* hw, int ht,
               const int32_t* ws, const float* ww, int wt, int accumulate, cudaStream_t st) {
  B200_REQUIRE(x->n == y->n && x->c == y->c && x->dtype == y->dtype, B200_ERR_BAD_ARG,
               "resample2d: batch/channel/dtype mismatch");
  TView xv = view_of(x), yv = view_of(y);
  const bool vec = vec_aligned(x, 8) && vec_aligned(y, 8);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    if (vec && y->h <= 65535 && y->n <= 65535) {
      dim3 grid((unsigned)(((long long)y->w * (y->c / 8) + NT - 1) / NT), (unsigned)y->h, (unsigned)y->n);
      resample_vec_kernel<T><<<grid, NT, 0, st>>>(xv, yv, hs, hw, ht, ws, ww, wt, accumulate);
    } else {
      long long total = (long long)y->n * y->h * y->w * y->c;
      resample_scalar_kernel<T><<<grid_for(total), NT, 0, st>>>(xv, yv, hs, hw, ht, ws, ww, wt, accumulate, total);
    }
  });
  return check_launch("resample_kernel");
}

int maxpool2_fwd(const b200_tensor* x, const b200_tensor* y, cudaStream_t st) {
  B200_REQUIRE(x->n == y->n && x->c == y->c && y->h == x->h / 2 && y->w == x->w / 2 && x->dtype == y->dtype,
               B200_ERR_BAD_ARG, "maxpool2_fwd: shape mismatch");
  long long total = (long long)y->n * y->h * y->w * y->c;
  TView xv = view_of(x), yv = view_of(y);
  B200_DISPATCH_DTYPE(x->dtype, T, { maxpool2_fwd_kernel<T><<<grid_for(total), NT, 0, st>>>(xv, yv, total); });
  return check_launch("maxpool2_fwd_kernel");
}

int maxpool2_bwd(const b200_tensor* x, const b200_tensor* y, const b200_tensor* dy, const b200_tensor* dx,
                 int accumulate, cudaStream_t st) {
  B200_REQUIRE(same_shape(x, dx) && same_shape(y, dy) && x->dtype == dx->dtype && y->dtype == dy->dtype,
               B200_ERR_BAD_ARG, "maxpool2_bwd: shape mismatch");
  long long total = (long long)x->n * x->h * x->w * x->c;
  TView xv = view_of(x), yv = view_of(y), dyv = view_of(dy), dxv = view_of(dx);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    maxpool2_bwd_kernel<T><<<grid_for(total), NT, 0, st>>>(xv, yv, dyv, dxv, accumulate, total);
  });
  return check_launch("maxpool2_bwd_kernel");
}

}  // namespace b200
