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
#include <stdlib.h>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int NT = 256;

// grid = (ceil(OW * C/8 / 256), OH, N): one block per output-row segment, one thread per
// (output pixel, 8-channel chunk).  No per-element divisions by H/W; the row's h-taps are block-uniform.
template <typename T, bool BATCHED>
__global__ void __launch_bounds__(NT)
resample_vec_kernel(TView x, TView y, const int* __restrict__ hs, const float* __restrict__ hw, int ht,
                    const int* __restrict__ ws, const float* __restrict__ ww, int wt, int accumulate) {
  pdl_sync();
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

// ---------------------------------------------------------------------------
// Row-marching bf16 kernels.  The gather kernel above spends ~380 instructions per 16-byte output
// chunk and is bound by instruction issue (ncu: issue slots 84 % busy at 20 % of HBM peak), so these
// two do the separable filter as "horizontal taps once per input row, vertical taps from registers":
// a thread owns one (output column, 8-channel chunk) and walks a segment of rows.
//   * resample_up_kernel   (vertical taps HT <= 3: up-sampling and the transpose of down-sampling):
//     driven by output rows; the horizontally filtered input rows h0..h0+HT-1 stay in registers and
//     are shifted when h0 advances.
//   * resample_down_kernel (many vertical taps: antialiased down-sampling and the transpose of
//     up-sampling): driven by input rows; each horizontally filtered row is scattered into the (at
//     most NSLOT) output rows whose span contains it; a row is stored when its span ends.
// ---------------------------------------------------------------------------
struct RsArgs {
  const __nv_bfloat16* x; long long x_sn, x_sh, x_sw; int xh, xw;
  __nv_bfloat16* y; long long y_sn, y_sh, y_sw; int yh, yw;
  int chunks;
  const int* hs; const float* hw; int ht;
  const int* ws; const float* ww; int wt;
  int seg, accumulate;
};

__device__ __forceinline__ float2 rs_unpack(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t rs_pack(float2 v) {
  __nv_bfloat162 h = __float22bfloat162_rn(v);
  return *reinterpret_cast<uint32_t*>(&h);
}

// horizontally filtered input row: r = sum_b ww[b] * x[row][w0 + b]; taps four at a time, zero taps skipped
__device__ __forceinline__ void rs_hrow(const __nv_bfloat16* __restrict__ row, long long x_sw, int xw, int w0,
                                        const float* __restrict__ wrow, int wt, float2 (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = make_float2(0.f, 0.f);
  for (int b0 = 0; b0 < wt; b0 += 4) {
    float wb[4];
    uint4 raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wb[k] = (b0 + k < wt) ? wrow[b0 + k] : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (wb[k] != 0.f) raw[k] = *reinterpret_cast<const uint4*>(row + (long long)min(w0 + b0 + k, xw - 1) * x_sw);
      else raw[k] = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 w2 = make_float2(wb[k], wb[k]);
      const uint32_t* q = reinterpret_cast<const uint32_t*>(&raw[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = __ffma2_rn(rs_unpack(q[i]), w2, r[i]);
    }
  }
}

// same with the (at most NW) horizontal weights already in registers: every load of the row is issued back to back
template <int NW>
__device__ __forceinline__ void rs_hrow_reg(const __nv_bfloat16* __restrict__ row, long long x_sw, int xw, int w0,
                                            const float (&wv)[NW], float2 (&r)[4]) {
  uint4 raw[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    if (wv[k] != 0.f) raw[k] = *reinterpret_cast<const uint4*>(row + (long long)min(w0 + k, xw - 1) * x_sw);
    else raw[k] = make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const float2 w2 = make_float2(wv[k], wv[k]);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(&raw[k]);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = __ffma2_rn(rs_unpack(q[i]), w2, r[i]);
  }
}

__device__ __forceinline__ void rs_store(__nv_bfloat16* dst, const float2 (&v)[4], int accumulate) {
  float2 o[4] = {v[0], v[1], v[2], v[3]};
  if (accumulate) {
    const uint4 e = *reinterpret_cast<const uint4*>(dst);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(&e);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __fadd2_rn(o[i], rs_unpack(q[i]));
  }
  uint4 out;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
  for (int i = 0; i < 4; ++i) ow[i] = rs_pack(o[i]);
  *reinterpret_cast<uint4*>(dst) = out;
}

template <int HT>
__global__ void __launch_bounds__(NT)
resample_up_kernel(const RsArgs a) {
  pdl_sync();
  const int item = blockIdx.x * NT + threadIdx.x;
  const int ow = item / a.chunks, j = item - ow * a.chunks;
  if (ow >= a.yw) return;
  const int n = blockIdx.z;
  const int oh0 = blockIdx.y * a.seg, oh1 = min(oh0 + a.seg, a.yh);
  const __nv_bfloat16* xp = a.x + (long long)n * a.x_sn + j * 8;
  __nv_bfloat16* yp = a.y + (long long)n * a.y_sn + (long long)ow * a.y_sw + j * 8;
  const int w0 = a.ws[ow];
  const float* wrow = a.ww + ow * a.wt;
  float2 cache[HT][4];
  int cur = -(1 << 30);
  const bool regw = a.wt <= 4;
  float wv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) wv[k] = (regw && k < a.wt) ? wrow[k] : 0.f;
  auto hrow = [&](int ih, float2 (&dst)[4]) {
    const __nv_bfloat16* row = xp + (long long)min(ih, a.xh - 1) * a.x_sh;
    if (regw) rs_hrow_reg<4>(row, a.x_sw, a.xw, w0, wv, dst);
    else rs_hrow(row, a.x_sw, a.xw, w0, wrow, a.wt, dst);
  };
  for (int oh = oh0; oh < oh1; ++oh) {
    const int h0 = a.hs[oh];                 // block-uniform
    if (h0 != cur) {
      if (HT > 1 && h0 == cur + 1) {
#pragma unroll
        for (int t = 0; t + 1 < HT; ++t)
#pragma unroll
          for (int i = 0; i < 4; ++i) cache[t][i] = cache[t + 1][i];
        hrow(h0 + HT - 1, cache[HT - 1]);
      } else {
#pragma unroll
        for (int t = 0; t < HT; ++t) hrow(h0 + t, cache[t]);
      }
      cur = h0;
    }
    float2 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < HT; ++t) {
      const float wa = a.hw[oh * HT + t];
      const float2 w2 = make_float2(wa, wa);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = __ffma2_rn(cache[t][i], w2, acc[i]);
    }
    rs_store(yp + (long long)oh * a.y_sh, acc, a.accumulate);
  }
}

template <int NSLOT, int NW>
__global__ void __launch_bounds__(NT)
resample_down_kernel(const RsArgs a) {
  pdl_sync();
  const int item = blockIdx.x * NT + threadIdx.x;
  const int ow = item / a.chunks, j = item - ow * a.chunks;
  if (ow >= a.yw) return;
  const int n = blockIdx.z;
  const int oh0 = blockIdx.y * a.seg, oh1 = min(oh0 + a.seg, a.yh);
  const __nv_bfloat16* xp = a.x + (long long)n * a.x_sn + j * 8;
  __nv_bfloat16* yp = a.y + (long long)n * a.y_sn + (long long)ow * a.y_sw + j * 8;
  const int w0 = a.ws[ow];
  const float* wrow = a.ww + ow * a.wt;
  constexpr int FAR = 1 << 29;
  int o_s[NSLOT], lo_s[NSLOT];
  float2 acc[NSLOT][4];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    o_s[s] = oh0 + s;
    lo_s[s] = o_s[s] < oh1 ? a.hs[o_s[s]] : FAR;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[s][i] = make_float2(0.f, 0.f);
  }
  float wv[NW > 0 ? NW : 1];
  if (NW > 0) {
#pragma unroll
    for (int k = 0; k < NW; ++k) wv[k] = k < a.wt ? wrow[k] : 0.f;
  }
  const int ih_end = a.hs[oh1 - 1] + a.ht;
  for (int ih = a.hs[oh0]; ih < ih_end; ++ih) {        // block-uniform trip count and slot state
    float2 r[4];
    if (NW > 0) rs_hrow_reg<(NW > 0 ? NW : 1)>(xp + (long long)min(ih, a.xh - 1) * a.x_sh, a.x_sw, a.xw, w0, wv, r);
    else rs_hrow(xp + (long long)min(ih, a.xh - 1) * a.x_sh, a.x_sw, a.xw, w0, wrow, a.wt, r);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int k = ih - lo_s[s];
      if (k >= 0 && k < a.ht) {
        const float wa = a.hw[o_s[s] * a.ht + k];
        const float2 w2 = make_float2(wa, wa);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[s][i] = __ffma2_rn(r[i], w2, acc[s][i]);
        if (k == a.ht - 1) {
          rs_store(yp + (long long)o_s[s] * a.y_sh, acc[s], a.accumulate);
          o_s[s] += NSLOT;
          lo_s[s] = o_s[s] < oh1 ? a.hs[o_s[s]] : FAR;
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[s][i] = make_float2(0.f, 0.f);
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
resample_scalar_kernel(TView x, TView y, const int* __restrict__ hs, const float* __restrict__ hw, int ht,
                       const int* __restrict__ ws, const float* __restrict__ ww, int wt, int accumulate,
                       long long total) {
  pdl_sync();
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
  pdl_sync();
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
  pdl_sync();
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

// vector forms (C % 8 == 0, 16-byte aligned, even H and W): one thread per (output pixel, 8-channel chunk) moves the
// 2x2 window as four 16-byte loads -- the scalar kernels above issue 2-byte accesses and three divisions per element
template <typename T>
__global__ void __launch_bounds__(NT)
maxpool2_fwd_vec_kernel(TView x, TView y, long long total) {
  pdl_sync();
  const T* xp = reinterpret_cast<const T*>(x.data);
  T* yp = reinterpret_cast<T*>(y.data);
  const int chunks = y.c / 8;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % chunks) * 8;
    long long p = i / chunks;
    const int ow = (int)(p % y.w); p /= y.w;
    const int oh = (int)(p % y.h);
    const int n = (int)(p / y.h);
    const T* s = xp + pix_offset(x, n, 2 * oh, 2 * ow) + c;
    float a[8], b[8], d[8], e[8];
    Vec8<T>::load(s, a); Vec8<T>::load(s + x.sw, b); Vec8<T>::load(s + x.sh, d); Vec8<T>::load(s + x.sh + x.sw, e);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(d[k], e[k]));
    Vec8<T>::store(yp + pix_offset(y, n, oh, ow) + c, a);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
maxpool2_bwd_vec_kernel(TView x, TView y, TView dy, TView dx, int accumulate, long long total) {
  pdl_sync();
  const T* xp = reinterpret_cast<const T*>(x.data);
  const T* yp = reinterpret_cast<const T*>(y.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  T* dxp = reinterpret_cast<T*>(dx.data);
  const int chunks = y.c / 8;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % chunks) * 8;
    long long p = i / chunks;
    const int ow = (int)(p % y.w); p /= y.w;
    const int oh = (int)(p % y.h);
    const int n = (int)(p / y.h);
    float m[8], g[8], v[4][8];
    Vec8<T>::load(yp + pix_offset(y, n, oh, ow) + c, m);
    Vec8<T>::load(dyp + pix_offset(dy, n, oh, ow) + c, g);
    const long long xo = pix_offset(x, n, 2 * oh, 2 * ow) + c;
    Vec8<T>::load(xp + xo, v[0]); Vec8<T>::load(xp + xo + x.sw, v[1]);
    Vec8<T>::load(xp + xo + x.sh, v[2]); Vec8<T>::load(xp + xo + x.sh + x.sw, v[3]);
    float o[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // the gradient goes to the FIRST maximum of the window in row-major order (TF MaxPoolGrad)
      const int first = v[0][k] == m[k] ? 0 : (v[1][k] == m[k] ? 1 : (v[2][k] == m[k] ? 2 : 3));
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][k] = first == q ? g[k] : 0.f;
    }
    const long long dO = pix_offset(dx, n, 2 * oh, 2 * ow) + c;
    const long long offs[4] = {0, dx.sw, dx.sh, dx.sh + dx.sw};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      T* dst = dxp + dO + offs[q];
      if (accumulate) {
        float e[8];
        Vec8<T>::load(dst, e);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[q][k] += e[k];
      }
      Vec8<T>::store(dst, o[q]);
    }
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

// Shift leading zero weights out of every row (start += shift) and return the effective tap count
// (largest non-zero extent); rows keep their pitch `taps`.  Host tables, in place.
int resample_compact(int n_out, int taps, int32_t* starts, float* weights) {
  int eff = 1;
  for (int o = 0; o < n_out; ++o) {
    float* w = weights + (size_t)o * taps;
    int first = 0;
    while (first < taps - 1 && w[first] == 0.f) ++first;
    int last = taps - 1;
    while (last > first && w[last] == 0.f) --last;
    if (first > 0) {
      for (int k = 0; k + first < taps; ++k) w[k] = w[k + first];
      for (int k = taps - first; k < taps; ++k) w[k] = 0.f;
      starts[o] += first;
    }
    if (last - first + 1 > eff) eff = last - first + 1;
  }
  return eff;
}

// Which marching kernel may run the ROW axis of a (compacted) table: 1 = up (taps <= 3),
// 2 / 3 = down with 4 / 6 slots (output o + NSLOT must start after output o has ended), 0 = gather only.
int resample_mode(int n_out, int taps, const int32_t* starts) {
  if (taps <= 3) return 1;
  for (int o = 1; o < n_out; ++o)
    if (starts[o] < starts[o - 1]) return 0;
  const int slots[2] = {4, 6};
  for (int m = 0; m < 2; ++m) {
    bool ok = true;
    for (int o = 0; o + slots[m] < n_out && ok; ++o) ok = starts[o + slots[m]] - starts[o] >= taps;
    if (ok) return 2 + m;
  }
  return 0;
}

int resample2d(const b200_tensor* x, const b200_tensor* y, const int32_t* hs, const float* hw, int ht,
               const int32_t* ws, const float* ww, int wt, int accumulate, int mode, cudaStream_t st) {
  B200_REQUIRE(x->n == y->n && x->c == y->c && x->dtype == y->dtype, B200_ERR_BAD_ARG,
               "resample2d: batch/channel/dtype mismatch");
  TView xv = view_of(x), yv = view_of(y);
  const bool vec = vec_aligned(x, 8) && vec_aligned(y, 8);
  static const int allow_march = getenv("B200_RESAMPLE_V") ? atoi(getenv("B200_RESAMPLE_V")) : 1;
  if (allow_march && mode > 0 && vec && x->dtype == B200_BF16 && y->n <= 65535 && (mode != 1 || ht <= 3)) {
    RsArgs a;
    a.x = reinterpret_cast<const __nv_bfloat16*>(x->data); a.x_sn = x->stride_n; a.x_sh = x->stride_h; a.x_sw = x->stride_w;
    a.xh = x->h; a.xw = x->w;
    a.y = reinterpret_cast<__nv_bfloat16*>(y->data); a.y_sn = y->stride_n; a.y_sh = y->stride_h; a.y_sw = y->stride_w;
    a.yh = y->h; a.yw = y->w;
    a.chunks = y->c / 8;
    a.hs = hs; a.hw = hw; a.ht = ht; a.ws = ws; a.ww = ww; a.wt = wt;
    a.accumulate = accumulate;
    const int bx = (int)(((long long)y->w * a.chunks + NT - 1) / NT);
    // rows per thread: long enough to amortise the first rows of a segment, short enough to fill the GPU
    const long long base_blocks = (long long)bx * y->n;
    static const int blocks_per_sm = getenv("B200_RS_BLOCKS") ? atoi(getenv("B200_RS_BLOCKS")) : 6;
    static const int min_rows_down = getenv("B200_RS_MINROWS") ? atoi(getenv("B200_RS_MINROWS")) : 4;
    int nseg = (int)(((long long)blocks_per_sm * sm_count() + base_blocks - 1) / base_blocks);
    const int min_rows = mode == 1 ? 8 : min_rows_down;
    if (nseg > (y->h + min_rows - 1) / min_rows) nseg = (y->h + min_rows - 1) / min_rows;
    if (nseg < 1) nseg = 1;
    a.seg = (y->h + nseg - 1) / nseg;
    nseg = (y->h + a.seg - 1) / a.seg;
    dim3 grid((unsigned)bx, (unsigned)nseg, (unsigned)y->n);
    if (mode == 1) {
      if (ht == 1) launch_pdl(resample_up_kernel<1>, grid, NT, 0, st, a);
      else if (ht == 2) launch_pdl(resample_up_kernel<2>, grid, NT, 0, st, a);
      else launch_pdl(resample_up_kernel<3>, grid, NT, 0, st, a);
    } else if (mode == 2) {
      if (wt <= 4) launch_pdl(resample_down_kernel<4, 4>, grid, NT, 0, st, a);
      else if (wt <= 8) launch_pdl(resample_down_kernel<4, 8>, grid, NT, 0, st, a);
      else launch_pdl(resample_down_kernel<4, 0>, grid, NT, 0, st, a);
    } else {
      if (wt <= 4) launch_pdl(resample_down_kernel<6, 4>, grid, NT, 0, st, a);
      else launch_pdl(resample_down_kernel<6, 0>, grid, NT, 0, st, a);
    }
    return check_launch("resample_march_kernel");
  }
  B200_DISPATCH_DTYPE(x->dtype, T, {
    if (vec && y->h <= 65535 && y->n <= 65535) {
      dim3 grid((unsigned)(((long long)y->w * (y->c / 8) + NT - 1) / NT), (unsigned)y->h, (unsigned)y->n);
      launch_pdl(resample_vec_kernel<T, false>, grid, NT, 0, st, xv, yv, hs, hw, ht, ws, ww, wt, accumulate);
    } else {
      long long total = (long long)y->n * y->h * y->w * y->c;
      launch_pdl(resample_scalar_kernel<T>, grid_for(total), NT, 0, st, xv, yv, hs, hw, ht, ws, ww, wt, accumulate, total);
    }
  });
  return check_launch("resample_kernel");
}

// 8-channel vectors need 16-byte aligned pixels for bf16, 32-byte (two 16-byte halves: 16 is enough) for fp32
static bool pool_vec_ok(const b200_tensor* t) {
  const size_t es = dtype_size(t->dtype);
  return t->c % 8 == 0 && reinterpret_cast<uintptr_t>(t->data) % 16 == 0 && (t->stride_w * es) % 16 == 0 &&
         (t->stride_h * es) % 16 == 0 && (t->stride_n * es) % 16 == 0;
}

int maxpool2_fwd(const b200_tensor* x, const b200_tensor* y, cudaStream_t st) {
  B200_REQUIRE(x->n == y->n && x->c == y->c && y->h == x->h / 2 && y->w == x->w / 2 && x->dtype == y->dtype,
               B200_ERR_BAD_ARG, "maxpool2_fwd: shape mismatch");
  long long total = (long long)y->n * y->h * y->w * y->c;
  TView xv = view_of(x), yv = view_of(y);
  if (pool_vec_ok(x) && pool_vec_ok(y)) {
    total /= 8;
    B200_DISPATCH_DTYPE(x->dtype, T, { launch_pdl(maxpool2_fwd_vec_kernel<T>, grid_for(total), NT, 0, st, xv, yv, total); });
    return check_launch("maxpool2_fwd_vec_kernel");
  }
  B200_DISPATCH_DTYPE(x->dtype, T, { launch_pdl(maxpool2_fwd_kernel<T>, grid_for(total), NT, 0, st, xv, yv, total); });
  return check_launch("maxpool2_fwd_kernel");
}

int maxpool2_bwd(const b200_tensor* x, const b200_tensor* y, const b200_tensor* dy, const b200_tensor* dx,
                 int accumulate, cudaStream_t st) {
  B200_REQUIRE(same_shape(x, dx) && same_shape(y, dy) && x->dtype == dx->dtype && y->dtype == dy->dtype,
               B200_ERR_BAD_ARG, "maxpool2_bwd: shape mismatch");
  long long total = (long long)x->n * x->h * x->w * x->c;
  TView xv = view_of(x), yv = view_of(y), dyv = view_of(dy), dxv = view_of(dx);
  if (x->h % 2 == 0 && x->w % 2 == 0 && pool_vec_ok(x) && pool_vec_ok(y) && pool_vec_ok(dy) && pool_vec_ok(dx)) {
    const long long items = (long long)y->n * y->h * y->w * (y->c / 8);
    B200_DISPATCH_DTYPE(x->dtype, T, {
      launch_pdl(maxpool2_bwd_vec_kernel<T>, grid_for(items), NT, 0, st, xv, yv, dyv, dxv, accumulate, items);
    });
    return check_launch("maxpool2_bwd_vec_kernel");
  }
  B200_DISPATCH_DTYPE(x->dtype, T, {
    launch_pdl(maxpool2_bwd_kernel<T>, grid_for(total), NT, 0, st, xv, yv, dyv, dxv, accumulate, total);
  });
  return check_launch("maxpool2_bwd_kernel");
}

}  // namespace b200
