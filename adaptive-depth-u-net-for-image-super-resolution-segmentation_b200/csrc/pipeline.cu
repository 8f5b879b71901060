// Device-side patch pipeline: random/grid crop out of a resident image and the LR synthesis
// (cv2 INTER_AREA shrink + INTER_CUBIC enlargement) of shared/pipeline.py:79-136 of the reference.
//
// OpenCV's resize is separable with a short list of taps per output index; the host helpers below build
// the same (index, weight) tables OpenCV builds (computeResizeAreaTab / interpolateCubic with A = -0.75,
// border taps clamped) and ONE gather kernel applies them: horizontal taps first, then vertical, fp32,
// which is the order OpenCV reduces in.  All of it is HBM-bound byte moving: a batch of 64 128x128 patches
// is 3 MB of uint8 read and 2 x 12.6 MB of fp32 written.
#include <math.h>
#include <float.h>

#include "common.cuh"

namespace b200 {

// ---- host: OpenCV tap tables ------------------------------------------------------------------
static bool area_is_fast(int in_size, int out_size, int* iscale) {
  const double scale = (double)in_size / out_size;
  const int is = (int)lrint(scale);
  *iscale = is;
  return fabs(scale - is) < DBL_EPSILON;
}

// one row of the INTER_AREA table; returns the tap count (idx / w may be NULL to count only)
static int area_row(int in_size, int out_size, int d, int32_t* idx, float* w) {
  int is = 1;
  if (area_is_fast(in_size, out_size, &is)) {
    for (int k = 0; k < is && idx; ++k) { idx[k] = d * is + k; w[k] = (float)(1.0 / is); }
    return is;
  }
  const double scale = (double)in_size / out_size;
  const double f1 = d * scale, f2 = f1 + scale;
  const double cell = fmin(scale, in_size - f1);
  int s1 = (int)ceil(f1), s2 = (int)floor(f2);
  if (s2 > in_size - 1) s2 = in_size - 1;
  if (s1 > s2) s1 = s2;
  int n = 0;
  if (s1 - f1 > 1e-3) { if (idx) { idx[n] = s1 - 1; w[n] = (float)((s1 - f1) / cell); } ++n; }
  for (int s = s1; s < s2; ++s) { if (idx) { idx[n] = s; w[n] = (float)(1.0 / cell); } ++n; }
  if (f2 - s2 > 1e-3) { if (idx) { idx[n] = s2; w[n] = (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell); } ++n; }
  return n;
}

int cv_resize_taps(int in_size, int out_size, int interp) {
  if (interp == B200_CV_INTER_CUBIC) return 4;
  int taps = 0;
  for (int d = 0; d < out_size; ++d) {
    const int n = area_row(in_size, out_size, d, nullptr, nullptr);
    if (n > taps) taps = n;
  }
  return taps;
}

int cv_resize_plan(int in_size, int out_size, int interp, int32_t* idx, float* w, int taps) {
  if (interp == B200_CV_INTER_CUBIC) {
    const double scale = (double)in_size / out_size;
    const float A = -0.75f;
    for (int d = 0; d < out_size; ++d) {
      volatile float fx = (float)((d + 0.5) * scale - 0.5);
      const int sx = (int)floorf(fx);
      volatile float x = fx - (float)sx;
      // interpolateCubic, every intermediate rounded to float (volatile keeps the host compiler from fusing)
      volatile float x1 = x + 1.0f;
      volatile float t0 = A * x1;          t0 = t0 - 5.0f * A;  t0 = t0 * x1;  t0 = t0 + 8.0f * A;  t0 = t0 * x1;
      volatile float c0 = t0 - 4.0f * A;
      volatile float t1 = (A + 2.0f) * x;  t1 = t1 - (A + 3.0f); t1 = t1 * x;  t1 = t1 * x;
      volatile float c1 = t1 + 1.0f;
      volatile float u = 1.0f - x;
      volatile float t2 = (A + 2.0f) * u;  t2 = t2 - (A + 3.0f); t2 = t2 * u;  t2 = t2 * u;
      volatile float c2 = t2 + 1.0f;
      volatile float c3 = 1.0f - c0;       c3 = c3 - c1;         c3 = c3 - c2;
      const float c[4] = {c0, c1, c2, c3};
      for (int k = 0; k < 4; ++k) {
        int s = sx - 1 + k;
        s = s < 0 ? 0 : (s > in_size - 1 ? in_size - 1 : s);
        idx[d * taps + k] = s;
        w[d * taps + k] = c[k];
      }
      for (int k = 4; k < taps; ++k) { idx[d * taps + k] = idx[d * taps + 3]; w[d * taps + k] = 0.0f; }
    }
    return B200_OK;
  }
  for (int d = 0; d < out_size; ++d) {
    const int n = area_row(in_size, out_size, d, idx + (size_t)d * taps, w + (size_t)d * taps);
    for (int k = n; k < taps; ++k) { idx[d * taps + k] = idx[d * taps + n - 1]; w[d * taps + k] = 0.0f; }
  }
  return B200_OK;
}

// ---- crop: resident image (uint8 or fp32 HxWx3) -> fp32 patches ---------------------------------
template <typename SrcT> __device__ __forceinline__ float to_unit(SrcT v);
template <> __device__ __forceinline__ float to_unit<unsigned char>(unsigned char v) { return (float)v / 255.0f; }
template <> __device__ __forceinline__ float to_unit<float>(float v) { return v; }

// one thread per element of a patch row (P*3 contiguous floats in both the image and the patch)
template <typename SrcT>
__global__ void __launch_bounds__(256)
patch_extract_kernel(const SrcT* __restrict__ img, int img_h, int img_w, const int32_t* __restrict__ origins,
                     float* __restrict__ hr, int n, int p, long long sn, long long sh) {
  pdl_sync();
  const int row_elems = p * 3;
  const long long total = (long long)n * p * row_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % row_elems);
    const long long q = i / row_elems;
    const int y = (int)(q % p);
    const int b = (int)(q / p);
    int top = origins[2 * b], left = origins[2 * b + 1];
    top = min(max(top, 0), img_h - p);          // an origin outside the image cannot read out of bounds
    left = min(max(left, 0), img_w - p);
    const SrcT v = img[((long long)(top + y) * img_w + left) * 3 + e];
    hr[(long long)b * sn + (long long)y * sh + e] = to_unit<SrcT>(v);
  }
}

int patch_extract(const void* image, int image_dtype, int img_h, int img_w, const int32_t* origins, const b200_tensor* hr,
                  cudaStream_t st) {
  const int p = hr->h;
  const long long total = (long long)hr->n * p * p * 3;
  const int block = 256;
  long long want = (total + block - 1) / block;
  const int grid = (int)(want < (long long)sm_count() * 16 ? want : (long long)sm_count() * 16);
  if (image_dtype == B200_U8)
    launch_pdl(patch_extract_kernel<unsigned char>, grid, block, 0, st, static_cast<const unsigned char*>(image), img_h, img_w,
                                                                origins, static_cast<float*>(hr->data), hr->n, p,
                                                                hr->stride_n, hr->stride_h);
  else
    launch_pdl(patch_extract_kernel<float>, grid, block, 0, st, static_cast<const float*>(image), img_h, img_w, origins,
                                                        static_cast<float*>(hr->data), hr->n, p, hr->stride_n,
                                                        hr->stride_h);
  return check_launch("patch_extract_kernel");
}

// ---- separable gather over explicit tap tables --------------------------------------------------
// y[n,oy,ox,c] = sum_j hw[oy][j] * ( sum_k ww[ox][k] * f(x[n, hi[oy][j], wi[ox][k], c]) ),  f = clip to [0,1] or id
template <int C, bool CLIP>
__global__ void __launch_bounds__(256)
gather2d_kernel(TView x, TView y, const int32_t* __restrict__ hi, const float* __restrict__ hw, int ht,
                const int32_t* __restrict__ wi, const float* __restrict__ ww, int wt) {
  pdl_sync();
  const long long total = (long long)y.n * y.h * y.w;
  const float* __restrict__ xp = static_cast<const float*>(x.data);
  float* __restrict__ yp = static_cast<float*>(y.data);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % y.w);
    const long long q = i / y.w;
    const int oy = (int)(q % y.h);
    const int n = (int)(q / y.h);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.0f;
    for (int j = 0; j < ht; ++j) {
      const int iy = hi[oy * ht + j];
      const float wy = hw[oy * ht + j];
      const float* row = xp + (long long)n * x.sn + (long long)iy * x.sh;
      float r[C];
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = 0.0f;
      for (int k = 0; k < wt; ++k) {
        const float* px = row + (long long)wi[ox * wt + k] * x.sw;
        const float wx = ww[ox * wt + k];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float v = px[c];
          if (CLIP) v = fminf(fmaxf(v, 0.0f), 1.0f);
          r[c] += v * wx;
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += r[c] * wy;
    }
    float* out = yp + (long long)n * y.sn + (long long)oy * y.sh + (long long)ox * y.sw;
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = acc[c];
  }
}

int gather2d(const b200_tensor* x, const b200_tensor* y, const int32_t* hi, const float* hw, int ht, const int32_t* wi,
             const float* ww, int wt, int clip01, cudaStream_t st) {
  const TView xv = view_of(x), yv = view_of(y);
  const long long total = (long long)y->n * y->h * y->w;
  const int block = 256;
  long long want = (total + block - 1) / block;
  const int grid = (int)(want < (long long)sm_count() * 32 ? want : (long long)sm_count() * 32);
#define B200_G2D(CH, CL) launch_pdl(gather2d_kernel<CH, CL>, grid, block, 0, st, xv, yv, hi, hw, ht, wi, ww, wt)
  if (x->c == 3) { if (clip01) B200_G2D(3, true); else B200_G2D(3, false); }
  else if (x->c == 1) { if (clip01) B200_G2D(1, true); else B200_G2D(1, false); }
  else return fail(B200_ERR_UNSUPPORTED, "gather2d: %d channels (1 or 3 supported)", x->c);
#undef B200_G2D
  return check_launch("gather2d_kernel");
}

// ---- row gather / scatter between the shuffle pool and a batch ----------------------------------
template <typename V>
__global__ void __launch_bounds__(256)
copy_rows_kernel(const V* __restrict__ src, const int32_t* __restrict__ src_rows, V* __restrict__ dst,
                 const int32_t* __restrict__ dst_rows, long long row_vecs) {
  pdl_sync();
  const int r = blockIdx.y;
  const long long s = (src_rows ? (long long)src_rows[r] : (long long)r) * row_vecs;
  const long long d = (dst_rows ? (long long)dst_rows[r] : (long long)r) * row_vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_vecs; i += (long long)gridDim.x * blockDim.x)
    dst[d + i] = src[s + i];
}

int copy_rows(const float* src, const int32_t* src_rows, float* dst, const int32_t* dst_rows, int n_rows,
              long long row_elems, cudaStream_t st) {
  const bool vec = row_elems % 4 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0;
  const long long units = vec ? row_elems / 4 : row_elems;
  long long bx = (units + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)n_rows);
  if (vec)
    launch_pdl(copy_rows_kernel<float4>, grid, 256, 0, st, reinterpret_cast<const float4*>(src), src_rows,
                                                   reinterpret_cast<float4*>(dst), dst_rows, units);
  else
    launch_pdl(copy_rows_kernel<float>, grid, 256, 0, st, src, src_rows, dst, dst_rows, units);
  return check_launch("copy_rows_kernel");
}

}  // namespace b200
