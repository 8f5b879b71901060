// Adam (keras semantics: eps 1e-7, bias-corrected step size), dtype casts, strided
// copies and the SIMT Conv2DTranspose(2, strides=2) kernels.
//
// Replaces tf.keras.optimizers.Adam at Super_resolution/code/train_adaptive_unet.py:490
// and keras Conv2DTranspose at Segmenation/code/unet_vinillia.py:67.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int NT = 256;

inline int grid_for(long long items, int per_thread = 1) {
  long long b = (items + (long long)NT * per_thread - 1) / ((long long)NT * per_thread);
  long long cap = 16LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Dynamic loss scaling (keras LossScaleOptimizer, the optimizer wrapper of the reference's mixed_float16 policy,
// Super_resolution/code/train_adaptive_unet.py:471-477): device-resident state ls = {scale, good steps, found_inf, -}.
// The loss gradient is multiplied by `scale`; a step whose gradients hold an inf / nan is skipped (no moment, weight or
// step-counter update) and halves the scale; `growth` finite steps in a row double it.
__global__ void adam_advance_kernel(int* step, const float* __restrict__ ls) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0 && !(ls && ls[2] != 0.f)) *step += 1;
}

__global__ void __launch_bounds__(NT) scale_by_device_kernel(float* p, size_t count, const float* __restrict__ ls) {
  pdl_sync();
  const float s = ls[0];
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < count; i += (size_t)gridDim.x * NT) p[i] *= s;
}
__global__ void __launch_bounds__(NT) scale_by_device_bf16_kernel(__nv_bfloat16* p, size_t count, const float* __restrict__ ls) {
  pdl_sync();
  const float s = ls[0];
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < count; i += (size_t)gridDim.x * NT)
    p[i] = __float2bfloat16_rn(__bfloat162float(p[i]) * s);
}

__global__ void __launch_bounds__(NT) finite_check_kernel(const float* __restrict__ g, size_t count, float* __restrict__ ls) {
  pdl_sync();
  bool bad = false;
  const size_t n4 = count / 4;
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * NT) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * NT + threadIdx.x; i < count; i += (size_t)gridDim.x * NT) bad |= !isfinite(g[i]);
  if (__syncthreads_or(bad) && threadIdx.x == 0) ls[2] = 1.f;      // every writer stores the same value
}

__global__ void loss_scale_update_kernel(float* ls, float growth) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (ls[2] != 0.f) {
    ls[0] = fmaxf(ls[0] * 0.5f, 1.f);
    ls[1] = 0.f;
    ls[2] = 0.f;
    ls[3] += 1.f;                                                   // skipped steps so far
  } else {
    ls[1] += 1.f;
    if (ls[1] >= growth) { ls[0] *= 2.f; ls[1] = 0.f; }
  }
}

// alpha = lr * sqrt(1 - b2^t) / (1 - b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= alpha*m/(sqrt(v)+eps)
__global__ void __launch_bounds__(NT)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            size_t count, const float* __restrict__ hyper, const int* __restrict__ step,
            __nv_bfloat16* __restrict__ shadow, const float* __restrict__ ls) {
  pdl_sync();
  if (ls && ls[2] != 0.f) return;                 // loss scaling: a non-finite gradient skips the whole step
  const float inv_scale = ls ? 1.f / ls[0] : 1.f;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const float omb1 = hyper[4], omb2 = hyper[5];  // (1-beta) evaluated in double by the host, as keras does
  const float t = (float)(*step);
  const float alpha = lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
  const size_t n4 = count / 4;
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * NT) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    gv.x *= inv_scale; gv.y *= inv_scale; gv.z *= inv_scale; gv.w *= inv_scale;
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gg = &gv.x; float* mm = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mm[k] += (gg[k] - mm[k]) * omb1;
      vp[k] += (gg[k] * gg[k] - vp[k]) * omb2;
      pp[k] -= alpha * mm[k] / (sqrtf(vp[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      __nv_bfloat162 a = __floats2bfloat162_rn(pv.x, pv.y), b = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  // tail
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * NT + threadIdx.x; i < count; i += (size_t)gridDim.x * NT) {
    const float gi = g[i] * inv_scale;
    float mm = m[i] + (gi - m[i]) * omb1;
    float vv = v[i] + (gi * gi - v[i]) * omb2;
    float pp = p[i] - alpha * mm / (sqrtf(vv) + eps);
    m[i] = mm; v[i] = vv; p[i] = pp;
    if (shadow) shadow[i] = __float2bfloat16_rn(pp);
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(NT) cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, size_t count) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < count; i += (size_t)gridDim.x * NT) stf(d + i, ldf(s + i));
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(NT) copy_tensor_kernel(TView s, TView d, long long total) {
  pdl_sync();
  const TS* sp = reinterpret_cast<const TS*>(s.data);
  TD* dp = reinterpret_cast<TD*>(d.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % d.c);
    long long p = i / d.c;
    int w = (int)(p % d.w); p /= d.w;
    int h = (int)(p % d.h);
    int n = (int)(p / d.h);
    stf(dp + pix_offset(d, n, h, w) + c, ldf(sp + pix_offset(s, n, h, w) + c));
  }
}

__global__ void __launch_bounds__(NT) scale_kernel(float* p, size_t count, float s) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < count; i += (size_t)gridDim.x * NT) p[i] *= s;
}

// ---- Conv2DTranspose(k=2, s=2): out[2i+a,2j+b,o] = sum_c in[i,j,c] K[a,b,o,c] + bias[o] -------------
template <typename T>
__global__ void __launch_bounds__(NT)
convT2_fprop_kernel(TView x, const T* __restrict__ k, const float* __restrict__ bias, TView y, long long total) {
  pdl_sync();
  const T* xp = reinterpret_cast<const T*>(x.data);
  T* yp = reinterpret_cast<T*>(y.data);
  const int Cin = x.c, Cout = y.c;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int o = (int)(i % Cout);
    long long p = i / Cout;
    int ow = (int)(p % y.w); p /= y.w;
    int oh = (int)(p % y.h);
    int n = (int)(p / y.h);
    const T* src = xp + pix_offset(x, n, oh / 2, ow / 2);
    const T* kk = k + (((long long)(oh & 1) * 2 + (ow & 1)) * Cout + o) * Cin;
    float acc = bias ? bias[o] : 0.f;
    for (int c = 0; c < Cin; ++c) acc += ldf(src + c) * ldf(kk + c);
    stf(yp + pix_offset(y, n, oh, ow) + o, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
convT2_dgrad_kernel(TView dy, const T* __restrict__ k, TView dx, long long total) {
  pdl_sync();
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  T* dxp = reinterpret_cast<T*>(dx.data);
  const int Cin = dx.c, Cout = dy.c;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % Cin);
    long long p = i / Cin;
    int w = (int)(p % dx.w); p /= dx.w;
    int h = (int)(p % dx.h);
    int n = (int)(p / dx.h);
    float acc = 0.f;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const T* g = dyp + pix_offset(dy, n, 2 * h + a, 2 * w + b);
        const T* kk = k + ((long long)(a * 2 + b) * Cout) * Cin + c;
        for (int o = 0; o < Cout; ++o) acc += ldf(g + o) * ldf(kk + (long long)o * Cin);
      }
    stf(dxp + pix_offset(dx, n, h, w) + c, acc);
  }
}

// dK[a,b,o,c] = sum_{n,i,j} dy[n,2i+a,2j+b,o] * x[n,i,j,c]; grid = (4*Cout, pixel splits), threads over c, one atomic
// per (block, c) into the zeroed dk  (fallback for channel counts the tensor-core path does not take)
template <typename T>
__global__ void __launch_bounds__(NT)
convT2_wgrad_kernel(TView x, TView dy, float* __restrict__ dk) {
  pdl_sync();
  const T* xp = reinterpret_cast<const T*>(x.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  const int Cin = x.c, Cout = dy.c;
  const int o = blockIdx.x % Cout, ab = blockIdx.x / Cout;
  const int a = ab / 2, b = ab % 2;
  const long long npix = (long long)x.n * x.h * x.w;
  const long long p0 = npix * blockIdx.y / gridDim.y, p1 = npix * (blockIdx.y + 1) / gridDim.y;
  for (int c = threadIdx.x; c < Cin; c += NT) {
    float acc = 0.f;
    for (long long p = p0; p < p1; ++p) {
      int w = (int)(p % x.w);
      long long q = p / x.w;
      int h = (int)(q % x.h);
      int n = (int)(q / x.h);
      acc += ldf(dyp + pix_offset(dy, n, 2 * h + a, 2 * w + b) + o) * ldf(xp + pix_offset(x, n, h, w) + c);
    }
    atomicAdd(dk + ((long long)ab * Cout + o) * Cin + c, acc);
  }
}

}  // namespace

int adam_advance(int32_t* step, const float* ls, cudaStream_t st) {
  launch_pdl(adam_advance_kernel, 1, 32, 0, st, step, ls);
  return check_launch("adam_advance_kernel");
}

int loss_scale_apply(void* data, int dtype, size_t count, const float* ls, cudaStream_t st) {
  if (dtype == B200_BF16) launch_pdl(scale_by_device_bf16_kernel, grid_for((long long)count), NT, 0, st, reinterpret_cast<__nv_bfloat16*>(data), count, ls);
  else launch_pdl(scale_by_device_kernel, grid_for((long long)count), NT, 0, st, reinterpret_cast<float*>(data), count, ls);
  return check_launch("scale_by_device_kernel");
}
int loss_scale_check(const float* g, size_t count, float* ls, cudaStream_t st) {
  launch_pdl(finite_check_kernel, grid_for((long long)count, 8), NT, 0, st, g, count, ls);
  return check_launch("finite_check_kernel");
}
int loss_scale_update(float* ls, float growth, cudaStream_t st) {
  launch_pdl(loss_scale_update_kernel, 1, 32, 0, st, ls, growth);
  return check_launch("loss_scale_update_kernel");
}

int adam_step(float* p, const float* g, float* m, float* v, size_t count, const float* hyper, const int32_t* step,
              void* shadow, const float* ls, cudaStream_t st) {
  B200_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                   ((uintptr_t)v % 16 == 0) && ((uintptr_t)shadow % 8 == 0),
               B200_ERR_BAD_ARG, "adam_step: buffers must be 16-byte aligned");
  launch_pdl(adam_kernel, grid_for((long long)count, 4), NT, 0, st, p, g, m, v, count, hyper, step,
                                                            reinterpret_cast<__nv_bfloat16*>(shadow), ls);
  return check_launch("adam_kernel");
}

int cast(const void* src, int sdt, void* dst, int ddt, size_t count, cudaStream_t st) {
  const int grid = grid_for((long long)count);
  if (sdt == B200_F32 && ddt == B200_BF16)
    launch_pdl(cast_kernel<float, __nv_bfloat16>, grid, NT, 0, st, (const float*)src, (__nv_bfloat16*)dst, count);
  else if (sdt == B200_BF16 && ddt == B200_F32)
    launch_pdl(cast_kernel<__nv_bfloat16, float>, grid, NT, 0, st, (const __nv_bfloat16*)src, (float*)dst, count);
  else if (sdt == B200_F32 && ddt == B200_F32)
    launch_pdl(cast_kernel<float, float>, grid, NT, 0, st, (const float*)src, (float*)dst, count);
  else
    launch_pdl(cast_kernel<__nv_bfloat16, __nv_bfloat16>, grid, NT, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, count);
  return check_launch("cast_kernel");
}

int copy_tensor(const b200_tensor* s, const b200_tensor* d, cudaStream_t st) {
  B200_REQUIRE(same_shape(s, d), B200_ERR_BAD_ARG, "copy_tensor: shape mismatch");
  long long total = (long long)d->n * d->h * d->w * d->c;
  TView sv = view_of(s), dv = view_of(d);
  const int grid = grid_for(total);
  if (s->dtype == B200_F32 && d->dtype == B200_BF16) launch_pdl(copy_tensor_kernel<float, __nv_bfloat16>, grid, NT, 0, st, sv, dv, total);
  else if (s->dtype == B200_BF16 && d->dtype == B200_F32) launch_pdl(copy_tensor_kernel<__nv_bfloat16, float>, grid, NT, 0, st, sv, dv, total);
  else if (s->dtype == B200_F32) launch_pdl(copy_tensor_kernel<float, float>, grid, NT, 0, st, sv, dv, total);
  else launch_pdl(copy_tensor_kernel<__nv_bfloat16, __nv_bfloat16>, grid, NT, 0, st, sv, dv, total);
  return check_launch("copy_tensor_kernel");
}

int scale_inplace(float* p, size_t count, float s, cudaStream_t st) {
  launch_pdl(scale_kernel, grid_for((long long)count), NT, 0, st, p, count, s);
  return check_launch("scale_kernel");
}

// ---- tensor-core form: Conv2DTranspose(k2, s2) is four 1x1 convolutions, one per output parity (a, b), whose
// output (fprop) / input (dgrad, wgrad) is the strided view y[:, a::2, b::2, :] -- the tensor descriptors and the
// TMA maps take arbitrary strides, so the pixel shuffle costs nothing.
bool conv_tc_supported(const b200_tensor* x, int cin, int cout, const b200_tensor* y, int ks);
int conv_tc_launch(const b200_tensor* x, const void* wmat, int cin, int cout, int tap_rev, int b_mn, const float* bias,
                   const b200_tensor* y, int act, int accumulate, cudaStream_t st, const struct ConvLnArgs* ln, int ks,
                   void* ws, size_t ws_bytes, const void* wmat_k = nullptr, int allow_pairs = 1, const struct ConvLnBwdArgs* lnb = nullptr);
bool conv_gemm_wanted(const b200_tensor* x, int cin, int cout, int ks);
size_t conv_gemm_workspace(const b200_tensor* x, int cin, int cout, int ks);
bool wgrad_tc_supported(const b200_tensor* x, const b200_tensor* dy, int ks);
int wgrad_tc_launch(const b200_tensor* x, const b200_tensor* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st,
                    int ks, int atomic);
int bias_act_bwd(const b200_tensor* dy, const b200_tensor* y, int act, const b200_tensor* dz, float* dbias, cudaStream_t st);

static b200_tensor parity_view(const b200_tensor* y, int a, int b) {
  b200_tensor v = *y;
  v.data = reinterpret_cast<char*>(y->data) + ((long long)a * y->stride_h + (long long)b * y->stride_w) * (long long)dtype_size(y->dtype);
  v.h = y->h / 2; v.w = y->w / 2;
  v.stride_h = 2 * y->stride_h; v.stride_w = 2 * y->stride_w;
  return v;
}

// scratch bytes the four 1x1 launches of the tensor-core form want (they run back to back on one stream and share it)
size_t convT2_workspace(const b200_tensor* x, int cin, int cout, int dgrad) {
  if (x->dtype != B200_BF16) return 0;
  b200_tensor v = dgrad ? parity_view(x, 0, 0) : *x;      // dgrad: x is dy, read through its parity views
  return conv_gemm_wanted(&v, cin, cout, 1) ? conv_gemm_workspace(&v, cin, cout, 1) : 0;
}

int convT2_fprop(const b200_tensor* x, const void* kernel, const float* bias, int cout, const b200_tensor* y,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  B200_REQUIRE(y->h == 2 * x->h && y->w == 2 * x->w && y->n == x->n && y->c == cout && x->dtype == y->dtype,
               B200_ERR_BAD_ARG, "convT2x2_fprop: shape mismatch");
  {
    b200_tensor y00 = parity_view(y, 0, 0);
    if (x->dtype == B200_BF16 && conv_tc_supported(x, x->c, cout, &y00, 1)) {
      for (int ab = 0; ab < 4; ++ab) {
        b200_tensor yv = parity_view(y, ab / 2, ab % 2);
        const __nv_bfloat16* k_ab = reinterpret_cast<const __nv_bfloat16*>(kernel) + (long long)ab * cout * x->c;   // [cout][cin], K-major
        int rc = conv_tc_launch(x, k_ab, x->c, cout, 0, 0, bias, &yv, B200_ACT_NONE, 0, st, nullptr, 1, ws, ws_bytes);
        if (rc) return rc;
      }
      return B200_OK;
    }
  }
  long long total = (long long)y->n * y->h * y->w * y->c;
  TView xv = view_of(x), yv = view_of(y);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    launch_pdl(convT2_fprop_kernel<T>, grid_for(total), NT, 0, st, xv, (const T*)kernel, bias, yv, total);
  });
  return check_launch("convT2_fprop_kernel");
}

int convT2_dgrad(const b200_tensor* dy, const void* kernel, int cout, const b200_tensor* dx, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  B200_REQUIRE(dy->h == 2 * dx->h && dy->w == 2 * dx->w && dy->n == dx->n && dy->c == cout && dx->dtype == dy->dtype,
               B200_ERR_BAD_ARG, "convT2x2_dgrad: shape mismatch");
  {
    b200_tensor d00 = parity_view(dy, 0, 0);
    if (dx->dtype == B200_BF16 && conv_tc_supported(&d00, cout, dx->c, dx, 1)) {
      for (int ab = 0; ab < 4; ++ab) {
        b200_tensor dv = parity_view(dy, ab / 2, ab % 2);
        const __nv_bfloat16* k_ab = reinterpret_cast<const __nv_bfloat16*>(kernel) + (long long)ab * cout * dx->c;  // [K = cout][N = cin], MN-major
        int rc = conv_tc_launch(&dv, k_ab, cout, dx->c, 0, 1, nullptr, dx, B200_ACT_NONE, ab > 0 ? 1 : 0, st, nullptr, 1, ws, ws_bytes);
        if (rc) return rc;
      }
      return B200_OK;
    }
  }
  long long total = (long long)dx->n * dx->h * dx->w * dx->c;
  TView dv = view_of(dy), xv = view_of(dx);
  B200_DISPATCH_DTYPE(dx->dtype, T, {
    launch_pdl(convT2_dgrad_kernel<T>, grid_for(total), NT, 0, st, dv, (const T*)kernel, xv, total);
  });
  return check_launch("convT2_dgrad_kernel");
}

int convT2_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dk, float* dbias, cudaStream_t st) {
  B200_REQUIRE(dy->h == 2 * x->h && dy->w == 2 * x->w && dy->n == x->n && x->dtype == dy->dtype, B200_ERR_BAD_ARG,
               "convT2x2_wgrad: shape mismatch");
  const long long kcount = 4LL * dy->c * x->c;
  cudaMemsetAsync(dk, 0, sizeof(float) * kcount, st);
  b200_tensor d00 = parity_view(dy, 0, 0);
  if (x->dtype == B200_BF16 && wgrad_tc_supported(&d00, x, 1)) {
    // dK[a,b][o][c] = sum_pixels dy_ab[p][o] * x[p][c]: the 1x1 filter gradient with dy_ab in the activation role
    for (int ab = 0; ab < 4; ++ab) {
      b200_tensor dv = parity_view(dy, ab / 2, ab % 2);
      int rc = wgrad_tc_launch(&dv, x, dk + (long long)ab * dy->c * x->c, nullptr, 0, st, 1, 1);
      if (rc) return rc;
    }
  } else {
    TView xv = view_of(x), dv = view_of(dy);
    const long long npix = (long long)x->n * x->h * x->w;
    int splits = (int)((npix + 2047) / 2048);
    if (splits > 64) splits = 64;
    dim3 grid(4 * dy->c, splits);
    B200_DISPATCH_DTYPE(x->dtype, T, { launch_pdl(convT2_wgrad_kernel<T>, grid, NT, 0, st, xv, dv, dk); });
    int rc = check_launch("convT2_wgrad_kernel");
    if (rc) return rc;
  }
  if (dbias) return bias_act_bwd(dy, dy, B200_ACT_NONE, dy, dbias, st);   // column sums of dy, added to dbias
  return B200_OK;
}

}  // namespace b200
