// Channel LayerNorm, BatchNorm and bias/activation backward -- HBM-bound kernels.
//
// One thread owns 8 consecutive channels of one pixel (16 B of bf16); the C/8
// threads of a pixel sit in one warp (C <= 256) or loop (C > 256) and reduce with
// shuffles.  Replaces keras LayerNormalization(axis=-1)+ReLU
// (Super_resolution/code/train_adaptive_unet.py:203-204,208-209) and
// BatchNormalization+ReLU (Segmenation/code/train_adaptive_unet.py:327-331).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ void pix_decode(long long p, int H, int W, int& n, int& h, int& w) {
  w = (int)(p % W); p /= W;
  h = (int)(p % H);
  n = (int)(p / H);
}

__device__ __forceinline__ float group_sum(float v, int tpp) {
  for (int o = tpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// Per-channel sums over the pixel groups of a block without shared-memory float atomics (which are
// compare-and-swap loops and serialise badly when ppb groups hit the same channel): every thread parks
// its 8 channel values, tpp*8 threads add the ppb rows and issue ONE global atomic per channel.
// Thread layout: threadIdx.x = pixel_group * tpp + lane_g.  s_tmp holds NT*8 floats.
__device__ __forceinline__ void block_channel_sum(const float (&v)[8], int tpp, int chunk_base, int C,
                                                  float* s_tmp, float* __restrict__ dst) {
#pragma unroll
  for (int i = 0; i < 8; ++i) s_tmp[threadIdx.x * 8 + i] = v[i];
  __syncthreads();
  if (dst && (int)threadIdx.x < tpp * 8) {
    const int ppb = NT / tpp;
    float a = 0.f;
    for (int g = 0; g < ppb; ++g) a += s_tmp[g * tpp * 8 + threadIdx.x];
    const int c = chunk_base * 8 + threadIdx.x;   // chunk_base: first 8-channel chunk of lane_g == 0
    if (c < C) atomicAdd(dst + c, a);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// LayerNorm forward: TPP threads per pixel (power of two <= 32), CPT chunks/thread
// ---------------------------------------------------------------------------
template <typename T, int CPT>
__global__ void __launch_bounds__(NT)
ln_fwd_kernel(TView z, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
              TView y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int tpp, long long npix) {
  pdl_sync();
  const int lane_g = threadIdx.x % tpp;
  const int ppb = NT / tpp;
  const int C = z.c, chunks = C / 8;
  const T* zp = reinterpret_cast<const T*>(z.data);
  T* yp = reinterpret_cast<T*>(y.data);
  const float invC = 1.f / (float)C;
  float gam[CPT][8], bet[CPT][8];   // this thread's channels never change: keep gamma/beta in registers
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int j = lane_g + k * tpp;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      gam[k][i] = j < chunks ? gamma[j * 8 + i] : 0.f;
      bet[k][i] = j < chunks ? beta[j * 8 + i] : 0.f;
    }
  }
  for (long long base = (long long)blockIdx.x * ppb; base < npix; base += (long long)gridDim.x * ppb) {
    // warp-uniform trip count: groups past the end recompute the last pixel and skip their stores
    long long p = base + threadIdx.x / tpp;
    const bool valid = p < npix;
    if (!valid) p = npix - 1;
    const T* src = zp + pix_offset_flat(z, p);
    float v[CPT][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      int j = lane_g + k * tpp;
      if (j < chunks) {
        Vec8<T>::load(src + j * 8, v[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[k][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
      }
    }
    const float mu = group_sum(s, tpp) * invC;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      int j = lane_g + k * tpp;
      if (j < chunks) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float d = v[k][i] - mu; q += d * d; }
      }
    }
    const float var = group_sum(q, tpp) * invC;
    const float rs = rsqrtf(var + eps);
    T* dst = yp + pix_offset_flat(y, p);
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      int j = lane_g + k * tpp;
      if (j < chunks && valid) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = (v[k][i] - mu) * rs * gam[k][i] + bet[k][i];
          o[i] = relu ? fmaxf(t, 0.f) : t;
        }
        Vec8<T>::store(dst + j * 8, o);
      }
    }
    if (lane_g == 0 && valid) { mean_out[p] = mu; rstd_out[p] = rs; }
  }
}

// LayerNorm backward.  Per-channel partials live in registers across the
// grid-stride loop, then go through shared memory to one global atomic per block.
template <typename T, int CPT>
__global__ void __launch_bounds__(NT, CPT == 1 ? 3 : 1)
ln_bwd_kernel(TView dy, TView z, const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ gamma, const float* __restrict__ beta, int relu, TView dz,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int tpp,
              long long npix) {
  pdl_sync();
  __shared__ float s_acc[NT * 8];
  const int lane_g = threadIdx.x % tpp;
  const int ppb = NT / tpp;
  const int C = z.c, chunks = C / 8;
  const T* zp = reinterpret_cast<const T*>(z.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  T* dzp = reinterpret_cast<T*>(dz.data);
  const float invC = 1.f / (float)C;
  float a_g[CPT][8], a_b[CPT][8], a_z[CPT][8];
#pragma unroll
  for (int k = 0; k < CPT; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) { a_g[k][i] = 0.f; a_b[k][i] = 0.f; a_z[k][i] = 0.f; }

  float gam[CPT][8], bet[CPT][8];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int j = lane_g + k * tpp;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      gam[k][i] = j < chunks ? gamma[j * 8 + i] : 0.f;
      bet[k][i] = j < chunks ? beta[j * 8 + i] : 0.f;
    }
  }
  for (long long base = (long long)blockIdx.x * ppb; base < npix; base += (long long)gridDim.x * ppb) {
    // warp-uniform trip count: groups past the end recompute the last pixel and skip their stores
    long long p = base + threadIdx.x / tpp;
    const bool valid = p < npix;
    if (!valid) p = npix - 1;
    const T* zs = zp + pix_offset_flat(z, p);
    const T* ds = dyp + pix_offset_flat(dy, p);
    const float mu = mean[p], rs = rstd[p];
    float xh[CPT][8], g[CPT][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      int j = lane_g + k * tpp;
      if (j < chunks) {
        float zv[8], dv[8];
        Vec8<T>::load(zs + j * 8, zv);
        Vec8<T>::load(ds + j * 8, dv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = (zv[i] - mu) * rs;
          float ga = gam[k][i];
          float t = x * ga + bet[k][i];
          float d = (!valid || (relu && !(t > 0.f))) ? 0.f : dv[i];
          a_g[k][i] += d * x;
          a_b[k][i] += d;
          float gg = d * ga;
          xh[k][i] = x; g[k][i] = gg;
          s1 += gg; s2 += gg * x;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { xh[k][i] = 0.f; g[k][i] = 0.f; }
      }
    }
    const float m1 = group_sum(s1, tpp) * invC;
    const float m2 = group_sum(s2, tpp) * invC;
    T* dd = dzp + pix_offset_flat(dz, p);
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      int j = lane_g + k * tpp;
      if (j < chunks && valid) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = rs * (g[k][i] - m1 - xh[k][i] * m2);
          a_z[k][i] += o[i];
        }
        Vec8<T>::store(dd + j * 8, o);
      }
    }
  }
  // chunk k of lane_g covers channels (lane_g + k*tpp)*8 ..: reduce one (chunk row, quantity) at a time
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    block_channel_sum(a_g[k], tpp, k * tpp, C, s_acc, dgamma);
    block_channel_sum(a_b[k], tpp, k * tpp, C, s_acc, dbeta);
    block_channel_sum(a_z[k], tpp, k * tpp, C, s_acc, dbias);
  }
}

// LayerNorm backward, bf16, C = 8*TPP (64/128/256), evenly spaced pixels.  These kernels are bound by
// instruction issue, not by HBM (ncu: issue slots 58 % busy at 46 % of HBM peak for the generic kernel
// above), so this variant is written for instruction count: 32-bit pixel indices, packed f32x2
// arithmetic (FFMA2/FADD2/FMUL2 of sm_100), bf16 pairs unpacked with one shift / one mask, two
// pixels in flight per thread.
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float2 v) {
  __nv_bfloat162 h = __float22bfloat162_rn(v);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int TPP>
__global__ void __launch_bounds__(NT, 3)
ln_bwd_lean_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_sw, const __nv_bfloat16* __restrict__ z,
                   long long z_sw, const float* __restrict__ mean, const float* __restrict__ rstd,
                   const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                   __nv_bfloat16* __restrict__ dz, long long dz_sw, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, float* __restrict__ dbias, int npix) {
  pdl_sync();
  constexpr int C = 8 * TPP, PPB = NT / TPP;
  __shared__ float s_acc[NT * 8];
  const int lane_g = threadIdx.x % TPP;
  const float invC = 1.f / (float)C;
  float2 gam[4], bet[4], a_g[4], a_b[4], a_z[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    gam[i] = make_float2(gamma[lane_g * 8 + 2 * i], gamma[lane_g * 8 + 2 * i + 1]);
    bet[i] = make_float2(beta[lane_g * 8 + 2 * i], beta[lane_g * 8 + 2 * i + 1]);
    a_g[i] = a_b[i] = a_z[i] = make_float2(0.f, 0.f);
  }
  const __nv_bfloat16* zq = z + lane_g * 8;
  const __nv_bfloat16* dq = dy + lane_g * 8;
  __nv_bfloat16* oq = dz + lane_g * 8;
  const int step = gridDim.x * PPB;
  // two pixels per iteration (p and p + step); warp-uniform trip count, the tail pixel is masked
  for (int base = blockIdx.x * PPB; base < npix; base += 2 * step) {
    int pp[2] = {base + (int)(threadIdx.x / TPP), base + (int)(threadIdx.x / TPP) + step};
    bool ok[2];
    uint4 zr[2], dr[2];
    float mu[2], rs[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ok[u] = pp[u] < npix;
      if (!ok[u]) pp[u] = npix - 1;
      zr[u] = *reinterpret_cast<const uint4*>(zq + (long long)pp[u] * z_sw);
      dr[u] = *reinterpret_cast<const uint4*>(dq + (long long)pp[u] * dy_sw);
      mu[u] = mean[pp[u]];
      rs[u] = rstd[pp[u]];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t* zw = reinterpret_cast<const uint32_t*>(&zr[u]);
      const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dr[u]);
      const float2 rs2 = make_float2(rs[u], rs[u]);
      const float2 nmr = make_float2(-mu[u] * rs[u], -mu[u] * rs[u]);
      float2 x[4], gg[4];
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[i] = __ffma2_rn(unpack_bf16x2(zw[i]), rs2, nmr);            // (z - mu) * rstd
        float2 d = unpack_bf16x2(dw[i]);
        if (relu) {
          const float2 t = __ffma2_rn(x[i], gam[i], bet[i]);
          d.x = t.x > 0.f ? d.x : 0.f;
          d.y = t.y > 0.f ? d.y : 0.f;
        }
        if (!ok[u]) d = make_float2(0.f, 0.f);
        a_g[i] = __ffma2_rn(d, x[i], a_g[i]);
        a_b[i] = __fadd2_rn(a_b[i], d);
        gg[i] = __fmul2_rn(d, gam[i]);
        s1 = __fadd2_rn(s1, gg[i]);
        s2 = __ffma2_rn(gg[i], x[i], s2);
      }
      float m1 = s1.x + s1.y, m2 = s2.x + s2.y;
#pragma unroll
      for (int o = TPP >> 1; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      }
      // dz = rstd * (g - mean(g) - xhat * mean(g * xhat))
      const float2 c1 = make_float2(-m1 * invC * rs[u], -m1 * invC * rs[u]);
      const float2 c2 = make_float2(-m2 * invC * rs[u], -m2 * invC * rs[u]);
      uint4 out;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 o2 = __ffma2_rn(x[i], c2, __ffma2_rn(gg[i], rs2, c1));
        a_z[i] = __fadd2_rn(a_z[i], o2);     // masked tail pixels have d == 0 for the whole group => o2 == 0
        ow[i] = pack_bf16x2(o2);
      }
      if (ok[u]) *reinterpret_cast<uint4*>(oq + (long long)pp[u] * dz_sw) = out;
    }
  }
  float v[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = a_g[i].x; v[2 * i + 1] = a_g[i].y; }
  block_channel_sum(v, TPP, 0, C, s_acc, dgamma);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = a_b[i].x; v[2 * i + 1] = a_b[i].y; }
  block_channel_sum(v, TPP, 0, C, s_acc, dbeta);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = a_z[i].x; v[2 * i + 1] = a_z[i].y; }
  block_channel_sum(v, TPP, 0, C, s_acc, dbias);
}

// ---------------------------------------------------------------------------
// Generic element-wise backward of bias+activation (scalar path, any C)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
bias_act_bwd_kernel(TView dy, TView y, int act, TView dz, float* __restrict__ dbias, long long total) {
  pdl_sync();
  extern __shared__ float s_acc[];  // [C]
  const int C = y.c;
  for (int i = threadIdx.x; i < C; i += NT) s_acc[i] = 0.f;
  __syncthreads();
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  const T* yp = reinterpret_cast<const T*>(y.data);
  T* dzp = reinterpret_cast<T*>(dz.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % C);
    int n, h, w;
    pix_decode(i / C, y.h, y.w, n, h, w);
    float d = ldf(dyp + pix_offset(dy, n, h, w) + c);
    float yv = ldf(yp + pix_offset(y, n, h, w) + c);
    if (act == B200_ACT_RELU) d = yv > 0.f ? d : 0.f;
    else if (act == B200_ACT_SIGMOID) d = d * yv * (1.f - yv);
    stf(dzp + pix_offset(dz, n, h, w) + c, d);
    if (dbias) atomicAdd(&s_acc[c], d);
  }
  __syncthreads();
  if (dbias)
    for (int i = threadIdx.x; i < C; i += NT) atomicAdd(dbias + i, s_acc[i]);
}

// narrow tensors (C <= 4: the RGB / mask heads): one thread per pixel, warp-reduced channel sums
template <typename T, int C>
__global__ void __launch_bounds__(NT)
bias_act_bwd_narrow_kernel(TView dy, TView y, int act, TView dz, float* __restrict__ dbias, long long npix) {
  pdl_sync();
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  const T* yp = reinterpret_cast<const T*>(y.data);
  T* dzp = reinterpret_cast<T*>(dz.data);
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (long long p = (long long)blockIdx.x * NT + threadIdx.x; p < npix; p += (long long)gridDim.x * NT) {
    const long long o1 = pix_offset_flat(dy, p), o2 = pix_offset_flat(y, p), o3 = pix_offset_flat(dz, p);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float d = ldf(dyp + o1 + c);
      const float yv = ldf(yp + o2 + c);
      if (act == B200_ACT_RELU) d = yv > 0.f ? d : 0.f;
      else if (act == B200_ACT_SIGMOID) d = d * yv * (1.f - yv);
      stf(dzp + o3 + c, d);
      acc[c] += d;
    }
  }
  if (dbias) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = warp_sum(acc[c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(dbias + c, v);
    }
  }
}

// vectorised variant: C % 8 == 0, C/8 a power of two <= NT
template <typename T>
__global__ void __launch_bounds__(NT)
bias_act_bwd_vec_kernel(TView dy, TView y, int act, TView dz, float* __restrict__ dbias, long long npix) {
  pdl_sync();
  extern __shared__ float s_acc[];
  const int C = y.c, chunks = C / 8;
  for (int i = threadIdx.x; i < C; i += NT) s_acc[i] = 0.f;
  __syncthreads();
  const int j = threadIdx.x % chunks;
  const int ppb = NT / chunks;
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  const T* yp = reinterpret_cast<const T*>(y.data);
  T* dzp = reinterpret_cast<T*>(dz.data);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (long long p = (long long)blockIdx.x * ppb + threadIdx.x / chunks; p < npix; p += (long long)gridDim.x * ppb) {
    float d[8], yv[8];
    Vec8<T>::load(dyp + pix_offset_flat(dy, p) + j * 8, d);
    Vec8<T>::load(yp + pix_offset_flat(y, p) + j * 8, yv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (act == B200_ACT_RELU) d[i] = yv[i] > 0.f ? d[i] : 0.f;
      else if (act == B200_ACT_SIGMOID) d[i] = d[i] * yv[i] * (1.f - yv[i]);
      acc[i] += d[i];
    }
    Vec8<T>::store(dzp + pix_offset_flat(dz, p) + j * 8, d);
  }
  if (dbias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_acc[j * 8 + i], acc[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += NT) atomicAdd(dbias + i, s_acc[i]);
  }
}

// ---------------------------------------------------------------------------
// BatchNorm
// ---------------------------------------------------------------------------
// mode 0: sums of z and z^2.  mode 1: sums of g=dy*mask and g*xhat (needs mean/rstd).
template <typename T, int MODE>
__global__ void __launch_bounds__(NT)
bn_stats_kernel(TView z, TView dy, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                double* __restrict__ stats, long long npix) {
  pdl_sync();
  extern __shared__ float s_acc[];  // [2][C]
  const int C = z.c;
  for (int i = threadIdx.x; i < 2 * C; i += NT) s_acc[i] = 0.f;
  __syncthreads();
  const T* zp = reinterpret_cast<const T*>(z.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  // thread -> fixed channel set: c = threadIdx.x % C when C <= NT, else loop
  const int cstep = C < NT ? C : NT;
  const int ppb = C < NT ? NT / C : 1;
  const long long chunk = 64;  // pixels per block iteration group
  for (long long p0 = (long long)blockIdx.x * chunk; p0 < npix; p0 += (long long)gridDim.x * chunk) {
    for (int c = threadIdx.x % cstep; c < C; c += cstep) {
      float a = 0.f, b = 0.f;
      for (long long p = p0 + threadIdx.x / cstep; threadIdx.x / cstep < ppb && p < p0 + chunk && p < npix; p += ppb) {
        int n, h, w;
        pix_decode(p, z.h, z.w, n, h, w);
        float zv = ldf(zp + pix_offset(z, n, h, w) + c);
        if (MODE == 0) { a += zv; b += zv * zv; }
        else {
          float x = (zv - mean[c]) * rstd[c];
          float t = x * gamma[c] + beta[c];
          float d = ldf(dyp + pix_offset(dy, n, h, w) + c);
          if (relu && !(t > 0.f)) d = 0.f;
          a += d; b += d * x;
        }
      }
      atomicAdd(&s_acc[c], a);
      atomicAdd(&s_acc[C + c], b);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += NT) atomicAdd(stats + i, (double)s_acc[i]);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double count, float eps, float momentum,
                                   float* save_mean, float* save_rstd, float* moving_mean, float* moving_var) {
  pdl_sync();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = stats[c] / count;
  double var = stats[C + c] / count - m * m;
  if (var < 0) var = 0;
  save_mean[c] = (float)m;
  save_rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (moving_mean) moving_mean[c] = moving_mean[c] * momentum + (float)m * (1.f - momentum);
  if (moving_var) moving_var[c] = moving_var[c] * momentum + (float)var * (1.f - momentum);
}

template <typename T>
__global__ void __launch_bounds__(NT)
bn_apply_kernel(TView z, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, int relu, TView y,
                long long total, const float* __restrict__ mmean, const float* __restrict__ mvar, float eps) {
  pdl_sync();
  const int C = z.c;
  const T* zp = reinterpret_cast<const T*>(z.data);
  T* yp = reinterpret_cast<T*>(y.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % C);
    int n, h, w;
    pix_decode(i / C, z.h, z.w, n, h, w);
    float mu = mean ? mean[c] : mmean[c];
    float rs = rstd ? rstd[c] : rsqrtf(mvar[c] + eps);
    float t = (ldf(zp + pix_offset(z, n, h, w) + c) - mu) * rs * gamma[c] + beta[c];
    if (relu) t = fmaxf(t, 0.f);
    stf(yp + pix_offset(y, n, h, w) + c, t);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
bn_bwd_apply_kernel(TView dy, TView z, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, int relu, TView dz,
                    const double* __restrict__ stats, double count, long long total) {
  pdl_sync();
  const int C = z.c;
  const T* zp = reinterpret_cast<const T*>(z.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);
  T* dzp = reinterpret_cast<T*>(dz.data);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    int c = (int)(i % C);
    int n, h, w;
    pix_decode(i / C, z.h, z.w, n, h, w);
    float x = (ldf(zp + pix_offset(z, n, h, w) + c) - mean[c]) * rstd[c];
    float t = x * gamma[c] + beta[c];
    float d = ldf(dyp + pix_offset(dy, n, h, w) + c);
    if (relu && !(t > 0.f)) d = 0.f;
    float sb = (float)(stats[c] / count), sg = (float)(stats[C + c] / count);
    stf(dzp + pix_offset(dz, n, h, w) + c, gamma[c] * rstd[c] * (d - sb - x * sg));
  }
}

__global__ void bn_bwd_params_kernel(const double* __restrict__ stats, int C, float* dgamma, float* dbeta) {
  pdl_sync();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] += (float)stats[c];
  if (dgamma) dgamma[c] += (float)stats[C + c];
}

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = 8LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int layernorm_fwd(const b200_tensor* z, const float* gamma, const float* beta, float eps, int relu,
                  const b200_tensor* y, float* mean, float* rstd, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, y) && z->dtype == y->dtype, B200_ERR_BAD_ARG, "layernorm_fwd: shape/dtype mismatch");
  const int C = z->c;
  B200_REQUIRE(C % 8 == 0 && vec_aligned(z, 8) && vec_aligned(y, 8), B200_ERR_UNSUPPORTED,
               "layernorm_fwd: C=%d must be a multiple of 8 with 16-byte aligned pixels", C);
  const int chunks = C / 8;
  int tpp = 1;
  while (tpp < chunks && tpp < 32) tpp <<= 1;
  const int cpt = (chunks + tpp - 1) / tpp;
  B200_REQUIRE(cpt <= 8, B200_ERR_UNSUPPORTED, "layernorm_fwd: C=%d too wide (max 2048)", C);
  const long long npix = (long long)z->n * z->h * z->w;
  const int grid = grid_for(npix, NT / tpp);
  TView zv = view_of(z), yv = view_of(y);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    if (cpt == 1) launch_pdl(ln_fwd_kernel<T, 1>, grid, NT, 0, st, zv, gamma, beta, eps, relu, yv, mean, rstd, tpp, npix);
    else if (cpt == 2) launch_pdl(ln_fwd_kernel<T, 2>, grid, NT, 0, st, zv, gamma, beta, eps, relu, yv, mean, rstd, tpp, npix);
    else if (cpt <= 4) launch_pdl(ln_fwd_kernel<T, 4>, grid, NT, 0, st, zv, gamma, beta, eps, relu, yv, mean, rstd, tpp, npix);
    else launch_pdl(ln_fwd_kernel<T, 8>, grid, NT, 0, st, zv, gamma, beta, eps, relu, yv, mean, rstd, tpp, npix);
  });
  return check_launch("ln_fwd_kernel");
}

int layernorm_bwd(const b200_tensor* dy, const b200_tensor* z, const float* mean, const float* rstd,
                  const float* gamma, const float* beta, int relu, const b200_tensor* dz, float* dgamma,
                  float* dbeta, float* dbias, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, dy) && same_shape(z, dz) && z->dtype == dy->dtype && z->dtype == dz->dtype,
               B200_ERR_BAD_ARG, "layernorm_bwd: shape/dtype mismatch");
  const int C = z->c;
  B200_REQUIRE(C % 8 == 0 && vec_aligned(z, 8) && vec_aligned(dy, 8) && vec_aligned(dz, 8), B200_ERR_UNSUPPORTED,
               "layernorm_bwd: C=%d must be a multiple of 8 with 16-byte aligned pixels", C);
  const int chunks = C / 8;
  int tpp = 1;
  while (tpp < chunks && tpp < 32) tpp <<= 1;
  const int cpt = (chunks + tpp - 1) / tpp;
  B200_REQUIRE(cpt <= 8, B200_ERR_UNSUPPORTED, "layernorm_bwd: C=%d too wide (max 2048)", C);
  const long long npix = (long long)z->n * z->h * z->w;
  long long blocks = (npix + (NT / tpp) - 1) / (NT / tpp);
  long long cap = 6LL * sm_count();
  const int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
  const size_t smem = sizeof(float) * 3 * C;
  TView dyv = view_of(dy), zv = view_of(z), dzv = view_of(dz);
  static const int lean_variant = getenv("B200_LN_BWD_LEAN") ? atoi(getenv("B200_LN_BWD_LEAN")) : 1;
  if (lean_variant && z->dtype == B200_BF16 && (C == 64 || C == 128 || C == 256) && zv.lin && dyv.lin && dzv.lin &&
      npix < (1LL << 30)) {
    const int ppb = NT / (C / 8);
    // at least eight pixels per thread: each block ends with 3*C global atomics on the same addresses
    long long nb = (npix + 8LL * ppb - 1) / (8LL * ppb);
    const long long capl = 3LL * sm_count();
    const int gridl = (int)(nb < capl ? nb : capl);
    const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy->data);
    const __nv_bfloat16* zp = reinterpret_cast<const __nv_bfloat16*>(z->data);
    __nv_bfloat16* dzp = reinterpret_cast<__nv_bfloat16*>(dz->data);
    if (C == 64) launch_pdl(ln_bwd_lean_kernel<8>, gridl, NT, 0, st, dyp, dy->stride_w, zp, z->stride_w, mean, rstd, gamma, beta, relu, dzp, dz->stride_w, dgamma, dbeta, dbias, (int)npix);
    else if (C == 128) launch_pdl(ln_bwd_lean_kernel<16>, gridl, NT, 0, st, dyp, dy->stride_w, zp, z->stride_w, mean, rstd, gamma, beta, relu, dzp, dz->stride_w, dgamma, dbeta, dbias, (int)npix);
    else launch_pdl(ln_bwd_lean_kernel<32>, gridl, NT, 0, st, dyp, dy->stride_w, zp, z->stride_w, mean, rstd, gamma, beta, relu, dzp, dz->stride_w, dgamma, dbeta, dbias, (int)npix);
    return check_launch("ln_bwd_lean_kernel");
  }
  B200_DISPATCH_DTYPE(z->dtype, T, {
    if (cpt == 1) launch_pdl(ln_bwd_kernel<T, 1>, grid, NT, smem, st, dyv, zv, mean, rstd, gamma, beta, relu, dzv, dgamma, dbeta, dbias, tpp, npix);
    else if (cpt == 2) launch_pdl(ln_bwd_kernel<T, 2>, grid, NT, smem, st, dyv, zv, mean, rstd, gamma, beta, relu, dzv, dgamma, dbeta, dbias, tpp, npix);
    else if (cpt <= 4) launch_pdl(ln_bwd_kernel<T, 4>, grid, NT, smem, st, dyv, zv, mean, rstd, gamma, beta, relu, dzv, dgamma, dbeta, dbias, tpp, npix);
    else launch_pdl(ln_bwd_kernel<T, 8>, grid, NT, smem, st, dyv, zv, mean, rstd, gamma, beta, relu, dzv, dgamma, dbeta, dbias, tpp, npix);
  });
  return check_launch("ln_bwd_kernel");
}

int bias_sum_c3(const b200_tensor* dy, float* dbias, cudaStream_t st);

int bias_act_bwd(const b200_tensor* dy, const b200_tensor* y, int act, const b200_tensor* dz, float* dbias,
                 cudaStream_t st) {
  B200_REQUIRE(same_shape(y, dy) && same_shape(y, dz) && y->dtype == dy->dtype && y->dtype == dz->dtype,
               B200_ERR_BAD_ARG, "bias_act_bwd: shape/dtype mismatch");
  {  // linear layer, in place: nothing to write, the call only wants the bias gradient (the RGB head)
    TView dv = view_of(dy);
    if (act == B200_ACT_NONE && dz->data == dy->data && dy->c == 3 && dy->dtype == B200_BF16 && dv.lin && dv.sw == 3 &&
        reinterpret_cast<uintptr_t>(dy->data) % 16 == 0) {
      if (!dbias) return B200_OK;
      return bias_sum_c3(dy, dbias, st);
    }
  }
  const int C = y->c;
  const long long npix = (long long)y->n * y->h * y->w;
  TView dyv = view_of(dy), yv = view_of(y), dzv = view_of(dz);
  const size_t smem = sizeof(float) * C;
  const bool vec = C % 8 == 0 && is_pow2(C / 8) && C / 8 <= NT && vec_aligned(dy, 8) && vec_aligned(y, 8) &&
                   vec_aligned(dz, 8);
  B200_DISPATCH_DTYPE(y->dtype, T, {
    if (C == 1 || C == 3) {
      const int grid = grid_for(npix, NT);
      if (C == 1) launch_pdl(bias_act_bwd_narrow_kernel<T, 1>, grid, NT, 0, st, dyv, yv, act, dzv, dbias, npix);
      else launch_pdl(bias_act_bwd_narrow_kernel<T, 3>, grid, NT, 0, st, dyv, yv, act, dzv, dbias, npix);
    } else if (vec) {
      long long blocks = (npix + (NT / (C / 8)) - 1) / (NT / (C / 8));
      long long cap = 4LL * sm_count();
      int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
      launch_pdl(bias_act_bwd_vec_kernel<T>, grid, NT, smem, st, dyv, yv, act, dzv, dbias, npix);
    } else {
      long long total = npix * C;
      int grid = grid_for(total, NT);
      launch_pdl(bias_act_bwd_kernel<T>, grid, NT, smem, st, dyv, yv, act, dzv, dbias, total);
    }
  });
  return check_launch("bias_act_bwd_kernel");
}

int batchnorm_fwd_train(const b200_tensor* z, const float* gamma, const float* beta, float eps, float momentum,
                        int relu, const b200_tensor* y, float* save_mean, float* save_rstd, float* moving_mean,
                        float* moving_var, double* stats_ws, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, y) && z->dtype == y->dtype, B200_ERR_BAD_ARG, "batchnorm_fwd: shape/dtype mismatch");
  const int C = z->c;
  const long long npix = (long long)z->n * z->h * z->w;
  cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * C, st);
  TView zv = view_of(z), yv = view_of(y);
  const int sgrid = grid_for(npix, 64);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    launch_pdl(bn_stats_kernel<T, 0>, sgrid, NT, sizeof(float) * 2 * C, st, zv, zv, nullptr, nullptr, nullptr, nullptr, 0, stats_ws, npix);
    launch_pdl(bn_finalize_kernel, (C + 127) / 128, 128, 0, st, stats_ws, C, (double)npix, eps, momentum, save_mean, save_rstd, moving_mean, moving_var);
    launch_pdl(bn_apply_kernel<T>, grid_for(npix * C, NT), NT, 0, st, zv, save_mean, save_rstd, gamma, beta, relu, yv, npix * C, nullptr, nullptr, eps);
  });
  count_launches(2);
  return check_launch("batchnorm_fwd_train");
}

// ---- BatchNorm in phases (synchronised BatchNorm under data parallelism) -------------------------------------------
// stats -> [caller all-reduces the 2*C doubles over the ranks] -> apply with the GLOBAL pixel count.  The backward
// statistics kernel also adds the LOCAL sums to dgamma / dbeta (they join the ordinary gradient exchange).
int batchnorm_stats(const b200_tensor* z, const b200_tensor* dy, const float* save_mean, const float* save_rstd,
                    const float* gamma, const float* beta, int relu, double* stats_ws, float* dgamma, float* dbeta,
                    cudaStream_t st) {
  const int C = z->c;
  const long long npix = (long long)z->n * z->h * z->w;
  cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * C, st);
  TView zv = view_of(z);
  if (!dy) {
    B200_DISPATCH_DTYPE(z->dtype, T, {
      launch_pdl(bn_stats_kernel<T, 0>, grid_for(npix, 64), NT, sizeof(float) * 2 * C, st, zv, zv, nullptr, nullptr, nullptr, nullptr, 0, stats_ws, npix);
    });
    return check_launch("bn_stats_kernel");
  }
  B200_REQUIRE(same_shape(z, dy) && z->dtype == dy->dtype && save_mean && save_rstd && gamma && beta, B200_ERR_BAD_ARG,
               "batchnorm_stats (backward): dy / statistics missing or mismatched");
  TView dyv = view_of(dy);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    launch_pdl(bn_stats_kernel<T, 1>, grid_for(npix, 64), NT, sizeof(float) * 2 * C, st, zv, dyv, save_mean, save_rstd, gamma, beta, relu, stats_ws, npix);
  });
  launch_pdl(bn_bwd_params_kernel, (C + 127) / 128, 128, 0, st, stats_ws, C, dgamma, dbeta);
  count_launches(1);
  return check_launch("bn_stats_kernel");
}

int batchnorm_fwd_apply(const b200_tensor* z, const float* gamma, const float* beta, float eps, float momentum, int relu,
                        const b200_tensor* y, float* save_mean, float* save_rstd, float* moving_mean, float* moving_var,
                        const double* stats_ws, double count, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, y) && z->dtype == y->dtype && count > 0, B200_ERR_BAD_ARG, "batchnorm_fwd_apply: bad argument");
  const int C = z->c;
  const long long npix = (long long)z->n * z->h * z->w;
  TView zv = view_of(z), yv = view_of(y);
  launch_pdl(bn_finalize_kernel, (C + 127) / 128, 128, 0, st, stats_ws, C, count, eps, momentum, save_mean, save_rstd, moving_mean, moving_var);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    launch_pdl(bn_apply_kernel<T>, grid_for(npix * C, NT), NT, 0, st, zv, save_mean, save_rstd, gamma, beta, relu, yv, npix * C, nullptr, nullptr, eps);
  });
  count_launches(1);
  return check_launch("batchnorm_fwd_apply");
}

int batchnorm_bwd_apply(const b200_tensor* dy, const b200_tensor* z, const float* save_mean, const float* save_rstd,
                        const float* gamma, const float* beta, int relu, const b200_tensor* dz, const double* stats_ws,
                        double count, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, dy) && same_shape(z, dz) && z->dtype == dy->dtype && z->dtype == dz->dtype && count > 0,
               B200_ERR_BAD_ARG, "batchnorm_bwd_apply: bad argument");
  const long long npix = (long long)z->n * z->h * z->w;
  TView zv = view_of(z), dyv = view_of(dy), dzv = view_of(dz);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    launch_pdl(bn_bwd_apply_kernel<T>, grid_for(npix * z->c, NT), NT, 0, st, dyv, zv, save_mean, save_rstd, gamma, beta, relu, dzv, stats_ws, count, npix * z->c);
  });
  return check_launch("batchnorm_bwd_apply");
}

int batchnorm_fwd_infer(const b200_tensor* z, const float* gamma, const float* beta, float eps, int relu,
                        const float* moving_mean, const float* moving_var, const b200_tensor* y, cudaStream_t st) {
  B200_REQUIRE(same_shape(z, y) && z->dtype == y->dtype, B200_ERR_BAD_ARG, "batchnorm_infer: shape/dtype mismatch");
  const long long total = (long long)z->n * z->h * z->w * z->c;
  TView zv = view_of(z), yv = view_of(y);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    launch_pdl(bn_apply_kernel<T>, grid_for(total, NT), NT, 0, st, zv, nullptr, nullptr, gamma, beta, relu, yv, total, moving_mean, moving_var, eps);
  });
  return check_launch("batchnorm_fwd_infer");
}

int batchnorm_bwd(const b200_tensor* dy, const b200_tensor* z, const float* save_mean, const float* save_rstd,
                  const float* gamma, const float* beta, int relu, const b200_tensor* dz, float* dgamma,
                  float* dbeta, float* dbias, double* stats_ws, cudaStream_t st) {
  (void)dbias;  // sum(dz) over N,H,W is identically zero after BatchNorm
  B200_REQUIRE(same_shape(z, dy) && same_shape(z, dz) && z->dtype == dy->dtype && z->dtype == dz->dtype,
               B200_ERR_BAD_ARG, "batchnorm_bwd: shape/dtype mismatch");
  const int C = z->c;
  const long long npix = (long long)z->n * z->h * z->w;
  cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * C, st);
  TView zv = view_of(z), dyv = view_of(dy), dzv = view_of(dz);
  B200_DISPATCH_DTYPE(z->dtype, T, {
    launch_pdl(bn_stats_kernel<T, 1>, grid_for(npix, 64), NT, sizeof(float) * 2 * C, st, zv, dyv, save_mean, save_rstd, gamma, beta, relu, stats_ws, npix);
    launch_pdl(bn_bwd_apply_kernel<T>, grid_for(npix * C, NT), NT, 0, st, dyv, zv, save_mean, save_rstd, gamma, beta, relu, dzv, stats_ws, (double)npix, npix * C);
    launch_pdl(bn_bwd_params_kernel, (C + 127) / 128, 128, 0, st, stats_ws, C, dgamma, dbeta);
  });
  count_launches(2);
  return check_launch("batchnorm_bwd");
}

}  // namespace b200
