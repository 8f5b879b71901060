// Specialised SIMT kernels for the two convolution shapes tensor cores cannot help:
//   * the RGB stem: 3x3 conv with Cin <= 4 (K = 27), fprop and wgrad   -- FMA/LDS bound
//   * the 1x1 heads with Cout <= 4 (residual_rgb 64->3, lesion_mask 64->1): fprop, dgrad, wgrad -- HBM bound
// (the stem needs no dgrad: the network input has no gradient).
//
// Replaces keras Conv2D at Super_resolution/code/train_adaptive_unet.py:202 (first call) and :267-274,
// Segmenation/code/train_adaptive_unet.py:326 (first call) and :361 (reference: /root/reference).
#include "common.cuh"

namespace b200 {
namespace {

// ---------------------------------------------------------------------------------------------
// Stem fprop: tile of 16x16 pixels per block, 256 threads = 128 pixel pairs x 2 halves of 32 couts.
// Weights [27][64] sit in shared memory and are read as warp-wide broadcasts.
// ---------------------------------------------------------------------------------------------
constexpr int ST_T = 16;           // tile edge
constexpr int ST_IW = ST_T + 2;

template <typename T>
__global__ void __launch_bounds__(256)
stem_fprop_kernel(TView x, const T* __restrict__ wgt, const float* __restrict__ bias, TView y, int act, int tiles_w,
                  int tiles_h) {
  pdl_sync();
  __shared__ float s_x[3][ST_IW][ST_IW + 1];
  __shared__ __align__(16) float s_w[27][64];
  __shared__ float s_b[64];
  const int tid = threadIdx.x;
  int bt = blockIdx.x;
  const int tw = bt % tiles_w; bt /= tiles_w;
  const int th = bt % tiles_h;
  const int n = bt / tiles_h;
  const int h0 = th * ST_T, w0 = tw * ST_T;
  const int co0 = blockIdx.y * 64;
  const int Cin = x.c, Cout = y.c;
  const T* xp = reinterpret_cast<const T*>(x.data);

  for (int i = tid; i < 3 * ST_IW * ST_IW; i += 256) {
    const int c = i % 3, pix = i / 3;
    const int r = pix / ST_IW, q = pix % ST_IW;
    const int ih = h0 + r - 1, iw = w0 + q - 1;
    float v = 0.f;
    if (c < Cin && ih >= 0 && ih < x.h && iw >= 0 && iw < x.w) v = ldf(xp + pix_offset(x, n, ih, iw) + c);
    s_x[c][r][q] = v;
  }
  for (int i = tid; i < 27 * 64; i += 256) {
    const int o = i % 64, k = i / 64;          // k = tap*3 + c
    const int tap = k / 3, c = k % 3;
    float v = 0.f;
    if (c < Cin && co0 + o < Cout) v = ldf(wgt + ((long long)tap * Cin + c) * Cout + co0 + o);
    s_w[k][o] = v;
  }
  if (tid < 64) s_b[tid] = (bias && co0 + tid < Cout) ? bias[co0 + tid] : 0.f;
  __syncthreads();

  const int half = tid / 128;                  // warp-uniform: which 32 couts
  const int pp = tid % 128;                    // pixel pair: row pp/8, cols 2*(pp%8), +1
  const int pr = pp / 8, pc = (pp % 8) * 2;
  float acc[2][32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { acc[0][j] = s_b[half * 32 + j]; acc[1][j] = acc[0][j]; }
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float x0 = s_x[c][pr + kh][pc + kw], x1 = s_x[c][pr + kh][pc + kw + 1];
        const float4* w4 = reinterpret_cast<const float4*>(&s_w[(kh * 3 + kw) * 3 + c][half * 32]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = w4[j];
          acc[0][4 * j + 0] += x0 * w.x; acc[0][4 * j + 1] += x0 * w.y; acc[0][4 * j + 2] += x0 * w.z; acc[0][4 * j + 3] += x0 * w.w;
          acc[1][4 * j + 0] += x1 * w.x; acc[1][4 * j + 1] += x1 * w.y; acc[1][4 * j + 2] += x1 * w.z; acc[1][4 * j + 3] += x1 * w.w;
        }
      }
  T* yp = reinterpret_cast<T*>(y.data);
  const int oh = h0 + pr;
  if (oh >= y.h) return;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int ow = w0 + pc + p;
    if (ow >= y.w) continue;
    T* dst = yp + pix_offset(y, n, oh, ow) + co0 + half * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = acc[p][g * 8 + i];
        o[i] = act == B200_ACT_RELU ? fmaxf(v, 0.f) : v;
      }
      if (co0 + half * 32 + g * 8 + 8 <= Cout) Vec8<T>::store(dst + g * 8, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stem wgrad: dW[k = tap*3+c][co] = sum_pixels patch[k] * dz[co].  256 threads = 4 pixel streams x
// (16 groups of 4 couts) x (4 groups of 7 k); register tile 7 x 4 per thread; per-block reduction in
// shared memory, then one atomic per output per block.
// ---------------------------------------------------------------------------------------------
constexpr int SW_TH = 8, SW_TW = 16;      // 128 pixels per tile

template <typename T>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(TView x, TView dy, float* __restrict__ dw, int tiles_w, int tiles_h, int total_tiles) {
  pdl_sync();
  __shared__ float s_x[3][SW_TH + 2][SW_TW + 3];
  __shared__ __align__(16) float s_dz[SW_TH * SW_TW][64];
  const int tid = threadIdx.x;
  const int stream = tid / 64;
  const int cg = tid % 16, kg = (tid % 64) / 16;
  const int co0 = blockIdx.y * 64;
  const int Cin = x.c, Cout = dy.c;
  const T* xp = reinterpret_cast<const T*>(x.data);
  const T* dp = reinterpret_cast<const T*>(dy.data);
  float acc[7][4];
#pragma unroll
  for (int a = 0; a < 7; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
  // this thread's (kh, kw, c) triples
  int t_kh[7], t_kw[7], t_c[7];
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    int k = kg * 7 + a;
    if (k > 26) k = 26;                   // the 28th slot duplicates k=26 and is dropped at the end
    t_kh[a] = (k / 3) / 3; t_kw[a] = (k / 3) % 3; t_c[a] = k % 3;
  }
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int bt = tile;
    const int tw = bt % tiles_w; bt /= tiles_w;
    const int th = bt % tiles_h;
    const int n = bt / tiles_h;
    const int h0 = th * SW_TH, w0 = tw * SW_TW;
    __syncthreads();
    for (int i = tid; i < 3 * (SW_TH + 2) * (SW_TW + 2); i += 256) {
      const int c = i % 3, pix = i / 3;
      const int r = pix / (SW_TW + 2), q = pix % (SW_TW + 2);
      const int ih = h0 + r - 1, iw = w0 + q - 1;
      float v = 0.f;
      if (c < Cin && ih >= 0 && ih < x.h && iw >= 0 && iw < x.w) v = ldf(xp + pix_offset(x, n, ih, iw) + c);
      s_x[c][r][q] = v;
    }
    for (int i = tid; i < SW_TH * SW_TW * 8; i += 256) {   // 16-byte loads: 8 channels per thread
      const int o8 = i % 8, pix = i / 8;
      const int oh = h0 + pix / SW_TW, ow = w0 + pix % SW_TW;
      float v[8];
      if (oh < dy.h && ow < dy.w && co0 + o8 * 8 + 8 <= Cout) Vec8<T>::load(dp + pix_offset(dy, n, oh, ow) + co0 + o8 * 8, v);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
      }
      *reinterpret_cast<float4*>(&s_dz[pix][o8 * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&s_dz[pix][o8 * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    for (int p = stream; p < SW_TH * SW_TW; p += 4) {
      const int pr = p / SW_TW, pc = p % SW_TW;
      const float4 d = *reinterpret_cast<const float4*>(&s_dz[p][cg * 4]);
#pragma unroll
      for (int a = 0; a < 7; ++a) {
        const float xv = s_x[t_c[a]][pr + t_kh[a]][pc + t_kw[a]];
        acc[a][0] += xv * d.x; acc[a][1] += xv * d.y; acc[a][2] += xv * d.z; acc[a][3] += xv * d.w;
      }
    }
  }
  // reduce the 4 streams through shared memory (reuse s_dz as [27][64] accumulator)
  __syncthreads();
  float* red = &s_dz[0][0];
  for (int i = tid; i < 27 * 64; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    const int k = kg * 7 + a;
    if (k <= 26) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&red[k * 64 + cg * 4 + j], acc[a][j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < 27 * 64; i += 256) {
    const int o = i % 64, k = i / 64;
    const int tap = k / 3, c = k % 3;
    if (c < Cin && co0 + o < Cout) atomicAdd(dw + ((long long)tap * Cin + c) * Cout + co0 + o, red[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// 1x1 heads with COUT <= 4: one thread per pixel (fprop, dgrad), 8 threads per pixel (wgrad).
// ---------------------------------------------------------------------------------------------
// C/8 threads per pixel (each 8 channels = one 16-byte load), shuffle reduction, lane 0 of the group stores
template <typename T, int COUT>
__global__ void __launch_bounds__(256)
head_fprop_kernel(TView x, const T* __restrict__ wgt, const float* __restrict__ bias, TView y, int act, long long npix) {
  pdl_sync();
  const int Cin = x.c, chunks = Cin / 8;
  const int j = threadIdx.x % chunks, ppb = 256 / chunks;
  float w[8][COUT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) w[i][o] = ldf(wgt + (j * 8 + i) * COUT + o);
  const T* xp = reinterpret_cast<const T*>(x.data);
  T* yp = reinterpret_cast<T*>(y.data);
  for (long long base = (long long)blockIdx.x * ppb; base < npix; base += (long long)gridDim.x * ppb) {
    long long p = base + threadIdx.x / chunks;      // warp-uniform trip count (shuffles below)
    const bool valid = p < npix;
    if (!valid) p = npix - 1;
    float v[8];
    Vec8<T>::load(xp + pix_offset_flat(x, p) + j * 8, v);
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a += v[i] * w[i][o];
      for (int s = chunks >> 1; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      acc[o] = a;
    }
    if (j == 0 && valid) {
      T* dst = yp + pix_offset_flat(y, p);
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        float r = acc[o] + (bias ? bias[o] : 0.f);
        if (act == B200_ACT_RELU) r = fmaxf(r, 0.f);
        else if (act == B200_ACT_SIGMOID) r = 1.f / (1.f + __expf(-r));
        stf(dst + o, r);
      }
    }
  }
}

// dx[c] (+)= sum_o dz[o] * W[c][o]
template <typename T, int COUT>
__global__ void __launch_bounds__(256)
head_dgrad_kernel(TView dy, const T* __restrict__ wgt, TView dx, int accumulate, long long npix) {
  pdl_sync();
  const int Cin = dx.c, chunks = Cin / 8;
  const int j = threadIdx.x % chunks, ppb = 256 / chunks;
  float w[8][COUT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) w[i][o] = ldf(wgt + (j * 8 + i) * COUT + o);
  const T* dp = reinterpret_cast<const T*>(dy.data);
  T* xp = reinterpret_cast<T*>(dx.data);
  for (long long p = (long long)blockIdx.x * ppb + threadIdx.x / chunks; p < npix; p += (long long)gridDim.x * ppb) {
    const T* src = dp + pix_offset_flat(dy, p);
    float d[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) d[o] = ldf(src + o);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = 0.f;
#pragma unroll
      for (int o = 0; o < COUT; ++o) a += d[o] * w[i][o];
      v[i] = a;
    }
    T* dst = xp + pix_offset_flat(dx, p) + j * 8;
    if (accumulate) {
      float e[8];
      Vec8<T>::load(dst, e);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += e[i];
    }
    Vec8<T>::store(dst, v);
  }
}

// dW[c][o] = sum_pixels x[c] * dz[o]; chunks = Cin/8 threads per pixel (power of two <= 32)
template <typename T, int COUT>
__global__ void __launch_bounds__(256)
head_wgrad_kernel(TView x, TView dy, float* __restrict__ dw, long long npix) {
  pdl_sync();
  extern __shared__ float s_acc[];   // [Cin][COUT]
  const int Cin = x.c, chunks = Cin / 8;
  for (int i = threadIdx.x; i < Cin * COUT; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int j = threadIdx.x % chunks, ppb = 256 / chunks;
  const T* xp = reinterpret_cast<const T*>(x.data);
  const T* dp = reinterpret_cast<const T*>(dy.data);
  float acc[8][COUT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[i][o] = 0.f;
  for (long long p = (long long)blockIdx.x * ppb + threadIdx.x / chunks; p < npix; p += (long long)gridDim.x * ppb) {
    float v[8], d[COUT];
    Vec8<T>::load(xp + pix_offset_flat(x, p) + j * 8, v);
    const T* src = dp + pix_offset_flat(dy, p);
#pragma unroll
    for (int o = 0; o < COUT; ++o) d[o] = ldf(src + o);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[i][o] += v[i] * d[o];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) atomicAdd(&s_acc[(j * 8 + i) * COUT + o], acc[i][o]);
  __syncthreads();
  for (int i = threadIdx.x; i < Cin * COUT; i += 256) atomicAdd(dw + i, s_acc[i]);
}


// ---------------------------------------------------------------------------------------------
// Lean bf16 head kernels (Cin = 64, COUT <= 4, evenly spaced pixels): 32-bit pixel indices, U pixels in
// flight per thread (these kernels were latency bound with one 16-byte load per thread), packed bf16
// unpacking.  8 threads per pixel, one 16-byte chunk each.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(q[i] << 16); v[2 * i + 1] = __uint_as_float(q[i] & 0xffff0000u); }
}

template <int COUT, int U>
__global__ void __launch_bounds__(256)
head_fprop_lean_kernel(const __nv_bfloat16* __restrict__ x, long long x_sw, const __nv_bfloat16* __restrict__ wgt,
                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, long long y_sw, int act, int npix) {
  pdl_sync();
  const int j = threadIdx.x & 7;
  float w[8][COUT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) w[i][o] = __bfloat162float(wgt[(j * 8 + i) * COUT + o]);
  const int step = gridDim.x * 32;
  for (int base = blockIdx.x * 32; base < npix; base += U * step) {      // warp-uniform trip count
    int p[U];
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      p[u] = base + u * step + (threadIdx.x >> 3);
      raw[u] = *reinterpret_cast<const uint4*>(x + (long long)min(p[u], npix - 1) * x_sw + j * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[8];
      unpack8(raw[u], v);
      float acc[COUT];
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a += v[i] * w[i][o];
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        acc[o] = a;
      }
      if (j == 0 && p[u] < npix) {
        __nv_bfloat16* dst = y + (long long)p[u] * y_sw;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          float r = acc[o] + (bias ? bias[o] : 0.f);
          if (act == B200_ACT_RELU) r = fmaxf(r, 0.f);
          else if (act == B200_ACT_SIGMOID) r = 1.f / (1.f + __expf(-r));
          dst[o] = __float2bfloat16_rn(r);
        }
      }
    }
  }
}

// dW[c][o] = sum_p x[p][c] * d[p][o] (+ optional db[o] = sum_p d[p][o])
template <int COUT, int U>
__global__ void __launch_bounds__(256)
head_wgrad_lean_kernel(const __nv_bfloat16* __restrict__ x, long long x_sw, const __nv_bfloat16* __restrict__ d,
                       long long d_sw, float* __restrict__ dw, float* __restrict__ dbias, int npix) {
  pdl_sync();
  __shared__ float s_red[256 * 8];
  const int j = threadIdx.x & 7;
  float acc[8][COUT], bacc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    bacc[o] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][o] = 0.f;
  }
  const int step = gridDim.x * 32;
  for (int base = blockIdx.x * 32; base < npix; base += U * step) {
    uint4 raw[U];
    float dv[U][COUT];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = base + u * step + (threadIdx.x >> 3);
      const bool ok = p < npix;
      raw[u] = ok ? *reinterpret_cast<const uint4*>(x + (long long)p * x_sw + j * 8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int o = 0; o < COUT; ++o) dv[u][o] = ok ? __bfloat162float(d[(long long)p * d_sw + o]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[8];
      unpack8(raw[u], v);
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        bacc[o] += dv[u][o];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][o] += v[i] * dv[u][o];
      }
    }
  }
  // block reduction over the 32 pixel groups, one output column at a time (no shared float atomics)
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[threadIdx.x * 8 + i] = acc[i][o];
    __syncthreads();
    if (threadIdx.x < 64) {
      float a = 0.f;
      for (int g = 0; g < 32; ++g) a += s_red[g * 64 + threadIdx.x];
      atomicAdd(dw + threadIdx.x * COUT + o, a);
    }
    __syncthreads();
  }
  if (dbias) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float a = j == 0 ? bacc[o] : 0.f;
      a = warp_sum(a);
      if ((threadIdx.x & 31) == 0) atomicAdd(dbias + o, a);
    }
  }
}

// column sums of a dense [npix][3] bf16 tensor (the bias gradient of the RGB head): 48 contiguous bytes
// (8 pixels) per thread so that the channel of every element is a compile-time constant
__global__ void __launch_bounds__(256)
bias_sum_c3_kernel(const __nv_bfloat16* __restrict__ d, long long n_elem, float* __restrict__ dbias) {
  pdl_sync();
  float a[3] = {0.f, 0.f, 0.f};
  const long long groups = n_elem / 24;
  for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < groups; g += (long long)gridDim.x * 256) {
    const uint4* src = reinterpret_cast<const uint4*>(d + g * 24);
    const uint4 r0 = src[0], r1 = src[1], r2 = src[2];
    float v[24];
    float t[8];
    unpack8(r0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = t[i];
    unpack8(r1, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 + i] = t[i];
    unpack8(r2, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[16 + i] = t[i];
#pragma unroll
    for (int i = 0; i < 24; ++i) a[i % 3] += v[i];
  }
  if (blockIdx.x == 0) {   // tail elements (fewer than 24)
    for (long long e = groups * 24 + threadIdx.x; e < n_elem; e += 256) a[e % 3] += __bfloat162float(d[e]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = warp_sum(a[c]);
    if ((threadIdx.x & 31) == 0) atomicAdd(dbias + c, s);
  }
}

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }


// ---------------------------------------------------------------------------
// im2col of a narrow input (Cin*9 <= 64, i.e. the RGB stem): xcol[n,h,w, tap*Cin + c] = x[n, h+kh-1, w+kw-1, c]
// (zero outside the image, zero for channels >= 9*Cin), 64 bf16 channels per pixel.  The 3x3 stem
// convolution then IS a 1x1 convolution of xcol with the HWIO kernel read as a [9*Cin -> 64][Cout]
// matrix, which runs on the tcgen05 kernels (fprop with the fused LayerNorm epilogue, wgrad) instead
// of the FMA-bound SIMT stem kernels.  8 threads per pixel, one 16-byte chunk each: a warp writes
// 512 contiguous bytes.
// ---------------------------------------------------------------------------
// block = 32 pixels of one image row x 8 chunks.  The three input rows the segment needs are staged in
// shared memory as fp32 (each row of the patch is 3*Cin CONTIGUOUS input values, so channel k of a pixel is
// s_rows[k / (3*Cin)][pixel*Cin + k % (3*Cin)]); the thread then only does shared loads and one 16-byte store.
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
im2col3x3_kernel(TView x, __nv_bfloat16* __restrict__ xcol, long long col_sw, int H, int W, int cin_rt) {
  pdl_sync();
  constexpr int SEG = 32, MAXC = 7;
  __shared__ float s_rows[3][(SEG + 2) * MAXC];
  const int cin = CIN > 0 ? CIN : cin_rt;
  const int rowlen = (SEG + 2) * cin;
  const int w0 = blockIdx.x * SEG, h = blockIdx.y, n = blockIdx.z;
  const T* img = reinterpret_cast<const T*>(x.data) + (long long)n * x.sn;
  for (int i = threadIdx.x; i < 3 * rowlen; i += 256) {
    const int r = i / rowlen, e = i - r * rowlen;
    const int px = e / cin, c = e - px * cin;
    const int ih = h + r - 1, iw = w0 + px - 1;
    s_rows[r][e] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? ldf(img + (long long)ih * x.sh + (long long)iw * x.sw + c) : 0.f;
  }
  __syncthreads();
  const int pl = threadIdx.x >> 3, q = threadIdx.x & 7;
  const int w = w0 + pl;
  if (w >= W) return;
  const int k3 = 3 * cin, kmax = 9 * cin;
  uint32_t out[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = q * 8 + i * 2 + e;
      const int r = k / k3, m = k - r * k3;
      v[e] = k < kmax ? s_rows[r][pl * cin + m] : 0.f;
    }
    __nv_bfloat162 b = __floats2bfloat162_rn(v[0], v[1]);
    out[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  const long long p = ((long long)n * H + h) * W + w;
  *reinterpret_cast<uint4*>(xcol + p * col_sw + q * 8) = make_uint4(out[0], out[1], out[2], out[3]);
}

}  // namespace

// ---- dispatch helpers ---------------------------------------------------------------------------
bool stem_supported(const b200_tensor* x, const b200_tensor* y, int ks) {
  return ks == 3 && x->c <= 3 && y->c % 8 == 0 && vec_aligned(y, 8) && x->dtype == y->dtype;   // y doubles as dy for wgrad
}

int stem_fprop(const b200_tensor* x, const void* wgt, const float* bias, const b200_tensor* y, int act, cudaStream_t st) {
  const int tiles_w = (y->w + ST_T - 1) / ST_T, tiles_h = (y->h + ST_T - 1) / ST_T;
  dim3 grid(tiles_w * tiles_h * y->n, (y->c + 63) / 64);
  TView xv = view_of(x), yv = view_of(y);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    launch_pdl(stem_fprop_kernel<T>, grid, 256, 0, st, xv, reinterpret_cast<const T*>(wgt), bias, yv, act, tiles_w, tiles_h);
  });
  return check_launch("stem_fprop_kernel");
}

int stem_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw, cudaStream_t st) {
  const int tiles_w = (dy->w + SW_TW - 1) / SW_TW, tiles_h = (dy->h + SW_TH - 1) / SW_TH;
  const int total = tiles_w * tiles_h * dy->n;
  int gx = 4 * sm_count();
  if (gx > total) gx = total;
  cudaMemsetAsync(dw, 0, sizeof(float) * 9 * x->c * dy->c, st);
  dim3 grid(gx, (dy->c + 63) / 64);
  TView xv = view_of(x), dv = view_of(dy);
  B200_DISPATCH_DTYPE(x->dtype, T, { launch_pdl(stem_wgrad_kernel<T>, grid, 256, 0, st, xv, dv, dw, tiles_w, tiles_h, total); });
  return check_launch("stem_wgrad_kernel");
}

// xcol: dense or evenly spaced [N,H,W,64] bf16
int im2col3x3(const b200_tensor* x, const b200_tensor* xcol, cudaStream_t st) {
  B200_REQUIRE(x->c * 9 <= 64 && xcol->c == 64 && xcol->dtype == B200_BF16 && x->n == xcol->n && x->h == xcol->h &&
                   x->w == xcol->w,
               B200_ERR_BAD_ARG, "im2col3x3: need Cin*9 <= 64 and a [N,H,W,64] bf16 destination of the same extent");
  TView cv = view_of(xcol);
  B200_REQUIRE(cv.lin && (reinterpret_cast<uintptr_t>(xcol->data) % 16 == 0) && (xcol->stride_w * 2) % 16 == 0,
               B200_ERR_UNSUPPORTED, "im2col3x3: destination pixels must be evenly spaced and 16-byte aligned");
  B200_REQUIRE(x->h <= 65535 && x->n <= 65535, B200_ERR_UNSUPPORTED, "im2col3x3: H and N must be <= 65535");
  TView xv = view_of(x);
  dim3 grid((unsigned)((x->w + 31) / 32), (unsigned)x->h, (unsigned)x->n);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(xcol->data);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    if (x->c == 3) launch_pdl(im2col3x3_kernel<T, 3>, grid, 256, 0, st, xv, dst, xcol->stride_w, x->h, x->w, 3);
    else launch_pdl(im2col3x3_kernel<T, 0>, grid, 256, 0, st, xv, dst, xcol->stride_w, x->h, x->w, x->c);
  });
  return check_launch("im2col3x3_kernel");
}

bool head_supported(const b200_tensor* x, const b200_tensor* y, int ks) {
  return ks == 1 && (y->c == 1 || y->c == 3) && x->c % 8 == 0 && pow2(x->c / 8) && x->c / 8 <= 32 && vec_aligned(x, 8) &&
         x->dtype == y->dtype;
}

#define B200_HEAD_DISPATCH(cout, ...)            \
  do {                                           \
    if ((cout) == 1) { constexpr int COUT = 1; __VA_ARGS__ } \
    else { constexpr int COUT = 3; __VA_ARGS__ }  \
  } while (0)

static int head_grid(long long npix, int per_block) {
  long long b = (npix + per_block - 1) / per_block;
  long long cap = 8LL * sm_count();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static bool head_lean_ok(const b200_tensor* x, const b200_tensor* y) {
  TView xv = view_of(x), yv = view_of(y);
  return x->dtype == B200_BF16 && x->c == 64 && y->c == 3 && xv.lin && yv.lin && (long long)x->n * x->h * x->w < (1LL << 30);
}

// column sums of a dense bf16 [.., 3] tensor into dbias (+=)
int bias_sum_c3(const b200_tensor* dy, float* dbias, cudaStream_t st) {
  const long long n_elem = (long long)dy->n * dy->h * dy->w * 3;
  long long blocks = (n_elem / 24 + 255) / 256;
  if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
  if (blocks < 1) blocks = 1;
  launch_pdl(bias_sum_c3_kernel, (int)blocks, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dy->data), n_elem, dbias);
  return check_launch("bias_sum_c3_kernel");
}

int head_fprop(const b200_tensor* x, const void* wgt, const float* bias, const b200_tensor* y, int act, cudaStream_t st) {
  if (head_lean_ok(x, y)) {
    const int npix = x->n * x->h * x->w;
    long long blocks = (npix + 4 * 32 - 1) / (4 * 32);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    launch_pdl(head_fprop_lean_kernel<3, 4>, (int)blocks, 256, 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(x->data), x->stride_w, reinterpret_cast<const __nv_bfloat16*>(wgt), bias,
        reinterpret_cast<__nv_bfloat16*>(y->data), y->stride_w, act, npix);
    return check_launch("head_fprop_lean_kernel");
  }
  const long long npix = (long long)x->n * x->h * x->w;
  TView xv = view_of(x), yv = view_of(y);
  const size_t smem = sizeof(float) * x->c * y->c;
  B200_DISPATCH_DTYPE(x->dtype, T, {
    B200_HEAD_DISPATCH(y->c, {
      launch_pdl(head_fprop_kernel<T, COUT>, head_grid(npix, 256 / (x->c / 8)), 256, 0, st, xv, reinterpret_cast<const T*>(wgt), bias, yv, act, npix);
    });
  });
  return check_launch("head_fprop_kernel");
}

int head_dgrad(const b200_tensor* dy, const void* wgt, const b200_tensor* dx, int accumulate, cudaStream_t st) {
  const long long npix = (long long)dx->n * dx->h * dx->w;
  TView dv = view_of(dy), xv = view_of(dx);
  const size_t smem = sizeof(float) * dx->c * dy->c;
  B200_DISPATCH_DTYPE(dx->dtype, T, {
    B200_HEAD_DISPATCH(dy->c, {
      launch_pdl(head_dgrad_kernel<T, COUT>, head_grid(npix, 256 / (dx->c / 8)), 256, 0, st, dv, reinterpret_cast<const T*>(wgt), xv, accumulate, npix);
    });
  });
  return check_launch("head_dgrad_kernel");
}

int head_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw, cudaStream_t st) {
  if (head_lean_ok(x, dy)) {
    const int npix_i = x->n * x->h * x->w;
    cudaMemsetAsync(dw, 0, sizeof(float) * 64 * 3, st);
    long long blocks = (npix_i + 4 * 32 - 1) / (4 * 32);
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    launch_pdl(head_wgrad_lean_kernel<3, 4>, (int)blocks, 256, 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(x->data), x->stride_w, reinterpret_cast<const __nv_bfloat16*>(dy->data),
        dy->stride_w, dw, nullptr, npix_i);
    return check_launch("head_wgrad_lean_kernel");
  }
  const long long npix = (long long)x->n * x->h * x->w;
  TView xv = view_of(x), dv = view_of(dy);
  const size_t smem = sizeof(float) * x->c * dy->c;
  cudaMemsetAsync(dw, 0, smem, st);
  const int ppb = 256 / (x->c / 8);
  long long blocks = (npix + ppb - 1) / ppb;
  if (blocks > 6LL * sm_count()) blocks = 6LL * sm_count();
  B200_DISPATCH_DTYPE(x->dtype, T, {
    B200_HEAD_DISPATCH(dy->c, { launch_pdl(head_wgrad_kernel<T, COUT>, (int)blocks, 256, smem, st, xv, dv, dw, npix); });
  });
  return check_launch("head_wgrad_kernel");
}

}  // namespace b200
