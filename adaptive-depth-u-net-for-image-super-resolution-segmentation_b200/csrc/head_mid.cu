// 1x1 convolutions with a narrow output (4 <= Cout <= 32) from 32 or 64 channels, bf16: the multi-class segmentation
// head `Conv2D(num_classes, 1)` of unet_vinillia.py:89-90 (BASELINE config 4: 32 -> 21 classes at 256x256).
//
// Cout = 21 has no 16-byte-aligned pixel pitch (42 B), so the TMA / tcgen05 path cannot take it, and the work is tiny
// (1.3 KFLOP per pixel against 106 B of HBM traffic): these are HBM-bound SIMT kernels.
//   fprop : one thread per pixel; the pixel's Cin values in registers, the [Cin][Cout] weights as fp32 in shared memory
//           (broadcast float4 reads, Cout padded to a multiple of 8), the warp's 32 x Cout outputs staged in shared
//           memory and written as contiguous 32-bit words (a pixel's 42 bytes are not a vector store)
//   dgrad : one thread per pixel; dy from the warp's contiguous bytes, dx stored as 16-byte vectors (+ accumulate)
//   wgrad : 64-pixel tiles of x and dy converted to fp32 in shared memory; a thread owns one input channel and 8
//           outputs (1 + 2 shared loads per 8 FMAs), block partial sums added to dw with atomics
#include "common.cuh"

namespace b200 {

namespace {

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(q[i] << 16); v[2 * i + 1] = __uint_as_float(q[i] & 0xffff0000u); }
}

// weights [CIN][cout] bf16 -> shared fp32 [CIN][CP] (padding columns zero)
template <int CIN, int CP>
__device__ __forceinline__ void load_weights(const __nv_bfloat16* __restrict__ wgt, int cout, float (*w_s)[CP]) {
  for (int i = threadIdx.x; i < CIN * CP; i += blockDim.x) {
    const int c = i / CP, o = i % CP;
    w_s[c][o] = o < cout ? __bfloat162float(wgt[c * cout + o]) : 0.f;
  }
}

template <int CIN, int CP>
__global__ void __launch_bounds__(256)
head_mid_fprop_kernel(const __nv_bfloat16* __restrict__ x, long long x_sw, const __nv_bfloat16* __restrict__ wgt,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int cout, int act, long long npix) {
  pdl_sync();
  __shared__ __align__(16) float w_s[CIN][CP];
  __shared__ float b_s[CP];
  __shared__ __align__(16) __nv_bfloat16 out_s[8][32 * CP];
  load_weights<CIN, CP>(wgt, cout, w_s);
  for (int o = threadIdx.x; o < CP; o += blockDim.x) b_s[o] = (bias && o < cout) ? bias[o] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * 256; base < npix; base += (long long)gridDim.x * 256) {
    const long long p0 = base + warp * 32;          // first pixel of this warp (warp-uniform)
    if (p0 >= npix) continue;
    const long long p = p0 + lane;
    const bool valid = p < npix;
    float acc[CP];
#pragma unroll
    for (int o = 0; o < CP; ++o) acc[o] = b_s[o];
    const uint4* src = reinterpret_cast<const uint4*>(x + (valid ? p : p0) * x_sw);
#pragma unroll
    for (int ch = 0; ch < CIN / 8; ++ch) {
      float v[8];
      unpack8(src[ch], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4* wr = reinterpret_cast<const float4*>(w_s[ch * 8 + i]);
#pragma unroll
        for (int o4 = 0; o4 < CP / 4; ++o4) {
          const float4 w = wr[o4];
          acc[o4 * 4 + 0] += v[i] * w.x; acc[o4 * 4 + 1] += v[i] * w.y;
          acc[o4 * 4 + 2] += v[i] * w.z; acc[o4 * 4 + 3] += v[i] * w.w;
        }
      }
    }
    if (act == B200_ACT_RELU) {
#pragma unroll
      for (int o = 0; o < CP; ++o) acc[o] = fmaxf(acc[o], 0.f);
    } else if (act == B200_ACT_SIGMOID) {
#pragma unroll
      for (int o = 0; o < CP; ++o) acc[o] = 1.f / (1.f + __expf(-acc[o]));
    }
    // stage the warp's pixels (pixel-major, cout values each) and write them as contiguous 32-bit words
    __nv_bfloat16* st = out_s[warp];
#pragma unroll
    for (int o = 0; o < CP; ++o)
      if (o < cout) st[lane * cout + o] = __float2bfloat16_rn(acc[o]);
    __syncwarp();
    const long long left = npix - p0;
    const int npx = left < 32 ? (int)left : 32;
    const int nwords = npx * cout / 2;               // 32 * cout is even; a ragged tail may leave one odd element
    uint32_t* dst = reinterpret_cast<uint32_t*>(y + p0 * cout);     // p0 % 32 == 0 -> 4-byte aligned
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(st);
    for (int i = lane; i < nwords; i += 32) dst[i] = sw[i];
    if (lane == 0 && ((npx * cout) & 1)) y[p0 * cout + npx * cout - 1] = st[npx * cout - 1];
    __syncwarp();
  }
}

// dx[p][c] (+)= sum_o dy[p][o] * W[c][o]
template <int CIN, int CP>
__global__ void __launch_bounds__(256)
head_mid_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ wgt,
                      __nv_bfloat16* __restrict__ dx, long long dx_sw, int cout, int accumulate, long long npix) {
  pdl_sync();
  __shared__ __align__(16) float w_s[CIN][CP];
  load_weights<CIN, CP>(wgt, cout, w_s);
  __syncthreads();
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < npix; p += (long long)gridDim.x * 256) {
    float d[CP];
    const __nv_bfloat16* src = dy + p * cout;
#pragma unroll
    for (int o = 0; o < CP; ++o) d[o] = o < cout ? __bfloat162float(src[o]) : 0.f;
    uint4* dst = reinterpret_cast<uint4*>(dx + p * dx_sw);
#pragma unroll 1
    for (int ch = 0; ch < CIN / 8; ++ch) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4* wr = reinterpret_cast<const float4*>(w_s[ch * 8 + i]);
        float a = 0.f;
#pragma unroll
        for (int o4 = 0; o4 < CP / 4; ++o4) {
          const float4 w = wr[o4];
          a += d[o4 * 4 + 0] * w.x + d[o4 * 4 + 1] * w.y + d[o4 * 4 + 2] * w.z + d[o4 * 4 + 3] * w.w;
        }
        v[i] = a;
      }
      if (accumulate) {
        float e[8];
        unpack8(dst[ch], e);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += e[i];
      }
      uint4 r;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      dst[ch] = r;
    }
  }
}

// dw[c][o] += sum_p x[p][c] * dy[p][o]   (dw zeroed by the launcher)
constexpr int WG_TP = 64;   // pixels per shared-memory tile
template <int CIN, int CP>
__global__ void __launch_bounds__(256)
head_mid_wgrad_kernel(const __nv_bfloat16* __restrict__ x, long long x_sw, const __nv_bfloat16* __restrict__ dy,
                      float* __restrict__ dw, int cout, long long npix) {
  pdl_sync();
  __shared__ __align__(16) float x_s[WG_TP][CIN];
  __shared__ __align__(16) float d_s[WG_TP][CP];
  constexpr int OG = CP / 8;                       // output groups of 8
  const int c = threadIdx.x % CIN, og = threadIdx.x / CIN;
  const bool owner = og < OG;                      // CIN * OG <= 256 threads own accumulators, all threads load
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int i = threadIdx.x; i < WG_TP * CP; i += 256) (&d_s[0][0])[i] = 0.f;     // padding columns stay zero
  const long long ntiles = (npix + WG_TP - 1) / WG_TP;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long p0 = t * WG_TP;
    __syncthreads();
    for (int i = threadIdx.x; i < WG_TP * (CIN / 8); i += 256) {
      const int px = i / (CIN / 8), ch = i % (CIN / 8);
      float v[8];
      if (p0 + px < npix) {
        unpack8(reinterpret_cast<const uint4*>(x + (p0 + px) * x_sw)[ch], v);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(&x_s[px][ch * 8]);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int i = threadIdx.x; i < WG_TP * cout; i += 256) {
      const int px = i / cout, o = i % cout;
      d_s[px][o] = (p0 + px < npix) ? __bfloat162float(dy[p0 * cout + i]) : 0.f;
    }
    __syncthreads();
    if (owner) {
#pragma unroll 8
      for (int px = 0; px < WG_TP; ++px) {
        const float xv = x_s[px][c];
        const float4 a = *reinterpret_cast<const float4*>(&d_s[px][og * 8]);
        const float4 b = *reinterpret_cast<const float4*>(&d_s[px][og * 8 + 4]);
        acc[0] += xv * a.x; acc[1] += xv * a.y; acc[2] += xv * a.z; acc[3] += xv * a.w;
        acc[4] += xv * b.x; acc[5] += xv * b.y; acc[6] += xv * b.z; acc[7] += xv * b.w;
      }
    }
  }
  if (owner) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int o = og * 8 + k;
      if (o < cout) atomicAdd(dw + c * cout + o, acc[k]);
    }
  }
}

inline int mid_grid(long long items, int per_block, int per_sm) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)per_sm * sm_count();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// CIN in {32, 64} x CP in {8, 16, 24, 32}
#define B200_MID_DISPATCH(cin, cout, ...)                                                      \
  do {                                                                                         \
    const int cp_ = ((cout) + 7) / 8 * 8;                                                      \
    if ((cin) == 32) {                                                                         \
      constexpr int CIN = 32;                                                                  \
      if (cp_ == 8) { constexpr int CP = 8; __VA_ARGS__ } else if (cp_ == 16) { constexpr int CP = 16; __VA_ARGS__ } \
      else if (cp_ == 24) { constexpr int CP = 24; __VA_ARGS__ } else { constexpr int CP = 32; __VA_ARGS__ }         \
    } else {                                                                                   \
      constexpr int CIN = 64;                                                                  \
      if (cp_ == 8) { constexpr int CP = 8; __VA_ARGS__ } else if (cp_ == 16) { constexpr int CP = 16; __VA_ARGS__ } \
      else if (cp_ == 24) { constexpr int CP = 24; __VA_ARGS__ } else { constexpr int CP = 32; __VA_ARGS__ }         \
    }                                                                                          \
  } while (0)

}  // namespace

// x: [.., Cin] (32 or 64 channels, evenly spaced pixels, 16-byte aligned); y: dense [.., Cout], 4 <= Cout <= 32; bf16
bool head_mid_supported(const b200_tensor* x, const b200_tensor* y, int ks) {
  if (ks != 1 || x->dtype != B200_BF16 || y->dtype != B200_BF16) return false;
  if ((x->c != 32 && x->c != 64) || y->c < 4 || y->c > 32) return false;
  const TView xv = view_of(x), yv = view_of(y);
  if (!xv.lin || !yv.lin || y->stride_w != y->c) return false;
  if (reinterpret_cast<uintptr_t>(x->data) % 16 != 0 || (x->stride_w * 2) % 16 != 0) return false;
  return reinterpret_cast<uintptr_t>(y->data) % 4 == 0;
}

int head_mid_fprop(const b200_tensor* x, const void* wgt, const float* bias, const b200_tensor* y, int act, cudaStream_t st) {
  const long long npix = (long long)x->n * x->h * x->w;
  B200_MID_DISPATCH(x->c, y->c, {
    launch_pdl(head_mid_fprop_kernel<CIN, CP>, mid_grid(npix, 256, 8), 256, 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(x->data), x->stride_w, reinterpret_cast<const __nv_bfloat16*>(wgt), bias,
        reinterpret_cast<__nv_bfloat16*>(y->data), y->c, act, npix);
  });
  return check_launch("head_mid_fprop_kernel");
}

// dy: dense [.., Cout]; dx: [.., Cin]
int head_mid_dgrad(const b200_tensor* dy, const void* wgt, const b200_tensor* dx, int accumulate, cudaStream_t st) {
  const long long npix = (long long)dx->n * dx->h * dx->w;
  B200_MID_DISPATCH(dx->c, dy->c, {
    launch_pdl(head_mid_dgrad_kernel<CIN, CP>, mid_grid(npix, 256, 8), 256, 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(dy->data), reinterpret_cast<const __nv_bfloat16*>(wgt),
        reinterpret_cast<__nv_bfloat16*>(dx->data), dx->stride_w, dy->c, accumulate, npix);
  });
  return check_launch("head_mid_dgrad_kernel");
}

int head_mid_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw, cudaStream_t st) {
  const long long npix = (long long)x->n * x->h * x->w;
  cudaMemsetAsync(dw, 0, sizeof(float) * x->c * dy->c, st);
  B200_MID_DISPATCH(x->c, dy->c, {
    launch_pdl(head_mid_wgrad_kernel<CIN, CP>, mid_grid(npix, WG_TP, 4), 256, 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(x->data), x->stride_w, reinterpret_cast<const __nv_bfloat16*>(dy->data), dw,
        dy->c, npix);
  });
  return check_launch("head_mid_wgrad_kernel");
}

}  // namespace b200
