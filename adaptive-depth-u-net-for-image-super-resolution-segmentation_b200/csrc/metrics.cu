// Evaluation metrics of the reference's eval loops as fused reductions
// (Super_resolution/code/train_adaptive_unet.py:144-157, 673-721; evaluate_model.py:94-163):
//   luma_pair   : clip(pred) -> BT.601 luma of pred and hr -> border shave -> the two luma planes + per-image SSE
//   ssim_planes : tf.image.ssim (11x11 Gaussian, sigma 1.5, K1 .01, K2 .03, "valid") per (image, channel) plane of a
//                 channel-interleaved tensor: sums of the SSIM map and of its contrast-structure factor (for MS-SSIM)
//   avgpool2    : the 2x2 average pooling between MS-SSIM scales (odd extents padded symmetrically, as TF does)
// All HBM-bound: every plane is read once per kernel (plus the 10-pixel halo, which hits L2).
#include <math.h>

#include "common.cuh"

namespace b200 {

template <typename T>
__global__ void __launch_bounds__(256)
luma_pair_kernel(TView pred, TView hr, int shave, float* __restrict__ pred_y, float* __restrict__ hr_y,
                 float* __restrict__ sse) {
  pdl_sync();
  const int oh = pred.h - 2 * shave, ow = pred.w - 2 * shave;
  const int n = blockIdx.y;
  const int plane = oh * ow;
  const T* pp = static_cast<const T*>(pred.data);
  const float* hp = static_cast<const float*>(hr.data);
  float acc = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int y = i / ow + shave, x = i % ow + shave;
    const T* p = pp + pix_offset(pred, n, y, x);
    const float* h = hp + pix_offset(hr, n, y, x);
    // evaluate_model.py:106-107: prediction clipped to [0,1], HR taken as is; then rgb_to_luma_bt601 (:144-157)
    const float pr = fminf(fmaxf(ldf<T>(p), 0.0f), 1.0f), pg = fminf(fmaxf(ldf<T>(p + 1), 0.0f), 1.0f),
                pb = fminf(fmaxf(ldf<T>(p + 2), 0.0f), 1.0f);
    float ly = (pr * 65.481f + pg * 128.553f + pb * 24.966f + 16.0f) / 255.0f;
    float lh = (h[0] * 65.481f + h[1] * 128.553f + h[2] * 24.966f + 16.0f) / 255.0f;
    ly = fminf(fmaxf(ly, 0.0f), 1.0f);
    lh = fminf(fmaxf(lh, 0.0f), 1.0f);
    pred_y[(long long)n * plane + i] = ly;
    hr_y[(long long)n * plane + i] = lh;
    const float d = lh - ly;
    acc += d * d;
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(sse + n, v);
  }
}

int luma_pair(const b200_tensor* pred, const b200_tensor* hr, int shave, float* pred_y, float* hr_y, float* sse,
              cudaStream_t st) {
  const int oh = pred->h - 2 * shave, ow = pred->w - 2 * shave;
  cudaMemsetAsync(sse, 0, sizeof(float) * pred->n, st);
  const int plane = oh * ow;
  int bx = (plane + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, pred->n);
  const TView pv = view_of(pred), hv = view_of(hr);
  B200_DISPATCH_DTYPE(pred->dtype, T, { launch_pdl(luma_pair_kernel<T>, grid, 256, 0, st, pv, hv, shave, pred_y, hr_y, sse); });
  return check_launch("luma_pair_kernel");
}

// ---- SSIM ---------------------------------------------------------------------------------------
constexpr int kWin = 11, kTile = 32, kIn = kTile + kWin - 1;   // 42x42 inputs per 32x32 outputs
struct Gauss { float g[kWin]; };

__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w, int ch, Gauss gw, float c1, float c2,
            float* __restrict__ out) {
  pdl_sync();
  __shared__ float sa[kIn][kIn + 1], sb[kIn][kIn + 1];
  __shared__ float hz[5][kIn][kTile];      // horizontally filtered a, b, a*a, b*b, a*b
  __shared__ float red[2][8];
  // plane n = (image n / ch, channel n % ch) of a channel-interleaved [images][h][w][ch] tensor (ch = 1: plain planes)
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const long long base = (long long)(n / ch) * h * w * ch + (n % ch);
  const float* pa = a + base;
  const float* pb = b + base;
  for (int i = threadIdx.x; i < kIn * kIn; i += blockDim.x) {
    const int r = i / kIn, c = i % kIn;
    const int y = y0 + r, x = x0 + c;
    const bool in = y < h && x < w;
    sa[r][c] = in ? pa[((long long)y * w + x) * ch] : 0.0f;
    sb[r][c] = in ? pb[((long long)y * w + x) * ch] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kIn * kTile; i += blockDim.x) {
    const int r = i / kTile, c = i % kTile;
    float fa = 0.f, fb = 0.f, faa = 0.f, fbb = 0.f, fab = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float va = sa[r][c + k], vb = sb[r][c + k], g = gw.g[k];
      fa += g * va; fb += g * vb; faa += g * (va * va); fbb += g * (vb * vb); fab += g * (va * vb);
    }
    hz[0][r][c] = fa; hz[1][r][c] = fb; hz[2][r][c] = faa; hz[3][r][c] = fbb; hz[4][r][c] = fab;
  }
  __syncthreads();
  const int vh = h - kWin + 1, vw = w - kWin + 1;     // "valid" window positions
  float s_ssim = 0.f, s_cs = 0.f;
  for (int i = threadIdx.x; i < kTile * kTile; i += blockDim.x) {
    const int r = i / kTile, c = i % kTile;
    if (y0 + r >= vh || x0 + c >= vw) continue;
    float ma = 0.f, mb = 0.f, eaa = 0.f, ebb = 0.f, eab = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float g = gw.g[k];
      ma += g * hz[0][r + k][c]; mb += g * hz[1][r + k][c];
      eaa += g * hz[2][r + k][c]; ebb += g * hz[3][r + k][c]; eab += g * hz[4][r + k][c];
    }
    // tf.image.ssim's _ssim_helper: luminance and contrast-structure factors
    const float num0 = ma * mb * 2.0f, den0 = ma * ma + mb * mb;
    const float lum = (num0 + c1) / (den0 + c1);
    const float num1 = eab * 2.0f, den1 = eaa + ebb;
    const float cs = (num1 - num0 + c2) / (den1 - den0 + c2);
    s_ssim += lum * cs;
    s_cs += cs;
  }
  s_ssim = warp_sum(s_ssim);
  s_cs = warp_sum(s_cs);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_ssim; red[1][threadIdx.x >> 5] = s_cs; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float v0 = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.0f, v1 = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.0f;
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    if (threadIdx.x == 0) { atomicAdd(out + 2 * n, v0); atomicAdd(out + 2 * n + 1, v1); }
  }
}

int ssim_planes(const float* a, const float* b, int n, int h, int w, int ch, float max_val, float* out, cudaStream_t st) {
  Gauss gw;
  double g[kWin], sum = 0.0;
  for (int k = 0; k < kWin; ++k) {
    const double x = k - (kWin - 1) / 2.0;
    g[k] = exp(-(x * x) / (2.0 * 1.5 * 1.5));
    sum += g[k];
  }
  for (int k = 0; k < kWin; ++k) gw.g[k] = (float)(g[k] / sum);
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  cudaMemsetAsync(out, 0, sizeof(float) * 2 * n * ch, st);
  const int vh = h - kWin + 1, vw = w - kWin + 1;
  dim3 grid((vw + kTile - 1) / kTile, (vh + kTile - 1) / kTile, n * ch);
  launch_pdl(ssim_kernel, grid, 256, 0, st, a, b, h, w, ch, gw, c1, c2, out);
  return check_launch("ssim_kernel");
}

// ---- 2x2 average pooling between MS-SSIM scales ---------------------------------------------------
__global__ void __launch_bounds__(256)
avgpool2_planes_kernel(const float* __restrict__ x, int n, int h, int w, int ch, float* __restrict__ y) {
  pdl_sync();
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  const long long total = (long long)n * oh * ow * ch;      // channel-interleaved [n][h][w][ch] -> [n][oh][ow][ch]
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ch);
    long long q = i / ch;
    const int ox = (int)(q % ow); q /= ow;
    const int oy = (int)(q % oh);
    const float* p = x + (q / oh) * (long long)h * w * ch + c;
    const int y0 = 2 * oy, y1 = min(2 * oy + 1, h - 1), x0 = 2 * ox, x1 = min(2 * ox + 1, w - 1);   // symmetric pad
    y[i] = (p[((long long)y0 * w + x0) * ch] + p[((long long)y0 * w + x1) * ch] + p[((long long)y1 * w + x0) * ch] +
            p[((long long)y1 * w + x1) * ch]) * 0.25f;
  }
}

int avgpool2_planes(const float* x, int n, int h, int w, int ch, float* y, cudaStream_t st) {
  const long long total = (long long)n * ((h + 1) / 2) * ((w + 1) / 2) * ch;
  long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)sm_count() * 16 ? want : (long long)sm_count() * 16);
  launch_pdl(avgpool2_planes_kernel, grid, 256, 0, st, x, n, h, w, ch, y);
  return check_launch("avgpool2_planes_kernel");
}

}  // namespace b200
