// SIMT direct convolution (stride 1, "same"), fp32 accumulate, fp32 or bf16 storage.
//
// These kernels cover what the tcgen05 implicit-GEMM path does not: the Cin=3
// stem, the Cout=3/1/21 1x1 heads, fp32 storage, and arbitrary strides.  They are
// also the on-device cross-check for the tensor-core kernels in tests.
//
// Replaces keras Conv2D at Super_resolution/code/train_adaptive_unet.py:202,207,259,267
// (reference: /root/reference) for those shapes.
#include "common.cuh"

namespace b200 {

namespace {

constexpr int TP = 8;        // output tile is TP x TP pixels
constexpr int TCO = 64;      // output channels per block
constexpr int CK = 8;        // input-channel chunk staged in smem
constexpr int NTHREADS = 256;

// y[n,h,w,o] = act(bias[o] + sum_{kh,kw,c} x[n,h+kh-p,w+kw-p,c] * W(kh,kw,c,o))
// DGRAD: W(kh,kw,c,o) = hwio[KS-1-kh][KS-1-kw][o][c]  (filter cin = our Cout, filter cout = our Cin)
template <typename T, int KS, bool DGRAD>
__global__ void __launch_bounds__(NTHREADS)
conv_direct_kernel(TView x, const T* __restrict__ wgt, const float* __restrict__ bias, TView y, int act,
                   int accumulate, int tiles_w, int tiles_h) {
  pdl_sync();
  constexpr int HALO = KS - 1;
  constexpr int PAD = KS / 2;
  constexpr int IW = TP + HALO;
  __shared__ float s_in[IW * IW][CK];
  __shared__ __align__(16) float s_w[KS * KS][CK][TCO];

  const int tid = threadIdx.x;
  const int cg = tid % 16;      // 4 couts each
  const int pg = tid / 16;      // 4 pixels each: row pg/2, cols (pg%2)*4..
  int bt = blockIdx.x;
  const int tw = bt % tiles_w; bt /= tiles_w;
  const int th = bt % tiles_h; bt /= tiles_h;
  const int n = bt;
  const int h0 = th * TP, w0 = tw * TP;
  const int co0 = blockIdx.y * TCO;
  const int Cin = x.c, Cout = y.c;
  const T* xp = reinterpret_cast<const T*>(x.data);

  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[p][j] = 0.f;

  const int py = pg / 2, px0 = (pg % 2) * 4;

  for (int c0 = 0; c0 < Cin; c0 += CK) {
    // stage input patch
    for (int i = tid; i < IW * IW * CK; i += NTHREADS) {
      int ck = i % CK, pix = i / CK;
      int ih = h0 + pix / IW - PAD, iw = w0 + pix % IW - PAD;
      int c = c0 + ck;
      float v = 0.f;
      if (ih >= 0 && ih < x.h && iw >= 0 && iw < x.w && c < Cin) v = ldf(xp + pix_offset(x, n, ih, iw) + c);
      s_in[pix][ck] = v;
    }
    // stage weights
    for (int i = tid; i < KS * KS * CK * TCO; i += NTHREADS) {
      int o = i % TCO, ck = (i / TCO) % CK, tap = i / (TCO * CK);
      int c = c0 + ck, oo = co0 + o;
      float v = 0.f;
      if (c < Cin && oo < Cout) {
        if (!DGRAD) v = ldf(wgt + ((long long)tap * Cin + c) * Cout + oo);
        else v = ldf(wgt + ((long long)(KS * KS - 1 - tap) * Cout + oo) * Cin + c);
      }
      s_w[tap][ck][o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kh = 0; kh < KS; ++kh)
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
#pragma unroll
        for (int ck = 0; ck < CK; ++ck) {
          float4 w4 = *reinterpret_cast<const float4*>(&s_w[kh * KS + kw][ck][cg * 4]);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float xv = s_in[(py + kh) * IW + px0 + p + kw][ck];
            acc[p][0] += xv * w4.x; acc[p][1] += xv * w4.y; acc[p][2] += xv * w4.z; acc[p][3] += xv * w4.w;
          }
        }
      }
    __syncthreads();
  }

  T* yp = reinterpret_cast<T*>(y.data);
  const int oh = h0 + py;
  if (oh >= y.h) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int ow = w0 + px0 + p;
    if (ow >= y.w) continue;
    T* dst = yp + pix_offset(y, n, oh, ow);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = co0 + cg * 4 + j;
      if (o >= Cout) continue;
      float v = acc[p][j] + (bias ? bias[o] : 0.f);
      if (act == B200_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == B200_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
      if (accumulate) v += ldf(dst + o);
      stf(dst + o, v);
    }
  }
}

// dW[tap][ci][co] += sum over this block's pixel tiles of x[.., ci] * dy[.., co]
template <typename T, int KS, int CI_T>
__global__ void __launch_bounds__(NTHREADS)
wgrad_direct_kernel(TView x, TView dy, float* __restrict__ dw, int tiles_w, int tiles_h, int total_tiles) {
  pdl_sync();
  constexpr int HALO = KS - 1;
  constexpr int PAD = KS / 2;
  constexpr int IW = TP + HALO;
  constexpr int CO_THREADS = NTHREADS / CI_T;
  constexpr int CO_PER = TCO / CO_THREADS;
  __shared__ float s_x[IW * IW][CI_T];
  __shared__ __align__(16) float s_dy[TP * TP][TCO];

  const int tid = threadIdx.x;
  const int ci_l = tid % CI_T;
  const int cot = tid / CI_T;
  const int ci0 = blockIdx.x * CI_T;
  const int co0 = blockIdx.y * TCO;
  const int Cin = x.c, Cout = dy.c;
  const T* xp = reinterpret_cast<const T*>(x.data);
  const T* dyp = reinterpret_cast<const T*>(dy.data);

  float acc[KS * KS][CO_PER];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t)
#pragma unroll
    for (int j = 0; j < CO_PER; ++j) acc[t][j] = 0.f;

  for (int tile = blockIdx.z; tile < total_tiles; tile += gridDim.z) {
    int bt = tile;
    const int tw = bt % tiles_w; bt /= tiles_w;
    const int th = bt % tiles_h; bt /= tiles_h;
    const int n = bt;
    const int h0 = th * TP, w0 = tw * TP;
    for (int i = tid; i < IW * IW * CI_T; i += NTHREADS) {
      int ck = i % CI_T, pix = i / CI_T;
      int ih = h0 + pix / IW - PAD, iw = w0 + pix % IW - PAD;
      int c = ci0 + ck;
      float v = 0.f;
      if (ih >= 0 && ih < x.h && iw >= 0 && iw < x.w && c < Cin) v = ldf(xp + pix_offset(x, n, ih, iw) + c);
      s_x[pix][ck] = v;
    }
    for (int i = tid; i < TP * TP * TCO; i += NTHREADS) {
      int o = i % TCO, pix = i / TCO;
      int oh = h0 + pix / TP, ow = w0 + pix % TP;
      float v = 0.f;
      if (oh < dy.h && ow < dy.w && co0 + o < Cout) v = ldf(dyp + pix_offset(dy, n, oh, ow) + co0 + o);
      s_dy[pix][o] = v;
    }
    __syncthreads();
    for (int p = 0; p < TP * TP; ++p) {
      const int py = p / TP, px = p % TP;
      float d[CO_PER];
#pragma unroll
      for (int j = 0; j < CO_PER; ++j) d[j] = s_dy[p][cot * CO_PER + j];
#pragma unroll
      for (int kh = 0; kh < KS; ++kh)
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          float xv = s_x[(py + kh) * IW + px + kw][ci_l];
#pragma unroll
          for (int j = 0; j < CO_PER; ++j) acc[kh * KS + kw][j] += xv * d[j];
        }
    }
    __syncthreads();
  }
  const int ci = ci0 + ci_l;
  if (ci < Cin) {
#pragma unroll
    for (int t = 0; t < KS * KS; ++t)
#pragma unroll
      for (int j = 0; j < CO_PER; ++j) {
        int o = co0 + cot * CO_PER + j;
        if (o < Cout) atomicAdd(dw + ((long long)t * Cin + ci) * Cout + o, acc[t][j]);
      }
  }
}

// [tap][cin][cout] -> [tap][cout][cin], 32x32 tiles through shared memory (coalesced both ways)
template <typename T>
__global__ void __launch_bounds__(256)
filter_pack_kernel(const T* __restrict__ hwio, T* __restrict__ ohwi, int cin, int cout) {
  pdl_sync();
  __shared__ T tile[32][33];
  const int t = blockIdx.z;
  const int c0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
  const T* src = hwio + (long long)t * cin * cout;
  T* dst = ohwi + (long long)t * cin * cout;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, o = o0 + tx;
    if (c < cin && o < cout) tile[r][tx] = src[(long long)c * cout + o];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, c = c0 + tx;
    if (o < cout && c < cin) dst[(long long)o * cin + c] = tile[tx][r];
  }
}

}  // namespace

int conv_simt_fprop(const b200_tensor* x, const b200_filter* f, const float* bias, const b200_tensor* y, int act,
                    int accumulate, bool dgrad, cudaStream_t st) {
  const int ks = f->kh;
  B200_REQUIRE(f->kh == f->kw && (ks == 1 || ks == 3), B200_ERR_UNSUPPORTED, "conv_simt: kernel %dx%d unsupported",
               f->kh, f->kw);
  B200_REQUIRE(x->dtype == y->dtype && x->dtype == f->dtype, B200_ERR_BAD_ARG, "conv_simt: dtype mismatch");
  const int tiles_w = (y->w + TP - 1) / TP, tiles_h = (y->h + TP - 1) / TP;
  dim3 grid(tiles_w * tiles_h * y->n, (y->c + TCO - 1) / TCO);
  TView xv = view_of(x), yv = view_of(y);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    const T* w = reinterpret_cast<const T*>(f->hwio);
    if (ks == 3) {
      if (dgrad) launch_pdl(conv_direct_kernel<T, 3, true>, grid, NTHREADS, 0, st, xv, w, bias, yv, act, accumulate, tiles_w, tiles_h);
      else launch_pdl(conv_direct_kernel<T, 3, false>, grid, NTHREADS, 0, st, xv, w, bias, yv, act, accumulate, tiles_w, tiles_h);
    } else {
      if (dgrad) launch_pdl(conv_direct_kernel<T, 1, true>, grid, NTHREADS, 0, st, xv, w, bias, yv, act, accumulate, tiles_w, tiles_h);
      else launch_pdl(conv_direct_kernel<T, 1, false>, grid, NTHREADS, 0, st, xv, w, bias, yv, act, accumulate, tiles_w, tiles_h);
    }
  });
  return check_launch("conv_direct_kernel");
}

int conv_simt_wgrad(const b200_tensor* x, const b200_tensor* dy, int ks, float* dw, cudaStream_t st) {
  B200_REQUIRE(ks == 1 || ks == 3, B200_ERR_UNSUPPORTED, "wgrad_simt: kernel %d unsupported", ks);
  B200_REQUIRE(x->dtype == dy->dtype, B200_ERR_BAD_ARG, "wgrad_simt: dtype mismatch");
  const int tiles_w = (dy->w + TP - 1) / TP, tiles_h = (dy->h + TP - 1) / TP;
  const int total = tiles_w * tiles_h * dy->n;
  const bool narrow = x->c <= 4;
  const int ci_t = narrow ? 4 : 16;
  const int gx = (x->c + ci_t - 1) / ci_t, gy = (dy->c + TCO - 1) / TCO;
  int gz = (4 * sm_count() + gx * gy - 1) / (gx * gy);
  if (gz > total) gz = total;
  if (gz < 1) gz = 1;
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)ks * ks * x->c * dy->c, st);
  dim3 grid(gx, gy, gz);
  TView xv = view_of(x), dv = view_of(dy);
  B200_DISPATCH_DTYPE(x->dtype, T, {
    if (ks == 3) {
      if (narrow) launch_pdl(wgrad_direct_kernel<T, 3, 4>, grid, NTHREADS, 0, st, xv, dv, dw, tiles_w, tiles_h, total);
      else launch_pdl(wgrad_direct_kernel<T, 3, 16>, grid, NTHREADS, 0, st, xv, dv, dw, tiles_w, tiles_h, total);
    } else {
      if (narrow) launch_pdl(wgrad_direct_kernel<T, 1, 4>, grid, NTHREADS, 0, st, xv, dv, dw, tiles_w, tiles_h, total);
      else launch_pdl(wgrad_direct_kernel<T, 1, 16>, grid, NTHREADS, 0, st, xv, dv, dw, tiles_w, tiles_h, total);
    }
  });
  return check_launch("wgrad_direct_kernel");
}

int filter_pack(const void* hwio, void* ohwi, int taps, int cin, int cout, int dtype, cudaStream_t st) {
  dim3 grid((cin + 31) / 32, (cout + 31) / 32, taps);
  B200_DISPATCH_DTYPE(dtype, T, {
    launch_pdl(filter_pack_kernel<T>, grid, 256, 0, st, reinterpret_cast<const T*>(hwio), reinterpret_cast<T*>(ohwi), cin, cout);
  });
  return check_launch("filter_pack_kernel");
}

}  // namespace b200
