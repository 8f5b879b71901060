// extern "C" boundary of libb200unet.so (see include/b200_unet.h).
#include <string.h>

#include "common.cuh"

namespace b200 {

// ---- error plumbing -------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static long long g_launches = 0;
void count_launches(int n) { g_launches += n; }
int check_launch(const char* what) {
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B200_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return B200_OK;
}

// implemented in the kernel translation units
int conv_simt_fprop(const b200_tensor*, const b200_filter*, const float*, const b200_tensor*, int, int, bool, cudaStream_t);
int conv_simt_wgrad(const b200_tensor*, const b200_tensor*, int, float*, cudaStream_t);
int filter_pack(const void*, void*, int, int, int, int, cudaStream_t);
bool conv_tc_supported(const b200_tensor*, int, int, const b200_tensor*, int);
struct ConvLnArgs {
  const float* gamma; const float* beta; float eps; int relu;
  const b200_tensor* z;
  float* mean; float* rstd;
};
bool conv_tc_ln_supported(int cout);
struct ConvLnBwdArgs {
  const b200_tensor* z;
  const float* mean; const float* rstd; const float* gamma; const float* beta; int relu;
  float* dgamma; float* dbeta; float* dbias;
};
bool conv_tc_dgrad_lnbwd_supported(const b200_tensor* dy, int cout, int cin, const b200_tensor* dx, int ks);
bool conv_gemm_wanted(const b200_tensor*, int, int, int);
size_t conv_gemm_workspace(const b200_tensor*, int, int, int);
int conv_tc_launch(const b200_tensor*, const void*, int, int, int, int, const float*, const b200_tensor*, int, int, cudaStream_t,
                   const ConvLnArgs* ln, int ks, void* ws, size_t ws_bytes, const void* wmat_k = nullptr, int allow_pairs = 1, const struct ConvLnBwdArgs* lnb = nullptr);
int umma_probe(const void*, int, const void*, int, int, int, int, float*, cudaStream_t);
int umma_rate(int, int, int, long long*, int, cudaStream_t);
bool stem_supported(const b200_tensor*, const b200_tensor*, int);
int stem_fprop(const b200_tensor*, const void*, const float*, const b200_tensor*, int, cudaStream_t);
int stem_wgrad(const b200_tensor*, const b200_tensor*, float*, cudaStream_t);
int im2col3x3(const b200_tensor*, const b200_tensor*, cudaStream_t);
bool head_supported(const b200_tensor*, const b200_tensor*, int);
int head_fprop(const b200_tensor*, const void*, const float*, const b200_tensor*, int, cudaStream_t);
int head_dgrad(const b200_tensor*, const void*, const b200_tensor*, int, cudaStream_t);
int head_wgrad(const b200_tensor*, const b200_tensor*, float*, cudaStream_t);
bool wgrad_tc_supported(const b200_tensor*, const b200_tensor*, int);
size_t wgrad_tc_workspace(const b200_tensor*, const b200_tensor*, int);
int wgrad_tc_launch(const b200_tensor*, const b200_tensor*, float*, void*, size_t, cudaStream_t, int, int atomic = 0);
int layernorm_fwd(const b200_tensor*, const float*, const float*, float, int, const b200_tensor*, float*, float*, cudaStream_t);
int layernorm_bwd(const b200_tensor*, const b200_tensor*, const float*, const float*, const float*, const float*, int,
                  const b200_tensor*, float*, float*, float*, cudaStream_t);
int bias_act_bwd(const b200_tensor*, const b200_tensor*, int, const b200_tensor*, float*, cudaStream_t);
int batchnorm_fwd_train(const b200_tensor*, const float*, const float*, float, float, int, const b200_tensor*, float*,
                        float*, float*, float*, double*, cudaStream_t);
int batchnorm_fwd_infer(const b200_tensor*, const float*, const float*, float, int, const float*, const float*,
                        const b200_tensor*, cudaStream_t);
int batchnorm_bwd(const b200_tensor*, const b200_tensor*, const float*, const float*, const float*, const float*, int,
                  const b200_tensor*, float*, float*, float*, double*, cudaStream_t);
int batchnorm_stats(const b200_tensor*, const b200_tensor*, const float*, const float*, const float*, const float*, int,
                    double*, float*, float*, cudaStream_t);
int batchnorm_fwd_apply(const b200_tensor*, const float*, const float*, float, float, int, const b200_tensor*, float*, float*,
                        float*, float*, const double*, double, cudaStream_t);
int batchnorm_bwd_apply(const b200_tensor*, const b200_tensor*, const float*, const float*, const float*, const float*, int,
                        const b200_tensor*, const double*, double, cudaStream_t);
int resize_extent(int, float);
int resample_taps(int, int, int);
int resample_plan(int, int, int, int32_t*, float*, int);
int resample_plan_transpose(int, int, int, const int32_t*, const float*, int32_t*, float*, int);
int resample2d(const b200_tensor*, const b200_tensor*, const int32_t*, const float*, int, const int32_t*, const float*,
               int, int, int, cudaStream_t);
int resample_compact(int, int, int32_t*, float*);
int resample_mode(int, int, const int32_t*);
int maxpool2_fwd(const b200_tensor*, const b200_tensor*, cudaStream_t);
int maxpool2_bwd(const b200_tensor*, const b200_tensor*, const b200_tensor*, const b200_tensor*, int, cudaStream_t);
int clipadd_fwd(const b200_tensor*, const b200_tensor*, const b200_tensor*, cudaStream_t);
int clipadd_bwd(const b200_tensor*, const b200_tensor*, const b200_tensor*, const b200_tensor*, cudaStream_t);
int sr_loss(const b200_tensor*, const b200_tensor*, int, float, float, float*, const b200_tensor*, float*, cudaStream_t);
int bce_dice_loss(const b200_tensor*, const b200_tensor*, float, float, float, float*, const b200_tensor*, float*,
                  cudaStream_t);
int binary_confusion(const b200_tensor*, const b200_tensor*, float, float*, cudaStream_t);
int softmax_fwd(const b200_tensor*, const b200_tensor*, cudaStream_t);
int softmax_ce_loss(const b200_tensor*, const int32_t*, float, float*, const b200_tensor*, float*, cudaStream_t);
int adam_advance(int32_t*, const float*, cudaStream_t);
int adam_step(float*, const float*, float*, float*, size_t, const float*, const int32_t*, void*, const float*, cudaStream_t);
int loss_scale_apply(void*, int, size_t, const float*, cudaStream_t);
int loss_scale_check(const float*, size_t, float*, cudaStream_t);
int loss_scale_update(float*, float, cudaStream_t);
int cast(const void*, int, void*, int, size_t, cudaStream_t);
int copy_tensor(const b200_tensor*, const b200_tensor*, cudaStream_t);
int scale_inplace(float*, size_t, float, cudaStream_t);
size_t convT2_workspace(const b200_tensor*, int, int, int);
int convT2_fprop(const b200_tensor*, const void*, const float*, int, const b200_tensor*, void*, size_t, cudaStream_t);
int convT2_dgrad(const b200_tensor*, const void*, int, const b200_tensor*, void*, size_t, cudaStream_t);
int convT2_wgrad(const b200_tensor*, const b200_tensor*, float*, float*, cudaStream_t);
bool head_mid_supported(const b200_tensor*, const b200_tensor*, int);
int head_mid_fprop(const b200_tensor*, const void*, const float*, const b200_tensor*, int, cudaStream_t);
int head_mid_dgrad(const b200_tensor*, const void*, const b200_tensor*, int, cudaStream_t);
int head_mid_wgrad(const b200_tensor*, const b200_tensor*, float*, cudaStream_t);
int cv_resize_taps(int, int, int);
int cv_resize_plan(int, int, int, int32_t*, float*, int);
int patch_extract(const void*, int, int, int, const int32_t*, const b200_tensor*, cudaStream_t);
int gather2d(const b200_tensor*, const b200_tensor*, const int32_t*, const float*, int, const int32_t*, const float*, int,
             int, cudaStream_t);
int copy_rows(const float*, const int32_t*, float*, const int32_t*, int, long long, cudaStream_t);
int luma_pair(const b200_tensor*, const b200_tensor*, int, float*, float*, float*, cudaStream_t);
int ssim_planes(const float*, const float*, int, int, int, int, float, float*, cudaStream_t);
int avgpool2_planes(const float*, int, int, int, int, float*, cudaStream_t);

}  // namespace b200

using namespace b200;

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define REQ_T(t, name) B200_REQUIRE(valid_tensor(t), B200_ERR_BAD_ARG, "%s: invalid tensor '%s'", __func__, name)

extern "C" {

const char* b200_version(void) { return "b200unet 0.1.0 (sm_100a)"; }
const char* b200_last_error(void) { return g_err; }
long long b200_launch_count(int reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int b200_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(B200_ERR_LAUNCH, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return fail(B200_ERR_LAUNCH, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (sms) *sms = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return B200_OK;
}

#define WS_PTR(ws) ((ws) ? (ws)->ptr : nullptr)
#define WS_BYTES(ws) ((ws) ? (ws)->bytes : (size_t)0)

int b200_conv2d_fprop(const b200_tensor* x, const b200_filter* f, const float* bias, const b200_tensor* y, int act,
                      int algo, const b200_scratch* ws, void* stream) {
  REQ_T(x, "x"); REQ_T(y, "y");
  B200_REQUIRE(f && f->hwio, B200_ERR_BAD_ARG, "conv2d_fprop: filter missing");
  B200_REQUIRE(x->c == f->cin && y->c == f->cout && x->n == y->n && x->h == y->h && x->w == y->w, B200_ERR_BAD_ARG,
               "conv2d_fprop: shapes x[%d,%d,%d,%d] y[%d,%d,%d,%d] filter cin=%d cout=%d disagree", x->n, x->h, x->w,
               x->c, y->n, y->h, y->w, y->c, f->cin, f->cout);
  const bool tc_ok = (f->kh == 3 || f->kh == 1) && f->kw == f->kh && f->dtype == B200_BF16 &&
                     (act == B200_ACT_NONE || act == B200_ACT_RELU) && conv_tc_supported(x, f->cin, f->cout, y, f->kh);
  if (algo == B200_ALGO_TCGEN05 || algo == B200_ALGO_TCGEN05_1CTA || (algo == B200_ALGO_AUTO && tc_ok)) {
    B200_REQUIRE(tc_ok, B200_ERR_UNSUPPORTED, "conv2d_fprop: tcgen05 path does not support this shape/dtype");
    return conv_tc_launch(x, f->hwio, f->cin, f->cout, 0, 1, bias, y, act, 0, ST(stream), nullptr, f->kh, WS_PTR(ws), WS_BYTES(ws),
                          f->ohwi, algo != B200_ALGO_TCGEN05_1CTA);
  }
  if (algo == B200_ALGO_AUTO && f->dtype == x->dtype && f->kh == f->kw) {
    if (stem_supported(x, y, f->kh) && act != B200_ACT_SIGMOID) return stem_fprop(x, f->hwio, bias, y, act, ST(stream));
    if (head_supported(x, y, f->kh)) return head_fprop(x, f->hwio, bias, y, act, ST(stream));
    if (head_mid_supported(x, y, f->kh)) return head_mid_fprop(x, f->hwio, bias, y, act, ST(stream));
  }
  return conv_simt_fprop(x, f, bias, y, act, 0, false, ST(stream));
}

int b200_conv2d_ln_fprop(const b200_tensor* x, const b200_filter* f, const float* bias, const float* gamma,
                         const float* beta, float eps, int relu, const b200_tensor* z, const b200_tensor* y, float* mean,
                         float* rstd, int algo, const b200_scratch* ws, void* stream) {
  REQ_T(x, "x"); REQ_T(y, "y");
  B200_REQUIRE(f && f->hwio && gamma && beta && mean && rstd, B200_ERR_BAD_ARG, "conv2d_ln_fprop: NULL argument");
  B200_REQUIRE(x->c == f->cin && y->c == f->cout && x->n == y->n && x->h == y->h && x->w == y->w, B200_ERR_BAD_ARG,
               "conv2d_ln_fprop: shapes disagree with filter cin=%d cout=%d", f->cin, f->cout);
  const bool have_z = z && z->data;
  if (have_z) B200_REQUIRE(same_shape(z, y) && z->dtype == y->dtype, B200_ERR_BAD_ARG, "conv2d_ln_fprop: z/y mismatch");
  const bool fused = (f->kh == 3 || f->kh == 1) && f->kw == f->kh && f->dtype == B200_BF16 &&
                     conv_tc_supported(x, f->cin, f->cout, y, f->kh) && conv_tc_ln_supported(f->cout) &&
                     algo != B200_ALGO_SIMT && !conv_gemm_wanted(x, f->cin, f->cout, f->kh);
  if (fused) {
    ConvLnArgs ln{gamma, beta, eps, relu, have_z ? z : nullptr, mean, rstd};
    return conv_tc_launch(x, f->hwio, f->cin, f->cout, 0, 1, bias, y, B200_ACT_NONE, 0, ST(stream), &ln, f->kh, nullptr, 0,
                          f->ohwi, algo != B200_ALGO_TCGEN05_1CTA);
  }
  // composition: convolution into z (or into y when the caller keeps no z), then the stand-alone LayerNorm
  const b200_tensor* zz = have_z ? z : y;
  int rc = b200_conv2d_fprop(x, f, bias, zz, B200_ACT_NONE, algo, ws, stream);
  if (rc) return rc;
  return layernorm_fwd(zz, gamma, beta, eps, relu, y, mean, rstd, ST(stream));
}

int b200_conv2d_dgrad(const b200_tensor* dy, const b200_filter* f, const b200_tensor* dx, int accumulate, int algo,
                      const b200_scratch* ws, void* stream) {
  REQ_T(dy, "dy"); REQ_T(dx, "dx");
  B200_REQUIRE(f && f->hwio, B200_ERR_BAD_ARG, "conv2d_dgrad: filter missing");
  B200_REQUIRE(dy->c == f->cout && dx->c == f->cin && dx->n == dy->n && dx->h == dy->h && dx->w == dy->w,
               B200_ERR_BAD_ARG, "conv2d_dgrad: shapes disagree with filter cin=%d cout=%d", f->cin, f->cout);
  // dgrad is a convolution of dy (K = cout) producing cin channels; its B operand [tap][cin][cout]
  // is the HWIO kernel itself, read with the tap order reversed.
  const bool tc_ok = (f->kh == 3 || f->kh == 1) && f->kw == f->kh && f->dtype == B200_BF16 &&
                     conv_tc_supported(dy, f->cout, f->cin, dx, f->kh);
  if (algo == B200_ALGO_TCGEN05 || algo == B200_ALGO_TCGEN05_1CTA || (algo == B200_ALGO_AUTO && tc_ok)) {
    B200_REQUIRE(tc_ok, B200_ERR_UNSUPPORTED, "conv2d_dgrad: tcgen05 path does not support this shape/dtype");
    return conv_tc_launch(dy, f->hwio, f->cout, f->cin, 1, 0, nullptr, dx, B200_ACT_NONE, accumulate, ST(stream), nullptr,
                          f->kh, WS_PTR(ws), WS_BYTES(ws), nullptr, algo != B200_ALGO_TCGEN05_1CTA);
  }
  if (algo == B200_ALGO_AUTO && f->dtype == dx->dtype && f->kh == f->kw && head_supported(dx, dy, f->kh))
    return head_dgrad(dy, f->hwio, dx, accumulate, ST(stream));
  if (algo == B200_ALGO_AUTO && f->dtype == dx->dtype && f->kh == f->kw && head_mid_supported(dx, dy, f->kh))
    return head_mid_dgrad(dy, f->hwio, dx, accumulate, ST(stream));
  return conv_simt_fprop(dy, f, nullptr, dx, B200_ACT_NONE, accumulate, true, ST(stream));
}

size_t b200_conv2d_workspace(const b200_tensor* x, const b200_filter* f, int dgrad) {
  if (!x || !f || f->dtype != B200_BF16 || x->dtype != B200_BF16 || f->kh != f->kw || (f->kh != 3 && f->kh != 1)) return 0;
  const int cin = dgrad ? f->cout : f->cin, cout = dgrad ? f->cin : f->cout;
  if (x->c != cin || !conv_gemm_wanted(x, cin, cout, f->kh)) return 0;
  return conv_gemm_workspace(x, cin, cout, f->kh);
}

size_t b200_conv2d_wgrad_workspace(const b200_tensor* x, const b200_tensor* dy, int kh, int kw, int algo) {
  if (algo == B200_ALGO_SIMT || kh != kw || (kh != 3 && kh != 1)) return 0;
  if (!wgrad_tc_supported(x, dy, kh)) return 0;
  return wgrad_tc_workspace(x, dy, kh);
}

int b200_conv2d_dgrad_ln_bwd_supported(const b200_tensor* dy, const b200_filter* f, const b200_tensor* dz) {
  if (!valid_tensor(dy) || !valid_tensor(dz) || !f || !f->hwio || f->dtype != B200_BF16 || f->kh != 3 || f->kw != 3) return 0;
  if (dy->c != f->cout || dz->c != f->cin || dy->n != dz->n || dy->h != dz->h || dy->w != dz->w) return 0;
  return conv_tc_dgrad_lnbwd_supported(dy, f->cout, f->cin, dz, 3) ? 1 : 0;
}

int b200_conv2d_dgrad_ln_bwd(const b200_tensor* dy, const b200_filter* f, const b200_tensor* z, const float* mean,
                             const float* rstd, const float* gamma, const float* beta, int relu, const b200_tensor* dz,
                             float* dgamma, float* dbeta, float* dbias, void* stream) {
  REQ_T(dy, "dy"); REQ_T(z, "z"); REQ_T(dz, "dz");
  B200_REQUIRE(f && f->hwio && mean && rstd && gamma && beta, B200_ERR_BAD_ARG, "conv2d_dgrad_ln_bwd: NULL argument");
  B200_REQUIRE(b200_conv2d_dgrad_ln_bwd_supported(dy, f, dz), B200_ERR_UNSUPPORTED,
               "conv2d_dgrad_ln_bwd: only 3x3 bf16 filters with 64 input channels on plain 16x8 tiles (see "
               "b200_conv2d_dgrad_ln_bwd_supported); run b200_conv2d_dgrad + b200_layernorm_bwd instead");
  ConvLnBwdArgs a{z, mean, rstd, gamma, beta, relu, dgamma, dbeta, dbias};
  return conv_tc_launch(dy, f->hwio, f->cout, f->cin, 1, 0, nullptr, dz, B200_ACT_NONE, 0, ST(stream), nullptr, 3, nullptr, 0,
                        nullptr, 1, &a);
}

int b200_conv2d_wgrad(const b200_tensor* x, const b200_tensor* dy, int kh, int kw, float* dw, void* ws, size_t ws_bytes,
                      int algo, void* stream) {
  REQ_T(x, "x"); REQ_T(dy, "dy");
  B200_REQUIRE(dw, B200_ERR_BAD_ARG, "conv2d_wgrad: dw is NULL");
  B200_REQUIRE(kh == kw && x->n == dy->n && x->h == dy->h && x->w == dy->w, B200_ERR_BAD_ARG,
               "conv2d_wgrad: shapes disagree");
  const bool tc_ok = wgrad_tc_supported(x, dy, kh);
  if (algo == B200_ALGO_TCGEN05 || (algo == B200_ALGO_AUTO && tc_ok)) {
    B200_REQUIRE(tc_ok, B200_ERR_UNSUPPORTED, "conv2d_wgrad: tcgen05 path does not support this shape/dtype");
    return wgrad_tc_launch(x, dy, dw, ws, ws_bytes, ST(stream), kh);
  }
  if (algo == B200_ALGO_AUTO) {
    if (stem_supported(x, dy, kh)) return stem_wgrad(x, dy, dw, ST(stream));
    if (head_supported(x, dy, kh)) return head_wgrad(x, dy, dw, ST(stream));
    if (head_mid_supported(x, dy, kh)) return head_mid_wgrad(x, dy, dw, ST(stream));
  }
  return conv_simt_wgrad(x, dy, kh, dw, ST(stream));
}

int b200_im2col3x3(const b200_tensor* x, const b200_tensor* xcol, void* stream) {
  REQ_T(x, "x"); REQ_T(xcol, "xcol");
  return im2col3x3(x, xcol, ST(stream));
}

int b200_conv2d_wgrad_atomic(const b200_tensor* x, const b200_tensor* dy, int kh, int kw, float* dw, void* stream) {
  REQ_T(x, "x"); REQ_T(dy, "dy");
  B200_REQUIRE(dw, B200_ERR_BAD_ARG, "conv2d_wgrad_atomic: dw is NULL");
  B200_REQUIRE(kh == kw && x->n == dy->n && x->h == dy->h && x->w == dy->w, B200_ERR_BAD_ARG,
               "conv2d_wgrad_atomic: shapes disagree");
  B200_REQUIRE(wgrad_tc_supported(x, dy, kh), B200_ERR_UNSUPPORTED,
               "conv2d_wgrad_atomic: only the tcgen05 shapes (bf16, channels multiples of 64, 3x3 or 1x1)");
  return wgrad_tc_launch(x, dy, dw, nullptr, 0, ST(stream), kh, 1);
}

int b200_filter_pack(const void* hwio, void* ohwi, int kh, int kw, int cin, int cout, int dtype, void* stream) {
  B200_REQUIRE(hwio && ohwi && kh > 0 && kw > 0 && cin > 0 && cout > 0, B200_ERR_BAD_ARG, "filter_pack: bad argument");
  return filter_pack(hwio, ohwi, kh * kw, cin, cout, dtype, ST(stream));
}

size_t b200_convT2x2_workspace(const b200_tensor* x, int cin, int cout, int dgrad) {
  if (!valid_tensor(x)) return 0;
  return convT2_workspace(x, cin, cout, dgrad);
}
int b200_convT2x2_fprop(const b200_tensor* x, const void* k, const float* bias, int cout, const b200_tensor* y,
                        const b200_scratch* ws, void* s) {
  REQ_T(x, "x"); REQ_T(y, "y");
  return convT2_fprop(x, k, bias, cout, y, WS_PTR(ws), WS_BYTES(ws), ST(s));
}
int b200_convT2x2_dgrad(const b200_tensor* dy, const void* k, int cout, const b200_tensor* dx, const b200_scratch* ws,
                        void* s) {
  REQ_T(dy, "dy"); REQ_T(dx, "dx");
  return convT2_dgrad(dy, k, cout, dx, WS_PTR(ws), WS_BYTES(ws), ST(s));
}
int b200_convT2x2_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dk, float* db, void* s) {
  REQ_T(x, "x"); REQ_T(dy, "dy");
  return convT2_wgrad(x, dy, dk, db, ST(s));
}

int b200_bias_act_bwd(const b200_tensor* dy, const b200_tensor* y, int act, const b200_tensor* dz, float* dbias, void* s) {
  REQ_T(dy, "dy"); REQ_T(y, "y"); REQ_T(dz, "dz");
  return bias_act_bwd(dy, y, act, dz, dbias, ST(s));
}

int b200_layernorm_fwd(const b200_tensor* z, const float* g, const float* b, float eps, int relu, const b200_tensor* y,
                       float* mean, float* rstd, void* s) {
  REQ_T(z, "z"); REQ_T(y, "y");
  B200_REQUIRE(g && b && mean && rstd, B200_ERR_BAD_ARG, "layernorm_fwd: NULL parameter");
  return layernorm_fwd(z, g, b, eps, relu, y, mean, rstd, ST(s));
}
int b200_layernorm_bwd(const b200_tensor* dy, const b200_tensor* z, const float* mean, const float* rstd, const float* g,
                       const float* b, int relu, const b200_tensor* dz, float* dg, float* db, float* dbias, void* s) {
  REQ_T(dy, "dy"); REQ_T(z, "z"); REQ_T(dz, "dz");
  B200_REQUIRE(g && b && mean && rstd, B200_ERR_BAD_ARG, "layernorm_bwd: NULL parameter");
  return layernorm_bwd(dy, z, mean, rstd, g, b, relu, dz, dg, db, dbias, ST(s));
}

int b200_batchnorm_fwd_train(const b200_tensor* z, const float* g, const float* b, float eps, float mom, int relu,
                             const b200_tensor* y, float* sm, float* sr, float* mm, float* mv, double* ws, void* s) {
  REQ_T(z, "z"); REQ_T(y, "y");
  B200_REQUIRE(g && b && sm && sr && ws, B200_ERR_BAD_ARG, "batchnorm_fwd_train: NULL parameter");
  return batchnorm_fwd_train(z, g, b, eps, mom, relu, y, sm, sr, mm, mv, ws, ST(s));
}
int b200_batchnorm_fwd_infer(const b200_tensor* z, const float* g, const float* b, float eps, int relu, const float* mm,
                             const float* mv, const b200_tensor* y, void* s) {
  REQ_T(z, "z"); REQ_T(y, "y");
  B200_REQUIRE(g && b && mm && mv, B200_ERR_BAD_ARG, "batchnorm_fwd_infer: NULL parameter");
  return batchnorm_fwd_infer(z, g, b, eps, relu, mm, mv, y, ST(s));
}
int b200_batchnorm_bwd(const b200_tensor* dy, const b200_tensor* z, const float* sm, const float* sr, const float* g,
                       const float* b, int relu, const b200_tensor* dz, float* dg, float* db, float* dbias, double* ws,
                       void* s) {
  REQ_T(dy, "dy"); REQ_T(z, "z"); REQ_T(dz, "dz");
  B200_REQUIRE(g && b && sm && sr && ws, B200_ERR_BAD_ARG, "batchnorm_bwd: NULL parameter");
  return batchnorm_bwd(dy, z, sm, sr, g, b, relu, dz, dg, db, dbias, ws, ST(s));
}

int b200_batchnorm_stats(const b200_tensor* z, const b200_tensor* dy, const float* sm, const float* sr, const float* g,
                         const float* b, int relu, double* ws, float* dg, float* db, void* s) {
  REQ_T(z, "z");
  B200_REQUIRE(ws, B200_ERR_BAD_ARG, "batchnorm_stats: NULL workspace");
  const bool bwd = dy && dy->data;
  if (bwd) REQ_T(dy, "dy");
  return batchnorm_stats(z, bwd ? dy : nullptr, sm, sr, g, b, relu, ws, dg, db, ST(s));
}
int b200_batchnorm_fwd_apply(const b200_tensor* z, const float* g, const float* b, float eps, float mom, int relu,
                             const b200_tensor* y, float* sm, float* sr, float* mm, float* mv, const double* ws,
                             double count, void* s) {
  REQ_T(z, "z"); REQ_T(y, "y");
  B200_REQUIRE(g && b && sm && sr && ws, B200_ERR_BAD_ARG, "batchnorm_fwd_apply: NULL parameter");
  return batchnorm_fwd_apply(z, g, b, eps, mom, relu, y, sm, sr, mm, mv, ws, count, ST(s));
}
int b200_batchnorm_bwd_apply(const b200_tensor* dy, const b200_tensor* z, const float* sm, const float* sr, const float* g,
                             const float* b, int relu, const b200_tensor* dz, const double* ws, double count, void* s) {
  REQ_T(dy, "dy"); REQ_T(z, "z"); REQ_T(dz, "dz");
  B200_REQUIRE(g && b && sm && sr && ws, B200_ERR_BAD_ARG, "batchnorm_bwd_apply: NULL parameter");
  return batchnorm_bwd_apply(dy, z, sm, sr, g, b, relu, dz, ws, count, ST(s));
}

int b200_resize_extent(int extent, float scale) { return resize_extent(extent, scale); }
int b200_resample_taps(int in_size, int out_size, int aa) { return resample_taps(in_size, out_size, aa); }
int b200_resample_plan(int in_size, int out_size, int aa, int32_t* starts, float* weights, int taps) {
  B200_REQUIRE(starts && weights, B200_ERR_BAD_ARG, "resample_plan: NULL table");
  return resample_plan(in_size, out_size, aa, starts, weights, taps);
}
int b200_resample_plan_transpose(int in_size, int out_size, int taps, const int32_t* starts, const float* weights,
                                 int32_t* ts, float* tw, int tt) {
  B200_REQUIRE(starts && weights, B200_ERR_BAD_ARG, "resample_plan_transpose: NULL table");
  return resample_plan_transpose(in_size, out_size, taps, starts, weights, ts, tw, tt);
}
int b200_resample2d(const b200_tensor* x, const b200_tensor* y, const int32_t* hs, const float* hw, int ht,
                    const int32_t* ws, const float* ww, int wt, int accumulate, void* s) {
  REQ_T(x, "x"); REQ_T(y, "y");
  B200_REQUIRE(hs && hw && ws && ww && ht > 0 && wt > 0, B200_ERR_BAD_ARG, "resample2d: NULL table");
  return resample2d(x, y, hs, hw, ht, ws, ww, wt, accumulate, 0, ST(s));
}
int b200_resample_compact(int n_out, int taps, int32_t* starts, float* weights) {
  B200_REQUIRE(starts && weights && n_out > 0 && taps > 0, B200_ERR_BAD_ARG, "resample_compact: bad argument");
  return resample_compact(n_out, taps, starts, weights);
}
int b200_resample_mode(int n_out, int taps, const int32_t* starts) {
  B200_REQUIRE(starts && n_out > 0 && taps > 0, B200_ERR_BAD_ARG, "resample_mode: bad argument");
  return resample_mode(n_out, taps, starts);
}
int b200_resample2d_ex(const b200_tensor* x, const b200_tensor* y, const int32_t* hs, const float* hw, int ht,
                       const int32_t* ws, const float* ww, int wt, int accumulate, int h_mode, void* s) {
  REQ_T(x, "x"); REQ_T(y, "y");
  B200_REQUIRE(hs && hw && ws && ww && ht > 0 && wt > 0, B200_ERR_BAD_ARG, "resample2d_ex: NULL table");
  B200_REQUIRE(h_mode >= 0 && h_mode <= 3, B200_ERR_BAD_ARG, "resample2d_ex: h_mode %d out of range", h_mode);
  return resample2d(x, y, hs, hw, ht, ws, ww, wt, accumulate, h_mode, ST(s));
}

int b200_maxpool2_fwd(const b200_tensor* x, const b200_tensor* y, void* s) {
  REQ_T(x, "x"); REQ_T(y, "y");
  return maxpool2_fwd(x, y, ST(s));
}
int b200_maxpool2_bwd(const b200_tensor* x, const b200_tensor* y, const b200_tensor* dy, const b200_tensor* dx, int acc,
                      void* s) {
  REQ_T(x, "x"); REQ_T(y, "y"); REQ_T(dy, "dy"); REQ_T(dx, "dx");
  return maxpool2_bwd(x, y, dy, dx, acc, ST(s));
}

int b200_clipadd_fwd(const b200_tensor* inp, const b200_tensor* res, const b200_tensor* y, void* s) {
  REQ_T(inp, "inp"); REQ_T(res, "res"); REQ_T(y, "y");
  return clipadd_fwd(inp, res, y, ST(s));
}
int b200_clipadd_bwd(const b200_tensor* inp, const b200_tensor* res, const b200_tensor* dy, const b200_tensor* dres,
                     void* s) {
  REQ_T(inp, "inp"); REQ_T(res, "res"); REQ_T(dy, "dy"); REQ_T(dres, "dres");
  return clipadd_bwd(inp, res, dy, dres, ST(s));
}

int b200_sr_loss(const b200_tensor* pred, const b200_tensor* target, int kind, float eps, float gs, float* out,
                 const b200_tensor* dpred, float* ws, void* s) {
  REQ_T(pred, "pred"); REQ_T(target, "target");
  B200_REQUIRE(out && ws, B200_ERR_BAD_ARG, "sr_loss: NULL out/ws");
  return sr_loss(pred, target, kind, eps, gs, out, dpred, ws, ST(s));
}
int b200_bce_dice_loss(const b200_tensor* pred, const b200_tensor* target, float bw, float dw, float gs, float* out,
                       const b200_tensor* dpred, float* ws, void* s) {
  REQ_T(pred, "pred"); REQ_T(target, "target");
  B200_REQUIRE(out && ws, B200_ERR_BAD_ARG, "bce_dice_loss: NULL out/ws");
  return bce_dice_loss(pred, target, bw, dw, gs, out, dpred, ws, ST(s));
}
int b200_binary_confusion(const b200_tensor* pred, const b200_tensor* target, float threshold, float* counts, void* s) {
  REQ_T(pred, "pred"); REQ_T(target, "target");
  B200_REQUIRE(counts, B200_ERR_BAD_ARG, "binary_confusion: NULL counts");
  return binary_confusion(pred, target, threshold, counts, ST(s));
}
int b200_softmax_fwd(const b200_tensor* z, const b200_tensor* p, void* s) {
  REQ_T(z, "z"); REQ_T(p, "p");
  return softmax_fwd(z, p, ST(s));
}
int b200_softmax_ce_loss(const b200_tensor* prob, const int32_t* labels, float gs, float* out, const b200_tensor* dl,
                         float* ws, void* s) {
  REQ_T(prob, "prob");
  B200_REQUIRE(labels && out && ws, B200_ERR_BAD_ARG, "softmax_ce_loss: NULL argument");
  return softmax_ce_loss(prob, labels, gs, out, dl, ws, ST(s));
}

int b200_adam_advance(int32_t* step, const float* loss_scale, void* s) {
  B200_REQUIRE(step, B200_ERR_BAD_ARG, "adam_advance: NULL step");
  return adam_advance(step, loss_scale, ST(s));
}
int b200_adam_step(float* p, const float* g, float* m, float* v, size_t count, const float* hyper, const int32_t* step,
                   void* shadow, const float* loss_scale, void* s) {
  B200_REQUIRE(p && g && m && v && hyper && step, B200_ERR_BAD_ARG, "adam_step: NULL argument");
  if (count == 0) return B200_OK;
  return adam_step(p, g, m, v, count, hyper, step, shadow, loss_scale, ST(s));
}
int b200_loss_scale_apply(void* data, int dtype, size_t count, const float* loss_scale, void* s) {
  B200_REQUIRE(data && loss_scale && (dtype == B200_F32 || dtype == B200_BF16), B200_ERR_BAD_ARG, "loss_scale_apply: bad argument");
  if (count == 0) return B200_OK;
  return loss_scale_apply(data, dtype, count, loss_scale, ST(s));
}
int b200_loss_scale_check(const float* g, size_t count, float* loss_scale, void* s) {
  B200_REQUIRE(g && loss_scale && (uintptr_t)g % 16 == 0, B200_ERR_BAD_ARG, "loss_scale_check: bad argument");
  if (count == 0) return B200_OK;
  return loss_scale_check(g, count, loss_scale, ST(s));
}
int b200_loss_scale_update(float* loss_scale, float growth_interval, void* s) {
  B200_REQUIRE(loss_scale && growth_interval >= 1.f, B200_ERR_BAD_ARG, "loss_scale_update: bad argument");
  return loss_scale_update(loss_scale, growth_interval, ST(s));
}

int b200_cast(const void* src, int sdt, void* dst, int ddt, size_t count, void* s) {
  B200_REQUIRE(src && dst, B200_ERR_BAD_ARG, "cast: NULL argument");
  if (count == 0) return B200_OK;
  return cast(src, sdt, dst, ddt, count, ST(s));
}
int b200_copy_tensor(const b200_tensor* src, const b200_tensor* dst, void* s) {
  REQ_T(src, "src"); REQ_T(dst, "dst");
  return copy_tensor(src, dst, ST(s));
}
int b200_scale_inplace(float* p, size_t count, float sc, void* s) {
  B200_REQUIRE(p, B200_ERR_BAD_ARG, "scale_inplace: NULL argument");
  if (count == 0) return B200_OK;
  return scale_inplace(p, count, sc, ST(s));
}

// ---- patch pipeline ---------------------------------------------------------------------------
int b200_cv_resize_taps(int in_size, int out_size, int interp) {
  B200_REQUIRE(in_size > 0 && out_size > 0, B200_ERR_BAD_ARG, "cv_resize_taps: sizes must be positive");
  B200_REQUIRE(interp == B200_CV_INTER_AREA || interp == B200_CV_INTER_CUBIC, B200_ERR_BAD_ARG,
               "cv_resize_taps: unknown interpolation %d", interp);
  B200_REQUIRE(interp == B200_CV_INTER_CUBIC || out_size <= in_size, B200_ERR_UNSUPPORTED,
               "cv_resize_taps: INTER_AREA tables are for shrinking (%d -> %d)", in_size, out_size);
  return cv_resize_taps(in_size, out_size, interp);
}
int b200_cv_resize_plan(int in_size, int out_size, int interp, int32_t* idx, float* weights, int taps) {
  B200_REQUIRE(idx && weights, B200_ERR_BAD_ARG, "cv_resize_plan: NULL table");
  const int need = b200_cv_resize_taps(in_size, out_size, interp);
  if (need < 0) return need;
  B200_REQUIRE(taps >= need, B200_ERR_BAD_ARG, "cv_resize_plan: %d taps given, %d needed", taps, need);
  return cv_resize_plan(in_size, out_size, interp, idx, weights, taps);
}
int b200_patch_extract(const void* image, int image_dtype, int img_h, int img_w, const int32_t* origins,
                       const b200_tensor* hr, void* s) {
  REQ_T(hr, "hr");
  B200_REQUIRE(image && origins, B200_ERR_BAD_ARG, "patch_extract: NULL argument");
  B200_REQUIRE(image_dtype == B200_U8 || image_dtype == B200_F32, B200_ERR_BAD_ARG, "patch_extract: image dtype %d",
               image_dtype);
  B200_REQUIRE(hr->dtype == B200_F32 && hr->c == 3 && hr->stride_w == 3 && hr->h == hr->w, B200_ERR_BAD_ARG,
               "patch_extract: hr must be fp32 [n,P,P,3] with dense rows");
  B200_REQUIRE(img_h >= hr->h && img_w >= hr->w, B200_ERR_BAD_ARG, "patch_extract: patch_size exceeds image dimensions");
  return patch_extract(image, image_dtype, img_h, img_w, origins, hr, ST(s));
}
int b200_gather2d(const b200_tensor* x, const b200_tensor* y, const int32_t* h_idx, const float* h_w, int h_taps,
                  const int32_t* w_idx, const float* w_w, int w_taps, int clip01, void* s) {
  REQ_T(x, "x"); REQ_T(y, "y");
  B200_REQUIRE(h_idx && h_w && w_idx && w_w && h_taps > 0 && w_taps > 0, B200_ERR_BAD_ARG, "gather2d: bad tables");
  B200_REQUIRE(x->dtype == B200_F32 && y->dtype == B200_F32 && x->c == y->c && x->n == y->n, B200_ERR_BAD_ARG,
               "gather2d: fp32 tensors with equal batch and channels expected");
  return gather2d(x, y, h_idx, h_w, h_taps, w_idx, w_w, w_taps, clip01, ST(s));
}
int b200_copy_rows(const float* src, const int32_t* src_rows, float* dst, const int32_t* dst_rows, int n_rows,
                   long long row_elems, void* s) {
  B200_REQUIRE(src && dst && row_elems > 0 && n_rows >= 0 && n_rows <= 65535, B200_ERR_BAD_ARG, "copy_rows: bad argument");
  if (n_rows == 0) return B200_OK;
  return copy_rows(src, src_rows, dst, dst_rows, n_rows, row_elems, ST(s));
}

// ---- evaluation metrics -------------------------------------------------------------------------
int b200_luma_pair(const b200_tensor* pred, const b200_tensor* hr, int shave, float* pred_y, float* hr_y, float* sse,
                   void* s) {
  REQ_T(pred, "pred_rgb"); REQ_T(hr, "hr_rgb");
  B200_REQUIRE(pred_y && hr_y && sse, B200_ERR_BAD_ARG, "luma_pair: NULL output");
  B200_REQUIRE(same_shape(pred, hr) && pred->c == 3 && hr->dtype == B200_F32, B200_ERR_BAD_ARG,
               "luma_pair: prediction and fp32 target must both be [n,h,w,3]");
  B200_REQUIRE(shave >= 0 && 2 * shave < pred->h && 2 * shave < pred->w && pred->n <= 65535, B200_ERR_BAD_ARG,
               "luma_pair: shave %d removes the full %dx%d frame", shave, pred->h, pred->w);
  return luma_pair(pred, hr, shave, pred_y, hr_y, sse, ST(s));
}
int b200_ssim_planes(const float* a, const float* b, int n, int h, int w, int channels, float max_val, float* out,
                     void* s) {
  B200_REQUIRE(a && b && out && n > 0 && channels > 0 && (long long)n * channels <= 65535, B200_ERR_BAD_ARG,
               "ssim_planes: bad argument");
  B200_REQUIRE(h >= 11 && w >= 11, B200_ERR_BAD_ARG, "ssim_planes: %dx%d is smaller than the 11x11 window", h, w);
  return ssim_planes(a, b, n, h, w, channels, max_val, out, ST(s));
}
int b200_avgpool2_planes(const float* x, int n, int h, int w, int channels, float* y, void* s) {
  B200_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && channels > 0, B200_ERR_BAD_ARG, "avgpool2_planes: bad argument");
  return avgpool2_planes(x, n, h, w, channels, y, ST(s));
}

int b200_debug_umma_probe(const void* a, int a_rows, const void* b, int start_bytes, int sbo_bytes, int lbo_bytes,
                          int mn_major, float* out, void* s) {
  B200_REQUIRE(a && b && out, B200_ERR_BAD_ARG, "umma_probe: NULL argument");
  return umma_probe(a, a_rows, b, start_bytes, sbo_bytes, lbo_bytes, mn_major, out, ST(s));
}

int b200_debug_umma_rate(int n, int iters, int a_stride_bytes, long long* cycles, int grid, void* s) {
  B200_REQUIRE(cycles && (n == 64 || n == 128 || n == 256) && grid > 0, B200_ERR_BAD_ARG, "umma_rate: bad argument");
  return umma_rate(n, iters, a_stride_bytes, cycles, grid, ST(s));
}

}  // extern "C"
