// Shared host/device helpers for the b200 U-Net kernels.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/b200_unet.h"

namespace b200 {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);   // also counts one kernel launch
void count_launches(int n);            // extra launches of multi-kernel entry points

#define B200_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) return ::b200::fail((code), __VA_ARGS__); \
  } while (0)

// ---- device view of a b200_tensor --------------------------------------------
struct TView {
  void* data;
  int n, h, w, c;
  long long sn, sh, sw;
  int lin;   // pixels are evenly spaced: offset(pixel p) = p * sw (dense tensors and channel slices)
};

inline TView view_of(const b200_tensor* t) {
  TView v;
  v.data = t->data;
  v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c;
  v.sn = t->stride_n; v.sh = t->stride_h; v.sw = t->stride_w;
  v.lin = (v.sh == (long long)v.w * v.sw && v.sn == (long long)v.h * v.sh) ? 1 : 0;
  return v;
}

inline bool same_shape(const b200_tensor* a, const b200_tensor* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}
inline bool valid_tensor(const b200_tensor* t) {
  return t && t->data && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 &&
         (t->dtype == B200_F32 || t->dtype == B200_BF16);
}
inline size_t dtype_size(int dt) { return dt == B200_BF16 ? 2 : 4; }
// channel vectors are 16-byte aligned when base and strides are
inline bool vec_aligned(const b200_tensor* t, int elems) {
  size_t es = dtype_size(t->dtype);
  return (reinterpret_cast<uintptr_t>(t->data) % 16 == 0) && ((t->c * es) % 16 == 0 || true) &&
         ((t->stride_w * es) % 16 == 0) && ((t->stride_h * es) % 16 == 0) && ((t->stride_n * es) % 16 == 0) &&
         (t->c % elems == 0);
}

// ---- element load/store in fp32 ------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// 8 consecutive channels as fp32 (16 B for bf16, 32 B for fp32)
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ long long pix_offset(const TView& t, int n, int h, int w) {
  return (long long)n * t.sn + (long long)h * t.sh + (long long)w * t.sw;
}
// offset of flat pixel index p (n-major, then h, then w); no divisions for evenly spaced pixels
__device__ __forceinline__ long long pix_offset_flat(const TView& t, long long p) {
  if (t.lin) return p * t.sw;
  const int w = (int)(p % t.w);
  const long long q = p / t.w;
  return (q / t.h) * t.sn + (q % t.h) * t.sh + (long long)w * t.sw;
}

inline int sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

// true the first time `slot` (a per-kernel id, < 16) is seen on the CURRENT device: function attributes
// (dynamic shared memory limits) are per device, so a process that drives several GPUs sets them on each
inline bool first_use_on_device(int slot) {
  static unsigned seen[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  const unsigned bit = 1u << slot;
  if (seen[dev] & bit) return false;
  seen[dev] |= bit;
  return true;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialisation attribute: its CTAs may be
// scheduled while the previous kernel of the stream is still draining (its last CTAs, its memory flush), so launch
// latency, block scheduling and -- in the tcgen05 kernels -- barrier init, TMEM allocation and tensor-map prefetch
// overlap the predecessor's tail.  The contract each kernel keeps: NO global-memory access (read or write) before
// pdl_sync().  griddepcontrol.wait returns once the preceding grid has completed and its writes are visible (so also
// everything that grid waited for: the dependency chain is transitive); launch_dependents right after it lets the
// NEXT kernel start its own prologue once all CTAs of this one are resident.  Cross-stream edges (events) and the
// launches of other libraries stay full dependencies.  B200_PDL=0 launches everything the classic way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_trigger(); }

// Measured (C2 / C1, profiles/r02_pdl_sweep.json): early-resident CTAs of the big memory-bound kernels (8 x 148 blocks)
// spin on SMs that the second stream's wgrad kernels would otherwise use -- C2 slows from 4.38 to 4.51-4.55 ms with PDL
// on every launch, while the launch-bound C1 gains 7 % (1.089 -> 1.011 ms).  Restricting the early start to launches of
// fewer than 2 x SM-count blocks (the persistent tcgen05 kernels, the deep levels, the small tensors) wins on both:
// C2 4.38 -> 4.25 ms, C1 1.089 -> 1.004 ms (bounds 148 / 296 / 448 / 600 / 1200: 4.30 / 4.25 / 4.34 / 4.41 / 4.55 ms).
// B200_PDL=0: never; B200_PDL_MAX_BLOCKS overrides the bound.
inline bool pdl_enabled(unsigned long long blocks = 0) {
  static const bool on = getenv("B200_PDL") ? atoi(getenv("B200_PDL")) != 0 : true;
  static const long long bound = getenv("B200_PDL_MAX_BLOCKS") ? atoll(getenv("B200_PDL_MAX_BLOCKS")) : -1;
  const unsigned long long lim = bound >= 0 ? (unsigned long long)bound : 2ull * (unsigned long long)sm_count();
  return on && blocks < lim;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled((unsigned long long)grid.x * grid.y * grid.z) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// dtype dispatch: DISPATCH_DTYPE(dt, T, { body using T })
#define B200_DISPATCH_DTYPE(dt, T, ...)                          \
  do {                                                            \
    if ((dt) == B200_BF16) { using T = __nv_bfloat16; __VA_ARGS__ } \
    else { using T = float; __VA_ARGS__ }                         \
  } while (0)

}  // namespace b200
