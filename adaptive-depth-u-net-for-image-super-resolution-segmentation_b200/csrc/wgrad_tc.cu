// 3x3 convolution weight gradient on tcgen05 tensor cores.
//
//   dW[tap][ci][co] = sum over pixels  x[pixel + tap][ci] * dz[pixel][co]
//
// is a GEMM whose reduction (K) dimension is the pixel index.  Both operands are read
// straight from their NHWC tiles in shared memory as MN-major UMMA operands (channels
// contiguous, one pixel per 128-byte row), so no transposed copy of the activations is ever
// made.  One CTA owns a 64(ci) x 64(co) block of all nine taps and walks a contiguous range of
// 16x8-pixel tiles; per tile it loads ONE halo window of x ({64ch,10,18} TMA box, shared with
// the fprop kernel's geometry) and one dz tile, and issues, per 16-pixel K step, five M=128
// MMAs: the 128 rows are TWO taps x 64 input channels -- the second tap is the same window
// shifted, expressed through the descriptor's leading-byte-offset.  The 9x64x64 fp32 partial
// sums stay in TMEM (5 x 64 columns) until the CTA has consumed all its tiles, then go to a
// workspace that a second kernel reduces deterministically over the CTAs (b200_conv2d_wgrad), or -- the
// training step's default, b200_conv2d_wgrad_atomic -- are added straight into the zeroed gradient buffer
// with `red.global.add.v4.f32` (no slabs, no reduce launch; the summation order then varies run to run).
//
// Replaces the autodiff (filter gradient) of keras Conv2D at
// Super_resolution/code/train_adaptive_unet.py:202,207,259 (reference: /root/reference).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

using namespace ptx;

int make_act_tmap(CUtensorMap* m, const b200_tensor* t, int box_w, int box_h, int box_n);

namespace {

constexpr int TILE_H = 16, TILE_W = 8;
constexpr int WIN_W = TILE_W + 2, WIN_H = TILE_H + 2;
constexpr int WIN_BYTES = WIN_W * WIN_H * 128;   // 23040
constexpr int WIN_STAGE = 23552;
constexpr int DZ_BYTES = TILE_H * TILE_W * 128;  // 16384
constexpr int STAGE = WIN_STAGE + DZ_BYTES;      // 39936
constexpr int NSTAGES = 5;
constexpr int NTHREADS = 192;
constexpr int NPAIRS = 5;

struct WgradParams {
  int N, H, W, Cin, Cout;
  int tiles_h, tiles_w, total_tiles, splits, cblocks, oblocks;
  int nb, ksteps, stage_bytes, zero_smem;   // stacked small images: nb images per tile, ksteps 16-pixel K steps
  int ntaps;   // 9 (3x3 filter) or 1 (1x1 filter: only the centre tap of the window is accumulated)
  float* partial;  // [splits][9][Cin][Cout]
  float* dw_atomic;   // non-NULL: skip the partial slabs and add straight into dw[9][Cin][Cout] with vector atomics
};

__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// first tap of each pair, and the distance (in window pixel rows) to the second tap
__device__ __forceinline__ int pair_first(int pi) { return pi < 4 ? 2 * pi : 7; }
__device__ __forceinline__ int tap_row(int t) { return (t / 3) * WIN_W + (t % 3); }

__global__ void __launch_bounds__(NTHREADS, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dz,
                   const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[NSTAGES], bar_empty[NSTAGES], bar_done;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // work decomposition: blockIdx.x -> (split, ci block, co block)
  int b = blockIdx.x;
  const int ob = b % p.oblocks; b /= p.oblocks;
  const int cb = b % p.cblocks;
  const int s = b / p.cblocks;
  const int t_begin = (int)((long long)p.total_tiles * s / p.splits);
  const int t_end = (int)((long long)p.total_tiles * (s + 1) / p.splits);

  if (p.zero_smem) {
    // stacked tiles leave K rows past the TMA boxes untouched: make them exact zeros once
    uint4* z = reinterpret_cast<uint4*>(smem_raw + (smem0 - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < NSTAGES * STAGE / 16; i += NTHREADS) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGES; ++i) { mbar_init(smem_u32(&bar_full[i]), 1); mbar_init(smem_u32(&bar_empty[i]), 1); }
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_dz); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  pdl_sync();      // (PDL) the prologue above overlapped the previous kernel; global memory only from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int q = t;
        const int tw = q % p.tiles_w; q /= p.tiles_w;
        const int th = q % p.tiles_h;
        const int n = (q / p.tiles_h) * p.nb;
        mbar_wait(smem_u32(&bar_empty[st]), ph ^ 1);
        const uint32_t base = smem0 + st * STAGE;
        mbar_arrive_expect_tx(smem_u32(&bar_full[st]), (uint32_t)p.stage_bytes);
        tma_load_4d(base, &tm_x, smem_u32(&bar_full[st]), cb * 64, tw * TILE_W - 1, th * TILE_H - 1, n);
        tma_load_4d(base + WIN_STAGE, &tm_dz, smem_u32(&bar_full[st]), ob * 64, tw * TILE_W, th * TILE_H, n);
        if (++st == NSTAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // all lanes run the loop (uniform control flow); one elected lane issues MMAs and commits
    {
      const uint32_t idesc = idesc_bf16(128, 64, 1, 1);
      const uint32_t b_hi = (uint32_t)(smem_desc_sw128(0, 0, 1024) >> 32);
      // A descriptor hi word is shared (SBO = window pitch); the lo word carries LBO = distance
      // between the two taps of the pair
      const uint32_t a_hi = (uint32_t)(smem_desc_sw128(0, 0, WIN_W * 128) >> 32);
      int st = 0, ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(smem_u32(&bar_full[st]), ph);
        tc_fence_after();
        const uint32_t win_lo = (smem0 + st * STAGE) >> 4;
        const uint32_t dz_lo = win_lo + (WIN_STAGE >> 4);
        if (elect_one()) {
#pragma unroll 1
          for (int ks = 0; ks < p.ksteps; ++ks) {
            const uint32_t acc = (t > t_begin || ks > 0) ? 1u : 0u;
            const uint32_t a_ks = win_lo + (uint32_t)ks * (2u * WIN_W * 128u >> 4);
            const uint32_t b_lo = dz_lo + (uint32_t)ks * (2048u >> 4);
#pragma unroll
            for (int pi = 0; pi < NPAIRS; ++pi) {
              if (p.ntaps == 1 && pi != 2) continue;   // 1x1 filter: the pair that starts at the centre tap only
              const int t0 = pair_first(pi);
              const uint32_t lbo = (uint32_t)(tap_row(t0 + 1) - tap_row(t0)) * 128u;
              const uint32_t a_off = ((uint32_t)tap_row(t0) * 128u >> 4) | ((lbo >> 4) << 16);
              umma_lohi(tmem_base + pi * 64, a_ks + a_off, a_hi, b_lo, b_hi, idesc, acc);
            }
          }
          umma_commit(smem_u32(&bar_empty[st]));
        }
        __syncwarp();
        if (++st == NSTAGES) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(smem_u32(&bar_done));
      __syncwarp();
    }
  } else {
    const int q = warp % 4;
    const int r = q * 32 + lane;          // TMEM lane = (tap of the pair, ci)
    const int half = r / 64, ci = r % 64;
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    for (int pi = 0; pi < NPAIRS; ++pi) {
      if (p.ntaps == 1 && pi != 2) continue;
      const int tap = pair_first(pi) + half;
      const bool dup = (pi == 4 && half == 0) || (p.ntaps == 1 && half == 1);   // rows 0..63 of the last pair repeat tap 7
      const int tslot = p.ntaps == 1 ? 0 : tap;
      float* dst = p.dw_atomic
                       ? p.dw_atomic + (((long long)tslot) * p.Cin + cb * 64 + ci) * p.Cout + ob * 64
                       : p.partial + (((long long)s * p.ntaps + tslot) * p.Cin + cb * 64 + ci) * p.Cout + ob * 64;
      const bool live = !dup && cb * 64 + ci < p.Cin;          // 32-channel inputs fill half of the 64 rows
      const int ncols = p.Cout - ob * 64;                      // 32-channel gradients fill half of the 64 columns
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pi * 64 + c0), v);
        tmem_ld_wait();
        if (live && c0 < ncols) {
          if (p.dw_atomic) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              red_add_v4(dst + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                         __uint_as_float(v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + c0 + i) =
                  make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                              __uint_as_float(v[i + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// dw[i] = sum over splits of partial[s][i].  A block owns 64 float4 outputs; its 256 threads are
// 4 split-groups x 64 outputs so that the split loop runs 4-wide, combined through shared memory.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, long long count, int splits,
                    int live_mask, long long per_tap) {
  pdl_sync();
  __shared__ float4 s_part[4][64];
  const long long n4 = count / 4;
  const int lane = threadIdx.x % 64, grp = threadIdx.x / 64;
  for (long long base = (long long)blockIdx.x * 64; base < n4; base += (long long)gridDim.x * 64) {
    const long long i = base + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      const int tap = (int)(i * 4 / per_tap);
      const int nsum = ((live_mask >> tap) & 1) ? splits : 0;   // taps that only ever see padding are exactly zero
      for (int s = grp; s < nsum; s += 4) {
        const float4 v = reinterpret_cast<const float4*>(partial + (long long)s * count)[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    s_part[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && i < n4) {
      float4 r = s_part[0][lane];
#pragma unroll
      for (int g = 1; g < 4; ++g) { const float4 v = s_part[g][lane]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
      reinterpret_cast<float4*>(dw)[i] = r;
    }
    __syncthreads();
  }
}

struct WgradGeom {
  b200_tensor x, dy;      // possibly re-viewed tensors
  int box_h, box_n, live_mask;
};

static bool flatten_1x1(const b200_tensor* t, b200_tensor* out) {
  if (t->h != 1 || t->w != 1) return false;
  *out = *t;
  out->n = 1;
  if (t->n % 8 == 0) { out->h = t->n / 8; out->w = 8; out->stride_w = t->stride_n; out->stride_h = 8 * t->stride_n; }
  else { out->h = 1; out->w = t->n; out->stride_w = t->stride_n; out->stride_h = t->stride_n * t->n; }
  out->stride_n = t->stride_n * t->n;
  return true;
}

inline void plan(const b200_tensor* x_in, const b200_tensor* dy_in, WgradParams& p, WgradGeom& g, int ks) {
  g.x = *x_in; g.dy = *dy_in;
  p.ntaps = ks == 1 ? 1 : 9;
  g.live_mask = 0;
  for (int t = 0; t < 9; ++t) {
    const bool dead = (x_in->h == 1 && t / 3 != 1) || (x_in->w == 1 && t % 3 != 1);
    if (!dead) g.live_mask |= 1 << t;
  }
  if (ks == 1) g.live_mask = 1;   // the reduce kernel sees a single tap block
  b200_tensor xf, yf;
  if (flatten_1x1(x_in, &xf) && flatten_1x1(dy_in, &yf)) { g.x = xf; g.dy = yf; }
  const b200_tensor* x = &g.x;
  p.N = x->n; p.H = x->h; p.W = x->w; p.Cin = x->c; p.Cout = g.dy.c;
  p.tiles_h = (p.H + TILE_H - 1) / TILE_H;
  p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
  p.nb = 1; p.ksteps = 8; p.zero_smem = 0;
  g.box_h = WIN_H; g.box_n = 1;
  p.stage_bytes = WIN_BYTES + DZ_BYTES;
  if (p.H + 2 <= 9 && p.N > 1) {   // stack small images: image b owns window rows [b*(H+2), (b+1)*(H+2))
    const int srows = p.H + 2;
    p.nb = WIN_H / srows;
    while ((p.nb - 1) * srows + p.H - 1 > TILE_H - 1) --p.nb;
    if (p.nb > TILE_H / srows) p.nb = TILE_H / srows;   // the dz tile holds 16 pixel rows
    if (p.nb > p.N) p.nb = p.N;
    g.box_h = srows; g.box_n = p.nb;
    p.ksteps = (p.nb * srows + 1) / 2;
    if (p.ksteps > 8) p.ksteps = 8;
    p.zero_smem = 1;
    p.stage_bytes = p.nb * srows * (WIN_W + TILE_W) * 128;
  }
  const int groups = (p.N + p.nb - 1) / p.nb;
  p.total_tiles = groups * p.tiles_h * p.tiles_w;
  p.cblocks = (p.Cin + 63) / 64;
  p.oblocks = (p.Cout + 63) / 64;
  // one wave: splits * (ci blocks * co blocks) CTAs must not exceed the SM count -- rounding UP here put 152 CTAs on
  // 148 SMs for 8 (ci, co) block pairs (256 -> 128 at 32x32), i.e. a second wave of 4 CTAs and twice the time
  const int pairs = p.cblocks * p.oblocks;
  int splits = sm_count() / pairs;
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
}

}  // namespace

// small-spatial path (conv_gemm.cu)
bool wgrad_small_wanted(const b200_tensor* x, const b200_tensor* dy, int ks);
size_t wgrad_small_workspace(const b200_tensor* x, const b200_tensor* dy, int ks);
int wgrad_small_launch(const b200_tensor* x, const b200_tensor* dy, float* out, int ks, int* splits_out, int* live_mask,
                       cudaStream_t st, int atomic = 0);

bool wgrad_tc_supported(const b200_tensor* x, const b200_tensor* dy, int ks) {
  if (ks != 3 && ks != 1) return false;
  if (x->dtype != B200_BF16 || dy->dtype != B200_BF16) return false;
  // 32-channel tensors: the 64-channel TMA boxes zero-fill the missing half, the epilogue skips those rows / columns
  if ((x->c % 64 != 0 && x->c != 32) || (dy->c % 64 != 0 && dy->c != 32)) return false;
  auto ok = [](const b200_tensor* t) {
    return ((uintptr_t)t->data % 16 == 0) && (t->stride_w * 2) % 16 == 0 && (t->stride_h * 2) % 16 == 0 &&
           (t->stride_n * 2) % 16 == 0;
  };
  return ok(x) && ok(dy);
}

size_t wgrad_tc_workspace(const b200_tensor* x, const b200_tensor* dy, int ks) {
  if (wgrad_small_wanted(x, dy, ks)) return wgrad_small_workspace(x, dy, ks);
  WgradParams p;
  WgradGeom g;
  plan(x, dy, p, g, ks);
  return sizeof(float) * (size_t)p.splits * p.ntaps * p.Cin * p.Cout;
}

int wgrad_tc_launch(const b200_tensor* x, const b200_tensor* dy, float* dw, void* ws, size_t ws_bytes,
                    cudaStream_t st, int ks, int atomic) {
  if (atomic && wgrad_small_wanted(x, dy, ks)) {
    int splits_s = 1, mask_s = 0;
    return wgrad_small_launch(x, dy, dw, ks, &splits_s, &mask_s, st, 1);
  }
  if (wgrad_small_wanted(x, dy, ks)) {
    // deep levels (images <= 8x8): one CTA per (tap, 128 ci, 128 co[, split]) over densely packed pixel boxes
    const size_t need_s = wgrad_small_workspace(x, dy, ks);
    B200_REQUIRE(need_s == 0 || (ws && ws_bytes >= need_s), B200_ERR_BAD_ARG,
                 "conv2d_wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need_s);
    const long long count_s = (long long)(ks == 1 ? 1 : 9) * x->c * dy->c;
    const bool direct = need_s == 0;
    if (direct) cudaMemsetAsync(dw, 0, sizeof(float) * count_s, st);   // taps that only see padding stay zero
    int splits_s = 1, mask_s = 0;
    int rcs = wgrad_small_launch(x, dy, direct ? dw : reinterpret_cast<float*>(ws), ks, &splits_s, &mask_s, st);
    if (rcs || direct) return rcs;
    long long blocks_s = (count_s / 4 + 63) / 64;
    if (blocks_s > 8LL * sm_count()) blocks_s = 8LL * sm_count();
    launch_pdl(wgrad_reduce_kernel, (int)blocks_s, 256, 0, st, reinterpret_cast<const float*>(ws), dw, count_s, splits_s, mask_s,
                                                       (long long)x->c * dy->c);
    return check_launch("wgrad_reduce_kernel");
  }
  WgradParams p;
  WgradGeom g;
  plan(x, dy, p, g, ks);
  const size_t need = sizeof(float) * (size_t)p.splits * p.ntaps * p.Cin * p.Cout;
  B200_REQUIRE(atomic || (ws && ws_bytes >= need), B200_ERR_BAD_ARG, "conv2d_wgrad: workspace too small (%zu < %zu bytes)",
               ws_bytes, need);
  B200_REQUIRE((atomic || (uintptr_t)ws % 16 == 0) && (uintptr_t)dw % 16 == 0, B200_ERR_BAD_ARG,
               "conv2d_wgrad: workspace/dw must be 16-byte aligned");
  p.partial = reinterpret_cast<float*>(ws);
  p.dw_atomic = atomic ? dw : nullptr;
  CUtensorMap tm_x, tm_dz;
  int rc = make_act_tmap(&tm_x, &g.x, WIN_W, g.box_h, g.box_n);
  if (rc) return rc;
  rc = make_act_tmap(&tm_dz, &g.dy, TILE_W, g.box_n > 1 || g.box_h != WIN_H ? g.box_h : TILE_H, g.box_n);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)NSTAGES * STAGE;
  if (first_use_on_device(2))
    cudaFuncSetAttribute(wgrad3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = p.splits * p.cblocks * p.oblocks;
  launch_pdl(wgrad3x3_tc_kernel, grid, NTHREADS, smem, st, tm_x, tm_dz, p);
  int rc2 = check_launch("wgrad3x3_tc_kernel");
  if (rc2 || atomic) return rc2;
  const long long count = (long long)p.ntaps * p.Cin * p.Cout;
  long long blocks = (count / 4 + 63) / 64;
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  launch_pdl(wgrad_reduce_kernel, (int)blocks, 256, 0, st, p.partial, dw, count, p.splits, g.live_mask, (long long)p.Cin * p.Cout);  // counted by check_launch
  return check_launch("wgrad_reduce_kernel");
}

}  // namespace b200
