// Small-spatial convolutions (images of at most 8x8 pixels: the deep levels of the U-Net) as
// split-K implicit GEMMs on tcgen05.
//
// At 2x2 or 1x1 pixels a 1024->512 layer is a GEMM with M = batch*H*W = 256 rows against 9.4 MB of
// weights: it is bound by streaming the weights, not by the tensor pipe, and the halo-window kernel of
// conv_tc.cu fits it badly (one 16x8 tile holds 16 valid pixels, and every tile re-reads all weights).
// Here
//   * M is packed densely: one TMA box {64ch, bw, bh, bn} with bw*bh*bn = 128 covers bn whole images,
//     and the box for filter tap (kh,kw) is simply loaded at coordinates (w0+kw-1, h0+kh-1): the
//     TMA's out-of-bounds zero fill is the "same" padding of every image in the box;
//   * K = (64-channel block, tap) units are split across CTAs so that all 148 SMs stream disjoint
//     slabs of the weights; partial sums go to an fp32 workspace [split][row][Cout];
//   * a finalize kernel adds the splits, the bias and the activation and writes bf16 NHWC
//     (channel-strided, so concat-in-place still works).
// The same scheme gives the filter gradient of these layers (K = the 128 pixels of a box, M x N =
// 128 ci x 128 co per tap, both operands MN-major straight from the NHWC boxes).
//
// Replaces the cuDNN kernels behind keras Conv2D at the deep levels of build_super_resolution_unet
// (Super_resolution/code/train_adaptive_unet.py:245-262, reference: /root/reference).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

using namespace ptx;

int make_act_tmap(CUtensorMap* m, const b200_tensor* t, int box_w, int box_h, int box_n);
int make_mat_tmap(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows);

namespace {

constexpr int NTHREADS = 192;
constexpr int A_BYTES = 128 * 128;   // 128 pixels x 64 channels
constexpr int MAX_STAGES = 6;

struct Geom {
  int bw, bh, bn, tiles_w, tiles_h, groups, m_tiles;
};

inline int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

inline Geom geom_of(const b200_tensor* t) {
  Geom g;
  g.bw = pow2ceil(t->w) < 8 ? pow2ceil(t->w) : 8;
  g.bh = pow2ceil(t->h) < 16 ? pow2ceil(t->h) : 16;
  while (g.bw * g.bh > 128) g.bh >>= 1;
  g.bn = 128 / (g.bw * g.bh);
  g.tiles_w = (t->w + g.bw - 1) / g.bw;
  g.tiles_h = (t->h + g.bh - 1) / g.bh;
  g.groups = (t->n + g.bn - 1) / g.bn;
  g.m_tiles = g.groups * g.tiles_h * g.tiles_w;
  return g;
}

struct GemmParams {
  int KB, BN, n_tiles, m_tiles, splits, units, nlive;
  int tiles_w, tiles_h, bw, bh, bn;
  int b_mn, tap_rev, Kc, Cout, ntaps, tap0;
  int live_taps[9];
  float* ws;   // [splits][m_tiles * 128][Cout]
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

// one CTA = one (split, m tile, n tile): accumulate its K units, dump the fp32 tile to the workspace
__global__ void __launch_bounds__(NTHREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wt_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = A_BYTES + wt_bytes;
  const int nstages = p.BN == 64 ? 6 : (p.BN == 128 ? 6 : 4);

  int item = blockIdx.x;
  const int j = item % p.n_tiles; item /= p.n_tiles;
  const int mt = item % p.m_tiles;
  const int split = item / p.m_tiles;
  const int u0 = (int)((long long)p.units * split / p.splits);
  const int u1 = (int)((long long)p.units * (split + 1) / p.splits);
  int t = mt;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h;
  const int n0 = (t / p.tiles_h) * p.bn;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(smem_u32(&bar_full[i]), 1); mbar_init(smem_u32(&bar_empty[i]), 1); }
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_b); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)p.BN < 32u ? 32u : (uint32_t)p.BN); tmem_relinquish(); }
  pdl_sync();      // (PDL) global memory only from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int u = u0; u < u1; ++u) {
        const int kb = u / p.nlive, tap = p.live_taps[u % p.nlive];
        mbar_wait(smem_u32(&bar_empty[st]), ph ^ 1);
        const uint32_t a_dst = smem0 + st * stage_bytes, b_dst = a_dst + A_BYTES;
        const uint32_t bar = smem_u32(&bar_full[st]);
        mbar_arrive_expect_tx(bar, stage_bytes);
        tma_load_4d(a_dst, &tm_x, bar, kb * 64, tw * p.bw + tap % 3 - 1, th * p.bh + tap / 3 - 1, n0);
        if (p.b_mn) {
          for (int a = 0; a < p.BN / 64; ++a)
            tma_load_2d(b_dst + a * 8192, &tm_b, bar, j * p.BN + a * 64, (tap - p.tap0) * p.Kc + kb * 64);
        } else {
          tma_load_2d(b_dst, &tm_b, bar, kb * 64, ((p.tap_rev ? 8 - tap : tap) - p.tap0) * p.Cout + j * p.BN);
        }
        if (++st == nstages) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16(128, p.BN, 0, p.b_mn);
    const uint32_t a_hi = (uint32_t)(smem_desc_sw128(0, 0, 1024) >> 32);
    const uint32_t b_hi = a_hi;
    const uint32_t b_ks = p.b_mn ? (2048u >> 4) : 2u;
    const uint32_t b_lbo = p.b_mn ? ((8192u >> 4) << 16) : 0u;
    int st = 0, ph = 0;
    for (int u = u0; u < u1; ++u) {
      mbar_wait(smem_u32(&bar_full[st]), ph);
      tc_fence_after();
      const uint32_t a_lo = (smem0 + st * stage_bytes) >> 4;
      const uint32_t b_lo = (((smem0 + st * stage_bytes + A_BYTES) >> 4) | b_lbo);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_lohi(tmem_base, a_lo + ks * 2, a_hi, b_lo + ks * b_ks, b_hi, idesc, (u > u0 || ks > 0) ? 1u : 0u);
        umma_commit(smem_u32(&bar_empty[st]));
      }
      __syncwarp();
      if (++st == nstages) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(smem_u32(&bar_done));
    __syncwarp();
  } else {
    const int q = warp % 4;
    const int r = q * 32 + lane;
    float* dst = p.ws + ((long long)split * p.m_tiles * 128 + (long long)mt * 128 + r) * p.Cout + j * p.BN;
    if (u1 > u0) {
      mbar_wait(smem_u32(&bar_done), 0);
      tc_fence_after();
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                 __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
      }
    } else {
      for (int c0 = 0; c0 < p.BN; c0 += 4) *reinterpret_cast<float4*>(dst + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.BN < 32u ? 32u : (uint32_t)p.BN); }
}

struct FinParams {
  const float* ws; int splits, m_tiles, Cout;
  int tiles_w, tiles_h, bw, bh, bn;
  int N, H, W;
  const float* bias; int act, accumulate;
  __nv_bfloat16* y; long long ysn, ysh, ysw;
};

// y[n,h,w,c] (+)= act(sum_s ws[s][row][c] + bias[c]).  A block owns 64 (row, 8-channel chunk) items; its
// 256 threads are 4 split-groups x 64 items, so the split loop runs 4-wide (and 2-deep unrolled) and is
// combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
conv_gemm_finalize_kernel(const FinParams p) {
  pdl_sync();
  __shared__ float s_part[3][64][9];
  const int chunks = p.Cout / 8;
  const long long total = (long long)p.m_tiles * 128 * chunks;
  const long long slab = (long long)p.m_tiles * 128 * p.Cout;
  const int item = threadIdx.x % 64, grp = threadIdx.x / 64;
  for (long long base = (long long)blockIdx.x * 64; base < total; base += (long long)gridDim.x * 64) {
    const long long i = base + item;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    int ck = 0;
    long long row = 0;
    if (i < total) {
      ck = (int)(i % chunks);
      row = i / chunks;
      const float* src = p.ws + row * p.Cout + ck * 8;
      int s = grp;
      for (; s + 4 < p.splits; s += 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(src + (long long)s * slab);
        const float4 b0 = *reinterpret_cast<const float4*>(src + (long long)s * slab + 4);
        const float4 a1 = *reinterpret_cast<const float4*>(src + (long long)(s + 4) * slab);
        const float4 b1 = *reinterpret_cast<const float4*>(src + (long long)(s + 4) * slab + 4);
        acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
        acc[4] += b0.x; acc[5] += b0.y; acc[6] += b0.z; acc[7] += b0.w;
        acc[0] += a1.x; acc[1] += a1.y; acc[2] += a1.z; acc[3] += a1.w;
        acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
      }
      for (; s < p.splits; s += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(src + (long long)s * slab);
        const float4 b0 = *reinterpret_cast<const float4*>(src + (long long)s * slab + 4);
        acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
        acc[4] += b0.x; acc[5] += b0.y; acc[6] += b0.z; acc[7] += b0.w;
      }
    }
    if (grp > 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s_part[grp - 1][item][k] = acc[k];
    }
    __syncthreads();
    if (grp == 0 && i < total) {
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += s_part[g][item][k];
      const int r = (int)(row % 128);
      int t = (int)(row / 128);
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      const int g = t / p.tiles_h;
      const int wi = r % p.bw, hi = (r / p.bw) % p.bh, ni = r / (p.bw * p.bh);
      const int n = g * p.bn + ni, h = th * p.bh + hi, w = tw * p.bw + wi;
      if (n < p.N && h < p.H && w < p.W) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += p.bias ? p.bias[ck * 8 + k] : 0.f;
        if (p.act == B200_ACT_RELU) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.f);
        }
        __nv_bfloat16* dst = p.y + (long long)n * p.ysn + (long long)h * p.ysh + (long long)w * p.ysw + ck * 8;
        if (p.accumulate) {
          float e[8];
          Vec8<__nv_bfloat16>::load(dst, e);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += e[k];
        }
        Vec8<__nv_bfloat16>::store(dst, acc);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// filter gradient of small-spatial layers: one CTA = (split, tap, 128 ci, 128 co)
// ---------------------------------------------------------------------------
constexpr int WG_STAGE = 4 * A_BYTES;   // x: 2 x 64 ci, dz: 2 x 64 co
constexpr int WG_STAGES = 3;

struct WgSmallParams {
  int Cin, Cout, cblocks, oblocks, nlive, m_tiles, splits;
  int tiles_w, tiles_h, bw, bh, bn, ntaps;
  int live_taps[9];
  float* out;   // [splits][ntaps][Cin][Cout]
  int atomic;   // 1: `out` is dw[ntaps][Cin][Cout] itself, every split adds into it with vector atomics
};

__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_small_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dz,
                   const WgSmallParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[WG_STAGES], bar_empty[WG_STAGES], bar_done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

  int b = blockIdx.x;
  const int ob = b % p.oblocks; b /= p.oblocks;
  const int cb = b % p.cblocks; b /= p.cblocks;
  const int tl = b % p.nlive;
  const int split = b / p.nlive;
  const int tap = p.live_taps[tl];
  const int t_begin = (int)((long long)p.m_tiles * split / p.splits);
  const int t_end = (int)((long long)p.m_tiles * (split + 1) / p.splits);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(smem_u32(&bar_full[i]), 1); mbar_init(smem_u32(&bar_empty[i]), 1); }
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_dz); }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), 128); tmem_relinquish(); }
  pdl_sync();      // (PDL) global memory only from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int mt = t_begin; mt < t_end; ++mt) {
        int t = mt;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int n0 = (t / p.tiles_h) * p.bn;
        mbar_wait(smem_u32(&bar_empty[st]), ph ^ 1);
        const uint32_t base = smem0 + st * WG_STAGE;
        const uint32_t bar = smem_u32(&bar_full[st]);
        mbar_arrive_expect_tx(bar, WG_STAGE);
        const int xw = tw * p.bw + tap % 3 - 1, xh = th * p.bh + tap / 3 - 1;
        tma_load_4d(base, &tm_x, bar, cb * 128, xw, xh, n0);
        tma_load_4d(base + A_BYTES, &tm_x, bar, cb * 128 + 64, xw, xh, n0);
        tma_load_4d(base + 2 * A_BYTES, &tm_dz, bar, ob * 128, tw * p.bw, th * p.bh, n0);
        tma_load_4d(base + 3 * A_BYTES, &tm_dz, bar, ob * 128 + 64, tw * p.bw, th * p.bh, n0);
        if (++st == WG_STAGES) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // both operands MN-major: rows of the boxes are K (pixels), 64-channel atoms A_BYTES apart (LBO)
    const uint32_t idesc = idesc_bf16(128, 128, 1, 1);
    const uint32_t hi = (uint32_t)(smem_desc_sw128(0, 0, 1024) >> 32);
    const uint32_t lbo = ((uint32_t)A_BYTES >> 4) << 16;
    int st = 0, ph = 0;
    for (int mt = t_begin; mt < t_end; ++mt) {
      mbar_wait(smem_u32(&bar_full[st]), ph);
      tc_fence_after();
      const uint32_t a_lo = ((smem0 + st * WG_STAGE) >> 4) | lbo;
      const uint32_t b_lo = ((smem0 + st * WG_STAGE + 2 * A_BYTES) >> 4) | lbo;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // 16 pixel rows = 2048 bytes per K step
          umma_lohi(tmem_base, a_lo + ks * (2048u >> 4), hi, b_lo + ks * (2048u >> 4), hi, idesc,
                    (mt > t_begin || ks > 0) ? 1u : 0u);
        umma_commit(smem_u32(&bar_empty[st]));
      }
      __syncwarp();
      if (++st == WG_STAGES) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(smem_u32(&bar_done));
    __syncwarp();
  } else {
    const int q = warp % 4;
    const int ci = q * 32 + lane;      // TMEM lane = input channel of the block
    const int tslot = p.ntaps == 1 ? 0 : tap;
    float* dst = p.out + (((long long)(p.atomic ? 0 : split) * p.ntaps + tslot) * p.Cin + cb * 128 + ci) * p.Cout + ob * 128;
    if (t_end > t_begin) {
      mbar_wait(smem_u32(&bar_done), 0);
      tc_fence_after();
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (p.atomic) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            red_add_v4(dst + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                       __uint_as_float(v[i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                   __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
      }
    } else if (!p.atomic) {
      for (int c0 = 0; c0 < 128; c0 += 4) *reinterpret_cast<float4*>(dst + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

inline int live_taps_of(const b200_tensor* x, int ks, int* taps) {
  int n = 0;
  for (int t = 0; t < 9; ++t) {
    const bool dead = (x->h == 1 && t / 3 != 1) || (x->w == 1 && t % 3 != 1) || (ks == 1 && t != 4);
    if (!dead) taps[n++] = t;
  }
  return n;
}

inline bool small_spatial(const b200_tensor* x) { return x->h <= 8 && x->w <= 8; }

struct GemmPlan { Geom g; int BN, n_tiles, KB, nlive, units, splits; int taps[9]; size_t ws_bytes; };

inline GemmPlan gemm_plan(const b200_tensor* x, int cin, int cout, int ks) {
  GemmPlan pl;
  pl.g = geom_of(x);
  pl.BN = cout % 128 == 0 ? 128 : 64;
  pl.n_tiles = cout / pl.BN;
  pl.KB = cin / 64;
  pl.nlive = live_taps_of(x, ks, pl.taps);
  pl.units = pl.KB * pl.nlive;
  const int base = pl.g.m_tiles * pl.n_tiles;
  int splits = (sm_count() + base - 1) / base;           // about one CTA per SM ...
  if (splits > pl.units / 4) splits = pl.units / 4;      // ... of at least four K units each
  if (splits < 1) splits = 1;
  pl.splits = splits;
  pl.ws_bytes = sizeof(float) * (size_t)splits * pl.g.m_tiles * 128 * cout;
  return pl;
}

}  // namespace

// ---- fprop / dgrad ---------------------------------------------------------------
bool conv_gemm_wanted(const b200_tensor* x, int cin, int cout, int ks) {
  static const int enabled = getenv("B200_CONV_GEMM") ? atoi(getenv("B200_CONV_GEMM")) : 1;
  // measured (tools/conv_table.py, C2): at 8x8 the halo-window kernel is still ahead for fprop/dgrad (its
  // activation traffic is 9x smaller); from 4x4 down the weight-streaming split-K form wins by 2-3x
  if (!enabled || x->h > 4 || x->w > 4 || cin % 64 != 0 || cout % 64 != 0) return false;
  // worth it when the weights, not the activations, dominate the traffic
  return (long long)cin * cout >= 128LL * 128;
}

size_t conv_gemm_workspace(const b200_tensor* x, int cin, int cout, int ks) {
  return gemm_plan(x, cin, cout, ks).ws_bytes;
}

// The split-K scratch is passed per call (the caller -- one Plan -- owns it and orders the launches that share it on one
// stream): which path a layer takes is a pure function of its shapes, never of what an earlier caller registered.
int conv_gemm_launch(const b200_tensor* x, const void* wmat, int cin, int cout, int tap_rev, int b_mn, const float* bias,
                     const b200_tensor* y, int act, int accumulate, int ks, void* ws, size_t ws_bytes, cudaStream_t st) {
  const GemmPlan pl = gemm_plan(x, cin, cout, ks);
  B200_REQUIRE(ws && ws_bytes >= pl.ws_bytes && (uintptr_t)ws % 16 == 0, B200_ERR_BAD_ARG,
               "conv (split-K, images <= 4x4): this layer needs %zu bytes of 16-byte aligned scratch (b200_conv2d_workspace), "
               "got %zu", pl.ws_bytes, ws ? ws_bytes : (size_t)0);
  GemmParams p;
  p.KB = pl.KB; p.BN = pl.BN; p.n_tiles = pl.n_tiles; p.m_tiles = pl.g.m_tiles; p.splits = pl.splits; p.units = pl.units;
  p.nlive = pl.nlive;
  for (int i = 0; i < 9; ++i) p.live_taps[i] = i < pl.nlive ? pl.taps[i] : 4;
  p.tiles_w = pl.g.tiles_w; p.tiles_h = pl.g.tiles_h; p.bw = pl.g.bw; p.bh = pl.g.bh; p.bn = pl.g.bn;
  p.b_mn = b_mn; p.tap_rev = tap_rev; p.Kc = cin; p.Cout = cout;
  p.ntaps = ks == 1 ? 1 : 9; p.tap0 = ks == 1 ? 4 : 0;
  p.ws = reinterpret_cast<float*>(ws);
  CUtensorMap tm_x, tm_b;
  int rc = make_act_tmap(&tm_x, x, p.bw, p.bh, p.bn);
  if (rc) return rc;
  rc = b_mn ? make_mat_tmap(&tm_b, wmat, (long long)p.ntaps * cin, cout, 64)
            : make_mat_tmap(&tm_b, wmat, (long long)p.ntaps * cout, cin, p.BN);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)MAX_STAGES * (A_BYTES + p.BN * 128);
  if (first_use_on_device(0))
    cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + MAX_STAGES * (A_BYTES + 128 * 128));
  const int grid = p.splits * p.m_tiles * p.n_tiles;
  launch_pdl(conv_gemm_kernel, grid, NTHREADS, smem, st, tm_x, tm_b, p);
  rc = check_launch("conv_gemm_kernel");
  if (rc) return rc;
  FinParams f;
  f.ws = p.ws; f.splits = p.splits; f.m_tiles = p.m_tiles; f.Cout = cout;
  f.tiles_w = p.tiles_w; f.tiles_h = p.tiles_h; f.bw = p.bw; f.bh = p.bh; f.bn = p.bn;
  f.N = y->n; f.H = y->h; f.W = y->w;
  f.bias = bias; f.act = act; f.accumulate = accumulate;
  f.y = reinterpret_cast<__nv_bfloat16*>(y->data); f.ysn = y->stride_n; f.ysh = y->stride_h; f.ysw = y->stride_w;
  const long long total = (long long)p.m_tiles * 128 * (cout / 8);
  long long blocks = (total + 63) / 64;
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  launch_pdl(conv_gemm_finalize_kernel, (int)blocks, 256, 0, st, f);
  return check_launch("conv_gemm_finalize_kernel");
}

// ---- wgrad -----------------------------------------------------------------------
struct WgSmallPlan { Geom g; int nlive, splits; int taps[9]; };

static WgSmallPlan wg_small_plan(const b200_tensor* x, const b200_tensor* dy, int ks) {
  WgSmallPlan pl;
  pl.g = geom_of(x);
  pl.nlive = live_taps_of(x, ks, pl.taps);
  const int base = pl.nlive * (x->c / 128) * (dy->c / 128);
  int splits = (sm_count() + base - 1) / base;
  if (splits > pl.g.m_tiles) splits = pl.g.m_tiles;
  if (splits < 1) splits = 1;
  pl.splits = splits;
  return pl;
}

bool wgrad_small_wanted(const b200_tensor* x, const b200_tensor* dy, int ks) {
  static const int enabled = getenv("B200_CONV_GEMM") ? atoi(getenv("B200_CONV_GEMM")) : 1;
  return enabled && small_spatial(x) && x->c % 128 == 0 && dy->c % 128 == 0;
}

size_t wgrad_small_workspace(const b200_tensor* x, const b200_tensor* dy, int ks) {
  const WgSmallPlan pl = wg_small_plan(x, dy, ks);
  if (pl.splits == 1) return 0;
  return sizeof(float) * (size_t)pl.splits * (ks == 1 ? 1 : 9) * x->c * dy->c;
}

// returns the number of splits written ([splits][ntaps][Cin][Cout] in `out`); dead taps are left untouched
int wgrad_small_launch(const b200_tensor* x, const b200_tensor* dy, float* out, int ks, int* splits_out, int* live_mask,
                       cudaStream_t st, int atomic) {
  const WgSmallPlan pl = wg_small_plan(x, dy, ks);
  WgSmallParams p;
  p.Cin = x->c; p.Cout = dy->c; p.cblocks = x->c / 128; p.oblocks = dy->c / 128; p.nlive = pl.nlive;
  p.m_tiles = pl.g.m_tiles; p.splits = pl.splits;
  p.tiles_w = pl.g.tiles_w; p.tiles_h = pl.g.tiles_h; p.bw = pl.g.bw; p.bh = pl.g.bh; p.bn = pl.g.bn;
  p.ntaps = ks == 1 ? 1 : 9;
  int mask = 0;
  for (int i = 0; i < 9; ++i) {
    p.live_taps[i] = i < pl.nlive ? pl.taps[i] : 4;
    if (i < pl.nlive) mask |= 1 << (ks == 1 ? 0 : pl.taps[i]);
  }
  p.out = out;
  p.atomic = atomic;
  CUtensorMap tm_x, tm_dz;
  int rc = make_act_tmap(&tm_x, x, p.bw, p.bh, p.bn);
  if (rc) return rc;
  rc = make_act_tmap(&tm_dz, dy, p.bw, p.bh, p.bn);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)WG_STAGES * WG_STAGE;
  if (first_use_on_device(1))
    cudaFuncSetAttribute(wgrad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = p.splits * p.nlive * p.cblocks * p.oblocks;
  launch_pdl(wgrad_small_kernel, grid, NTHREADS, smem, st, tm_x, tm_dz, p);
  *splits_out = p.splits;
  *live_mask = mask;
  return check_launch("wgrad_small_kernel");
}

}  // namespace b200
