// 3x3 "same" convolution as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma, fp32 accumulators in TMEM), operands staged by TMA.  One kernel
// serves fprop and dgrad (dgrad = the same convolution with the tap order reversed
// and the HWIO kernel read as [tap][cin][cout]).
//
// Replaces the cuDNN kernels TF dispatches for keras Conv2D at
// Super_resolution/code/train_adaptive_unet.py:202,207,259 (reference: /root/reference)
// and their autodiff (dgrad).
//
// Formulation (per CTA work item):
//   output tile  : 16 rows x 8 columns of one image = 128 GEMM rows (M), BN output channels (N)
//   A operand    : for each 64-channel block of Cin, ONE TMA box {64ch, 10, 18, 1} -- the tile plus
//                  its one-pixel halo -- lands in shared memory with the 128-byte swizzle; image
//                  borders are the TMA's out-of-bounds zero fill.  All nine filter taps read that
//                  one window through shifted UMMA descriptors: tap (kh,kw) starts (kh*10+kw) pixel
//                  rows (128 B each) into the window and strides 8-row groups by the window pitch
//                  (SBO = 10*128 B), so the activations are fetched from L2 once, not nine times.
//   B operand    : [tap][N][K] K-major weight tiles {64 x BN}; kept resident in shared memory for the
//                  whole (persistent) CTA when all of them fit, streamed through a ring otherwise.
//   accumulators : 2 x BN TMEM columns (double buffered: epilogue of item i overlaps MMAs of i+1).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM owner,
// warps 2..5 = epilogue (TMEM -> registers -> bias/activation -> bf16 -> global).
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

using namespace ptx;

// ---------------------------------------------------------------------------
// TMA descriptor helpers (driver entry point fetched through the runtime)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC bf16 activation view -> 4D map {C, W, H, N}, box {64, bw, bh, bn}, 128B swizzle, zero OOB fill
int make_act_tmap(CUtensorMap* m, const b200_tensor* t, int box_w, int box_h, int box_n) {
  EncodeTiledFn fn = encode_fn();
  B200_REQUIRE(fn, B200_ERR_LAUNCH, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->stride_w * 2, (cuuint64_t)t->stride_h * 2, (cuuint64_t)t->stride_n * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_LAUNCH, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return B200_OK;
}

// NHWC bf16 activation view -> 4D map with the image index BEFORE the row index: {C, W, N, H}, box {64, bw, bn, bh}.
// A box then lands in shared memory as [row][image][column][channel]: the rows of `bn` small images interleaved,
// which keeps the 8-pixel row groups of a two-image tile equally spaced (see "pair" tiles in conv3x3_tc_kernel).
int make_act_tmap_nh(CUtensorMap* m, const b200_tensor* t, int box_w, int box_n, int box_h) {
  EncodeTiledFn fn = encode_fn();
  B200_REQUIRE(fn, B200_ERR_LAUNCH, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->n, (cuuint64_t)t->h};
  cuuint64_t strides[3] = {(cuuint64_t)t->stride_w * 2, (cuuint64_t)t->stride_n * 2, (cuuint64_t)t->stride_h * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_n, (cuuint32_t)box_h};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_LAUNCH, "cuTensorMapEncodeTiled(activation, N before H) failed: %d", (int)r);
  return B200_OK;
}

// row-major bf16 matrix [rows][cols] -> 2D map, box {64 cols, box_rows}, 128B swizzle
int make_mat_tmap(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  B200_REQUIRE(fn, B200_ERR_LAUNCH, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_LAUNCH, "cuTensorMapEncodeTiled(matrix) failed: %d", (int)r);
  return B200_OK;
}

namespace {

constexpr int TILE_H = 16, TILE_W = 8;            // 128 output pixels
constexpr int WIN_W = TILE_W + 2, WIN_H = TILE_H + 2;
constexpr int WIN_BYTES = WIN_W * WIN_H * 128;    // 23040
constexpr int WIN_STAGE = 23552;                  // padded to a multiple of 1024
constexpr int WIN_PITCH = WIN_W * 128;            // 1280: byte distance between tile rows = SBO
constexpr int PAIR_WIN_STAGE = WIN_W * 2 * (8 + 2) * 128;   // 25600: two interleaved 8-row images + halos
constexpr int MAX_WSLOTS = 18;
constexpr int NTHREADS = 192;
constexpr int NTHREADS2 = 320;   // LayerNorm-64 instantiation: a second group of four epilogue warps (two warps per scheduler)
constexpr int SMEM_LIMIT = 232448;                // 227 KB

struct ConvTcParams {
  int N, H, W, Cout, KB, BN, n_tiles, tiles_h, tiles_w, total_items;
  int tap_rev, live_mask, resident, nsw, nsb, act, accumulate;
  int debug;  // B200_CONV_DEBUG: 1 = epilogue only waits/releases (profiling aid)
  // output path: 1 = tiles are staged in shared memory (128-byte swizzle) and written by TMA stores
  // (ring of `nslots` 16 KB slots, one slot = 128 pixels x 64 channels); 0 = per-thread stores (accumulate)
  int tma_store, nslots;
  // epilogue groups (EPI == 1 only): 2 = two 4-warp groups alternate tiles, group g owns accumulator stage g and half of
  // the slots, so the LayerNorm epilogue's dependent-instruction latency overlaps between two warps per scheduler
  int egroups;
  // CTA pairs (cta_group::2): the two CTAs of a cluster work on two neighbouring pixel tiles with ONE M = 256 MMA stream
  // issued by the leader; each CTA stages its own window and its HALF of every weight tile (wt_bytes = BN/2 rows)
  int cg2, wt_bytes;
  int b_mn;   // 1: B operand is MN-major (fprop reads the Keras HWIO kernel [tap][cin][cout] as is); 0: K-major (dgrad)
  int Kc;     // K total (input channels of this convolution)
  int ntaps, tap0;  // 9, 0 for a 3x3 filter; 1, 4 for a 1x1 filter (centre tap only; its weights are matrix block 0)
  // small images (H <= 7) are stacked: one tile holds `nb` images, each `srows` = H+2 window rows
  int nb, srows, win_bytes;
  // pair tiles (8-row images): two images per tile with their rows INTERLEAVED in the window ([row][image][10 px]), so
  // that row group g = row*2 + image sits g*1280 bytes into the window for every tap; tap (kh,kw) starts (kh*20+kw) pixels in
  int pair, win_stage;
  const float* bias;
  __nv_bfloat16* y;
  long long ysn, ysh, ysw;
  // fused LayerNormalization(axis=-1) [+ReLU] epilogue (single N tile only): z = bf16(conv + bias) is
  // stored for the backward pass, y = act(LN(z)) for the next layer, mean/rstd per pixel in fp32
  int ln, ln_relu;
  float ln_eps;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* z;          // may be NULL (inference)
  long long zsn, zsh, zsw;
  float* mean;
  float* rstd;
  // fused LayerNormalization(+ReLU) BACKWARD epilogue of a dgrad launch (EPI == 4; CTA pairs, N = 64): the accumulator is
  // dy of a LayerNorm output; with the saved z (p.z, read), mean / rstd (read), gamma / beta the epilogue writes dz instead
  // and sums d(gamma), d(beta) and d(bias of the conv that produced z) over its pixels
  int lnb;
  float* dgamma;
  float* dbeta;
  float* dbias;
};

// D[tmem] (+)= A * B^T with the descriptors given as (lo, hi) words: the hi words are loop
// invariants and the lo words advance by plain 32-bit adds in the single issuing thread.
template <bool CG2 = false>
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  if (CG2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// One {BN x 64} weight tile for (64-channel K block kb, tap, N tile j) into shared memory.
//   K-major (dgrad):  matrix rows = tap*Cout + n, cols = k      -> one box {64 k, BN rows}
//   MN-major (fprop): matrix rows = tap*K + k,   cols = n (HWIO) -> BN/64 boxes {64 n, 64 k rows}, 8 KB apart
// CTA pairs: CTA `crank` loads ITS half of the tile (N columns [crank*BN/2, (crank+1)*BN/2)) and signals the leader's barrier.
template <bool CG2>
__device__ __forceinline__ void load_weight_tile(const ConvTcParams& p, const CUtensorMap* tm_b, uint32_t dst,
                                                 uint32_t bar, int kb, int tap, int j, uint32_t crank) {
  if (CG2) {
    if (p.b_mn)   // BN = 128 only: one 64-column atom per CTA
      tma_load_2d_cg2(dst, tm_b, bar, j * p.BN + (int)crank * 64, (tap - p.tap0) * p.Kc + kb * 64);
    else
      tma_load_2d_cg2(dst, tm_b, bar, kb * 64,
                      ((p.tap_rev ? 8 - tap : tap) - p.tap0) * p.Cout + j * p.BN + (int)crank * (p.BN / 2));
  } else if (p.b_mn) {
    for (int a = 0; a < p.BN / 64; ++a)
      tma_load_2d(dst + a * 8192, tm_b, bar, j * p.BN + a * 64, (tap - p.tap0) * p.Kc + kb * 64);
  } else {
    tma_load_2d(dst, tm_b, bar, kb * 64, ((p.tap_rev ? 8 - tap : tap) - p.tap0) * p.Cout + j * p.BN);
  }
}

__device__ __forceinline__ void decode_item(const ConvTcParams& p, int item, int& j, int& tw, int& th, int& n) {
  j = item % p.n_tiles;
  int t = item / p.n_tiles;
  tw = t % p.tiles_w; t /= p.tiles_w;
  th = t % p.tiles_h;
  n = (t / p.tiles_h) * p.nb;     // first image of the tile
}

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int SLOT_BYTES = 128 * 128;   // 128 pixels x 64 bf16 channels
constexpr int EPI_THREADS = 128;

// Output staging ring of the epilogue warps.  A "job" is one 128-pixel x 64-channel box: the 128
// epilogue threads each write their pixel's 128 bytes into a slot (swizzled like the TMA expects),
// and ONE thread (the issuer) hands the slot to the TMA engine.  One named barrier per job: passing
// the barrier of job k proves that every thread finished (and proxy-fenced) its writes of job k-1,
// so the issuer launches the store of job k-1 right after it, and that the slot of job k is free
// (the issuer waited for the bulk group of job k-nslots before arriving).
struct StoreRing {
  uint32_t stg0;
  int nslots, slot;
  bool issuer;
  int d_have, d_which, d_c, d_w, d_h, d_n;
  uint32_t d_slot;
  const CUtensorMap* tm_y;
  const CUtensorMap* tm_z;
  uint32_t bar_id;     // named barrier of this epilogue group

  __device__ __forceinline__ void flush() {
    if (issuer && d_have) {
      if (dbg != 6) {
        tma_store_4d(d_which ? tm_z : tm_y, d_slot, d_c, d_w, d_h, d_n);
        bulk_commit();
      }
      d_have = 0;
    }
  }
  int dbg;
  __device__ __forceinline__ uint32_t begin() {
    if (dbg == 4) { flush(); return stg0 + (uint32_t)slot * SLOT_BYTES; }
    if (issuer) {
      if (nslots == 2) bulk_wait_read<0>();
      else if (nslots == 3) bulk_wait_read<1>();
      else bulk_wait_read<2>();
    }
    __syncwarp();
    named_bar_sync(bar_id, EPI_THREADS);
    flush();
    return stg0 + (uint32_t)slot * SLOT_BYTES;
  }
  __device__ __forceinline__ void end(uint32_t slot_addr, int which, int c, int w, int h, int n) {
    if (dbg != 3) fence_proxy_async();
    if (issuer) { d_have = 1; d_which = which; d_slot = slot_addr; d_c = c; d_w = w; d_h = h; d_n = n; }
    if (++slot == nslots) slot = 0;
  }
  __device__ __forceinline__ void drain() {
    __syncwarp();
    named_bar_sync(bar_id, EPI_THREADS);
    flush();
    if (issuer) bulk_wait_all();
  }
};

// this thread's pixel row (64 channels, value i = f(i)) -> its 128 bytes of the slot, 16-byte chunks
// XOR-swizzled by row (the layout a SWIZZLE_128B tensor map expects)
template <typename F>
__device__ __forceinline__ void stage_row(uint32_t slot_addr, int r, F f) {
  const uint32_t base = slot_addr + (uint32_t)r * 128u;
  const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st_shared_v4(base + (((uint32_t)c ^ sw) << 4), pack_bf16(f(c * 8 + 0), f(c * 8 + 1)), pack_bf16(f(c * 8 + 2), f(c * 8 + 3)),
                 pack_bf16(f(c * 8 + 4), f(c * 8 + 5)), pack_bf16(f(c * 8 + 6), f(c * 8 + 7)));
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// LayerNorm epilogue for one pixel (one thread) with all BN_T channels held in registers as packed
// bf16 pairs (z is stored in bf16 anyway, so the packed words ARE the z tile and cost half the registers):
//   z = bf16(acc + bias)            (what keras stores under the mixed policy; statistics are taken on it)
//   y = act((z - mean) * rstd * gamma + beta)
// Returns after the TMEM reads; the caller releases the accumulator stage, then stages z and y.
template <int BN_T>
__device__ __forceinline__ void ln_load(uint32_t taddr, const float* s_bias, uint32_t (&zp)[BN_T / 2], float& mean,
                                        float& rstd, float eps) {
  float sum = 0.f;
#pragma unroll
  for (int c0 = 0; c0 < BN_T; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const uint32_t w = pack_bf16(__uint_as_float(v[i]) + s_bias[c0 + i], __uint_as_float(v[i + 1]) + s_bias[c0 + i + 1]);
      zp[(c0 + i) / 2] = w;
      sum += bf16_lo(w) + bf16_hi(w);
    }
  }
  mean = sum * (1.f / BN_T);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < BN_T / 2; ++i) {
    const float d0 = bf16_lo(zp[i]) - mean, d1 = bf16_hi(zp[i]) - mean;
    q += d0 * d0 + d1 * d1;
  }
  rstd = rsqrtf(q * (1.f / BN_T) + eps);
}

// EPI selects which epilogues an instantiation contains (the LayerNorm epilogue's code generation is sensitive to
// what else lives in the kernel): 0 = all, 1 = LayerNorm over 64 channels, 2 = LayerNorm over 128, 3 = no LayerNorm
template <bool PAIR, int EPI, bool CG2 = false>
__global__ void __launch_bounds__((EPI == 1 || EPI == 4) ? NTHREADS2 : NTHREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_z,
                  const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full_w[8], bar_empty_w[8];
  __shared__ __align__(8) uint64_t bar_full_b[MAX_WSLOTS], bar_empty_b[MAX_WSLOTS];
  __shared__ __align__(8) uint64_t bar_tmem_full[2], bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t win0 = smem0;
  const uint32_t wt0 = smem0 + (uint32_t)p.nsw * (uint32_t)(PAIR ? PAIR_WIN_STAGE : WIN_STAGE);
  const uint32_t wt_bytes = (uint32_t)p.wt_bytes;
  const uint32_t tmem_cols = 2u * (uint32_t)p.BN;
  const uint32_t stg0 = wt0 + (uint32_t)p.nsb * wt_bytes;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (stg0 - smem_u32(smem_raw)) + (size_t)p.nslots * SLOT_BYTES);
  // CTA pair: rank within the cluster (0 = leader, the MMA issuer), and the work-item walk of this CTA: item q of the
  // cluster is the tile pair (2t, 2t+1) for N tile j; this CTA takes tile 2t + crank (a tile past the end is all padding)
  const uint32_t crank = CG2 ? cluster_ctarank() : 0u;
  const int q0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int qstep = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto item_of = [&](int q) { return CG2 ? ((q / p.n_tiles) * 2 + (int)crank) * p.n_tiles + q % p.n_tiles : q; };
  // "full" barriers and the accumulator-free barriers live in the LEADER and count arrivals of both CTAs
  auto lead = [&](uint64_t* bar) { return CG2 ? mapa_shared(smem_u32(bar), 0) : smem_u32(bar); };
  // (plain CTA-scope waits also for barriers the peer arrives on: what they order is the async proxy's shared-memory /
  // tensor-memory traffic, fenced by tcgen05.fence; a cluster-scope acquire would cost an L1 invalidate -- CCTL.IVALL -- per wait)
  auto wait = [&](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };

  if (threadIdx.x == 0) {
    const uint32_t np = CG2 ? 2u : 1u;
    for (int i = 0; i < p.nsw; ++i) { mbar_init(smem_u32(&bar_full_w[i]), np); mbar_init(smem_u32(&bar_empty_w[i]), 1); }
    for (int i = 0; i < p.nsb; ++i) { mbar_init(smem_u32(&bar_full_b[i]), np); mbar_init(smem_u32(&bar_empty_b[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tmem_full[i]), 1); mbar_init(smem_u32(&bar_tmem_empty[i]), 4 * np); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_x); prefetch_tmap(&tm_b);
    if (p.tma_store) { prefetch_tmap(&tm_y); if (p.z) prefetch_tmap(&tm_z); }
  }
  if (CG2) {   // both CTAs are running and their barriers initialised before anyone arrives remotely / allocates as a pair
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tmem_alloc_cg2(smem_u32(&tmem_base_smem), tmem_cols); tmem_relinquish_cg2(); }
  } else if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), tmem_cols); tmem_relinquish(); }
  // ---- everything above overlapped the previous kernel's tail (PDL); global memory is touched only from here on ----
  pdl_sync();
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
    s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (p.ln || p.lnb) { s_bias[p.Cout + i] = p.gamma[i]; s_bias[2 * p.Cout + i] = p.beta[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int sw = 0, pw = 0, sb = 0, pb = 0;
      auto expect = [&](uint32_t bar, uint32_t bytes) {
        if (CG2) mbar_arrive_expect_tx_cluster(bar, bytes); else mbar_arrive_expect_tx(bar, bytes);
      };
      if (p.resident) {
        for (int kb = 0; kb < p.KB; ++kb)
          for (int tap = 0; tap < 9; ++tap) {
            if (!((p.live_mask >> tap) & 1)) continue;
            const int slot = kb * p.ntaps + tap - p.tap0;
            const uint32_t bar = lead(&bar_full_b[slot]);
            expect(bar, wt_bytes);
            load_weight_tile<CG2>(p, &tm_b, wt0 + slot * wt_bytes, bar, kb, tap, 0, crank);
          }
      }
      for (int q = q0; q < p.total_items; q += qstep) {
        int j, tw, th, n;
        decode_item(p, item_of(q), j, tw, th, n);
        for (int kb = 0; kb < p.KB; ++kb) {
          wait(smem_u32(&bar_empty_w[sw]), pw ^ 1);
          const uint32_t bar = lead(&bar_full_w[sw]);
          expect(bar, (uint32_t)p.win_bytes);
          if (PAIR)   // map dims are (C, W, N, H)
            tma_load_4d(win0 + sw * PAIR_WIN_STAGE, &tm_x, bar, kb * 64, tw * TILE_W - 1, n, -1);
          else if (CG2)
            tma_load_4d_cg2(win0 + sw * WIN_STAGE, &tm_x, bar, kb * 64, tw * TILE_W - 1, th * TILE_H - 1, n);
          else
            tma_load_4d(win0 + sw * WIN_STAGE, &tm_x, bar, kb * 64, tw * TILE_W - 1, th * TILE_H - 1, n);
          if (++sw == p.nsw) { sw = 0; pw ^= 1; }
          if (!p.resident) {
            for (int tap = 0; tap < 9; ++tap) {
              if (!((p.live_mask >> tap) & 1)) continue;
              wait(smem_u32(&bar_empty_b[sb]), pb ^ 1);
              const uint32_t bb = lead(&bar_full_b[sb]);
              expect(bb, wt_bytes);
              load_weight_tile<CG2>(p, &tm_b, wt0 + sb * wt_bytes, bb, kb, tap, j, crank);
              if (++sb == p.nsb) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer ==============================
    // All 32 lanes run the loop so that control flow stays warp-uniform (descriptor words are then
    // computed on the uniform datapath); one elected lane issues the MMAs and the commits.
    // CTA pairs: only the leader's warp issues (M = 256 over both CTAs); its commits reach the barriers of both.
    if (!CG2 || crank == 0) {
      const uint32_t idesc = idesc_bf16(CG2 ? 256 : 128, p.BN, 0, p.b_mn);
      const uint32_t a_hi = (uint32_t)(smem_desc_sw128(0, 0, WIN_PITCH) >> 32);
      const uint32_t b_hi = (uint32_t)(smem_desc_sw128(0, 0, 1024) >> 32);
      // per-K-step (16 channels) advance of the B descriptor and its LBO field:
      //   K-major: +32 B inside the 128-byte swizzle row; MN-major: +16 rows of 128 B, 64-column atoms 8 KB apart
      const uint32_t b_ks = p.b_mn ? (2048u >> 4) : 2u;
      const uint32_t b_lbo = p.b_mn ? ((8192u >> 4) << 16) : 0u;
      const uint32_t wt_step = wt_bytes >> 4;
      const bool all_live = p.live_mask == 0x1FF;
      auto commit = [&](uint64_t* bar) { if (CG2) umma_commit_cg2(smem_u32(bar)); else umma_commit(smem_u32(bar)); };
      int sw = 0, pw = 0, sb = 0, pb = 0, as = 0, pa = 0;
      if (p.resident) {   // the weights are loaded once: wait for every tile up front
        for (int kb = 0; kb < p.KB; ++kb)
          for (int tap = 0; tap < 9; ++tap)
            if ((p.live_mask >> tap) & 1) wait(smem_u32(&bar_full_b[kb * p.ntaps + tap - p.tap0]), 0);
        tc_fence_after();
      }
      for (int q = q0; q < p.total_items; q += qstep) {
        wait(smem_u32(&bar_tmem_empty[as]), pa ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < p.KB; ++kb) {
          wait(smem_u32(&bar_full_w[sw]), pw);
          tc_fence_after();
          const uint32_t a_lo0 = (win0 + sw * (PAIR ? PAIR_WIN_STAGE : WIN_STAGE)) >> 4;
          if (!PAIR && p.resident && all_live) {
            // hot path of the full-resolution layers: 36 MMAs, descriptor words by 32-bit adds only
            const uint32_t b_lo0 = ((wt0 >> 4) | b_lbo) + (uint32_t)(kb * 9) * wt_step;
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t a_lo = a_lo0 + (uint32_t)((tap / 3) * WIN_W + (tap % 3)) * 8u;
                const uint32_t b_lo = b_lo0 + (uint32_t)tap * wt_step;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_lohi<CG2>(d_tmem, a_lo + ks * 2, a_hi, b_lo + ks * b_ks, b_hi, idesc, (tap | ks) ? 1u : accumulate);
              }
            }
            accumulate = 1;
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              if (!((p.live_mask >> tap) & 1)) continue;
              uint32_t b_lo;
              if (p.resident) {
                b_lo = ((wt0 >> 4) | b_lbo) + (uint32_t)(kb * p.ntaps + tap - p.tap0) * wt_step;
              } else {
                wait(smem_u32(&bar_full_b[sb]), pb);
                tc_fence_after();
                b_lo = ((wt0 >> 4) | b_lbo) + (uint32_t)sb * wt_step;
              }
              const uint32_t a_lo = a_lo0 + (uint32_t)((tap / 3) * (PAIR ? 2 * WIN_W : WIN_W) + (tap % 3)) * 8u;
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_lohi<CG2>(d_tmem, a_lo + ks * 2, a_hi, b_lo + ks * b_ks, b_hi, idesc, ks ? 1u : accumulate);
                if (!p.resident) commit(&bar_empty_b[sb]);
              }
              accumulate = 1;
              __syncwarp();
              if (!p.resident) { if (++sb == p.nsb) { sb = 0; pb ^= 1; } }
            }
          }
          if (elect_one()) commit(&bar_empty_w[sw]);
          __syncwarp();
          if (++sw == p.nsw) { sw = 0; pw ^= 1; }
        }
        if (elect_one()) commit(&bar_tmem_full[as]);
        __syncwarp();
        if (++as == 2) { as = 0; pa ^= 1; }
      }
    }
  } else {
    // ============================ epilogue ================================
    const int wq = warp % 4;                // TMEM lane quarter this warp may read
    const int r = wq * 32 + lane;           // GEMM row = pixel of the tile
    const int ty = r / TILE_W, tx = r % TILE_W;
    // stacked small images (srows huge otherwise) or pair tiles (rows of two images interleaved)
    const int sb_img = PAIR ? (ty & 1) : ty / p.srows, sb_row = PAIR ? (ty >> 1) : ty % p.srows;
    const int group = (warp - 2) >> 2;      // 0: warps 2..5, 1: warps 6..9 (EPI == 1 launches only)
    const int egroups = (EPI == 1 || EPI == 4) ? p.egroups : 1;
    StoreRing ring;
    ring.nslots = p.nslots / egroups; ring.slot = 0; ring.issuer = threadIdx.x == 64 + group * EPI_THREADS;
    ring.stg0 = stg0 + (uint32_t)(group * ring.nslots) * SLOT_BYTES;
    ring.bar_id = 1 + group;
    ring.d_have = 0; ring.d_which = 0; ring.d_c = ring.d_w = ring.d_h = ring.d_n = 0; ring.d_slot = 0;
    ring.tm_y = &tm_y; ring.tm_z = &tm_z; ring.dbg = p.debug;
    int it = -1;
    float lnb_acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // (EPI == 4) column sums of dz, d, d*xhat for this thread's channel pair
    uint4 zr[EPI == 4 ? 8 : 1], zn[EPI == 4 ? 8 : 1];       // (EPI == 4) saved pre-norm row of the current / the next own tile
    float ln_mu = 0.f, ln_rs = 0.f, ln_mu_n = 0.f, ln_rs_n = 0.f;
    bool lnb_primed = false;
    if (group < egroups)
    for (int q = q0; q < p.total_items; q += qstep) {
      const int item = item_of(q);
      ++it;
      if (egroups == 2 && (it & 1) != group) continue;        // the other group's tile
      const int as = it & 1, pa = (it >> 1) & 1;              // accumulator stage and its mbarrier phase
      int j, tw, th, n;
      decode_item(p, item, j, tw, th, n);
      const int oh = th * TILE_H + sb_row, ow = tw * TILE_W + tx;
      const int img = n + sb_img;
      const bool valid = oh < p.H && ow < p.W && sb_img < p.nb && img < p.N;
      const float* bias = s_bias + j * p.BN;
      // fused LayerNorm backward: the operands that do not come from the MMAs (this pixel's 64 saved pre-norm values, its
      // mean and rstd) are software-pipelined one tile ahead: the loads for the group's NEXT tile are issued here and
      // consumed an epilogue later (the epilogue, not the MMA stream, paces this kernel, so a load issued at the top of
      // its own tile would be waited for in full)
      if (EPI == 4) {
        auto load_z = [&](int qq, uint4* dst, float& mu, float& rs) {
          bool ok = false;
          if (qq < p.total_items) {
            int j2, tw2, th2, n2;
            decode_item(p, item_of(qq), j2, tw2, th2, n2);
            const int oh2 = th2 * TILE_H + sb_row, ow2 = tw2 * TILE_W + tx;
            if (oh2 < p.H && ow2 < p.W && n2 < p.N) {
              ok = true;
              const uint4* zp = reinterpret_cast<const uint4*>(p.z + (long long)n2 * p.zsn + (long long)oh2 * p.zsh + (long long)ow2 * p.zsw);
#pragma unroll
              for (int c = 0; c < 8; ++c) dst[c] = zp[c];
              const long long pixb = ((long long)n2 * p.H + oh2) * p.W + ow2;
              mu = p.mean[pixb]; rs = p.rstd[pixb];
            }
          }
          if (!ok) {
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = make_uint4(0u, 0u, 0u, 0u);
            mu = 0.f; rs = 0.f;
          }
        };
        if (!lnb_primed) { load_z(q, zr, ln_mu, ln_rs); lnb_primed = true; }
        load_z(q + egroups * qstep, zn, ln_mu_n, ln_rs_n);
      }
      wait(smem_u32(&bar_tmem_full[as]), pa);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * p.BN);
      const uint32_t tmem_empty = lead(&bar_tmem_empty[as]);     // CTA pairs: the leader's barrier counts both CTAs' warps
      // the accumulator stage goes back to the MMA warp as soon as this warp holds its values in registers
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG2) mbar_arrive_cluster(tmem_empty); else mbar_arrive(tmem_empty); }
      };
      // box origin (stacked tiles: th == 0, box = {64, 8, srows, nb}); pair tiles: the store map's dims are (C, W, N, H),
      // so the image index goes where the row coordinate normally is and the row origin (0) last
      const int cw = tw * TILE_W, ch = PAIR ? n : th * TILE_H, cn = PAIR ? 0 : n;
      if (p.debug == 1) {
        release_tmem();
      } else if (EPI == 4) {
        // ---- dgrad + LayerNorm(+ReLU) backward: dy (accumulator) -> dz, plus the column sums of d, d*xhat and dz ----
        // pass 1 consumes the accumulator 32 columns at a time and keeps d = masked dy as packed bf16 pairs (what the
        // unfused path stores between the two kernels): 32 registers instead of 64 next to the two pipelined z rows
        uint32_t dp[32];
        const float* gam = s_bias + p.Cout;
        const float* bet = s_bias + 2 * p.Cout;
        const bool relu = p.ln_relu != 0;
        const float nmr = -ln_mu * ln_rs;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        const uint32_t* zw = reinterpret_cast<const uint32_t*>(zr);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t v[32];
          tmem_ld32(taddr + hf * 32, v);
          tmem_ld_wait();
          if (hf == 1) release_tmem();
#pragma unroll
          for (int k = 0; k < 16; ++k) {         // d = masked dy, sums of g = d*gamma and g*xhat
            const int i = hf * 16 + k;
            const float x0 = fmaf(bf16_lo(zw[i]), ln_rs, nmr), x1 = fmaf(bf16_hi(zw[i]), ln_rs, nmr);
            const float g0 = gam[2 * i], g1 = gam[2 * i + 1];
            float d0 = __uint_as_float(v[2 * k]), d1 = __uint_as_float(v[2 * k + 1]);
            if (relu) {
              d0 = fmaf(x0, g0, bet[2 * i]) > 0.f ? d0 : 0.f;
              d1 = fmaf(x1, g1, bet[2 * i + 1]) > 0.f ? d1 : 0.f;
            }
            if (!valid) { d0 = 0.f; d1 = 0.f; }
            dp[i] = pack_bf16(d0, d1);
            d0 = bf16_lo(dp[i]); d1 = bf16_hi(dp[i]);
            const float e0 = d0 * g0, e1 = d1 * g1;
            s1a += e0; s1b += e1;                // (two independent chains per sum)
            s2a = fmaf(e0, x0, s2a); s2b = fmaf(e1, x1, s2b);
          }
        }
        const float c1 = -(s1a + s1b) * (1.f / 64.f) * ln_rs, c2 = -(s2a + s2b) * (1.f / 64.f) * ln_rs;
        // slots of this group: [0] dz (also the TMA store's source), [1] d, [2] d*xhat -- all bf16, 128-byte swizzled rows
        const uint32_t slotA = ring.stg0, slotB = slotA + SLOT_BYTES, slotC = slotB + SLOT_BYTES;
        // barrier 1: everyone finished the column sums of the previous tile and the previous dz store has read its slot
        if (ring.issuer) bulk_wait_read<0>();
        __syncwarp();
        named_bar_sync(ring.bar_id, EPI_THREADS);
        {
          const uint32_t rbase = (uint32_t)r * 128u, swz = (uint32_t)(r & 7);
#pragma unroll
          for (int c = 0; c < 8; ++c) {          // pass 2: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), 8 channels per 16-byte chunk
            uint32_t oa[4], ob[4], oc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i = c * 4 + k;
              const float x0 = fmaf(bf16_lo(zw[i]), ln_rs, nmr), x1 = fmaf(bf16_hi(zw[i]), ln_rs, nmr);
              const float d0 = bf16_lo(dp[i]), d1 = bf16_hi(dp[i]);
              const float z0 = fmaf(x0, c2, fmaf(d0 * gam[2 * i], ln_rs, c1));
              const float z1 = fmaf(x1, c2, fmaf(d1 * gam[2 * i + 1], ln_rs, c1));
              oa[k] = pack_bf16(valid ? z0 : 0.f, valid ? z1 : 0.f);
              ob[k] = dp[i];
              oc[k] = pack_bf16(d0 * x0, d1 * x1);
            }
            const uint32_t off = rbase + (((uint32_t)c ^ swz) << 4);
            st_shared_v4(slotA + off, oa[0], oa[1], oa[2], oa[3]);
            st_shared_v4(slotB + off, ob[0], ob[1], ob[2], ob[3]);
            st_shared_v4(slotC + off, oc[0], oc[1], oc[2], oc[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        named_bar_sync(ring.bar_id, EPI_THREADS);      // barrier 2: the three tiles are complete
        if (ring.issuer) { tma_store_4d(&tm_y, slotA, 0, cw, ch, cn); bulk_commit(); }
        {
          // column sums: this thread owns channels (2 cp, 2 cp + 1) over 32 of the tile's 128 pixel rows
          const int tg = (int)threadIdx.x - 64 - group * EPI_THREADS;
          const uint32_t cp = (uint32_t)(tg & 31), row0 = (uint32_t)(tg >> 5) * 32u;
#pragma unroll 4
          for (uint32_t rr = 0; rr < 32; ++rr) {
            const uint32_t row = row0 + rr;
            const uint32_t off = row * 128u + ((((cp >> 2) ^ (row & 7u))) << 4) + ((cp & 3u) << 2);
            const uint32_t wa = ld_shared_u32(slotA + off), wb = ld_shared_u32(slotB + off), wc = ld_shared_u32(slotC + off);
            lnb_acc[0] += bf16_lo(wa); lnb_acc[1] += bf16_hi(wa);
            lnb_acc[2] += bf16_lo(wb); lnb_acc[3] += bf16_hi(wb);
            lnb_acc[4] += bf16_lo(wc); lnb_acc[5] += bf16_hi(wc);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) zr[c] = zn[c];        // rotate the pipeline: the next own tile's operands become current
        ln_mu = ln_mu_n; ln_rs = ln_rs_n;
      } else if (EPI != 3 && EPI != 4 && p.ln) {
        const long long pix = ((long long)img * p.H + oh) * p.W + ow;
        const float* gam = s_bias + p.Cout;
        const float* bet = s_bias + 2 * p.Cout;
        auto run = [&](auto tag) {
          constexpr int BN_T = decltype(tag)::value;
          uint32_t zp[BN_T / 2];
          float mean, rstd;
          ln_load<BN_T>(taddr, s_bias, zp, mean, rstd, p.ln_eps);
          release_tmem();
          if (valid) { p.mean[pix] = mean; p.rstd[pix] = rstd; }
          const bool relu = p.ln_relu != 0;
          const uint32_t rbase = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
#pragma unroll
          for (int hf = 0; hf < BN_T / 64; ++hf) {
            if (p.z) {
              const uint32_t slot = ring.begin() + rbase;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                st_shared_v4(slot + (((uint32_t)c ^ sw) << 4), zp[hf * 32 + c * 4], zp[hf * 32 + c * 4 + 1],
                             zp[hf * 32 + c * 4 + 2], zp[hf * 32 + c * 4 + 3]);
              ring.end(slot - rbase, 1, hf * 64, cw, ch, cn);
            }
            const uint32_t slot = ring.begin() + rbase;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int ci = hf * 64 + c * 8 + k * 2;
                const uint32_t w = zp[ci / 2];
                float t0 = (bf16_lo(w) - mean) * rstd * gam[ci] + bet[ci];
                float t1 = (bf16_hi(w) - mean) * rstd * gam[ci + 1] + bet[ci + 1];
                if (relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); }
                o[k] = pack_bf16(t0, t1);
              }
              st_shared_v4(slot + (((uint32_t)c ^ sw) << 4), o[0], o[1], o[2], o[3]);
            }
            ring.end(slot - rbase, 0, hf * 64, cw, ch, cn);
          }
        };
        if (EPI == 1) run(std::integral_constant<int, 64>{});
        else if (EPI == 2) run(std::integral_constant<int, 128>{});
        else if (p.BN == 64) run(std::integral_constant<int, 64>{});
        else run(std::integral_constant<int, 128>{});
      } else if ((EPI == 0 || EPI == 3) && p.tma_store) {
        const int halves = p.BN / 64;
        for (int hf = 0; hf < halves; ++hf) {
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr + hf * 64, v0);
          tmem_ld32(taddr + hf * 64 + 32, v1);
          tmem_ld_wait();
          if (hf == halves - 1) release_tmem();
          const bool relu = p.act == B200_ACT_RELU;
          const float* bh = bias + hf * 64;
          const uint32_t slot = ring.begin();
          stage_row(slot, r, [&](int i) {
            const float t = __uint_as_float(i < 32 ? v0[i] : v1[i - 32]) + bh[i];
            return relu ? fmaxf(t, 0.f) : t;
          });
          ring.end(slot, 0, j * p.BN + hf * 64, cw, ch, cn);
        }
      } else if (EPI == 0 || EPI == 3) {
        // per-thread read-modify-write stores (gradient accumulation into an existing tensor)
        __nv_bfloat16* dst = p.y + (long long)img * p.ysn + (long long)oh * p.ysh + (long long)ow * p.ysw + j * p.BN;
        const int cmax = p.Cout - j * p.BN;     // 32-channel outputs fill half of a 64-column tile
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (valid && c0 < cmax) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[g * 8 + i]) + bias[c0 + g * 8 + i];
              if (p.act == B200_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
              }
              __nv_bfloat16* d8 = dst + c0 + g * 8;
              if (p.accumulate) {
                float e[8];
                Vec8<__nv_bfloat16>::load(d8, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += e[i];
              }
              Vec8<__nv_bfloat16>::store(d8, o);
            }
          }
        }
        release_tmem();
      }
    }
    if (EPI == 4 && group < egroups) {
      // per-group reduction of the column sums over the four 32-row quarters, then one atomic per (quantity, channel)
      if (ring.issuer) bulk_wait_all();
      __syncwarp();
      named_bar_sync(ring.bar_id, EPI_THREADS);
      float* red = reinterpret_cast<float*>(smem_raw + (ring.stg0 - smem_u32(smem_raw)));     // [4 quarters][3][64]
      const int tg = (int)threadIdx.x - 64 - group * EPI_THREADS;
      const int cp = tg & 31, qt = tg >> 5;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        red[(qt * 3 + k) * 64 + 2 * cp] = lnb_acc[2 * k];
        red[(qt * 3 + k) * 64 + 2 * cp + 1] = lnb_acc[2 * k + 1];
      }
      named_bar_sync(ring.bar_id, EPI_THREADS);
      for (int i = tg; i < 192; i += EPI_THREADS) {
        const int k = i / 64, c = i % 64;
        const float sum = red[(0 * 3 + k) * 64 + c] + red[(1 * 3 + k) * 64 + c] + red[(2 * 3 + k) * 64 + c] + red[(3 * 3 + k) * 64 + c];
        float* dst = k == 0 ? p.dbias : (k == 1 ? p.dbeta : p.dgamma);
        if (dst) atomicAdd(dst + c, sum);
      }
    } else if (p.tma_store && group < egroups) ring.drain();
  }

  tc_fence_before();
  __syncthreads();
  if (CG2) cluster_sync_all();   // the peer's MMAs read this CTA's shared memory and its epilogue arrives on our barriers
  if (warp == 1) { tc_fence_after(); if (CG2) tmem_dealloc_cg2(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------
// Descriptor probe: D[128 x 64] = A[rows start.., K=64] * B[64 x 64]^T with a caller-chosen
// start offset and stride-byte-offset, used by tests to pin the UMMA addressing model.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int a_rows,
                  int start_bytes, int sbo_bytes, int lbo_bytes, int mn_major, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem0, b_smem = smem0 + 65536;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar_load), 1); mbar_init(smem_u32(&bar_mma), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_smem), 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    // A: a_rows x 64 (rows of 128 B), loaded in boxes of 64 rows; B: 64 x 64
    mbar_arrive_expect_tx(smem_u32(&bar_load), (uint32_t)(a_rows * 128 + 64 * 128));
    for (int r0 = 0; r0 < a_rows; r0 += 64) tma_load_2d(a_smem + r0 * 128, &tm_a, smem_u32(&bar_load), 0, r0);
    tma_load_2d(b_smem, &tm_b, smem_u32(&bar_load), 0, 0);
    mbar_wait(smem_u32(&bar_load), 0);
    tc_fence_after();
    if (mn_major == 2) {
      // weight-stationary pairs: D0 = A[rows 0..127] * B, D1 = A[rows 128..255] * B; B filled once per K step
      const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bd = smem_desc_sw128(b_smem + ks * 32, 0, 1024);
        umma_ws_bf16<false>(tmem_base, smem_desc_sw128(a_smem + start_bytes + ks * 32, lbo_bytes, sbo_bytes), bd, idesc, ks > 0);
        umma_ws_bf16<true>(tmem_base + 64, smem_desc_sw128(a_smem + 128 * 128 + start_bytes + ks * 32, lbo_bytes, sbo_bytes), bd,
                           idesc, ks > 0);
      }
    } else if (!mn_major) {
      const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base, smem_desc_sw128(a_smem + start_bytes + ks * 32, lbo_bytes, sbo_bytes),
                  smem_desc_sw128(b_smem + ks * 32, 0, 1024), idesc, ks > 0);
    } else {
      // wgrad-style: both operands MN-major.  A rows are K (pixels): M = 2 atoms of 64 channels that
      // are `lbo_bytes` apart (the same window shifted); B = 64 x 64 tile, K rows.
      const uint32_t idesc = idesc_bf16(128, 64, 1, 1);
      for (int ks = 0; ks < 4; ++ks)   // K = 64 pixels = 4 steps of 16 rows = 2 groups of 8 rows
        umma_bf16(tmem_base, smem_desc_sw128(a_smem + start_bytes + ks * 2 * sbo_bytes, lbo_bytes, sbo_bytes),
                  smem_desc_sw128(b_smem + ks * 2048, 0, 1024), idesc, ks > 0);
    }
    umma_commit(smem_u32(&bar_mma));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < (mn_major == 2 ? 128 : 64); c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[((c0 / 64) * 128 + row) * 64 + (c0 % 64) + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ---------------------------------------------------------------------------
// MMA issue-rate probe: every CTA issues `iters` x 4 tcgen05.mma (M=128, N=n, K=16 each) on fixed
// shared-memory operands and reports the cycles one thread saw from first issue to completion.
// Used to find the real tensor-pipe ceiling of the SS-mode (both operands in smem) N=64/128/256 shapes.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int n, int iters, int a_stride_bytes, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x / 32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem0, b_smem = smem0 + 32768;
  for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[i] = 0x3c003c00u;  // bf16 ~0.0078
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar_mma), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16(128, n, 0, 0);
    const long long t0 = clock64();
    if (a_stride_bytes == 2048) {     // weight-stationary pairs: two A tiles per B tile (fill + lastuse)
      for (int it = 0; it < iters; it += 2) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = smem_desc_sw128(b_smem + ks * 32, 0, 1024);
          umma_ws_bf16<false>(tmem_base, smem_desc_sw128(a_smem + ks * 32, 0, 1024), bd, idesc, 1);
          umma_ws_bf16<true>(tmem_base + (uint32_t)n, smem_desc_sw128(a_smem + 16384 + ks * 32, 0, 1024), bd, idesc, 1);
        }
      }
    } else if (a_stride_bytes == 2049) {     // the same two-tile loop with ordinary MMAs (B read from shared memory twice)
      for (int it = 0; it < iters; it += 2) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = smem_desc_sw128(b_smem + ks * 32, 0, 1024);
          umma_bf16(tmem_base, smem_desc_sw128(a_smem + ks * 32, 0, 1024), bd, idesc, 1);
          umma_bf16(tmem_base + (uint32_t)n, smem_desc_sw128(a_smem + 16384 + ks * 32, 0, 1024), bd, idesc, 1);
        }
      }
    } else
    for (int it = 0; it < iters; ++it) {
      // a_stride 1280 = the convolution's access pattern: per tap a shifted window start and its own B tile
      const int tap = a_stride_bytes == 1280 ? it % 9 : 0;
      const uint32_t a_off = (uint32_t)((tap / 3) * 10 + tap % 3) * 128u;
      const uint32_t b_off = a_stride_bytes == 1280 && n == 64 ? (uint32_t)(tap % 4) * 8192u : 0u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base + (uint32_t)((it & 1) * n), smem_desc_sw128(a_smem + a_off + ks * 32, 0, a_stride_bytes),
                  smem_desc_sw128(b_smem + b_off + ks * 32, 0, 1024), idesc, 1);
    }
    umma_commit(smem_u32(&bar_mma));
    mbar_wait(smem_u32(&bar_mma), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

bool conv_tc_supported(const b200_tensor* x, int cin, int cout, const b200_tensor* y, int ks) {
  if (ks != 3 && ks != 1) return false;
  if (x->dtype != B200_BF16 || y->dtype != B200_BF16) return false;
  // 32-channel tensors ride the 64-channel tiles: the TMA boxes are 64 channels wide and everything past the
  // tensor's 32 is the out-of-bounds zero fill on loads and clipped on stores (half of the MMA work is padding)
  if ((cin % 64 != 0 && cin != 32) || (cout % 64 != 0 && cout != 32)) return false;
  if (cout > 64 && cout % 128 != 0) return false;
  auto ok = [](const b200_tensor* t) {
    return ((uintptr_t)t->data % 16 == 0) && (t->stride_w * 2) % 16 == 0 && (t->stride_h * 2) % 16 == 0 &&
           (t->stride_n * 2) % 16 == 0;
  };
  return ok(x) && ok(y);
}

// Images of 1x1 pixels only ever use the centre tap, so the batch can be laid out as the pixels of
// one image: [N,1,1,C] -> [1, N/8, 8, C] (or [1,1,N,C]); 128 images then share one MMA tile.
static bool flatten_1x1(const b200_tensor* t, b200_tensor* out) {
  if (t->h != 1 || t->w != 1) return false;
  *out = *t;
  out->n = 1;
  if (t->n % 8 == 0) { out->h = t->n / 8; out->w = 8; out->stride_w = t->stride_n; out->stride_h = 8 * t->stride_n; }
  else { out->h = 1; out->w = t->n; out->stride_w = t->stride_n; out->stride_h = t->stride_n * t->n; }
  out->stride_n = t->stride_n * t->n;
  return true;
}

bool conv_gemm_wanted(const b200_tensor* x, int cin, int cout, int ks);

// shapes the fused dgrad + LayerNorm-backward epilogue takes (mirrors conv_tc_launch: the CTA-pair kernel on plain
// 16x8 tiles, one 64-channel N tile, weights resident next to six staging slots: K = cout of the filter <= 128)
bool conv_tc_dgrad_lnbwd_supported(const b200_tensor* dy, int k_channels, int n_channels, const b200_tensor* dx, int ks) {
  static const int allow_cg2 = getenv("B200_CONV_CG2") ? atoi(getenv("B200_CONV_CG2")) : 1;
  if (!allow_cg2 || ks != 3 || n_channels != 64 || k_channels % 64 != 0 || k_channels > 128) return false;
  if (!conv_tc_supported(dy, k_channels, n_channels, dx, ks) || conv_gemm_wanted(dy, k_channels, n_channels, ks)) return false;
  if (dx->h == 8 && dx->n > 1) return false;                 // pair tiles (two 8-row images per tile)
  if (dx->h + 2 <= 9 && dx->n > 1) return false;             // stacked small images
  if (dx->h == 1 && dx->w == 1) return false;                // flattened 1x1 batches
  const long long tiles = (long long)dx->n * ((dx->h + TILE_H - 1) / TILE_H) * ((dx->w + TILE_W - 1) / TILE_W);
  return tiles >= 2;
}

// x: input activations (C = K total); y: output (C = cout).
//   b_mn = 1: wmat is the Keras HWIO kernel [9][cin][cout] (fprop, B read MN-major)
//   b_mn = 0: wmat is [9][cout][cin] K-major (dgrad passes the HWIO kernel with cin/cout swapped + tap_rev)
struct ConvLnArgs {
  const float* gamma; const float* beta; float eps; int relu;
  const b200_tensor* z;   // may be NULL
  float* mean; float* rstd;
};

bool conv_tc_ln_supported(int cout) { return cout == 64 || cout == 128; }

// dgrad with the LayerNorm(+ReLU) backward of the layer that PRODUCED the convolution's input fused into its epilogue
struct ConvLnBwdArgs {
  const b200_tensor* z;            // saved pre-norm activations of that layer (same shape as dx)
  const float* mean; const float* rstd; const float* gamma; const float* beta; int relu;
  float* dgamma; float* dbeta; float* dbias;   // accumulated into (atomics); any may be NULL
};

// small-spatial split-K path (conv_gemm.cu)
bool conv_gemm_wanted(const b200_tensor* x, int cin, int cout, int ks);
int conv_gemm_launch(const b200_tensor* x, const void* wmat, int cin, int cout, int tap_rev, int b_mn, const float* bias,
                     const b200_tensor* y, int act, int accumulate, int ks, void* ws, size_t ws_bytes, cudaStream_t st);

// wmat_k: optional K-major copy [tap][cout][cin] of an MN-major `wmat` (fprop: the b200_filter's ohwi pack); it lets
// Cout = 64 layers run as CTA pairs, whose 32-column weight halves have no 128-byte-swizzled MN-major form.
int conv_tc_launch(const b200_tensor* x_in, const void* wmat, int cin, int cout, int tap_rev, int b_mn, const float* bias,
                   const b200_tensor* y_in, int act, int accumulate, cudaStream_t st, const ConvLnArgs* ln, int ks,
                   void* ws, size_t ws_bytes, const void* wmat_k, int allow_pairs, const ConvLnBwdArgs* lnb) {
  B200_REQUIRE(conv_tc_supported(x_in, cin, cout, y_in, ks), B200_ERR_UNSUPPORTED,
               "conv3x3 tcgen05: unsupported shape cin=%d cout=%d (need bf16, multiples of 64, 16-byte aligned)", cin,
               cout);
  B200_REQUIRE(x_in->n == y_in->n && x_in->h == y_in->h && x_in->w == y_in->w && x_in->c == cin && y_in->c == cout,
               B200_ERR_BAD_ARG, "conv3x3 tcgen05: tensor shapes do not match the filter");
  // deep levels (images <= 4x4): weight-streaming split-K GEMM.  The choice depends on the shapes only; a caller that
  // passes too little scratch gets an error, never another algorithm.
  if (!ln && conv_gemm_wanted(x_in, cin, cout, ks))
    return conv_gemm_launch(x_in, wmat, cin, cout, tap_rev, b_mn, bias, y_in, act, accumulate, ks, ws, ws_bytes, st);
  b200_tensor xf, yf;
  const b200_tensor *x = x_in, *y = y_in;
  int live_mask = 0;
  for (int t = 0; t < 9; ++t) {
    const bool dead = (x_in->h == 1 && t / 3 != 1) || (x_in->w == 1 && t % 3 != 1) || (ks == 1 && t != 4);
    if (!dead) live_mask |= 1 << t;
  }
  b200_tensor zf;
  const b200_tensor* z = ln ? ln->z : nullptr;
  if (flatten_1x1(x_in, &xf) && flatten_1x1(y_in, &yf)) {   // live_mask stays centre-only
    x = &xf; y = &yf;
    if (z && flatten_1x1(z, &zf)) z = &zf;
  }
  ConvTcParams p;
  p.N = y->n; p.H = y->h; p.W = y->w; p.Cout = cout;
  p.KB = (cin + 63) / 64;
  p.BN = cout <= 64 ? 64 : 128;
  p.n_tiles = (cout + p.BN - 1) / p.BN;
  p.tiles_h = (p.H + TILE_H - 1) / TILE_H;
  p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
  // stack several small images in one tile: image b occupies window rows [b*(H+2), (b+1)*(H+2))
  p.nb = 1; p.srows = 1 << 20;
  int box_h = WIN_H, box_n = 1;
  if (p.H + 2 <= 9 && p.N > 1) {
    p.srows = p.H + 2;
    p.nb = WIN_H / p.srows;
    while ((p.nb - 1) * p.srows + p.H - 1 > TILE_H - 1) --p.nb;
    if (p.nb > p.N) p.nb = p.N;
    box_h = p.srows; box_n = p.nb;
  }
  p.win_bytes = WIN_W * box_h * box_n * 128;
  p.pair = 0; p.win_stage = WIN_STAGE;
  static const int allow_pair = getenv("B200_CONV_PAIR") ? atoi(getenv("B200_CONV_PAIR")) : 1;
  if (allow_pair && p.H == 8 && p.N > 1 && p.nb == 1) {
    // 8-row images fill half a 16x8 tile: put two images in one tile with interleaved rows
    p.pair = 1; p.nb = 2; p.srows = 1 << 20;
    p.win_bytes = WIN_W * 2 * (8 + 2) * 128;          // 25600
    p.win_stage = PAIR_WIN_STAGE;
    p.tiles_h = 1;
  }
  const int groups = (p.N + p.nb - 1) / p.nb;
  const int spatial_tiles = groups * p.tiles_h * p.tiles_w;
  p.total_items = spatial_tiles * p.n_tiles;
  // CTA pairs (cta_group::2): plain 16x8 tiles (no stacked / paired small images), TMA-store epilogues, 64-multiple
  // channel counts; MN-major weights need 64-column halves (BN = 128) or the K-major pack
  static const int allow_cg2 = getenv("B200_CONV_CG2") ? atoi(getenv("B200_CONV_CG2")) : 1;
  p.cg2 = 0;
  if (allow_cg2 && allow_pairs && !p.pair && p.nb == 1 && !accumulate && spatial_tiles >= 2 && cin % 64 == 0 && cout % 64 == 0 &&
      (!b_mn || p.BN == 128 || wmat_k)) {
    p.cg2 = 1;
    if (b_mn && p.BN == 64) { wmat = wmat_k; b_mn = 0; tap_rev = 0; }   // [tap][cout][cin]: K-major, natural tap order
    p.total_items = ((spatial_tiles + 1) / 2) * p.n_tiles;              // work items of a cluster = tile pairs
  }
  p.lnb = 0; p.dgamma = p.dbeta = p.dbias = nullptr;
  if (lnb) {
    B200_REQUIRE(p.cg2 && !ln && cout == 64 && p.n_tiles == 1 && ks == 3 && !accumulate && lnb->z && same_shape(lnb->z, y_in) &&
                     lnb->z->dtype == B200_BF16 && (uintptr_t)lnb->z->data % 16 == 0 && lnb->z->stride_w % 8 == 0 &&
                     lnb->z->stride_h % 8 == 0 && lnb->z->stride_n % 8 == 0,
                 B200_ERR_UNSUPPORTED, "conv3x3 dgrad + LayerNorm backward: needs the CTA-pair kernel, 64 channels, bf16, aligned z");
    p.lnb = 1;
  }
  p.tap_rev = tap_rev;
  { const char* dbg = getenv("B200_CONV_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  p.b_mn = b_mn;
  p.Kc = cin;
  p.ntaps = ks == 1 ? 1 : 9;
  p.tap0 = ks == 1 ? 4 : 0;
  p.live_mask = live_mask;
  const int wt_bytes = (p.cg2 ? p.BN / 2 : p.BN) * 128;
  p.wt_bytes = wt_bytes;
  // (>= 2 KB: a stacked-image store box may span 18 tile rows, i.e. read 2 KB past its 16-row slot; those rows are
  // clipped by the TMA, the bytes only have to exist)
  int bias_bytes = ((cout * 4 * (ln ? 3 : 1) + 1023) / 1024) * 1024;
  if (bias_bytes < 2048) bias_bytes = 2048;
  const int budget = SMEM_LIMIT - 1024 /*align slack*/ - 1024 /*static*/ - bias_bytes;
  const int all_w = p.ntaps * p.KB * wt_bytes;
  // output staging ring (TMA stores); gradient accumulation keeps the per-thread read-modify-write path
  p.tma_store = accumulate ? 0 : 1;
  B200_REQUIRE(!(ln && accumulate), B200_ERR_BAD_ARG, "conv3x3+LayerNorm tcgen05: accumulate is not supported");
  const int jobs_per_tile = (p.BN / 64) * ((ln && z) ? 2 : 1);
  auto plan_smem = [&](int nslots, int min_windows) -> bool {
    const int left = budget - nslots * SLOT_BYTES;
    p.nslots = nslots;
    p.resident = (p.n_tiles == 1 && p.ntaps * p.KB <= MAX_WSLOTS && left - all_w >= min_windows * p.win_stage) ? 1 : 0;
    if (p.resident) {
      p.nsb = p.ntaps * p.KB;
      p.nsw = (left - all_w) / p.win_stage;
    } else {
      p.nsb = p.cg2 ? 8 : 4;     // (pairs: half-size tiles, and a slot is free only after the round trip through the peer)
      p.nsw = (left - p.nsb * wt_bytes) / p.win_stage;
    }
    if (p.nsw > 8) p.nsw = 8;
    return p.nsw >= min_windows;
  };
  if (!p.tma_store) {
    plan_smem(0, 2);
  } else if (p.lnb) {
    // three fixed slots (dz, d, d*xhat) for each of the two epilogue groups
    B200_REQUIRE(plan_smem(6, 2) && p.resident, B200_ERR_UNSUPPORTED,
                 "conv3x3 dgrad + LayerNorm backward: weights do not fit next to the six staging slots");
  } else {
    // prefer resident weights; within that, as many slots as one tile's jobs while keeping 3 window stages
    static const int max_slots = getenv("B200_CONV_SLOTS") ? atoi(getenv("B200_CONV_SLOTS")) : 4;
    const int want = max_slots < 2 ? 2 : (max_slots > 4 ? 4 : max_slots);
    (void)jobs_per_tile;
    bool ok = false;
    for (int ns = want; ns >= 2 && !ok; --ns) ok = plan_smem(ns, 3) && p.resident;
    for (int ns = want; ns >= 2 && !ok; --ns) ok = plan_smem(ns, 2) && p.resident;
    if (!ok) ok = plan_smem(want, 3);
    if (!ok) plan_smem(2, 2);
  }
  B200_REQUIRE(p.nsw >= 2, B200_ERR_UNSUPPORTED, "conv3x3 tcgen05: shared memory budget too small");
  static const int want_groups = getenv("B200_CONV_EGROUPS") ? atoi(getenv("B200_CONV_EGROUPS")) : 2;
  p.egroups = (ln && p.BN == 64 && !p.pair && p.tma_store && p.nslots >= 4 && want_groups == 2) ? 2 : 1;
  if (p.lnb) p.egroups = 2;
  p.act = act; p.accumulate = accumulate; p.bias = bias;
  p.y = reinterpret_cast<__nv_bfloat16*>(y->data);
  p.ysn = y->stride_n; p.ysh = y->stride_h; p.ysw = y->stride_w;
  p.ln = ln ? 1 : 0;
  p.z = nullptr; p.zsn = p.zsh = p.zsw = 0;
  if (lnb) {
    p.ln_relu = lnb->relu; p.ln_eps = 0.f; p.gamma = lnb->gamma; p.beta = lnb->beta;
    p.mean = const_cast<float*>(lnb->mean); p.rstd = const_cast<float*>(lnb->rstd);
    p.z = reinterpret_cast<__nv_bfloat16*>(lnb->z->data);
    p.zsn = lnb->z->stride_n; p.zsh = lnb->z->stride_h; p.zsw = lnb->z->stride_w;
    p.dgamma = lnb->dgamma; p.dbeta = lnb->dbeta; p.dbias = lnb->dbias;
    z = nullptr;
  }
  if (ln) {
    B200_REQUIRE(p.n_tiles == 1 && conv_tc_ln_supported(cout), B200_ERR_UNSUPPORTED,
                 "conv3x3+LayerNorm tcgen05: Cout=%d needs a single N tile (64 or 128)", cout);
    p.ln_relu = ln->relu; p.ln_eps = ln->eps; p.gamma = ln->gamma; p.beta = ln->beta; p.mean = ln->mean; p.rstd = ln->rstd;
    if (z) {
      p.z = reinterpret_cast<__nv_bfloat16*>(z->data);
      p.zsn = z->stride_n; p.zsh = z->stride_h; p.zsw = z->stride_w;
    }
  }

  CUtensorMap tm_x, tm_b, tm_y, tm_z;
  int rc = p.pair ? make_act_tmap_nh(&tm_x, x, WIN_W, 2, 8 + 2) : make_act_tmap(&tm_x, x, WIN_W, box_h, box_n);
  if (rc) return rc;
  rc = b_mn ? make_mat_tmap(&tm_b, wmat, (long long)p.ntaps * cin, cout, 64)
            : make_mat_tmap(&tm_b, wmat, (long long)p.ntaps * cout, cin, p.cg2 ? p.BN / 2 : p.BN);
  if (rc) return rc;
  // output boxes: {64 ch, 8, 16, 1}, or {64 ch, 8, H+2, nb} for stacked small images (rows past H are clipped)
  const int obox_h = p.nb > 1 || box_h != WIN_H ? box_h : TILE_H;
  rc = p.pair ? make_act_tmap_nh(&tm_y, y, TILE_W, 2, 8) : make_act_tmap(&tm_y, y, TILE_W, obox_h, box_n);
  if (rc) return rc;
  rc = p.pair ? make_act_tmap_nh(&tm_z, z ? z : y, TILE_W, 2, 8) : make_act_tmap(&tm_z, z ? z : y, TILE_W, obox_h, box_n);
  if (rc) return rc;

  const size_t smem = 1024 + (size_t)p.nsw * p.win_stage + (size_t)p.nsb * wt_bytes + (size_t)p.nslots * SLOT_BYTES + bias_bytes;
  if (first_use_on_device(3)) {
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    cudaFuncSetAttribute(conv3x3_tc_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
  }
  if (first_use_on_device(4)) {
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 1024);
  }
  if (p.cg2) {
    // one cluster of two CTAs (the two SMs of a TPC) per tile pair; persistent over 74 clusters
    int clusters = p.total_items < sm_count() / 2 ? p.total_items : sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3((p.ln && p.BN == 64) || p.lnb ? NTHREADS2 : NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled(2ull * clusters) ? 2 : 1;
    cudaError_t e;
    if (p.lnb) e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<false, 4, true>, tm_x, tm_b, tm_y, tm_z, p);
    else if (p.ln && p.BN == 64) e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<false, 1, true>, tm_x, tm_b, tm_y, tm_z, p);
    else if (p.ln) e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<false, 2, true>, tm_x, tm_b, tm_y, tm_z, p);
    else e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<false, 3, true>, tm_x, tm_b, tm_y, tm_z, p);
    if (e != cudaSuccess) { check_launch("conv3x3_tc_kernel (CTA pairs)"); return fail(B200_ERR_LAUNCH, "conv3x3_tc_kernel (CTA pairs): %s", cudaGetErrorString(e)); }
    return check_launch("conv3x3_tc_kernel");
  }
  int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  if (p.pair) launch_pdl(conv3x3_tc_kernel<true, 0>, grid, NTHREADS, smem, st, tm_x, tm_b, tm_y, tm_z, p);
  else if (p.ln && p.BN == 64) launch_pdl(conv3x3_tc_kernel<false, 1>, grid, NTHREADS2, smem, st, tm_x, tm_b, tm_y, tm_z, p);
  else if (p.ln) launch_pdl(conv3x3_tc_kernel<false, 2>, grid, NTHREADS, smem, st, tm_x, tm_b, tm_y, tm_z, p);
  else launch_pdl(conv3x3_tc_kernel<false, 3>, grid, NTHREADS, smem, st, tm_x, tm_b, tm_y, tm_z, p);
  return check_launch("conv3x3_tc_kernel");
}

// a: [a_rows][64] bf16 (row-major), b: [64][64] bf16, out: [128][64] fp32
int umma_probe(const void* a, int a_rows, const void* b, int start_bytes, int sbo_bytes, int lbo_bytes, int mn_major,
               float* out, cudaStream_t st) {
  B200_REQUIRE(a_rows % 64 == 0 && a_rows <= 512, B200_ERR_BAD_ARG, "umma_probe: a_rows must be a multiple of 64 <= 512");
  CUtensorMap tm_a, tm_b;
  int rc = make_mat_tmap(&tm_a, a, a_rows, 64, 64);
  if (rc) return rc;
  rc = make_mat_tmap(&tm_b, b, 64, 64, 64);
  if (rc) return rc;
  const size_t smem = 1024 + 65536 + 8192;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_probe_kernel<<<1, 128, smem, st>>>(tm_a, tm_b, a_rows, start_bytes, sbo_bytes, lbo_bytes, mn_major, out);
  return check_launch("umma_probe_kernel");
}

}  // namespace b200

namespace b200 {
int umma_rate(int n, int iters, int a_stride_bytes, long long* cycles, int grid, cudaStream_t st) {
  const size_t smem = 1024 + 65536;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_rate_kernel<<<grid, 128, smem, st>>>(n, iters, a_stride_bytes, cycles);
  return check_launch("umma_rate_kernel");
}
}  // namespace b200
