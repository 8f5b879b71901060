// Gradient / weight exchange of the data-parallel step over NVLink peer memory.
//
// The reference trains on one GPU; data parallelism is this framework's own design, so there is no reference
// line to cite beyond the train step it shards (Super_resolution/code/train_adaptive_unet.py:622-632).
//
// NCCL's reduce-scatter / all-gather kernels run on SMs next to the convolutions of the step they hide behind: with
// the exchange active for most of the step (C3: 0.5 GB of gradients per rank) they cost the step 6..8 %.  Here the
// bytes move on the COPY ENGINES and only the sum runs on SMs:
//
//   every rank's flat gradient buffer G, bf16 weight shadow S and a block of 64-bit counters live in cudaMalloc'ed
//   memory exported with cudaIpcGetMemHandle and mapped by every peer (one node, NVSwitch);
//   reduce-scatter of a bucket  = signal "my bucket is complete" (counter += 1), wait until every peer's counter has
//       caught up, pull my shard of the bucket out of each peer's G into a local staging buffer
//       (cudaMemcpyPeerAsync: copy engine, no SM), then ONE kernel adds the staged slices into my shard;
//   all-gather of the shadow    = signal "my Adam is done", wait for the peers, pull every peer's shard of S into mine.
//
// Write-after-read safety is the caller's protocol (keras/model.py::_run_step_p2p): a rank zeroes G for the next step
// only after every peer signalled its Adam (which is ordered behind that peer's pulls), and rewrites its shard of S
// only after the next step's gradient counters of every peer caught up (ordered behind that peer's shadow pulls).
#include "common.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// counter += 1, visible to the peers (system scope) after everything this stream did before
__global__ void peer_signal_kernel(unsigned long long* flag) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd_system(flag, 1ULL);
  }
}

// thread t spins until peer t's counter has reached mine; a peer that never arrives (crashed rank) trips the
// timeout and the kernel traps instead of hanging the box
__global__ void peer_wait_kernel(const unsigned long long* const* peers, int n, const unsigned long long* own,
                                 unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t >= n) return;
  const unsigned long long target = *reinterpret_cast<const volatile unsigned long long*>(own);
  const unsigned long long* p = peers[t];
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(p) < target) {
    __nanosleep(200);
    if (global_ns() - t0 > timeout_ns) {
      printf("b200 peer_wait: peer %d stuck at %llu, waiting for %llu\n", t, ld_acquire_sys(p), target);
      __trap();
    }
  }
}

// acc[i] += sum_j staged[j * count + i]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) peer_sum_kernel(float* __restrict__ acc, const float* __restrict__ staged, int n,
                                                       size_t count) {
  const size_t nv = count / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 a = reinterpret_cast<const float4*>(acc)[i];
    for (int j = 0; j < n; ++j) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(staged + (size_t)j * count) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(acc)[i] = a;
  }
}

int cuda_fail(const char* what, cudaError_t e) { return fail(B200_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e)); }

}  // namespace
}  // namespace b200

using namespace b200;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int b200_peer_alloc(void** out, size_t bytes) {
  B200_REQUIRE(out && bytes > 0, B200_ERR_BAD_ARG, "peer_alloc: bad argument");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail("peer_alloc: cudaMalloc", e);
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail("peer_alloc: cudaMemset", e); }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail("peer_alloc: cudaDeviceSynchronize", e); }
  *out = p;
  return B200_OK;
}

int b200_peer_free(void* p) {
  if (!p) return B200_OK;
  cudaError_t e = cudaFree(p);
  if (e != cudaSuccess) return cuda_fail("peer_free: cudaFree", e);
  return B200_OK;
}

int b200_peer_export(const void* p, void* handle64) {
  B200_REQUIRE(p && handle64, B200_ERR_BAD_ARG, "peer_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == B200_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(p));
  if (e != cudaSuccess) return cuda_fail("peer_export: cudaIpcGetMemHandle", e);
  memcpy(handle64, &h, sizeof(h));
  return B200_OK;
}

int b200_peer_open(const void* handle64, void** out) {
  B200_REQUIRE(handle64 && out, B200_ERR_BAD_ARG, "peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail("peer_open: cudaIpcOpenMemHandle", e);
  *out = p;
  return B200_OK;
}

int b200_peer_close(void* p) {
  if (!p) return B200_OK;
  cudaError_t e = cudaIpcCloseMemHandle(p);
  if (e != cudaSuccess) return cuda_fail("peer_close: cudaIpcCloseMemHandle", e);
  return B200_OK;
}

int b200_peer_signal(unsigned long long* counter, void* stream) {
  B200_REQUIRE(counter, B200_ERR_BAD_ARG, "peer_signal: NULL counter");
  peer_signal_kernel<<<1, 32, 0, ST(stream)>>>(counter);
  return check_launch("peer_signal_kernel");
}

int b200_peer_wait(const unsigned long long* const* peer_counters, int n_peers, const unsigned long long* own,
                   double timeout_s, void* stream) {
  B200_REQUIRE(peer_counters && own && n_peers >= 1 && n_peers <= 32, B200_ERR_BAD_ARG, "peer_wait: bad argument");
  const unsigned long long ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 30.0) * 1e9);
  peer_wait_kernel<<<1, 32, 0, ST(stream)>>>(peer_counters, n_peers, own, ns);
  return check_launch("peer_wait_kernel");
}

int b200_peer_pull(void* const* dst, const void* const* src, const int* src_device, const size_t* bytes, int n,
                   int my_device, void* stream) {
  B200_REQUIRE(dst && src && src_device && bytes && n >= 0, B200_ERR_BAD_ARG, "peer_pull: bad argument");
  for (int j = 0; j < n; ++j) {
    if (bytes[j] == 0) continue;
    cudaError_t e = cudaMemcpyPeerAsync(dst[j], my_device, src[j], src_device[j], bytes[j], ST(stream));
    if (e != cudaSuccess) return cuda_fail("peer_pull: cudaMemcpyPeerAsync", e);
  }
  return B200_OK;
}

int b200_peer_gather_sum(float* acc, float* staging, const void* const* peer_src, const int* peer_device, int n_peers,
                         size_t count, int my_device, void* stream) {
  B200_REQUIRE(acc && staging && peer_src && peer_device && n_peers >= 1, B200_ERR_BAD_ARG, "peer_gather_sum: bad argument");
  B200_REQUIRE(count % 4 == 0 && (uintptr_t)acc % 16 == 0 && (uintptr_t)staging % 16 == 0, B200_ERR_BAD_ARG,
               "peer_gather_sum: count must be a multiple of 4 and the buffers 16-byte aligned");
  if (count == 0) return B200_OK;
  for (int j = 0; j < n_peers; ++j) {
    cudaError_t e = cudaMemcpyPeerAsync(staging + (size_t)j * count, my_device, peer_src[j], peer_device[j],
                                        count * sizeof(float), ST(stream));
    if (e != cudaSuccess) return cuda_fail("peer_gather_sum: cudaMemcpyPeerAsync", e);
  }
  const size_t nv = count / 4;
  const long long want = (long long)((nv + 255) / 256);
  const int grid = (int)(want < 2LL * sm_count() ? (want > 0 ? want : 1) : 2LL * sm_count());
  peer_sum_kernel<<<grid, 256, 0, ST(stream)>>>(acc, staging, n_peers, count);
  return check_launch("peer_sum_kernel");
}

}  // extern "C"
