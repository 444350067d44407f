// Device-side infrastructure: buffers, launch accounting, deterministic block reductions.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>

#include "fb_internal.h"

// NVTX ranges around the phases of a step (visible in Nsight Systems / ncu --nvtx; no cost without a tool attached).
// Header-only NVTX3 from the CUDA toolkit; the library does not link libnvToolsExt.
#include <nvtx3/nvToolsExt.h>
struct fb_nvtx_range {
  explicit fb_nvtx_range(const char *name) { nvtxRangePushA(name); }
  ~fb_nvtx_range() { nvtxRangePop(); }
  fb_nvtx_range(const fb_nvtx_range &) = delete;
  fb_nvtx_range &operator=(const fb_nvtx_range &) = delete;
};
#define FB_NVTX_CAT2(a, b) a##b
#define FB_NVTX_CAT(a, b) FB_NVTX_CAT2(a, b)
#define FB_NVTX(name) fb_nvtx_range FB_NVTX_CAT(_fb_nvtx_, __LINE__)(name)

struct fb_cuda_error : std::runtime_error {
  int status;
  fb_cuda_error(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};

#define FB_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw fb_cuda_error(_e == cudaErrorMemoryAllocation ? FB_ENOMEM : FB_ECUDA,                        \
                          std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                              std::to_string(__LINE__));                                                \
  } while (0)

// All kernels go through this so the context can report how many of OUR kernels ran.  Every
// kernel in the library is a grid-stride loop, so the requested grid is clamped to one full
// wave (resident blocks per SM x SM count, from the occupancy calculator): a grid larger than
// that only adds a partially filled tail wave.
int fb_clamp_grid(fb_ctx *ctx, const void *kernel, int grid, int block, size_t smem);

#define FB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                               \
  do {                                                                                               \
    const int _fb_g = fb_clamp_grid((ctx), (const void *)(kernel), (grid), (block), (smem));          \
    kernel<<<_fb_g, (block), (smem), (ctx)->dev->stream>>>(__VA_ARGS__);                             \
    (ctx)->launches++;                                                                               \
  } while (0)

constexpr int FB_NSLOTS = 48;          // reduction result slots
constexpr int FB_MAX_RED_BLOCKS = 2048;  // max blocks of a reducing kernel

struct fb_device_state {
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  double *red = nullptr;          // [FB_NSLOTS] reduction results
  double *partials = nullptr;     // [FB_NSLOTS][FB_MAX_RED_BLOCKS]
  unsigned int *counter = nullptr;  // arrival counter of the reducing kernel in flight
  int *flag = nullptr;            // Krylov status word: 0 running, 1 converged, 2 breakdown
  int *iters = nullptr;           // iteration at which flag was raised
  double *host_pinned = nullptr;  // small pinned staging area (64 doubles)
  cudaEvent_t ev[12];  // 0-5: fb_ns_step phases, 6-7: fb_ctx_timer, 8-11: local stopwatches
};

// device blocks of DBuf come from a per-(device, size) cache of released blocks (fb_kernels.cu): no cudaFree /
// cudaMalloc pair when an object of the same shape is rebuilt
void *fb_block_alloc(size_t bytes);
void fb_block_free(void *p, size_t bytes);

template <typename T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) fb_block_free(p, n * sizeof(T));
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    if (count == n && p) return;
    release();
    if (count == 0) return;
    p = static_cast<T *>(fb_block_alloc(count * sizeof(T)));
    n = count;
  }
  void upload(const T *src, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) FB_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) {
    if (n) FB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

#if defined(__CUDACC__)
__device__ __forceinline__ double fb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction of NR values per thread into red[slot0 .. slot0+NR).
// Every block writes its partial; the last block to arrive sums the partials in a fixed
// order, stores the results and resets the counter.  Returns true in the finishing
// block (all its threads), after the results are visible to it.
template <int NR>
__device__ __forceinline__ bool fb_grid_reduce(double (&v)[NR], double *partials, unsigned int *counter, double *red,
                                               int slot0) {
  __shared__ double s_part[NR][32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const double w = fb_warp_sum(v[r]);
    if (lane == 0) s_part[r][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double w = lane < nwarps ? s_part[r][lane] : 0.0;
      w = fb_warp_sum(w);
      if (lane == 0) partials[(size_t)(slot0 + r) * FB_MAX_RED_BLOCKS + blockIdx.x] = w;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int arrived = atomicAdd(counter, 1u);
    s_last = (arrived == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double acc = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
      acc += __ldcg(&partials[(size_t)(slot0 + r) * FB_MAX_RED_BLOCKS + b]);
    const double w = fb_warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_part[r][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double w = lane < nwarps ? s_part[r][lane] : 0.0;
      w = fb_warp_sum(w);
      if (lane == 0) red[slot0 + r] = w;
      v[r] = w;
    }
  }
  if (threadIdx.x == 0) *counter = 0u;
  __syncthreads();
  return true;
}
#endif
