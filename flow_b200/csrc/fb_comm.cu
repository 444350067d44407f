// Multi-GPU plumbing: one process per GPU; (1) the ghost-dof halo exchange in front of every SpMV /
// assembly and (2) the all-reduce of the Krylov dot products.
//
// Hot path: our own kernels over NVLink peer memory.  Every rank owns a "window" (flags + staging area),
// IPC-mapped into all other ranks of the node (fb_comm_window_create / _open).
//   halo exchange = k_halo_push (gathers the owned interface values and STORES them straight into the
//                   neighbours' windows over NVLink, then publishes an epoch flag)
//                 + k_halo_pull (waits for the neighbours' flags, copies the staged values into the ghost
//                   segment of x, acknowledges).
//   all-reduce    = k_peer_allreduce: one block writes its <= 32 partial sums into every window, waits for
//                   all contributions and sums them in rank order (bit-identical result on all ranks).
// Both cost a few microseconds instead of an NCCL launch each (the Krylov iterations of a partitioned
// run are latency bound: 560 pressure iterations of ~50 us each at 8 GPUs).  NCCL (grouped send/recv,
// ncclAllReduce) remains the set-up / fallback transport: no window (FB_NO_P2P=1, IPC refused) or a halo
// larger than the staging area.
// Replaces PETSc's VecScatter + MPI_Allreduce [EXT] (the reference tree has no MPI-aware code;
// under mpirun DOLFIN/PETSc would partition implicitly).
//
// NCCL is dlopen'ed (libnccl.so.2) so that the library loads on machines without it and binds to
// whichever NCCL the process already has (torch bundles its own).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "fb_ops.h"

constexpr int FB_MAX_PEERS = 8;  // one NVSwitch domain
constexpr size_t FB_WIN_HEADER = 8192;
constexpr int FB_AR_MAX = 32;  // doubles per peer all-reduce (FGMRES: up to m + 2 inner products at once)

// Header of every rank's window.  All words are written by remote GPUs with system-scope stores and read
// locally with volatile loads (the local L2 is the point of coherence for this memory).
struct PeerHeader {
  unsigned long long halo_flag[FB_MAX_PEERS];  // [sender]  : epoch of the sender's last completed push into this window
  unsigned long long halo_ack[FB_MAX_PEERS];   // [receiver]: epoch of my last push that this receiver has consumed
  unsigned long long ar_flag[2][FB_MAX_PEERS];  // [parity][rank]
  double ar_data[2][FB_MAX_PEERS][FB_AR_MAX];
  unsigned long long gather_flag[FB_MAX_PEERS];  // [sender]: epoch of the sender's last completed fb_peer_vec_gather
};
static_assert(sizeof(PeerHeader) <= FB_WIN_HEADER, "window header too large");

struct fb_comm {
  void *lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-memory windows
  bool p2p = false;
  char *win[FB_MAX_PEERS] = {nullptr};  // win[rank] is the local allocation, the others are IPC mappings
  size_t stage_per_peer = 0;            // bytes of staging reserved for each sender in every window
  unsigned long long halo_epoch[FB_MAX_PEERS] = {0};  // exchanges done with each peer (same count on both sides)
  unsigned long long ar_epoch = 0;
  unsigned long long gather_epoch = 0;
  unsigned int *arrive = nullptr;       // grid arrival counters of the push / pull kernels
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

// Loads libnccl and resolves the entry points; on failure `why` holds dlerror()'s text (read ONCE: dlerror() clears its
// state) and the handle is closed again.
static bool load_nccl(fb_comm &c, std::string &why) {
  if (c.lib) return true;
  c.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!c.lib) c.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!c.lib) {
    const char *e = dlerror();
    why = e ? e : "libnccl.so.2 not found";
    return false;
  }
#define FB_SYM(name)                                                        \
  c.name = reinterpret_cast<decltype(c.name)>(dlsym(c.lib, "nccl" #name)); \
  if (!c.name) {                                                            \
    const char *e = dlerror();                                              \
    why = e ? e : "missing symbol nccl" #name;                              \
    dlclose(c.lib);                                                         \
    c.lib = nullptr;                                                        \
    return false;                                                           \
  }
  FB_SYM(GetUniqueId)
  FB_SYM(CommInitRank)
  FB_SYM(CommDestroy)
  FB_SYM(AllReduce)
  FB_SYM(AllGather)
  FB_SYM(Broadcast)
  FB_SYM(Send)
  FB_SYM(Recv)
  FB_SYM(GroupStart)
  FB_SYM(GroupEnd)
  FB_SYM(GetErrorString)
#undef FB_SYM
  return true;
}

#define FB_API_BEGIN(ctxexpr) \
  fb_ctx *_ctx = (ctxexpr);   \
  try {
#define FB_API_END                                \
  }                                               \
  catch (const fb_cuda_error &e) {                \
    return fb_fail(_ctx, e.status, e.what());     \
  }                                               \
  catch (const std::exception &e) {               \
    return fb_fail(_ctx, FB_ECUDA, e.what());     \
  }                                               \
  return FB_OK;

#define FB_NCCL(c, expr)                                                                                     \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) throw fb_cuda_error(FB_ENCCL, std::string(#expr) + ": " + (c)->GetErrorString(_r)); \
  } while (0)

extern "C" {

int fb_comm_unique_id(void *id128) {
  if (!id128) return FB_EINVAL;
  fb_comm c;
  std::string why;
  if (!load_nccl(c, why)) return FB_ENCCL;
  ncclUniqueId id;
  if (c.GetUniqueId(&id) != ncclSuccess) return FB_ENCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "unexpected ncclUniqueId size");
  std::memcpy(id128, &id, sizeof(id));
  return FB_OK;
}

int fb_comm_init(fb_ctx *ctx, int rank, int nranks, const void *id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return FB_EINVAL;
  if (!ctx->dev) return fb_fail(ctx, FB_ENODEVICE, "fb_comm_init: host-only context");
  if (ctx->comm) return fb_fail(ctx, FB_EINVAL, "fb_comm_init: communicator already initialised");
  fb_comm *c = new fb_comm();
  std::string why;
  if (!load_nccl(*c, why)) {
    delete c;
    return fb_fail(ctx, FB_ENCCL, "fb_comm_init: cannot load NCCL: " + why);
  }
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  cudaSetDevice(ctx->device);
  ncclResult_t r = c->CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    std::string msg = std::string("ncclCommInitRank: ") + c->GetErrorString(r);
    delete c;
    return fb_fail(ctx, FB_ENCCL, msg);
  }
  c->rank = rank;
  c->nranks = nranks;
  ctx->comm = c;
  return FB_OK;
}

int fb_comm_destroy(fb_ctx *ctx) {
  if (!ctx || !ctx->comm) return FB_OK;
  fb_comm *c = ctx->comm;
  if (ctx->dev) cudaStreamSynchronize(ctx->dev->stream);
  for (int r = 0; r < c->nranks && r < FB_MAX_PEERS; ++r)
    if (c->win[r] && r != c->rank) cudaIpcCloseMemHandle(c->win[r]);
  if (c->win[c->rank]) cudaFree(c->win[c->rank]);
  if (c->arrive) cudaFree(c->arrive);
  c->CommDestroy(c->comm);
  delete c;
  ctx->comm = nullptr;
  return FB_OK;
}

/* Allocate this rank's window (header + nranks staging areas of staging_bytes_per_peer) and return its
 * 64-byte CUDA IPC handle; the launcher all-gathers the handles and passes them to fb_comm_window_open. */
int fb_comm_window_create(fb_ctx *ctx, int64_t staging_bytes_per_peer, void *handle64) {
  if (!ctx || !ctx->comm || !handle64 || staging_bytes_per_peer < 0) return FB_EINVAL;
  fb_comm *c = ctx->comm;
  if (c->nranks > FB_MAX_PEERS) return fb_fail(ctx, FB_EINVAL, "fb_comm_window_create: more than 8 ranks");
  if (c->win[c->rank]) return fb_fail(ctx, FB_EINVAL, "fb_comm_window_create: window exists");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected cudaIpcMemHandle_t size");
  FB_API_BEGIN(ctx)
  cudaSetDevice(ctx->device);
  c->stage_per_peer = ((size_t)staging_bytes_per_peer + 255) / 256 * 256;
  const size_t bytes = FB_WIN_HEADER + c->stage_per_peer * c->nranks;
  void *p = nullptr;
  FB_CUDA(cudaMalloc(&p, bytes));
  FB_CUDA(cudaMemset(p, 0, bytes));
  FB_CUDA(cudaMalloc((void **)&c->arrive, 4 * sizeof(unsigned int)));
  FB_CUDA(cudaMemset(c->arrive, 0, 4 * sizeof(unsigned int)));
  FB_CUDA(cudaDeviceSynchronize());
  c->win[c->rank] = static_cast<char *>(p);
  cudaIpcMemHandle_t h;
  FB_CUDA(cudaIpcGetMemHandle(&h, p));
  std::memcpy(handle64, &h, sizeof(h));
  FB_API_END
}

/* handles: nranks x 64 bytes, rank-ordered.  Maps every peer's window; afterwards halo exchanges and the
 * dot-product all-reduces run over peer memory.  Collective: every rank must call it (or none). */
int fb_comm_window_open(fb_ctx *ctx, const void *handles) {
  if (!ctx || !ctx->comm || !handles) return FB_EINVAL;
  fb_comm *c = ctx->comm;
  if (!c->win[c->rank]) return fb_fail(ctx, FB_EINVAL, "fb_comm_window_open: create the local window first");
  FB_API_BEGIN(ctx)
  cudaSetDevice(ctx->device);
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char *>(handles) + 64 * r, sizeof(h));
    void *p = nullptr;
    FB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->win[r] = static_cast<char *>(p);
  }
  c->p2p = true;
  FB_API_END
}

/* Back to NCCL (e.g. when another rank failed to map the windows: all ranks must use the same transport). */
int fb_comm_window_disable(fb_ctx *ctx) {
  if (!ctx || !ctx->comm) return FB_EINVAL;
  if (ctx->dev) cudaStreamSynchronize(ctx->dev->stream);
  ctx->comm->p2p = false;
  return FB_OK;
}

/* 1 if the peer-memory transport is active on this context */
int fb_comm_uses_peer_memory(fb_ctx *ctx) { return ctx && ctx->comm && ctx->comm->p2p ? 1 : 0; }

}  // extern "C"

bool fb_is_distributed(const fb_ctx *ctx) { return ctx->comm != nullptr && ctx->comm->nranks > 1; }

// ---- peer-memory kernels ------------------------------------------------------------------------
namespace {

// clock ticks (~60 s; ranks can be seconds apart after host-side set-up): a lost peer ends in NaNs, not in a hang
constexpr long long FB_SPIN_LIMIT = 120000000000LL;

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const unsigned long long *p, unsigned long long epoch) {
  const long long t0 = clock64();
  while (ld_sys(p) < epoch)
    if (clock64() - t0 > FB_SPIN_LIMIT) return false;
  return true;
}

struct PeerPtrs {
  char *win[FB_MAX_PEERS];
};

// One block.  Thread r < nranks publishes this rank's `count` partial sums in window r, all threads wait for the
// nranks contributions to the local window, thread c sums component c in rank order.
__global__ void k_peer_allreduce(PeerPtrs pp, int rank, int nranks, unsigned long long epoch, double *red, int slot0,
                                 int count) {
  const int parity = (int)(epoch & 1ull);
  const int t = threadIdx.x;
  __shared__ int s_ok;
  if (t == 0) s_ok = 1;
  __syncthreads();
  if (t < nranks) {
    PeerHeader *h = reinterpret_cast<PeerHeader *>(pp.win[t]);
    for (int c = 0; c < count; ++c) h->ar_data[parity][rank][c] = red[slot0 + c];
    __threadfence_system();
    st_sys(&h->ar_flag[parity][rank], epoch);
  }
  PeerHeader *me = reinterpret_cast<PeerHeader *>(pp.win[rank]);
  if (t < nranks && !spin_until(&me->ar_flag[parity][t], epoch)) s_ok = 0;
  __syncthreads();
  if (t < count) {
    double acc = 0.0;
    for (int r = 0; r < nranks; ++r) acc += *reinterpret_cast<volatile double *>(&me->ar_data[parity][r][t]);
    red[slot0 + t] = s_ok ? acc : __longlong_as_double(0x7ff8000000000000LL);
  }
}

struct HaloArgs {
  int nneigh;
  int peer[FB_MAX_PEERS];            // neighbour ranks
  long long send_ptr[FB_MAX_PEERS + 1];  // node offsets into send_nodes
  long long recv_ptr[FB_MAX_PEERS + 1];  // node offsets into the ghost segment
  unsigned long long epoch[FB_MAX_PEERS];
};

// ONE kernel per halo exchange (all blocks co-resident: grid <= SM count):
//   1. wait until every neighbour has consumed my previous push (its ack in my window),
//   2. gather the owned interface values of x and STORE them into the neighbours' staging areas (region `rank`
//      of their windows) over NVLink; the last block to finish publishes the epoch flag in each neighbour's window,
//   3. wait for the neighbours' flags in my window and copy their staged values into the ghost segment of x,
//   4. the last block to finish acknowledges to every neighbour.
__global__ void k_halo_exchange(PeerPtrs pp, HaloArgs ha, int rank, size_t stage_per_peer, int ncomp, long long n_owned,
                                const int *__restrict__ nodes, double *__restrict__ x, unsigned int *arrive) {
  PeerHeader *me = reinterpret_cast<PeerHeader *>(pp.win[rank]);
  __shared__ int s_ok;
  __shared__ bool s_last;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < ha.nneigh && !spin_until(&me->halo_ack[ha.peer[threadIdx.x]], ha.epoch[threadIdx.x] - 1)) s_ok = 0;
  __syncthreads();
  const long long nsend = ha.send_ptr[ha.nneigh] * ncomp;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nsend; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / ncomp;
    const int c = (int)(t - k * ncomp);
    int q = 0;
    while (k >= ha.send_ptr[q + 1]) ++q;
    double *stage = reinterpret_cast<double *>(pp.win[ha.peer[q]] + FB_WIN_HEADER + stage_per_peer * rank);
    stage[(k - ha.send_ptr[q]) * ncomp + c] = s_ok ? x[(long long)nodes[k] * ncomp + c] : nan;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&arrive[0], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < ha.nneigh) {
      PeerHeader *h = reinterpret_cast<PeerHeader *>(pp.win[ha.peer[threadIdx.x]]);
      st_sys(&h->halo_flag[rank], ha.epoch[threadIdx.x]);
    }
    if (threadIdx.x == 0) arrive[0] = 0u;
  }
  // ---- receive side
  if (threadIdx.x < ha.nneigh && !spin_until(&me->halo_flag[ha.peer[threadIdx.x]], ha.epoch[threadIdx.x])) s_ok = 0;
  __syncthreads();
  const long long nrecv = ha.recv_ptr[ha.nneigh] * ncomp;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nrecv; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / ncomp;
    int q = 0;
    while (k >= ha.recv_ptr[q + 1]) ++q;
    const double *stage = reinterpret_cast<const double *>(pp.win[rank] + FB_WIN_HEADER + stage_per_peer * ha.peer[q]);
    const double v = __ldcv(&stage[t - ha.recv_ptr[q] * ncomp]);  // written by a remote GPU: never from L1
    x[n_owned * ncomp + t] = s_ok ? v : nan;
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&arrive[1], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < ha.nneigh) {
      PeerHeader *h = reinterpret_cast<PeerHeader *>(pp.win[ha.peer[threadIdx.x]]);
      st_sys(&h->halo_ack[rank], ha.epoch[threadIdx.x]);
    }
    if (threadIdx.x == 0) arrive[1] = 0u;
  }
}

// fb_peer_vec_gather: every rank scatters its owned entries (global index l2g[i]) into the current buffer of ALL
// ranks, publishes its epoch, and waits for everybody else's: afterwards the local buffer holds the whole vector.
// The all-gather and the permutation to global numbering are one kernel of remote stores (grid <= SM count).
__global__ void k_peer_gather(PeerPtrs bufs, PeerPtrs wins, int rank, int nranks, unsigned long long epoch, long long n_owned,
                              const int *__restrict__ l2g, const double *__restrict__ owned, unsigned int *arrive, int *ok_out) {
  __shared__ bool s_last;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_owned; i += (long long)gridDim.x * blockDim.x) {
    const double v = owned[i];
    const long long g = l2g[i];
    for (int q = 0; q < nranks; ++q) reinterpret_cast<double *>(bufs.win[q])[g] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&arrive[2], 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < nranks) st_sys(&reinterpret_cast<PeerHeader *>(wins.win[threadIdx.x])->gather_flag[rank], epoch);
    if (threadIdx.x == 0) arrive[2] = 0u;
  }
  PeerHeader *me = reinterpret_cast<PeerHeader *>(wins.win[rank]);
  if (threadIdx.x < nranks && !spin_until(&me->gather_flag[threadIdx.x], epoch)) *ok_out = 0;
}

PeerPtrs peer_ptrs(const fb_comm *c) {
  PeerPtrs pp;
  for (int r = 0; r < FB_MAX_PEERS; ++r) pp.win[r] = c->win[r];
  return pp;
}

}  // namespace

// sum red[slot0 .. slot0+count) over all ranks (in place, on the library stream)
void fb_allreduce_slots(fb_ctx *ctx, int slot0, int count) {
  if (!fb_is_distributed(ctx)) return;
  ctx->allreduce_calls++;
  fb_comm *c = ctx->comm;
  double *p = ctx->dev->red + slot0;
  if (c->p2p && count <= FB_AR_MAX) {
    ++c->ar_epoch;
    k_peer_allreduce<<<1, 32, 0, ctx->dev->stream>>>(peer_ptrs(c), c->rank, c->nranks, c->ar_epoch, ctx->dev->red, slot0, count);
    ctx->launches++;
    return;
  }
  FB_NCCL(c, c->AllReduce(p, p, (size_t)count, ncclDouble, ncclSum, c->comm, ctx->dev->stream));
}

// ---- IPC-shared, double-buffered global vectors (replicated coarse problems) -------------------------
struct fb_peer_vec {
  int64_t n = 0;
  char *buf[2][FB_MAX_PEERS] = {{nullptr}};  // [parity][rank]; [.][my rank] is the local allocation
  int *ok = nullptr;                          // device flag: 0 after a time-out
  int rank = 0, nranks = 1;
};

void fb_peer_vec_destroy(fb_peer_vec *v);

// collective; returns null if the peer-memory transport is not active
fb_peer_vec *fb_peer_vec_create(fb_ctx *ctx, int64_t n_global) {
  if (!fb_is_distributed(ctx) || !ctx->comm->p2p) return nullptr;
  fb_comm *c = ctx->comm;
  cudaStream_t st = ctx->dev->stream;
  // owned by a guard until every buffer is mapped: a CUDA / NCCL error below must not leak the allocations
  struct Guard {
    fb_peer_vec *v;
    ~Guard() {
      if (v) fb_peer_vec_destroy(v);
    }
  } guard{new fb_peer_vec()};
  fb_peer_vec *v = guard.v;
  v->n = n_global;
  v->rank = c->rank;
  v->nranks = c->nranks;
  FB_CUDA(cudaMalloc((void **)&v->ok, sizeof(int)));
  const int one = 1;
  FB_CUDA(cudaMemcpy(v->ok, &one, sizeof(int), cudaMemcpyHostToDevice));
  DBuf<char> dh;  // handles of all ranks, exchanged with NCCL (set-up only)
  dh.alloc((size_t)64 * c->nranks);
  std::vector<char> hh((size_t)64 * c->nranks);
  for (int par = 0; par < 2; ++par) {
    void *p = nullptr;
    FB_CUDA(cudaMalloc(&p, sizeof(double) * (size_t)n_global));
    FB_CUDA(cudaMemset(p, 0, sizeof(double) * (size_t)n_global));
    v->buf[par][c->rank] = static_cast<char *>(p);
    cudaIpcMemHandle_t h;
    FB_CUDA(cudaIpcGetMemHandle(&h, p));
    FB_CUDA(cudaMemcpyAsync(dh.p + 64 * c->rank, &h, 64, cudaMemcpyHostToDevice, st));
    FB_NCCL(c, c->AllGather(dh.p + 64 * c->rank, dh.p, 64, ncclChar, c->comm, st));
    FB_CUDA(cudaMemcpyAsync(hh.data(), dh.p, hh.size(), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < c->nranks; ++r) {
      if (r == c->rank) continue;
      std::memcpy(&h, hh.data() + 64 * r, 64);
      void *q = nullptr;
      FB_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      v->buf[par][r] = static_cast<char *>(q);
    }
  }
  guard.v = nullptr;
  return v;
}

void fb_peer_vec_destroy(fb_peer_vec *v) {
  if (!v) return;
  for (int par = 0; par < 2; ++par)
    for (int r = 0; r < v->nranks; ++r)
      if (v->buf[par][r]) {
        if (r == v->rank) cudaFree(v->buf[par][r]);
        else cudaIpcCloseMemHandle(v->buf[par][r]);
      }
  if (v->ok) cudaFree(v->ok);
  delete v;
}

// Assemble the global vector from the ranks' owned parts; returns the local copy (valid until the next-but-one call).
const double *fb_peer_vec_gather(fb_ctx *ctx, fb_peer_vec *v, const double *owned, const int *l2g, int64_t n_owned) {
  fb_comm *c = ctx->comm;
  const unsigned long long epoch = ++c->gather_epoch;
  const int par = (int)(epoch & 1ull);
  PeerPtrs bufs;
  for (int r = 0; r < FB_MAX_PEERS; ++r) bufs.win[r] = v->buf[par][r];
  const int g = (int)std::min<int64_t>(std::max<int64_t>((n_owned + 255) / 256, 1), ctx->dev->sm_count);
  k_peer_gather<<<g, 256, 0, ctx->dev->stream>>>(bufs, peer_ptrs(c), c->rank, c->nranks, epoch, (long long)n_owned, l2g, owned,
                                                 c->arrive, v->ok);
  ctx->launches++;
  return reinterpret_cast<const double *>(v->buf[par][c->rank]);
}

// device buffer of rank `root` copied to all ranks (set-up only; NCCL)
void fb_broadcast_device(fb_ctx *ctx, void *p, size_t bytes, int root) {
  if (!fb_is_distributed(ctx)) return;
  fb_comm *c = ctx->comm;
  FB_NCCL(c, c->Broadcast(p, p, bytes, ncclChar, root, c->comm, ctx->dev->stream));
}

// host value summed over all ranks (synchronous; set-up decisions that every rank must take alike)
double fb_allreduce_host_sum(fb_ctx *ctx, double v) {
  if (!fb_is_distributed(ctx)) return v;
  fb_device_state *dv = ctx->dev;
  const int slot = FB_NSLOTS - 1;
  FB_CUDA(cudaMemcpyAsync(dv->red + slot, &v, sizeof(double), cudaMemcpyHostToDevice, dv->stream));
  fb_allreduce_slots(ctx, slot, 1);
  FB_CUDA(cudaMemcpyAsync(&v, dv->red + slot, sizeof(double), cudaMemcpyDeviceToHost, dv->stream));
  FB_CUDA(cudaStreamSynchronize(dv->stream));
  return v;
}

__global__ void k_halo_pack(int64_t nsend, int ncomp, const int *__restrict__ nodes, const double *__restrict__ x,
                            double *__restrict__ buf) {
  const int64_t total = nsend * ncomp;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / ncomp;
    const int c = (int)(t - k * ncomp);
    buf[t] = x[(int64_t)nodes[k] * ncomp + c];
  }
}

// refresh x on the ghost nodes [n_owned, nnodes) from their owners
void halo_exchange(fb_ctx *ctx, DevSpace &sp, double *x, int ncomp) {
  if (!fb_is_distributed(ctx) || sp.halo_ranks.empty()) return;
  ctx->halo_calls++;
  fb_comm *c = ctx->comm;
  cudaStream_t st = ctx->dev->stream;
  const int nneigh = (int)sp.halo_ranks.size();
  const int64_t nsend = sp.halo_send_ptr[nneigh];
  if (c->p2p && nneigh <= FB_MAX_PEERS) {
    // largest message of this exchange in either direction must fit the per-sender staging area
    int64_t biggest = 0;
    for (int k = 0; k < nneigh; ++k) {
      biggest = std::max(biggest, sp.halo_send_ptr[k + 1] - sp.halo_send_ptr[k]);
      biggest = std::max(biggest, sp.halo_recv_ptr[k + 1] - sp.halo_recv_ptr[k]);
    }
    if ((size_t)biggest * ncomp * sizeof(double) <= c->stage_per_peer) {
      HaloArgs ha;
      ha.nneigh = nneigh;
      for (int k = 0; k < nneigh; ++k) {
        ha.peer[k] = sp.halo_ranks[k];
        ha.epoch[k] = ++c->halo_epoch[sp.halo_ranks[k]];
      }
      for (int k = 0; k <= nneigh; ++k) {
        ha.send_ptr[k] = sp.halo_send_ptr[k];
        ha.recv_ptr[k] = sp.halo_recv_ptr[k];
      }
      const PeerPtrs pp = peer_ptrs(c);
      const int64_t nrecv = sp.halo_recv_ptr[nneigh];
      const int64_t work = std::max(nsend, nrecv) * ncomp;
      const int g = (int)std::min<int64_t>(std::max<int64_t>((work + 255) / 256, 1), ctx->dev->sm_count);
      k_halo_exchange<<<g, 256, 0, st>>>(pp, ha, c->rank, c->stage_per_peer, ncomp, (long long)sp.n_owned,
                                         sp.halo_send_nodes.p, x, c->arrive);
      ctx->launches += 1;
      return;
    }
  }
  sp.halo_buf.alloc((size_t)nsend * 3);
  if (nsend > 0) {
    int64_t g = (nsend * ncomp + 255) / 256;
    if (g > 4096) g = 4096;
    FB_LAUNCH(ctx, k_halo_pack, (int)g, 256, 0, nsend, ncomp, sp.halo_send_nodes.p, x, sp.halo_buf.p);
  }
  FB_NCCL(c, c->GroupStart());
  for (int k = 0; k < nneigh; ++k) {
    const int64_t s0 = sp.halo_send_ptr[k], s1 = sp.halo_send_ptr[k + 1];
    const int64_t r0 = sp.halo_recv_ptr[k], r1 = sp.halo_recv_ptr[k + 1];
    if (s1 > s0)
      FB_NCCL(c, c->Send(sp.halo_buf.p + s0 * ncomp, (size_t)((s1 - s0) * ncomp), ncclDouble, sp.halo_ranks[k], c->comm, st));
    if (r1 > r0)
      FB_NCCL(c, c->Recv(x + (sp.n_owned + r0) * ncomp, (size_t)((r1 - r0) * ncomp), ncclDouble, sp.halo_ranks[k], c->comm, st));
  }
  FB_NCCL(c, c->GroupEnd());
}
