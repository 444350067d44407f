// Multi-GPU plumbing: one process per GPU, NCCL over NVLink for (1) the ghost-dof halo exchange
// in front of every SpMV / assembly and (2) the all-reduce of the Krylov dot products.
// Replaces PETSc's VecScatter + MPI_Allreduce [EXT] (the reference tree has no MPI-aware code;
// under mpirun DOLFIN/PETSc would partition implicitly).
//
// NCCL is dlopen'ed (libnccl.so.2) so that the library loads on machines without it and binds to
// whichever NCCL the process already has (torch bundles its own).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "fb_ops.h"

struct fb_comm {
  void *lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static bool load_nccl(fb_comm &c) {
  if (c.lib) return true;
  c.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!c.lib) c.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!c.lib) return false;
#define FB_SYM(name)                                                   \
  c.name = reinterpret_cast<decltype(c.name)>(dlsym(c.lib, "nccl" #name)); \
  if (!c.name) return false;
  FB_SYM(GetUniqueId)
  FB_SYM(CommInitRank)
  FB_SYM(CommDestroy)
  FB_SYM(AllReduce)
  FB_SYM(Send)
  FB_SYM(Recv)
  FB_SYM(GroupStart)
  FB_SYM(GroupEnd)
  FB_SYM(GetErrorString)
#undef FB_SYM
  return true;
}

#define FB_NCCL(c, expr)                                                                                     \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) throw fb_cuda_error(FB_ENCCL, std::string(#expr) + ": " + (c)->GetErrorString(_r)); \
  } while (0)

extern "C" {

int fb_comm_unique_id(void *id128) {
  if (!id128) return FB_EINVAL;
  fb_comm c;
  if (!load_nccl(c)) return FB_ENCCL;
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
  if (!load_nccl(*c)) {
    delete c;
    return fb_fail(ctx, FB_ENCCL, std::string("fb_comm_init: cannot load NCCL: ") + (dlerror() ? dlerror() : "missing symbol"));
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
  if (ctx->dev) cudaStreamSynchronize(ctx->dev->stream);
  ctx->comm->CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
  return FB_OK;
}

}  // extern "C"

bool fb_is_distributed(const fb_ctx *ctx) { return ctx->comm != nullptr && ctx->comm->nranks > 1; }

// sum red[slot0 .. slot0+count) over all ranks (in place, on the library stream)
void fb_allreduce_slots(fb_ctx *ctx, int slot0, int count) {
  if (!fb_is_distributed(ctx)) return;
  fb_comm *c = ctx->comm;
  double *p = ctx->dev->red + slot0;
  FB_NCCL(c, c->AllReduce(p, p, (size_t)count, ncclDouble, ncclSum, c->comm, ctx->dev->stream));
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
  fb_comm *c = ctx->comm;
  cudaStream_t st = ctx->dev->stream;
  const int nneigh = (int)sp.halo_ranks.size();
  const int64_t nsend = sp.halo_send_ptr[nneigh];
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
