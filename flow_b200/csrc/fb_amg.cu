// Smoothed-aggregation AMG preconditioner for the pressure Poisson solve.
//
// Replaces hypre BoomerAMG as used by the reference (pressure_correction.py:331, :414-419) [EXT].
// The reference relaxes with Jacobi on the coarsest grid because an LU factorisation there breaks on
// the singular pure-Neumann operator (:399-418).  Here the coarsest operator (a few hundred rows) is
// inverted densely once on the host as a PSEUDO-inverse: A+ = (A + s w w^T)^-1 - w w^T / s with w the
// coarse image of the constant vector when the fine operator annihilates constants, the plain inverse
// otherwise (Dirichlet variant).  Damped-Jacobi smoothing on all other levels; the cycle is a fixed
// symmetric positive semi-definite operator, as CG requires.
//
// Set-up (host, once per matrix): strength-of-connection filter, greedy aggregation (Vanek et al.),
// tentative prolongator carrying the near-null-space vector from level to level, prolongator smoothing
// P = (I - w D^-1 A) P0 with w = 4 / (3 rho), Galerkin coarse operators R A P by row-wise sparse
// products.  Cycle (device): V(1,1); lanes per row are chosen per operator from its mean row length
// (the Galerkin operators and the restrictions of the deep levels have rows of 100+ entries).
// Multi-GPU: every rank builds the hierarchy of its owned x owned block (additive Schwarz with AMG
// sub-solves): no communication inside the preconditioner; CG's operator stays the global one.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <utility>
#include <vector>

#include "fb_amg_host.h"
#include "fb_ops.h"

using fb_amg_host::HostCsr;

namespace {

struct DevCsr {
  int nrows = 0, ncols = 0;
  DBuf<int> ptr, col;
  DBuf<double> val;
  void upload(const HostCsr &h, cudaStream_t st) {
    nrows = h.nrows;
    ncols = h.ncols;
    ptr.upload(h.ptr.data(), h.ptr.size(), st);
    col.upload(h.col.data(), h.col.size(), st);
    val.upload(h.val.data(), h.val.size(), st);
  }
};

// ---- device kernels: AT lanes per row, AT chosen per operator from its mean row length
// y = M x (MODE 0), y += M x (MODE 1), y = b - M x (MODE 2)
template <int AT, int MODE>
__global__ void k_amg_spmv(int nrows, const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                           const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ y) {
  const int lane = threadIdx.x % AT;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / AT;
  const int ngroups = (gridDim.x * blockDim.x) / AT;
  const int npad = ((nrows + ngroups - 1) / ngroups) * ngroups;
  for (int row = group; row < npad; row += ngroups) {
    double acc = 0.0;
    if (row < nrows)
      for (int k = ptr[row] + lane; k < ptr[row + 1]; k += AT) acc += val[k] * x[col[k]];
#pragma unroll
    for (int o = AT / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < nrows && lane == 0) {
      if (MODE == 0) y[row] = acc;
      if (MODE == 1) y[row] += acc;
      if (MODE == 2) y[row] = b[row] - acc;
    }
  }
}

// damped Jacobi from a zero initial guess: x = w D^-1 b
__global__ void k_amg_smooth0(int n, double w, const double *__restrict__ dinv, const double *__restrict__ b, double *x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = w * dinv[i] * b[i];
}

// xout = x + w D^-1 (b - A x)
template <int AT>
__global__ void k_amg_jacobi(int nrows, const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                             double w, const double *__restrict__ dinv, const double *__restrict__ b, const double *__restrict__ x,
                             double *__restrict__ xout) {
  const int lane = threadIdx.x % AT;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / AT;
  const int ngroups = (gridDim.x * blockDim.x) / AT;
  const int npad = ((nrows + ngroups - 1) / ngroups) * ngroups;
  for (int row = group; row < npad; row += ngroups) {
    double acc = 0.0;
    if (row < nrows)
      for (int k = ptr[row] + lane; k < ptr[row + 1]; k += AT) acc += val[k] * x[col[k]];
#pragma unroll
    for (int o = AT / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < nrows && lane == 0) xout[row] = x[row] + w * dinv[row] * (b[row] - acc);
  }
}

// coarsest level: x = Ainv b with the dense (pseudo-)inverse, one warp per row
__global__ void k_amg_dense(int n, const double *__restrict__ Ainv, const double *__restrict__ b, double *__restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += nwarps) {
    double acc = 0.0;
    for (int k = lane; k < n; k += 32) acc += Ainv[(size_t)i * n + k] * b[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) x[i] = acc;
  }
}

// In-place inverse of a dense symmetric positive definite matrix (row-major, n x n) by Cholesky; false if a pivot fails.
int lanes_for(const HostCsr &M) {
  const double mean = M.nrows > 0 ? (double)M.val.size() / M.nrows : 0.0;
  // The kernels are latency bound (row pointer -> entries -> x[col], three dependent loads per row): fewer lanes per row
  // put more rows in flight per warp.  FB_AMG_LANES=0: the first thresholds (12 / 32 / 96).
  static const int mode = getenv("FB_AMG_LANES") ? atoi(getenv("FB_AMG_LANES")) : 1;
  int at;
  if (mode == 0)
    at = mean <= 12.0 ? 4 : (mean <= 32.0 ? 8 : (mean <= 96.0 ? 16 : 32));
  else
    at = mean <= 20.0 ? 4 : (mean <= 56.0 ? 8 : (mean <= 160.0 ? 16 : 32));
  // few rows: spend more lanes per row rather than leave the GPU idle
  while (at < 32 && (int64_t)M.nrows * at < 148 * 256) at *= 2;
  return at;
}

}  // namespace

struct AmgLevel {
  DevCsr A, P, R;
  int lanes_A = 4, lanes_P = 4, lanes_R = 4;
  DBuf<double> dinv, x, b, r, tmp;
  DBuf<double> Ainv;  // coarsest level only: dense (pseudo-)inverse
  double omega = 0.67;
  int n = 0;
};

struct fb_amg {
  fb_ctx *ctx = nullptr;
  std::vector<AmgLevel *> levels;
  // replicated mode (partitioned runs)
  fb_peer_vec *gather = nullptr;
  DBuf<int> l2g;
  DBuf<double> zg;
  int n_owned = 0;
  double operator_complexity = 1.0;
  bool singular = false;
  // The V-cycle is a chain of ~16 kernels of 3 - 30 us each: replayed as ONE CUDA graph per (input, output) pair, the
  // launch gaps between them shrink (the work vectors of the PCG iteration that calls it keep their addresses).
  std::map<std::pair<const double *, double *>, cudaGraphExec_t> graphs;
  std::map<std::pair<const double *, double *>, int> seen;  // direct runs of a pair before it is captured
  int graph_kernels = 0;                                    // kernels per cycle (for the launch counter)
  bool graphs_enabled = true;
  void drop_graphs() {
    for (auto &g : graphs) cudaGraphExecDestroy(g.second);
    graphs.clear();
    seen.clear();
  }
  ~fb_amg() {
    drop_graphs();
    for (auto *l : levels) delete l;
    if (gather) fb_peer_vec_destroy(gather);
  }
};

void amg_set_replicated(fb_amg *amg, fb_peer_vec *gather, const int *l2g_host, int n_owned) {
  amg->gather = gather;
  amg->n_owned = n_owned;
  amg->l2g.upload(l2g_host, (size_t)n_owned, amg->ctx->dev->stream);
  amg->zg.alloc((size_t)amg->levels[0]->n);
  FB_CUDA(cudaStreamSynchronize(amg->ctx->dev->stream));
}

namespace {
__global__ void k_amg_extract(int n_owned, const int *__restrict__ l2g, const double *__restrict__ zg, double *__restrict__ z) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_owned; i += gridDim.x * blockDim.x) z[i] = zg[l2g[i]];
}
}  // namespace

void amg_destroy(fb_amg *amg) { delete amg; }

// Build the hierarchy of the n x n leading block of the CSR matrix (rowptr, col, val) given on the host
// (fb_amg_host.h) and upload it.
fb_amg *amg_setup(fb_ctx *ctx, int n, const int *rowptr, const int *col, const double *val) {
  cudaStream_t st = ctx->dev->stream;
  fb_amg *amg = new fb_amg();
  amg->ctx = ctx;
  std::vector<fb_amg_host::HostLevel> host;
  fb_amg_host::build_hierarchy(n, rowptr, col, val, host, amg->singular, amg->operator_complexity);
  for (size_t l = 0; l < host.size(); ++l) {
    const fb_amg_host::HostLevel &H = host[l];
    AmgLevel *L = new AmgLevel();
    amg->levels.push_back(L);
    L->n = H.n;
    L->omega = H.omega;
    L->A.upload(H.A, st);
    L->lanes_A = lanes_for(H.A);
    L->dinv.upload(H.dinv.data(), H.dinv.size(), st);
    for (DBuf<double> *v : {&L->x, &L->b, &L->r, &L->tmp}) v->alloc((size_t)std::max(1, H.n));
    if (!H.Ainv.empty()) L->Ainv.upload(H.Ainv.data(), H.Ainv.size(), st);
    if (l + 1 < host.size()) {
      L->P.upload(H.P, st);
      L->R.upload(H.R, st);
      L->lanes_P = lanes_for(H.P);
      L->lanes_R = lanes_for(H.R);
    }
    FB_CUDA(cudaStreamSynchronize(st));  // the host arrays of this level go away with `host`
  }
  return amg;
}

static inline int agrid(int n, int lanes) {
  int64_t g = ((int64_t)n * lanes + 255) / 256;
  return g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : (int)g);
}

template <int MODE>
static void amg_spmv(fb_ctx *ctx, const DevCsr &M, int lanes, const double *x, const double *b, double *y) {
  const int g = agrid(M.nrows, lanes);
  switch (lanes) {
    case 4: FB_LAUNCH(ctx, (k_amg_spmv<4, MODE>), g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, x, b, y); break;
    case 8: FB_LAUNCH(ctx, (k_amg_spmv<8, MODE>), g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, x, b, y); break;
    case 16: FB_LAUNCH(ctx, (k_amg_spmv<16, MODE>), g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, x, b, y); break;
    default: FB_LAUNCH(ctx, (k_amg_spmv<32, MODE>), g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, x, b, y); break;
  }
}

static void amg_jacobi(fb_ctx *ctx, const AmgLevel &L, const double *b, const double *x, double *xout) {
  const DevCsr &M = L.A;
  const int g = agrid(M.nrows, L.lanes_A);
  switch (L.lanes_A) {
    case 4: FB_LAUNCH(ctx, k_amg_jacobi<4>, g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, L.omega, L.dinv.p, b, x, xout); break;
    case 8: FB_LAUNCH(ctx, k_amg_jacobi<8>, g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, L.omega, L.dinv.p, b, x, xout); break;
    case 16: FB_LAUNCH(ctx, k_amg_jacobi<16>, g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, L.omega, L.dinv.p, b, x, xout); break;
    default: FB_LAUNCH(ctx, k_amg_jacobi<32>, g, 256, 0, M.nrows, M.ptr.p, M.col.p, M.val.p, L.omega, L.dinv.p, b, x, xout); break;
  }
}

// z = V-cycle(r): r and z are device vectors of the fine level (n entries).  Per level the working iterate lives
// in L.tmp (pre-smooth, residual, coarse correction) and the post-smoothed result in L.x (z on the fine level).
static void amg_cycle(fb_amg *amg, const double *r, double *z);

// amg_cycle through a CUDA graph: the first two calls with a given (r, z) run directly (they also fill the occupancy cache
// of FB_LAUNCH), the third is captured, later ones replay it.  Any failure of the graph API switches the hierarchy back to
// direct launches for good.  FB_AMG_GRAPH=0 disables the graphs.
static void amg_cycle_graphed(fb_amg *amg, const double *r, double *z) {
  static const bool env_on = !(getenv("FB_AMG_GRAPH") && atoi(getenv("FB_AMG_GRAPH")) == 0);
  fb_ctx *ctx = amg->ctx;
  cudaStream_t st = ctx->dev->stream;
  if (!env_on || !amg->graphs_enabled) return amg_cycle(amg, r, z);
  const auto key = std::make_pair(r, z);
  auto it = amg->graphs.find(key);
  if (it != amg->graphs.end()) {
    if (cudaGraphLaunch(it->second, st) == cudaSuccess) {
      ctx->launches += amg->graph_kernels;
      return;
    }
    cudaGetLastError();
    amg->graphs_enabled = false;
    amg->drop_graphs();
    return amg_cycle(amg, r, z);
  }
  if (amg->seen[key]++ < 2) return amg_cycle(amg, r, z);
  if (amg->graphs.size() >= 8) amg->drop_graphs();  // callers with ever-changing vectors: start over
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  const long long before = ctx->launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    amg->graphs_enabled = false;
    return amg_cycle(amg, r, z);
  }
  bool ok = true;
  try {
    amg_cycle(amg, r, z);
  } catch (...) {
    ok = false;
  }
  if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) ok = false;
  amg->graph_kernels = (int)(ctx->launches - before);
  ctx->launches = before;
  if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError();
    if (exec) cudaGraphExecDestroy(exec);
    amg->graphs_enabled = false;
    return amg_cycle(amg, r, z);  // nothing ran during the failed capture
  }
  amg->graphs[key] = exec;
  if (cudaGraphLaunch(exec, st) != cudaSuccess) {
    cudaGetLastError();
    amg->graphs_enabled = false;
    amg->drop_graphs();
    return amg_cycle(amg, r, z);
  }
  ctx->launches += amg->graph_kernels;
}

void amg_apply(fb_amg *amg, const double *r, double *z) {
  if (!amg->gather) return amg_cycle_graphed(amg, r, z);
  fb_ctx *ctx = amg->ctx;
  const double *rg = fb_peer_vec_gather(ctx, amg->gather, r, amg->l2g.p, amg->n_owned);
  amg_cycle_graphed(amg, rg, amg->zg.p);
  FB_LAUNCH(ctx, k_amg_extract, agrid(amg->n_owned, 1), 256, 0, amg->n_owned, amg->l2g.p, amg->zg.p, z);
}

static void amg_cycle(fb_amg *amg, const double *r, double *z) {
  fb_ctx *ctx = amg->ctx;
  const int nl = (int)amg->levels.size();
  if (nl == 1) {  // no coarse level: two Jacobi sweeps
    AmgLevel &L = *amg->levels[0];
    FB_LAUNCH(ctx, k_amg_smooth0, agrid(L.n, 1), 256, 0, L.n, L.omega, L.dinv.p, r, L.tmp.p);
    amg_jacobi(ctx, L, r, L.tmp.p, z);
    return;
  }
  // down sweep
  for (int l = 0; l < nl; ++l) {
    AmgLevel &L = *amg->levels[l];
    const double *b = (l == 0) ? r : L.b.p;
    if (l == nl - 1) {
      if (L.Ainv.p) {
        FB_LAUNCH(ctx, k_amg_dense, agrid(L.n, 32), 256, 0, L.n, L.Ainv.p, b, L.x.p);
      } else {  // coarsening stalled on a large level: a few Jacobi sweeps (result must end in L.x)
        FB_LAUNCH(ctx, k_amg_smooth0, agrid(L.n, 1), 256, 0, L.n, L.omega, L.dinv.p, b, L.x.p);
        for (int s = 0; s < 2; ++s) {
          amg_jacobi(ctx, L, b, L.x.p, L.tmp.p);
          amg_jacobi(ctx, L, b, L.tmp.p, L.x.p);
        }
      }
      break;
    }
    AmgLevel &C = *amg->levels[l + 1];
    FB_LAUNCH(ctx, k_amg_smooth0, agrid(L.n, 1), 256, 0, L.n, L.omega, L.dinv.p, b, L.tmp.p);  // pre-smooth from zero
    amg_spmv<2>(ctx, L.A, L.lanes_A, L.tmp.p, b, L.r.p);                                        // r = b - A x
    amg_spmv<0>(ctx, L.R, L.lanes_R, L.r.p, nullptr, C.b.p);                                    // restrict
  }
  // up sweep
  for (int l = nl - 2; l >= 0; --l) {
    AmgLevel &L = *amg->levels[l];
    AmgLevel &C = *amg->levels[l + 1];
    const double *b = (l == 0) ? r : L.b.p;
    double *out = (l == 0) ? z : L.x.p;
    amg_spmv<1>(ctx, L.P, L.lanes_P, C.x.p, nullptr, L.tmp.p);  // x += P xc
    amg_jacobi(ctx, L, b, L.tmp.p, out);                        // post-smooth
  }
}

int amg_num_levels(const fb_amg *amg) { return (int)amg->levels.size(); }
double amg_complexity(const fb_amg *amg) { return amg->operator_complexity; }
int amg_level_size(const fb_amg *amg, int l) { return amg->levels[l]->n; }
