// CUDA kernels of the hot path (sm_100a): pattern/scatter-map setup, element assembly
// (P1 stiffness, P2 mass, momentum residual + Jacobian, pressure / correction right-hand
// sides, heat operator), Dirichlet application, fp64 CSR / block-CSR SpMV with fused dot
// products, and the vector kernels.  fp64 throughout; bandwidth / atomic bound, so no
// tensor cores (BASELINE.json north_star).
//
// Storage of a D x D block matrix ("row-planar block CSR"): block row I with nb blocks
// holds D consecutive scalar rows of nb*D values each, i.e. the value array IS the
// scalar CSR value array of the interleaved system; only the column indices are
// compressed to one int per block.  SpMV lanes stream each scalar row contiguously
// (coalesced) and share one x gather between the D rows.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "fb_element.cuh"
#include <memory>
#include <mutex>
#include <numeric>

#include "fb_ops.h"

namespace hq {
#define FB_TABLE static const
#include "fb_quadrature.h"
#undef FB_TABLE
}  // namespace hq
namespace dq {
#define FB_TABLE static __constant__ const
#include "fb_quadrature.h"
#undef FB_TABLE
}  // namespace dq
#define FB_TABLE static __device__ const
#include "fb_p2_tables.h"  // FB_M3_TRI / FB_M3_TET: copied to shared memory by k_momentum_J_cf
#undef FB_TABLE

// ---- quadrature accessors ---------------------------------------------------
template <int D>
struct Q5;  // degree-5 cell rule
template <>
struct Q5<2> {
  static constexpr int NQ = dq::TRI_D5_NQ;
  __device__ static double lam(int q, int m) { return dq::TRI_D5_LAM[q][m]; }
  __device__ static double w(int q) { return dq::TRI_D5_W[q]; }
  static double hlam(int q, int m) { return hq::TRI_D5_LAM[q][m]; }
  static double hw(int q) { return hq::TRI_D5_W[q]; }
};
template <>
struct Q5<3> {
  static constexpr int NQ = dq::TET_D5_NQ;
  __device__ static double lam(int q, int m) { return dq::TET_D5_LAM[q][m]; }
  __device__ static double w(int q) { return dq::TET_D5_W[q]; }
  static double hlam(int q, int m) { return hq::TET_D5_LAM[q][m]; }
  static double hw(int q) { return hq::TET_D5_W[q]; }
};
template <int D>
struct Q2;  // degree-2 cell rule
template <>
struct Q2<2> {
  static constexpr int NQ = dq::TRI_D2_NQ;
  __device__ static const double *lam_ptr() { return &dq::TRI_D2_LAM[0][0]; }
  __device__ static const double *w_ptr() { return dq::TRI_D2_W; }
  __device__ static double lam(int q, int m) { return dq::TRI_D2_LAM[q][m]; }
  __device__ static double w(int q) { return dq::TRI_D2_W[q]; }
};
template <>
struct Q2<3> {
  static constexpr int NQ = dq::TET_D2_NQ;
  __device__ static const double *lam_ptr() { return &dq::TET_D2_LAM[0][0]; }
  __device__ static const double *w_ptr() { return dq::TET_D2_W; }
  __device__ static double lam(int q, int m) { return dq::TET_D2_LAM[q][m]; }
  __device__ static double w(int q) { return dq::TET_D2_W[q]; }
};
template <int D>
struct QF;  // facet rule (degree 5) in facet barycentrics (D entries)
template <>
struct QF<2> {
  static constexpr int NQ = dq::SEG_D5_NQ;
  __device__ static const double *lam_ptr() { return &dq::SEG_D5_LAM[0][0]; }
  __device__ static const double *w_ptr() { return dq::SEG_D5_W; }
};
template <>
struct QF<3> {
  static constexpr int NQ = dq::TRI_D5_NQ;
  __device__ static const double *lam_ptr() { return &dq::TRI_D5_LAM[0][0]; }
  __device__ static const double *w_ptr() { return dq::TRI_D5_W; }
};

__constant__ double c_mref2[6 * 6];    // P2 reference mass / |K|, triangle
__constant__ double c_mref3[10 * 10];  // tetrahedron
static bool g_mref_uploaded = false;

template <int D>
static void upload_mref_dim(double *dst_symbol_host) {
  constexpr int NL = Elem<D>::NL2;
  for (int a = 0; a < NL; ++a)
    for (int b = 0; b < NL; ++b) {
      double s = 0;
      for (int q = 0; q < Q5<D>::NQ; ++q) {
        double lam[D + 1];
        for (int m = 0; m <= D; ++m) lam[m] = Q5<D>::hlam(q, m);
        s += Q5<D>::hw(q) * fb_p2_phi<D>(a, lam) * fb_p2_phi<D>(b, lam);
      }
      dst_symbol_host[a * NL + b] = s;
    }
}

static void upload_mref() {
  if (g_mref_uploaded) return;
  double m2[36], m3[100];
  upload_mref_dim<2>(m2);
  upload_mref_dim<3>(m3);
  FB_CUDA(cudaMemcpyToSymbol(c_mref2, m2, sizeof(m2)));
  FB_CUDA(cudaMemcpyToSymbol(c_mref3, m3, sizeof(m3)));
  g_mref_uploaded = true;
}

int fb_clamp_grid(fb_ctx *ctx, const void *kernel, int grid, int block, size_t smem) {
  static std::unordered_map<const void *, int> cache;  // resident blocks per SM of each kernel
  static std::mutex cache_mutex;                       // contexts of several host threads share the cache
  std::lock_guard<std::mutex> guard(cache_mutex);
  const uint64_t key_block = (uint64_t)block;
  const void *key = (const void *)((uintptr_t)kernel ^ (key_block << 48));
  auto it = cache.find(key);
  int per_sm;
  if (it == cache.end()) {
    per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cache[key] = per_sm;
  } else {
    per_sm = it->second;
  }
  const int wave = per_sm * ctx->dev->sm_count;
  return grid < wave ? grid : wave;
}

// ---- device block cache behind DBuf (fb_device.cuh) ---------------------------------------------------------------
// cudaFree synchronises the device and, like cudaMalloc, costs 0.1 - 1 ms: drivers that rebuild an operator every step
// (a new Heat per Banach iteration, tests/test_boussinesq.py:220-227) or solvers whose work vectors alternate between two
// sizes paid that dozens of times per step.  Released blocks are kept per (device, size) and handed out again; every
// user of a context enqueues on the context's one stream, so a recycled block is ordered behind its previous uses.
namespace {
struct BlockCache {
  std::mutex m;
  std::unordered_multimap<uint64_t, void *> free_blocks;  // key: device << 48 | bytes / 256
  size_t cached = 0;
};
BlockCache &block_cache() {
  static BlockCache c;
  return c;
}
constexpr size_t FB_CACHE_TOTAL = size_t(8) << 30, FB_CACHE_BLOCK = size_t(1) << 30;
inline uint64_t cache_key(int dev, size_t bytes) { return ((uint64_t)dev << 48) | (uint64_t)(bytes >> 8); }
void cache_flush(int dev) {  // caller holds the lock
  BlockCache &c = block_cache();
  for (auto it = c.free_blocks.begin(); it != c.free_blocks.end();) {
    if ((int)(it->first >> 48) == dev) {
      cudaFree(it->second);
      c.cached -= (size_t)(it->first & ((uint64_t(1) << 48) - 1)) << 8;
      it = c.free_blocks.erase(it);
    } else {
      ++it;
    }
  }
}
}  // namespace

void *fb_block_alloc(size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  int dev = 0;
  FB_CUDA(cudaGetDevice(&dev));
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> guard(c.m);
  auto it = c.free_blocks.find(cache_key(dev, bytes));
  if (it != c.free_blocks.end()) {
    void *p = it->second;
    c.free_blocks.erase(it);
    c.cached -= bytes;
    return p;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try again
    cudaGetLastError();
    cache_flush(dev);
    e = cudaMalloc(&p, bytes);
  }
  FB_CUDA(e);
  return p;
}

void fb_block_free(void *p, size_t bytes) {
  if (!p) return;
  bytes = (bytes + 255) & ~size_t(255);
  BlockCache &c = block_cache();
  cudaPointerAttributes attr;  // the device that owns the block (the caller's current device may be another one)
  if (bytes <= FB_CACHE_BLOCK && cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
    const int dev = attr.device;
    std::lock_guard<std::mutex> guard(c.m);
    if (c.cached + bytes <= FB_CACHE_TOTAL) {
      c.free_blocks.emplace(cache_key(dev, bytes), p);
      c.cached += bytes;
      return;
    }
  }
  cudaFree(p);
}

static inline int grid_for(int64_t work_items, int block, int cap) {
  int64_t g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// =============================================================================
// setup kernels
// =============================================================================
__global__ void k_scatter_map(int64_t nc, int nl, const int *__restrict__ cell_nodes, const int *__restrict__ rowptr,
                              const int *__restrict__ col, int *__restrict__ smap) {
  const int64_t total = nc * nl * nl;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / (nl * nl);
    const int r = (int)(t - c * nl * nl);
    const int a = r / nl, b = r - a * nl;
    const int I = cell_nodes[c * nl + a], J = cell_nodes[c * nl + b];
    int lo = rowptr[I], hi = rowptr[I + 1] - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col[mid] < J)
        lo = mid + 1;
      else
        hi = mid;
    }
    smap[t] = lo;
  }
}

__global__ void k_diag_slot(int64_t n, const int *__restrict__ rowptr, const int *__restrict__ col, int *__restrict__ diag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = rowptr[i], hi = rowptr[i + 1] - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col[mid] < (int)i)
        lo = mid + 1;
      else
        hi = mid;
    }
    diag[i] = lo;
  }
}

void dev_space_build(fb_space *s, DevSpace &d) {
  fb_ctx *ctx = s->mesh->ctx;
  cudaStream_t st = ctx->dev->stream;
  upload_mref();
  fb_space_build_pattern(s);
  const fb_mesh *m = s->mesh;
  d.ctx = ctx;
  d.host = s;
  d.dim = m->dim;
  d.nl = s->nl;
  d.degree = s->degree;
  d.nnodes = s->nnodes;
  d.nc = m->nc;
  d.nnz = (int64_t)s->indices.size();
  if (d.nnz * (int64_t)(m->dim * m->dim) > (int64_t)8e9 || d.nnz > INT32_MAX)
    throw fb_cuda_error(FB_EINVAL, "pattern too large for one GPU (partition the mesh)");
  d.xyz.upload(m->xyz.data(), m->xyz.size(), st);
  // The device copy stores the cells in Morton order of their centroids (a private order: nothing indexed by cell
  // crosses the ABI).  Cells that share matrix rows then sit next to each other, so the scatter-adds of the assembly
  // kernels to one row meet in L2 instead of arriving a mesh plane apart, and the node gathers of neighbouring
  // threads hit the same lines.  FB_CELL_ORDER=0 keeps the mesh's own order.
  const std::vector<int32_t> &corder = fb_mesh_cell_order(s->mesh);
  {
    const int nl = s->nl, nv = m->dim + 1;
    std::vector<int32_t> cn((size_t)m->nc * nl), cv((size_t)m->nc * nv);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < m->nc; ++c) {
      const int64_t o = corder[c];
      std::memcpy(&cn[c * nl], &s->cell_nodes[o * nl], sizeof(int32_t) * nl);
      std::memcpy(&cv[c * nv], &m->cells[o * nv], sizeof(int32_t) * nv);
    }
    d.cell_nodes.upload(cn.data(), cn.size(), st);
    d.cells.upload(cv.data(), cv.size(), st);
    FB_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<int> rp(s->nnodes + 1);
  for (int64_t i = 0; i <= s->nnodes; ++i) rp[i] = (int)s->indptr[i];
  d.rowptr.upload(rp.data(), rp.size(), st);
  d.col.upload(s->indices.data(), s->indices.size(), st);
  d.diag.alloc(s->nnodes);
  d.smap.alloc((size_t)d.nc * d.nl * d.nl);
  d.n_owned = s->n_owned;
  d.halo_ranks.assign(s->halo_ranks.begin(), s->halo_ranks.end());
  d.halo_send_ptr = s->halo_send_ptr;
  d.halo_recv_ptr = s->halo_recv_ptr;
  if (!s->halo_send_nodes.empty()) d.halo_send_nodes.upload(s->halo_send_nodes.data(), s->halo_send_nodes.size(), st);
  d.nbf = (int64_t)m->bf_cell.size();
  {
    std::vector<int32_t> inv((size_t)m->nc), bfc(m->bf_cell.size());
    for (int64_t c = 0; c < m->nc; ++c) inv[corder[c]] = (int32_t)c;
    for (size_t k = 0; k < bfc.size(); ++k) bfc[k] = inv[m->bf_cell[k]];
    // facets sorted by (cell, local facet): a cell's facets are consecutive (k_momentum_J_facets_elem)
    std::vector<size_t> ord(bfc.size());
    std::iota(ord.begin(), ord.end(), (size_t)0);
    std::sort(ord.begin(), ord.end(), [&](size_t x, size_t y) {
      return bfc[x] != bfc[y] ? bfc[x] < bfc[y] : m->bf_local[x] < m->bf_local[y];
    });
    std::vector<int32_t> bfc2(bfc.size()), bfl2(bfc.size());
    for (size_t k = 0; k < ord.size(); ++k) {
      bfc2[k] = bfc[ord[k]];
      bfl2[k] = m->bf_local[ord[k]];
    }
    d.bf_cell.upload(bfc2.data(), bfc2.size(), st);
    d.bf_local.upload(bfl2.data(), bfl2.size(), st);
    FB_CUDA(cudaStreamSynchronize(st));
  }
  FB_LAUNCH(ctx, k_scatter_map, grid_for(d.nc * d.nl * d.nl, 256, 148 * 16), 256, 0, d.nc, d.nl, d.cell_nodes.p,
            d.rowptr.p, d.col.p, d.smap.p);
  FB_LAUNCH(ctx, k_diag_slot, grid_for(d.nnodes, 256, 148 * 16), 256, 0, d.nnodes, d.rowptr.p, d.col.p, d.diag.p);
  FB_CUDA(cudaStreamSynchronize(st));  // rp is a temporary
}

LinOp make_linop(const fb_mat &m, int ncomp, const uint8_t *mask) {
  LinOp A;
  A.block = m.block;
  A.ncomp = m.block > 1 ? 1 : ncomp;
  A.nrows = m.sp->n_owned;
  A.nlocal = m.sp->nnodes;
  A.halo = (fb_is_distributed(m.ctx) && !m.sp->halo_ranks.empty()) ? m.sp : nullptr;
  A.rowptr = m.sp->rowptr.p;
  A.col = m.sp->col.p;
  A.val = m.val.p;
  A.mask = mask;
  if (m.block == 1 && m.tval.p && m.sp->tile && tile_enabled()) {
    A.tile = m.sp->tile;
    A.tval = m.tval.p;
  }
  return A;
}

void mat_enable_tile(fb_ctx *ctx, fb_mat &m) {
  if (m.block != 1 || !tile_enabled()) return;
  DevSpace *sp = m.sp;
  if (!sp->tile) {
    std::unique_ptr<TileFormat> tf(new TileFormat());
    tile_format_build(sp->host, *tf, ctx->dev->stream);
    sp->tile = tf.release();
  }
  m.tval.alloc((size_t)sp->tile->nent);
  tile_pack(ctx, *sp->tile, m.val.p, m.tval.p);
}

void mat_repack(fb_ctx *ctx, fb_mat &m) {
  if (m.tval.p && m.sp->tile) tile_pack(ctx, *m.sp->tile, m.val.p, m.tval.p);
}

// =============================================================================
// constant operators: one thread per (cell, a, b)
// =============================================================================
template <int D, int DEG, int KIND>
__global__ void k_assemble_constant(int64_t nc, const int *__restrict__ cells, const double *__restrict__ xyz,
                                    const int *__restrict__ smap, double *__restrict__ val) {
  constexpr int NL = DEG == 1 ? Elem<D>::NL1 : Elem<D>::NL2;
  const int64_t total = nc * NL * NL;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / (NL * NL);
    const int r = (int)(t - c * NL * NL);
    const int a = r / NL, b = r - a * NL;
    double X[(D + 1) * D];
#pragma unroll
    for (int v = 0; v <= D; ++v) {
      const int node = cells[c * (D + 1) + v];
#pragma unroll
      for (int k = 0; k < D; ++k) X[v * D + k] = xyz[(int64_t)node * D + k];
    }
    double glam[D + 1][D], vol;
    fb_geometry<D>(X, glam, vol);
    double e = 0.0;
    if (KIND == 1) {  // mass
      if (DEG == 1)
        e = vol * (a == b ? 2.0 : 1.0) / ((D + 1) * (D + 2));
      else
        e = vol * (D == 2 ? c_mref2[a * NL + b] : c_mref3[a * NL + b]);
    } else {  // stiffness
      if (DEG == 1) {
#pragma unroll
        for (int k = 0; k < D; ++k) e += glam[a][k] * glam[b][k];
        e *= vol;
      } else {
        for (int q = 0; q < Q2<D>::NQ; ++q) {
          double lam[D + 1], ga[D], gb[D];
#pragma unroll
          for (int m = 0; m <= D; ++m) lam[m] = Q2<D>::lam(q, m);
          fb_p2_grad<D>(a, lam, glam, ga);
          fb_p2_grad<D>(b, lam, glam, gb);
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s += ga[k] * gb[k];
          e += Q2<D>::w(q) * s;
        }
        e *= vol;
      }
    }
    atomicAdd(&val[smap[t]], e);
  }
}

void assemble_constant(fb_ctx *ctx, DevSpace &sp, int kind, double *val) {
  FB_CUDA(cudaMemsetAsync(val, 0, sizeof(double) * sp.nnz, ctx->dev->stream));
  const int g = grid_for(sp.nc * sp.nl * sp.nl, 256, 148 * 32);
#define FB_AC(D, DEG, KIND)                                                                                     \
  FB_LAUNCH(ctx, (k_assemble_constant<D, DEG, KIND>), g, 256, 0, sp.nc, sp.cells.p, sp.xyz.p, sp.smap.p, val)
  if (sp.dim == 2 && sp.degree == 1 && kind == 0) FB_AC(2, 1, 0);
  else if (sp.dim == 2 && sp.degree == 1) FB_AC(2, 1, 1);
  else if (sp.dim == 2 && kind == 0) FB_AC(2, 2, 0);
  else if (sp.dim == 2) FB_AC(2, 2, 1);
  else if (sp.degree == 1 && kind == 0) FB_AC(3, 1, 0);
  else if (sp.degree == 1) FB_AC(3, 1, 1);
  else if (kind == 0) FB_AC(3, 2, 0);
  else FB_AC(3, 2, 1);
#undef FB_AC
}

template <int D>
__global__ void k_lumped(int64_t nc, int nl, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                         const double *__restrict__ xyz, double *__restrict__ diag) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    double X[(D + 1) * D];
    int nodes[D + 1];
    for (int v = 0; v <= D; ++v) {
      nodes[v] = cell_nodes[c * nl + v];
      for (int k = 0; k < D; ++k) X[v * D + k] = xyz[(int64_t)cells[c * (D + 1) + v] * D + k];
    }
    double glam[D + 1][D], vol;
    fb_geometry<D>(X, glam, vol);
    for (int v = 0; v <= D; ++v) atomicAdd(&diag[nodes[v]], vol / (D + 1));
  }
}

void assemble_lumped(fb_ctx *ctx, DevSpace &sp, double *diag) {
  FB_CUDA(cudaMemsetAsync(diag, 0, sizeof(double) * sp.nnodes, ctx->dev->stream));
  const int g = grid_for(sp.nc, 256, 148 * 16);
  if (sp.dim == 2)
    FB_LAUNCH(ctx, k_lumped<2>, g, 256, 0, sp.nc, sp.nl, sp.cell_nodes.p, sp.cells.p, sp.xyz.p, diag);
  else
    FB_LAUNCH(ctx, k_lumped<3>, g, 256, 0, sp.nc, sp.nl, sp.cell_nodes.p, sp.cells.p, sp.xyz.p, diag);
}

// =============================================================================
// SpMV
// =============================================================================
// scalar CSR times NC interleaved vectors; T lanes cooperate on a row.
template <int NC, int T, int DOT>
__global__ void __launch_bounds__(256)
    k_spmm(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
           const uint8_t *__restrict__ mask, const double *__restrict__ x, double *__restrict__ y,
           const double *__restrict__ w, double *partials, unsigned int *counter, double *red, int slot,
           const int *__restrict__ flag) {
  if (flag && *flag) return;
  const int lane = threadIdx.x % T;
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
  double d[3] = {0.0, 0.0, 0.0};
  const int64_t nrows_pad = ((nrows + ngroups - 1) / ngroups) * ngroups;  // keep shuffles convergent
  for (int64_t row = group; row < nrows_pad; row += ngroups) {
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    double wv = 0.0, xd = 0.0;  // epilogue operands fetched up front (see k_bspmv_u)
    bool masked = false;
    if (row < nrows && lane < NC) {
      const int64_t dof = row * NC + lane;
      if (DOT >= 1) wv = w[dof];
      if (DOT == 3 || mask) xd = x[dof];
      if (mask) masked = mask[dof] != 0;
    }
    if (row < nrows) {
      const int r0 = rowptr[row], r1 = rowptr[row + 1];
      for (int k = r0 + lane; k < r1; k += T) {
        const double a = __ldcs(&val[k]);
        const int64_t j = col[k];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += a * x[j * NC + c];
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (row < nrows && lane < NC) {
      // lane c finalises component c
      double yc = acc[0];
#pragma unroll
      for (int c = 1; c < NC; ++c)
        if (lane == c) yc = acc[c];
      const int64_t dof = row * NC + lane;
      if (masked) yc = xd;
      y[dof] = yc;
      if (DOT == 1 || DOT == 2) d[0] += wv * yc;
      if (DOT == 2) d[1] += yc * yc;
      if (DOT == 3) {  // single-reduction CG: (w.x, y.x, x.x) with x the multiplied vector
        d[0] += wv * xd;
        d[1] += yc * xd;
        d[2] += xd * xd;
      }
    }
  }
  if (DOT >= 1) {
    if (DOT == 1) {
      double v1[1] = {d[0]};
      fb_grid_reduce<1>(v1, partials, counter, red, slot);
    } else if (DOT == 2) {
      double v2[2] = {d[0], d[1]};
      fb_grid_reduce<2>(v2, partials, counter, red, slot);
    } else {
      fb_grid_reduce<3>(d, partials, counter, red, slot);
    }
  }
}

// Same product with U row chunks per lane in flight (all index / value / x loads of a row are issued before the
// FMAs) and, if CHUNKED, a contiguous range of rows per block instead of a grid-stride walk: neighbouring rows
// share most of their columns, so a block that sweeps consecutive rows finds its x gathers in L1.
template <int NC, int T, int DOT, int U, int CHUNKED, int PAIR = 0>
__global__ void __launch_bounds__(256)
    k_spmm_u(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
             const uint8_t *__restrict__ mask, const double *__restrict__ x, double *__restrict__ y,
             const double *__restrict__ w, double *partials, unsigned int *counter, double *red, int slot,
             const int *__restrict__ flag, int64_t xlen) {
  if (flag && *flag) return;
  const int lane = threadIdx.x % T;
  constexpr int RPB = 256 / T;  // rows per block per sweep
  const int64_t per_block = (((nrows + gridDim.x - 1) / gridDim.x + RPB - 1) / RPB) * RPB;
  int64_t row, row_end, stride;
  if (CHUNKED) {
    row = blockIdx.x * per_block + threadIdx.x / T;
    row_end = (blockIdx.x + 1) * per_block;  // padded: all groups of the block run the same trip count
    stride = RPB;
  } else {
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
    row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
    row_end = ((nrows + ngroups - 1) / ngroups) * ngroups;
    stride = ngroups;
  }
  double d[3] = {0.0, 0.0, 0.0};
  for (; row < row_end; row += stride) {
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    // epilogue operands fetched up front (see k_bspmv_u)
    double wv = 0.0, xd = 0.0;
    bool masked = false;
    if (row < nrows && lane < NC) {
      const int64_t dof = row * NC + lane;
      if (DOT >= 1) wv = w[dof];
      if (DOT == 3 || mask) xd = x[dof];
      if (mask) masked = mask[dof] != 0;
    }
    if (row < nrows) {
      const int r0 = rowptr[row], r1 = rowptr[row + 1];
      for (int k0 = r0 + lane; k0 < r1; k0 += T * U) {
        int64_t j[U];
        double a[U], xv[U][NC];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = (k0 + u * T < r1) ? col[k0 + u * T] : -1;
        if (PAIR && NC == 3) {
          // the three values of node j sit at doubles 3j .. 3j+2: two aligned 16-byte loads cover them for either
          // parity of j (instead of three 8-byte loads; the gather is bound by L1 wavefronts, not by bytes)
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (j[u] >= 0) {
              const int64_t e = 3 * j[u];
              const double2 lo = *reinterpret_cast<const double2 *>(x + (e & ~int64_t(1)));
              double2 hi;
              if ((e | 1) + 3 <= xlen)
                hi = *reinterpret_cast<const double2 *>(x + (e & ~int64_t(1)) + 2);
              else
                hi = make_double2(x[e + 2], 0.0);  // last node of an odd-sized vector: stay inside the array
              const bool odd = e & 1;
              xv[u][0] = odd ? lo.y : lo.x;
              xv[u][1] = odd ? hi.x : lo.y;
              xv[u][2] = odd ? hi.y : hi.x;
            } else {
              xv[u][0] = xv[u][1] = xv[u][2] = 0.0;
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int c = 0; c < NC; ++c) xv[u][c] = (j[u] >= 0) ? x[j[u] * NC + c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (j[u] >= 0) ? __ldcs(&val[k0 + u * T]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] += a[u] * xv[u][c];
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (row < nrows && lane < NC) {
      double yc = acc[0];
#pragma unroll
      for (int c = 1; c < NC; ++c)
        if (lane == c) yc = acc[c];
      const int64_t dof = row * NC + lane;
      if (masked) yc = xd;
      y[dof] = yc;
      if (DOT == 1 || DOT == 2) d[0] += wv * yc;
      if (DOT == 2) d[1] += yc * yc;
      if (DOT == 3) {
        d[0] += wv * xd;
        d[1] += yc * xd;
        d[2] += xd * xd;
      }
    }
  }
  if (DOT >= 1) {
    if (DOT == 1) {
      double v1[1] = {d[0]};
      fb_grid_reduce<1>(v1, partials, counter, red, slot);
    } else if (DOT == 2) {
      double v2[2] = {d[0], d[1]};
      fb_grid_reduce<2>(v2, partials, counter, red, slot);
    } else {
      fb_grid_reduce<3>(d, partials, counter, red, slot);
    }
  }
}

// D x D row-planar block CSR; one warp (T = 32) or half warp per block row.
template <int D, int T, int DOT>
__global__ void __launch_bounds__(256)
    k_bspmv(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
            const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, double *partials,
            unsigned int *counter, double *red, int slot, const int *__restrict__ flag) {
  if (flag && *flag) return;
  const int lane = threadIdx.x % T;
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
  double d[2] = {0.0, 0.0};
  const int64_t nrows_pad = ((nrows + ngroups - 1) / ngroups) * ngroups;
  for (int64_t row = group; row < nrows_pad; row += ngroups) {
    double acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.0;
    if (row < nrows) {
      const int r0 = rowptr[row];
      const int len = (rowptr[row + 1] - r0) * D;
      const double *v = val + (int64_t)r0 * (D * D);
      for (int t = lane; t < len; t += T) {
        const int blk = t / D;
        const int j = t - blk * D;
        const double xv = x[(int64_t)col[r0 + blk] * D + j];
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] += __ldcs(&v[i * len + t]) * xv;
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (row < nrows && lane < D) {
      double yc = acc[0];
#pragma unroll
      for (int i = 1; i < D; ++i)
        if (lane == i) yc = acc[i];
      const int64_t dof = row * D + lane;
      y[dof] = yc;
      if (DOT >= 1) d[0] += w[dof] * yc;
      if (DOT >= 2) d[1] += yc * yc;
    }
  }
  if (DOT >= 1) {
    if (DOT == 1) {
      double v1[1] = {d[0]};
      fb_grid_reduce<1>(v1, partials, counter, red, slot);
    } else {
      fb_grid_reduce<2>(d, partials, counter, red, slot);
    }
  }
}

// Same product, U independent row chunks per lane in flight (index -> x gather -> values are
// issued for all chunks before the FMAs) to cover DRAM latency with fewer resident warps.
template <int D, int T, int DOT, int U, int MINB, typename VT = double, int CHUNKED = 0>
__global__ void __launch_bounds__(256, MINB)
    k_bspmv_u(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col, const VT *__restrict__ val,
              const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ w, double *partials,
              unsigned int *counter, double *red, int slot, const int *__restrict__ flag) {
  if (flag && *flag) return;
  const int lane = threadIdx.x % T;
  double d[2] = {0.0, 0.0};
  int64_t row, nrows_pad, stride;
  if (CHUNKED) {  // a contiguous range of rows per block (see k_spmm_u)
    constexpr int RPB = 256 / T;
    const int64_t per_block = (((nrows + gridDim.x - 1) / gridDim.x + RPB - 1) / RPB) * RPB;
    row = blockIdx.x * per_block + threadIdx.x / T;
    nrows_pad = (blockIdx.x + 1) * per_block;
    stride = RPB;
  } else {
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
    row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
    nrows_pad = ((nrows + ngroups - 1) / ngroups) * ngroups;
    stride = ngroups;
  }
  for (; row < nrows_pad; row += stride) {
    double acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.0;
    // the dot-product operand is fetched up front: loaded in the epilogue its latency would sit between two rows
    double wv = 0.0;
    if (DOT >= 1 && row < nrows && lane < D) wv = w[row * D + lane];
    if (row < nrows) {
      const int r0 = rowptr[row];
      const int len = (rowptr[row + 1] - r0) * D;
      const VT *v = val + (int64_t)r0 * (D * D);
      for (int t0 = lane; t0 < len; t0 += T * U) {
        int c[U];
        double xv[U];
        VT a[U][D];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int t = t0 + u * T;
          c[u] = (t < len) ? col[r0 + t / D] * D + (t % D) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = (c[u] >= 0) ? x[c[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int t = t0 + u * T;
#pragma unroll
          for (int i = 0; i < D; ++i) a[u][i] = (t < len) ? __ldcs(&v[i * len + t]) : VT(0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < D; ++i) acc[i] += (double)a[u][i] * xv[u];
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (row < nrows && lane < D) {
      double yc = acc[0];
#pragma unroll
      for (int i = 1; i < D; ++i)
        if (lane == i) yc = acc[i];
      const int64_t dof = row * D + lane;
      y[dof] = yc;
      if (DOT >= 1) d[0] += wv * yc;
      if (DOT >= 2) d[1] += yc * yc;
    }
  }
  if (DOT >= 1) {
    if (DOT == 1) {
      double v1[1] = {d[0]};
      fb_grid_reduce<1>(v1, partials, counter, red, slot);
    } else {
      fb_grid_reduce<2>(d, partials, counter, red, slot);
    }
  }
}

template <int D, int T, int U, int MINB>
static void launch_bspmv_u(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
                           const int *flag) {
  fb_device_state *dv = ctx->dev;
  const int block = 256;
  const int64_t rows_per_block = block / T;
  const int g = grid_for((A.nrows + rows_per_block - 1) / rows_per_block * block, block, dv->sm_count * 8);
#define FB_BSU(DOT)                                                                                                    \
  FB_LAUNCH(ctx, (k_bspmv_u<D, T, DOT, U, MINB>), g, block, 0, A.nrows, A.rowptr, A.col, A.val, x, y, w, dv->partials, \
            dv->counter, dv->red, slot, flag)
#define FB_BSU32(DOT)                                                                                                   \
  FB_LAUNCH(ctx, (k_bspmv_u<D, T, DOT, U, MINB, float>), g, block, 0, A.nrows, A.rowptr, A.col, A.val32, x, y, w,        \
            dv->partials, dv->counter, dv->red, slot, flag)
  if (A.val32) {  // fp32-stored operator (mixed-precision chord Jacobian); accumulation stays fp64
    if (dot_mode == 0) FB_BSU32(0);
    else if (dot_mode == 1) FB_BSU32(1);
    else FB_BSU32(2);
    return;
  }
  if (dot_mode == 0) FB_BSU(0);
  else if (dot_mode == 1) FB_BSU(1);
  else FB_BSU(2);
#undef FB_BSU
#undef FB_BSU32
}

template <int NC, int T>
static void launch_spmm(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
                        const int *flag) {
  fb_device_state *dv = ctx->dev;
  const int block = 256;
  const int64_t rows_per_block = block / T;
  const int g = grid_for((A.nrows + rows_per_block - 1) / rows_per_block * block, block, dv->sm_count * 8);
#define FB_SP(DOT)                                                                                                  \
  FB_LAUNCH(ctx, (k_spmm<NC, T, DOT>), g, block, 0, A.nrows, A.rowptr, A.col, A.val, A.mask, x, y, w, dv->partials, \
            dv->counter, dv->red, slot, flag)
  if (dot_mode == 0) FB_SP(0);
  else if (dot_mode == 1) FB_SP(1);
  else if (dot_mode == 2) FB_SP(2);
  else FB_SP(3);
#undef FB_SP
}

template <int NC, int T, int U, int CHUNKED, int PAIR = 0>
static void launch_spmm_u(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
                          const int *flag) {
  fb_device_state *dv = ctx->dev;
  const int block = 256;
  const int64_t rows_per_block = block / T;
  const int g = grid_for((A.nrows + rows_per_block - 1) / rows_per_block * block, block, dv->sm_count * 8);
#define FB_SPU(DOT)                                                                                                      \
  FB_LAUNCH(ctx, (k_spmm_u<NC, T, DOT, U, CHUNKED, PAIR>), g, block, 0, A.nrows, A.rowptr, A.col, A.val, A.mask, x, y, w,        \
            dv->partials, dv->counter, dv->red, slot, flag, A.nlocal * NC)
  if (dot_mode == 0) FB_SPU(0);
  else if (dot_mode == 1) FB_SPU(1);
  else if (dot_mode == 2) FB_SPU(2);
  else FB_SPU(3);
#undef FB_SPU
}

template <int D, int T>
static void launch_bspmv(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
                         const int *flag) {
  fb_device_state *dv = ctx->dev;
  const int block = 256;
  const int64_t rows_per_block = block / T;
  const int g = grid_for((A.nrows + rows_per_block - 1) / rows_per_block * block, block, dv->sm_count * 8);
#define FB_BS(DOT)                                                                                             \
  FB_LAUNCH(ctx, (k_bspmv<D, T, DOT>), g, block, 0, A.nrows, A.rowptr, A.col, A.val, x, y, w, dv->partials,   \
            dv->counter, dv->red, slot, flag)
  if (dot_mode == 0) FB_BS(0);
  else if (dot_mode == 1) FB_BS(1);
  else FB_BS(2);
#undef FB_BS
}

static void spmv_local(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
                       const int *flag);

void spmv(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
          const int *flag) {
  // distributed: the ghost entries of x are refreshed from their owners first (x is logically const:
  // only its ghost copies are rewritten), the fused dot products are summed over the ranks afterwards
  if (A.halo) halo_exchange(ctx, *A.halo, const_cast<double *>(x), A.dofs_per_node());
  spmv_local(ctx, A, x, y, dot_mode, w, slot, flag);
  if (dot_mode > 0) fb_allreduce_slots(ctx, slot, dot_mode);
}

static void spmv_local(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
                       const int *flag) {
  if (A.block == 2) return launch_bspmv_u<2, 8, 4, 1>(ctx, A, x, y, dot_mode, w, slot, flag);
  if (A.block == 3) {
    // Variants measured on B200 at n = 74 (profiles/r1_spmv_variants.txt): one chunk per lane, 32 lanes per row 1.76 ms;
    // T=32/U=2 1.37 ms; T=16/U=3 1.28 ms; T=16/U=4 1.25 ms (7.9 GB of DRAM traffic at 6.4 TB/s); T=16/U=6 1.25 ms;
    // T=8/U=4 1.44 ms; contiguous row ranges per block 1.96 ms.  FB_BSPMV_V=0 selects the first version.
    static const int V = getenv("FB_BSPMV_V") ? atoi(getenv("FB_BSPMV_V")) : 9;
    if (V == 0) return launch_bspmv<3, 32>(ctx, A, x, y, dot_mode, w, slot, flag);
    return launch_bspmv_u<3, 16, 4, 1>(ctx, A, x, y, dot_mode, w, slot, flag);
  }
  if (A.tile && A.tval) return tile_spmm(ctx, A, x, y, dot_mode, w, slot, flag);
  // scalar: pick lanes per row from the average row length
  switch (A.ncomp) {
    case 1:
      return launch_spmm<1, 8>(ctx, A, x, y, dot_mode, w, slot, flag);
    case 2:
      return launch_spmm<2, 8>(ctx, A, x, y, dot_mode, w, slot, flag);
    case 3: {
      // measured on B200 at n = 74 (3.3 M rows, 29 entries per row; profiles/r1_spmm3_variants.txt): T=16 0.559 ms,
      // T=8/U=2 0.423 ms, T=8/U=4 0.424 ms, T=4/U=8 0.474 ms; contiguous row ranges per block, paired 16-byte and
      // padded 32-byte gathers are all slower.  FB_SPMM_V=0 selects the first version.
      static const int V = getenv("FB_SPMM_V") ? atoi(getenv("FB_SPMM_V")) : 7;
      if (V == 0) return launch_spmm<3, 16>(ctx, A, x, y, dot_mode, w, slot, flag);
      return launch_spmm_u<3, 8, 2, 0>(ctx, A, x, y, dot_mode, w, slot, flag);
    }
    default:
      throw fb_cuda_error(FB_EINVAL, "spmv: ncomp must be 1..3");
  }
}

// =============================================================================
// vector kernels
// =============================================================================
__global__ void k_fill(double *x, double a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = a;
}
__global__ void k_axpy(double *y, double a, const double *__restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void k_to_float(float *__restrict__ dst, const double *__restrict__ src, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}
__global__ void k_axpby(double *z, double a, const double *x, double b, const double *y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = a * x[i] + b * y[i];
}
__global__ void k_dot(const double *__restrict__ x, const double *__restrict__ y, int64_t n, double *partials,
                      unsigned int *counter, double *red, int slot) {
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[0] += x[i] * y[i];
  fb_grid_reduce<1>(v, partials, counter, red, slot);
}
__global__ void k_mask_set(uint8_t *mask, const int64_t *__restrict__ dofs, int64_t nbc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbc; i += (int64_t)gridDim.x * blockDim.x)
    mask[dofs[i]] = 1;
}
__global__ void k_set_at(double *x, const int64_t *__restrict__ dofs, const double *__restrict__ vals, int64_t nbc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbc; i += (int64_t)gridDim.x * blockDim.x)
    x[dofs[i]] = vals ? vals[i] : 0.0;
}
__global__ void k_copy_at(double *dst, const double *__restrict__ src, const int64_t *__restrict__ dofs, int64_t nbc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbc; i += (int64_t)gridDim.x * blockDim.x)
    dst[dofs[i]] = src[dofs[i]];
}
__global__ void k_bc_residual(double *F, const double *__restrict__ x, const int64_t *__restrict__ dofs,
                              const double *__restrict__ vals, int64_t nbc) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbc; i += (int64_t)gridDim.x * blockDim.x)
    F[dofs[i]] = x[dofs[i]] - vals[i];
}

static inline int vgrid(fb_ctx *ctx, int64_t n) { return grid_for(n, 256, ctx->dev->sm_count * 8); }

void vec_fill(fb_ctx *ctx, double *x, double a, int64_t n) { FB_LAUNCH(ctx, k_fill, vgrid(ctx, n), 256, 0, x, a, n); }
void vec_axpy(fb_ctx *ctx, double *y, double a, const double *x, int64_t n) {
  FB_LAUNCH(ctx, k_axpy, vgrid(ctx, n), 256, 0, y, a, x, n);
}
void vec_to_float(fb_ctx *ctx, float *dst, const double *src, int64_t n) {
  FB_LAUNCH(ctx, k_to_float, vgrid(ctx, n), 256, 0, dst, src, n);
}
void vec_axpby(fb_ctx *ctx, double *z, double a, const double *x, double b, const double *y, int64_t n) {
  FB_LAUNCH(ctx, k_axpby, vgrid(ctx, n), 256, 0, z, a, x, b, y, n);
}
void vec_dot(fb_ctx *ctx, const double *x, const double *y, int64_t n, int slot) {
  fb_device_state *dv = ctx->dev;
  FB_LAUNCH(ctx, k_dot, vgrid(ctx, n), 256, 0, x, y, n, dv->partials, dv->counter, dv->red, slot);
  fb_allreduce_slots(ctx, slot, 1);
}
double vec_norm2_sync(fb_ctx *ctx, const double *x, int64_t n) {
  fb_device_state *dv = ctx->dev;
  vec_dot(ctx, x, x, n, FB_NSLOTS - 1);
  FB_CUDA(cudaMemcpyAsync(dv->host_pinned, dv->red + (FB_NSLOTS - 1), sizeof(double), cudaMemcpyDeviceToHost, dv->stream));
  FB_CUDA(cudaStreamSynchronize(dv->stream));
  return sqrt(dv->host_pinned[0]);
}
void mask_build(fb_ctx *ctx, uint8_t *mask, int64_t ndofs, const int64_t *dofs, int64_t nbc) {
  FB_CUDA(cudaMemsetAsync(mask, 0, ndofs, ctx->dev->stream));
  if (nbc > 0) FB_LAUNCH(ctx, k_mask_set, vgrid(ctx, nbc), 256, 0, mask, dofs, nbc);
}
void vec_set_at(fb_ctx *ctx, double *x, const int64_t *dofs, const double *vals, int64_t nbc) {
  if (nbc > 0) FB_LAUNCH(ctx, k_set_at, vgrid(ctx, nbc), 256, 0, x, dofs, vals, nbc);
}
void vec_zero_at(fb_ctx *ctx, double *x, const int64_t *dofs, int64_t nbc) {
  if (nbc > 0) FB_LAUNCH(ctx, k_set_at, vgrid(ctx, nbc), 256, 0, x, dofs, (const double *)nullptr, nbc);
}
void vec_copy_at(fb_ctx *ctx, double *dst, const double *src, const int64_t *dofs, int64_t nbc) {
  if (nbc > 0) FB_LAUNCH(ctx, k_copy_at, vgrid(ctx, nbc), 256, 0, dst, src, dofs, nbc);
}
// Lift the constrained unknowns of a system whose constrained rows are identity rows:
// xg = b on the constrained dofs (0 elsewhere), b <- b - A xg with b = 0 on constrained rows.
// Afterwards every Krylov vector vanishes on those dofs; add xg back to the solution.
void lift_identity_rows(fb_ctx *ctx, const LinOp &A, double *b, const int64_t *dofs, int64_t nbc, double *xg, double *tmp) {
  const int64_t n = A.ndofs();
  vec_fill(ctx, xg, 0.0, n);
  vec_copy_at(ctx, xg, b, dofs, nbc);
  spmv(ctx, A, xg, tmp);
  vec_axpy(ctx, b, -1.0, tmp, n);
  vec_zero_at(ctx, b, dofs, nbc);
}
void bc_residual(fb_ctx *ctx, double *F, const double *x, const int64_t *dofs, const double *vals, int64_t nbc) {
  if (nbc > 0) FB_LAUNCH(ctx, k_bc_residual, vgrid(ctx, nbc), 256, 0, F, x, dofs, vals, nbc);
}

// =============================================================================
// Dirichlet application on matrices  (DirichletBC.apply / assemble_system [EXT])
// =============================================================================
// Newton Jacobian rows -> identity (non-symmetric apply, pressure_correction.py:226)
template <int D>
__global__ void k_bc_rows_blocked(const int *__restrict__ rowptr, const int *__restrict__ diag, double *__restrict__ val,
                                  const int64_t *__restrict__ dofs, int64_t nbc) {
  // one warp per constrained dof
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < nbc; b += nwarps) {
    const int64_t dof = dofs[b];
    const int64_t I = dof / D;
    const int i = (int)(dof - I * D);
    const int r0 = rowptr[I];
    const int len = (rowptr[I + 1] - r0) * D;
    double *row = val + (int64_t)r0 * (D * D) + (int64_t)i * len;
    const int dpos = (diag[I] - r0) * D + i;
    for (int t = lane; t < len; t += 32) row[t] = (t == dpos) ? 1.0 : 0.0;
  }
}

void bc_rows_identity_blocked(fb_ctx *ctx, const DevSpace &sp, int D, double *val, const int64_t *dofs, int64_t nbc) {
  if (nbc <= 0) return;
  const int g = grid_for(nbc * 32, 256, ctx->dev->sm_count * 8);
  if (D == 2)
    FB_LAUNCH(ctx, k_bc_rows_blocked<2>, g, 256, 0, sp.rowptr.p, sp.diag.p, val, dofs, nbc);
  else
    FB_LAUNCH(ctx, k_bc_rows_blocked<3>, g, 256, 0, sp.rowptr.p, sp.diag.p, val, dofs, nbc);
}

// scalar matrix: SYM = 1 zero row + column with unit diagonal, SYM = 0 row only
template <int SYM>
__global__ void k_bc_scalar(int64_t n, const int *__restrict__ rowptr, const int *__restrict__ col,
                            double *__restrict__ val, const uint8_t *__restrict__ mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool mi = mask[i];
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int j = col[k];
      if (mi)
        val[k] = (j == (int)i) ? 1.0 : 0.0;
      else if (SYM && mask[j])
        val[k] = 0.0;
    }
  }
}

void bc_symmetric_scalar(fb_ctx *ctx, const DevSpace &sp, double *val, const uint8_t *mask) {
  FB_LAUNCH(ctx, k_bc_scalar<1>, vgrid(ctx, sp.nnodes), 256, 0, sp.nnodes, sp.rowptr.p, sp.col.p, val, mask);
}
void bc_rows_identity_scalar(fb_ctx *ctx, const DevSpace &sp, double *val, const uint8_t *mask) {
  FB_LAUNCH(ctx, k_bc_scalar<0>, vgrid(ctx, sp.nnodes), 256, 0, sp.nnodes, sp.rowptr.p, sp.col.p, val, mask);
}

// =============================================================================
// Jacobi / block-Jacobi setup
// =============================================================================
__global__ void k_jacobi_scalar(int64_t n, int ncomp, const int *__restrict__ diag, const double *__restrict__ val,
                                const uint8_t *__restrict__ mask, double *__restrict__ dinv) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * ncomp; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / ncomp;
    const double dd = val[diag[i]];
    dinv[t] = (mask && mask[t]) ? 1.0 : 1.0 / dd;
  }
}

void jacobi_setup_scalar(fb_ctx *ctx, const DevSpace &sp, const double *val, int ncomp, const uint8_t *mask, double *dinv) {
  FB_LAUNCH(ctx, k_jacobi_scalar, vgrid(ctx, sp.n_owned * ncomp), 256, 0, sp.n_owned, ncomp, sp.diag.p, val, mask, dinv);
}

template <int D>
__global__ void k_jacobi_blocked(int64_t n, const int *__restrict__ rowptr, const int *__restrict__ diag,
                                 const double *__restrict__ val, int block_mode, double *__restrict__ binv) {
  for (int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; I < n; I += (int64_t)gridDim.x * blockDim.x) {
    const int r0 = rowptr[I];
    const int len = (rowptr[I + 1] - r0) * D;
    const double *base = val + (int64_t)r0 * (D * D) + (int64_t)(diag[I] - r0) * D;
    double A[D][D], B[D][D];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) A[i][j] = base[(int64_t)i * len + j];
    if (block_mode == 0) {
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) B[i][j] = (i == j) ? 1.0 / A[i][i] : 0.0;
    } else if (D == 2) {
      const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
      const double inv = 1.0 / det;
      B[0][0] = A[1][1] * inv;
      B[0][1] = -A[0][1] * inv;
      B[1][0] = -A[1][0] * inv;
      B[1][1] = A[0][0] * inv;
    } else {
      const double c00 = A[1][1] * A[2 % D][2 % D] - A[1][2 % D] * A[2 % D][1];
      const double c01 = A[1][2 % D] * A[2 % D][0] - A[1][0] * A[2 % D][2 % D];
      const double c02 = A[1][0] * A[2 % D][1] - A[1][1] * A[2 % D][0];
      const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2 % D] * c02;
      const double inv = 1.0 / det;
      B[0][0] = c00 * inv;
      B[1][0] = c01 * inv;
      B[2 % D][0] = c02 * inv;
      B[0][1] = (A[0][2 % D] * A[2 % D][1] - A[0][1] * A[2 % D][2 % D]) * inv;
      B[1][1] = (A[0][0] * A[2 % D][2 % D] - A[0][2 % D] * A[2 % D][0]) * inv;
      B[2 % D][1] = (A[0][1] * A[2 % D][0] - A[0][0] * A[2 % D][1]) * inv;
      B[0][2 % D] = (A[0][1] * A[1][2 % D] - A[0][2 % D] * A[1][1]) * inv;
      B[1][2 % D] = (A[0][2 % D] * A[1][0] - A[0][0] * A[1][2 % D]) * inv;
      B[2 % D][2 % D] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * inv;
    }
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) binv[I * (D * D) + i * D + j] = B[i][j];
  }
}

void jacobi_setup_blocked(fb_ctx *ctx, const DevSpace &sp, int D, const double *val, int block_mode, double *binv) {
  const int g = vgrid(ctx, sp.n_owned);
  if (D == 2)
    FB_LAUNCH(ctx, k_jacobi_blocked<2>, g, 256, 0, sp.n_owned, sp.rowptr.p, sp.diag.p, val, block_mode, binv);
  else
    FB_LAUNCH(ctx, k_jacobi_blocked<3>, g, 256, 0, sp.n_owned, sp.rowptr.p, sp.diag.p, val, block_mode, binv);
}

// =============================================================================
// momentum residual and Jacobian: one warp per cell
// =============================================================================
constexpr int MOM_WARPS = 4;

// Shared-memory working set of the warp-per-cell momentum kernels.  Everything that is
// indexed with a lane-dependent index lives here (constant memory would serialise, register
// arrays would spill to local memory).
template <int D>
struct MomShared {
  static constexpr int NL = Elem<D>::NL2;
  static constexpr int NQ = Q5<D>::NQ;
  double phi[NQ][NL];       // P2 basis at the quadrature points (cell independent)
  double qlam[NQ][D + 1];   // quadrature points
  double qw[NQ];
  double glam[MOM_WARPS][D + 1][D];
  double U[MOM_WARPS][NL][D];
  double g[MOM_WARPS][NQ][NL][D];
  double uq[MOM_WARPS][NQ][D];
  double gu[MOM_WARPS][NQ][D][D];
  double p0q[MOM_WARPS][NQ];
};

template <int D>
__device__ __forceinline__ void mom_fill_tables(MomShared<D> &s) {
  constexpr int NL = Elem<D>::NL2, NQ = Q5<D>::NQ;
  for (int t = threadIdx.x; t < NQ * (D + 1); t += blockDim.x) s.qlam[t / (D + 1)][t % (D + 1)] = Q5<D>::lam(t / (D + 1), t % (D + 1));
  for (int t = threadIdx.x; t < NQ; t += blockDim.x) s.qw[t] = Q5<D>::w(t);
  __syncthreads();
  for (int t = threadIdx.x; t < NQ * NL; t += blockDim.x) {
    const int q = t / NL, a = t - q * NL;
    s.phi[q][a] = fb_p2_phi<D>(a, s.qlam[q]);
  }
  __syncthreads();
}

// geometry of the warp's cell: grad(lambda) to shared memory, returns the volume
template <int D>
__device__ __forceinline__ double mom_geometry(MomShared<D> &s, int wid, int lane, const int *__restrict__ cn,
                                               const double *__restrict__ xyz) {
  double X[(D + 1) * D];
#pragma unroll
  for (int v = 0; v <= D; ++v) {
    const int64_t node = cn[v];
#pragma unroll
    for (int k = 0; k < D; ++k) X[v * D + k] = xyz[node * D + k];
  }
  double glam[D + 1][D], vol;
  fb_geometry<D>(X, glam, vol);
  if (lane == 0) {
#pragma unroll
    for (int m = 0; m <= D; ++m)
#pragma unroll
      for (int k = 0; k < D; ++k) s.glam[wid][m][k] = glam[m][k];
  }
  __syncwarp();
  return vol;
}

// gather the cell's coefficients of `u`, tabulate basis gradients (first call per cell) and
// evaluate u and grad u at all quadrature points
template <int D>
__device__ __forceinline__ void mom_phase_a(MomShared<D> &s, int wid, int lane, const int *__restrict__ cn,
                                            const double *__restrict__ u, bool first, bool need_grad) {
  constexpr int NL = Elem<D>::NL2, NQ = Q5<D>::NQ;
  if (lane < NL) {
    const int64_t node = cn[lane];
#pragma unroll
    for (int i = 0; i < D; ++i) s.U[wid][lane][i] = u[node * D + i];
  }
  if (first) {
    for (int t = lane; t < NQ * NL; t += 32) {
      const int q = t / NL, a = t - q * NL;
      double g[D];
      fb_p2_grad<D>(a, s.qlam[q], s.glam[wid], g);
#pragma unroll
      for (int k = 0; k < D; ++k) s.g[wid][q][a][k] = g[k];
    }
  }
  __syncwarp();
  constexpr int PER_Q = D * (D + 1);
  for (int t = lane; t < NQ * PER_Q; t += 32) {
    const int q = t / PER_Q;
    const int r = t - q * PER_Q;
    const int i = r / (D + 1), kk = r - i * (D + 1);
    double acc = 0.0;
    if (kk == 0) {
#pragma unroll
      for (int a = 0; a < NL; ++a) acc += s.U[wid][a][i] * s.phi[q][a];
      s.uq[wid][q][i] = acc;
    } else if (need_grad) {
#pragma unroll
      for (int a = 0; a < NL; ++a) acc += s.U[wid][a][i] * s.g[wid][q][a][kk - 1];
      s.gu[wid][q][i][kk - 1] = acc;
    }
  }
  __syncwarp();
}

template <int D>
__device__ __forceinline__ void cell_geometry(const int *__restrict__ cn, const double *__restrict__ xyz,
                                              double glam[D + 1][D], double &vol) {
  double X[(D + 1) * D];
#pragma unroll
  for (int v = 0; v <= D; ++v) {
    const int64_t node = cn[v];
#pragma unroll
    for (int k = 0; k < D; ++k) X[v * D + k] = xyz[node * D + k];
  }
  fb_geometry<D>(X, glam, vol);
}

// Cell part of one state's contribution to F1 (pressure_correction.py:169-190):
//   F[(a,i)] += cm (u, phi_a e_i) - cr dt/rho R_cell(u; phi_a e_i)
// called with (ui, 1, theta) every Newton iteration and with (u0, -1, 1-theta) once per step.
template <int D>
__global__ void __launch_bounds__(MOM_WARPS * 32)
    k_momentum_F(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells, const double *__restrict__ xyz, double dt, double rho,
                 double mu, const double *__restrict__ u, const double *__restrict__ p0, const int *__restrict__ pcn, double cm,
                 double cr, double *__restrict__ F) {
  constexpr int NL = Elem<D>::NL2, NQ = Q5<D>::NQ;
  __shared__ MomShared<D> s;
  mom_fill_tables<D>(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp0 = blockIdx.x * (int64_t)MOM_WARPS + wid;
  const int64_t nwarps = (int64_t)gridDim.x * MOM_WARPS;
  const double cdt = cr * dt / rho;
  const bool need_R = (cr != 0.0);
  for (int64_t c = warp0; c < nc; c += nwarps) {
    const int *cn = cell_nodes + c * NL;
    const double vol = mom_geometry<D>(s, wid, lane, cells + c * (D + 1), xyz);
    if (need_R && lane < NQ) {
      double p = 0.0;
#pragma unroll
      for (int v = 0; v <= D; ++v) p += p0[pcn[c * (D + 1) + v]] * s.qlam[lane][v];
      s.p0q[wid][lane] = p;
    }
    mom_phase_a<D>(s, wid, lane, cn, u, need_R, need_R);
    if (lane < NL * D) {
      const int ta = lane / D, ti = lane - ta * D;
      double acc = 0.0;
      for (int q = 0; q < NQ; ++q) {
        const double w = s.qw[q] * vol;
        const double pa = s.phi[q][ta];
        acc += cm * w * pa * s.uq[wid][q][ti];
        if (need_R)  // lane-dependent (a, i): index the shared tables directly
          acc -= cdt * w * fb_rhs_point<D>(ti, rho, mu, pa, s.g[wid][q][ta], s.uq[wid][q], s.gu[wid][q], s.p0q[wid][q]);
      }
      atomicAdd(&F[(int64_t)cn[ta] * D + ti], acc);
    }
    __syncwarp();
  }
}

// boundary-facet part of F: -dt/rho [ -(p0 n, v)_ds + mu ((grad u)^T n, v)_ds ]  (pressure_correction.py:142-143)
// evaluated for the blended state theta*ui + (1-theta)*u0 (the facet terms are linear in u)
template <int D>
__global__ void k_momentum_F_facets(int64_t nbf, const int *__restrict__ bf_cell, const int *__restrict__ bf_local,
                                    const int *__restrict__ cell_nodes, const int *__restrict__ cells, const double *__restrict__ xyz, MomentumArgs a,
                                    double *__restrict__ F) {
  constexpr int NL = Elem<D>::NL2;
  const int64_t total = nbf * NL * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fidx = t / (NL * D);
    const int r = (int)(t - fidx * NL * D);
    const int ta = r / D, ti = r - ta * D;
    const int64_t c = bf_cell[fidx];
    const int f = bf_local[fidx];
    const int *cn = cell_nodes + c * NL;
    if (!fb_node_on_facet<D>(ta, f)) continue;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double Ue[NL * D], p0e[D + 1];
    for (int b = 0; b < NL; ++b) {
      const int64_t node = cn[b];
      for (int k = 0; k < D; ++k)
        Ue[b * D + k] = a.theta * a.ui[node * D + k] + (1.0 - a.theta) * a.u0[node * D + k];
    }
    for (int v = 0; v <= D; ++v) p0e[v] = a.p0[a.pcn[c * (D + 1) + v]];
    const double acc = fb_facet_F<D>(ta, ti, f, glam, vol, QF<D>::lam_ptr(), QF<D>::w_ptr(), QF<D>::NQ, Ue, p0e, a.mu);
    atomicAdd(&F[(int64_t)cn[ta] * D + ti], -(a.dt / a.rho) * acc);
  }
}

// Same contribution, one THREAD per cell with everything in registers: the warp-per-cell version above
// is bound by shared-memory bandwidth (ncu: L1/shared pipe 90 % busy, 9.6 ms at 2.43 M cells); here the
// basis functions are recomputed from the barycentric point (a few FMAs each) instead of being staged.
template <int D>
__global__ void __launch_bounds__(128)
    k_momentum_F_thread(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                        const double *__restrict__ xyz, double dt, double rho, double mu, const double *__restrict__ u,
                        const double *__restrict__ p0, const int *__restrict__ pcn, double cm, double cr,
                        double *__restrict__ F, const double *__restrict__ adv) {
  constexpr int NL = Elem<D>::NL2, NQ = Q5<D>::NQ;
  const double cdt = cr * dt / rho;
  const bool need_R = (cr != 0.0);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double U[NL][D], acc[NL][D], p0e[D + 1];
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      const int64_t node = cn[a];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        U[a][i] = u[node * D + i];
        acc[a][i] = 0.0;
      }
    }
#pragma unroll
    for (int v = 0; v <= D; ++v) p0e[v] = need_R ? p0[pcn[c * (D + 1) + v]] : 0.0;
    for (int q = 0; q < NQ; ++q) {
      double lam[D + 1];
#pragma unroll
      for (int m = 0; m <= D; ++m) lam[m] = Q5<D>::lam(q, m);
      const double w = Q5<D>::w(q) * vol;
      double uq[D], gu[D][D], pq = 0.0, wq[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        uq[i] = 0.0;
        wq[i] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) gu[i][k] = 0.0;
      }
      if (adv && need_R) {  // advecting velocity of the semi-implicit linearisation at the point
#pragma unroll
        for (int a = 0; a < NL; ++a) {
          const double pa = fb_p2_phi<D>(a, lam);
          const int64_t node = cn[a];
#pragma unroll
          for (int i = 0; i < D; ++i) wq[i] += adv[node * D + i] * pa;
        }
      }
#pragma unroll
      for (int v = 0; v <= D; ++v) pq += p0e[v] * lam[v];
#pragma unroll
      for (int a = 0; a < NL; ++a) {
        const double pa = fb_p2_phi<D>(a, lam);
        double ga[D];
        if (need_R) fb_p2_grad<D>(a, lam, glam, ga);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          uq[i] += U[a][i] * pa;
          if (need_R) {
#pragma unroll
            for (int k = 0; k < D; ++k) gu[i][k] += U[a][i] * ga[k];
          }
        }
      }
      // point quantities: the integrand of (a, i) is  phi_a X_i + sum_k d_k phi_a Z_ik  with
      //   X_i  = w [ cm u_i + cdt rho/2 ((grad u) b)_i ],
      //   Z_ik = -w cdt [ rho/2 b_k u_i - mu (d_k u_i + d_i u_k) + p0 delta_ik ]        (b: advecting velocity)
      // (fb_rhs_point written out, with everything that does not depend on the test node hoisted: 4 FMAs per (a, i))
      double X[D], Z[D][D];
      {
        const double *bq = (adv && need_R) ? wq : uq;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double conv = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) conv += gu[i][k] * bq[k];
          X[i] = w * (cm * uq[i] + (need_R ? cdt * 0.5 * rho * conv : 0.0));
#pragma unroll
          for (int k = 0; k < D; ++k)
            Z[i][k] = need_R ? -w * cdt * (0.5 * rho * bq[k] * uq[i] - mu * (gu[i][k] + gu[k][i]) + (i == k ? pq : 0.0)) : 0.0;
        }
      }
#pragma unroll
      for (int a = 0; a < NL; ++a) {
        const double pa = fb_p2_phi<D>(a, lam);
        double ga[D];
        if (need_R) fb_p2_grad<D>(a, lam, glam, ga);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          double v = pa * X[i];
          if (need_R) {
#pragma unroll
            for (int k = 0; k < D; ++k) v += ga[k] * Z[i][k];
          }
          acc[a][i] += v;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      const int64_t node = cn[a];
#pragma unroll
      for (int i = 0; i < D; ++i) atomicAdd(&F[node * D + i], acc[a][i]);
    }
  }
}

template <int D>
static void momentum_F_cells(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, const double *u, double cm, double cr,
                             double *F, const double *adv = nullptr) {
  static const int variant = getenv("FB_F_KERNEL") ? atoi(getenv("FB_F_KERNEL")) : 1;  // 0: warp per cell, 1: thread per cell
  if (variant == 0 && !adv) {
    const int g = grid_for(W.nc * 32, MOM_WARPS * 32, ctx->dev->sm_count * 16);
    FB_LAUNCH(ctx, k_momentum_F<D>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, a.dt, a.rho, a.mu, u, a.p0, a.pcn, cm, cr, F);
  } else {
    const int g = grid_for(W.nc, 128, ctx->dev->sm_count * 16);
    FB_LAUNCH(ctx, k_momentum_F_thread<D>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, a.dt, a.rho, a.mu, u, a.p0, a.pcn, cm, cr, F, adv);
  }
}

// F += state-u0 part: -(u0, v) - dt/rho (1-theta) R_cell(u0; v).  For backward Euler this is -M u0 and
// the caller uses the assembled mass matrix instead.
void assemble_momentum_F_old_state(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *F) {
  if (W.dim == 2)
    momentum_F_cells<2>(ctx, W, a, a.u0, -1.0, 1.0 - a.theta, F);
  else
    momentum_F_cells<3>(ctx, W, a, a.u0, -1.0, 1.0 - a.theta, F);
}

// F += (ui, v) - dt/rho theta R_cell(ui; v) + boundary-facet terms of the blended state
void assemble_momentum_F_new_state(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *F) {
  const int gf = grid_for(W.nbf * W.nl * W.dim, 128, ctx->dev->sm_count * 16);
  if (W.dim == 2) {
    momentum_F_cells<2>(ctx, W, a, a.ui, 1.0, a.theta, F, a.adv);
    if (W.nbf && !a.skip_facets) FB_LAUNCH(ctx, k_momentum_F_facets<2>, gf, 128, 0, W.nbf, W.bf_cell.p, W.bf_local.p, W.cell_nodes.p, W.cells.p, W.xyz.p, a, F);
  } else {
    momentum_F_cells<3>(ctx, W, a, a.ui, 1.0, a.theta, F, a.adv);
    if (W.nbf && !a.skip_facets) FB_LAUNCH(ctx, k_momentum_F_facets<3>, gf, 128, 0, W.nbf, W.bf_cell.p, W.bf_local.p, W.cell_nodes.p, W.cells.p, W.xyz.p, a, F);
  }
}

void assemble_momentum_F(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *F) {
  FB_CUDA(cudaMemsetAsync(F, 0, sizeof(double) * W.nnodes * W.dim, ctx->dev->stream));
  assemble_momentum_F_old_state(ctx, W, a, F);
  assemble_momentum_F_new_state(ctx, W, a, F);
}

template <int D>
__global__ void __launch_bounds__(MOM_WARPS * 32)
    k_momentum_J(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells, const double *__restrict__ xyz,
                 const int *__restrict__ rowptr, const int *__restrict__ smap, MomentumArgs a, double *__restrict__ val) {
  constexpr int NL = Elem<D>::NL2, NQ = Q5<D>::NQ, NP = NL * NL, R = (NP + 31) / 32;
  __shared__ MomShared<D> s;
  mom_fill_tables<D>(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp0 = blockIdx.x * (int64_t)MOM_WARPS + wid;
  const int64_t nwarps = (int64_t)gridDim.x * MOM_WARPS;
  const double c1 = 0.5 * a.theta * a.dt, c2 = a.theta * a.dt * a.mu / a.rho;
  for (int64_t c = warp0; c < nc; c += nwarps) {
    const int *cn = cell_nodes + c * NL;
    const double vol = mom_geometry<D>(s, wid, lane, cells + c * (D + 1), xyz);
    mom_phase_a<D>(s, wid, lane, cn, a.adv ? a.adv : a.ui, true, true);
    double acc[R][D][D];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) acc[r][i][j] = 0.0;
    for (int q = 0; q < NQ; ++q) {
      const double w = s.qw[q] * vol;
      double u[D], gu[D][D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        u[k] = s.uq[wid][q][k];
#pragma unroll
        for (int l = 0; l < D; ++l) gu[k][l] = s.gu[wid][q][k][l];
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int p = lane + 32 * r;
        if (p < NP) {
          const int pa_i = p / NL, pb_i = p - pa_i * NL;
          double ga[D], gb[D];
#pragma unroll
          for (int k = 0; k < D; ++k) {
            ga[k] = s.g[wid][q][pa_i][k];
            gb[k] = s.g[wid][q][pb_i][k];
          }
          fb_jac_point<D>(w, c1, c2, s.phi[q][pa_i], s.phi[q][pb_i], ga, gb, u, gu, acc[r], a.adv ? 0.0 : 1.0);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int p = lane + 32 * r;
      if (p < NP) {
        const int pa_i = p / NL;
        const int I = cn[pa_i];
        const int r0 = rowptr[I];
        const int len = (rowptr[I + 1] - r0) * D;
        const int slot = smap[c * NP + p];
        double *base = val + (int64_t)r0 * (D * D) + (int64_t)(slot - r0) * D;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) atomicAdd(base + (int64_t)i * len + j, acc[r][i][j]);
      }
    }
    __syncwarp();
  }
}

// Closed-form element Jacobians (fb_jac_pair, fb_element.cuh): one warp per cell.
//   phase A (cooperative, shared memory): nodal velocities U, the vertex values GV of the affine basis gradients
//            and their sums S, the vertex values GU of grad u, the moments WT[a][w][k] = sum_c M3[a][c][w] U_c[k];
//   phase B: lane = (test node a, group g); the a-dependent operands stay in registers, the lane walks
//            the trial nodes b = g, g + NG, ... and scatters each D x D block with fp64 atomics through smap.
// ~160 fused multiply-adds per block in 3D against ~700 of the quadrature kernel k_momentum_J, no quadrature
// tables, no hand-off buffer from the residual kernel.
template <int D>
struct JcfShared {
  static constexpr int NL = Elem<D>::NL2, NV = D + 1;
  double M3[NL * NL * NV];
  double U[MOM_WARPS][NL][D];
  double GV[MOM_WARPS][NL][NV * D];
  double S[MOM_WARPS][NL][D];
  double WT[MOM_WARPS][NL][NV * D];
  double GU[MOM_WARPS][NV * D * D];
};

// ELEM = 0: scatter-add into the matrix (fp64 atomics through smap); ELEM = 1: store the element blocks cell by cell
// (val = element buffer, nc * NL * NL * D * D doubles) for the deterministic gather pass k_jac_gather
template <int D, int ELEM = 0, int MINB = 1>
__global__ void __launch_bounds__(MOM_WARPS * 32, MINB)
    k_momentum_J_cf(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                    const double *__restrict__ xyz, const int *__restrict__ rowptr, const int *__restrict__ smap,
                    MomentumArgs a, double *__restrict__ val) {
  constexpr int NL = Elem<D>::NL2, NV = D + 1, NP = NL * NL;
  constexpr int NG = 32 / NL;               // lanes per test node (3 in 3D, 5 in 2D); 32 - NG * NL lanes idle in phase B
  constexpr int ROUNDS = (NL + NG - 1) / NG;
  __shared__ JcfShared<D> s;
  {
    const double *tab = (D == 2) ? FB_M3_TRI : FB_M3_TET;
    for (int t = threadIdx.x; t < NL * NL * NV; t += blockDim.x) s.M3[t] = tab[t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp0 = blockIdx.x * (int64_t)MOM_WARPS + wid;
  const int64_t nwarps = (int64_t)gridDim.x * MOM_WARPS;
  const double c1 = 0.5 * a.theta * a.dt, c2 = a.theta * a.dt * a.mu / a.rho;
  const int ta = lane % NL, grp = lane / NL;
  const bool active = lane < NG * NL;
  for (int64_t c = warp0; c < nc; c += nwarps) {
    const int *cn = cell_nodes + c * NL;
    // scatter addresses first: their load latency hides behind phase A
    int r0 = 0, len = 0;
    if (active && !ELEM) {
      const int I = cn[ta];
      r0 = rowptr[I];
      len = (rowptr[I + 1] - r0) * D;
    }
    // ---- phase A
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);  // every lane: broadcast loads, no staging
    // semi-implicit linearisation: the Jacobian only sees the advecting field (and drops the terms that
    // differentiate it: c1t = 0)
    if (lane < NL * D) s.U[wid][lane / D][lane % D] = (a.adv ? a.adv : a.ui)[(int64_t)cn[lane / D] * D + lane % D];
    for (int t = lane; t < NL * NV; t += 32) {
      const int n = t / NV, w = t - n * NV;
      double g[D];
      fb_p2_vertex_grad<D>(n, w, glam, g);
#pragma unroll
      for (int k = 0; k < D; ++k) s.GV[wid][n][w * D + k] = g[k];
    }
    __syncwarp();
    if (lane < NL * D) {
      const int n = lane / D, k = lane % D;
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NV; ++w) sum += s.GV[wid][n][w * D + k];
      s.S[wid][n][k] = sum;
    }
    for (int t = lane; t < NV * D * D; t += 32) {  // GU[v][i][j] = sum_c U_c[i] GV_c[v][j]
      const int v = t / (D * D), i = (t / D) % D, j = t % D;
      double sum = 0.0;
#pragma unroll
      for (int cc = 0; cc < NL; ++cc) sum += s.U[wid][cc][i] * s.GV[wid][cc][v * D + j];
      s.GU[wid][t] = sum;
    }
    for (int t = lane; t < NL * NV * D; t += 32) {  // WT[n][w][k] = sum_c M3[n][c][w] U_c[k]
      const int n = t / (NV * D), w = (t / D) % NV, k = t % D;
      double sum = 0.0;
#pragma unroll
      for (int cc = 0; cc < NL; ++cc) sum += s.M3[(n * NL + cc) * NV + w] * s.U[wid][cc][k];
      s.WT[wid][n][w * D + k] = sum;
    }
    __syncwarp();
    // ---- phase B
    if (active) {
      double GA[NV * D], SA[D], WA[NV * D];  // GU is read from shared memory: one address per warp -> broadcast
#pragma unroll
      for (int t = 0; t < NV * D; ++t) {
        GA[t] = s.GV[wid][ta][t];
        WA[t] = s.WT[wid][ta][t];
      }
#pragma unroll
      for (int k = 0; k < D; ++k) SA[k] = s.S[wid][ta][k];
#pragma unroll 1
      for (int r = 0; r < ROUNDS; ++r) {
        const int tb = grp + r * NG;
        if (tb < NL) {
          const int slot = ELEM ? 0 : smap[c * NP + ta * NL + tb];  // issued before the block's arithmetic
          double J[D][D];
          fb_jac_pair<D>(vol, c1, c2, GA, SA, s.GV[wid][tb], s.S[wid][tb], WA, s.WT[wid][tb], s.GU[wid], &s.M3[(ta * NL + tb) * NV], J,
                         a.adv ? 0.0 : c1);
          if (ELEM) {
            double *dst = val + ((int64_t)c * NP + ta * NL + tb) * (D * D);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
              for (int j = 0; j < D; ++j) __stcs(dst + i * D + j, J[i][j]);
          } else {
            double *base = val + (int64_t)r0 * (D * D) + (int64_t)(slot - r0) * D;
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
              for (int j = 0; j < D; ++j) atomicAdd(base + (int64_t)i * len + j, J[i][j]);
          }
        }
      }
    }
    __syncwarp();
  }
}

// Same closed form, other work split (experiment, FB_J_KERNEL=3): lane = (test node a, column j) -- NL * D lanes --
// walks ALL trial nodes b and produces column j of every block (a, b).  The lane keeps only the j-slices of the
// test-side operands in registers (GA[.][j], WA[.][j], S_a[j], GU[.][.][j]: 21 doubles instead of 27 + the block), every
// trial-side operand is one address per warp (broadcast), M3 is read through a transposed copy (conflict free).  The
// scalar part of the block (mass + skew convection + trace of the viscous part) needs the sum over the D lanes of a test
// node: one quantity, two shuffles.
template <int D>
struct Jcf2Shared {
  static constexpr int NL = Elem<D>::NL2, NV = D + 1;
  double M3T[NL * NV * NL];  // [(b * NV + v) * NL + a]
  double U[MOM_WARPS][NL][D];
  double GV[MOM_WARPS][NL][NV * D];
  double S[MOM_WARPS][NL][D];
  double WT[MOM_WARPS][NL][NV * D];
  double GU[MOM_WARPS][NV * D * D];
};

template <int D>
__global__ void __launch_bounds__(MOM_WARPS * 32)
    k_momentum_J_cf2(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                     const double *__restrict__ xyz, const int *__restrict__ rowptr, const int *__restrict__ smap,
                     MomentumArgs a, double *__restrict__ val) {
  constexpr int NL = Elem<D>::NL2, NV = D + 1, NP = NL * NL;
  static_assert(NL * D <= 32, "one lane per (test node, column)");
  __shared__ Jcf2Shared<D> s;
  const double *tab = (D == 2) ? FB_M3_TRI : FB_M3_TET;
  for (int t = threadIdx.x; t < NL * NL * NV; t += blockDim.x) {
    const int ta = t / (NL * NV), r = t - ta * (NL * NV), tb = r / NV, v = r - tb * NV;
    s.M3T[(tb * NV + v) * NL + ta] = tab[t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp0 = blockIdx.x * (int64_t)MOM_WARPS + wid;
  const int64_t nwarps = (int64_t)gridDim.x * MOM_WARPS;
  const double c1 = 0.5 * a.theta * a.dt, c2 = a.theta * a.dt * a.mu / a.rho;
  const double c1t = a.adv ? 0.0 : c1;
  const double lm = 1.0 / ((D + 1) * (D + 2));
  const bool active = lane < NL * D;
  const int ta = active ? lane / D : 0, tj = active ? lane % D : 0;
  const int lane0 = ta * D;  // first lane of this test node
  for (int64_t c = warp0; c < nc; c += nwarps) {
    const int *cn = cell_nodes + c * NL;
    int r0 = 0, len = 0;
    if (active) {
      const int I = cn[ta];
      r0 = rowptr[I];
      len = (rowptr[I + 1] - r0) * D;
    }
    // ---- phase A (as k_momentum_J_cf)
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    if (lane < NL * D) s.U[wid][lane / D][lane % D] = (a.adv ? a.adv : a.ui)[(int64_t)cn[lane / D] * D + lane % D];
    for (int t = lane; t < NL * NV; t += 32) {
      const int n = t / NV, w = t - n * NV;
      double g[D];
      fb_p2_vertex_grad<D>(n, w, glam, g);
#pragma unroll
      for (int k = 0; k < D; ++k) s.GV[wid][n][w * D + k] = g[k];
    }
    __syncwarp();
    if (lane < NL * D) {
      const int n = lane / D, k = lane % D;
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NV; ++w) sum += s.GV[wid][n][w * D + k];
      s.S[wid][n][k] = sum;
    }
    for (int t = lane; t < NV * D * D; t += 32) {
      const int v = t / (D * D), i = (t / D) % D, j = t % D;
      double sum = 0.0;
#pragma unroll
      for (int cc = 0; cc < NL; ++cc) sum += s.U[wid][cc][i] * s.GV[wid][cc][v * D + j];
      s.GU[wid][t] = sum;
    }
    for (int t = lane; t < NL * NV * D; t += 32) {
      const int n = t / (NV * D), w = (t / D) % NV, k = t % D;
      double sum = 0.0;
#pragma unroll
      for (int cc = 0; cc < NL; ++cc) sum += s.M3T[(cc * NV + w) * NL + n] * s.U[wid][cc][k];  // M3[n][cc][w]
      s.WT[wid][n][w * D + k] = sum;
    }
    __syncwarp();
    // ---- phase B
    double GAj[NV], WAj[NV], GUj[NV][D], SAj;
#pragma unroll
    for (int w = 0; w < NV; ++w) {
      GAj[w] = s.GV[wid][ta][w * D + tj];
      WAj[w] = s.WT[wid][ta][w * D + tj];
#pragma unroll
      for (int i = 0; i < D; ++i) GUj[w][i] = s.GU[wid][(w * D + i) * D + tj];
    }
    SAj = s.S[wid][ta][tj];
#pragma unroll 2
    for (int tb = 0; tb < NL; ++tb) {
      const int slot = active ? smap[c * NP + ta * NL + tb] : 0;
      double m3[NV], mab = 0.0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        m3[v] = s.M3T[(tb * NV + v) * NL + ta];
        mab += m3[v];
      }
      const double *GB = s.GV[wid][tb];  // [v * D + i], one address per warp
      const double *WB = s.WT[wid][tb];
      double Jc[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double g = s.S[wid][tb][i] * SAj, t2 = 0.0, t3 = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          g += GB[v * D + i] * GAj[v];
          t2 += m3[v] * GUj[v][i];
          t3 += GAj[v] * WB[v * D + i];
        }
        Jc[i] = c1t * (t2 - t3) + c2 * lm * g;
      }
      // this lane's share of the scalar part: c1 (C_ab - C_ba) + c2 tr G, restricted to k = j
      double q = s.S[wid][tb][tj] * SAj, cab = 0.0, cba = 0.0;
#pragma unroll
      for (int w = 0; w < NV; ++w) {
        const double gbj = GB[w * D + tj], wbj = WB[w * D + tj];
        q += gbj * GAj[w];
        cab += gbj * WAj[w];
        cba += GAj[w] * wbj;
      }
      q = c1 * (cab - cba) + c2 * lm * q;
      double dg = q;
#pragma unroll
      for (int k = 1; k < D; ++k) dg += __shfl_sync(0xffffffffu, q, lane0 + (tj + k) % D);
      dg += mab;
      if (active) {
        double *base = val + (int64_t)r0 * (D * D) + (int64_t)(slot - r0) * D + tj;
#pragma unroll
        for (int i = 0; i < D; ++i) atomicAdd(base + (int64_t)i * len, vol * (Jc[i] + (i == tj ? dg : 0.0)));
      }
    }
    __syncwarp();
  }
}

// Column-per-lane closed form with merged trial-side operands (FB_J_KERNEL=5).
// ncu of k_momentum_J_cf2 (profiles/r2_k_momentum_J_cf2_ncu_details.txt): bound by the shared-memory data pipe (91 %,
// FP64 33 %) -- 410 LDS instructions and 825 wavefronts per cell: a 16-byte broadcast load (which the compiler forms from
// neighbouring operands) costs 4 wavefronts, an 8-byte broadcast costs 1.  Here
//   * the trial-side operands of a block enter through ONE array E[b][v][i] = c2 lm GB[v][i] - c1t WB[v][i]
//     (v = NV: c2 lm SB[i]), written in phase A: the column is  J[i] = sum_v m3[v] (c1t GU[v][i][j]) + sum_v GA[v][j] E[b][v][i],
//     and, since the scalar part shares the same sums, q_j = c1 sum_w GB[w][j] WA[w][j] + (second sum at i = j)
//     (plus - c1 sum_w GA[w][j] WB[w][j] in the semi-implicit variant, where c1t = 0);
//   * every shared-memory operand is read by an explicit 8-byte load (lds64): 23 loads and ~36 FMAs per (lane, trial node)
//     against 30 loads (13 of them 16-byte) and 62 FMAs;
//   * the cell geometry is staged in shared memory (the dynamic indexing of fb_p2_vertex_grad went to local memory).
__device__ __forceinline__ double lds64(const double *p) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}

template <int D>
struct Jcf3Shared {
  static constexpr int NL = Elem<D>::NL2, NV = D + 1, NVP = 4;
  double M3T[NL * NV * NL];  // [(b * NV + v) * NL + a]
  double X[MOM_WARPS][(D + 1) * D];
  double GL[MOM_WARPS][D + 1][D];
  double U[MOM_WARPS][NL][D];
  double GV[MOM_WARPS][NL][NV * D];
  double GVT[MOM_WARPS][NL][D][NVP];  // GVT[n][k][w] = GV[n][w * D + k]
  double WTT[MOM_WARPS][NL][D][NVP];  // WTT[n][k][w] = (1/|K|) int lambda_w phi_n u_k
  double E[MOM_WARPS][NL][NV + 1][D];
  double S[MOM_WARPS][NL][D];
  double GU[MOM_WARPS][NV * D * D];
};

// The global operands of a cell (vertex ids -> coordinates, node ids -> velocities and row extents: two dependent loads
// each) are fetched one cell ahead and held in registers, one value per lane: in the first version 16 % of all warp
// samples waited for them at the top of a cell (ncu source page, long scoreboard) with only 4 warps per scheduler to
// hide it.  The cell's row of the scatter map is read at the top as well (NL ints per lane), not slot by slot.
template <int D>
__global__ void __launch_bounds__(MOM_WARPS * 32)
    k_momentum_J_cf3(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                     const double *__restrict__ xyz, const int *__restrict__ rowptr, const int *__restrict__ smap,
                     MomentumArgs a, double *__restrict__ val) {
  using SH = Jcf3Shared<D>;
  constexpr int NL = Elem<D>::NL2, NV = D + 1, NP = NL * NL;
  static_assert(NL * D <= 32, "one lane per (test node, column)");
  static_assert(NL % 2 == 0, "scatter-map rows are read as int2");
  __shared__ SH s;
  const double *tab = (D == 2) ? FB_M3_TRI : FB_M3_TET;
  for (int t = threadIdx.x; t < NL * NL * NV; t += blockDim.x) {
    const int ta = t / (NL * NV), r = t - ta * (NL * NV), tb = r / NV, v = r - tb * NV;
    s.M3T[(tb * NV + v) * NL + ta] = tab[t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp0 = blockIdx.x * (int64_t)MOM_WARPS + wid;
  const int64_t nwarps = (int64_t)gridDim.x * MOM_WARPS;
  const double c1 = 0.5 * a.theta * a.dt, c2 = a.theta * a.dt * a.mu / a.rho;
  const bool semi = a.adv != nullptr;
  const double *usrc = semi ? a.adv : a.ui;
  const double c1t = semi ? 0.0 : c1;
  const double c2lm = c2 / ((D + 1) * (D + 2));
  const bool active = lane < NL * D;
  const int ta = active ? lane / D : 0, tj = active ? lane % D : 0;
  const int lane0 = ta * D;  // first lane of this test node
  const bool is_vertex = ta < NV;
  const int pn = is_vertex ? ta : edge_v<D>(ta - NV, 0), qn = is_vertex ? ta : edge_v<D>(ta - NV, 1);
  // ---- software pipeline over the cells of this warp
  int n_vid = 0, n_node = 0, n_r0 = 0, n_r1 = 0;
  double n_x = 0.0, n_u = 0.0;
  auto fetch_ids = [&](int64_t cc) {
    if (cc < nc) {
      if (lane < D + 1) n_vid = cells[cc * (D + 1) + lane];
      if (active) n_node = cell_nodes[cc * NL + ta];
    }
  };
  auto fetch_values = [&](int64_t cc) {
    if (cc < nc) {  // warp-uniform
      const int vid = __shfl_sync(0xffffffffu, n_vid, (lane / D) % (D + 1));
      if (lane < (D + 1) * D) n_x = xyz[(int64_t)vid * D + lane % D];
      if (active) {
        n_u = usrc[(int64_t)n_node * D + tj];
        n_r0 = rowptr[n_node];
        n_r1 = rowptr[n_node + 1];
      }
    }
  };
  fetch_ids(warp0);
  fetch_values(warp0);
  for (int64_t c = warp0; c < nc; c += nwarps) {
    const int r0 = n_r0, len = (n_r1 - n_r0) * D;
    if (lane < (D + 1) * D) s.X[wid][lane] = n_x;
    if (active) s.U[wid][ta][tj] = n_u;
    int slot[NL];
    {
      const int2 *row = reinterpret_cast<const int2 *>(smap + c * NP + ta * NL);
#pragma unroll
      for (int h = 0; h < NL / 2; ++h) {
        const int2 v2 = row[h];
        slot[2 * h] = v2.x;
        slot[2 * h + 1] = v2.y;
      }
    }
    fetch_ids(c + nwarps);
    __syncwarp();
    // ---- phase A
    double vol;
    {
      double X[(D + 1) * D], glam[D + 1][D];
#pragma unroll
      for (int t = 0; t < (D + 1) * D; ++t) X[t] = lds64(&s.X[wid][t]);
      fb_geometry<D>(X, glam, vol);  // every lane; one lane stages the gradients (dynamic indexing below)
      if (lane == 0) {
#pragma unroll
        for (int v = 0; v < D + 1; ++v)
#pragma unroll
          for (int k = 0; k < D; ++k) s.GL[wid][v][k] = glam[v][k];
      }
    }
    __syncwarp();
    // vertex values of grad phi_n (fb_p2_vertex_grad), lane = (n, k): A_w glam[pn][k] + B_w glam[qn][k] with
    // (A_w, B_w) = (3 or -1, 0) for a vertex node n = pn = qn, (4 [w == qn], 4 [w == pn]) for the edge node (pn, qn)
    double gw[NV];
    {
      const double Gp = lds64(&s.GL[wid][pn][tj]), Gq = lds64(&s.GL[wid][qn][tj]);
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < NV; ++w) {
        const double Aw = is_vertex ? (w == pn ? 3.0 : -1.0) : (w == qn ? 4.0 : 0.0);
        const double Bw = is_vertex ? 0.0 : (w == pn ? 4.0 : 0.0);
        gw[w] = Aw * Gp + Bw * Gq;
        sum += gw[w];
        if (active) {
          s.GV[wid][ta][w * D + tj] = gw[w];
          s.GVT[wid][ta][tj][w] = gw[w];
        }
      }
      if (active) {
        s.S[wid][ta][tj] = sum;
        s.E[wid][ta][NV][tj] = c2lm * sum;
      }
    }
    __syncwarp();
    for (int t = lane; t < NV * D * D; t += 32) {  // GU[v][i][j] = sum_c U_c[i] GV_c[v][j], scaled by c1t
      const int v = t / (D * D), i = (t / D) % D, j = t % D;
      double sum = 0.0;
#pragma unroll
      for (int cc = 0; cc < NL; ++cc) sum += s.U[wid][cc][i] * s.GV[wid][cc][v * D + j];
      s.GU[wid][t] = c1t * sum;
    }
    {  // WT[n][w][k] = sum_c M3[n][c][w] U_c[k], lane = (n, k): the U_c[k] stay in registers over the NV moments
      double Uk[NL];
#pragma unroll
      for (int cc = 0; cc < NL; ++cc) Uk[cc] = lds64(&s.U[wid][cc][tj]);
#pragma unroll
      for (int w = 0; w < NV; ++w) {
        double sum = 0.0;
#pragma unroll
        for (int cc = 0; cc < NL; ++cc) sum += lds64(&s.M3T[(cc * NV + w) * NL + ta]) * Uk[cc];
        if (active) {
          s.WTT[wid][ta][tj][w] = sum;
          s.E[wid][ta][w][tj] = c2lm * gw[w] - c1t * sum;
        }
      }
    }
    __syncwarp();
    fetch_values(c + nwarps);  // lands during phase B
    // ---- phase B: test-side operands of this lane (column tj of test node ta)
    double GAj[NV], WAj[NV], GUj[NV][D], SAj;
#pragma unroll
    for (int w = 0; w < NV; ++w) {
      GAj[w] = lds64(&s.GVT[wid][ta][tj][w]);
      WAj[w] = c1 * lds64(&s.WTT[wid][ta][tj][w]);
#pragma unroll
      for (int i = 0; i < D; ++i) GUj[w][i] = lds64(&s.GU[wid][(w * D + i) * D + tj]);
    }
    SAj = lds64(&s.S[wid][ta][tj]);
#pragma unroll
    for (int tb = 0; tb < NL; ++tb) {
      double m3[NV], mab = 0.0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        m3[v] = lds64(&s.M3T[(tb * NV + v) * NL + ta]);
        mab += m3[v];
      }
      const double *Eb = &s.E[wid][tb][0][0];  // one address per warp
      double Jc[D], ytj = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double y = SAj * lds64(Eb + NV * D + i), t2 = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          y += GAj[v] * lds64(Eb + v * D + i);
          t2 += m3[v] * GUj[v][i];
        }
        if (i == tj) ytj = y;
        Jc[i] = y + t2;
      }
      // this lane's share of the scalar part c1 (C_ab - C_ba) + c2 tr G (k = j): the second and third term are ytj
      double q = ytj;
#pragma unroll
      for (int w = 0; w < NV; ++w) q += lds64(&s.GVT[wid][tb][tj][w]) * WAj[w];
      if (semi) {  // c1t = 0: E carries no convection term
#pragma unroll
        for (int w = 0; w < NV; ++w) q -= c1 * GAj[w] * lds64(&s.WTT[wid][tb][tj][w]);
      }
      double dg = q;
#pragma unroll
      for (int k = 1; k < D; ++k) dg += __shfl_sync(0xffffffffu, q, lane0 + (tj + k) % D);
      dg += mab;
      if (active) {
        double *base = val + (int64_t)r0 * (D * D) + (int64_t)(slot[tb] - r0) * D + tj;
#pragma unroll
        for (int i = 0; i < D; ++i) atomicAdd(base + (int64_t)i * len, vol * (Jc[i] + (i == tj ? dg : 0.0)));
      }
    }
    __syncwarp();
  }
}

// boundary-facet part of J: -theta dt/rho mu ((grad delta)^T n, v)_ds
template <int D>
__global__ void k_momentum_J_facets(int64_t nbf, const int *__restrict__ bf_cell, const int *__restrict__ bf_local,
                                    const int *__restrict__ cell_nodes, const int *__restrict__ cells, const double *__restrict__ xyz,
                                    const int *__restrict__ rowptr, const int *__restrict__ smap, MomentumArgs a,
                                    double *__restrict__ val) {
  constexpr int NL = Elem<D>::NL2, NP = NL * NL;
  const int64_t total = nbf * NP;
  const double coef = -a.theta * a.dt / a.rho * a.mu;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fidx = t / NP;
    const int p = (int)(t - fidx * NP);
    const int ta = p / NL, tb = p - ta * NL;
    const int64_t c = bf_cell[fidx];
    const int f = bf_local[fidx];
    const int *cn = cell_nodes + c * NL;
    if (!fb_node_on_facet<D>(ta, f)) continue;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double B[D][D];
    fb_facet_J<D>(ta, tb, f, glam, vol, QF<D>::lam_ptr(), QF<D>::w_ptr(), QF<D>::NQ, B);
    const int I = cn[ta];
    const int r0 = rowptr[I];
    const int len = (rowptr[I + 1] - r0) * D;
    const int slot = smap[c * NP + p];
    double *base = val + (int64_t)r0 * (D * D) + (int64_t)(slot - r0) * D;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) atomicAdd(base + (int64_t)i * len + j, coef * B[i][j]);
  }
}


// ---- deterministic two-pass assembly of the cell part of J --------------------------------------------------------
// Pass 1 (k_momentum_J_cf<D, 1>) stores every element block; pass 2 sums, for each block of the matrix, its element
// contributions in a FIXED order (ascending cell) and writes the block once: no atomics (2.2 G fp64 atomicAdds at
// n = 74, the throughput limit of the one-pass kernel), no memset, bit-reproducible values.  The gather lists are the
// inverse of smap, built once per space.
__global__ void k_gather_count(int64_t total, const int *__restrict__ smap, int *__restrict__ count) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&count[smap[t]], 1);
}
__global__ void k_gather_fill(int64_t total, const int *__restrict__ smap, const int *__restrict__ gptr, int *__restrict__ cursor,
                              int *__restrict__ gsrc) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int s = smap[t];
    gsrc[gptr[s] + atomicAdd(&cursor[s], 1)] = (int)t;
  }
}
__global__ void k_gather_sort(int64_t nnz, const int *__restrict__ gptr, int *__restrict__ gsrc) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nnz; s += (int64_t)gridDim.x * blockDim.x) {
    const int b = gptr[s], e = gptr[s + 1];
    for (int i = b + 1; i < e; ++i) {  // insertion sort: the lists hold a few dozen entries at most
      const int v = gsrc[i];
      int j = i - 1;
      while (j >= b && gsrc[j] > v) {
        gsrc[j + 1] = gsrc[j];
        --j;
      }
      gsrc[j + 1] = v;
    }
  }
}
__global__ void k_exclusive_scan_serial(int64_t n, const int *__restrict__ in, int *__restrict__ out) {
  // one block, blocked scan: set-up only (n = number of matrix blocks)
  __shared__ long long carry;
  __shared__ int part[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int add = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n) out[i] = (int)(carry + part[threadIdx.x] - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = (int)carry;
}

// T lanes per block row; lane t sums the contributions of block t, t + T, ... and writes the D x D values (row-planar)
template <int D, int T>
__global__ void __launch_bounds__(256)
    k_jac_gather(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ gptr, const int *__restrict__ gsrc,
                 const double *__restrict__ ebuf, double *__restrict__ val) {
  const int lane = threadIdx.x % T;
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
  for (int64_t row = group; row < nrows; row += ngroups) {
    const int r0 = rowptr[row], nb = rowptr[row + 1] - r0;
    const int len = nb * D;
    double *base = val + (int64_t)r0 * (D * D);
    for (int t = lane; t < nb; t += T) {
      const int b = gptr[r0 + t], e = gptr[r0 + t + 1];
      double acc[D * D];
#pragma unroll
      for (int k = 0; k < D * D; ++k) acc[k] = 0.0;
      for (int q = b; q < e; ++q) {
        const double *src = ebuf + (int64_t)gsrc[q] * (D * D);
#pragma unroll
        for (int k = 0; k < D * D; ++k) acc[k] += __ldcs(src + k);
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) base[(int64_t)i * len + t * D + j] = acc[i * D + j];
    }
  }
}

static void build_gather_map(fb_ctx *ctx, DevSpace &W) {
  const int64_t total = W.nc * W.nl * W.nl;
  if (total > INT32_MAX) throw fb_cuda_error(FB_EINVAL, "gather map: too many element blocks for one GPU");
  cudaStream_t st = ctx->dev->stream;
  DBuf<int> count;
  count.alloc((size_t)W.nnz);
  count.zero(st);
  W.gptr.alloc((size_t)W.nnz + 1);
  W.gsrc.alloc((size_t)total);
  const int g = grid_for(total, 256, ctx->dev->sm_count * 16);
  FB_LAUNCH(ctx, k_gather_count, g, 256, 0, total, W.smap.p, count.p);
  k_exclusive_scan_serial<<<1, 1024, 0, st>>>(W.nnz, count.p, W.gptr.p);
  ctx->launches++;
  count.zero(st);
  FB_LAUNCH(ctx, k_gather_fill, g, 256, 0, total, W.smap.p, W.gptr.p, count.p, W.gsrc.p);
  FB_LAUNCH(ctx, k_gather_sort, grid_for(W.nnz, 256, ctx->dev->sm_count * 16), 256, 0, W.nnz, W.gptr.p, W.gsrc.p);
  FB_CUDA(cudaStreamSynchronize(st));
}

// Two-pass deterministic assembly: opt-in (fb_ns_opts.deterministic_assembly or FB_J_TWO_PASS=1).  Measured on B200 at
// n = 74 (profiles/r2_jacobian_assembly.txt): element pass 12.0 ms + gather pass 10.3 ms against 17 ms for the one-pass
// atomic kernel -- the closed-form element arithmetic (162 registers, 17 % occupancy), not the atomics, is what the
// assembly costs, so the default stays one-pass.
static bool g_two_pass_option = false;
void set_deterministic_assembly(bool on) { g_two_pass_option = on; }
static bool use_two_pass(const DevSpace &W) {
  const char *e = getenv("FB_J_TWO_PASS");
  return e ? atoi(e) != 0 : g_two_pass_option;
}

// boundary-facet part of J added to the ELEMENT blocks (two-pass mode): the facets are sorted by cell, the thread of a
// cell's first facet adds all of that cell's facets in order -> one writer per element block, fixed order
template <int D>
__global__ void k_momentum_J_facets_elem(int64_t nbf, const int *__restrict__ bf_cell, const int *__restrict__ bf_local,
                                         const int *__restrict__ cells, const double *__restrict__ xyz, MomentumArgs a,
                                         double *__restrict__ ebuf) {
  constexpr int NL = Elem<D>::NL2, NP = NL * NL;
  const int64_t total = nbf * NP;
  const double coef = -a.theta * a.dt / a.rho * a.mu;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fidx = t / NP;
    const int p = (int)(t - fidx * NP);
    const int ta = p / NL, tb = p - ta * NL;
    const int64_t c = bf_cell[fidx];
    if (fidx > 0 && bf_cell[fidx - 1] == c) continue;  // not the cell's first facet
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double acc[D][D];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) acc[i][j] = 0.0;
    bool any = false;
    for (int64_t f2 = fidx; f2 < nbf && bf_cell[f2] == c; ++f2) {
      const int f = bf_local[f2];
      if (!fb_node_on_facet<D>(ta, f)) continue;
      double B[D][D];
      fb_facet_J<D>(ta, tb, f, glam, vol, QF<D>::lam_ptr(), QF<D>::w_ptr(), QF<D>::NQ, B);
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) acc[i][j] += coef * B[i][j];
      any = true;
    }
    if (!any) continue;
    double *dst = ebuf + ((int64_t)c * NP + p) * (D * D);
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) dst[i * D + j] += acc[i][j];
  }
}

void assemble_momentum_J(fb_ctx *ctx, const DevSpace &W_, const MomentumArgs &a, double *Jval) {
  DevSpace &W = const_cast<DevSpace &>(W_);
  const int D = W.dim;
  static const int variant0 = getenv("FB_J_KERNEL") ? atoi(getenv("FB_J_KERNEL")) : 2;
  if (!((variant0 == 2 || variant0 == 4) && use_two_pass(W)))  // the gather pass writes every block (rows without cells do not exist)
    FB_CUDA(cudaMemsetAsync(Jval, 0, sizeof(double) * W.nnz * D * D, ctx->dev->stream));
  const int g = grid_for(W.nc * 32, MOM_WARPS * 32, ctx->dev->sm_count * 16);
  const int gf = grid_for(W.nbf * W.nl * W.nl, 128, ctx->dev->sm_count * 16);
  // 2 (default): closed form (k_momentum_J_cf); 0: degree-5 quadrature (k_momentum_J), kept as the cross-check.
  // Measured at n = 74 on B200 (2.43 M cells): 14.7 ms vs 29.8 ms per launch.
  static const int variant = getenv("FB_J_KERNEL") ? atoi(getenv("FB_J_KERNEL")) : 2;
  // closed form, lane = (test node, column): default in 3D (30 of 32 lanes; in 2D only 12 lanes would work, the
  // (test node, trial group) kernel stays).  Measured at n = 74, ms per assembly (ncu): (test node, trial group) 15.5,
  // column lanes k_momentum_J_cf2 11.70, merged trial-side operands + one-cell-ahead prefetch k_momentum_J_cf3 11.27
  // (profiles/r2_jacobian_assembly.txt).  FB_J_KERNEL=3 / 4 / 5 force cf2 / the trial-group kernel / cf3.
  const bool column_lanes = variant == 3 || variant == 5 || (variant == 2 && D == 3);
  if ((variant == 5 || (variant == 2 && D == 3)) && !use_two_pass(W)) {
    if (D == 2)
      FB_LAUNCH(ctx, k_momentum_J_cf3<2>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
    else
      FB_LAUNCH(ctx, k_momentum_J_cf3<3>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
  } else if (column_lanes && variant != 4 && !use_two_pass(W)) {
    if (D == 2)
      FB_LAUNCH(ctx, k_momentum_J_cf2<2>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
    else
      FB_LAUNCH(ctx, k_momentum_J_cf2<3>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
  } else if ((variant == 2 || variant == 4) && use_two_pass(W)) {
    if (!W.gptr.p) build_gather_map(ctx, W);
    W.ebuf.alloc((size_t)W.nc * W.nl * W.nl * D * D);
    const int gg = grid_for(W.n_owned * 8, 256, ctx->dev->sm_count * 16);
    const bool facets = W.nbf && a.theta != 0.0 && !a.skip_facets;
    if (D == 2) {
      FB_LAUNCH(ctx, (k_momentum_J_cf<2, 1>), g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, W.ebuf.p);
      if (facets) FB_LAUNCH(ctx, k_momentum_J_facets_elem<2>, gf, 128, 0, W.nbf, W.bf_cell.p, W.bf_local.p, W.cells.p, W.xyz.p, a, W.ebuf.p);
      FB_LAUNCH(ctx, (k_jac_gather<2, 8>), gg, 256, 0, W.nnodes, W.rowptr.p, W.gptr.p, W.gsrc.p, W.ebuf.p, Jval);
    } else {
      FB_LAUNCH(ctx, (k_momentum_J_cf<3, 1>), g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, W.ebuf.p);
      if (facets) FB_LAUNCH(ctx, k_momentum_J_facets_elem<3>, gf, 128, 0, W.nbf, W.bf_cell.p, W.bf_local.p, W.cells.p, W.xyz.p, a, W.ebuf.p);
      FB_LAUNCH(ctx, (k_jac_gather<3, 8>), gg, 256, 0, W.nnodes, W.rowptr.p, W.gptr.p, W.gsrc.p, W.ebuf.p, Jval);
    }
    return;  // the facet terms are in
  } else if (variant == 2 || variant == 4) {
    // resident blocks per SM the register allocation is bounded for (FB_J_MINB, experiments).  Measured at n = 74, ms per
    // assembly: unbounded (198 registers, 3 blocks/SM) 16.2; 4 blocks (128 registers, spills) 21.2; 5 (96) 32.7; 6 (80) 49.6
    static const int minb = getenv("FB_J_MINB") ? atoi(getenv("FB_J_MINB")) : 3;
    if (D == 2)
      FB_LAUNCH(ctx, k_momentum_J_cf<2>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
    else if (minb <= 3)
      FB_LAUNCH(ctx, k_momentum_J_cf<3>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
    else if (minb == 4)
      FB_LAUNCH(ctx, (k_momentum_J_cf<3, 0, 4>), g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
    else if (minb == 5)
      FB_LAUNCH(ctx, (k_momentum_J_cf<3, 0, 5>), g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
    else
      FB_LAUNCH(ctx, (k_momentum_J_cf<3, 0, 6>), g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
  } else if (D == 2) {
    FB_LAUNCH(ctx, k_momentum_J<2>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
  } else {
    FB_LAUNCH(ctx, k_momentum_J<3>, g, MOM_WARPS * 32, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, W.rowptr.p, W.smap.p, a, Jval);
  }
  if (W.nbf && a.theta != 0.0 && !a.skip_facets) {
    if (D == 2)
      FB_LAUNCH(ctx, k_momentum_J_facets<2>, gf, 128, 0, W.nbf, W.bf_cell.p, W.bf_local.p, W.cell_nodes.p, W.cells.p, W.xyz.p,
                W.rowptr.p, W.smap.p, a, Jval);
    else
      FB_LAUNCH(ctx, k_momentum_J_facets<3>, gf, 128, 0, W.nbf, W.bf_cell.p, W.bf_local.p, W.cell_nodes.p, W.cells.p, W.xyz.p,
                W.rowptr.p, W.smap.p, a, Jval);
  }
}

// =============================================================================
// pressure / correction right-hand sides: one thread per cell
// =============================================================================
template <int D>
__global__ void k_pressure_rhs(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                               const int *__restrict__ pcn, const double *__restrict__ xyz, double dt,
                               double rho, double mu, int rotational, const double *__restrict__ ui,
                               const double *__restrict__ p0, double *__restrict__ b) {
  constexpr int NL = Elem<D>::NL2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double Ue[NL * D], p0e[D + 1], be[D + 1];
    for (int a = 0; a < NL; ++a)
      for (int i = 0; i < D; ++i) Ue[a * D + i] = ui[(int64_t)cn[a] * D + i];
    for (int v = 0; v <= D; ++v) p0e[v] = p0[pcn[c * (D + 1) + v]];
    fb_pressure_rhs_cell<D>(glam, vol, Q2<D>::lam_ptr(), Q2<D>::w_ptr(), Q2<D>::NQ, Ue, p0e, dt, rho, mu, rotational, be);
    for (int v = 0; v <= D; ++v) atomicAdd(&b[pcn[c * (D + 1) + v]], be[v]);
  }
}

void assemble_pressure_rhs(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, double dt, double rho, double mu,
                           int rotational, const double *ui, const double *p0, double *b) {
  FB_CUDA(cudaMemsetAsync(b, 0, sizeof(double) * P.nnodes, ctx->dev->stream));
  const int g = grid_for(W.nc, 128, ctx->dev->sm_count * 32);
  if (W.dim == 2)
    FB_LAUNCH(ctx, k_pressure_rhs<2>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, dt, rho, mu, rotational, ui, p0, b);
  else
    FB_LAUNCH(ctx, k_pressure_rhs<3>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, dt, rho, mu, rotational, ui, p0, b);
}

template <int D>
__global__ void k_correction_grad(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                                  const int *__restrict__ pcn, const double *__restrict__ xyz,
                                  double dt, double rho, double mu, int rotational, const double *__restrict__ ui,
                                  const double *__restrict__ p1, const double *__restrict__ p0, double *__restrict__ b) {
  constexpr int NL = Elem<D>::NL2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double Ue[NL * D], dpe[D + 1], gphi[D];
    if (rotational)
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i) Ue[a * D + i] = ui[(int64_t)cn[a] * D + i];
    for (int v = 0; v <= D; ++v) dpe[v] = p1[pcn[c * (D + 1) + v]] - p0[pcn[c * (D + 1) + v]];
    fb_correction_gradphi<D>(glam, Ue, dpe, mu, rotational, gphi);
    const double coef = -dt / rho * vol;
    for (int a = 0; a < NL; ++a) {
      const double m = coef * fb_p2_mean<D>(a);
      if (m != 0.0)
        for (int k = 0; k < D; ++k) atomicAdd(&b[(int64_t)cn[a] * D + k], m * gphi[k]);
    }
  }
}

void assemble_correction_grad(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, double dt, double rho, double mu,
                              int rotational, const double *ui, const double *p1, const double *p0, double *b) {
  const int g = grid_for(W.nc, 128, ctx->dev->sm_count * 32);
  if (W.dim == 2)
    FB_LAUNCH(ctx, k_correction_grad<2>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, dt, rho, mu, rotational, ui, p1, p0, b);
  else
    FB_LAUNCH(ctx, k_correction_grad<3>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, dt, rho, mu, rotational, ui, p1, p0, b);
}

// =============================================================================
// heat operator (heat.py:54-58): one thread per (cell, a, b)
// =============================================================================
template <int D, int DEG>
__global__ void k_heat(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ wcell_nodes,
                       const int *__restrict__ cells, const double *__restrict__ xyz, const int *__restrict__ smap, const double *__restrict__ conv,
                       double kdiff, double *__restrict__ val) {
  constexpr int NL = DEG == 1 ? Elem<D>::NL1 : Elem<D>::NL2;
  constexpr int NLW = Elem<D>::NL2;
  const int64_t total = nc * NL * NL;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / (NL * NL);
    const int r = (int)(t - c * NL * NL);
    const int a = r / NL, b = r - a * NL;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double e = 0.0;
    for (int q = 0; q < Q5<D>::NQ; ++q) {
      double lam[D + 1], ga[D], gb[D];
      for (int m = 0; m <= D; ++m) lam[m] = Q5<D>::lam(q, m);
      double pa;
      if (DEG == 1) {
        pa = lam[a];
        for (int k = 0; k < D; ++k) {
          ga[k] = glam[a][k];
          gb[k] = glam[b][k];
        }
      } else {
        pa = fb_p2_phi<D>(a, lam);
        fb_p2_grad<D>(a, lam, glam, ga);
        fb_p2_grad<D>(b, lam, glam, gb);
      }
      double s = 0.0;
      for (int k = 0; k < D; ++k) s -= kdiff * ga[k] * gb[k];
      if (conv) {
        const int *wn = wcell_nodes + c * NLW;
        double cg = 0.0;
        for (int n = 0; n < NLW; ++n) {
          const double pn = fb_p2_phi<D>(n, lam);
          for (int k = 0; k < D; ++k) cg += pn * conv[(int64_t)wn[n] * D + k] * gb[k];
        }
        s -= cg * pa;
      }
      e += Q5<D>::w(q) * s;
    }
    atomicAdd(&val[smap[t]], e * vol);
  }
}

void assemble_heat(fb_ctx *ctx, const DevSpace &V, const DevSpace *W, const double *conv, double kdiff, double *val) {
  FB_CUDA(cudaMemsetAsync(val, 0, sizeof(double) * V.nnz, ctx->dev->stream));
  const int g = grid_for(V.nc * V.nl * V.nl, 128, ctx->dev->sm_count * 32);
  const int *wcn = W ? W->cell_nodes.p : nullptr;
  if (!W) conv = nullptr;
#define FB_HT(D, DEG) \
  FB_LAUNCH(ctx, (k_heat<D, DEG>), g, 128, 0, V.nc, V.cell_nodes.p, wcn, V.cells.p, V.xyz.p, V.smap.p, conv, kdiff, val)
  if (V.dim == 2 && V.degree == 1) FB_HT(2, 1);
  else if (V.dim == 2) FB_HT(2, 2);
  else if (V.degree == 1) FB_HT(3, 1);
  else FB_HT(3, 2);
#undef FB_HT
}

// SUPG terms of the heat operator (heat.py:60-86), triangles only like the reference's SupgStab:
//   Msupg[a][b] = int phi_b tau (conv.grad phi_a)
//   A[a][b]    += int [ kappa Lap(phi_b)/rho_cp - conv.grad phi_b ] tau (conv.grad phi_a)
//   b[a]       += int source/rho_cp tau (conv.grad phi_a)
// tau is the degree-1 Expression of stabilization.py: evaluated at the cell's vertices (with the
// nodal convection there) and interpolated linearly.  Degree-7 rule (tau P1 * conv P2 * P1 * P2 * P1).
template <int DEG>
__global__ void k_heat_supg(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ wcell_nodes,
                            const int *__restrict__ cells, const double *__restrict__ xyz, const int *__restrict__ smap,
                            const double *__restrict__ conv, double kappa, double rho_cp, double source, int pdeg,
                            double *__restrict__ Aval, double *__restrict__ Mval, double *__restrict__ bvec, int *err) {
  constexpr int D = 2;
  constexpr int NL = DEG == 1 ? Elem<D>::NL1 : Elem<D>::NL2;
  constexpr int NLW = Elem<D>::NL2;
  const int64_t total = nc * NL * NL;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / (NL * NL);
    const int r = (int)(t - c * NL * NL);
    const int a = r / NL, b = r - a * NL;
    const int *cv = cells + c * 3;
    const int *wn = wcell_nodes + c * NLW;
    double X[6];
    for (int v = 0; v < 3; ++v)
      for (int k = 0; k < 2; ++k) X[v * 2 + k] = xyz[(int64_t)cv[v] * 2 + k];
    double glam[3][2], vol;
    fb_geometry<2>(X, glam, vol);
    double tauv[3];
    for (int v = 0; v < 3; ++v) {
      const double cvv[2] = {conv[(int64_t)wn[v] * 2], conv[(int64_t)wn[v] * 2 + 1]};
      tauv[v] = fb_supg_tau(X, cvv, kappa, pdeg);
      if (tauv[v] < 0.0) {
        *err = 1;
        tauv[v] = 0.0;
      }
    }
    double lap_b = 0.0;  // Laplacian of phi_b (constant per cell; zero for P1)
    if (DEG == 2) {
      double H[2][2];
      fb_p2_hess<2>(b, glam, H);
      lap_b = H[0][0] + H[1][1];
    }
    double em = 0.0, ea = 0.0, eb = 0.0;
    for (int q = 0; q < dq::TRI_D7_NQ; ++q) {
      double lam[3], ga[2], gb[2];
      for (int m = 0; m < 3; ++m) lam[m] = dq::TRI_D7_LAM[q][m];
      double pb;
      if (DEG == 1) {
        pb = lam[b];
        for (int k = 0; k < 2; ++k) {
          ga[k] = glam[a][k];
          gb[k] = glam[b][k];
        }
      } else {
        pb = fb_p2_phi<2>(b, lam);
        fb_p2_grad<2>(a, lam, glam, ga);
        fb_p2_grad<2>(b, lam, glam, gb);
      }
      double cq[2] = {0.0, 0.0};
      for (int n = 0; n < NLW; ++n) {
        const double pn = fb_p2_phi<2>(n, lam);
        cq[0] += pn * conv[(int64_t)wn[n] * 2];
        cq[1] += pn * conv[(int64_t)wn[n] * 2 + 1];
      }
      const double tau = tauv[0] * lam[0] + tauv[1] * lam[1] + tauv[2] * lam[2];
      const double test = tau * (cq[0] * ga[0] + cq[1] * ga[1]) * dq::TRI_D7_W[q];
      em += pb * test;
      ea += (kappa * lap_b / rho_cp - (cq[0] * gb[0] + cq[1] * gb[1])) * test;
      if (b == 0) eb += source / rho_cp * test;
    }
    atomicAdd(&Mval[smap[t]], em * vol);
    atomicAdd(&Aval[smap[t]], ea * vol);
    if (b == 0 && source != 0.0) atomicAdd(&bvec[cell_nodes[c * NL + a]], eb * vol);
  }
}

int assemble_heat_supg(fb_ctx *ctx, const DevSpace &V, const DevSpace &W, const double *conv, double kappa, double rho_cp,
                       double source, double *Aval, double *Mval, double *bvec) {
  if (V.dim != 2) throw fb_cuda_error(FB_EINVAL, "SUPG stabilisation is defined for triangles only (stabilization.py:84-92)");
  int *err = ctx->dev->flag;
  FB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), ctx->dev->stream));
  const int g = grid_for(V.nc * V.nl * V.nl, 128, ctx->dev->sm_count * 32);
  if (V.degree == 1)
    FB_LAUNCH(ctx, k_heat_supg<1>, g, 128, 0, V.nc, V.cell_nodes.p, W.cell_nodes.p, V.cells.p, V.xyz.p, V.smap.p, conv, kappa,
              rho_cp, source, 1, Aval, Mval, bvec, err);
  else
    FB_LAUNCH(ctx, k_heat_supg<2>, g, 128, 0, V.nc, V.cell_nodes.p, W.cell_nodes.p, V.cells.p, V.xyz.p, V.smap.p, conv, kappa,
              rho_cp, source, 2, Aval, Mval, bvec, err);
  int h = 0;
  FB_CUDA(cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, ctx->dev->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->dev->stream));
  FB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), ctx->dev->stream));
  return h;
}

// =============================================================================
// Stokes divergence block, matrix-free (stokes.py:40-42):  (B u)_a = -int psi_a div u,  (B^T p)_(b,j) = -int p d_j phi_b
// =============================================================================
template <int D, int TRANSPOSE>
__global__ void k_stokes_div(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                             const int *__restrict__ pcn, const double *__restrict__ xyz,
                             const double *__restrict__ in, double *__restrict__ out) {
  constexpr int NL = Elem<D>::NL2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    if (!TRANSPOSE) {
      double be[D + 1];
      for (int v = 0; v <= D; ++v) be[v] = 0.0;
      for (int q = 0; q < Q2<D>::NQ; ++q) {
        double lam[D + 1];
        for (int m = 0; m <= D; ++m) lam[m] = Q2<D>::lam(q, m);
        double div = 0.0;
        for (int a = 0; a < NL; ++a) {
          double g[D];
          fb_p2_grad<D>(a, lam, glam, g);
          for (int i = 0; i < D; ++i) div += in[(int64_t)cn[a] * D + i] * g[i];
        }
        for (int v = 0; v <= D; ++v) be[v] -= Q2<D>::w(q) * vol * div * lam[v];
      }
      for (int v = 0; v <= D; ++v) atomicAdd(&out[pcn[c * (D + 1) + v]], be[v]);
    } else {
      for (int q = 0; q < Q2<D>::NQ; ++q) {
        double lam[D + 1];
        for (int m = 0; m <= D; ++m) lam[m] = Q2<D>::lam(q, m);
        double p = 0.0;
        for (int v = 0; v <= D; ++v) p += in[pcn[c * (D + 1) + v]] * lam[v];
        const double w = -Q2<D>::w(q) * vol * p;
        for (int a = 0; a < NL; ++a) {
          double g[D];
          fb_p2_grad<D>(a, lam, glam, g);
          for (int i = 0; i < D; ++i) atomicAdd(&out[(int64_t)cn[a] * D + i], w * g[i]);
        }
      }
    }
  }
}

void stokes_div(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, const double *u, double *out_p) {
  const int g = grid_for(W.nc, 128, ctx->dev->sm_count * 32);
  if (W.dim == 2)
    FB_LAUNCH(ctx, (k_stokes_div<2, 0>), g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, u, out_p);
  else
    FB_LAUNCH(ctx, (k_stokes_div<3, 0>), g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, u, out_p);
}
void stokes_grad(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, const double *p, double *out_u) {
  const int g = grid_for(W.nc, 128, ctx->dev->sm_count * 32);
  if (W.dim == 2)
    FB_LAUNCH(ctx, (k_stokes_div<2, 1>), g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, p, out_u);
  else
    FB_LAUNCH(ctx, (k_stokes_div<3, 1>), g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, P.cell_nodes.p, W.xyz.p, p, out_u);
}


// =============================================================================
// driver-side quantities (tests/test_karman_vortex_street.py:261-287, tests/test_sealed_box.py:134-141)
// =============================================================================
// b_a = int |u_h| phi_a dx with the caller's quadrature rule (the integrand is not polynomial: the rule is part of the
// definition; the reference asks FFC for degree 4): one thread per cell
template <int D>
__global__ void k_magnitude_load(int64_t nc, const int *__restrict__ cell_nodes, const int *__restrict__ cells,
                                 const double *__restrict__ xyz, const double *__restrict__ u, int nq, const double *__restrict__ qlam,
                                 const double *__restrict__ qw, double *__restrict__ b) {
  constexpr int NL = Elem<D>::NL2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    cell_geometry<D>(cells + c * (D + 1), xyz, glam, vol);
    double U[NL][D], acc[NL];
#pragma unroll
    for (int a = 0; a < NL; ++a) {
      acc[a] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) U[a][i] = u[(int64_t)cn[a] * D + i];
    }
    for (int q = 0; q < nq; ++q) {
      double lam[D + 1], phi[NL], m2 = 0.0;
#pragma unroll
      for (int m = 0; m <= D; ++m) lam[m] = qlam[q * (D + 1) + m];
#pragma unroll
      for (int a = 0; a < NL; ++a) phi[a] = fb_p2_phi<D>(a, lam);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double ui = 0.0;
#pragma unroll
        for (int a = 0; a < NL; ++a) ui += U[a][i] * phi[a];
        m2 += ui * ui;
      }
      const double wm = qw[q] * vol * sqrt(m2);
#pragma unroll
      for (int a = 0; a < NL; ++a) acc[a] += wm * phi[a];
    }
#pragma unroll
    for (int a = 0; a < NL; ++a) atomicAdd(&b[cn[a]], acc[a]);
  }
}

// out[0] = max_i |x_i| over n entries; out[1] = max over nodes of the Euclidean norm of the D components of u (may be null)
__global__ void k_max_norms(int64_t n, const double *__restrict__ x, int64_t nnodes, int D, const double *__restrict__ u,
                            unsigned long long *out) {
  double m0 = 0.0, m1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m0 = fmax(m0, fabs(x[i]));
  if (u)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnodes; i += (int64_t)gridDim.x * blockDim.x) {
      double s2 = 0.0;
      for (int k = 0; k < D; ++k) s2 += u[i * D + k] * u[i * D + k];
      m1 = fmax(m1, sqrt(s2));
    }
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(m1));
  }
}

void assemble_magnitude_load(fb_ctx *ctx, const DevSpace &W, const double *u, int nq, const double *qlam, const double *qw, double *b) {
  FB_CUDA(cudaMemsetAsync(b, 0, sizeof(double) * W.nnodes, ctx->dev->stream));
  const int g = grid_for(W.nc, 128, ctx->dev->sm_count * 16);
  if (W.dim == 2)
    FB_LAUNCH(ctx, k_magnitude_load<2>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, u, nq, qlam, qw, b);
  else
    FB_LAUNCH(ctx, k_magnitude_load<3>, g, 128, 0, W.nc, W.cell_nodes.p, W.cells.p, W.xyz.p, u, nq, qlam, qw, b);
}

void vec_max_norms(fb_ctx *ctx, const double *x, int64_t n, const double *u, int64_t nnodes, int D, double *out2_host) {
  fb_device_state *dv = ctx->dev;
  unsigned long long *slot = reinterpret_cast<unsigned long long *>(dv->red + 46);
  FB_CUDA(cudaMemsetAsync(slot, 0, 2 * sizeof(unsigned long long), dv->stream));
  FB_LAUNCH(ctx, k_max_norms, vgrid(ctx, std::max<int64_t>(n, nnodes)), 256, 0, n, x, nnodes, D, u, slot);
  FB_CUDA(cudaMemcpyAsync(dv->host_pinned + 42, slot, 2 * sizeof(double), cudaMemcpyDeviceToHost, dv->stream));
  FB_CUDA(cudaStreamSynchronize(dv->stream));
  out2_host[0] = dv->host_pinned[42];
  out2_host[1] = dv->host_pinned[43];
}


// tau at the three vertices of every cell (flow/stabilization.py:13-152), for inspection / parity: what k_heat_supg uses
__global__ void k_supg_tau_values(int64_t nc, const int *__restrict__ cells, const int *__restrict__ wcell_nodes, int wnl,
                                  const double *__restrict__ xyz, const double *__restrict__ conv, double eps, int p,
                                  double *__restrict__ out, int *__restrict__ too_large) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    double X[6];
    for (int v = 0; v < 3; ++v) {
      const int node = cells[c * 3 + v];
      X[2 * v] = xyz[(int64_t)node * 2];
      X[2 * v + 1] = xyz[(int64_t)node * 2 + 1];
    }
    for (int v = 0; v < 3; ++v) {
      const int64_t n = wcell_nodes[c * wnl + v];
      const double cv[2] = {conv[n * 2], conv[n * 2 + 1]};
      const double t = fb_supg_tau(X, cv, eps, p);
      if (t < 0.0) *too_large = 1;
      out[c * 3 + v] = t;
    }
  }
}

// device cells are stored in Morton order: `order` maps device position -> mesh cell for the scatter back
int supg_tau_values(fb_ctx *ctx, const DevSpace &W, const double *conv, double eps, int p, const int *order_dev, double *out_mesh_order) {
  DBuf<double> tmp;
  DBuf<int> flag;
  tmp.alloc((size_t)W.nc * 3);
  flag.alloc(1);
  flag.zero(ctx->dev->stream);
  FB_LAUNCH(ctx, k_supg_tau_values, grid_for(W.nc, 256, ctx->dev->sm_count * 8), 256, 0, W.nc, W.cells.p, W.cell_nodes.p, W.nl, W.xyz.p, conv, eps, p,
            tmp.p, flag.p);
  std::vector<double> h((size_t)W.nc * 3);
  int bad = 0;
  FB_CUDA(cudaMemcpyAsync(h.data(), tmp.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->dev->stream));
  FB_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->dev->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->dev->stream));
  const std::vector<int32_t> &corder = fb_mesh_cell_order(W.host->mesh);
  for (int64_t c = 0; c < W.nc; ++c)
    for (int v = 0; v < 3; ++v) out_mesh_order[(int64_t)corder[c] * 3 + v] = h[c * 3 + v];
  (void)order_dev;
  return bad;
}
