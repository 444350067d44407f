// Tile-CSR SpMM for the scalar Lagrange operators (P2 mass M, S = M + theta dt nu K, P1/P2 stiffness, heat operator)
// applied to NC interleaved components: the PETSc MatMult of the reference's CG solves
// (pressure_correction.py:451-464 velocity correction, and the inner solves of the momentum preconditioner).
//
// Why a second format.  The row-wise CSR kernel (k_spmm_u) gathers x[col] straight from global memory: ~29 gathers of
// 24 bytes per row, each a separate 32-byte sector.  ncu showed it bound by L1 wavefronts (L1/TEX 89 % busy, DRAM 41 %),
// not by HBM.  Here the rows are grouped into TILES of spatially close nodes (Morton order of the node coordinates,
// a private row order of this format -- x and y keep the canonical numbering of the ABI):
//   * the union of the columns a tile touches (~3 nodes per row instead of 29) is gathered ONCE into shared memory
//     with cp.async (LDGSTS), one tile ahead;
//   * the entries are stored tile by tile as fp64 value + 16-bit index into that union (10 bytes per entry instead
//     of 12) and streamed into shared memory by the TMA unit (cp.async.bulk + mbarrier), one tile ahead, with an
//     L2 evict-first policy so that x stays L2 resident;
//   * the products run from shared memory; L1 only serves the union gather.
// One persistent CTA per SM, two stages.  Algorithmic bytes stay those of scalar CSR.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "fb_ops.h"

namespace {

constexpr int TILE_THREADS = 1024;
constexpr int TILE_T = 4;                           // lanes per row
constexpr int TILE_RCAP = TILE_THREADS / TILE_T;    // rows per tile: one pass of the block
constexpr int TILE_ECAP = 6144;                     // entries per tile (multiple of 8)
constexpr int TILE_UCAP = 1280;                     // union columns per tile (NC = 3: 30 KB of x per stage)
constexpr int TILE_DESC = 8;                        // ints per tile descriptor

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier; evict-first in L2 (streamed once)
__device__ __forceinline__ void tma_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

struct TileArgs {
  int ntiles;
  const int *desc;          // ntiles * TILE_DESC: row0, nr, e0, ne_pad, u0, nu, -, -
  const int *rowid;         // tile-order row -> canonical row
  const int *rptr;          // per tile nr + 1 offsets relative to e0, stored at row0 + tile
  const int *ucol;          // union column lists
  const uint16_t *lidx;     // entries: index into the tile's union
  const double *tval;       // entries: values (packed from the CSR values, see k_tile_pack)
  const uint8_t *mask;      // per dof: identity row (may be null)
  const double *x;
  double *y;
  const double *w;
  double *partials;
  unsigned int *counter;
  double *red;
  int slot;
  const int *flag;
};

template <int NC>
struct TileSmem {
  static constexpr int VAL_BYTES = TILE_ECAP * 8;
  static constexpr int X_BYTES = TILE_UCAP * NC * 8;
  static constexpr int IDX_BYTES = TILE_ECAP * 2;
  static constexpr int STAGE = VAL_BYTES + X_BYTES + IDX_BYTES;
  static constexpr int TOTAL = 2 * STAGE + 64;
};

// DOT: 0 none; 1: red[slot] = w.y; 2: + red[slot+1] = y.y; 3: (w.x, y.x, x.x) -- as k_spmm_u
template <int NC, int DOT>
__global__ void __launch_bounds__(TILE_THREADS, 1) k_tile_spmm(const TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = TileSmem<NC>;
  if (a.flag && *a.flag) return;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * L::STAGE);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  auto stage_val = [&](int s) { return reinterpret_cast<double *>(smem + s * L::STAGE); };
  auto stage_x = [&](int s) { return reinterpret_cast<double *>(smem + s * L::STAGE + L::VAL_BYTES); };
  auto stage_idx = [&](int s) { return reinterpret_cast<uint16_t *>(smem + s * L::STAGE + L::VAL_BYTES + L::X_BYTES); };

  // union columns of a tile, two per thread (TILE_UCAP <= 2 * TILE_THREADS), held in registers one tile ahead
  auto load_ucols = [&](int t, int &c0, int &c1) {
    c0 = c1 = -1;
    if (t < a.ntiles) {
      const int u0 = a.desc[t * TILE_DESC + 4], nu = a.desc[t * TILE_DESC + 5];
      if (tid < nu) c0 = a.ucol[u0 + tid];
      if (tid + TILE_THREADS < nu) c1 = a.ucol[u0 + tid + TILE_THREADS];
    }
  };
  // start the loads of tile t into stage s: TMA for the entries, cp.async gather for the x union
  auto issue = [&](int t, int s, int c0, int c1) {
    if (t < a.ntiles) {
      if (tid == 0) {
        const int e0 = a.desc[t * TILE_DESC + 2], ne = a.desc[t * TILE_DESC + 3];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this stage are done
        mbar_expect_tx(&bar[s], (uint32_t)ne * 10u);
        tma_bulk_load(stage_val(s), a.tval + e0, (uint32_t)ne * 8u, &bar[s], policy);
        tma_bulk_load(stage_idx(s), a.lidx + e0, (uint32_t)ne * 2u, &bar[s], policy);
      }
      double *xs = stage_x(s);
      if (c0 >= 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) cp_async8(xs + tid * NC + c, a.x + (int64_t)c0 * NC + c);
      }
      if (c1 >= 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) cp_async8(xs + (tid + TILE_THREADS) * NC + c, a.x + (int64_t)c1 * NC + c);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double d[3] = {0.0, 0.0, 0.0};
  int t = blockIdx.x;
  int c0, c1;
  load_ucols(t, c0, c1);
  issue(t, 0, c0, c1);
  load_ucols(t + gridDim.x, c0, c1);
  uint32_t phasebits = 0u;  // bit s: parity the next wait on bar[s] expects
  const int lane = tid % TILE_T, rl = tid / TILE_T;
  for (int s = 0; t < a.ntiles; t += gridDim.x, s ^= 1) {
    // one tile ahead: entries by TMA, x union by cp.async; two tiles ahead: the union column ids into registers
    issue(t + gridDim.x, s ^ 1, c0, c1);
    load_ucols(t + 2 * gridDim.x, c0, c1);
    // this tile's row data (global, coalesced) while its stage completes
    const int row0 = a.desc[t * TILE_DESC + 0], nr = a.desc[t * TILE_DESC + 1];
    int eb = 0, ee = 0, row = -1;
    if (rl < nr) {
      eb = a.rptr[row0 + t + rl];
      ee = a.rptr[row0 + t + rl + 1];
      row = a.rowid[row0 + rl];
    }
    double wv = 0.0, xd = 0.0;
    bool masked = false;
    if (row >= 0 && lane < NC) {
      const int64_t dof = (int64_t)row * NC + lane;
      if (DOT >= 1) wv = a.w[dof];
      if (DOT == 3 || a.mask) xd = a.x[dof];
      if (a.mask) masked = a.mask[dof] != 0;
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();  // every thread's part of the x union has landed
    mbar_wait(&bar[s], (phasebits >> s) & 1u);
    phasebits ^= 1u << s;
    const double *vs = stage_val(s);
    const double *xs = stage_x(s);
    const uint16_t *is = stage_idx(s);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll 4
    for (int k = eb + lane; k < ee; k += TILE_T) {
      const double av = vs[k];
      const int j = is[k];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] += av * xs[j * NC + c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int o = TILE_T / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (row >= 0 && lane < NC) {
      double yc = acc[0];
#pragma unroll
      for (int c = 1; c < NC; ++c)
        if (lane == c) yc = acc[c];
      if (masked) yc = xd;
      a.y[(int64_t)row * NC + lane] = yc;
      if (DOT == 1 || DOT == 2) d[0] += wv * yc;
      if (DOT == 2) d[1] += yc * yc;
      if (DOT == 3) {
        d[0] += wv * xd;
        d[1] += yc * xd;
        d[2] += xd * xd;
      }
    }
    __syncthreads();  // stage s is free again (it is refilled at the top of the next iteration)
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (DOT >= 1) {
    if (DOT == 1) {
      double v1[1] = {d[0]};
      fb_grid_reduce<1>(v1, a.partials, a.counter, a.red, a.slot);
    } else if (DOT == 2) {
      double v2[2] = {d[0], d[1]};
      fb_grid_reduce<2>(v2, a.partials, a.counter, a.red, a.slot);
    } else {
      fb_grid_reduce<3>(d, a.partials, a.counter, a.red, a.slot);
    }
  }
}

__global__ void k_tile_pack(int64_t nent, const int *__restrict__ src, const double *__restrict__ val, double *__restrict__ tval) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nent; k += (int64_t)gridDim.x * blockDim.x) {
    const int s = src[k];
    tval[k] = s >= 0 ? val[s] : 0.0;
  }
}

template <int NC, int DOT>
void launch_tile(fb_ctx *ctx, const TileArgs &a) {
  using L = TileSmem<NC>;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(k_tile_spmm<NC, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int grid = std::min(a.ntiles, ctx->dev->sm_count);
  k_tile_spmm<NC, DOT><<<grid, TILE_THREADS, L::TOTAL, ctx->dev->stream>>>(a);
  ctx->launches++;
}

template <int NC>
void launch_tile_nc(fb_ctx *ctx, const TileArgs &a, int dot_mode) {
  if (dot_mode == 0) launch_tile<NC, 0>(ctx, a);
  else if (dot_mode == 1) launch_tile<NC, 1>(ctx, a);
  else if (dot_mode == 2) launch_tile<NC, 2>(ctx, a);
  else launch_tile<NC, 3>(ctx, a);
}

inline uint64_t spread3(uint64_t v) {  // 21 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}
inline uint64_t spread2(uint64_t v) {  // 31 bits -> every second bit
  v &= 0x7fffffffull;
  v = (v | v << 16) & 0x0000ffff0000ffffull;
  v = (v | v << 8) & 0x00ff00ff00ff00ffull;
  v = (v | v << 4) & 0x0f0f0f0f0f0f0f0full;
  v = (v | v << 2) & 0x3333333333333333ull;
  v = (v | v << 1) & 0x5555555555555555ull;
  return v;
}

}  // namespace

bool tile_enabled() {
  static const int on = getenv("FB_TILE") ? atoi(getenv("FB_TILE")) : 1;
  return on != 0;
}

// Host side of the format: tiles of the owned rows of s's node pattern (pure host code, once per space).
struct HostTile {
  std::vector<int> desc, order, rptr, ucol, src;
  std::vector<uint16_t> lidx;
  int ntiles = 0;
};

static void tile_format_build_host(fb_space *s, HostTile &h) {
  fb_space_build_pattern(s);
  const int dim = s->mesh->dim;
  const int64_t n = s->n_owned, nn = s->nnodes;
  const std::vector<int64_t> &ip = s->indptr;
  const std::vector<int32_t> &ix = s->indices;
  // Morton order of the owned rows
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < dim; ++k) {
      lo[k] = std::min(lo[k], s->coords[i * dim + k]);
      hi[k] = std::max(hi[k], s->coords[i * dim + k]);
    }
  double ext = 0.0;
  for (int k = 0; k < dim; ++k) ext = std::max(ext, hi[k] - lo[k]);
  if (!(ext > 0.0)) ext = 1.0;
  const int bits = dim == 3 ? 21 : 31;
  const double scale = (double)((1ull << bits) - 1) / ext;
  std::vector<uint64_t> key((size_t)n);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    uint64_t q[3] = {0, 0, 0};
    for (int k = 0; k < dim; ++k) q[k] = (uint64_t)((s->coords[i * dim + k] - lo[k]) * scale);
    key[i] = dim == 3 ? (spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2) : (spread2(q[0]) | spread2(q[1]) << 1);
  }
  std::vector<int> &order = h.order;
  order.resize((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  key.clear();
  key.shrink_to_fit();

  std::vector<int> &desc = h.desc, &rptr = h.rptr, &ucol = h.ucol, &src = h.src;
  std::vector<uint16_t> &lidx = h.lidx;
  rptr.reserve((size_t)n + n / 128 + 16);
  src.reserve(ix.size() + ix.size() / 64);
  lidx.reserve(ix.size() + ix.size() / 64);
  std::vector<int> mark((size_t)nn, -1);
  std::vector<uint16_t> lid((size_t)nn, 0);
  std::vector<int> ulist;
  ulist.reserve(TILE_UCAP);
  int64_t pos = 0;
  int tile = 0;
  while (pos < n) {
    // grow the tile row by row in Morton order until one of the caps binds
    ulist.clear();
    int nr = 0, ne = 0;
    const int64_t row0 = pos;
    while (pos < n && nr < TILE_RCAP) {
      const int r = order[pos];
      const int len = (int)(ip[r + 1] - ip[r]);
      if (len > TILE_ECAP || len > TILE_UCAP) throw fb_cuda_error(FB_EINVAL, "tile format: matrix row too long");
      if (ne + len > TILE_ECAP) break;
      int fresh = 0;
      for (int64_t k = ip[r]; k < ip[r + 1]; ++k) fresh += mark[ix[k]] != tile;
      if ((int)ulist.size() + fresh > TILE_UCAP) break;
      for (int64_t k = ip[r]; k < ip[r + 1]; ++k)
        if (mark[ix[k]] != tile) {
          mark[ix[k]] = tile;
          ulist.push_back(ix[k]);
        }
      ne += len;
      ++nr;
      ++pos;
    }
    std::sort(ulist.begin(), ulist.end());
    for (size_t j = 0; j < ulist.size(); ++j) lid[ulist[j]] = (uint16_t)j;
    const int e0 = (int)lidx.size();
    int off = 0;
    for (int i = 0; i < nr; ++i) {
      const int r = order[row0 + i];
      rptr.push_back(off);
      for (int64_t k = ip[r]; k < ip[r + 1]; ++k) {
        lidx.push_back(lid[ix[k]]);
        src.push_back((int)k);
      }
      off += (int)(ip[r + 1] - ip[r]);
    }
    rptr.push_back(off);
    const int ne_pad = (off + 7) & ~7;
    for (int k = off; k < ne_pad; ++k) {
      lidx.push_back(0);
      src.push_back(-1);
    }
    const int d8[TILE_DESC] = {(int)row0, nr, e0, ne_pad, (int)ucol.size(), (int)ulist.size(), 0, 0};
    desc.insert(desc.end(), d8, d8 + TILE_DESC);
    ucol.insert(ucol.end(), ulist.begin(), ulist.end());
    ++tile;
    if ((int64_t)lidx.size() > (int64_t)INT32_MAX - TILE_ECAP) throw fb_cuda_error(FB_EINVAL, "tile format: too many entries");
  }
  h.ntiles = tile;
}

void tile_format_build(fb_space *s, TileFormat &tf, cudaStream_t st) {
  HostTile h;
  tile_format_build_host(s, h);
  tf.nrows = s->n_owned;
  tf.ntiles = h.ntiles;
  tf.nent = (int64_t)h.lidx.size();
  tf.union_total = (int64_t)h.ucol.size();
  tf.desc.upload(h.desc.data(), h.desc.size(), st);
  tf.rowid.upload(h.order.data(), h.order.size(), st);
  tf.rptr.upload(h.rptr.data(), h.rptr.size(), st);
  tf.ucol.upload(h.ucol.data(), h.ucol.size(), st);
  tf.lidx.upload(h.lidx.data(), h.lidx.size(), st);
  tf.src.upload(h.src.data(), h.src.size(), st);
  FB_CUDA(cudaStreamSynchronize(st));  // the host vectors are temporaries
}

// Host-only self check (works in a context without a device): build the format and verify that it is the CSR pattern
// -- every owned row exactly once, ucol[lidx] == CSR columns in CSR order, src == CSR slots, caps and TMA alignment
// respected.  stats: ntiles, entries incl. padding, sum of union sizes, max rows, max entries, max union.
extern "C" int fb_space_tile_check(fb_space *s, int64_t *stats) {
  if (!s) return FB_EINVAL;
  try {
    HostTile h;
    tile_format_build_host(s, h);
    const int64_t n = s->n_owned;
    std::vector<uint8_t> seen((size_t)n, 0);
    int64_t max_r = 0, max_e = 0, max_u = 0, rp = 0;
    for (int t = 0; t < h.ntiles; ++t) {
      const int *d = &h.desc[(size_t)t * TILE_DESC];
      const int row0 = d[0], nr = d[1], e0 = d[2], ne = d[3], u0 = d[4], nu = d[5];
      if (nr < 1 || nr > TILE_RCAP || ne > TILE_ECAP || nu > TILE_UCAP || (e0 & 7) || (ne & 7)) return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: caps/alignment");
      if (rp != row0 + t) return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: rptr layout");
      for (int j = 1; j < nu; ++j)
        if (h.ucol[u0 + j - 1] >= h.ucol[u0 + j]) return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: union not ascending");
      for (int i = 0; i < nr; ++i) {
        const int r = h.order[row0 + i];
        if (r < 0 || r >= n || seen[r]) return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: row covered twice");
        seen[r] = 1;
        const int eb = h.rptr[rp + i], ee = h.rptr[rp + i + 1];
        if (ee - eb != (int)(s->indptr[r + 1] - s->indptr[r])) return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: row length");
        for (int k = eb; k < ee; ++k) {
          const int64_t slot = s->indptr[r] + (k - eb);
          if (h.src[e0 + k] != (int)slot || h.lidx[e0 + k] >= nu || h.ucol[u0 + h.lidx[e0 + k]] != s->indices[slot])
            return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: entry mismatch");
        }
      }
      rp += nr + 1;
      max_r = std::max<int64_t>(max_r, nr);
      max_e = std::max<int64_t>(max_e, ne);
      max_u = std::max<int64_t>(max_u, nu);
    }
    for (int64_t i = 0; i < n; ++i)
      if (!seen[i]) return fb_fail(s->mesh->ctx, FB_EINVAL, "tile check: row missing");
    if (stats) {
      stats[0] = h.ntiles;
      stats[1] = (int64_t)h.lidx.size();
      stats[2] = (int64_t)h.ucol.size();
      stats[3] = max_r;
      stats[4] = max_e;
      stats[5] = max_u;
    }
  } catch (const std::exception &e) {
    return fb_fail(s->mesh->ctx, FB_EINVAL, e.what());
  }
  return FB_OK;
}

void tile_pack(fb_ctx *ctx, const TileFormat &tf, const double *val, double *tval) {
  const int64_t n = tf.nent;
  int64_t g = (n + 255) / 256;
  g = std::max<int64_t>(1, std::min<int64_t>(g, ctx->dev->sm_count * 8));
  FB_LAUNCH(ctx, k_tile_pack, (int)g, 256, 0, n, tf.src.p, val, tval);
}

void tile_spmm(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
               const int *flag) {
  const TileFormat &tf = *A.tile;
  fb_device_state *dv = ctx->dev;
  TileArgs a;
  a.ntiles = (int)tf.ntiles;
  a.desc = tf.desc.p;
  a.rowid = tf.rowid.p;
  a.rptr = tf.rptr.p;
  a.ucol = tf.ucol.p;
  a.lidx = tf.lidx.p;
  a.tval = A.tval;
  a.mask = A.mask;
  a.x = x;
  a.y = y;
  a.w = w;
  a.partials = dv->partials;
  a.counter = dv->counter;
  a.red = dv->red;
  a.slot = slot;
  a.flag = flag;
  switch (A.ncomp) {
    case 1: return launch_tile_nc<1>(ctx, a, dot_mode);
    case 2: return launch_tile_nc<2>(ctx, a, dot_mode);
    case 3: return launch_tile_nc<3>(ctx, a, dot_mode);
    default: throw fb_cuda_error(FB_EINVAL, "tile_spmm: ncomp must be 1..3");
  }
}
