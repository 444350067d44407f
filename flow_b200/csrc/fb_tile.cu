// Tile-CSR SpMM for the scalar Lagrange operators (P2 mass M, S = M + theta dt nu K, P1/P2 stiffness, heat operator)
// applied to NC interleaved components: the PETSc MatMult of the reference's CG solves
// (pressure_correction.py:451-464 velocity correction, and the inner solves of the momentum preconditioner).
//
// Why a second format.  The row-wise CSR kernel (k_spmm_u) gathers x[col] straight from global memory: ~29 gathers of
// 24 bytes per row, each a separate 32-byte sector.  ncu showed it bound by the L1 data pipe (L1/TEX 89 % busy, DRAM
// 41 %), not by HBM.  Here the rows are grouped into TILES of spatially close nodes (Morton order of the node
// coordinates, a private row order of this format -- x and y keep the canonical numbering of the ABI):
//   * the union of the columns a tile touches (3.5 - 4.5 nodes per row instead of 29) is gathered ONCE into shared
//     memory with cp.async (LDGSTS, 16 + 8 bytes per node), one tile ahead;
//   * entries are (fp64 value, 16-bit slot in that union): 10 bytes instead of 12.  Inside a tile the rows are sorted
//     by length and stored in groups of 8 rows x 4 lanes, step-major (SELL-8x4): step s of a group is 32 consecutive
//     entries, one per lane of the warp that owns the group -> the value / index streams are perfectly coalesced and
//     a warp's trip count is uniform;
//   * the entries of each row are ordered so that the 32 shared-memory gathers of one step spread evenly over the
//     16 bank pairs (see tile_format_build_host): ~2.3 wavefronts per gather instruction instead of 4.5;
//   * the entry stream reaches the SM either through the TMA unit (cp.async.bulk + mbarrier into a shared-memory
//     stage, evict-first in L2) or by coalesced streaming loads (ld.global.cs) -- TileCfg::TMA.  Measured on B200
//     (profiles/r2_tile_spmm_*.txt): the TMA writes into shared memory occupy the same L1 data pipe as the loads that
//     read them back, so which one wins is decided by measurement, not by principle.
// Algorithmic bytes stay those of scalar CSR.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "fb_ops.h"

namespace {

constexpr int TILE_T = 4;      // lanes per row
constexpr int TILE_G = 8;      // rows per group (one warp)
constexpr int TILE_DESC = 8;   // ints per tile descriptor: row0, nr, e0, ne, u0, nu, g0, ng

// Kernel configurations.  A tile is one pass of the block: at most THREADS / TILE_T rows, one warp per group of 8.
template <int ID>
struct TileCfg;
template <>
struct TileCfg<0> {  // TMA-streamed entries, two CTAs per SM, one tile ahead each
  static constexpr int THREADS = 512, ECAP = 3328, UCAP = 768, NSTAGE = 2, CTAS = 2, TMA = 1;
};
template <>
struct TileCfg<1> {  // streaming loads, two CTAs per SM
  static constexpr int THREADS = 512, ECAP = 1 << 20, UCAP = 768, NSTAGE = 2, CTAS = 2, TMA = 0;
};
template <>
struct TileCfg<2> {  // streaming loads, three CTAs per SM
  static constexpr int THREADS = 512, ECAP = 1 << 20, UCAP = 768, NSTAGE = 2, CTAS = 3, TMA = 0;
};
template <>
struct TileCfg<3> {  // streaming loads, one big CTA per SM (smallest column unions)
  static constexpr int THREADS = 1024, ECAP = 1 << 20, UCAP = 1280, NSTAGE = 2, CTAS = 1, TMA = 0;
};
template <>
struct TileCfg<4> {  // streaming loads, six small CTAs per SM
  static constexpr int THREADS = 256, ECAP = 1 << 20, UCAP = 448, NSTAGE = 2, CTAS = 6, TMA = 0;
};
template <>
struct TileCfg<5> {  // TMA-streamed entries, one big CTA per SM
  static constexpr int THREADS = 1024, ECAP = 6656, UCAP = 1280, NSTAGE = 2, CTAS = 1, TMA = 1;
};
constexpr int TILE_NCFG = 6;

struct TileCaps {
  int threads, ecap, ucap, nstage, ctas, tma;
};
template <int ID>
constexpr TileCaps caps_of() {
  return TileCaps{TileCfg<ID>::THREADS, TileCfg<ID>::ECAP, TileCfg<ID>::UCAP, TileCfg<ID>::NSTAGE, TileCfg<ID>::CTAS, TileCfg<ID>::TMA};
}
inline TileCaps tile_caps(int id) {
  switch (id) {
    case 0: return caps_of<0>();
    case 1: return caps_of<1>();
    case 2: return caps_of<2>();
    case 3: return caps_of<3>();
    case 4: return caps_of<4>();
    default: return caps_of<5>();
  }
}

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
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// the NC values of node `col` -> xs[slot * NC ..]; NC = 3: 24 bytes at a multiple of 24, i.e. one 16-byte and one
// 8-byte piece whose order depends on the parity of the node; slot parity == column parity by construction
template <int NC>
__device__ __forceinline__ void gather_node(double *xs, int slot, const double *x, int col) {
  double *dst = xs + slot * NC;
  const double *src = x + (int64_t)col * NC;
  if (NC == 3) {
    if (col & 1) {
      cp_async8(dst, src);
      cp_async16(dst + 1, src + 1);
    } else {
      cp_async16(dst, src);
      cp_async8(dst + 2, src + 2);
    }
  } else if (NC == 2) {
    cp_async16(dst, src);
  } else {
    cp_async8(dst, src);
  }
}

struct TileArgs {
  int ntiles;
  const int *desc;          // ntiles * TILE_DESC
  const int *rowid;         // tile-order row -> canonical row
  const int *gptr;          // per tile ng + 1 entry offsets of its row groups, relative to e0 (multiples of 32)
  const int *ucol;          // union column lists (-1: unused slot)
  const uint16_t *lidx;     // entries: slot in the tile's union
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
  // Chebyshev epilogue (DOT == 4): x = d_k; r' = r - A d_k; d' = c_dd d_k + c_r dinv r'; z' = z + d_k (+ d' in the
  // last step); y is not written.  The vectors that only ever appear as ROW operands of the recurrence -- r, z, dinv --
  // live in TILE ORDER (entry (row0 + rl) * NC + c: the rows of a tile are contiguous, the accesses coalesce); d is the
  // gather source of the next product and keeps the canonical numbering, like the input (ch_rin of the first step:
  // ch_rin_canonical) and the result (ch_zout of the last step).
  const double *ch_rin = nullptr, *ch_zin = nullptr, *ch_dinv = nullptr;
  double *ch_rout = nullptr, *ch_zout = nullptr, *ch_dout = nullptr;
  double ch_cdd = 0.0, ch_cr = 0.0;
  int ch_last = 0, ch_rin_canonical = 0;
};

template <int NC, class CFG>
struct TileSmem {
  static constexpr int VAL_BYTES = CFG::TMA ? CFG::ECAP * 8 : 0;
  static constexpr int X_BYTES = CFG::UCAP * NC * 8;
  static constexpr int IDX_BYTES = CFG::TMA ? CFG::ECAP * 2 : 0;
  static constexpr int STAGE = (VAL_BYTES + X_BYTES + IDX_BYTES + 127) / 128 * 128;
  static constexpr int TOTAL = CFG::NSTAGE * STAGE + 64;
};

// DOT: 0 none; 1: red[slot] = w.y; 2: + red[slot+1] = y.y; 3: (w.x, y.x, x.x) -- as k_spmm_u
template <int NC, int DOT, class CFG>
__global__ void __launch_bounds__(CFG::THREADS, CFG::CTAS) k_tile_spmm(const TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = TileSmem<NC, CFG>;
  constexpr int NS = CFG::NSTAGE, THREADS = CFG::THREADS;
  constexpr bool TMA = CFG::TMA != 0;
  static_assert(CFG::UCAP <= 2 * THREADS, "two union columns per thread");
  if (a.flag && *a.flag) return;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + NS * L::STAGE);
  const int tid = threadIdx.x;
  if (TMA) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  uint64_t policy = 0;
  if (TMA) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  auto stage_val = [&](int s) { return reinterpret_cast<double *>(smem + s * L::STAGE); };
  auto stage_x = [&](int s) { return reinterpret_cast<double *>(smem + s * L::STAGE + L::VAL_BYTES); };
  auto stage_idx = [&](int s) { return reinterpret_cast<uint16_t *>(smem + s * L::STAGE + L::VAL_BYTES + L::X_BYTES); };

  // union columns of a tile, two per thread, fetched into registers one iteration before they are used
  auto load_ucols = [&](int t, int &c0, int &c1) {
    c0 = c1 = -1;
    if (t < a.ntiles) {
      const int u0 = a.desc[t * TILE_DESC + 4], nu = a.desc[t * TILE_DESC + 5];
      if (tid < nu) c0 = a.ucol[u0 + tid];
      if (tid + THREADS < nu) c1 = a.ucol[u0 + tid + THREADS];
    }
  };
  // start the loads of tile t into stage s: TMA for the entries (if staged), cp.async gather for the x union
  auto issue = [&](int t, int s, int c0, int c1) {
    if (t < a.ntiles) {
      if (TMA && tid == 0) {
        const int e0 = a.desc[t * TILE_DESC + 2], ne = a.desc[t * TILE_DESC + 3];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this stage are done
        mbar_expect_tx(&bar[s], (uint32_t)ne * 10u);
        tma_bulk_load(stage_val(s), a.tval + e0, (uint32_t)ne * 8u, &bar[s], policy);
        tma_bulk_load(stage_idx(s), a.lidx + e0, (uint32_t)ne * 2u, &bar[s], policy);
      }
      double *xs = stage_x(s);
      if (c0 >= 0) gather_node<NC>(xs, tid, a.x, c0);
      if (c1 >= 0) gather_node<NC>(xs, tid + THREADS, a.x, c1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double d[3] = {0.0, 0.0, 0.0};
  const int G = gridDim.x;
  int t = blockIdx.x;
  int c0, c1;
  // prologue: NS - 1 tiles in flight
#pragma unroll
  for (int k = 0; k < NS - 1; ++k) {
    load_ucols(t + k * G, c0, c1);
    issue(t + k * G, k, c0, c1);
  }
  load_ucols(t + (NS - 1) * G, c0, c1);
  const int lane32 = tid & 31, warp = tid >> 5, lane = tid % TILE_T, rl = tid / TILE_T;
  for (int it = 0; t < a.ntiles; t += G, ++it) {
    const int s = it % NS;
    // NS - 1 tiles ahead: entries by TMA, x union by cp.async; NS tiles ahead: the union column ids into registers
    issue(t + (NS - 1) * G, (it + NS - 1) % NS, c0, c1);
    load_ucols(t + NS * G, c0, c1);
    // this tile's row data (global, coalesced) while its stage completes
    const int *dsc = a.desc + t * TILE_DESC;
    const int row0 = dsc[0], nr = dsc[1], e0 = dsc[2], g0 = dsc[6], ng = dsc[7];
    int kb = 0, ke = 0, row = -1;
    if (warp < ng) {
      kb = a.gptr[g0 + warp];
      ke = a.gptr[g0 + warp + 1];
    }
    if (rl < nr) row = a.rowid[row0 + rl];
    double wv = 0.0, xd = 0.0, dinv = 0.0, zin = 0.0;  // all row operands of the epilogue are in flight before the wait
    bool masked = false;
    if (row >= 0 && lane < NC) {
      const int64_t dof = (int64_t)row * NC + lane;
      if (DOT >= 1 && DOT <= 3) wv = a.w[dof];
      if (DOT == 4) {
        const int64_t tdof = (int64_t)(row0 + rl) * NC + lane;
        wv = a.ch_rin[a.ch_rin_canonical ? dof : tdof];
        dinv = a.ch_dinv[tdof];
        if (a.ch_zin) zin = a.ch_zin[tdof];
      }
      if (DOT >= 3 || a.mask) xd = a.x[dof];
      if (a.mask) masked = a.mask[dof] != 0;
    }
    cp_async_wait<NS - 1>();
    __syncthreads();  // every thread's part of the x union has landed
    if (TMA) mbar_wait(&bar[s], (uint32_t)(it / NS) & 1u);
    const double *xs = stage_x(s);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    if (TMA) {
      const double *vs = stage_val(s);
      const uint16_t *is = stage_idx(s);
#pragma unroll 4
      for (int k = kb + lane32; k < ke; k += 32) {
        const double av = vs[k];
        const int j = is[k];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += av * xs[j * NC + c];
      }
    } else {
      const double *vg = a.tval + e0;
      const uint16_t *ig = a.lidx + e0;
#pragma unroll 4
      for (int k = kb + lane32; k < ke; k += 32) {
        const double av = __ldcs(vg + k);
        const int j = __ldcs(ig + k);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += av * xs[j * NC + c];
      }
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
      if (DOT == 4) {
        const int64_t dof = (int64_t)row * NC + lane, tdof = (int64_t)(row0 + rl) * NC + lane;
        const double rn = wv - yc;
        const double dn = a.ch_cdd * xd + a.ch_cr * dinv * rn;
        double z = zin + xd;
        if (a.ch_last) {
          a.ch_zout[dof] = z + dn;
        } else {
          a.ch_zout[tdof] = z;
          a.ch_rout[tdof] = rn;
          a.ch_dout[dof] = dn;
        }
      } else {
        a.y[(int64_t)row * NC + lane] = yc;
      }
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
  cp_async_wait<0>();
  if (DOT >= 1 && DOT <= 3) {
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

template <int NC, int DOT, class CFG>
void launch_tile(fb_ctx *ctx, const TileArgs &a) {
  using L = TileSmem<NC, CFG>;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(k_tile_spmm<NC, DOT, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int grid = std::min(a.ntiles, ctx->dev->sm_count * CFG::CTAS);
  k_tile_spmm<NC, DOT, CFG><<<grid, CFG::THREADS, L::TOTAL, ctx->dev->stream>>>(a);
  ctx->launches++;
}

template <int NC, class CFG>
void launch_tile_dot(fb_ctx *ctx, const TileArgs &a, int dot_mode) {
  if (dot_mode == 0) launch_tile<NC, 0, CFG>(ctx, a);
  else if (dot_mode == 1) launch_tile<NC, 1, CFG>(ctx, a);
  else if (dot_mode == 2) launch_tile<NC, 2, CFG>(ctx, a);
  else if (dot_mode == 3) launch_tile<NC, 3, CFG>(ctx, a);
  else launch_tile<NC, 4, CFG>(ctx, a);
}

template <int NC>
void launch_tile_nc(fb_ctx *ctx, const TileArgs &a, int dot_mode, int cfg) {
  switch (cfg) {
    case 0: return launch_tile_dot<NC, TileCfg<0>>(ctx, a, dot_mode);
    case 1: return launch_tile_dot<NC, TileCfg<1>>(ctx, a, dot_mode);
    case 2: return launch_tile_dot<NC, TileCfg<2>>(ctx, a, dot_mode);
    case 3: return launch_tile_dot<NC, TileCfg<3>>(ctx, a, dot_mode);
    case 4: return launch_tile_dot<NC, TileCfg<4>>(ctx, a, dot_mode);
    default: return launch_tile_dot<NC, TileCfg<5>>(ctx, a, dot_mode);
  }
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

// kernel configuration of the formats built from now on (FB_TILE_CFG, experiments; default: see TileCfg)
static int tile_default_cfg() {
  const char *e = getenv("FB_TILE_CFG");
  const int c = e ? atoi(e) : 0;  // measured best on B200 at n = 74 (profiles/r2_tile_spmm_configs.txt)
  return (c >= 0 && c < TILE_NCFG) ? c : 0;
}

// Host side of the format: tiles of the owned rows of s's node pattern (pure host code, once per space).
struct HostTile {
  std::vector<int> desc, order, gptr, ucol, src;
  std::vector<uint16_t> lidx;
  int ntiles = 0;
};

static void tile_format_build_host(fb_space *s, HostTile &h, const TileCaps &cap) {
  const int TILE_RCAP = cap.threads / TILE_T, TILE_ECAP = cap.ecap, TILE_UCAP = cap.ucap;
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

  std::vector<int> &desc = h.desc, &gptr = h.gptr, &ucol = h.ucol, &src = h.src;
  std::vector<uint16_t> &lidx = h.lidx;
  auto rowlen = [&](int r) { return (int)(ip[r + 1] - ip[r]); };
  // entries of a tile whose rows (sorted by length, descending) are order[row0 .. row0 + nr): groups of 8 rows, each
  // padded to the steps of its longest row
  auto tile_entries = [&](int64_t row0, int nr) {
    int total = 0;
    for (int i = 0; i < nr; i += TILE_G) total += 32 * ((rowlen(order[row0 + i]) + TILE_T - 1) / TILE_T);
    return total;
  };
  // ---- phase 1 (sequential): tiles = runs of rows in Morton order, grown until one of the caps binds
  std::vector<int> mark((size_t)nn, -1);
  std::vector<int> ulist;
  ulist.reserve(TILE_UCAP);
  int64_t pos = 0, etotal = 0;
  int tile = 0, stamp = 0;
  while (pos < n) {
    const int64_t row0 = pos;
    int rcap = TILE_RCAP, nr = 0, n_even = 0, n_odd = 0;
    for (;;) {  // (re)grow with a smaller row cap until the padded entries fit the stage
      ulist.clear();
      nr = n_even = n_odd = 0;
      pos = row0;
      int ne = 0;
      ++stamp;
      while (pos < n && nr < rcap) {
        const int r = order[pos];
        const int len = rowlen(r);
        if (len > TILE_UCAP || len > 256 || 32 * ((len + 3) / 4) > TILE_ECAP) throw fb_cuda_error(FB_EINVAL, "tile format: matrix row too long");
        if (ne + ((len + 3) & ~3) > TILE_ECAP) break;
        // slots keep the parity of their column (16-byte gathers, see gather_node): the union occupies
        // 2 * max(#even, #odd) slots
        int fe = 0, fo = 0;
        for (int64_t k = ip[r]; k < ip[r + 1]; ++k)
          if (mark[ix[k]] != stamp) (ix[k] & 1) ? ++fo : ++fe;
        if (2 * std::max(n_even + fe, n_odd + fo) > TILE_UCAP) break;
        n_even += fe;
        n_odd += fo;
        for (int64_t k = ip[r]; k < ip[r + 1]; ++k)
          if (mark[ix[k]] != stamp) {
            mark[ix[k]] = stamp;
            ulist.push_back(ix[k]);
          }
        ne += (len + 3) & ~3;
        ++nr;
        ++pos;
      }
      // rows of similar length next to each other: a warp runs the trip count of the longest of its 8 rows
      // (vertex rows of a P2 pattern are ~3x longer than edge rows)
      std::stable_sort(order.begin() + row0, order.begin() + row0 + nr, [&](int a, int b) { return rowlen(a) > rowlen(b); });
      if (tile_entries(row0, nr) <= TILE_ECAP || nr <= TILE_G) break;
      // undo the sort's effect on later attempts: the run is re-sorted by Morton position implicitly because the
      // retry takes a PREFIX of the same rows only if they are back in Morton order
      std::stable_sort(order.begin() + row0, order.begin() + row0 + nr, [&](int a, int b) { return a < b; });
      rcap = std::max(TILE_G, (nr - TILE_G) & ~(TILE_G - 1));
    }
    std::sort(ulist.begin(), ulist.end());
    {  // even columns -> even slots, odd columns -> odd slots (ascending within each class), -1 in unused slots
      std::vector<int> slots((size_t)2 * std::max(n_even, n_odd), -1);
      int se = 0, so = 1;
      for (int c : ulist) {
        if (c & 1) {
          slots[so] = c;
          so += 2;
        } else {
          slots[se] = c;
          se += 2;
        }
      }
      ulist.swap(slots);
    }
    const int ng = (nr + TILE_G - 1) / TILE_G;
    const int g0 = (int)gptr.size();
    int off = 0;
    for (int g = 0; g < ng; ++g) {
      gptr.push_back(off);
      off += 32 * ((rowlen(order[row0 + g * TILE_G]) + TILE_T - 1) / TILE_T);
    }
    gptr.push_back(off);
    const int d8[TILE_DESC] = {(int)row0, nr, (int)etotal, off, (int)ucol.size(), (int)ulist.size(), g0, ng};
    desc.insert(desc.end(), d8, d8 + TILE_DESC);
    ucol.insert(ucol.end(), ulist.begin(), ulist.end());
    etotal += off;
    ++tile;
    if (etotal > (int64_t)INT32_MAX - (1 << 22)) throw fb_cuda_error(FB_EINVAL, "tile format: too many entries");
  }
  mark.clear();
  mark.shrink_to_fit();
  lidx.assign((size_t)etotal, 0);
  src.assign((size_t)etotal, -1);
  // ---- phase 2 (parallel over tiles): entries, step-major inside each group of 8 rows: entry (step s, row i, lane l)
  // of a group sits at 32 s + 4 i + l, i.e. at the lane of the warp that reads it.
  // Entry order inside the rows.  At step s lane l of row i gathers x from shared memory at slot * NC * 8 bytes.
  // Shared memory serves 16 distinct 8-byte bank pairs per wavefront, and (NC * slot + c) mod 16 is a bijection of
  // slot mod 16 for NC = 1, 3, so the 32 gathers of one instruction need the minimum of 2 wavefronts iff every residue
  // class slot mod 16 is used exactly once per half-warp (64-bit accesses are served half-warp by half-warp).  The sum
  // over a row does not depend on the order of its entries: the rows of a group are emitted step by step, every lane
  // taking the remaining entry of its row whose class is used least so far in its half-warp (CSR order: 3.8-way
  // conflicts measured).  Padding entries: zero value, a gathered slot.
#pragma omp parallel
  {
    std::vector<uint16_t> lid((size_t)nn, 0);
#pragma omp for schedule(dynamic, 16)
    for (int t = 0; t < tile; ++t) {
      const int *d = &desc[(size_t)t * TILE_DESC];
      const int row0 = d[0], nr = d[1], e0 = d[2], u0 = d[4], nu = d[5], g0 = d[6], ng = d[7];
      int any_slot = 0;
      for (int j = 0; j < nu; ++j)
        if (ucol[u0 + j] >= 0) {
          lid[ucol[u0 + j]] = (uint16_t)j;
          any_slot = j;
        }
      for (int g = 0; g < ng; ++g) {
        const int i0 = g * TILE_G, i1 = std::min(nr, i0 + TILE_G);
        const int gb = gptr[g0 + g], nsteps = (gptr[g0 + g + 1] - gb) / 32;
        int ent[TILE_G][256], cls[TILE_G][256], left[TILE_G], pad_slot[TILE_G];
        for (int q = 0; q < TILE_G; ++q) {
          left[q] = 0;
          pad_slot[q] = any_slot;
          if (i0 + q >= i1) continue;
          const int r = order[row0 + i0 + q];
          int m = 0;
          for (int64_t k = ip[r]; k < ip[r + 1]; ++k, ++m) {
            ent[q][m] = (int)k;
            cls[q][m] = lid[ix[k]] & 15;
          }
          left[q] = m;
          if (m) pad_slot[q] = lid[ix[ip[r]]];
        }
        for (int st = 0; st < nsteps; ++st) {
          int count[16] = {0};
          for (int q = 0; q < TILE_G; ++q) {
            // 64-bit shared loads are served per half-warp (16 lanes = 4 rows): the classes are balanced within each
            if (q == TILE_G / 2) std::fill(count, count + 16, 0);
            for (int l = 0; l < TILE_T; ++l) {
              const size_t at = (size_t)e0 + gb + 32 * st + TILE_T * q + l;
              if (left[q] == 0) {  // padding
                lidx[at] = (uint16_t)pad_slot[q];
                src[at] = -1;
                count[pad_slot[q] & 15]++;
                continue;
              }
              int best = 0;
              for (int m = 1; m < left[q]; ++m)
                if (count[cls[q][m]] < count[cls[q][best]]) best = m;
              const int k = ent[q][best];
              lidx[at] = lid[ix[k]];
              src[at] = k;
              count[cls[q][best]]++;
              ent[q][best] = ent[q][left[q] - 1];
              cls[q][best] = cls[q][left[q] - 1];
              --left[q];
            }
          }
        }
      }
    }
  }
  h.ntiles = tile;
}

void tile_format_build(fb_space *s, TileFormat &tf, cudaStream_t st) {
  HostTile h;
  tf.cfg = tile_default_cfg();
  tile_format_build_host(s, h, tile_caps(tf.cfg));
  tf.nrows = s->n_owned;
  tf.ntiles = h.ntiles;
  tf.nent = (int64_t)h.lidx.size();
  tf.union_total = (int64_t)h.ucol.size();
  tf.desc.upload(h.desc.data(), h.desc.size(), st);
  tf.rowid.upload(h.order.data(), h.order.size(), st);
  tf.gptr.upload(h.gptr.data(), h.gptr.size(), st);
  tf.ucol.upload(h.ucol.data(), h.ucol.size(), st);
  tf.lidx.upload(h.lidx.data(), h.lidx.size(), st);
  tf.src.upload(h.src.data(), h.src.size(), st);
  FB_CUDA(cudaStreamSynchronize(st));  // the host vectors are temporaries
}

// Host-only self check (works in a context without a device): build the format and verify that it is the CSR pattern
// -- every owned row exactly once, each row's entries a permutation of its CSR row (ucol[lidx] == CSR column,
// src == CSR slot) at the positions its lanes read, padding only zero-valued, caps and TMA alignment respected.
// stats: ntiles, entries incl. padding, sum of union slots, max rows, max entries, max union, row / entry / union caps.
extern "C" int fb_space_tile_check(fb_space *s, int64_t *stats) {
  if (!s) return FB_EINVAL;
  try {
    HostTile h;
    const TileCaps cap = tile_caps(tile_default_cfg());
    const int TILE_RCAP = cap.threads / TILE_T, TILE_ECAP = cap.ecap, TILE_UCAP = cap.ucap;
    tile_format_build_host(s, h, cap);
    const int64_t n = s->n_owned;
    fb_ctx *ctx = s->mesh->ctx;
    std::vector<uint8_t> seen((size_t)n, 0);
    int64_t max_r = 0, max_e = 0, max_u = 0, rows = 0;
    for (int t = 0; t < h.ntiles; ++t) {
      const int *d = &h.desc[(size_t)t * TILE_DESC];
      const int row0 = d[0], nr = d[1], e0 = d[2], ne = d[3], u0 = d[4], nu = d[5], g0 = d[6], ng = d[7];
      if (nr < 1 || nr > TILE_RCAP || ne > TILE_ECAP || nu > TILE_UCAP || (e0 & 31) || (ne & 31)) return fb_fail(ctx, FB_EINVAL, "tile check: caps/alignment");
      if (row0 != rows || ng != (nr + TILE_G - 1) / TILE_G || h.gptr[g0] != 0 || h.gptr[g0 + ng] != ne) return fb_fail(ctx, FB_EINVAL, "tile check: layout");
      for (int j = 0; j < nu; ++j) {  // slot parity == column parity, ascending within each parity class, gaps are -1
        const int c = h.ucol[u0 + j];
        if (c >= 0 && ((c ^ j) & 1)) return fb_fail(ctx, FB_EINVAL, "tile check: slot parity");
        if (c >= 0 && j >= 2 && h.ucol[u0 + j - 2] >= c) return fb_fail(ctx, FB_EINVAL, "tile check: union order");
      }
      for (int i = 0; i < nr; ++i) {
        const int r = h.order[row0 + i];
        if (r < 0 || r >= n || seen[r]) return fb_fail(ctx, FB_EINVAL, "tile check: row covered twice");
        seen[r] = 1;
        const int g = i / TILE_G, q = i % TILE_G;
        const int gb = h.gptr[g0 + g], nsteps = (h.gptr[g0 + g + 1] - gb) / 32;
        const int len = (int)(s->indptr[r + 1] - s->indptr[r]);
        std::vector<int> slots_seen;
        for (int st = 0; st < nsteps; ++st)
          for (int l = 0; l < TILE_T; ++l) {
            const size_t at = (size_t)e0 + gb + 32 * st + TILE_T * q + l;
            const int sl = h.src[at], li = h.lidx[at];
            if (li >= nu || h.ucol[u0 + li] < 0) return fb_fail(ctx, FB_EINVAL, "tile check: entry points at an empty slot");
            if (sl < 0) continue;  // padding: zero value, any gathered slot
            if (sl < s->indptr[r] || sl >= s->indptr[r + 1] || h.ucol[u0 + li] != s->indices[sl]) return fb_fail(ctx, FB_EINVAL, "tile check: entry mismatch");
            slots_seen.push_back(sl);
          }
        std::sort(slots_seen.begin(), slots_seen.end());
        if ((int)slots_seen.size() != len || std::adjacent_find(slots_seen.begin(), slots_seen.end()) != slots_seen.end())
          return fb_fail(ctx, FB_EINVAL, "tile check: row entries are not a permutation of the CSR row");
      }
      // lanes of rows that do not exist (last group of the tile) only hold padding
      for (int i = nr; i < ng * TILE_G; ++i) {
        const int g = i / TILE_G, q = i % TILE_G;
        const int gb = h.gptr[g0 + g], nsteps = (h.gptr[g0 + g + 1] - gb) / 32;
        for (int st = 0; st < nsteps; ++st)
          for (int l = 0; l < TILE_T; ++l) {
            const size_t at = (size_t)e0 + gb + 32 * st + TILE_T * q + l;
            if (h.src[at] >= 0 || h.lidx[at] >= nu || h.ucol[u0 + h.lidx[at]] < 0) return fb_fail(ctx, FB_EINVAL, "tile check: padding lane");
          }
      }
      rows += nr;
      max_r = std::max<int64_t>(max_r, nr);
      max_e = std::max<int64_t>(max_e, ne);
      max_u = std::max<int64_t>(max_u, nu);
    }
    for (int64_t i = 0; i < n; ++i)
      if (!seen[i]) return fb_fail(ctx, FB_EINVAL, "tile check: row missing");
    if (stats) {
      stats[0] = h.ntiles;
      stats[1] = (int64_t)h.lidx.size();
      stats[2] = (int64_t)h.ucol.size();
      stats[3] = max_r;
      stats[4] = max_e;
      stats[5] = max_u;
      stats[6] = TILE_RCAP;
      stats[7] = TILE_ECAP;
      stats[8] = TILE_UCAP;
    }
  } catch (const std::exception &e) {
    return fb_fail(s->mesh->ctx, FB_EINVAL, e.what());
  }
  return FB_OK;
}

namespace {
__global__ void k_tile_order(int64_t nrows, int ncomp, const int *__restrict__ rowid, const double *__restrict__ x, double *__restrict__ xt) {
  const int64_t total = nrows * ncomp;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / ncomp;
    xt[t] = x[(int64_t)rowid[q] * ncomp + (t - q * ncomp)];
  }
}
}  // namespace

// xt = x in the tile order of the format's rows (owned rows only)
void tile_to_tile_order(fb_ctx *ctx, const TileFormat &tf, int ncomp, const double *x, double *xt) {
  const int64_t n = (int64_t)tf.rowid.n * ncomp;
  int64_t g = (n + 255) / 256;
  g = std::max<int64_t>(1, std::min<int64_t>(g, ctx->dev->sm_count * 8));
  FB_LAUNCH(ctx, k_tile_order, (int)g, 256, 0, (int64_t)tf.rowid.n, ncomp, tf.rowid.p, x, xt);
}

void tile_pack(fb_ctx *ctx, const TileFormat &tf, const double *val, double *tval) {
  const int64_t n = tf.nent;
  int64_t g = (n + 255) / 256;
  g = std::max<int64_t>(1, std::min<int64_t>(g, ctx->dev->sm_count * 8));
  FB_LAUNCH(ctx, k_tile_pack, (int)g, 256, 0, n, tf.src.p, val, tval);
}

static TileArgs tile_args(fb_ctx *ctx, const LinOp &A, const double *x, double *y, const double *w, int slot, const int *flag) {
  const TileFormat &tf = *A.tile;
  fb_device_state *dv = ctx->dev;
  TileArgs a;
  a.ntiles = (int)tf.ntiles;
  a.desc = tf.desc.p;
  a.rowid = tf.rowid.p;
  a.gptr = tf.gptr.p;
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
  return a;
}

static void tile_launch(fb_ctx *ctx, const LinOp &A, const TileArgs &a, int dot_mode) {
  switch (A.ncomp) {
    case 1: return launch_tile_nc<1>(ctx, a, dot_mode, A.tile->cfg);
    case 2: return launch_tile_nc<2>(ctx, a, dot_mode, A.tile->cfg);
    case 3: return launch_tile_nc<3>(ctx, a, dot_mode, A.tile->cfg);
    default: throw fb_cuda_error(FB_EINVAL, "tile_spmm: ncomp must be 1..3");
  }
}

void tile_spmm(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
               const int *flag) {
  tile_launch(ctx, A, tile_args(ctx, A, x, y, w, slot, flag), dot_mode);
}

// One Chebyshev step fused into the product (see TileArgs): the caller refreshed the ghosts of d.
// rin: canonical numbering if rin_canonical (the preconditioner's input), tile order otherwise; rout, zin, dinv_t: tile
// order; zout: tile order, canonical in the last step; d, dout: canonical.
void tile_cheb_step(fb_ctx *ctx, const LinOp &A, const double *d, const double *rin, bool rin_canonical, double *rout,
                    const double *zin, double *zout, double *dout, const double *dinv_t, double cdd, double cr, bool last) {
  TileArgs a = tile_args(ctx, A, d, nullptr, nullptr, 0, nullptr);
  a.ch_rin = rin;
  a.ch_rout = rout;
  a.ch_zin = zin;
  a.ch_zout = zout;
  a.ch_dout = dout;
  a.ch_dinv = dinv_t;
  a.ch_rin_canonical = rin_canonical ? 1 : 0;
  a.ch_cdd = cdd;
  a.ch_cr = cr;
  a.ch_last = last ? 1 : 0;
  tile_launch(ctx, A, a, 4);
}
