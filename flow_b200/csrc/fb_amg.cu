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
#include <cstring>
#include <numeric>
#include <vector>

#include "fb_ops.h"

namespace {

struct HostCsr {
  int nrows = 0, ncols = 0;
  std::vector<int> ptr, col;
  std::vector<double> val;
};

HostCsr transpose(const HostCsr &A) {
  HostCsr T;
  T.nrows = A.ncols;
  T.ncols = A.nrows;
  T.ptr.assign(T.nrows + 1, 0);
  for (int c : A.col) T.ptr[c + 1]++;
  for (int i = 0; i < T.nrows; ++i) T.ptr[i + 1] += T.ptr[i];
  T.col.resize(A.col.size());
  T.val.resize(A.val.size());
  std::vector<int> fill(T.ptr.begin(), T.ptr.end() - 1);
  for (int i = 0; i < A.nrows; ++i)
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int p = fill[A.col[k]]++;
      T.col[p] = i;
      T.val[p] = A.val[k];
    }
  return T;
}

// C = A * B (Gustavson), columns sorted
HostCsr multiply(const HostCsr &A, const HostCsr &B) {
  HostCsr C;
  C.nrows = A.nrows;
  C.ncols = B.ncols;
  C.ptr.assign(A.nrows + 1, 0);
  std::vector<double> acc(B.ncols, 0.0);
  std::vector<int> mark(B.ncols, -1), cols;
  for (int i = 0; i < A.nrows; ++i) {
    cols.clear();
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      const int a = A.col[k];
      const double av = A.val[k];
      for (int l = B.ptr[a]; l < B.ptr[a + 1]; ++l) {
        const int j = B.col[l];
        if (mark[j] != i) {
          mark[j] = i;
          acc[j] = 0.0;
          cols.push_back(j);
        }
        acc[j] += av * B.val[l];
      }
    }
    std::sort(cols.begin(), cols.end());
    for (int j : cols) {
      C.col.push_back(j);
      C.val.push_back(acc[j]);
    }
    C.ptr[i + 1] = (int)C.col.size();
  }
  return C;
}

// greedy aggregation on the strength graph; returns aggregate id per node (-1: isolated, e.g. Dirichlet rows)
int aggregate(const HostCsr &A, double theta, std::vector<int> &agg) {
  const int n = A.nrows;
  std::vector<double> diag(n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (A.col[k] == i) diag[i] = std::fabs(A.val[k]);
  auto strong = [&](int i, int k) {
    const int j = A.col[k];
    return j != i && A.val[k] * A.val[k] > theta * theta * diag[i] * diag[j];
  };
  agg.assign(n, -1);
  std::vector<char> has_strong(n, 0);
  for (int i = 0; i < n; ++i)
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k)) {
        has_strong[i] = 1;
        break;
      }
  int nagg = 0;
  // pass 1: root nodes whose strong neighbourhood is entirely free
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1 || !has_strong[i]) continue;
    bool free_nbhd = true;
    for (int k = A.ptr[i]; k < A.ptr[i + 1] && free_nbhd; ++k)
      if (strong(i, k) && agg[A.col[k]] != -1) free_nbhd = false;
    if (!free_nbhd) continue;
    agg[i] = nagg;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k)) agg[A.col[k]] = nagg;
    ++nagg;
  }
  // pass 2: attach the rest to the aggregate of their strongest aggregated neighbour (as of pass 1)
  std::vector<int> agg1(agg);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1 || !has_strong[i]) continue;
    double best = 0.0;
    int who = -1;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k) && agg1[A.col[k]] != -1 && std::fabs(A.val[k]) > best) {
        best = std::fabs(A.val[k]);
        who = agg1[A.col[k]];
      }
    if (who != -1) agg[i] = who;
  }
  // pass 3: leftovers form aggregates with their free strong neighbours
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1 || !has_strong[i]) continue;
    agg[i] = nagg;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k) && agg[A.col[k]] == -1) agg[A.col[k]] = nagg;
    ++nagg;
  }
  return nagg;
}

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
bool spd_inverse(std::vector<double> &a, int n) {
  // A = L L^T (L stored in the lower triangle)
  for (int j = 0; j < n; ++j) {
    double d = a[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    a[(size_t)j * n + j] = d;
#pragma omp parallel for schedule(static)
    for (int i = j + 1; i < n; ++i) {
      double s = a[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
      a[(size_t)i * n + j] = s / d;
    }
  }
  // X = L^-1 (lower), column by column
  std::vector<double> li((size_t)n * n, 0.0);
#pragma omp parallel for schedule(dynamic, 8)
  for (int c = 0; c < n; ++c) {
    li[(size_t)c * n + c] = 1.0 / a[(size_t)c * n + c];
    for (int i = c + 1; i < n; ++i) {
      double s = 0.0;
      for (int k = c; k < i; ++k) s -= a[(size_t)i * n + k] * li[(size_t)k * n + c];
      li[(size_t)i * n + c] = s / a[(size_t)i * n + i];
    }
  }
  // A^-1 = L^-T L^-1
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
      for (int k = i; k < n; ++k) s += li[(size_t)k * n + i] * li[(size_t)k * n + j];
      a[(size_t)i * n + j] = s;
      a[(size_t)j * n + i] = s;
    }
  return true;
}

int lanes_for(const HostCsr &M) {
  const double mean = M.nrows > 0 ? (double)M.val.size() / M.nrows : 0.0;
  int at = mean <= 12.0 ? 4 : (mean <= 32.0 ? 8 : (mean <= 96.0 ? 16 : 32));
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
  ~fb_amg() {
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

constexpr int AMG_COARSE_TARGET = 400;   // coarsen until at most this many rows ...
constexpr int AMG_DENSE_MAX = 2500;      // ... and invert densely if the coarsest level has at most this many

// Build the hierarchy of the n x n leading block of the CSR matrix (rowptr, col, val) given on the host.
fb_amg *amg_setup(fb_ctx *ctx, int n, const int *rowptr, const int *col, const double *val) {
  cudaStream_t st = ctx->dev->stream;
  fb_amg *amg = new fb_amg();
  amg->ctx = ctx;
  HostCsr A;
  A.nrows = A.ncols = n;
  A.ptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] < n) {  // drop ghost columns: owned x owned block
        A.col.push_back(col[k]);
        A.val.push_back(val[k]);
      }
    A.ptr[i + 1] = (int)A.col.size();
  }
  // near-null-space vector: constants.  The operator is treated as singular if it annihilates them.
  std::vector<double> B(n, 1.0);
  {
    double worst = 0.0;
    for (int i = 0; i < n; ++i) {
      double s = 0.0, d = 0.0;
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        s += A.val[k];
        if (A.col[k] == i) d = std::fabs(A.val[k]);
      }
      if (d > 0.0) worst = std::max(worst, std::fabs(s) / d);
    }
    amg->singular = worst < 1e-10;
  }
  const double nnz0 = (double)A.val.size();
  double nnz_total = 0.0;
  double theta = 0.08;
  for (int lev = 0; lev < 12; ++lev) {
    AmgLevel *L = new AmgLevel();
    amg->levels.push_back(L);
    L->n = A.nrows;
    nnz_total += (double)A.val.size();
    std::vector<double> dinv(A.nrows, 1.0);
    double rho = 0.0;  // Gershgorin bound of rho(D^-1 A)
    for (int i = 0; i < A.nrows; ++i) {
      double d = 0.0, s = 0.0;
      for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        if (A.col[k] == i) d = A.val[k];
        s += std::fabs(A.val[k]);
      }
      if (d > 0.0) {
        dinv[i] = 1.0 / d;
        rho = std::max(rho, s / d);
      }
    }
    if (rho <= 0.0) rho = 2.0;
    L->omega = 4.0 / (3.0 * rho);
    L->A.upload(A, st);
    L->lanes_A = lanes_for(A);
    L->dinv.upload(dinv.data(), dinv.size(), st);
    for (DBuf<double> *v : {&L->x, &L->b, &L->r, &L->tmp}) v->alloc((size_t)std::max(1, A.nrows));
    bool last = (A.nrows <= AMG_COARSE_TARGET || lev == 11);
    std::vector<int> agg;
    int nagg = 0;
    if (!last) {
      nagg = aggregate(A, theta, agg);
      theta *= 0.5;
      if (nagg == 0 || nagg > 0.8 * A.nrows) last = true;  // coarsening stalled: this level is the coarsest
    }
    if (last) {
      if (A.nrows <= AMG_DENSE_MAX && lev > 0) {
        const int m = A.nrows;
        std::vector<double> dense((size_t)m * m, 0.0);
        for (int i = 0; i < m; ++i)
          for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) dense[(size_t)i * m + A.col[k]] = A.val[k];
        for (int i = 0; i < m; ++i)  // symmetrise the rounding of the triple products
          for (int j = 0; j < i; ++j) {
            const double s = 0.5 * (dense[(size_t)i * m + j] + dense[(size_t)j * m + i]);
            dense[(size_t)i * m + j] = dense[(size_t)j * m + i] = s;
          }
        double shift = 0.0, bb = 0.0;
        std::vector<double> w(B);
        if (amg->singular) {
          for (int i = 0; i < m; ++i) {
            shift += dense[(size_t)i * m + i];
            bb += w[i] * w[i];
          }
          shift /= m;
          for (int i = 0; i < m; ++i) w[i] /= std::sqrt(bb);
          for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) dense[(size_t)i * m + j] += shift * w[i] * w[j];
        }
        // rows without a diagonal (isolated) get a unit diagonal so that the factorisation exists
        for (int i = 0; i < m; ++i)
          if (!(dense[(size_t)i * m + i] > 0.0)) dense[(size_t)i * m + i] = 1.0;
        if (spd_inverse(dense, m)) {
          if (amg->singular)
            for (int i = 0; i < m; ++i)
              for (int j = 0; j < m; ++j) dense[(size_t)i * m + j] -= w[i] * w[j] / shift;
          L->Ainv.upload(dense.data(), dense.size(), st);
          FB_CUDA(cudaStreamSynchronize(st));
        }
      }
      break;
    }
    // tentative prolongator: column a carries B restricted to aggregate a, normalised; coarse B = the norms
    std::vector<double> nrm(nagg, 0.0);
    for (int i = 0; i < A.nrows; ++i)
      if (agg[i] >= 0) nrm[agg[i]] += B[i] * B[i];
    for (double &v : nrm) v = std::sqrt(v);
    HostCsr P0;
    P0.nrows = A.nrows;
    P0.ncols = nagg;
    P0.ptr.assign(A.nrows + 1, 0);
    for (int i = 0; i < A.nrows; ++i) {
      if (agg[i] >= 0 && nrm[agg[i]] > 0.0) {
        P0.col.push_back(agg[i]);
        P0.val.push_back(B[i] / nrm[agg[i]]);
      }
      P0.ptr[i + 1] = (int)P0.col.size();
    }
    // P = (I - w D^-1 A) P0
    HostCsr AP0 = multiply(A, P0);
    HostCsr P;
    P.nrows = A.nrows;
    P.ncols = nagg;
    P.ptr.assign(A.nrows + 1, 0);
    for (int i = 0; i < A.nrows; ++i) {
      // merge row i of P0 (0 or 1 entry) with -w dinv_i * row i of AP0
      const int pc = (P0.ptr[i + 1] > P0.ptr[i]) ? P0.col[P0.ptr[i]] : -1;
      const double pv = (pc >= 0) ? P0.val[P0.ptr[i]] : 0.0;
      bool placed = (pc < 0);
      for (int k = AP0.ptr[i]; k < AP0.ptr[i + 1]; ++k) {
        const int j = AP0.col[k];
        double v = -L->omega * dinv[i] * AP0.val[k];
        if (!placed && pc < j) {
          P.col.push_back(pc);
          P.val.push_back(pv);
          placed = true;
        }
        if (j == pc) {
          v += pv;
          placed = true;
        }
        P.col.push_back(j);
        P.val.push_back(v);
      }
      if (!placed) {
        P.col.push_back(pc);
        P.val.push_back(pv);
      }
      P.ptr[i + 1] = (int)P.col.size();
    }
    HostCsr R = transpose(P);
    HostCsr Ac = multiply(R, multiply(A, P));
    L->P.upload(P, st);
    L->R.upload(R, st);
    L->lanes_P = lanes_for(P);
    L->lanes_R = lanes_for(R);
    FB_CUDA(cudaStreamSynchronize(st));
    A = std::move(Ac);
    B.swap(nrm);
  }
  FB_CUDA(cudaStreamSynchronize(st));
  amg->operator_complexity = nnz_total / std::max(1.0, nnz0);
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

void amg_apply(fb_amg *amg, const double *r, double *z) {
  if (!amg->gather) return amg_cycle(amg, r, z);
  fb_ctx *ctx = amg->ctx;
  const double *rg = fb_peer_vec_gather(ctx, amg->gather, r, amg->l2g.p, amg->n_owned);
  amg_cycle(amg, rg, amg->zg.p);
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
