// Host-side set-up of the smoothed-aggregation AMG hierarchy (no CUDA in this file): strength filter, greedy
// aggregation, tentative prolongator carrying the near-null-space vector, prolongator smoothing, Galerkin
// products, dense (pseudo-)inverse of the coarsest operator.  fb_amg.cu uploads the levels and runs the V-cycle
// on the device; tests/hostsim compiles the same routines for the CPU-only test tier (it brings its own host
// V-cycle, which is test code and not part of the product).
// Replaces the set-up phase of hypre BoomerAMG as used by pressure_correction.py:331, :414-419 [EXT].
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace fb_amg_host {

struct HostCsr {
  int nrows = 0, ncols = 0;
  std::vector<int> ptr, col;
  std::vector<double> val;
};

inline HostCsr transpose(const HostCsr &A) {
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
inline HostCsr multiply(const HostCsr &A, const HostCsr &B) {
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
inline int aggregate(const HostCsr &A, double theta, std::vector<int> &agg) {
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

inline bool spd_inverse(std::vector<double> &a, int n) {
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


struct HostLevel {
  HostCsr A, P, R;            // operator, prolongator to this level from the next coarser one, its transpose
  std::vector<double> dinv;   // inverse diagonal
  std::vector<double> Ainv;   // coarsest level only: dense (pseudo-)inverse, row-major n x n (empty: Jacobi sweeps)
  double omega = 0.67;        // damping of the Jacobi smoother: 4 / (3 rho), rho = Gershgorin bound of D^-1 A
  int n = 0;
};

constexpr int AMG_COARSE_TARGET = 400;   // coarsen until at most this many rows ...
constexpr int AMG_DENSE_MAX = 2500;      // ... and invert densely if the coarsest level has at most this many

// Hierarchy of the n x n leading block of the CSR matrix (rowptr, col, val); columns >= n (ghosts) are dropped.
// singular: the operator annihilates constants (pure Neumann problem); complexity: sum of nnz over the levels / nnz.
inline void build_hierarchy(int n, const int *rowptr, const int *col, const double *val, std::vector<HostLevel> &levels,
                            bool &singular, double &complexity) {
  HostCsr A;
  A.nrows = A.ncols = n;
  A.ptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] < n) {
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
    singular = worst < 1e-10;
  }
  const double nnz0 = (double)A.val.size();
  double nnz_total = 0.0;
  double theta = 0.08;
  levels.clear();
  for (int lev = 0; lev < 12; ++lev) {
    levels.emplace_back();
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
    const double omega = 4.0 / (3.0 * rho);
    levels.back().n = A.nrows;
    levels.back().omega = omega;
    levels.back().dinv = dinv;
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
        if (singular) {
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
          if (singular)
            for (int i = 0; i < m; ++i)
              for (int j = 0; j < m; ++j) dense[(size_t)i * m + j] -= w[i] * w[j] / shift;
          levels.back().Ainv.swap(dense);
        }
      }
      levels.back().A = std::move(A);
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
        double v = -omega * dinv[i] * AP0.val[k];
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
    levels.back().A = std::move(A);
    levels.back().P = std::move(P);
    levels.back().R = std::move(R);
    A = std::move(Ac);
    B.swap(nrm);
  }
  complexity = nnz_total / std::max(1.0, nnz0);
}

}  // namespace fb_amg_host
