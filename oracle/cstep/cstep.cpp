// CPU implementation of one IPCS / backward-Euler step on tetrahedra in C++/OpenMP
// (TEST / BENCH INFRASTRUCTURE: bench.py's cpu_baseline and --impl reference legs; never loaded by flow_b200).
//
// The reference's own stack (DOLFIN + PETSc + hypre) cannot be installed here; this file restates the
// arithmetic of flow/navier_stokes/pressure_correction.py on the host cores so that the GPU numbers have a compiled,
// multi-threaded CPU number beside them:
//   * _compute_tentative_velocity (:147-255): Newton on F1 (backward Euler, :175-180) with the Jacobian of the current
//     iterate (:202), |F|_2 < 1e-10 absolute, <= 10 iterations (:228-236).  The reference's LU update is replaced by
//     Jacobi-BiCGStab on the 3x3-block CSR Jacobian, solved to max(1e-13, 1e-6 |rhs|);
//   * _compute_pressure (:258-433), pure-Neumann branch: CG on the P1 stiffness matrix, PETSc's preconditioned-norm
//     test with rtol = tol (:419-424); Jacobi instead of hypre BoomerAMG;
//   * _compute_velocity_correction (:436-465): CG on the P2 mass matrix (one scalar matrix for the 3 components),
//     Dirichlet dofs eliminated symmetrically.
// Self-contained: the element integrals are written out here from the forms (_rhs_weak :135-144, derivative(F1, ui)),
// the quadrature rule and the P2 basis tables are handed in by the numpy oracle (oracle/fem.py) -- nothing is shared
// with the product's sources.  Sparsity patterns and the constant matrices are built here as well (the numpy/scipy
// route needs > 40 GB at 10 M dofs).  Full Dirichlet velocity conditions, p_bcs = [] (the benchmark configuration;
// the exterior-facet terms of _rhs_weak then only touch constrained rows and are skipped).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
constexpr int D = 3, NL = 10, NV = 4, MAXQ = 32;

struct Csr {  // node-level pattern
  int64_t n = 0;
  std::vector<int64_t> ptr;
  std::vector<int32_t> col;
  int64_t find(int64_t row, int32_t c) const {
    int64_t lo = ptr[row], hi = ptr[row + 1] - 1;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (col[mid] < c) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
};

struct Problem {
  int64_t nc = 0, nn = 0, nv = 0;  // cells, P2 nodes, vertices (= P1 nodes; the first nv P2 nodes)
  std::vector<int32_t> cn;         // nc * 10
  std::vector<double> xyz;         // nv * 3
  int nq = 0;
  double lam[MAXQ][NV], w[MAXQ], phi[MAXQ][NL], dphi[MAXQ][NL][NV], mean[NL], mref[NL][NL];
  Csr P2, P1;
  std::vector<double> M, Ap, J;    // scalar P2 mass, P1 stiffness, Jacobian blocks (9 per P2 pattern entry)
  std::vector<int64_t> bc_dofs;
  std::vector<double> bc_vals;
  std::vector<uint8_t> mask;       // per velocity dof
};

void geometry(const Problem &p, const int32_t *cn, double glam[NV][D], double &vol) {
  const double *x0 = &p.xyz[(int64_t)cn[0] * 3];
  double Jm[3][3];
  for (int v = 1; v < 4; ++v)
    for (int k = 0; k < 3; ++k) Jm[k][v - 1] = p.xyz[(int64_t)cn[v] * 3 + k] - x0[k];
  const double det = Jm[0][0] * (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) - Jm[0][1] * (Jm[1][0] * Jm[2][2] - Jm[1][2] * Jm[2][0]) +
                     Jm[0][2] * (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]);
  vol = std::fabs(det) / 6.0;
  // rows of the inverse of Jm are the gradients of lambda_1..3
  double inv[3][3];
  inv[0][0] = (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) / det;
  inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) / det;
  inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) / det;
  inv[1][0] = (Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2]) / det;
  inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) / det;
  inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) / det;
  inv[2][0] = (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]) / det;
  inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) / det;
  inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) / det;
  for (int k = 0; k < 3; ++k) {
    glam[1][k] = inv[0][k];
    glam[2][k] = inv[1][k];
    glam[3][k] = inv[2][k];
    glam[0][k] = -(inv[0][k] + inv[1][k] + inv[2][k]);
  }
}

// node pattern of the first `nl` local nodes of every cell (nl = 10: P2, nl = 4: P1), columns ascending
void build_pattern(const Problem &p, int nl, int64_t nrows, Csr &A) {
  std::vector<int64_t> cnt(nrows + 1, 0);
  for (int64_t c = 0; c < p.nc; ++c)
    for (int a = 0; a < nl; ++a) cnt[p.cn[c * NL + a] + 1]++;
  for (int64_t i = 0; i < nrows; ++i) cnt[i + 1] += cnt[i];
  std::vector<int32_t> cells(cnt[nrows]);  // node -> cells
  std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
  for (int64_t c = 0; c < p.nc; ++c)
    for (int a = 0; a < nl; ++a) cells[fill[p.cn[c * NL + a]]++] = (int32_t)c;
  A.n = nrows;
  A.ptr.assign(nrows + 1, 0);
  std::vector<int32_t> len(nrows);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < nrows; ++i) {
      tmp.clear();
      for (int64_t k = cnt[i]; k < cnt[i + 1]; ++k)
        for (int b = 0; b < nl; ++b) tmp.push_back(p.cn[(int64_t)cells[k] * NL + b]);
      std::sort(tmp.begin(), tmp.end());
      len[i] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  }
  for (int64_t i = 0; i < nrows; ++i) A.ptr[i + 1] = A.ptr[i] + len[i];
  A.col.resize(A.ptr[nrows]);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < nrows; ++i) {
      tmp.clear();
      for (int64_t k = cnt[i]; k < cnt[i + 1]; ++k)
        for (int b = 0; b < nl; ++b) tmp.push_back(p.cn[(int64_t)cells[k] * NL + b]);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      std::copy(tmp.begin(), tmp.end(), A.col.begin() + A.ptr[i]);
    }
  }
}

void assemble_constant(Problem &p) {
  p.M.assign(p.P2.col.size(), 0.0);
  p.Ap.assign(p.P1.col.size(), 0.0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < p.nc; ++c) {
    const int32_t *cn = &p.cn[c * NL];
    double glam[NV][D], vol;
    geometry(p, cn, glam, vol);
    for (int a = 0; a < NL; ++a)
      for (int b = 0; b < NL; ++b) {
        const int64_t k = p.P2.find(cn[a], cn[b]);
#pragma omp atomic
        p.M[k] += vol * p.mref[a][b];
      }
    for (int a = 0; a < NV; ++a)
      for (int b = 0; b < NV; ++b) {
        const int64_t k = p.P1.find(cn[a], cn[b]);
        const double e = vol * (glam[a][0] * glam[b][0] + glam[a][1] * glam[b][1] + glam[a][2] * glam[b][2]);
#pragma omp atomic
        p.Ap[k] += e;
      }
  }
}

// F1(ui; v) = (ui - u0, v) - dt/rho R(ui; v), R = _rhs_weak cell terms (:135-144); J = dF1/dui (blocks on the P2 pattern)
void assemble_momentum(Problem &p, double dt, double rho, double mu, const double *ui, const double *u0, const double *p0,
                       double *F, bool withJ) {
  const double c1 = 0.5 * dt, c2 = dt * mu / rho, cdt = dt / rho;
#pragma omp parallel
  {
    std::vector<double> Je(withJ ? NL * NL * 9 : 0);
#pragma omp for schedule(dynamic, 128)
    for (int64_t c = 0; c < p.nc; ++c) {
      const int32_t *cn = &p.cn[c * NL];
      double glam[NV][D], vol;
      geometry(p, cn, glam, vol);
      double U[NL][D], U0[NL][D], Fe[NL][D] = {{0}};
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i) {
          U[a][i] = ui[(int64_t)cn[a] * D + i];
          U0[a][i] = u0[(int64_t)cn[a] * D + i];
        }
      if (withJ) std::fill(Je.begin(), Je.end(), 0.0);
      for (int q = 0; q < p.nq; ++q) {
        const double w = p.w[q] * vol;
        const double *phi = p.phi[q];
        double g[NL][D];
        for (int a = 0; a < NL; ++a)
          for (int k = 0; k < D; ++k) {
            double s = 0.0;
            for (int m = 0; m < NV; ++m) s += p.dphi[q][a][m] * glam[m][k];
            g[a][k] = s;
          }
        double uq[D] = {0}, du[D] = {0}, gu[D][D] = {{0}}, p0q = 0.0;
        for (int v = 0; v < NV; ++v) p0q += p0[cn[v]] * p.lam[q][v];
        for (int a = 0; a < NL; ++a)
          for (int i = 0; i < D; ++i) {
            uq[i] += U[a][i] * phi[a];
            du[i] += (U[a][i] - U0[a][i]) * phi[a];
            for (int k = 0; k < D; ++k) gu[i][k] += U[a][i] * g[a][k];
          }
        double conv[D];  // (grad u) u
        for (int i = 0; i < D; ++i) conv[i] = gu[i][0] * uq[0] + gu[i][1] * uq[1] + gu[i][2] * uq[2];
        double ug[NL];   // u . grad phi_a
        for (int a = 0; a < NL; ++a) ug[a] = uq[0] * g[a][0] + uq[1] * g[a][1] + uq[2] * g[a][2];
        for (int a = 0; a < NL; ++a)
          for (int i = 0; i < D; ++i) {
            double eps_g = 0.0;  // eps(u)_ik d_k phi_a
            for (int k = 0; k < D; ++k) eps_g += 0.5 * (gu[i][k] + gu[k][i]) * g[a][k];
            const double R = -0.5 * rho * (conv[i] * phi[a] - ug[a] * uq[i]) - 2.0 * mu * eps_g + p0q * g[a][i];
            Fe[a][i] += w * (phi[a] * du[i] - cdt * R);
          }
        if (withJ)
          for (int a = 0; a < NL; ++a)
            for (int b = 0; b < NL; ++b) {
              double *B = &Je[(a * NL + b) * 9];
              const double gab = g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2];
              const double diag = phi[a] * phi[b] + c1 * (ug[b] * phi[a] - ug[a] * phi[b]) + c2 * gab;
              for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j)
                  B[i * 3 + j] += w * ((i == j ? diag : 0.0) + c1 * (phi[a] * phi[b] * gu[i][j] - g[a][j] * phi[b] * uq[i]) +
                                       c2 * g[b][i] * g[a][j]);
            }
      }
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i) {
#pragma omp atomic
          F[(int64_t)cn[a] * D + i] += Fe[a][i];
        }
      if (withJ)
        for (int a = 0; a < NL; ++a)
          for (int b = 0; b < NL; ++b) {
            double *dst = &p.J[p.P2.find(cn[a], cn[b]) * 9];
            const double *B = &Je[(a * NL + b) * 9];
            for (int e = 0; e < 9; ++e) {
#pragma omp atomic
              dst[e] += B[e];
            }
          }
    }
  }
}

// y = J x (3x3 blocks on the P2 node pattern); rows of constrained dofs act as identity
void bsr_spmv(const Problem &p, const double *x, double *y, bool masked) {
#pragma omp parallel for schedule(static)
  for (int64_t I = 0; I < p.nn; ++I) {
    double s[3] = {0, 0, 0};
    for (int64_t k = p.P2.ptr[I]; k < p.P2.ptr[I + 1]; ++k) {
      const double *B = &p.J[k * 9];
      const double *xj = &x[(int64_t)p.P2.col[k] * 3];
      for (int i = 0; i < 3; ++i) s[i] += B[i * 3] * xj[0] + B[i * 3 + 1] * xj[1] + B[i * 3 + 2] * xj[2];
    }
    for (int i = 0; i < 3; ++i) y[I * 3 + i] = (masked && p.mask[I * 3 + i]) ? x[I * 3 + i] : s[i];
  }
}

// y = (A (x) I_nc) x with a scalar node matrix; constrained dofs: y = x if mask
void scalar_spmm(const Csr &A, const double *val, int nc, const double *x, double *y, const uint8_t *mask) {
#pragma omp parallel for schedule(static)
  for (int64_t I = 0; I < A.n; ++I) {
    double s[3] = {0, 0, 0};
    for (int64_t k = A.ptr[I]; k < A.ptr[I + 1]; ++k) {
      const double a = val[k];
      const double *xj = &x[(int64_t)A.col[k] * nc];
      for (int i = 0; i < nc; ++i) s[i] += a * xj[i];
    }
    for (int i = 0; i < nc; ++i) y[I * nc + i] = (mask && mask[I * nc + i]) ? x[I * nc + i] : s[i];
  }
}

double dot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// right-preconditioned Jacobi-BiCGStab on the block Jacobian with identity rows on the mask; x0 = 0; ||r|| <= atol
int bicgstab(const Problem &p, const double *dinv, const double *b, double *x, double atol, int maxit) {
  const int64_t n = p.nn * 3;
  std::vector<double> r(b, b + n), rh(b, b + n), pv(n, 0.0), v(n, 0.0), ph(n), sh(n), t(n);
  std::fill(x, x + n, 0.0);
  double rho = 1, alpha = 1, omega = 1;
  if (std::sqrt(dot(n, r.data(), r.data())) <= atol) return 0;
  for (int it = 1; it <= maxit; ++it) {
    const double rho_new = dot(n, rh.data(), r.data());
    const double beta = it == 1 ? 0.0 : (rho_new / rho) * (alpha / omega);
    rho = rho_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      pv[i] = r[i] + beta * (pv[i] - omega * v[i]);
      ph[i] = dinv[i] * pv[i];
    }
    bsr_spmv(p, ph.data(), v.data(), true);
    alpha = rho / dot(n, rh.data(), v.data());
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      r[i] -= alpha * v[i];
      sh[i] = dinv[i] * r[i];
    }
    bsr_spmv(p, sh.data(), t.data(), true);
    const double tt = dot(n, t.data(), t.data());
    omega = tt > 0 ? dot(n, t.data(), r.data()) / tt : 0.0;
    double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * ph[i] + omega * sh[i];
      r[i] -= omega * t[i];
      rr += r[i] * r[i];
    }
    if (std::sqrt(rr) <= atol) return it;
    if (omega == 0.0 || rho == 0.0) return -1;
  }
  return -1;
}

// Jacobi-PCG, x0 = 0, ||D^-1 r|| <= rtol ||D^-1 b|| (PETSc default: preconditioned norm)
int pcg(const Csr &A, const double *val, int nc, const uint8_t *mask, const double *b, double *x, double rtol, int maxit) {
  const int64_t n = A.n * nc;
  std::vector<double> dinv(n), r(b, b + n), z(n), pv(n), Ap(n);
#pragma omp parallel for schedule(static)
  for (int64_t I = 0; I < A.n; ++I) {
    const double dgl = val[A.find(I, (int32_t)I)];
    for (int i = 0; i < nc; ++i) dinv[I * nc + i] = (mask && mask[I * nc + i]) ? 1.0 : 1.0 / dgl;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    x[i] = 0.0;
    z[i] = dinv[i] * r[i];
    pv[i] = z[i];
  }
  double rz = dot(n, r.data(), z.data());
  const double ref = std::sqrt(dot(n, z.data(), z.data()));
  if (ref == 0.0) return 0;
  for (int it = 1; it <= maxit; ++it) {
    scalar_spmm(A, val, nc, pv.data(), Ap.data(), mask);
    const double alpha = rz / dot(n, pv.data(), Ap.data());
    double rz_new = 0.0, zz = 0.0;
#pragma omp parallel for reduction(+ : rz_new, zz) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * pv[i];
      r[i] -= alpha * Ap[i];
      z[i] = dinv[i] * r[i];
      rz_new += r[i] * z[i];
      zz += z[i] * z[i];
    }
    if (std::sqrt(zz) <= rtol * ref) return it;
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) pv[i] = z[i] + beta * pv[i];
  }
  return -1;
}
}  // namespace

extern "C" {

int cs_num_threads() { return omp_get_max_threads(); }
void cs_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

// tables: lam[nq][4], w[nq] (sum = 1), phi[nq][10], dphi[nq][10][4] from oracle/fem.py (degree-5 rule)
void *cs_create(int64_t nc, const int32_t *cell_nodes, int64_t nn, int64_t nv, const double *xyz, int nq, const double *lam,
                const double *w, const double *phi, const double *dphi, int64_t nbc, const int64_t *bc_dofs,
                const double *bc_vals) {
  if (nq > MAXQ) return nullptr;
  Problem *p = new Problem();
  p->nc = nc;
  p->nn = nn;
  p->nv = nv;
  p->cn.assign(cell_nodes, cell_nodes + nc * NL);
  p->xyz.assign(xyz, xyz + nv * 3);
  p->nq = nq;
  std::memcpy(p->lam, lam, sizeof(double) * nq * NV);
  std::memcpy(p->w, w, sizeof(double) * nq);
  std::memcpy(p->phi, phi, sizeof(double) * nq * NL);
  std::memcpy(p->dphi, dphi, sizeof(double) * nq * NL * NV);
  for (int a = 0; a < NL; ++a) {
    p->mean[a] = 0.0;
    for (int q = 0; q < nq; ++q) p->mean[a] += p->w[q] * p->phi[q][a];
    for (int b = 0; b < NL; ++b) {
      p->mref[a][b] = 0.0;
      for (int q = 0; q < nq; ++q) p->mref[a][b] += p->w[q] * p->phi[q][a] * p->phi[q][b];
    }
  }
  build_pattern(*p, NL, nn, p->P2);
  build_pattern(*p, NV, nv, p->P1);
  assemble_constant(*p);
  p->J.assign(p->P2.col.size() * 9, 0.0);
  p->bc_dofs.assign(bc_dofs, bc_dofs + nbc);
  p->bc_vals.assign(bc_vals, bc_vals + nbc);
  p->mask.assign(nn * 3, 0);
  for (int64_t k = 0; k < nbc; ++k) p->mask[bc_dofs[k]] = 1;
  return p;
}

void cs_destroy(void *h) { delete static_cast<Problem *>(h); }

int64_t cs_nnz(void *h, int which) {
  Problem *p = static_cast<Problem *>(h);
  return which == 0 ? (int64_t)p->P1.col.size() : (int64_t)p->P2.col.size();
}

// stats: newton its, momentum Krylov its, pressure its, correction its
int cs_step(void *h, double dt, double rho, double mu, const double *u0, const double *p0, double tol, double *u1, double *p1,
            int *stats) {
  Problem &p = *static_cast<Problem *>(h);
  const int64_t nu = p.nn * 3, np_ = p.nv, nbc = (int64_t)p.bc_dofs.size();
  std::vector<double> ui(u0, u0 + nu), F(nu), delta(nu), dinv(nu), dg(nu), tmp(nu);
  auto residual = [&](bool withJ) {
    std::fill(F.begin(), F.end(), 0.0);
    if (withJ) std::fill(p.J.begin(), p.J.end(), 0.0);
    assemble_momentum(p, dt, rho, mu, ui.data(), u0, p0, F.data(), withJ);
    for (int64_t k = 0; k < nbc; ++k) F[p.bc_dofs[k]] = ui[p.bc_dofs[k]] - p.bc_vals[k];
    return std::sqrt(dot(nu, F.data(), F.data()));
  };
  double r = residual(true);
  int newton = 0, kits = 0;
  while (r >= 1e-10) {  // :228-236
    if (newton >= 10) return 2;
    if (newton > 0) residual(true);
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < p.nn; ++I) {
      const double *B = &p.J[p.P2.find(I, (int32_t)I) * 9];
      for (int i = 0; i < 3; ++i) dinv[I * 3 + i] = p.mask[I * 3 + i] ? 1.0 : 1.0 / B[i * 4];
    }
    // lift the Dirichlet dofs (their update is ui - g exactly), solve on the free ones
    std::fill(dg.begin(), dg.end(), 0.0);
    for (int64_t k = 0; k < nbc; ++k) dg[p.bc_dofs[k]] = F[p.bc_dofs[k]];
    bsr_spmv(p, dg.data(), tmp.data(), false);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nu; ++i) F[i] = p.mask[i] ? 0.0 : F[i] - tmp[i];
    const double bnorm = std::sqrt(dot(nu, F.data(), F.data()));
    const int its = bicgstab(p, dinv.data(), F.data(), delta.data(), std::max(1e-13, 1e-6 * bnorm), 2000);
    if (its < 0) return 3;
    kits += its;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nu; ++i) ui[i] -= delta[i] + dg[i];
    ++newton;
    r = residual(false);
  }
  // pressure (:317-323, non-rotational): b = -rho/dt (div ui, q) + (grad p0, grad q)
  std::vector<double> bp(np_, 0.0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < p.nc; ++c) {
    const int32_t *cn = &p.cn[c * NL];
    double glam[NV][D], vol, be[NV] = {0, 0, 0, 0};
    geometry(p, cn, glam, vol);
    for (int q = 0; q < p.nq; ++q) {
      double div = 0.0;
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i) {
          double gai = 0.0;
          for (int m = 0; m < NV; ++m) gai += p.dphi[q][a][m] * glam[m][i];
          div += ui[(int64_t)cn[a] * D + i] * gai;
        }
      for (int v = 0; v < NV; ++v) be[v] += -rho / dt * p.w[q] * vol * div * p.lam[q][v];
    }
    double gp[D] = {0, 0, 0};
    for (int v = 0; v < NV; ++v)
      for (int k = 0; k < D; ++k) gp[k] += p0[cn[v]] * glam[v][k];
    for (int v = 0; v < NV; ++v) {
      be[v] += vol * (gp[0] * glam[v][0] + gp[1] * glam[v][1] + gp[2] * glam[v][2]);
#pragma omp atomic
      bp[cn[v]] += be[v];
    }
  }
  const int pits = pcg(p.P1, p.Ap.data(), 1, nullptr, bp.data(), p1, tol, 100000);
  if (pits < 0) return 3;
  // correction (:441-449): M u1 = M ui - dt/rho (grad(p1 - p0), v); symmetric elimination of the Dirichlet dofs
  std::vector<double> bu(nu), xg(nu, 0.0), wv(nu);
  scalar_spmm(p.P2, p.M.data(), 3, ui.data(), bu.data(), nullptr);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < p.nc; ++c) {
    const int32_t *cn = &p.cn[c * NL];
    double glam[NV][D], vol, gphi[D] = {0, 0, 0};
    geometry(p, cn, glam, vol);
    for (int v = 0; v < NV; ++v)
      for (int k = 0; k < D; ++k) gphi[k] += (p1[cn[v]] - p0[cn[v]]) * glam[v][k];
    for (int a = 0; a < NL; ++a)
      for (int k = 0; k < D; ++k) {
#pragma omp atomic
        bu[(int64_t)cn[a] * D + k] += -dt / rho * vol * p.mean[a] * gphi[k];
      }
  }
  for (int64_t k = 0; k < nbc; ++k) xg[p.bc_dofs[k]] = p.bc_vals[k];
  scalar_spmm(p.P2, p.M.data(), 3, xg.data(), wv.data(), nullptr);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nu; ++i) bu[i] = p.mask[i] ? 0.0 : bu[i] - wv[i];
  // masked operator == symmetric elimination when all Krylov vectors vanish on the constrained dofs (they do: b = 0 there)
  const int cits = pcg(p.P2, p.M.data(), 3, p.mask.data(), bu.data(), u1, tol, 5000);
  if (cits < 0) return 3;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nu; ++i) u1[i] += xg[i];
  stats[0] = newton;
  stats[1] = kits;
  stats[2] = pits;
  stats[3] = cits;
  return 0;
}

// OpenMP SpMV throughput on the two matrices that carry the step: the momentum Jacobian (3x3 blocks) and the scalar P2
// mass matrix x 3 components.  gbs[0], gbs[1]: algorithmic GB/s with the scalar-CSR byte count of SURVEY.md 8d
// (12 B per scalar non-zero + 20 B per scalar row); ms[0], ms[1]: time per product.
void cs_spmv_bench(void *h, int reps, double *gbs, double *ms) {
  Problem &p = *static_cast<Problem *>(h);
  const int64_t nu = p.nn * 3;
  std::vector<double> x(nu), y(nu);
  for (int64_t i = 0; i < nu; ++i) x[i] = std::sin((double)i);
  const double nnzb = (double)p.P2.col.size();
  for (int which = 0; which < 2; ++which) {
    for (int k = 0; k < 2; ++k) which == 0 ? bsr_spmv(p, x.data(), y.data(), false) : scalar_spmm(p.P2, p.M.data(), 3, x.data(), y.data(), nullptr);
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < reps; ++k) which == 0 ? bsr_spmv(p, x.data(), y.data(), false) : scalar_spmm(p.P2, p.M.data(), 3, x.data(), y.data(), nullptr);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
    const double bytes = which == 0 ? nnzb * 9.0 * 12.0 + nu * 20.0 : 3.0 * (nnzb * 12.0 + p.nn * 20.0);
    gbs[which] = bytes / sec / 1e9;
    ms[which] = sec * 1e3;
  }
}
}
