// CPU baseline of the whole IPCS step in C++/OpenMP (BENCH INFRASTRUCTURE, not a product path, never
// loaded by flow_b200).  Same algorithm as the GPU path -- Newton with Jacobi-BiCGStab for the tentative
// velocity (pressure_correction.py:147-255), Jacobi-PCG for the pressure Poisson (:258-433) and for the
// velocity correction (:436-465) -- so that bench.py's cpu_baseline / --impl reference legs time a
// multi-threaded compiled implementation instead of numpy.  Element integrals come from
// flow_b200/csrc/fb_element.cuh compiled for the host (checked against the independent numpy oracle in
// tests/test_hostsim.py); CSR patterns and the constant matrices are handed in by the numpy oracle.
// Backward Euler, full Dirichlet velocity conditions, p_bcs = [] (the benchmark configuration).
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../flow_b200/csrc/fb_element.cuh"

namespace hq {
#define FB_TABLE static const
#include "../../flow_b200/csrc/fb_quadrature.h"
#undef FB_TABLE
}  // namespace hq

extern "C" {
void cb_spmv(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data, const double *x, double *y);
int cb_pcg(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data, const double *dinv,
           const double *b, double *x, double rtol, int maxit);
int cb_bicgstab(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data, const double *dinv,
                const double *b, double *x, double atol, int maxit);
}

namespace {
constexpr int D = 3, NL = 10, NQ = hq::TET_D5_NQ;

struct Csr {
  int64_t n;
  const int64_t *indptr;
  const int32_t *indices;
  double *data;
  int64_t find(int64_t row, int32_t col) const {
    int64_t lo = indptr[row], hi = indptr[row + 1] - 1;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (indices[mid] < col) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
};

void geom(const int *cv, const double *xyz, double glam[4][3], double &vol) {
  double X[12];
  for (int v = 0; v < 4; ++v)
    for (int k = 0; k < 3; ++k) X[v * 3 + k] = xyz[(int64_t)cv[v] * 3 + k];
  fb_geometry<3>(X, glam, vol);
}

// F = (ui - u0, v) - dt/rho R(ui; v)  (cell terms; the cavity has Dirichlet data on the whole boundary, so the
// facet terms only touch constrained rows) and, if J, the Jacobian (scalar CSR of the interleaved system)
void assemble(int64_t nc, const int *cell_nodes, const double *xyz, double dt, double rho, double mu, const double *ui,
              const double *u0, const double *p0, double *F, Csr *J) {
  const double c1 = 0.5 * dt, c2 = dt * mu / rho, cdt = dt / rho;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < nc; ++c) {
    const int *cn = cell_nodes + c * NL;
    double glam[4][3], vol;
    geom(cn, xyz, glam, vol);
    double Fe[NL][D] = {{0}};
    static thread_local std::vector<double> Je;
    if (J) Je.assign(NL * D * NL * D, 0.0);
    for (int q = 0; q < NQ; ++q) {
      const double *lam = &hq::TET_D5_LAM[q][0];
      const double w = hq::TET_D5_W[q] * vol;
      double phi[NL], g[NL][D];
      for (int a = 0; a < NL; ++a) {
        phi[a] = fb_p2_phi<3>(a, lam);
        fb_p2_grad<3>(a, lam, glam, g[a]);
      }
      double p0q = 0.0;
      for (int v = 0; v < 4; ++v) p0q += p0[cn[v]] * lam[v];
      double uq[D] = {0}, u0q[D] = {0}, gu[D][D] = {{0}};
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i) {
          const double ua = ui[(int64_t)cn[a] * D + i];
          uq[i] += ua * phi[a];
          u0q[i] += u0[(int64_t)cn[a] * D + i] * phi[a];
          for (int k = 0; k < D; ++k) gu[i][k] += ua * g[a][k];
        }
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i)
          Fe[a][i] += w * phi[a] * (uq[i] - u0q[i]) - cdt * w * fb_rhs_point<3>(i, rho, mu, phi[a], g[a], uq, gu, p0q);
      if (J)
        for (int a = 0; a < NL; ++a)
          for (int b = 0; b < NL; ++b) {
            double blk[D][D] = {{0}};
            fb_jac_point<3>(w, c1, c2, phi[a], phi[b], g[a], g[b], uq, gu, blk);
            for (int i = 0; i < D; ++i)
              for (int j = 0; j < D; ++j) Je[((a * D + i) * NL + b) * D + j] += blk[i][j];
          }
    }
    for (int a = 0; a < NL; ++a)
      for (int i = 0; i < D; ++i) {
#pragma omp atomic
        F[(int64_t)cn[a] * D + i] += Fe[a][i];
      }
    if (J)
      for (int a = 0; a < NL; ++a)
        for (int i = 0; i < D; ++i) {
          const int64_t row = (int64_t)cn[a] * D + i;
          for (int b = 0; b < NL; ++b) {
            const int64_t k0 = J->find(row, cn[b] * D);  // the D columns of node b are contiguous
            for (int j = 0; j < D; ++j) {
#pragma omp atomic
              J->data[k0 + j] += Je[((a * D + i) * NL + b) * D + j];
            }
          }
        }
  }
}

void rows_identity(Csr &A, const int64_t *dofs, int64_t n) {
#pragma omp parallel for
  for (int64_t k = 0; k < n; ++k) {
    const int64_t r = dofs[k];
    for (int64_t e = A.indptr[r]; e < A.indptr[r + 1]; ++e) A.data[e] = (A.indices[e] == r) ? 1.0 : 0.0;
  }
}

double norm2(int64_t n, const double *x) {
  double s = 0;
#pragma omp parallel for reduction(+ : s)
  for (int64_t i = 0; i < n; ++i) s += x[i] * x[i];
  return std::sqrt(s);
}
}  // namespace

extern "C" {

int cs_num_threads() { return omp_get_max_threads(); }

// One IPCS / backward-Euler step on tetrahedra.  J pattern = scalar CSR of the interleaved vector-P2 system,
// Ap = P1 stiffness, Mv = vector-P2 mass (interleaved CSR).  stats: newton its, momentum its, pressure its, correction its.
int cs_ipcs_step(int64_t nc, const int *cell_nodes, const double *xyz, int64_t nu, int64_t np_, const int64_t *Jptr,
                 const int32_t *Jidx, double *Jval, const int64_t *Aptr, const int32_t *Aidx, const double *Aval,
                 const int64_t *Mptr, const int32_t *Midx, const double *Mval, double dt, double rho, double mu,
                 const double *u0, const double *p0, int64_t nbc, const int64_t *bc_dofs, const double *bc_vals, double tol,
                 double *u1, double *p1, int *stats) {
  std::vector<double> ui(u0, u0 + nu), F(nu), delta(nu), dinv(nu), dg(nu), tmp(nu);
  Csr J{nu, Jptr, Jidx, Jval};
  auto residual = [&](bool withJ) {
    std::fill(F.begin(), F.end(), 0.0);
    if (withJ) std::memset(Jval, 0, sizeof(double) * Jptr[nu]);
    assemble(nc, cell_nodes, xyz, dt, rho, mu, ui.data(), u0, p0, F.data(), withJ ? &J : nullptr);
    for (int64_t k = 0; k < nbc; ++k) F[bc_dofs[k]] = ui[bc_dofs[k]] - bc_vals[k];
    return norm2(nu, F.data());
  };
  double r = residual(true);
  int newton = 0, kits = 0;
  while (r >= 1e-10) {
    if (newton >= 10) return 2;
    if (newton > 0) residual(true);
    rows_identity(J, bc_dofs, nbc);
#pragma omp parallel for
    for (int64_t i = 0; i < nu; ++i) dinv[i] = 1.0 / Jval[J.find(i, (int32_t)i)];
    // lift the Dirichlet dofs, solve on the free ones
    std::fill(dg.begin(), dg.end(), 0.0);
    for (int64_t k = 0; k < nbc; ++k) dg[bc_dofs[k]] = F[bc_dofs[k]];
    cb_spmv(nu, Jptr, Jidx, Jval, dg.data(), tmp.data());
#pragma omp parallel for
    for (int64_t i = 0; i < nu; ++i) F[i] -= tmp[i];
    for (int64_t k = 0; k < nbc; ++k) F[bc_dofs[k]] = 0.0;
    const int its = cb_bicgstab(nu, Jptr, Jidx, Jval, dinv.data(), F.data(), delta.data(), 1e-11, 1000);
    if (its < 0) return 3;
    kits += its;
#pragma omp parallel for
    for (int64_t i = 0; i < nu; ++i) ui[i] -= delta[i] + dg[i];
    ++newton;
    r = residual(false);
  }
  // pressure: -rho/dt (div ui, q) + (grad p0, grad q)
  std::vector<double> bp(np_, 0.0), dp(np_);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < nc; ++c) {
    const int *cn = cell_nodes + c * NL;
    double glam[4][3], vol, Ue[NL * D], p0e[4], be[4];
    geom(cn, xyz, glam, vol);
    for (int a = 0; a < NL; ++a)
      for (int i = 0; i < D; ++i) Ue[a * D + i] = ui[(int64_t)cn[a] * D + i];
    for (int v = 0; v < 4; ++v) p0e[v] = p0[cn[v]];
    fb_pressure_rhs_cell<3>(glam, vol, &hq::TET_D2_LAM[0][0], hq::TET_D2_W, hq::TET_D2_NQ, Ue, p0e, dt, rho, mu, 0, be);
    for (int v = 0; v < 4; ++v) {
#pragma omp atomic
      bp[cn[v]] += be[v];
    }
  }
  for (int64_t i = 0; i < np_; ++i) dp[i] = 1.0 / Aval[Csr{np_, Aptr, Aidx, const_cast<double *>(Aval)}.find(i, (int32_t)i)];
  const int pits = cb_pcg(np_, Aptr, Aidx, Aval, dp.data(), bp.data(), p1, tol, 50000);
  if (pits < 0) return 3;
  // correction: M u1 = M ui - dt/rho (grad(p1 - p0), v), Dirichlet dofs eliminated symmetrically (masked CG)
  std::vector<double> bu(nu), xg(nu, 0.0), w(nu), mdinv(nu);
  cb_spmv(nu, Mptr, Midx, Mval, ui.data(), bu.data());
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t c = 0; c < nc; ++c) {
    const int *cn = cell_nodes + c * NL;
    double glam[4][3], vol, dpe[4], gphi[3];
    geom(cn, xyz, glam, vol);
    for (int v = 0; v < 4; ++v) dpe[v] = p1[cn[v]] - p0[cn[v]];
    fb_correction_gradphi<3>(glam, nullptr, dpe, mu, 0, gphi);
    for (int a = 0; a < NL; ++a)
      for (int k = 0; k < D; ++k) {
#pragma omp atomic
        bu[(int64_t)cn[a] * D + k] += -dt / rho * vol * fb_p2_mean<3>(a) * gphi[k];
      }
  }
  // symmetric elimination on a copy of M
  std::vector<double> Mbc(Mval, Mval + Mptr[nu]);
  std::vector<uint8_t> mask(nu, 0);
  for (int64_t k = 0; k < nbc; ++k) {
    mask[bc_dofs[k]] = 1;
    xg[bc_dofs[k]] = bc_vals[k];
  }
  cb_spmv(nu, Mptr, Midx, Mval, xg.data(), w.data());
#pragma omp parallel for
  for (int64_t i = 0; i < nu; ++i) {
    bu[i] = mask[i] ? xg[i] : bu[i] - w[i];
    for (int64_t e = Mptr[i]; e < Mptr[i + 1]; ++e)
      if (mask[i]) Mbc[e] = (Midx[e] == i) ? 1.0 : 0.0; else if (mask[Midx[e]]) Mbc[e] = 0.0;
  }
  Csr Mc{nu, Mptr, Midx, Mbc.data()};
#pragma omp parallel for
  for (int64_t i = 0; i < nu; ++i) mdinv[i] = 1.0 / Mbc[Mc.find(i, (int32_t)i)];
  const int cits = cb_pcg(nu, Mptr, Midx, Mbc.data(), mdinv.data(), bu.data(), u1, tol, 5000);
  if (cits < 0) return 3;
  stats[0] = newton;
  stats[1] = kits;
  stats[2] = pits;
  stats[3] = cits;
  return 0;
}
}
