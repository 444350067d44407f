// TEST-ONLY host harness: compiles the FB_HD element routines of
// flow_b200/csrc/fb_element.cuh (the exact code the CUDA kernels call) with g++ so that
// the CPU-only test tier can check the element math against the oracle.  It is built
// into tests/hostsim/_hostsim.so by tests/conftest.py, is never imported by flow_b200
// and is not a compute path of the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../flow_b200/csrc/fb_amg_host.h"
#include "../../flow_b200/csrc/fb_element.cuh"

namespace hq {
#define FB_TABLE static const
#include "../../flow_b200/csrc/fb_quadrature.h"
#include "../../flow_b200/csrc/fb_p2_tables.h"
#undef FB_TABLE
}  // namespace hq
template <int D> static const double *m3_table() { return D == 2 ? hq::FB_M3_TRI : hq::FB_M3_TET; }
static int g_closed_form = 0;  // 1: cell part of J from fb_jac_pair (closed form) instead of fb_jac_point (quadrature)

template <int D> struct HQ5;
template <> struct HQ5<2> { static constexpr int NQ = hq::TRI_D5_NQ; static const double *lam() { return &hq::TRI_D5_LAM[0][0]; } static const double *w() { return hq::TRI_D5_W; } };
template <> struct HQ5<3> { static constexpr int NQ = hq::TET_D5_NQ; static const double *lam() { return &hq::TET_D5_LAM[0][0]; } static const double *w() { return hq::TET_D5_W; } };
template <int D> struct HQ2;
template <> struct HQ2<2> { static constexpr int NQ = hq::TRI_D2_NQ; static const double *lam() { return &hq::TRI_D2_LAM[0][0]; } static const double *w() { return hq::TRI_D2_W; } };
template <> struct HQ2<3> { static constexpr int NQ = hq::TET_D2_NQ; static const double *lam() { return &hq::TET_D2_LAM[0][0]; } static const double *w() { return hq::TET_D2_W; } };
template <int D> struct HQF;
template <> struct HQF<2> { static constexpr int NQ = hq::SEG_D5_NQ; static const double *lam() { return &hq::SEG_D5_LAM[0][0]; } static const double *w() { return hq::SEG_D5_W; } };
template <> struct HQF<3> { static constexpr int NQ = hq::TRI_D5_NQ; static const double *lam() { return &hq::TRI_D5_LAM[0][0]; } static const double *w() { return hq::TRI_D5_W; } };

template <int D>
static void geom(const int *cn, const double *xyz, double glam[D + 1][D], double &vol) {
  double X[(D + 1) * D];
  for (int v = 0; v <= D; ++v)
    for (int k = 0; k < D; ++k) X[v * D + k] = xyz[(int64_t)cn[v] * D + k];
  fb_geometry<D>(X, glam, vol);
}

template <int D>
static void momentum(int64_t nc, const int *cell_nodes, const double *xyz, int64_t nbf, const int *bf_cell,
                     const int *bf_local, double dt, double rho, double mu, double theta, const double *ui,
                     const double *u0, const double *p0, int64_t ndofs, double *F, double *J) {
  constexpr int NL = Elem<D>::NL2, NQ = HQ5<D>::NQ;
  const double c1 = 0.5 * theta * dt, c2 = theta * dt * mu / rho, cdt = dt / rho;
  std::memset(F, 0, sizeof(double) * ndofs);
  if (J) std::memset(J, 0, sizeof(double) * ndofs * ndofs);
  for (int64_t c = 0; c < nc; ++c) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    geom<D>(cn, xyz, glam, vol);
    for (int q = 0; q < NQ; ++q) {
      const double *lam = HQ5<D>::lam() + q * (D + 1);
      const double w = HQ5<D>::w()[q] * vol;
      double phi[NL], g[NL][D];
      for (int a = 0; a < NL; ++a) {
        phi[a] = fb_p2_phi<D>(a, lam);
        fb_p2_grad<D>(a, lam, glam, g[a]);
      }
      double p0q = 0.0;
      for (int v = 0; v <= D; ++v) p0q += p0[cn[v]] * lam[v];
      for (int state = 0; state < 2; ++state) {
        const double *u = state == 0 ? ui : u0;
        double uq[D], gu[D][D];
        for (int i = 0; i < D; ++i) {
          uq[i] = 0;
          for (int k = 0; k < D; ++k) gu[i][k] = 0;
          for (int a = 0; a < NL; ++a) {
            const double ua = u[(int64_t)cn[a] * D + i];
            uq[i] += ua * phi[a];
            for (int k = 0; k < D; ++k) gu[i][k] += ua * g[a][k];
          }
        }
        const double wt = state == 0 ? theta : 1.0 - theta;
        for (int a = 0; a < NL; ++a)
          for (int i = 0; i < D; ++i) {
            double acc = (state == 0 ? 1.0 : -1.0) * w * phi[a] * uq[i];
            if (wt != 0.0) acc -= cdt * wt * w * fb_rhs_point<D>(i, rho, mu, phi[a], g[a], uq, gu, p0q);
            F[(int64_t)cn[a] * D + i] += acc;
          }
        if (J && state == 0 && !g_closed_form) {
          for (int a = 0; a < NL; ++a)
            for (int b = 0; b < NL; ++b) {
              double blk[D][D];
              for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) blk[i][j] = 0;
              fb_jac_point<D>(w, c1, c2, phi[a], phi[b], g[a], g[b], uq, gu, blk);
              for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) J[((int64_t)cn[a] * D + i) * ndofs + (int64_t)cn[b] * D + j] += blk[i][j];
            }
        }
      }
    }
  }
  if (J && g_closed_form) {
    // the staging the CUDA kernel k_momentum_J_cf does per cell, then one fb_jac_pair per block
    constexpr int NV = D + 1;
    const double *M3 = m3_table<D>();
    for (int64_t c = 0; c < nc; ++c) {
      const int *cn = cell_nodes + c * NL;
      double glam[D + 1][D], vol;
      geom<D>(cn, xyz, glam, vol);
      double U[NL][D], GV[NL][NV * D], S[NL][D], Wt[NL][NV * D], GU[NV * D * D];
      for (int a = 0; a < NL; ++a) {
        for (int i = 0; i < D; ++i) {
          U[a][i] = ui[(int64_t)cn[a] * D + i];
          S[a][i] = 0.0;
        }
        for (int w = 0; w < NV; ++w) {
          fb_p2_vertex_grad<D>(a, w, glam, &GV[a][w * D]);
          for (int k = 0; k < D; ++k) S[a][k] += GV[a][w * D + k];
        }
      }
      for (int v = 0; v < NV; ++v)
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            double s2 = 0.0;
            for (int cc = 0; cc < NL; ++cc) s2 += U[cc][i] * GV[cc][v * D + j];
            GU[(v * D + i) * D + j] = s2;
          }
      for (int a = 0; a < NL; ++a)
        for (int w = 0; w < NV; ++w)
          for (int k = 0; k < D; ++k) {
            double s2 = 0.0;
            for (int cc = 0; cc < NL; ++cc) s2 += M3[(a * NL + cc) * NV + w] * U[cc][k];
            Wt[a][w * D + k] = s2;
          }
      for (int a = 0; a < NL; ++a)
        for (int b = 0; b < NL; ++b) {
          double blk[D][D];
          fb_jac_pair<D>(vol, c1, c2, GV[a], S[a], GV[b], S[b], Wt[a], Wt[b], GU, M3 + (a * NL + b) * NV, blk);
          for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) J[((int64_t)cn[a] * D + i) * ndofs + (int64_t)cn[b] * D + j] += blk[i][j];
        }
    }
  }
  for (int64_t fi = 0; fi < nbf; ++fi) {
    const int64_t c = bf_cell[fi];
    const int f = bf_local[fi];
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    geom<D>(cn, xyz, glam, vol);
    double Ue[NL * D], p0e[D + 1];
    for (int b = 0; b < NL; ++b)
      for (int k = 0; k < D; ++k)
        Ue[b * D + k] = theta * ui[(int64_t)cn[b] * D + k] + (1.0 - theta) * u0[(int64_t)cn[b] * D + k];
    for (int v = 0; v <= D; ++v) p0e[v] = p0[cn[v]];
    for (int a = 0; a < NL; ++a) {
      if (!fb_node_on_facet<D>(a, f)) continue;
      for (int i = 0; i < D; ++i)
        F[(int64_t)cn[a] * D + i] -= cdt * fb_facet_F<D>(a, i, f, glam, vol, HQF<D>::lam(), HQF<D>::w(), HQF<D>::NQ, Ue, p0e, mu);
      if (J && theta != 0.0)
        for (int b = 0; b < NL; ++b) {
          double B[D][D];
          fb_facet_J<D>(a, b, f, glam, vol, HQF<D>::lam(), HQF<D>::w(), HQF<D>::NQ, B);
          for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j)
              J[((int64_t)cn[a] * D + i) * ndofs + (int64_t)cn[b] * D + j] += -theta * cdt * mu * B[i][j];
        }
    }
  }
}

template <int D>
static void rhs(int64_t nc, const int *cell_nodes, const double *xyz, double dt, double rho, double mu, int rotational,
                const double *ui, const double *p1, const double *p0, double *bp, double *bu_grad) {
  constexpr int NL = Elem<D>::NL2;
  for (int64_t c = 0; c < nc; ++c) {
    const int *cn = cell_nodes + c * NL;
    double glam[D + 1][D], vol;
    geom<D>(cn, xyz, glam, vol);
    double Ue[NL * D], p0e[D + 1], dpe[D + 1], be[D + 1], gphi[D];
    for (int a = 0; a < NL; ++a)
      for (int i = 0; i < D; ++i) Ue[a * D + i] = ui[(int64_t)cn[a] * D + i];
    for (int v = 0; v <= D; ++v) {
      p0e[v] = p0[cn[v]];
      dpe[v] = p1[cn[v]] - p0[cn[v]];
    }
    fb_pressure_rhs_cell<D>(glam, vol, HQ2<D>::lam(), HQ2<D>::w(), HQ2<D>::NQ, Ue, p0e, dt, rho, mu, rotational, be);
    for (int v = 0; v <= D; ++v) bp[cn[v]] += be[v];
    fb_correction_gradphi<D>(glam, Ue, dpe, mu, rotational, gphi);
    for (int a = 0; a < NL; ++a)
      for (int k = 0; k < D; ++k) bu_grad[(int64_t)cn[a] * D + k] += -dt / rho * vol * fb_p2_mean<D>(a) * gphi[k];
  }
}

// ---- host V-cycle + PCG on the hierarchy built by fb_amg_host::build_hierarchy (the set-up code the library
// uploads to the GPU); mirrors amg_cycle / krylov_pcg_single_reduction of the CUDA path operation by operation
namespace {
using fb_amg_host::HostCsr;
using fb_amg_host::HostLevel;

void csr_mv(const HostCsr &M, const std::vector<double> &x, std::vector<double> &y) {
  y.assign(M.nrows, 0.0);
  for (int i = 0; i < M.nrows; ++i) {
    double s = 0.0;
    for (int k = M.ptr[i]; k < M.ptr[i + 1]; ++k) s += M.val[k] * x[M.col[k]];
    y[i] = s;
  }
}

void vcycle(const std::vector<HostLevel> &lv, size_t l, const std::vector<double> &b, std::vector<double> &x) {
  const HostLevel &L = lv[l];
  const int n = L.n;
  x.assign(n, 0.0);
  if (l + 1 == lv.size()) {
    if (!L.Ainv.empty()) {
      for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += L.Ainv[(size_t)i * n + k] * b[k];
        x[i] = s;
      }
    } else {  // a single level (or a stalled coarsening): damped Jacobi sweeps
      std::vector<double> ax;
      for (int i = 0; i < n; ++i) x[i] = L.omega * L.dinv[i] * b[i];
      for (int sweep = 0; sweep < (lv.size() == 1 ? 1 : 4); ++sweep) {
        csr_mv(L.A, x, ax);
        for (int i = 0; i < n; ++i) x[i] += L.omega * L.dinv[i] * (b[i] - ax[i]);
      }
    }
    return;
  }
  std::vector<double> ax, r(n), bc, xc, px;
  for (int i = 0; i < n; ++i) x[i] = L.omega * L.dinv[i] * b[i];  // pre-smooth from zero
  csr_mv(L.A, x, ax);
  for (int i = 0; i < n; ++i) r[i] = b[i] - ax[i];
  csr_mv(L.R, r, bc);
  vcycle(lv, l + 1, bc, xc);
  csr_mv(L.P, xc, px);
  for (int i = 0; i < n; ++i) x[i] += px[i];
  csr_mv(L.A, x, ax);
  for (int i = 0; i < n; ++i) x[i] += L.omega * L.dinv[i] * (b[i] - ax[i]);  // post-smooth
}
}  // namespace

extern "C" {
// PCG with the AMG V-cycle, PETSc's default test ||M^-1 r|| <= rtol ||M^-1 b||.  Returns the iteration count
// (-1: not converged); levels/sizes/complexity/singular describe the hierarchy.
int hs_amg_pcg(int n, const int *rowptr, const int *col, const double *val, const double *b, double *x, double rtol,
               int maxit, int *levels, int *sizes, double *complexity, int *singular, int *has_dense) {
  std::vector<HostLevel> lv;
  bool sing = false;
  double cx = 1.0;
  fb_amg_host::build_hierarchy(n, rowptr, col, val, lv, sing, cx);
  *levels = (int)lv.size();
  for (size_t l = 0; l < lv.size() && l < 12; ++l) sizes[l] = lv[l].n;
  *complexity = cx;
  *singular = sing ? 1 : 0;
  *has_dense = lv.back().Ainv.empty() ? 0 : 1;
  std::vector<double> r(b, b + n), z, p, Ap, xs(n, 0.0);
  vcycle(lv, 0, r, z);
  p = z;
  double rz = 0.0, zz0 = 0.0;
  for (int i = 0; i < n; ++i) {
    rz += r[i] * z[i];
    zz0 += z[i] * z[i];
  }
  int it = 0;
  for (; it < maxit; ++it) {
    csr_mv(lv[0].A, p, Ap);
    double pAp = 0.0;
    for (int i = 0; i < n; ++i) pAp += p[i] * Ap[i];
    if (!(pAp > 0.0)) break;
    const double alpha = rz / pAp;
    for (int i = 0; i < n; ++i) {
      xs[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
    }
    vcycle(lv, 0, r, z);
    double rz2 = 0.0, zz = 0.0;
    for (int i = 0; i < n; ++i) {
      rz2 += r[i] * z[i];
      zz += z[i] * z[i];
    }
    if (zz <= rtol * rtol * zz0) {
      ++it;
      for (int i = 0; i < n; ++i) x[i] = xs[i];
      return it;
    }
    const double beta = rz2 / rz;
    rz = rz2;
    for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  for (int i = 0; i < n; ++i) x[i] = xs[i];
  return -1;
}
void hs_set_closed_form(int on) { g_closed_form = on; }
double hs_supg_tau(const double *X, const double *v, double eps, int p) { return fb_supg_tau(X, v, eps, p); }
int hs_momentum(int dim, int64_t nc, const int *cell_nodes, const double *xyz, int64_t nbf, const int *bf_cell,
                const int *bf_local, double dt, double rho, double mu, double theta, const double *ui, const double *u0,
                const double *p0, int64_t ndofs, double *F, double *J) {
  if (dim == 2) momentum<2>(nc, cell_nodes, xyz, nbf, bf_cell, bf_local, dt, rho, mu, theta, ui, u0, p0, ndofs, F, J);
  else momentum<3>(nc, cell_nodes, xyz, nbf, bf_cell, bf_local, dt, rho, mu, theta, ui, u0, p0, ndofs, F, J);
  return 0;
}
int hs_rhs(int dim, int64_t nc, const int *cell_nodes, const double *xyz, double dt, double rho, double mu, int rotational,
           const double *ui, const double *p1, const double *p0, double *bp, double *bu_grad) {
  if (dim == 2) rhs<2>(nc, cell_nodes, xyz, dt, rho, mu, rotational, ui, p1, p0, bp, bu_grad);
  else rhs<3>(nc, cell_nodes, xyz, dt, rho, mu, rotational, ui, p1, p0, bp, bu_grad);
  return 0;
}
}
