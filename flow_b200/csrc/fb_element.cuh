// Element-level math of the hot-path forms: affine simplex geometry, P1/P2 Lagrange
// bases in barycentric coordinates (UFC ordering) and the pointwise
// integrands of the momentum residual / Jacobian (pressure_correction.py:135-144,
// :169-190, :202).  Everything here is FB_HD so that tests/hostsim can compile the
// very same routines for the CPU-only unit tests; the shipped library only ever
// calls them from CUDA kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

template <int D>
struct Elem {
  static constexpr int NV = D + 1;
  static constexpr int NE = (D == 2) ? 3 : 6;
  static constexpr int NL1 = NV;
  static constexpr int NL2 = NV + NE;
};

// local edge e of a simplex joins local vertices edge_v<D>(e,0) < edge_v<D>(e,1)
template <int D>
FB_HD int edge_v(int e, int k) {
  if (D == 2) {
    // (1,2), (0,2), (0,1)
    return k == 0 ? (e == 0 ? 1 : 0) : (e == 2 ? 1 : 2);
  } else {
    // (2,3), (1,3), (1,2), (0,3), (0,2), (0,1)
    const int a = (e == 0) ? 2 : ((e == 1 || e == 2) ? 1 : 0);
    const int b = (e == 0 || e == 1 || e == 3) ? 3 : ((e == 2 || e == 4) ? 2 : 1);
    return k == 0 ? a : b;
  }
}

// grad(lambda_m) (rows) and cell volume from the D+1 vertex coordinates X[v*D+k]
template <int D>
FB_HD void fb_geometry(const double *X, double glam[D + 1][D], double &vol) {
  if (D == 2) {
    const double ax = X[2] - X[0], ay = X[3] - X[1];
    const double bx = X[4] - X[0], by = X[5] - X[1];
    const double det = ax * by - bx * ay;
    const double inv = 1.0 / det;
    glam[1][0] = by * inv;
    glam[1][1] = -bx * inv;
    glam[2][0] = -ay * inv;
    glam[2][1] = ax * inv;
    glam[0][0] = -(glam[1][0] + glam[2][0]);
    glam[0][1] = -(glam[1][1] + glam[2][1]);
    vol = 0.5 * (det < 0 ? -det : det);
  } else {
    const double a0 = X[3] - X[0], a1 = X[4] - X[1], a2 = X[5] - X[2];
    const double b0 = X[6] - X[0], b1 = X[7] - X[1], b2 = X[8] - X[2];
    const double c0 = X[9] - X[0], c1 = X[10] - X[1], c2 = X[11] - X[2];
    // cross products
    const double bc0 = b1 * c2 - b2 * c1, bc1 = b2 * c0 - b0 * c2, bc2 = b0 * c1 - b1 * c0;
    const double ca0 = c1 * a2 - c2 * a1, ca1 = c2 * a0 - c0 * a2, ca2 = c0 * a1 - c1 * a0;
    const double ab0 = a1 * b2 - a2 * b1, ab1 = a2 * b0 - a0 * b2, ab2 = a0 * b1 - a1 * b0;
    const double det = a0 * bc0 + a1 * bc1 + a2 * bc2;
    const double inv = 1.0 / det;
    glam[1][0] = bc0 * inv;
    glam[1][1] = bc1 * inv;
    glam[1][2] = bc2 * inv;
    glam[2][0] = ca0 * inv;
    glam[2][1] = ca1 * inv;
    glam[2][2] = ca2 * inv;
    glam[3][0] = ab0 * inv;
    glam[3][1] = ab1 * inv;
    glam[3][2] = ab2 * inv;
    for (int k = 0; k < 3; ++k) glam[0][k] = -(glam[1][k] + glam[2][k] + glam[3][k]);
    vol = (det < 0 ? -det : det) / 6.0;
  }
}

// P2 basis function `a` at barycentric point lam
template <int D>
FB_HD double fb_p2_phi(int a, const double *lam) {
  if (a <= D) return lam[a] * (2.0 * lam[a] - 1.0);
  const int e = a - (D + 1);
  return 4.0 * lam[edge_v<D>(e, 0)] * lam[edge_v<D>(e, 1)];
}

// physical gradient of P2 basis function `a` at lam
template <int D>
FB_HD void fb_p2_grad(int a, const double *lam, const double glam[D + 1][D], double g[D]) {
  if (a <= D) {
    const double s = 4.0 * lam[a] - 1.0;
    for (int k = 0; k < D; ++k) g[k] = s * glam[a][k];
  } else {
    const int e = a - (D + 1);
    const int va = edge_v<D>(e, 0), vb = edge_v<D>(e, 1);
    const double sa = 4.0 * lam[vb], sb = 4.0 * lam[va];
    for (int k = 0; k < D; ++k) g[k] = sa * glam[va][k] + sb * glam[vb][k];
  }
}

// second derivatives d_k d_i phi_a (constant per cell)
template <int D>
FB_HD void fb_p2_hess(int a, const double glam[D + 1][D], double H[D][D]) {
  if (a <= D) {
    for (int i = 0; i < D; ++i)
      for (int k = 0; k < D; ++k) H[i][k] = 4.0 * glam[a][i] * glam[a][k];
  } else {
    const int e = a - (D + 1);
    const int va = edge_v<D>(e, 0), vb = edge_v<D>(e, 1);
    for (int i = 0; i < D; ++i)
      for (int k = 0; k < D; ++k) H[i][k] = 4.0 * (glam[va][i] * glam[vb][k] + glam[vb][i] * glam[va][k]);
  }
}

// integral of P2 basis function a over the cell divided by the cell volume
template <int D>
FB_HD double fb_p2_mean(int a) {
  if (D == 2) return a <= D ? 0.0 : 1.0 / 3.0;
  return a <= D ? -1.0 / 20.0 : 1.0 / 5.0;
}

// ---- momentum integrands ----------------------------------------------------
// One quadrature point's contribution to the D x D Jacobian block (test a, trial b):
//   J[i][j] += w * { delta_ij [ phi_a phi_b + c1 ((u.grad phi_b) phi_a - (u.grad phi_a) phi_b) + c2 grad phi_a.grad phi_b ]
//                    + c1 ( phi_a phi_b d_j u_i - d_j phi_a phi_b u_i ) + c2 d_i phi_b d_j phi_a }
// with c1 = theta*dt/2 and c2 = theta*dt*mu/rho   (derivative(F1, ui), pressure_correction.py:202)
template <int D>
// c1t: coefficient of the two terms that differentiate the ADVECTING velocity, ((grad u) delta, v) - ((grad v) delta, u);
// c1t = c1 for the reference's fully implicit convection, 0 for the semi-implicit linearisation (u = advecting field)
FB_HD void fb_jac_point(double w, double c1, double c2, double pa, double pb, const double ga[D], const double gb[D],
                        const double u[D], const double gu[D][D], double J[D][D], double c1t_over_c1 = 1.0) {
  double uga = 0.0, ugb = 0.0, gab = 0.0;
  for (int k = 0; k < D; ++k) {
    uga += u[k] * ga[k];
    ugb += u[k] * gb[k];
    gab += ga[k] * gb[k];
  }
  const double m = w * pa * pb;
  const double dg = m + w * (c1 * (ugb * pa - uga * pb) + c2 * gab);
  const double c1m = c1 * c1t_over_c1 * m, c1wb = c1 * c1t_over_c1 * w * pb, c2w = c2 * w;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double v = c1m * gu[i][j] - c1wb * ga[j] * u[i] + c2w * gb[i] * ga[j];
      if (i == j) v += dg;
      J[i][j] += v;
    }
}

// ---- closed-form element Jacobian ----------------------------------------------------
// All integrands of the Jacobian are polynomials on an affine cell, and grad phi_a is affine: it is the P1
// interpolant of its D+1 vertex values.  With
//   GA[w][k] = d_k phi_a at vertex w,   SA[k] = sum_w GA[w][k],
//   WA[w][k] = (1/|K|) int lambda_w phi_a u_k = sum_c M3[a][c][w] U_c[k]      (M3: fb_p2_tables.h),
//   GU[v][i][j] = d_j u_i at vertex v = sum_c U_c[i] GC[v][j],
//   int lambda_v lambda_w = |K| (1 + delta_vw) / ((D+1)(D+2))
// the block (test a, trial b) is, exactly (same value as the degree-5 quadrature of fb_jac_point, ~160
// instead of ~700 fused multiply-adds per block in 3D):
//   J[i][j] = |K| { delta_ij [ m_ab + c1 (C_ab - C_ba) + c2 tr G ] + c1 (T2[i][j] - T3[i][j]) + c2 G[i][j] }
//   m_ab = sum_v M3[a][b][v]                 C_ab = sum_{w,k} GB[w][k] WA[w][k]   (= int (u.grad phi_b) phi_a / |K|)
//   G[i][j] = (SB[i] SA[j] + sum_v GB[v][i] GA[v][j]) / ((D+1)(D+2))              (= int d_i phi_b d_j phi_a / |K|)
//   T2[i][j] = sum_v M3[a][b][v] GU[v][i][j]                                      (= int phi_a phi_b d_j u_i / |K|)
//   T3[i][j] = sum_w GA[w][j] WB[w][i]                                            (= int phi_b u_i d_j phi_a / |K|)
template <int D>
FB_HD void fb_p2_vertex_grad(int a, int w, const double glam[D + 1][D], double g[D]) {
  if (a <= D) {
    const double s = (a == w) ? 3.0 : -1.0;  // 4 lambda_a - 1 at vertex w
    for (int k = 0; k < D; ++k) g[k] = s * glam[a][k];
  } else {
    const int e = a - (D + 1);
    const int va = edge_v<D>(e, 0), vb = edge_v<D>(e, 1);
    const double sa = (w == vb) ? 4.0 : 0.0, sb = (w == va) ? 4.0 : 0.0;
    for (int k = 0; k < D; ++k) g[k] = sa * glam[va][k] + sb * glam[vb][k];
  }
}

template <int D>
FB_HD void fb_jac_pair(double vol, double c1, double c2, const double *GA, const double *SA, const double *GB,
                       const double *SB, const double *WA, const double *WB, const double *GU, const double *m3,
                       double J[D][D], double c1t = -1.0) {
  if (c1t < 0.0) c1t = c1;  // coefficient of T2 - T3 (terms differentiating the advecting velocity); 0: semi-implicit
  constexpr int NV = D + 1;
  const double lm = 1.0 / ((D + 1) * (D + 2));
  double gb[NV * D], wb[NV * D], mv[NV], sb[D];  // trial-node operands in registers (shared memory on the device)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int t = 0; t < NV * D; ++t) {
    gb[t] = GB[t];
    wb[t] = WB[t];
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int v = 0; v < NV; ++v) mv[v] = m3[v];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < D; ++k) sb[k] = SB[k];
  double mab = 0.0, cab = 0.0, cba = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int v = 0; v < NV; ++v) mab += mv[v];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int t = 0; t < NV * D; ++t) {
    cab += gb[t] * WA[t];
    cba += GA[t] * wb[t];
  }
  double kab = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < D; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < D; ++j) {
      double g = sb[i] * SA[j], t2 = 0.0, t3 = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int v = 0; v < NV; ++v) {
        g += gb[v * D + i] * GA[v * D + j];
        t2 += mv[v] * GU[(v * D + i) * D + j];
        t3 += GA[v * D + j] * wb[v * D + i];
      }
      g *= lm;
      if (i == j) kab += g;
      J[i][j] = c1t * (t2 - t3) + c2 * g;
    }
  }
  const double dg = mab + c1 * (cab - cba) + c2 * kab;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < D; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < D; ++j) J[i][j] = vol * (J[i][j] + (i == j ? dg : 0.0));
  }
}

// One quadrature point's contribution to  coef * R_cell(u; phi_a e_i)  without the forcing term:
//   R = -rho/2 [ ((grad u)u)_i phi_a - (u.grad phi_a) u_i ] - 2 mu eps(u)_ik d_k phi_a + p0 d_i phi_a
// (pressure_correction.py:138-141).  Returns the value for component i.
template <int D>
// adv: advecting velocity at the point (null: u itself, the reference's form; semi-implicit linearisation: u0)
FB_HD double fb_rhs_point(int i, double rho, double mu, double pa, const double ga[D], const double u[D],
                          const double gu[D][D], double p0, const double *adv = nullptr) {
  const double *wv = adv ? adv : u;
  double conv = 0.0, uga = 0.0, visc = 0.0;
  for (int k = 0; k < D; ++k) {
    conv += gu[i][k] * wv[k];
    uga += wv[k] * ga[k];
    visc += (gu[i][k] + gu[k][i]) * ga[k];
  }
  return -0.5 * rho * (conv * pa - uga * u[i]) - mu * visc + p0 * ga[i];
}

// ---- exterior-facet integrands (pressure_correction.py:142-143) ------------------
// phi_a vanishes identically on facet f (opposite local vertex f) unless node a lies on it
template <int D>
FB_HD bool fb_node_on_facet(int a, int f) {
  if (a <= D) return a != f;
  return edge_v<D>(a - (D + 1), 0) != f && edge_v<D>(a - (D + 1), 1) != f;
}

// barycentric point of the parent cell from facet-rule point q (qlam: nq x D)
template <int D>
FB_HD void fb_facet_point(int f, const double *qlam, int q, double lam[D + 1]) {
  int n = 0;
  for (int m = 0; m <= D; ++m) lam[m] = (m == f) ? 0.0 : qlam[q * D + n++];
}

// R_facet(u; phi_a e_i) = sum_q w_q phi_a [ -p0 n_i + mu ((grad u)^T n)_i ] |facet|
// Ue: NL2 x D cell coefficients of u, p0e: D+1 vertex values of p0
template <int D>
FB_HD double fb_facet_F(int ta, int ti, int f, const double glam[D + 1][D], double vol, const double *qlam,
                        const double *qw, int nq, const double *Ue, const double *p0e, double mu) {
  constexpr int NL = Elem<D>::NL2;
  double nA[D];  // outward normal times facet measure
  for (int k = 0; k < D; ++k) nA[k] = -glam[f][k] * (D * vol);
  double acc = 0.0;
  for (int q = 0; q < nq; ++q) {
    double lam[D + 1];
    fb_facet_point<D>(f, qlam, q, lam);
    const double pa = fb_p2_phi<D>(ta, lam);
    double p0 = 0.0;
    for (int v = 0; v <= D; ++v) p0 += p0e[v] * lam[v];
    double gtn = 0.0;  // sum_k d_i u_k n_k
    for (int b = 0; b < NL; ++b) {
      double gb[D];
      fb_p2_grad<D>(b, lam, glam, gb);
      double un = 0.0;
      for (int k = 0; k < D; ++k) un += Ue[b * D + k] * nA[k];
      gtn += gb[ti] * un;
    }
    acc += qw[q] * pa * (-p0 * nA[ti] + mu * gtn);
  }
  return acc;
}

// B[i][j] = sum_q w_q phi_a d_i phi_b n_j |facet|   (derivative of ((grad u)^T n, v)_ds)
template <int D>
FB_HD void fb_facet_J(int ta, int tb, int f, const double glam[D + 1][D], double vol, const double *qlam,
                      const double *qw, int nq, double B[D][D]) {
  double nA[D], gs[D];
  for (int k = 0; k < D; ++k) {
    nA[k] = -glam[f][k] * (D * vol);
    gs[k] = 0.0;
  }
  for (int q = 0; q < nq; ++q) {
    double lam[D + 1], gb[D];
    fb_facet_point<D>(f, qlam, q, lam);
    const double pa = fb_p2_phi<D>(ta, lam);
    fb_p2_grad<D>(tb, lam, glam, gb);
    for (int k = 0; k < D; ++k) gs[k] += qw[q] * pa * gb[k];
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) B[i][j] = gs[i] * nA[j];
}

// ---- pressure / correction right-hand sides per cell -----------------------------
// be[v] = -rho/dt (div u, psi_v) + (grad p0, grad psi_v) [- mu (grad div u, grad psi_v)]
// (pressure_correction.py:318-323); q2lam/q2w: degree-2 cell rule
template <int D>
FB_HD void fb_pressure_rhs_cell(const double glam[D + 1][D], double vol, const double *q2lam, const double *q2w, int nq,
                                const double *Ue, const double *p0e, double dt, double rho, double mu, int rotational,
                                double be[D + 1]) {
  constexpr int NL = Elem<D>::NL2;
  for (int v = 0; v <= D; ++v) be[v] = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double *lam = q2lam + q * (D + 1);
    double div = 0.0;
    for (int a = 0; a < NL; ++a) {
      double g[D];
      fb_p2_grad<D>(a, lam, glam, g);
      for (int i = 0; i < D; ++i) div += Ue[a * D + i] * g[i];
    }
    const double w = -rho / dt * q2w[q] * vol * div;
    for (int v = 0; v <= D; ++v) be[v] += w * lam[v];
  }
  double gv[D];
  for (int k = 0; k < D; ++k) {
    gv[k] = 0.0;
    for (int v = 0; v <= D; ++v) gv[k] += p0e[v] * glam[v][k];
  }
  if (rotational) {
    for (int a = 0; a < NL; ++a) {
      double H[D][D];
      fb_p2_hess<D>(a, glam, H);
      for (int k = 0; k < D; ++k)
        for (int i = 0; i < D; ++i) gv[k] -= mu * Ue[a * D + i] * H[i][k];
    }
  }
  for (int v = 0; v <= D; ++v) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) s += gv[k] * glam[v][k];
    be[v] += vol * s;
  }
}

// grad(phi), phi = p1 - p0 [+ mu div u]  (pressure_correction.py:444-446), constant per cell
template <int D>
FB_HD void fb_correction_gradphi(const double glam[D + 1][D], const double *Ue, const double *dpe, double mu,
                                 int rotational, double gphi[D]) {
  constexpr int NL = Elem<D>::NL2;
  for (int k = 0; k < D; ++k) {
    gphi[k] = 0.0;
    for (int v = 0; v <= D; ++v) gphi[k] += dpe[v] * glam[v][k];
  }
  if (rotational) {
    for (int a = 0; a < NL; ++a) {
      double H[D][D];
      fb_p2_hess<D>(a, glam, H);
      for (int k = 0; k < D; ++k)
        for (int i = 0; i < D; ++i) gphi[k] += mu * Ue[a * D + i] * H[i][k];
    }
  }
}

// ---- SUPG stabilisation parameter (flow/stabilization.py:50-143, triangles only) -------------
// tau = h^2/(4 eps p) xi(Pe),  h = element diameter in the direction of the convection v
// (:74-111), Pe = |v| h / (2 p eps), xi = (coth Pe - 1/Pe)/Pe with its Taylor branch (:123-125).
// X: the 3 vertex coordinates (x0,y0,x1,y1,x2,y2).  Returns a negative value if tau > 1e3,
// where the reference throws (:132-140).
FB_HD double fb_supg_tau(const double *X, const double v[2], double eps, int p) {
  const double conv_norm = sqrt(v[0] * v[0] + v[1] * v[1]);
  if (conv_norm < 1.0e-10) return 0.0;  // :64-68
  const double ax = X[2] - X[0], ay = X[3] - X[1], bx = X[4] - X[0], by = X[5] - X[1];
  const double det = ax * by - bx * ay;
  const double area = 0.5 * (det < 0 ? -det : det);
  double sum = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j) {
      const double e0 = X[2 * i] - X[2 * j], e1 = X[2 * i + 1] - X[2 * j + 1];
      const double t = e1 * v[0] - e0 * v[1];
      sum += t < 0 ? -t : t;
    }
  const double h = 4.0 * conv_norm * area / sum;
  const double Pe = 0.5 * conv_norm * h / (p * eps);
  const double xi = Pe > 1.0e-5 ? (1.0 / tanh(Pe) - 1.0 / Pe) / Pe : 1.0 / 3.0 - Pe * Pe / 45.0 + 2.0 / 945.0 * Pe * Pe * Pe * Pe;
  const double tau = h * h / 4.0 / eps / p * xi;
  return tau > 1.0e3 ? -1.0 : tau;
}
