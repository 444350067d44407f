/* CPU throughput baseline (TEST / BENCH INFRASTRUCTURE, not a product path).
 *
 * Multi-threaded fp64 CSR SpMV, Jacobi-PCG and Jacobi-BiCGStab: what PETSc's MatMult /
 * KSPCG / KSPBCGS would do for the reference on the host cores [EXT, PETSc is not
 * installable here, SURVEY.md 8c-d].  Used by oracle/solvers.py when the oracle runs
 * in "krylov" mode and by bench.py's cpu_baseline / --impl reference legs.
 * Build: make -C oracle/cbaseline  ->  oracle/_cbaseline.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int cb_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void cb_spmv(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) s += data[k] * x[indices[k]];
    y[i] = s;
  }
}

static double dot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* PCG, x0 = 0, stop when ||D^-1 r|| <= rtol ||D^-1 b|| (PETSc default preconditioned norm).
 * Returns iterations, or -1 if maxit was reached. */
int cb_pcg(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data, const double *dinv,
           const double *b, double *x, double rtol, int maxit) {
  double *r = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n),
         *Ap = malloc(sizeof(double) * n);
  int it, result = -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = dinv[i] * b[i];
    p[i] = z[i];
  }
  double rz = dot(n, r, z);
  const double ref = sqrt(dot(n, z, z));
  if (ref == 0.0) {
    result = 0;
    goto done;
  }
  for (it = 1; it <= maxit; ++it) {
    cb_spmv(n, indptr, indices, data, p, Ap);
    const double alpha = rz / dot(n, p, Ap);
    double rz_new = 0.0, zz = 0.0;
#pragma omp parallel for reduction(+ : rz_new, zz) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
      z[i] = dinv[i] * r[i];
      rz_new += r[i] * z[i];
      zz += z[i] * z[i];
    }
    if (sqrt(zz) <= rtol * ref) {
      result = it;
      goto done;
    }
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
done:
  free(r);
  free(z);
  free(p);
  free(Ap);
  return result;
}

/* right-preconditioned BiCGStab with Jacobi, x0 = 0, stop when ||r||_2 <= atol. */
int cb_bicgstab(int64_t n, const int64_t *indptr, const int32_t *indices, const double *data, const double *dinv,
                const double *b, double *x, double atol, int maxit) {
  double *r = malloc(sizeof(double) * n), *rh = malloc(sizeof(double) * n), *p = calloc(n, sizeof(double)),
         *v = calloc(n, sizeof(double)), *ph = malloc(sizeof(double) * n), *sh = malloc(sizeof(double) * n),
         *t = malloc(sizeof(double) * n);
  int it, result = -1;
  memcpy(r, b, sizeof(double) * n);
  memcpy(rh, b, sizeof(double) * n);
  memset(x, 0, sizeof(double) * n);
  double rho = 1, alpha = 1, omega = 1;
  if (sqrt(dot(n, r, r)) <= atol) {
    result = 0;
    goto done;
  }
  for (it = 1; it <= maxit; ++it) {
    const double rho_new = dot(n, rh, r);
    const double beta = (it == 1) ? 0.0 : (rho_new / rho) * (alpha / omega);
    rho = rho_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      p[i] = r[i] + beta * (p[i] - omega * v[i]);
      ph[i] = dinv[i] * p[i];
    }
    cb_spmv(n, indptr, indices, data, ph, v);
    alpha = rho / dot(n, rh, v);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      r[i] -= alpha * v[i];
      sh[i] = dinv[i] * r[i];
    }
    cb_spmv(n, indptr, indices, data, sh, t);
    const double tt = dot(n, t, t);
    omega = tt > 0 ? dot(n, t, r) / tt : 0.0;
    double rr = 0.0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * ph[i] + omega * sh[i];
      r[i] -= omega * t[i];
      rr += r[i] * r[i];
    }
    if (sqrt(rr) <= atol) {
      result = it;
      goto done;
    }
    if (omega == 0.0 || rho == 0.0) break;
  }
done:
  free(r);
  free(rh);
  free(p);
  free(v);
  free(ph);
  free(sh);
  free(t);
  return result;
}
