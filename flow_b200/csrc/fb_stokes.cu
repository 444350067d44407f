// Stokes bootstrap solve (stokes.py:13-148) -- placeholder until the GMRES driver lands.
#include "fb_ops.h"

extern "C" int fb_stokes_solve(fb_space *W, fb_space *P, double mu, int forcing, const double *f, int64_t n_ubc,
                               const int64_t *ubc_dofs, const double *ubc_vals, int64_t n_pbc, const int64_t *pbc_dofs,
                               const double *pbc_vals, double tol, int maxit, double *u, double *p, int *iterations) {
  if (!W) return FB_EINVAL;
  return fb_fail(W->mesh->ctx, FB_EINVAL, "fb_stokes_solve: not implemented yet");
}
