// fp32 inner solver of the momentum preconditioner (opts.inner_fp32): a fixed number of Jacobi-preconditioned CG
// iterations on the scalar operator S = M + theta dt nu K applied to the D velocity components, with S, the
// vectors and the products in fp32 and the reductions in fp64.  It is used ONLY as the variable preconditioner of
// the flexible GMRES in fb_krylov.cu: the outer iteration, its residual test, the Newton residual and every vector
// that leaves the solver are fp64, so a less accurate z = S^-1 v costs outer iterations, not accuracy.
// Single-GPU only (the halo exchange moves fp64 vectors); partitioned runs use the fp64 inner CG.
#include "fb_ops.h"

namespace {

constexpr int S32_RZ = 42;   // red[42 + parity]: r.z of the current / next iterate
constexpr int S32_PAP = 44;  // p.(S p)

inline int vgrid32(fb_ctx *ctx, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int cap = ctx->dev->sm_count * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// y = S x for NC interleaved components (identity on masked dofs), red[slot] = x.y ; T lanes per row, U chunks in flight
template <int NC, int T, int U>
__global__ void __launch_bounds__(256)
    k_spmm32(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
             const uint8_t *__restrict__ mask, const float *__restrict__ x, float *__restrict__ y, double *partials,
             unsigned int *counter, double *red, int slot) {
  const int lane = threadIdx.x % T;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / T;
  const int64_t npad = ((nrows + ngroups - 1) / ngroups) * ngroups;
  double d[1] = {0.0};
  for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / T; row < npad; row += ngroups) {
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0f;
    float xd = 0.0f;
    bool masked = false;
    if (row < nrows && lane < NC) {
      xd = x[row * NC + lane];
      if (mask) masked = mask[row * NC + lane] != 0;
    }
    if (row < nrows) {
      const int r0 = rowptr[row], r1 = rowptr[row + 1];
      for (int k0 = r0 + lane; k0 < r1; k0 += T * U) {
        int64_t j[U];
        float a[U], xv[U][NC];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = (k0 + u * T < r1) ? col[k0 + u * T] : -1;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int c = 0; c < NC; ++c) xv[u][c] = (j[u] >= 0) ? x[j[u] * NC + c] : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = (j[u] >= 0) ? __ldcs(&val[k0 + u * T]) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] += a[u] * xv[u][c];
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (row < nrows && lane < NC) {
      float yc = acc[0];
#pragma unroll
      for (int c = 1; c < NC; ++c)
        if (lane == c) yc = acc[c];
      if (masked) yc = xd;
      y[row * NC + lane] = yc;
      d[0] += (double)xd * (double)yc;
    }
  }
  fb_grid_reduce<1>(d, partials, counter, red, slot);
}

// r = v, z = D^-1 r, p = z, x = 0 ; red[S32_RZ] = r.z
__global__ void k32_start(int64_t n, const double *__restrict__ v, const float *__restrict__ dinv, float *r, float *z, float *p,
                          float *x, double *partials, unsigned int *counter, double *red) {
  double d[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float ri = (float)v[i];
    const float zi = dinv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    p[i] = zi;
    x[i] = 0.0f;
    d[0] += (double)ri * (double)zi;
  }
  fb_grid_reduce<1>(d, partials, counter, red, S32_RZ);
}

// x += alpha p, r -= alpha Ap, z = D^-1 r ; red[S32_RZ + next parity] = r.z
__global__ void k32_update(int64_t n, int it, const float *__restrict__ dinv, const float *__restrict__ p,
                           const float *__restrict__ Ap, float *x, float *r, float *z, double *partials, unsigned int *counter,
                           double *red) {
  const int par = it & 1;
  const double pAp = red[S32_PAP];
  const float alpha = pAp > 0.0 ? (float)(red[S32_RZ + par] / pAp) : 0.0f;
  double d[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const float ri = r[i] - alpha * Ap[i];
    const float zi = dinv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    d[0] += (double)ri * (double)zi;
  }
  fb_grid_reduce<1>(d, partials, counter, red, S32_RZ + (par ^ 1));
}

__global__ void k32_direction(int64_t n, int it, const float *__restrict__ z, float *p, const double *red) {
  const int par = it & 1;
  const double rz = red[S32_RZ + par];
  const float beta = rz > 0.0 ? (float)(red[S32_RZ + (par ^ 1)] / rz) : 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = z[i] + beta * p[i];
}

__global__ void k32_finish(int64_t n, const float *__restrict__ x, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)x[i];
}

template <int NC>
void spmm32(fb_ctx *ctx, const Inner32 &in, const float *x, float *y) {
  fb_device_state *dv = ctx->dev;
  constexpr int T = 8;
  int64_t g = ((in.nrows + (256 / T) - 1) / (256 / T));
  const int cap = dv->sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  FB_LAUNCH(ctx, (k_spmm32<NC, T, 2>), (int)g, 256, 0, in.nrows, in.rowptr, in.col, in.val, in.mask, x, y, dv->partials,
            dv->counter, dv->red, S32_PAP);
}

}  // namespace

void Inner32::ensure(int64_t n) {
  for (DBuf<float> *b : {&r, &z, &p, &Ap, &x}) b->alloc((size_t)n);
}

// z = (its CG iterations on S, from zero) applied to v.  No host synchronisation.
void inner32_apply(fb_ctx *ctx, Inner32 &in, const double *v, double *z_out) {
  fb_device_state *dv = ctx->dev;
  const int64_t n = in.nrows * in.ncomp;
  in.ensure(n);
  const int g = vgrid32(ctx, n);
  FB_LAUNCH(ctx, k32_start, g, 256, 0, n, v, in.dinv, in.r.p, in.z.p, in.p.p, in.x.p, dv->partials, dv->counter, dv->red);
  for (int it = 0; it < in.its; ++it) {
    if (in.ncomp == 2)
      spmm32<2>(ctx, in, in.p.p, in.Ap.p);
    else
      spmm32<3>(ctx, in, in.p.p, in.Ap.p);
    FB_LAUNCH(ctx, k32_update, g, 256, 0, n, it, in.dinv, in.p.p, in.Ap.p, in.x.p, in.r.p, in.z.p, dv->partials, dv->counter,
              dv->red);
    if (it + 1 < in.its) FB_LAUNCH(ctx, k32_direction, g, 256, 0, n, it, in.z.p, in.p.p, dv->red);
  }
  FB_LAUNCH(ctx, k32_finish, g, 256, 0, n, in.x.p, z_out);
}
