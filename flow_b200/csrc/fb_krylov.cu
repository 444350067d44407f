// Device-resident Krylov solvers: Jacobi-PCG (pressure Poisson, velocity correction,
// projections) and right-preconditioned BiCGStab (momentum Newton systems, heat).
//
// Replaces PETSc KSPCG / the LU solves of the reference (pressure_correction.py:325-339,
// :414-432, :451-464; Newton's linear solves :224-254; heat.py:117-121).  No host
// round-trip per iteration: every scalar (alpha, beta, omega, rho) is recomputed inside
// the kernels from reduction slots that live in device memory; the stopping test runs on
// the device (in the finishing block of the reducing kernel on one GPU, in a one-thread
// kernel behind the NCCL all-reduce on several) and raises a flag that turns all later
// kernels of the batch into no-ops, so the result does not depend on how many iterations
// the host enqueues between polls.  Multi-GPU: vectors hold owned entries first, ghosts
// after; all vector kernels and dots run over the owned part, the SpMV wrapper refreshes
// ghosts, each reducing kernel is followed by ONE all-reduce of its 1-2 slots.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "fb_ops.h"

namespace {

inline int vgrid(fb_ctx *ctx, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int cap = ctx->dev->sm_count * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

constexpr int S_PAP = 4, S_TOL2 = 5;
constexpr int S_BTOL2 = 12;
constexpr int S_REF = 40;  // warm starts: ||M^-1 b||^2 of the ORIGINAL right-hand side (the iteration runs on r0)

// ---------------------------------------------------------------- PCG
__device__ void cg_start_check(double *red, double rtol, double warm_extra2, int warm, int *flag, int *iters) {
  // cold: red[1] = ||M^-1 b||^2 (+ squared Dirichlet values, folded in by block 0)
  // warm: red[1] = ||M^-1 r0||^2, the reference norm comes from the original b
  const double ref2 = warm ? red[S_REF] + warm_extra2 : red[1];
  red[S_TOL2] = rtol * rtol * ref2;
  *iters = 0;
  *flag = (ref2 == 0.0) ? 1 : ((ref2 != ref2 || red[1] != red[1]) ? 3 : ((warm && red[1] <= red[S_TOL2]) ? 1 : 0));
}

__device__ void cg_update_check(const double *red, int it, int *flag, int *iters) {
  const int nxt = (it & 1) ^ 1;
  const double pAp = red[S_PAP], zz = red[2 * nxt + 1];
  if (!(pAp > 0.0) || zz != zz) {
    *flag = (zz != zz || pAp != pAp) ? 3 : 2;
    *iters = it + 1;
  } else if (zz <= red[S_TOL2]) {
    *flag = 1;
    *iters = it + 1;
  }
}

__global__ void k_cg_start(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv, double *r, double *z,
                           double *p, double *x, double rtol, double ref_extra2, int warm, int dist, double *partials,
                           unsigned int *counter, double *red, int *flag, int *iters) {
  double d[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = b[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    p[i] = zi;
    if (!warm) x[i] = 0.0;  // warm: x holds the initial guess and b its residual
    d[0] += ri * zi;
    d[1] += zi * zi;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && !warm) d[1] += ref_extra2;
  const bool last = fb_grid_reduce<2>(d, partials, counter, red, 0);
  if (last && threadIdx.x == 0 && !dist) cg_start_check(red, rtol, ref_extra2, warm, flag, iters);
}
__global__ void k_cg_start_check(double *red, double rtol, double warm_extra2, int warm, int *flag, int *iters) {
  cg_start_check(red, rtol, warm_extra2, warm, flag, iters);
}

// red[S_REF] = || M^-1 b ||^2 with M^-1 = diag(dinv), or || zb ||^2 if dinv == null (zb = AMG cycle applied to b)
__global__ void k_cg_ref2(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv, double *partials,
                          unsigned int *counter, double *red) {
  double d[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double zi = dinv ? dinv[i] * b[i] : b[i];
    d[0] += zi * zi;
  }
  fb_grid_reduce<1>(d, partials, counter, red, S_REF);
}

__global__ void k_cg_update(int64_t n, int it, const double *__restrict__ dinv, const double *__restrict__ p,
                            const double *__restrict__ Ap, double *x, double *r, double *z, int dist, double *partials,
                            unsigned int *counter, double *red, int *flag, int *iters) {
  if (*flag) return;
  const int par = it & 1, nxt = par ^ 1;
  const double alpha = red[2 * par] / red[S_PAP];
  double d[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * Ap[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    d[0] += ri * zi;
    d[1] += zi * zi;
  }
  const bool last = fb_grid_reduce<2>(d, partials, counter, red, 2 * nxt);
  if (last && threadIdx.x == 0 && !dist) cg_update_check(red, it, flag, iters);
}
__global__ void k_cg_update_check(const double *red, int it, int *flag, int *iters) {
  if (*flag) return;
  cg_update_check(red, it, flag, iters);
}

__global__ void k_cg_direction(int64_t n, int it, const double *__restrict__ z, double *p, const double *red,
                               const int *flag) {
  if (*flag) return;
  const int par = it & 1, nxt = par ^ 1;
  const double beta = red[2 * nxt] / red[2 * par];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = z[i] + beta * p[i];
}

// ---------------------------------------------------------------- BiCGStab
// slots of parity set s (base 6*s): RHO, RR, RHV, TS, TT
__device__ __forceinline__ int bs(int s, int k) { return 6 * s + k; }

template <int D>
__device__ __forceinline__ void apply_minv(const double *__restrict__ minv, int64_t node, const double in[D], double out[D]) {
  if (D == 1) {
    out[0] = minv[node] * in[0];
  } else {
    const double *B = minv + node * (D * D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) s += B[i * D + j] * in[j];
      out[i] = s;
    }
  }
}

__device__ void bi_start_check(double *red, double atol, int *flag, int *iters) {
  red[S_BTOL2] = atol * atol;
  *iters = 0;
  const double rr = red[bs(0, 1)];
  *flag = (rr != rr) ? 3 : (rr <= atol * atol ? 1 : 0);
}

__device__ void bi_update_check(const double *red, int it, int *flag, int *iters) {
  const int s = it & 1, nx = s ^ 1;
  const double tt = red[bs(s, 4)];
  const double omega = tt > 0.0 ? red[bs(s, 3)] / tt : 0.0;
  const double rr = red[bs(nx, 1)];
  if (rr != rr) {
    *flag = 3;
    *iters = it + 1;
  } else if (rr <= red[S_BTOL2]) {
    *flag = 1;
    *iters = it + 1;
  } else if (red[bs(nx, 0)] == 0.0 || omega == 0.0) {
    *flag = 2;  // breakdown: the host restarts from the true residual
    *iters = it + 1;
  }
}

__global__ void k_bi_start(int64_t n, const double *__restrict__ b, double *r, double *rhat, double *x, double atol,
                           int dist, double *partials, unsigned int *counter, double *red, int *flag, int *iters) {
  double d[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = b[i];
    r[i] = ri;
    rhat[i] = ri;
    x[i] = 0.0;
    d[0] += ri * ri;
    d[1] += ri * ri;
  }
  const bool last = fb_grid_reduce<2>(d, partials, counter, red, bs(0, 0));
  if (last && threadIdx.x == 0 && !dist) bi_start_check(red, atol, flag, iters);
}
__global__ void k_bi_start_check(double *red, double atol, int *flag, int *iters) { bi_start_check(red, atol, flag, iters); }

// p = r + beta (p - omega v);  phat = Minv p        (one thread per node)
template <int D>
__global__ void k_bi_direction(int64_t nnodes, int it, const double *__restrict__ minv, const double *__restrict__ r,
                               const double *__restrict__ v, double *p, double *phat, const double *red, const int *flag) {
  if (*flag) return;
  const int s = it & 1, pv = s ^ 1;
  double beta = 0.0, omega = 0.0;
  if (it > 0) {
    const double alpha_prev = red[bs(pv, 0)] / red[bs(pv, 2)];
    const double tt = red[bs(pv, 4)];
    omega = tt > 0.0 ? red[bs(pv, 3)] / tt : 0.0;
    beta = (red[bs(s, 0)] / red[bs(pv, 0)]) * (omega != 0.0 ? alpha_prev / omega : 0.0);
  }
  for (int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; node < nnodes; node += (int64_t)gridDim.x * blockDim.x) {
    double pn[D], ph[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int64_t k = node * D + i;
      pn[i] = (it == 0) ? r[k] : r[k] + beta * (p[k] - omega * v[k]);
      p[k] = pn[i];
    }
    apply_minv<D>(minv, node, pn, ph);
#pragma unroll
    for (int i = 0; i < D; ++i) phat[node * D + i] = ph[i];
  }
}

// r <- s = r - alpha v ;  shat = Minv s
template <int D>
__global__ void k_bi_half(int64_t nnodes, int it, const double *__restrict__ minv, const double *__restrict__ v, double *r,
                          double *shat, const double *red, int *flag, int *iters) {
  if (*flag) return;
  const int s = it & 1;
  const double rhv = red[bs(s, 2)];
  if (rhv == 0.0 || rhv != rhv) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      *flag = (rhv != rhv) ? 3 : 2;
      *iters = it + 1;
    }
    return;  // every block takes this branch (rhv is grid-uniform and identical on all ranks)
  }
  const double alpha = red[bs(s, 0)] / rhv;
  for (int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; node < nnodes; node += (int64_t)gridDim.x * blockDim.x) {
    double sn[D], sh[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int64_t k = node * D + i;
      sn[i] = r[k] - alpha * v[k];
      r[k] = sn[i];
    }
    apply_minv<D>(minv, node, sn, sh);
#pragma unroll
    for (int i = 0; i < D; ++i) shat[node * D + i] = sh[i];
  }
}

// x += alpha phat + omega shat ; r = s - omega t ; rho_next = rhat.r ; rr = r.r
__global__ void k_bi_update(int64_t n, int it, const double *__restrict__ phat, const double *__restrict__ shat,
                            const double *__restrict__ t, const double *__restrict__ rhat, double *x, double *r, int dist,
                            double *partials, unsigned int *counter, double *red, int *flag, int *iters) {
  if (*flag) return;
  const int s = it & 1, nx = s ^ 1;
  const double alpha = red[bs(s, 0)] / red[bs(s, 2)];
  const double tt = red[bs(s, 4)];
  const double omega = tt > 0.0 ? red[bs(s, 3)] / tt : 0.0;
  double d[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * phat[i] + omega * shat[i];
    const double ri = r[i] - omega * t[i];
    r[i] = ri;
    d[0] += rhat[i] * ri;
    d[1] += ri * ri;
  }
  const bool last = fb_grid_reduce<2>(d, partials, counter, red, bs(nx, 0));
  if (last && threadIdx.x == 0 && !dist) bi_update_check(red, it, flag, iters);
}
__global__ void k_bi_update_check(const double *red, int it, int *flag, int *iters) {
  if (*flag) return;
  bi_update_check(red, it, flag, iters);
}

int poll(fb_ctx *ctx, int *flag_out, int *iters_out) {
  fb_device_state *dv = ctx->dev;
  int *hp = reinterpret_cast<int *>(dv->host_pinned + 32);
  FB_CUDA(cudaMemcpyAsync(hp, dv->flag, sizeof(int), cudaMemcpyDeviceToHost, dv->stream));
  FB_CUDA(cudaMemcpyAsync(hp + 1, dv->iters, sizeof(int), cudaMemcpyDeviceToHost, dv->stream));
  FB_CUDA(cudaStreamSynchronize(dv->stream));
  *flag_out = hp[0];
  *iters_out = hp[1];
  return 0;
}

}  // namespace

int krylov_pcg_single_reduction(fb_ctx *ctx, const LinOp &A, const double *dinv, const double *b, double *x, double rtol,
                                double ref_extra2, int maxit, int check_every, KrylovWork &w, int *iters, fb_amg *amg, int warm);

int krylov_pcg(fb_ctx *ctx, const LinOp &A, const double *dinv, const double *b, double *x, double rtol, double ref_extra2,
               int maxit, int check_every, KrylovWork &w, int *iters, fb_amg *amg, bool warm_start) {
  const int warm = warm_start ? 1 : 0;
  if (warm) {
    // x holds an initial guess: iterate on r0 = b - A x with the SAME stopping test as a cold start
    // (||M^-1 r|| <= rtol ||M^-1 b||, PETSc's default also for a non-zero guess); only the starting point differs.
    fb_device_state *dv = ctx->dev;
    const int64_t n = A.ndofs();
    w.ensure(7, A.nlocal_dofs());
    double *r0 = w.v[5].p, *zb = w.v[6].p;
    const int g = std::min(vgrid(ctx, n), FB_MAX_RED_BLOCKS);
    if (amg) {
      amg_apply(amg, b, zb);
      FB_LAUNCH(ctx, k_cg_ref2, g, 256, 0, n, zb, (const double *)nullptr, dv->partials, dv->counter, dv->red);
    } else {
      FB_LAUNCH(ctx, k_cg_ref2, g, 256, 0, n, b, dinv, dv->partials, dv->counter, dv->red);
    }
    if (fb_is_distributed(ctx)) fb_allreduce_slots(ctx, S_REF, 1);
    spmv(ctx, A, x, r0);
    vec_axpby(ctx, r0, 1.0, b, -1.0, r0, n);
    b = r0;
  }
  if (amg) return krylov_pcg_single_reduction(ctx, A, dinv, b, x, rtol, ref_extra2, maxit, check_every, w, iters, amg, warm);
  // -1 (default): single-reduction CG where the iteration is latency bound (small systems, or several ranks:
  // one all-reduce instead of two), classic PCG where it is bandwidth bound (measured at n = 74: P2 mass x 3,
  // 9.9 M dofs: 0.81 vs 0.98 ms per iteration); 0 / 1 force one of them
  static const int knob = getenv("FB_CG") ? atoi(getenv("FB_CG")) : -1;
  const int variant = knob >= 0 ? knob : ((fb_is_distributed(ctx) || A.ndofs() < (int64_t(1) << 21)) ? 1 : 0);
  if (variant == 1 && A.block == 1)
    return krylov_pcg_single_reduction(ctx, A, dinv, b, x, rtol, ref_extra2, maxit, check_every, w, iters, nullptr, warm);
  fb_device_state *dv = ctx->dev;
  const int64_t n = A.ndofs();
  const int dist = fb_is_distributed(ctx) ? 1 : 0;
  if (dist && warm) {
    // warm starts take the reference norm from red[S_REF] (all-reduced) PLUS the squared Dirichlet values: that term
    // must be the global sum on every rank as well, or the ranks stop at different iterations (cold starts fold the
    // local value into the all-reduced start reduction instead)
    FB_CUDA(cudaMemcpyAsync(dv->red + 45, &ref_extra2, sizeof(double), cudaMemcpyHostToDevice, dv->stream));
    fb_allreduce_slots(ctx, 45, 1);
    FB_CUDA(cudaMemcpyAsync(dv->host_pinned + 41, dv->red + 45, sizeof(double), cudaMemcpyDeviceToHost, dv->stream));
    FB_CUDA(cudaStreamSynchronize(dv->stream));
    ref_extra2 = dv->host_pinned[41];
  }
  w.ensure(4, A.nlocal_dofs());
  double *r = w.v[0].p, *z = w.v[1].p, *p = w.v[2].p, *Ap = w.v[3].p;
  const int g = vgrid(ctx, n);
  FB_LAUNCH(ctx, k_cg_start, g, 256, 0, n, b, dinv, r, z, p, x, rtol, ref_extra2, warm, dist, dv->partials, dv->counter,
            dv->red, dv->flag, dv->iters);
  if (dist) {
    fb_allreduce_slots(ctx, 0, 2);
    FB_LAUNCH(ctx, k_cg_start_check, 1, 1, 0, dv->red, rtol, ref_extra2, warm, dv->flag, dv->iters);
  }
  int flag = 0, done = 0, it = 0;
  if (check_every < 1) check_every = 1;
  while (it < maxit) {
    const int batch = std::min(check_every, maxit - it);
    for (int k = 0; k < batch; ++k, ++it) {
      spmv(ctx, A, p, Ap, 1, p, S_PAP, dv->flag);
      FB_LAUNCH(ctx, k_cg_update, g, 256, 0, n, it, dinv, p, Ap, x, r, z, dist, dv->partials, dv->counter, dv->red,
                dv->flag, dv->iters);
      if (dist) {
        fb_allreduce_slots(ctx, 2 * ((it & 1) ^ 1), 2);
        FB_LAUNCH(ctx, k_cg_update_check, 1, 1, 0, dv->red, it, dv->flag, dv->iters);
      }
      FB_LAUNCH(ctx, k_cg_direction, g, 256, 0, n, it, z, p, dv->red, dv->flag);
    }
    poll(ctx, &flag, &done);
    if (flag) break;
  }
  if (iters) *iters = flag ? done : it;
  if (flag == 1) return FB_OK;
  if (flag >= 2) return FB_ENAN;
  return FB_ENOCONV_KRYLOV;
}

// ---------------------------------------------------------------- single-reduction PCG
// Chronopoulos-Gear CG: per iteration ONE SpMV (with the three dot products gamma = r.u, delta = w.u,
// |u|^2 fused into its epilogue), ONE all-reduce of those three numbers and ONE fused vector kernel.
// Same Krylov space and stopping test as krylov_pcg; used where the iteration is latency bound
// (pressure Poisson: 560 iterations of ~20 us kernels, and every NCCL call costs ~10 us).
namespace {
constexpr int S_CG3 = 20;   // gamma, delta, |u|^2 of the current iterate
constexpr int S_CGST = 24;  // state[2][3]: alpha, gamma, tol2 per parity

__global__ void k_cg3_start(int64_t n, const double *__restrict__ b, const double *__restrict__ dinv, double *r, double *u,
                            double *p, double *s, double *x, int warm) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = b[i];
    r[i] = ri;
    if (dinv) u[i] = dinv[i] * ri;  // null: the caller applies the AMG cycle to r
    p[i] = 0.0;
    s[i] = 0.0;
    if (!warm) x[i] = 0.0;  // warm: x holds the initial guess and b its residual
  }
}

// it: iteration being applied (0-based).  red[S_CG3..] holds gamma_it, delta_it, |u_it|^2 (all-reduced).
__global__ void k_cg3_update(int64_t n, int it, double rtol, double ref_extra2, int warm, const double *__restrict__ dinv,
                             const double *__restrict__ w, double *u, double *r, double *p, double *s, double *x,
                             double *red, int *flag, int *iters) {
  if (*flag) return;
  const int par = it & 1, prv = par ^ 1;
  const double gamma = red[S_CG3], delta = red[S_CG3 + 1], zz = red[S_CG3 + 2];
  double *st = red + S_CGST;
  const double ref2 = (warm ? red[S_REF] : zz) + ref_extra2;  // warm: norm of the original right-hand side
  const double tol2 = (it == 0) ? rtol * rtol * ref2 : st[3 * prv + 2];
  // every thread evaluates the stopping test on the same all-reduced numbers: uniform decision
  const bool bad = (zz != zz) || (gamma != gamma) || (delta != delta);
  const bool conv = (it == 0) ? (ref2 == 0.0 || (warm && zz <= tol2)) : (zz <= tol2);
  double alpha, beta;
  if (it == 0) {
    beta = 0.0;
    alpha = gamma / delta;
  } else {
    const double alpha_old = st[3 * prv], gamma_old = st[3 * prv + 1];
    beta = gamma / gamma_old;
    alpha = gamma / (delta - beta * gamma / alpha_old);
  }
  const bool breakdown = !bad && !conv && !(alpha > 0.0);  // (p, A p) <= 0
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st[3 * par] = alpha;
    st[3 * par + 1] = gamma;
    st[3 * par + 2] = tol2;
    if (bad || conv || breakdown) {
      *flag = bad ? 3 : (conv ? 1 : 2);
      *iters = it;
    }
  }
  if (bad || conv || breakdown) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double pi = u[i] + beta * p[i];
    const double si = w[i] + beta * s[i];
    p[i] = pi;
    s[i] = si;
    x[i] += alpha * pi;
    const double ri = r[i] - alpha * si;
    r[i] = ri;
    if (dinv) u[i] = dinv[i] * ri;
  }
}
}  // namespace

int krylov_pcg_single_reduction(fb_ctx *ctx, const LinOp &A, const double *dinv, const double *b, double *x, double rtol,
                                double ref_extra2, int maxit, int check_every, KrylovWork &w, int *iters, fb_amg *amg, int warm) {
  fb_device_state *dv = ctx->dev;
  if (amg) dinv = nullptr;
  const int64_t n = A.ndofs();
  w.ensure(5, A.nlocal_dofs());
  double *r = w.v[0].p, *u = w.v[1].p, *wv = w.v[2].p, *p = w.v[3].p, *s = w.v[4].p;
  const int g = vgrid(ctx, n);
  if (fb_is_distributed(ctx)) {  // the squared Dirichlet values entering the reference norm are a global sum
    FB_CUDA(cudaMemcpyAsync(dv->red + S_CG3 + 3, &ref_extra2, sizeof(double), cudaMemcpyHostToDevice, dv->stream));
    fb_allreduce_slots(ctx, S_CG3 + 3, 1);
    FB_CUDA(cudaMemcpyAsync(dv->host_pinned, dv->red + S_CG3 + 3, sizeof(double), cudaMemcpyDeviceToHost, dv->stream));
    FB_CUDA(cudaStreamSynchronize(dv->stream));
    ref_extra2 = dv->host_pinned[0];
  }
  FB_CUDA(cudaMemsetAsync(dv->flag, 0, sizeof(int), dv->stream));
  FB_CUDA(cudaMemsetAsync(dv->iters, 0, sizeof(int), dv->stream));
  FB_LAUNCH(ctx, k_cg3_start, g, 256, 0, n, b, dinv, r, u, p, s, x, warm);
  if (amg) amg_apply(amg, r, u);
  int flag = 0, done = 0, it = 0;
  if (check_every < 1) check_every = 1;
  while (it <= maxit) {
    const int batch = std::min(check_every, maxit + 1 - it);
    for (int k = 0; k < batch; ++k, ++it) {
      spmv(ctx, A, u, wv, 3, r, S_CG3, dv->flag);  // w = A u; gamma, delta, |u|^2 -> one all-reduce
      FB_LAUNCH(ctx, k_cg3_update, g, 256, 0, n, it, rtol, ref_extra2, warm, dinv, wv, u, r, p, s, x, dv->red, dv->flag,
                dv->iters);
      if (amg) amg_apply(amg, r, u);  // u = M^-1 r (runs unconditionally; harmless once the flag is up)
    }
    poll(ctx, &flag, &done);
    if (flag) break;
  }
  if (iters) *iters = flag ? done : maxit;
  if (flag == 1) return FB_OK;
  if (flag >= 2) return FB_ENAN;
  return FB_ENOCONV_KRYLOV;
}

constexpr int FB_BREAKDOWN = -100;  // internal: recoverable BiCGStab breakdown

template <int D>
static int bicgstab_impl(fb_ctx *ctx, const LinOp &A, const double *minv, const double *b, double *x, double atol,
                         int maxit, int check_every, KrylovWork &w, int *iters) {
  fb_device_state *dv = ctx->dev;
  const int64_t n = A.ndofs();
  const int64_t nnodes = n / D;
  const int dist = fb_is_distributed(ctx) ? 1 : 0;
  w.ensure(7, A.nlocal_dofs());
  double *r = w.v[0].p, *rhat = w.v[1].p, *p = w.v[2].p, *v = w.v[3].p, *phat = w.v[4].p, *shat = w.v[5].p, *t = w.v[6].p;
  const int g = vgrid(ctx, n), gn = vgrid(ctx, nnodes);
  FB_LAUNCH(ctx, k_bi_start, g, 256, 0, n, b, r, rhat, x, atol, dist, dv->partials, dv->counter, dv->red, dv->flag,
            dv->iters);
  if (dist) {
    fb_allreduce_slots(ctx, 0, 2);
    FB_LAUNCH(ctx, k_bi_start_check, 1, 1, 0, dv->red, atol, dv->flag, dv->iters);
  }
  int flag = 0, done = 0, it = 0;
  if (check_every < 1) check_every = 1;
  poll(ctx, &flag, &done);
  while (!flag && it < maxit) {
    const int batch = std::min(check_every, maxit - it);
    for (int k = 0; k < batch; ++k, ++it) {
      const int s = it & 1;
      FB_LAUNCH(ctx, k_bi_direction<D>, gn, 256, 0, nnodes, it, minv, r, v, p, phat, dv->red, dv->flag);
      spmv(ctx, A, phat, v, 1, rhat, 6 * s + 2, dv->flag);
      FB_LAUNCH(ctx, k_bi_half<D>, gn, 256, 0, nnodes, it, minv, v, r, shat, dv->red, dv->flag, dv->iters);
      spmv(ctx, A, shat, t, 2, r, 6 * s + 3, dv->flag);
      FB_LAUNCH(ctx, k_bi_update, g, 256, 0, n, it, phat, shat, t, rhat, x, r, dist, dv->partials, dv->counter, dv->red,
                dv->flag, dv->iters);
      if (dist) {
        fb_allreduce_slots(ctx, 6 * (s ^ 1), 2);
        FB_LAUNCH(ctx, k_bi_update_check, 1, 1, 0, dv->red, it, dv->flag, dv->iters);
      }
    }
    poll(ctx, &flag, &done);
  }
  if (iters) *iters = flag ? done : it;
  if (flag == 1) return FB_OK;
  if (flag == 2) return FB_BREAKDOWN;
  if (flag == 3) return FB_ENAN;
  return FB_ENOCONV_KRYLOV;
}

static int bicgstab_once(fb_ctx *ctx, const LinOp &A, const double *minv, const double *b, double *x, double atol,
                         int maxit, int check_every, KrylovWork &w, int *iters) {
  if (A.block == 2) return bicgstab_impl<2>(ctx, A, minv, b, x, atol, maxit, check_every, w, iters);
  if (A.block == 3) return bicgstab_impl<3>(ctx, A, minv, b, x, atol, maxit, check_every, w, iters);
  return bicgstab_impl<1>(ctx, A, minv, b, x, atol, maxit, check_every, w, iters);
}

// BiCGStab with restarts: on a (recoverable) breakdown the true residual b - A x is formed and the
// iteration restarted for the correction, the usual remedy for rho ~ 0 / omega ~ 0.
int krylov_bicgstab(fb_ctx *ctx, const LinOp &A, const double *minv, const double *b, double *x, double atol, int maxit,
                    int check_every, KrylovWork &w, int *iters) {
  const int64_t n = A.ndofs();
  int total = 0, its = 0;
  int status = bicgstab_once(ctx, A, minv, b, x, atol, maxit, check_every, w, &its);
  total += its;
  for (int restart = 0; status == FB_BREAKDOWN && restart < 20 && total < maxit; ++restart) {
    w.v[7].alloc((size_t)A.nlocal_dofs());
    w.v[8].alloc((size_t)A.nlocal_dofs());
    double *rhs = w.v[7].p, *e = w.v[8].p;
    spmv(ctx, A, x, rhs);
    vec_axpby(ctx, rhs, 1.0, b, -1.0, rhs, n);  // rhs = b - A x
    status = bicgstab_once(ctx, A, minv, rhs, e, atol, maxit - total, check_every, w, &its);
    total += its;
    vec_axpy(ctx, x, 1.0, e, n);
  }
  if (iters) *iters = total;
  if (status == FB_BREAKDOWN) return FB_ENOCONV_KRYLOV;
  return status;
}

// ---------------------------------------------------------------- flexible GMRES
// Right-preconditioned FGMRES(m) for the momentum Newton systems, with a VARIABLE preconditioner (a few CG
// iterations on the scalar operator S = M + theta dt nu K applied to the D components, see fb_api.cu): the
// Jacobian is S (x) I plus the viscous coupling and the linearised convection, so S^-1 J is well clustered and the
// 10 GB Jacobian is streamed ~15 times per step instead of ~80 (the inner products stream the 1.2 GB S).
// Replaces the sparse LU of the reference's Newton solver (pressure_correction.py:224-254) like krylov_bicgstab.
//
// Arnoldi: classical Gram-Schmidt, the j+2 inner products of one iteration are computed in chunks of four by one
// kernel each (w is re-read per chunk), all-reduced together, and the new basis vector is formed by one fused
// kernel; |w - V h|^2 comes from Pythagoras.  The small least-squares problem lives on the host: one D2H of j+2
// numbers per outer iteration.
namespace {

struct VecPtrs {
  const double *p[32];
};

// red[slot0 + k] = v_k . w for k < nv (nv <= 4); v_k == null -> w . w
__global__ void k_dot4(int64_t n, const double *__restrict__ w, const double *__restrict__ v0, const double *__restrict__ v1,
                       const double *__restrict__ v2, const double *__restrict__ v3, double *partials, unsigned int *counter,
                       double *red, int slot0) {
  double d[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
    d[0] += (v0 ? v0[i] : wi) * wi;
    if (v1) d[1] += v1[i] * wi;
    if (v2) d[2] += v2[i] * wi;
    if (v3) d[3] += v3[i] * wi;
  }
  fb_grid_reduce<4>(d, partials, counter, red, slot0);
}

// vnext = (w - sum_{i<nv} h_i V_i) / sqrt(ww - sum h_i^2), h_i = red[i], ww = red[nv]
__global__ void k_fg_orth(int64_t n, int nv, VecPtrs V, const double *__restrict__ w, const double *__restrict__ red,
                          double *__restrict__ vnext) {
  double h[32], s2 = 0.0;
  for (int k = 0; k < nv; ++k) {
    h[k] = red[k];
    s2 += h[k] * h[k];
  }
  const double nn = red[nv] - s2;
  const double sc = nn > 0.0 ? rsqrt(nn) : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = w[i];
    for (int k = 0; k < nv; ++k) acc -= h[k] * V.p[k][i];
    vnext[i] = sc * acc;
  }
}

struct Coefs {
  double y[32];
};
// x (+)= sum_{i<nv} y_i Z_i
__global__ void k_fg_combine(int64_t n, int nv, VecPtrs Z, Coefs c, int accumulate, double *__restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = accumulate ? x[i] : 0.0;
    for (int k = 0; k < nv; ++k) acc += c.y[k] * Z.p[k][i];
    x[i] = acc;
  }
}

__global__ void k_scale_to(int64_t n, double a, const double *__restrict__ in, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a * in[i];
}

}  // namespace

int krylov_fgmres(fb_ctx *ctx, const LinOp &A, const FgmresPrecond &pc, const double *b, double *x, double atol, int maxit,
                  int m, FgmresWork &fw, int *iters, int *inner_iters, bool warm_start) {
  fb_device_state *dv = ctx->dev;
  cudaStream_t st = dv->stream;
  const int64_t n = A.ndofs(), nl = A.nlocal_dofs();
  if (m > 30) m = 30;
  if (m < 2) m = 2;
  fw.ensure(m, nl);
  const int g = std::min(vgrid(ctx, n), FB_MAX_RED_BLOCKS);
  double *hp = dv->host_pinned;
  std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), gv(m + 1), yv(m);
  int total = 0, inner_total = 0;
  bool converged = false, first = !warm_start;  // warm start: x holds an initial guess, the first cycle starts from b - A x
  if (!warm_start) FB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * nl, st));
  double *w = fw.w.p;
  while (!converged && total < maxit) {
    // r = b - A x  (x = 0 in the first cycle)
    if (first) {
      FB_CUDA(cudaMemcpyAsync(w, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    } else {
      spmv(ctx, A, x, w);
      vec_axpby(ctx, w, 1.0, b, -1.0, w, n);
    }
    first = false;
    FB_LAUNCH(ctx, k_dot4, g, 256, 0, n, w, (const double *)nullptr, (const double *)nullptr, (const double *)nullptr,
              (const double *)nullptr, dv->partials, dv->counter, dv->red, 0);
    fb_allreduce_slots(ctx, 0, 1);
    FB_CUDA(cudaMemcpyAsync(hp, dv->red, sizeof(double), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    const double beta = std::sqrt(hp[0]);
    if (beta != beta) return FB_ENAN;
    if (beta <= atol) {
      converged = true;
      break;
    }
    FB_LAUNCH(ctx, k_scale_to, vgrid(ctx, n), 256, 0, n, 1.0 / beta, w, fw.V[0].p);
    std::fill(gv.begin(), gv.end(), 0.0);
    gv[0] = beta;
    int j = 0;
    for (; j < m && total < maxit; ++j, ++total) {
      int ii = 0;
      const int pst = pc.apply(pc.self, fw.V[j].p, fw.Z[j].p, &ii);
      inner_total += ii;
      if (pst == FB_ENAN) return FB_ENAN;
      spmv(ctx, A, fw.Z[j].p, w);
      // h_i = V_i . w (i <= j), ww = w . w  -> slots 0 .. j+1
      const int nd = j + 2;
      for (int c0 = 0; c0 < nd; c0 += 4) {
        const double *vp[4];
        for (int k = 0; k < 4; ++k) {
          const int idx = c0 + k;
          vp[k] = idx <= j ? fw.V[idx].p : (idx == j + 1 ? w : nullptr);
        }
        // slot idx == j+1 is w.w: pass w itself (v == w); unused entries of the chunk are null
        FB_LAUNCH(ctx, k_dot4, g, 256, 0, n, w, vp[0], vp[1], vp[2], vp[3], dv->partials, dv->counter, dv->red, c0);
      }
      fb_allreduce_slots(ctx, 0, nd);
      VecPtrs V;
      for (int k = 0; k <= j; ++k) V.p[k] = fw.V[k].p;
      FB_LAUNCH(ctx, k_fg_orth, vgrid(ctx, n), 256, 0, n, j + 1, V, w, dv->red, fw.V[j + 1].p);
      FB_CUDA(cudaMemcpyAsync(hp, dv->red, sizeof(double) * nd, cudaMemcpyDeviceToHost, st));
      FB_CUDA(cudaStreamSynchronize(st));
      double s2 = 0.0;
      for (int i = 0; i <= j; ++i) {
        H[(size_t)i * m + j] = hp[i];
        s2 += hp[i] * hp[i];
      }
      const double nn = hp[j + 1] - s2;
      if (hp[j + 1] != hp[j + 1]) return FB_ENAN;
      const double hlast = nn > 0.0 ? std::sqrt(nn) : 0.0;
      H[(size_t)(j + 1) * m + j] = hlast;
      for (int i = 0; i < j; ++i) {  // previous Givens rotations
        const double t = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = t;
      }
      const double a = H[(size_t)j * m + j], bb = H[(size_t)(j + 1) * m + j];
      const double rr = std::hypot(a, bb);
      if (rr == 0.0 || rr != rr) return FB_ENAN;
      cs[j] = a / rr;
      sn[j] = bb / rr;
      H[(size_t)j * m + j] = rr;
      H[(size_t)(j + 1) * m + j] = 0.0;
      gv[j + 1] = -sn[j] * gv[j];
      gv[j] = cs[j] * gv[j];
      // cancellation in the Pythagorean norm (w almost inside the basis): end the cycle, the next one restarts
      // from the true residual
      const bool degenerate = nn <= 1e-10 * hp[j + 1];
      if (std::fabs(gv[j + 1]) <= atol || degenerate) {
        converged = std::fabs(gv[j + 1]) <= atol;
        ++j;
        ++total;
        break;
      }
    }
    // x += Z y,  H y = g
    for (int i = j - 1; i >= 0; --i) {
      double s = gv[i];
      for (int k = i + 1; k < j; ++k) s -= H[(size_t)i * m + k] * yv[k];
      yv[i] = s / H[(size_t)i * m + i];
    }
    VecPtrs Z;
    Coefs c;
    for (int k = 0; k < j; ++k) {
      Z.p[k] = fw.Z[k].p;
      c.y[k] = yv[k];
    }
    FB_LAUNCH(ctx, k_fg_combine, vgrid(ctx, n), 256, 0, n, j, Z, c, 1, x);
  }
  if (iters) *iters = total;
  if (inner_iters) *inner_iters = inner_total;
  return converged ? FB_OK : FB_ENOCONV_KRYLOV;
}


// ---------------------------------------------------------------- Chebyshev polynomial preconditioner
namespace {
__global__ void k_cheb_start(int64_t n, double inv_theta, const double *__restrict__ dinv, const double *__restrict__ v,
                             double *__restrict__ d0, double *__restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = inv_theta * dinv[i] * v[i];
    d0[i] = d;
    if (z) z[i] = d;
  }
}
__global__ void k_lanczos_seed(int64_t n, const uint8_t *__restrict__ mask, double *__restrict__ r) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // deterministic pseudo-random start with all frequencies (hash of the index), zero on constrained dofs
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    r[i] = (mask && mask[i]) ? 0.0 : ((double)(h & 0xFFFFFF) / 8388608.0 - 1.0);
  }
}
double dot_sync(fb_ctx *ctx, const double *x, const double *y, int64_t n) {
  fb_device_state *dv = ctx->dev;
  vec_dot(ctx, x, y, n, 44);
  fb_allreduce_slots(ctx, 44, 1);
  FB_CUDA(cudaMemcpyAsync(dv->host_pinned + 40, dv->red + 44, sizeof(double), cudaMemcpyDeviceToHost, dv->stream));
  FB_CUDA(cudaStreamSynchronize(dv->stream));
  return dv->host_pinned[40];
}
// number of eigenvalues of the symmetric tridiagonal (a, b) below x (Sturm sequence)
int sturm_count(const std::vector<double> &a, const std::vector<double> &b, double x) {
  int cnt = 0;
  double q = 1.0;
  for (size_t i = 0; i < a.size(); ++i) {
    q = a[i] - x - (i > 0 ? b[i - 1] * b[i - 1] / (q != 0.0 ? q : 1e-300) : 0.0);
    if (q < 0.0) ++cnt;
  }
  return cnt;
}
}  // namespace

// Extreme eigenvalues of D^-1 A from `steps` iterations of preconditioned CG (Lanczos tridiagonal, host scalars; set-up
// only).  The largest Ritz value converges fast and from below, the smallest from above: the returned interval is
// widened by 5 % / 15 %.
void cheb_estimate_spectrum(fb_ctx *ctx, const LinOp &A, const double *dinv, int steps, double *lmin, double *lmax) {
  const int64_t n = A.ndofs(), nl = A.nlocal_dofs();
  DBuf<double> r, z, p, Ap;
  r.alloc((size_t)nl);
  z.alloc((size_t)nl);
  p.alloc((size_t)nl);
  Ap.alloc((size_t)nl);
  p.zero(ctx->dev->stream);
  FB_LAUNCH(ctx, k_lanczos_seed, vgrid(ctx, n), 256, 0, n, A.mask, r.p);
  // z = dinv .* r
  FB_LAUNCH(ctx, k_cheb_start, vgrid(ctx, n), 256, 0, n, 1.0, dinv, r.p, z.p, (double *)nullptr);
  vec_axpby(ctx, p.p, 1.0, z.p, 0.0, z.p, n);
  double rz = dot_sync(ctx, r.p, z.p, n);
  std::vector<double> alpha, beta;
  for (int k = 0; k < steps && rz > 0.0; ++k) {
    spmv(ctx, A, p.p, Ap.p);
    const double pAp = dot_sync(ctx, p.p, Ap.p, n);
    if (!(pAp > 0.0)) break;
    const double al = rz / pAp;
    vec_axpy(ctx, r.p, -al, Ap.p, n);
    FB_LAUNCH(ctx, k_cheb_start, vgrid(ctx, n), 256, 0, n, 1.0, dinv, r.p, z.p, (double *)nullptr);
    const double rz_new = dot_sync(ctx, r.p, z.p, n);
    const double be = rz_new / rz;
    alpha.push_back(al);
    beta.push_back(be);
    vec_axpby(ctx, p.p, 1.0, z.p, be, p.p, n);
    if (!(rz_new > 1e-28 * rz)) break;
    rz = rz_new;
  }
  const size_t m = alpha.size();
  if (m == 0) throw fb_cuda_error(FB_ENAN, "cheb_estimate_spectrum: operator is not positive definite");
  std::vector<double> a(m), b(m > 0 ? m - 1 : 0);
  for (size_t j = 0; j < m; ++j) {
    a[j] = 1.0 / alpha[j] + (j > 0 ? beta[j - 1] / alpha[j - 1] : 0.0);
    if (j + 1 < m) b[j] = std::sqrt(beta[j]) / alpha[j];
  }
  double lo = 0.0, hi = 0.0;
  for (size_t j = 0; j < m; ++j) {  // Gershgorin bracket of the tridiagonal
    const double rad = (j > 0 ? std::fabs(b[j - 1]) : 0.0) + (j + 1 < m ? std::fabs(b[j]) : 0.0);
    hi = std::max(hi, a[j] + rad);
    lo = j == 0 ? a[j] - rad : std::min(lo, a[j] - rad);
  }
  auto kth = [&](int k) {  // k-th smallest eigenvalue by bisection
    double x0 = lo, x1 = hi;
    for (int it = 0; it < 200; ++it) {
      const double xm = 0.5 * (x0 + x1);
      if (sturm_count(a, b, xm) > k) x1 = xm; else x0 = xm;
    }
    return 0.5 * (x0 + x1);
  };
  const double emin = kth(0), emax = kth((int)m - 1);
  *lmax = 1.05 * emax;
  *lmin = std::max(0.85 * emin, 1e-6 * emax);
}

void cheb_apply(fb_ctx *ctx, const LinOp &A, const double *dinv, double lmin, double lmax, int degree, const double *v, double *z,
                ChebWork &w) {
  if (!A.tile || !A.tval) throw fb_cuda_error(FB_EINVAL, "cheb_apply: operator is not in tile format");
  const int64_t n = A.ndofs(), nl = A.nlocal_dofs();
  if (w.d0.n != (size_t)nl) {
    w.r.alloc((size_t)nl);
    w.zt.alloc((size_t)nl);
    w.dinv_t.alloc((size_t)nl);
    w.dinv_valid = false;
    w.d0.alloc((size_t)nl);
    w.d1.alloc((size_t)nl);
    // ghost entries of the direction vectors: refreshed by the halo exchange, or -- rank-local polynomial
    // (w.local) -- left at zero for ever (the kernels only write owned rows)
    w.d0.zero(ctx->dev->stream);
    w.d1.zero(ctx->dev->stream);
  }
  if (!w.dinv_valid || w.dinv_src != dinv) {  // the inverse diagonal in the row order of the tiles
    tile_to_tile_order(ctx, *A.tile, A.dofs_per_node(), dinv, w.dinv_t.p);
    w.dinv_src = dinv;
    w.dinv_valid = true;
  }
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  if (degree < 1) degree = 1;
  FB_LAUNCH(ctx, k_cheb_start, vgrid(ctx, n), 256, 0, n, 1.0 / theta, dinv, v, w.d0.p, degree == 1 ? z : (double *)nullptr);
  double rho_prev = 1.0 / sigma;
  double *dk = w.d0.p, *dn = w.d1.p;
  for (int k = 0; k + 1 < degree; ++k) {
    const double rho = 1.0 / (2.0 * sigma - rho_prev);
    const bool last = k + 2 == degree;
    if (A.halo && !w.local) halo_exchange(ctx, *A.halo, dk, A.dofs_per_node());
    // r and the running sum z stay in tile order between the products; the input v and the result z are canonical
    tile_cheb_step(ctx, A, dk, k == 0 ? v : w.r.p, k == 0, w.r.p, k == 0 ? nullptr : w.zt.p, last ? z : w.zt.p, dn, w.dinv_t.p,
                   rho * rho_prev, 2.0 * rho / delta, last);
    std::swap(dk, dn);
    rho_prev = rho;
  }
}
