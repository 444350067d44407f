// Stokes bootstrap solve: replaces flow.stokes.solve (stokes.py:13-148).
//
//   a = mu (grad u, grad v) - (p, div v) - (q, div u),  L = (f, v)          (stokes.py:40-45)
//   A, b = assemble_system(a, L, bcs)   -- symmetric Dirichlet elimination    (stokes.py:46)
//   GMRES preconditioned with the matrix of  mu (grad u, grad v) - p q       (stokes.py:55-60)
//
// Here: restarted flexible GMRES (right preconditioning) on the saddle-point operator, whose
// velocity block is the assembled scalar P2 stiffness applied to the d interleaved components
// and whose divergence/gradient blocks are applied matrix-free per cell.  The preconditioner is
// the same block-diagonal operator as the reference's (mu K, M_p), each block inverted
// approximately: the velocity block component by component by PCG preconditioned with a smoothed-aggregation AMG
// V-cycle on mu K (fb_amg.cu; the reference uses one BoomerAMG V-cycle per block, stokes.py:59), the pressure mass
// block by Jacobi-PCG.  Meshes below 4096 nodes keep Jacobi-PCG for the velocity block as well.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "fb_ops.h"

namespace {

__global__ void k_axpy_slot(double *y, double sign, const double *red, const double *__restrict__ x, int64_t n) {
  const double a = sign * red[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void k_scale_rsqrt_slot(double *out, const double *__restrict__ in, const double *red, int64_t n) {
  const double a = red[0] > 0.0 ? 1.0 / sqrt(red[0]) : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * in[i];
}
// component c of an interleaved vector <-> contiguous vector (AMG on the scalar stiffness matrix, component by component)
__global__ void k_take_comp(int64_t nnodes, int D, int c, const double *__restrict__ v, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnodes; i += (int64_t)gridDim.x * blockDim.x) out[i] = v[i * D + c];
}
__global__ void k_put_comp(int64_t nnodes, int D, int c, const double *__restrict__ in, double *__restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnodes; i += (int64_t)gridDim.x * blockDim.x) z[i * D + c] = in[i];
}
__global__ void k_take_comp_mask(int64_t nnodes, int D, int c, const uint8_t *__restrict__ m, uint8_t *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnodes; i += (int64_t)gridDim.x * blockDim.x) out[i] = m[i * D + c];
}
__global__ void k_mask_zero(double *y, const uint8_t *__restrict__ mask, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (mask[i]) y[i] = 0.0;
}

inline int vgrid(fb_ctx *ctx, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int cap = ctx->dev->sm_count * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct Stokes {
  fb_ctx *ctx;
  DevSpace *W, *P;
  int D;
  int64_t nu, np, n;
  double mu;
  fb_mat K, Mp;
  DBuf<uint8_t> mask;  // nu + np
  DBuf<double> dinv_u, dinv_p, tmp;
  KrylovWork kw, kw_p;  // one set of work vectors per block size: DBuf::alloc reallocates (cudaFree + cudaMalloc) on every size change
  int64_t inner_k = 0, inner_m = 0;  // inner PCG iterations spent in the two blocks (FB_VERBOSE)
  // AMG on mu K, one hierarchy per component (each with that component's Dirichlet dofs eliminated symmetrically, so that
  // the V-cycle maps into the space the masked PCG operator works in); components with identical constraints share one
  fb_amg *amgK[3] = {nullptr, nullptr, nullptr};
  DBuf<uint8_t> mask_c;      // D contiguous per-component masks
  DBuf<double> dinv_c;       // Jacobi diagonal of the per-component masked operators (unused by the AMG path of krylov_pcg)
  DBuf<double> cin, cout;    // one component, contiguous
  int64_t nn = 0;
  ~Stokes() {
    for (int c = 0; c < 3; ++c) {
      bool shared = false;
      for (int e = 0; e < c; ++e) shared = shared || amgK[e] == amgK[c];
      if (amgK[c] && !shared) amg_destroy(amgK[c]);
    }
  }

  // y = (I-P) A x on free rows, 0 on constrained rows (x vanishes on constrained dofs)
  void apply(const double *x, double *y) {
    cudaStream_t st = ctx->dev->stream;
    LinOp Kop = make_linop(K, D, nullptr);
    spmv(ctx, Kop, x, y);                                   // K x_u (K already carries mu)
    stokes_grad(ctx, *W, *P, x + nu, y);                        // += B^T x_p
    FB_CUDA(cudaMemsetAsync(y + nu, 0, sizeof(double) * np, st));
    stokes_div(ctx, *W, *P, x, y + nu);                         // B x_u
    FB_LAUNCH(ctx, k_mask_zero, vgrid(ctx, n), 256, 0, y, mask.p, n);
  }

  // z = blockdiag(mu K, M_p)^-1 v, each block by masked Jacobi-PCG to a loose tolerance
  int precond(const double *v, double *z, double rtol_in) {
    int its = 0;
    int st = FB_OK;
    if (amgK[0]) {
      for (int c = 0; c < D; ++c) {
        FB_LAUNCH(ctx, k_take_comp, vgrid(ctx, nn), 256, 0, nn, D, c, v, cin.p);
        LinOp Kc = make_linop(K, 1, mask_c.p + (size_t)c * nn);
        st = krylov_pcg(ctx, Kc, dinv_c.p, cin.p, cout.p, rtol_in, 0.0, 200, 4, kw, &its, amgK[c]);
        inner_k += its;
        if (st == FB_ENAN) return st;
        FB_LAUNCH(ctx, k_put_comp, vgrid(ctx, nn), 256, 0, nn, D, c, cout.p, z);
      }
    } else {
      LinOp Kop = make_linop(K, D, mask.p);
      st = krylov_pcg(ctx, Kop, dinv_u.p, v, z, rtol_in, 0.0, 2000, 25, kw, &its);
      inner_k += its;
      if (st == FB_ENAN) return st;
    }
    LinOp Mop = make_linop(Mp, 1, mask.p + nu);
    st = krylov_pcg(ctx, Mop, dinv_p.p, v + nu, z + nu, rtol_in, 0.0, 500, 10, kw_p, &its);
    inner_m += its;
    if (st == FB_ENAN) return st;
    return FB_OK;
  }
};

}  // namespace

extern "C" int fb_stokes_solve(fb_space *Wsp, fb_space *Psp, double mu, int forcing, const double *f, int64_t n_ubc,
                               const int64_t *ubc_dofs, const double *ubc_vals, int64_t n_pbc, const int64_t *pbc_dofs,
                               const double *pbc_vals, double tol, int maxit, double *u_out, double *p_out,
                               int *iterations) {
  if (!Wsp || !Psp || !u_out || !p_out) return FB_EINVAL;
  fb_ctx *ctx = Wsp->mesh->ctx;
  if (!ctx || !ctx->dev)
    return fb_fail(ctx, FB_ENODEVICE, "this context has no CUDA device (host-only); no CPU compute path exists");
  if (!(mu > 0.0)) return fb_fail(ctx, FB_EINVAL, "fb_stokes_solve: mu must be > 0");  // stokes.py:23
  if (Wsp->mesh != Psp->mesh || Wsp->degree != 2 || Wsp->ncomp != Wsp->mesh->dim || Psp->degree != 1 || Psp->ncomp != 1)
    return fb_fail(ctx, FB_EINVAL, "fb_stokes_solve: need W = vector P2 and P = scalar P1 on one mesh");
  if ((n_ubc > 0 && (!ubc_dofs || !ubc_vals)) || (n_pbc > 0 && (!pbc_dofs || !pbc_vals)))
    return fb_fail(ctx, FB_EINVAL, "fb_stokes_solve: bad Dirichlet arrays");
  try {
    cudaStream_t st = ctx->dev->stream;
    fb_device_state *dv = ctx->dev;
    const auto t_begin = std::chrono::steady_clock::now();
    Stokes S;
    S.ctx = ctx;
    if (!Wsp->dev) {
      std::unique_ptr<DevSpace> d(new DevSpace());
      dev_space_build(Wsp, *d);
      Wsp->dev = d.release();
    }
    if (!Psp->dev) {
      std::unique_ptr<DevSpace> d(new DevSpace());
      dev_space_build(Psp, *d);
      Psp->dev = d.release();
    }
    S.W = static_cast<DevSpace *>(Wsp->dev);
    S.P = static_cast<DevSpace *>(Psp->dev);
    const int D = S.D = Wsp->mesh->dim;
    const int64_t nu = S.nu = Wsp->nnodes * D, np = S.np = Psp->nnodes, n = S.n = nu + np;
    S.mu = mu;
    for (int64_t i = 0; i < n_ubc; ++i)
      if (ubc_dofs[i] < 0 || ubc_dofs[i] >= nu) return fb_fail(ctx, FB_EINVAL, "fb_stokes_solve: velocity dof out of range");
    for (int64_t i = 0; i < n_pbc; ++i)
      if (pbc_dofs[i] < 0 || pbc_dofs[i] >= np) return fb_fail(ctx, FB_EINVAL, "fb_stokes_solve: pressure dof out of range");

    // operators: mu K (scalar P2 stiffness), M_p (P1 mass)
    S.K.ctx = S.Mp.ctx = ctx;
    S.K.sp = S.W;
    S.K.val.alloc((size_t)S.W->nnz);
    assemble_constant(ctx, *S.W, 0, S.K.val.p);
    vec_axpby(ctx, S.K.val.p, mu, S.K.val.p, 0.0, S.K.val.p, S.W->nnz);
    S.Mp.sp = S.P;
    S.Mp.val.alloc((size_t)S.P->nnz);
    assemble_constant(ctx, *S.P, 1, S.Mp.val.p);

    // Dirichlet data: mask, x_g
    std::vector<int64_t> dofs((size_t)(n_ubc + n_pbc));
    std::vector<double> vals((size_t)(n_ubc + n_pbc));
    for (int64_t i = 0; i < n_ubc; ++i) {
      dofs[i] = ubc_dofs[i];
      vals[i] = ubc_vals[i];
    }
    for (int64_t i = 0; i < n_pbc; ++i) {
      dofs[n_ubc + i] = nu + pbc_dofs[i];
      vals[n_ubc + i] = pbc_vals[i];
    }
    const int64_t nbc = n_ubc + n_pbc;
    DBuf<int64_t> ddofs;
    DBuf<double> dvals;
    ddofs.upload(dofs.data(), dofs.size(), st);
    dvals.upload(vals.data(), vals.size(), st);
    S.mask.alloc((size_t)n);
    mask_build(ctx, S.mask.p, n, ddofs.p, nbc);
    S.dinv_u.alloc((size_t)nu);
    S.dinv_p.alloc((size_t)np);
    jacobi_setup_scalar(ctx, *S.W, S.K.val.p, D, S.mask.p, S.dinv_u.p);
    jacobi_setup_scalar(ctx, *S.P, S.Mp.val.p, 1, S.mask.p + nu, S.dinv_p.p);
    S.nn = Wsp->nnodes;
    // FB_STOKES_AMG_MIN: smallest mesh (P2 nodes) that gets the hierarchy (tests force it on small meshes; a huge value = Jacobi-PCG)
    const int64_t amg_min = getenv("FB_STOKES_AMG_MIN") ? atoll(getenv("FB_STOKES_AMG_MIN")) : 4096;
    if (Wsp->nnodes >= amg_min && !fb_is_distributed(ctx)) {
      const int64_t nn = Wsp->nnodes;
      fb_space_build_pattern(Wsp);
      const int64_t nnz = (int64_t)Wsp->indices.size();
      std::vector<int> rp((size_t)nn + 1);
      for (int64_t i = 0; i <= nn; ++i) rp[(size_t)i] = (int)Wsp->indptr[(size_t)i];
      std::vector<double> k0((size_t)nnz), kv;
      FB_CUDA(cudaStreamSynchronize(st));
      FB_CUDA(cudaMemcpy(k0.data(), S.K.val.p, sizeof(double) * nnz, cudaMemcpyDeviceToHost));
      std::vector<int> bits((size_t)nn, 0);  // constrained components per node (a dof may be listed twice)
      for (int64_t i = 0; i < n_ubc; ++i) bits[(size_t)(ubc_dofs[i] / D)] |= 1 << (int)(ubc_dofs[i] % D);
      for (int c = 0; c < D; ++c) {
        for (int e = 0; e < c && !S.amgK[c]; ++e) {  // same constraint set as an earlier component?
          bool same = true;
          for (int64_t i = 0; i < nn && same; ++i) same = ((bits[(size_t)i] >> c) & 1) == ((bits[(size_t)i] >> e) & 1);
          if (same) S.amgK[c] = S.amgK[e];
        }
        if (S.amgK[c]) continue;
        kv = k0;
        for (int64_t i = 0; i < nn; ++i)
          for (int64_t k = rp[(size_t)i]; k < rp[(size_t)i + 1]; ++k) {
            const int64_t j = Wsp->indices[(size_t)k];
            if (((bits[(size_t)i] >> c) & 1) || ((bits[(size_t)j] >> c) & 1)) kv[(size_t)k] = (i == j) ? 1.0 : 0.0;
          }
        S.amgK[c] = amg_setup(ctx, (int)nn, rp.data(), Wsp->indices.data(), kv.data());
      }
      S.mask_c.alloc((size_t)nn * D);
      S.dinv_c.alloc((size_t)nn);
      S.cin.alloc((size_t)nn);
      S.cout.alloc((size_t)nn);
      for (int c = 0; c < D; ++c)
        FB_LAUNCH(ctx, k_take_comp_mask, vgrid(ctx, nn), 256, 0, nn, D, c, S.mask.p, S.mask_c.p + (size_t)c * nn);
      FB_LAUNCH(ctx, k_take_comp, vgrid(ctx, nn), 256, 0, nn, D, 0, S.dinv_u.p, S.dinv_c.p);
    }

    // right-hand side: b_u = (f, v), b_p = 0; lifted by the Dirichlet data
    DBuf<double> b, xg, x, w, tmp;
    b.alloc((size_t)n);
    xg.alloc((size_t)n);
    x.alloc((size_t)n);
    w.alloc((size_t)n);
    tmp.alloc((size_t)n);
    b.zero(st);
    if (forcing != FB_F_NONE && f) {
      if (forcing == FB_F_LOAD) {
        FB_CUDA(cudaMemcpyAsync(b.p, f, sizeof(double) * nu, cudaMemcpyHostToDevice, st));
      } else {
        fb_mat Mu;
        Mu.ctx = ctx;
        Mu.sp = S.W;
        Mu.val.alloc((size_t)S.W->nnz);
        assemble_constant(ctx, *S.W, 1, Mu.val.p);
        std::vector<double> fh((size_t)nu);
        if (forcing == FB_F_CONSTANT)
          for (int64_t i = 0; i < nu; ++i) fh[i] = f[i % D];
        else
          for (int64_t i = 0; i < nu; ++i) fh[i] = f[i];
        FB_CUDA(cudaMemcpyAsync(tmp.p, fh.data(), sizeof(double) * nu, cudaMemcpyHostToDevice, st));
        spmv(ctx, make_linop(Mu, D, nullptr), tmp.p, b.p);
        FB_CUDA(cudaStreamSynchronize(st));
      }
    }
    xg.zero(st);
    if (nbc > 0) {
      vec_set_at(ctx, xg.p, ddofs.p, dvals.p, nbc);
      // full operator on x_g (not masked): tmp = A x_g
      spmv(ctx, make_linop(S.K, D, nullptr), xg.p, tmp.p);
      stokes_grad(ctx, *S.W, *S.P, xg.p + nu, tmp.p);
      FB_CUDA(cudaMemsetAsync(tmp.p + nu, 0, sizeof(double) * np, st));
      stokes_div(ctx, *S.W, *S.P, xg.p, tmp.p + nu);
      vec_axpy(ctx, b.p, -1.0, tmp.p, n);
      vec_zero_at(ctx, b.p, ddofs.p, nbc);
    }

    FB_CUDA(cudaStreamSynchronize(st));
    const auto t_setup = std::chrono::steady_clock::now();
    // ---- FGMRES(m)
    const int m = 30;
    std::vector<DBuf<double>> V(m + 1), Z(m);
    for (auto &v : V) v.alloc((size_t)n);
    for (auto &z : Z) z.alloc((size_t)n);
    const int g = vgrid(ctx, n);
    const double bnorm = vec_norm2_sync(ctx, b.p, n);
    double g2 = 0.0;
    for (double v : vals) g2 += v * v;
    const double ref = std::sqrt(bnorm * bnorm + g2);  // reference includes the Dirichlet rows of the eliminated system
    x.zero(st);
    int total = 0;
    bool converged = (ref == 0.0);
    double resid = bnorm;
    std::vector<double> H((m + 1) * m), cs(m), sn(m), gv(m + 1), yv(m);
    double *hp = dv->host_pinned;
    while (!converged && total < maxit) {
      // r = b - A x
      if (total == 0) {
        FB_CUDA(cudaMemcpyAsync(w.p, b.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
      } else {
        S.apply(x.p, w.p);
        vec_axpby(ctx, w.p, 1.0, b.p, -1.0, w.p, n);
      }
      vec_dot(ctx, w.p, w.p, n, 40);
      FB_LAUNCH(ctx, k_scale_rsqrt_slot, g, 256, 0, V[0].p, w.p, dv->red + 40, n);
      FB_CUDA(cudaMemcpyAsync(hp, dv->red + 40, sizeof(double), cudaMemcpyDeviceToHost, st));
      FB_CUDA(cudaStreamSynchronize(st));
      const double beta = std::sqrt(hp[0]);
      resid = beta;
      if (beta <= tol * ref) {
        converged = true;
        break;
      }
      std::fill(gv.begin(), gv.end(), 0.0);
      gv[0] = beta;
      int j = 0;
      for (; j < m && total < maxit; ++j, ++total) {
        // inner tolerance: two digits below the current relative residual, at most 1e-2
        const double rtol_in = 1e-2;
        int pst = S.precond(V[j].p, Z[j].p, rtol_in);
        if (pst != FB_OK) return fb_fail(ctx, FB_ENAN, "fb_stokes_solve: preconditioner broke down");
        S.apply(Z[j].p, w.p);
        for (int i = 0; i <= j; ++i) {  // modified Gram-Schmidt, coefficients stay on the device
          vec_dot(ctx, V[i].p, w.p, n, i);
          FB_LAUNCH(ctx, k_axpy_slot, g, 256, 0, w.p, -1.0, dv->red + i, V[i].p, n);
        }
        vec_dot(ctx, w.p, w.p, n, j + 1);
        FB_LAUNCH(ctx, k_scale_rsqrt_slot, g, 256, 0, V[j + 1].p, w.p, dv->red + j + 1, n);
        FB_CUDA(cudaMemcpyAsync(hp, dv->red, sizeof(double) * (j + 2), cudaMemcpyDeviceToHost, st));
        FB_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i <= j; ++i) H[i * m + j] = hp[i];
        H[(j + 1) * m + j] = std::sqrt(hp[j + 1]);
        for (int i = 0; i < j; ++i) {  // previous Givens rotations
          const double t = cs[i] * H[i * m + j] + sn[i] * H[(i + 1) * m + j];
          H[(i + 1) * m + j] = -sn[i] * H[i * m + j] + cs[i] * H[(i + 1) * m + j];
          H[i * m + j] = t;
        }
        const double a = H[j * m + j], bb = H[(j + 1) * m + j];
        const double rr = std::hypot(a, bb);
        if (rr == 0.0 || rr != rr) return fb_fail(ctx, FB_ENAN, "fb_stokes_solve: GMRES breakdown");
        cs[j] = a / rr;
        sn[j] = bb / rr;
        H[j * m + j] = rr;
        H[(j + 1) * m + j] = 0.0;
        gv[j + 1] = -sn[j] * gv[j];
        gv[j] = cs[j] * gv[j];
        resid = std::fabs(gv[j + 1]);
        if (resid <= tol * ref) {
          ++j;
          ++total;
          converged = true;
          break;
        }
      }
      // x += Z y,  H y = g
      for (int i = j - 1; i >= 0; --i) {
        double s = gv[i];
        for (int k = i + 1; k < j; ++k) s -= H[i * m + k] * yv[k];
        yv[i] = s / H[i * m + i];
      }
      for (int i = 0; i < j; ++i) vec_axpy(ctx, x.p, yv[i], Z[i].p, n);
    }
    if (iterations) *iterations = total;
    if (getenv("FB_VERBOSE")) {
      FB_CUDA(cudaStreamSynchronize(st));
      const auto t_end = std::chrono::steady_clock::now();
      fprintf(stderr, "[flow_b200] stokes: %lld dofs, set-up %.3f s (AMG on the velocity block: %s), FGMRES %d iterations in %.3f s, inner PCG iterations: velocity %lld, pressure mass %lld\n",
              (long long)n, std::chrono::duration<double>(t_setup - t_begin).count(), S.amgK[0] ? "yes" : "no", total,
              std::chrono::duration<double>(t_end - t_setup).count(), (long long)S.inner_k, (long long)S.inner_m);
    }
    vec_axpy(ctx, x.p, 1.0, xg.p, n);
    FB_CUDA(cudaMemcpyAsync(u_out, x.p, sizeof(double) * nu, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaMemcpyAsync(p_out, x.p + nu, sizeof(double) * np, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    if (!converged) {
      char buf[160];
      snprintf(buf, sizeof buf, "fb_stokes_solve: GMRES did not converge in %d iterations (|r|/|b| = %.3e)", total,
               resid / (ref > 0 ? ref : 1.0));
      return fb_fail(ctx, FB_ENOCONV_KRYLOV, buf);
    }
  } catch (const fb_cuda_error &e) {
    return fb_fail(ctx, e.status, e.what());
  } catch (const std::exception &e) {
    return fb_fail(ctx, FB_ECUDA, e.what());
  }
  return FB_OK;
}
