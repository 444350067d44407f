/* flowb200.h -- C ABI of libflowb200.so, the B200-native engine behind flow's
 * time-stepping hot path.
 *
 * The reference (nschloe/flow) has no FFI of its own: its boundary to native
 * code is DOLFIN's Python API.  Each entry point below
 * names the reference call it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - every function returns an int status (FB_OK == 0); nothing throws or aborts
 *     across the ABI; fb_last_error(ctx) holds a message for the last failure.
 *   - the caller owns every input/output array (plain host buffers unless
 *     FB_DEVICE_PTRS is passed); the library owns handles and all device memory.
 *   - numbering is the canonical one (oracle/fem.py): vertex nodes in vertex
 *     order, then edge nodes in lexicographic (min,max) order; vector dofs are
 *     interleaved, dof = ncomp*node + comp.
 *   - a context created with device = -1 is host-only: mesh / space / pattern
 *     queries work, every compute entry point returns FB_ENODEVICE.  There is no
 *     CPU compute path in this library.
 */
#ifndef FLOWB200_H
#define FLOWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_ctx fb_ctx;
typedef struct fb_mesh fb_mesh;
typedef struct fb_space fb_space;
typedef struct fb_mat fb_mat;
typedef struct fb_ns fb_ns;
typedef struct fb_heat fb_heat;

enum fb_status {
  FB_OK = 0,
  FB_EINVAL = 1,         /* asserts at pressure_correction.py:488-489, stokes.py:23 */
  FB_ENOCONV_NEWTON = 2, /* error_on_nonconvergence, pressure_correction.py:236 */
  FB_ENOCONV_KRYLOV = 3, /* pressure_correction.py:337,424,462; stokes.py:139 */
  FB_ENAN = 4,
  FB_ECUDA = 5,
  FB_ENCCL = 6,
  FB_ENOMEM = 7,
  FB_ENODEVICE = 8
};

enum fb_scheme { FB_FORWARD_EULER = 0, FB_BACKWARD_EULER = 1, FB_CRANK_NICOLSON = 2 };
enum fb_flags { FB_DEVICE_PTRS = 1, FB_ROTATIONAL = 2, FB_CHORIN = 4 };
enum fb_forcing { FB_F_NONE = 0, FB_F_CONSTANT = 1, FB_F_NODAL = 2, FB_F_LOAD = 3 };
enum fb_krylov { FB_BICGSTAB = 0, FB_GMRES = 1, FB_CG = 2 };
enum fb_precond { FB_JACOBI = 0, FB_BLOCK_JACOBI = 1, FB_CHEBYSHEV = 2 /* see fb_ns_opts.inner_chebyshev */, FB_AMG = 3 };

/* ---- context ---------------------------------------------------------- */
int fb_version(void);
int fb_ctx_create(int device, fb_ctx **out);
int fb_ctx_destroy(fb_ctx *ctx);
const char *fb_last_error(fb_ctx *ctx);
const char *fb_status_string(int status);
/* number of kernels this context has launched so far (bench.py gpu_launches) */
int fb_ctx_launch_count(fb_ctx *ctx, int64_t *count);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel is launched on) */
int fb_ctx_timer_start(fb_ctx *ctx);
int fb_ctx_timer_stop(fb_ctx *ctx, double *ms);

/* page-locked host buffers for the caller's state vectors (faster H2D/D2H); plain malloc'd
 * or numpy memory is accepted everywhere as well */
int fb_host_alloc(fb_ctx *ctx, int64_t bytes, void **out);
int fb_host_free(fb_ctx *ctx, void *ptr);

/* ---- mesh: replaces dolfin.Mesh / UnitSquareMesh / RectangleMesh / BoxMesh
 * (tests/test_navier_stokes.py:82,144,176; tests/test_sealed_box.py:53).
 * cells: ncells*(gdim+1) vertex indices; they are sorted per cell (UFC order). */
int fb_mesh_create(fb_ctx *ctx, int gdim, int64_t nverts, const double *xyz, int64_t ncells,
                   const int32_t *cells, fb_mesh **out);
int fb_mesh_destroy(fb_mesh *mesh);
int fb_mesh_info(fb_mesh *mesh, int64_t *nverts, int64_t *ncells, int64_t *nedges, int64_t *nbfacets);
int fb_mesh_cells(fb_mesh *mesh, const int32_t **cells);
int fb_mesh_edges(fb_mesh *mesh, const int32_t **edges);
int fb_mesh_boundary_facets(fb_mesh *mesh, const int32_t **cell, const int32_t **local_facet);

/* ---- function space: replaces FunctionSpace / VectorFunctionSpace and their dof maps
 * (tests/test_navier_stokes.py:282-283; tests/test_sealed_box.py:59-61). degree in {1,2}. */
int fb_space_create(fb_mesh *mesh, int degree, int ncomp, fb_space **out);
int fb_space_destroy(fb_space *space);
int fb_space_info(fb_space *space, int64_t *nnodes, int64_t *ndofs, int *nodes_per_cell);
int fb_space_dofmap(fb_space *space, const int32_t **cell_nodes);
int fb_space_node_coords(fb_space *space, const double **xyz);
int fb_space_boundary_nodes(fb_space *space, const uint8_t **flags);
/* node-level CSR sparsity pattern (columns ascending) -- the pattern of assemble() */
int fb_space_pattern(fb_space *space, int64_t *nnz, const int64_t **indptr, const int32_t **indices);

/* ---- distributed runs (one process per GPU).  The caller partitions the mesh
 * (flow_b200/parallel.py: recursive coordinate bisection, owner-computes rows, one ghost-cell
 * layer), creates the rank-local mesh and passes the node numbering "owned first, ghosts grouped
 * by owner" plus the halo plan.  Replaces DOLFIN's implicit MPI partitioning + PETSc's VecScatter /
 * MPI_Allreduce [EXT]; the reference tree itself has no MPI-aware code. */
int fb_mesh_set_boundary_facets(fb_mesh *mesh, int64_t n, const int32_t *cell, const int32_t *local_facet);
int fb_space_create_numbered(fb_mesh *mesh, int degree, int ncomp, const int32_t *perm, int64_t n_owned,
                             fb_space **out);
int fb_space_set_halo(fb_space *space, int nneigh, const int32_t *ranks, const int64_t *send_ptr,
                      const int32_t *send_nodes, const int64_t *recv_ptr);
/* 128-byte NCCL unique id (create on rank 0, broadcast by the launcher, e.g. torch.distributed) */
int fb_comm_unique_id(void *id128);
int fb_comm_init(fb_ctx *ctx, int rank, int nranks, const void *id128);
int fb_comm_destroy(fb_ctx *ctx);
/* Peer-memory transport over NVLink (replaces NCCL on the hot path: halo exchange and dot-product all-reduce
 * become our own kernels storing into IPC-mapped windows of the other ranks).  Each rank creates its window
 * and gets a 64-byte CUDA IPC handle; the launcher all-gathers the handles (rank order) and every rank opens
 * them.  Optional: without it the same operations go through NCCL. */
int fb_comm_window_create(fb_ctx *ctx, int64_t staging_bytes_per_peer, void *handle64);
int fb_comm_window_open(fb_ctx *ctx, const void *handles);
int fb_comm_window_disable(fb_ctx *ctx);
int fb_comm_uses_peer_memory(fb_ctx *ctx);
/* partitioned runs: halo exchanges / all-reduces enqueued so far on this context, and the latency of one of each
 * (collective micro-benchmark) -- bench.py's communication share */
int fb_ctx_comm_counts(fb_ctx *ctx, int64_t *halo_exchanges, int64_t *allreduces);
int fb_space_bench_comm(fb_space *space, int ncomp, int reps, double *halo_us, double *allreduce_us);
/* refresh the ghost entries of a host vector (ncomp interleaved components) -- test/debug helper */
int fb_space_halo_exchange(fb_space *space, int ncomp, double *x);

/* ---- assembled operators: replaces dolfin.assemble(a) for the constant forms
 * u*v*dx (pressure_correction.py:442) and dot(grad p, grad q)*dx (:317), plus
 * the vertex-quadrature mass of heat.py:39-45. */
int fb_assemble_mass(fb_space *space, fb_mat **out);
int fb_assemble_stiffness(fb_space *space, fb_mat **out);
int fb_assemble_lumped_mass(fb_space *space, double *diag_out);
int fb_mat_destroy(fb_mat *mat);
/* block = 1: scalar node matrix; block = d: d x d blocks (momentum Jacobian).
 * values_out (caller-allocated, nnz*block*block doubles) receives scalar-CSR
 * values of the interleaved system, row-major within each block row. */
int fb_mat_info(fb_mat *mat, int64_t *nrows, int64_t *nnz_blocks, int *block);
int fb_mat_values(fb_mat *mat, double *values_out);
/* y = A x for `ncomp` interleaved components sharing a scalar matrix (block==1),
 * or the blocked product (block==d, ncomp ignored). Host buffers. */
int fb_mat_spmv(fb_mat *mat, int ncomp, const double *x, double *y);
/* Jacobi-PCG on a scalar SPD matrix applied to ncomp interleaved components with
 * symmetric Dirichlet elimination: replaces solve(a==L, bcs, 'cg', symmetric=True)
 * (pressure_correction.py:451-464) and project() in the tests. */
int fb_mat_solve_cg(fb_mat *mat, int ncomp, const double *b, double *x, int64_t nbc, const int64_t *bc_dofs,
                    const double *bc_vals, double rtol, int maxit, int *iterations);
/* Storage format of a scalar node matrix on the device.  FB_FORMAT_CSR: row-wise CSR kernels.  FB_FORMAT_TILE: tile-CSR
 * (csrc/fb_tile.cu): rows grouped into tiles of spatially close nodes, the union of a tile's columns staged in shared
 * memory, entries streamed by TMA bulk copies as (fp64 value, 16-bit local column).  Same results up to summation
 * order; the Navier-Stokes engine selects it by itself for its P2 operators on meshes of >= 16384 nodes. */
enum fb_format { FB_FORMAT_CSR = 0, FB_FORMAT_TILE = 1 };
int fb_mat_set_format(fb_mat *mat, int format);
/* host-only self check of the tile format of a space's pattern (no device needed): the format must reproduce the CSR
 * pattern exactly.  stats[9]: tiles, entries incl. padding, sum of union sizes, max rows / entries / union per tile, and
 * the row / entry / union caps of the kernel configuration */
int fb_space_tile_check(fb_space *space, int64_t *stats);
int fb_mat_format_info(fb_mat *mat, int *format, int64_t *ntiles, int64_t *entries, int64_t *union_columns);
/* SpMV micro-benchmark on resident data: runs `reps` products, returns avg ms and algorithmic bytes */
int fb_mat_bench_spmv(fb_mat *mat, int ncomp, int reps, double *ms_avg, double *bytes);

/* ---- Navier-Stokes pressure-correction step: replaces _step
 * (pressure_correction.py:468-518) = _compute_tentative_velocity (:147-255) +
 * _compute_pressure (:258-433) + _compute_velocity_correction (:436-465). */
typedef struct fb_ns_opts {
  int momentum_solver;   /* FB_BICGSTAB: block-Jacobi BiCGStab on the Jacobian; FB_GMRES: flexible GMRES whose
                            preconditioner is momentum_inner_its CG iterations on the constant scalar operator
                            M + theta dt nu K per component (the Jacobian is streamed ~5x less often) */
  int momentum_precond;  /* preconditioner of the FB_BICGSTAB solver: FB_JACOBI or FB_BLOCK_JACOBI (default) */
  int pressure_precond;  /* FB_AMG (default; smoothed aggregation, V(1,1), dense pseudo-inverse on the coarsest level;
                            replaces hypre BoomerAMG of pressure_correction.py:331,:414-419; systems of fewer than
                            4096 unknowns fall back to Jacobi) or FB_JACOBI */
  int newton_maxit;      /* 10, pressure_correction.py:232 */
  double newton_atol;    /* 1e-10, pressure_correction.py:499 */
  double momentum_rtol;  /* 1e-6: every Newton update is solved to max(newton_overshoot * newton_atol, momentum_rtol * |F|) */
  int momentum_maxit;    /* 1000 (commented-out block, pressure_correction.py:249) */
  int pressure_maxit;    /* Krylov cap for the Poisson solve; default 20000 (Jacobi needs more than AMG's 1000) */
  int correction_maxit;  /* default 1000 */
  int gmres_restart;     /* restart length of the FB_GMRES solver: default 30 (PETSc default), at most 20 vectors are kept */
  int check_every;       /* Krylov iterations enqueued between host convergence checks */
  int chebyshev_degree;  /* degree of the Chebyshev preconditioner (inner_chebyshev); 0 (default): chosen from the estimated
                            condition number of D^-1 S: 4 up to kappa = 50, then round(1.2 sqrt(kappa)), at most 12 */
  int jacobian_reuse;    /* 0 (default): the reference's Newton iteration -- Jacobian of the current iterate at every
                            iteration, first iterate with |F|_2 < newton_atol accepted, updates solved to the tolerance
                            above, so that the iterates are those of the reference's Newton + LU (pressure_correction.py:
                            224-254).  1: chord variant -- keep the step's first Jacobian for later iterations while the
                            residual contracts by > 10x per iteration and iterate to newton_overshoot * newton_atol: a
                            cheaper path to the ROOT of F1 (it does not reproduce the reference's last iterate when that
                            one happens to sit just below newton_atol) */
  int adaptive_forcing;  /* chord variant only.  1: the first linear solve of a step stops at the nonlinear remainder
                            observed at the previous step */
  int jacobian_across_steps; /* chord variant only.  1: the chord Jacobian also survives from one step to the next while dt,
                            rho, mu, the scheme and the constrained dofs are unchanged and the first update of the previous
                            step contracted |F| by > 100x */
  int warm_start;        /* 1 (default): the Poisson / correction CG iterations start from p0 / ui instead of 0; the
                            stopping test (relative to the preconditioned norm of b) is the reference's */
  int jacobian_fp32;     /* 0 (default).  1: the chord Jacobian used INSIDE the Krylov solves is stored in fp32 (half the
                            SpMV bytes); residuals, vectors, dots and the Newton test |F| < atol stay fp64, so the
                            accepted solution satisfies the same fp64 criterion (mixed-precision inexact Newton) */
  int extrapolate_guess; /* 0 (default): Newton starts from u0 (:220).  1: from u0 + dt/dt_prev (u0 - u0_prev) when that guess
                            has a residual below twice what the previous step started from (smooth time loops; no gain
                            on the impulsively started cavity of the benchmark) */
  int momentum_inner_its; /* CG iterations per preconditioner application of the FB_GMRES momentum solver (default 4) */
  int inner_fp32;        /* 0 (default).  1: the CG iterations of the FB_GMRES preconditioner run in fp32 (operator,
                            vectors, products; reductions fp64).  The flexible outer iteration, its residual test and all
                            results stay fp64: a looser preconditioner costs outer iterations, not accuracy.  1 GPU only. */
  double newton_overshoot; /* 1e-3.  The reference's updates are exact (LU) and its accepted iterate sits anywhere between
                            newton_atol and 1e-17; |F|_2 is not mesh-normalised, so on fine meshes two iterates that both
                            pass newton_atol differ visibly (3e-8 relative at 1.3e5 dofs, 1e-7 at 1e7).  Default mode: floor
                            of the linear tolerances = newton_overshoot * newton_atol.  Chord variant: the loop also runs
                            until |F|_2 < newton_overshoot * newton_atol (an iterate below newton_atol is still accepted
                            when further updates stagnate) */
  int inner_chebyshev;   /* 1 (default): on meshes whose P2 operators run from the tile format, the preconditioner of the
                            FB_GMRES momentum solver is a Chebyshev polynomial of degree chebyshev_degree in the Jacobi-scaled
                            S = M + theta dt nu K (spectrum estimated by 12 Lanczos steps whenever S changes): degree - 1
                            products with S, each ONE kernel (vector updates in the product's epilogue), no inner products,
                            no host synchronisation.  0: momentum_inner_its CG iterations on S */
  int semi_implicit;     /* 0 (default): the reference's fully implicit convection ((grad ui) ui, v) - ((grad v) ui, ui).
                            1: semi-implicit linearisation ((grad ui) u0, v) - ((grad v) u0, ui), the (u^k . grad) u^{k+1}
                            treatment the reference's notes recommend (pressure_correction.py:96-101, :204-219) but do not
                            implement: the tentative-velocity system becomes linear in ui (one assembly per step, the
                            second update -- if the first solve's tolerance leaves one -- reuses the matrix) and keeps the skew-symmetric form; an O(dt) different discretisation, NOT the
                            reference's numbers (parity: oracle variant of the same form) */
  int inner_local;       /* 0 (default).  1: in partitioned runs the Chebyshev preconditioner is the polynomial of each rank's
                            owned x owned block of S (no halo exchange inside the preconditioner); the flexible outer
                            iteration keeps the global operator.  Measured at N = 2, n = 74: 33.6 instead of 29.8 outer
                            iterations, 101 instead of 96 ms per step -- the saved exchanges do not pay for the weaker
                            preconditioner; experimental */
  int deterministic_assembly; /* 0 (default): the Jacobian is scatter-added with fp64 atomics (values reproducible to rounding,
                            not bit for bit).  1: two passes -- element blocks stored cell by cell, then every matrix block
                            sums its contributions in a fixed order (ascending cell): bit-reproducible Jacobian, no atomics,
                            +17.5 GB of scratch and ~1.3x the assembly time at 10 M dofs */
  double momentum_amg_kappa; /* 0 (default: never).  > 0: when the estimated condition number of the Jacobi-scaled
                            S = M + theta dt nu K exceeds it (diffusion-dominated steps, BASELINE.json config 2), the
                            preconditioner of the FB_GMRES momentum solver is one smoothed-aggregation AMG V-cycle on S per
                            velocity component (single GPU, >= 4096 nodes).  Measured on config 2 (n = 333): 90 instead of 103
                            outer iterations per step but 47 instead of 31 ms -- what S (x) I lacks there is the viscous
                            coupling of the components, not a better inverse of S; off by default */
  double momentum_rtol_loose; /* 1e-3 (default).  Upper bound of the relative tolerance of Newton updates that are predicted NOT to
                            be the last one (quadratic model |F_next| ~ C |F|^2, C from the previous step): only the last
                            update is part of the accepted iterate; an earlier one is solved just tightly enough that the
                            final iterate stays within 1e-9 of the reference's (tighter when the reference's final residual
                            is predicted close to newton_atol, see fb_api.cu).  A wrong prediction is caught: if the residual
                            after a loosely solved update is below 30 newton_atol, the same linear system is solved on to the
                            tight tolerance before the acceptance test is read.  <= momentum_rtol: every update tight */
} fb_ns_opts;

typedef struct fb_ns_stats {
  int newton_its;
  int momentum_its; /* total Krylov iterations over all Newton steps */
  int pressure_its;
  int correction_its;
  double newton_residual;
  double ms_tentative, ms_pressure, ms_correction, ms_total; /* CUDA-event times */
  double ms_assembly_J, ms_assembly_F, ms_momentum_solve;
  int64_t launches;
  double reserved[8]; /* [0..4]: |F| after k Newton updates; [5]: inner CG iterations of the FB_GMRES momentum
                         solver; [6]: 1 if the extrapolated start was used; [7]: Jacobian assemblies in this step */
} fb_ns_stats;

int fb_ns_opts_default(fb_ns_opts *opts);
/* W: vector P2 space (ncomp == gdim), P: scalar P1 space on the same mesh. */
int fb_ns_create(fb_space *W, fb_space *P, const fb_ns_opts *opts, fb_ns **out);
int fb_ns_destroy(fb_ns *ns);
/* Partitioned runs with the peer-memory transport: replace the per-rank (additive Schwarz) hierarchy by the
 * hierarchy of the GLOBAL P1 stiffness matrix, replicated on every rank; l2g[i] = global vertex of the i-th owned
 * local pressure dof.  The preconditioner, hence the CG iteration count, is then the single-GPU one for any
 * number of ranks.  Collective.  No-op when Jacobi is the pressure preconditioner or NCCL is the transport. */
int fb_ns_set_pressure_amg_global(fb_ns *ns, fb_space *Pglobal, fb_mat *Aglobal, const int64_t *l2g, int64_t n_owned);
/* AMG hierarchy of the pressure operator: number of levels, operator complexity, rows per level */
int fb_ns_amg_info(fb_ns *ns, int *levels, double *complexity, int *sizes, int max_sizes);
/* One step.  scheme: fb_scheme; flags: FB_ROTATIONAL | FB_CHORIN | FB_DEVICE_PTRS.
 * forcing: fb_forcing; f0/f1 = f at t_n / t_{n+1}: gdim doubles (CONSTANT), ndofs(W)
 * doubles of nodal values (NODAL) or of the load vector int f.v dx (LOAD); NULL == 0.
 * u_bc / p_bc: Dirichlet dofs + values (DirichletBC already evaluated at dof coordinates).
 * tol: the reference's `tol` argument (Krylov rtol of the Poisson and correction solves). */
int fb_ns_step(fb_ns *ns, double dt, double rho, double mu, int scheme, int flags, const double *u0, const double *p0,
               int forcing, const double *f0, const double *f1, int64_t n_ubc, const int64_t *ubc_dofs,
               const double *ubc_vals, int64_t n_pbc, const int64_t *pbc_dofs, const double *pbc_vals, double tol,
               double *u1, double *p1, fb_ns_stats *stats);
/* Driver-side quantity of the reference's time loops, on the device: the L2 projection of the velocity magnitude onto
 * the scalar P2 space and its max norm,
 *     unorm = project(sqrt(ux**2 + uy**2), FunctionSpace(mesh, 'Lagrange', 2), quadrature_degree 4); norm(unorm.vector(), 'linf')
 * (tests/test_karman_vortex_street.py:262-268 -- the CFL-like step-size control; tests/test_sealed_box.py:134-141).
 * The integrand is not polynomial, so the quadrature rule is part of the definition: nq barycentric points
 * lam[nq][gdim+1] and weights w (sum 1).  u: velocity dofs (host, or device with FB_DEVICE_PTRS); rtol: mass solve.
 * unorm_out (nnodes(P2) doubles) may be NULL; nodal_max = max over the nodes of |u(x_i)| (no projection).
 * Partitioned runs: the maxima are over the rank's owned nodes (reduce them with the launcher's max all-reduce). */
int fb_ns_velocity_magnitude(fb_ns *ns, const double *u, int flags, int nq, const double *lam, const double *w, double rtol,
                             double *unorm_out, double *linf, double *nodal_max);
/* parity/debug: F1(ui) (and J if want_J) of pressure_correction.py:169-202 without BCs.
 * theta = 0, 1, 0.5; load = time-weighted load vector or NULL. */
int fb_ns_residual(fb_ns *ns, double dt, double rho, double mu, double theta, const double *ui, const double *u0,
                   const double *p0, const double *load, double *F_out, int want_J);
/* which: 0 = P1 stiffness, 1 = scalar P2 mass, 2 = last momentum Jacobian (borrowed handles) */
int fb_ns_matrix(fb_ns *ns, int which, fb_mat **out);
/* parity/debug: right-hand sides L2 (pressure_correction.py:318-323) and L3 (:448-449) */
int fb_ns_pressure_rhs(fb_ns *ns, double dt, double rho, double mu, int rotational, const double *ui, const double *p0,
                       double *b_out);
int fb_ns_correction_rhs(fb_ns *ns, double dt, double rho, double mu, int rotational, const double *ui,
                         const double *p1, const double *p0, double *b_out);

/* ---- heat: replaces flow.heat.Heat (heat.py:20-122).  V scalar P1/P2, W vector P2 (conv) or NULL */
int fb_heat_create(fb_space *V, fb_space *W, const double *conv, double kappa, double rho, double cp,
                   const double *source_load, fb_heat **out);
/* with SUPG stabilisation (heat.py:60-86 + stabilization.py:13-152; triangles only like the reference);
 * source_value: the constant source entering the SUPG residual term */
int fb_heat_create_supg(fb_space *V, fb_space *W, const double *conv, double kappa, double rho, double cp,
                        const double *source_load, int supg, double source_value, fb_heat **out);
int fb_heat_supg_mass(fb_heat *heat, double *values_out);
/* flow.stabilization.supg (stabilization.py:13-152): tau at the 3 vertices of every cell (ncells*3 doubles, mesh cell order),
 * evaluated by the device routine the SUPG heat kernel uses; W: vector P2 space of the convection field (triangles) */
int fb_supg_tau(fb_space *W, const double *conv, double epsilon, int p, double *tau_out); /* parity getter: SUPG part of M on V's pattern */
int fb_heat_destroy(fb_heat *heat);
/* alpha*M*u + beta*(A*u + b)   (heat.py:92-101) */
int fb_heat_eval(fb_heat *heat, double alpha, double beta, const double *u, double *out);
/* solve (alpha*M + beta*A) x = b with DirichletBC.apply rows (heat.py:103-122); b is modified like the reference does */
int fb_heat_solve(fb_heat *heat, double alpha, double beta, double *b, int64_t nbc, const int64_t *bc_dofs,
                  const double *bc_vals, double rtol, int maxit, double *x, int *iterations);
int fb_heat_matrix(fb_heat *heat, int which, fb_mat **out); /* 0 = A */

/* ---- Stokes: replaces flow.stokes.solve (stokes.py:13-148); mixed dofs are passed split (u, p) */
int fb_stokes_solve(fb_space *W, fb_space *P, double mu, int forcing, const double *f, int64_t n_ubc,
                    const int64_t *ubc_dofs, const double *ubc_vals, int64_t n_pbc, const int64_t *pbc_dofs,
                    const double *pbc_vals, double tol, int maxit, double *u, double *p, int *iterations);

#ifdef __cplusplus
}
#endif
#endif /* FLOWB200_H */
