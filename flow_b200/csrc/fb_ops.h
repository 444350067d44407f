// Host-callable launchers of the CUDA kernels in fb_kernels.cu / fb_krylov.cu.
#pragma once
#include <vector>

#include "fb_device.cuh"

// Tile-CSR format of a node pattern (fb_tile.cu): rows grouped into tiles of spatially close nodes, entries stored
// as (fp64 value, 16-bit index into the tile's column union), the union gathered into shared memory once per tile.
struct TileFormat {
  int64_t nrows = 0, ntiles = 0, nent = 0, union_total = 0;
  int cfg = 0;           // kernel configuration the tiles were sized for (fb_tile.cu TileCfg)
  DBuf<int> desc;        // ntiles * 8: row0, nr, e0, ne, u0, nu, g0, ng
  DBuf<int> rowid;       // tile-order row -> canonical row
  DBuf<int> gptr;        // per tile ng + 1 entry offsets of its 8-row groups, relative to e0
  DBuf<int> ucol;        // column unions (canonical node ids)
  DBuf<uint16_t> lidx;   // nent
  DBuf<int> src;         // nent: CSR slot of the entry (-1: padding) -- used to (re)pack values
};

// Device copy of a node space: dof map, node-level CSR pattern, element->slot scatter map.
struct DevSpace {
  fb_ctx *ctx = nullptr;
  fb_space *host = nullptr;  // the space this is the device copy of
  TileFormat *tile = nullptr;  // built on first use (mat_enable_tile)
  ~DevSpace() { delete tile; }
  int dim = 0, nl = 0, degree = 0;
  int64_t nnodes = 0, nc = 0, nnz = 0, nbf = 0;
  DBuf<double> xyz;         // vertex coordinates (mesh.nv * dim)
  DBuf<int> cell_nodes;     // nc * nl, in the space's node numbering
  DBuf<int> cells;          // nc * (dim+1): mesh vertex connectivity (geometry)
  DBuf<int> rowptr;         // nnodes + 1
  DBuf<int> col;            // nnz
  DBuf<int> diag;           // nnodes: slot of the diagonal entry
  DBuf<int> smap;           // nc * nl * nl: CSR slot of (node_a, node_b)
  DBuf<int> bf_cell, bf_local;
  // deterministic two-pass assembly (fb_kernels.cu): per matrix block the list of element blocks that sum into it
  // (gptr: nnz + 1, gsrc: nc * nl * nl, ascending) and the element buffer of the last assembly
  DBuf<int> gptr, gsrc;
  DBuf<double> ebuf;
  // distributed: rows [0, n_owned) are computed here, [n_owned, nnodes) are ghosts (SpMV inputs only)
  int64_t n_owned = 0;
  std::vector<int> halo_ranks;
  std::vector<int64_t> halo_send_ptr, halo_recv_ptr;
  DBuf<int> halo_send_nodes;
  DBuf<double> halo_buf;
};

struct fb_mat {
  fb_ctx *ctx = nullptr;
  DevSpace *sp = nullptr;  // pattern owner (not owned)
  int block = 1;           // 1: scalar node matrix; D: D x D blocks stored row-planar
  DBuf<double> val;        // nnz * block * block
  DBuf<double> tval;       // block == 1 only: the same values in tile order (empty: CSR kernels only)
  bool owned_space = false;
};

// Linear operator view handed to SpMV / Krylov
struct LinOp {
  int block = 1;   // 1 or D
  int ncomp = 1;   // interleaved components sharing a scalar matrix (block == 1)
  int64_t nrows = 0;   // node rows computed here (owned)
  int64_t nlocal = 0;  // node columns stored here (owned + ghosts)
  DevSpace *halo = nullptr;  // non-null: refresh the ghosts of x before multiplying
  const int *rowptr = nullptr;
  const int *col = nullptr;
  const double *val = nullptr;
  const float *val32 = nullptr;   // block > 1 only: fp32 copy of val; when set the product streams this one
  const uint8_t *mask = nullptr;  // per dof: 1 -> identity row (Dirichlet); may be null
  const TileFormat *tile = nullptr;  // block == 1: tile format of the pattern and the values packed for it
  const double *tval = nullptr;      // (both set: spmv() runs the tile kernel)
  int64_t ndofs() const { return nrows * (block > 1 ? block : ncomp); }
  int64_t nlocal_dofs() const { return nlocal * (block > 1 ? block : ncomp); }
  int dofs_per_node() const { return block > 1 ? block : ncomp; }
};

LinOp make_linop(const fb_mat &m, int ncomp, const uint8_t *mask);

// ---- multi-GPU (fb_comm.cu): no-ops on a single rank
bool fb_is_distributed(const fb_ctx *ctx);
void fb_allreduce_slots(fb_ctx *ctx, int slot0, int count);
double fb_allreduce_host_sum(fb_ctx *ctx, double v);
void fb_broadcast_device(fb_ctx *ctx, void *p, size_t bytes, int root);
// IPC-shared global vector assembled from the ranks' owned parts by remote stores (peer-memory transport only)
struct fb_peer_vec;
fb_peer_vec *fb_peer_vec_create(fb_ctx *ctx, int64_t n_global);  // collective; null without peer memory
void fb_peer_vec_destroy(fb_peer_vec *v);
const double *fb_peer_vec_gather(fb_ctx *ctx, fb_peer_vec *v, const double *owned, const int *l2g, int64_t n_owned);
void halo_exchange(fb_ctx *ctx, DevSpace &sp, double *x, int ncomp);

// ---- setup
void dev_space_build(fb_space *s, DevSpace &d);
// kind: 0 P1/P2 stiffness, 1 mass
void assemble_constant(fb_ctx *ctx, DevSpace &sp, int kind, double *val);
void assemble_lumped(fb_ctx *ctx, DevSpace &sp, double *diag);

// ---- tile format (fb_tile.cu)
bool tile_enabled();  // FB_TILE=0 disables the format globally (CSR kernels everywhere)
void tile_format_build(fb_space *s, TileFormat &tf, cudaStream_t st);
void tile_pack(fb_ctx *ctx, const TileFormat &tf, const double *val, double *tval);  // tval[k] = val[src[k]]
void tile_spmm(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode, const double *w, int slot,
               const int *flag);
void tile_cheb_step(fb_ctx *ctx, const LinOp &A, const double *d, const double *rin, bool rin_canonical, double *rout,
                    const double *zin, double *zout, double *dout, const double *dinv_t, double cdd, double cr, bool last);
void tile_to_tile_order(fb_ctx *ctx, const TileFormat &tf, int ncomp, const double *x, double *xt);  // xt = x in the format's row order
// build the space's tile format if needed and pack m's values for it; mat_repack after m.val changed
void mat_enable_tile(fb_ctx *ctx, fb_mat &m);
void mat_repack(fb_ctx *ctx, fb_mat &m);

// ---- SpMV: y = A x; dot_mode 0 none, 1: red[slot] = w.y, 2: red[slot] = w.y and red[slot+1] = y.y
void spmv(fb_ctx *ctx, const LinOp &A, const double *x, double *y, int dot_mode = 0, const double *w = nullptr,
          int slot = 0, const int *flag = nullptr);

// ---- vector kernels
void vec_fill(fb_ctx *ctx, double *x, double a, int64_t n);
void vec_axpy(fb_ctx *ctx, double *y, double a, const double *x, int64_t n);          // y += a x
void vec_axpby(fb_ctx *ctx, double *z, double a, const double *x, double b, const double *y, int64_t n);  // z = a x + b y
void vec_to_float(fb_ctx *ctx, float *dst, const double *src, int64_t n);
void vec_dot(fb_ctx *ctx, const double *x, const double *y, int64_t n, int slot);     // red[slot] = x.y
double vec_norm2_sync(fb_ctx *ctx, const double *x, int64_t n);                       // host-synchronous
void mask_build(fb_ctx *ctx, uint8_t *mask, int64_t ndofs, const int64_t *dofs, int64_t nbc);
void vec_set_at(fb_ctx *ctx, double *x, const int64_t *dofs, const double *vals, int64_t nbc);   // x[dofs] = vals
void vec_zero_at(fb_ctx *ctx, double *x, const int64_t *dofs, int64_t nbc);
void vec_copy_at(fb_ctx *ctx, double *dst, const double *src, const int64_t *dofs, int64_t nbc);
void lift_identity_rows(fb_ctx *ctx, const LinOp &A, double *b, const int64_t *dofs, int64_t nbc, double *xg, double *tmp);
void bc_residual(fb_ctx *ctx, double *F, const double *x, const int64_t *dofs, const double *vals, int64_t nbc);

// ---- Dirichlet on matrices
void bc_rows_identity_blocked(fb_ctx *ctx, const DevSpace &sp, int D, double *val, const int64_t *dofs, int64_t nbc);
void bc_symmetric_scalar(fb_ctx *ctx, const DevSpace &sp, double *val, const uint8_t *mask);
void bc_rows_identity_scalar(fb_ctx *ctx, const DevSpace &sp, double *val, const uint8_t *mask);

// ---- preconditioners: inverse diagonal (block == 1: per node, replicated over comps by the caller's kernels)
void jacobi_setup_scalar(fb_ctx *ctx, const DevSpace &sp, const double *val, int ncomp, const uint8_t *mask, double *dinv);
// D x D inverse diagonal blocks (block_mode 0: point Jacobi stored as diagonal blocks, 1: full block inverse)
void jacobi_setup_blocked(fb_ctx *ctx, const DevSpace &sp, int D, const double *val, int block_mode, double *binv);

// ---- Navier-Stokes element kernels (D = dim of W)
struct MomentumArgs {
  double dt, rho, mu, theta;
  const double *ui, *u0, *p0;
  const int *pcn;  // cell -> P1 dofs of the pressure space
  bool skip_facets = false;     // every boundary dof is constrained: the exterior-facet terms only touch overwritten rows
  const double *adv = nullptr;  // semi-implicit linearisation: advecting velocity of the new-state convection (null: ui)
};
void assemble_momentum_F(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *F);  // zero + both parts
void assemble_momentum_F_old_state(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *F);  // += u0 part
void assemble_momentum_F_new_state(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *F);  // += ui part
void assemble_momentum_J(fb_ctx *ctx, const DevSpace &W, const MomentumArgs &a, double *Jval);
void set_deterministic_assembly(bool on);  // two-pass, fixed summation order (process-wide switch; FB_J_TWO_PASS overrides)
void assemble_pressure_rhs(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, double dt, double rho, double mu,
                           int rotational, const double *ui, const double *p0, double *b);
// adds -dt/rho (grad phi, v) to b (which already holds M ui)
void assemble_correction_grad(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, double dt, double rho, double mu,
                              int rotational, const double *ui, const double *p1, const double *p0, double *b);
// driver-side quantities: load vector of |u_h| against the scalar P2 basis, max norms
void assemble_magnitude_load(fb_ctx *ctx, const DevSpace &W, const double *u, int nq, const double *qlam, const double *qw, double *b);
void vec_max_norms(fb_ctx *ctx, const double *x, int64_t n, const double *u, int64_t nnodes, int D, double *out2_host);
// heat operator A (heat.py:54-58): -(kappa/rho_cp) grad u.grad v - (conv.grad u) v on V's pattern
void assemble_heat(fb_ctx *ctx, const DevSpace &V, const DevSpace *W, const double *conv, double kdiff, double *val);
// SUPG additions (heat.py:60-86): returns non-zero if tau exceeded 1e3 somewhere (the reference throws)
int assemble_heat_supg(fb_ctx *ctx, const DevSpace &V, const DevSpace &W, const double *conv, double kappa, double rho_cp,
                       double source, double *Aval, double *Mval, double *bvec);
int supg_tau_values(fb_ctx *ctx, const DevSpace &W, const double *conv, double eps, int p, const int *order_dev, double *out_mesh_order);
// B x and B^T y for the divergence block of stokes.py:40-42 (matrix-free)
void stokes_div(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, const double *u, double *out_p);
void stokes_grad(fb_ctx *ctx, const DevSpace &W, const DevSpace &P, const double *p, double *out_u);

// ---- Krylov (fb_krylov.cu).  All vectors are device pointers; return FB_OK / FB_ENOCONV_KRYLOV / FB_ENAN.
struct KrylovWork {
  DBuf<double> v[10];
  void ensure(int count, int64_t n) {
    for (int i = 0; i < count; ++i) v[i].alloc((size_t)n);
  }
};
// smoothed-aggregation AMG (fb_amg.cu): hierarchy of the n x n leading (owned) block of a host CSR matrix
struct fb_amg;
fb_amg *amg_setup(fb_ctx *ctx, int n, const int *rowptr, const int *col, const double *val);
void amg_apply(fb_amg *amg, const double *r, double *z);  // z = V(1,1)-cycle applied to r (device vectors)
// Partitioned runs: the hierarchy is the one of the GLOBAL matrix, replicated on every rank.  amg_apply then
// gathers the ranks' owned residuals into the global vector (fb_peer_vec_gather), runs the cycle redundantly and
// returns the owned part: the preconditioner -- hence the iteration count -- is the single-GPU one.
void amg_set_replicated(fb_amg *amg, fb_peer_vec *gather, const int *l2g_host, int n_owned);
void amg_destroy(fb_amg *amg);
int amg_num_levels(const fb_amg *amg);
double amg_complexity(const fb_amg *amg);
int amg_level_size(const fb_amg *amg, int l);

// PCG, stopping test ||M^-1 r|| <= rtol * ||M^-1 b|| (PETSc default); M^-1 = Jacobi (dinv) or, if amg != null, the
// AMG V-cycle (then dinv is ignored).  warm_start: x holds an initial guess on entry (the test is unchanged).
int krylov_pcg(fb_ctx *ctx, const LinOp &A, const double *dinv, const double *b, double *x, double rtol, double ref_extra2,
               int maxit, int check_every, KrylovWork &w, int *iters, fb_amg *amg = nullptr, bool warm_start = false);
// right-preconditioned BiCGStab with D x D block inverse (A.block > 1) or diagonal dinv (A.block == 1);
// stops when ||r||_2 <= atol
int krylov_bicgstab(fb_ctx *ctx, const LinOp &A, const double *minv, const double *b, double *x, double atol, int maxit,
                    int check_every, KrylovWork &w, int *iters);

// Flexible GMRES(m), right preconditioning with a variable preconditioner z = pc.apply(self, v): stops when the
// residual estimate is <= atol (absolute, like krylov_bicgstab).  inner_iters: preconditioner iterations spent.
struct FgmresPrecond {
  void *self = nullptr;
  int (*apply)(void *self, const double *v, double *z, int *inner_its) = nullptr;
};
struct FgmresWork {
  std::vector<DBuf<double>> V, Z;
  DBuf<double> w;
  void ensure(int m, int64_t n) {
    if ((int)V.size() < m + 1) {
      V = std::vector<DBuf<double>>(m + 1);
      Z = std::vector<DBuf<double>>(m);
    }
    for (auto &v : V) v.alloc((size_t)n);
    for (auto &z : Z) z.alloc((size_t)n);
    w.alloc((size_t)n);
  }
};
int krylov_fgmres(fb_ctx *ctx, const LinOp &A, const FgmresPrecond &pc, const double *b, double *x, double atol, int maxit,
                  int m, FgmresWork &fw, int *iters, int *inner_iters, bool warm_start = false);

// Chebyshev polynomial preconditioner z = p_m(D^-1 A) D^-1 v for an SPD operator in tile format (fb_krylov.cu):
// m - 1 products, each ONE kernel (the vector updates of the iteration live in the product's epilogue), no inner
// products, no host synchronisation.  [lmin, lmax] bound the spectrum of D^-1 A (cheb_estimate_spectrum).
struct ChebWork {
  DBuf<double> r, d0, d1;   // r: tile order (row operand only); d0, d1: canonical (gather sources)
  DBuf<double> zt, dinv_t;  // tile order: running sum of the directions, inverse diagonal
  const double *dinv_src = nullptr;  // dinv_t is the permuted copy of this array ...
  bool dinv_valid = false;           // ... and is rebuilt when the owner of the operator clears this flag
  double lmin = 0.0, lmax = 0.0;
  // partitioned runs: true = polynomial of the rank's owned x owned block of the operator (couplings to ghost
  // columns dropped, no halo exchange inside the preconditioner -- block Jacobi over the ranks); the outer flexible
  // GMRES iteration and its operator stay global, so only the iteration count can change, not the solution
  bool local = false;
};
void cheb_estimate_spectrum(fb_ctx *ctx, const LinOp &A, const double *dinv, int lanczos_steps, double *lmin, double *lmax);
void cheb_apply(fb_ctx *ctx, const LinOp &A, const double *dinv, double lmin, double lmax, int degree, const double *v, double *z,
                ChebWork &w);

// fp32 inner solver of the momentum preconditioner (fb_inner32.cu)
struct Inner32 {
  int64_t nrows = 0;  // node rows
  int ncomp = 3;
  int its = 4;
  const int *rowptr = nullptr, *col = nullptr;
  const float *val = nullptr, *dinv = nullptr;
  const uint8_t *mask = nullptr;
  DBuf<float> r, z, p, Ap, x;
  void ensure(int64_t n);
};
void inner32_apply(fb_ctx *ctx, Inner32 &in, const double *v, double *z_out);
