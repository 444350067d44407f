// Host-side mesh topology, dof maps and sparsity patterns (no CUDA in this unit).
//
// Replaces, for the hot path only, DOLFIN's Mesh / FunctionSpace / DofMap /
// SparsityPatternBuilder [EXT, SURVEY.md 2.2 E1, E2, E4].  The numbering is the
// canonical one defined in oracle/fem.py so that GPU and oracle arrays can be
// compared bit for bit.
#include <algorithm>
#include <array>
#include <cstring>
#include <cstdlib>
#include <numeric>

#include "fb_internal.h"

namespace {

const int TRI_EDGES[3][2] = {{1, 2}, {0, 2}, {0, 1}};
const int TET_EDGES[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};

struct FacetKey {
  int32_t v[3];
  int32_t cell, local;
  bool operator<(const FacetKey &o) const {
    if (v[0] != o.v[0]) return v[0] < o.v[0];
    if (v[1] != o.v[1]) return v[1] < o.v[1];
    return v[2] < o.v[2];
  }
  bool same(const FacetKey &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};

}  // namespace

int fb_fail(fb_ctx *ctx, int status, const std::string &msg) {
  if (ctx) ctx->err = msg;
  return status;
}

extern "C" {

int fb_mesh_create(fb_ctx *ctx, int gdim, int64_t nverts, const double *xyz, int64_t ncells, const int32_t *cells,
                   fb_mesh **out) {
  if (!ctx || !out || !xyz || !cells) return FB_EINVAL;
  if (gdim != 2 && gdim != 3) return fb_fail(ctx, FB_EINVAL, "fb_mesh_create: gdim must be 2 or 3");
  if (nverts <= 0 || ncells <= 0 || nverts > INT32_MAX) return fb_fail(ctx, FB_EINVAL, "fb_mesh_create: bad sizes");
  const int nvc = gdim + 1;
  fb_mesh *m = new fb_mesh();
  m->ctx = ctx;
  m->dim = gdim;
  m->nv = nverts;
  m->nc = ncells;
  m->xyz.assign(xyz, xyz + nverts * gdim);
  m->cells.assign(cells, cells + ncells * nvc);
  bool bad = false;
#pragma omp parallel for reduction(|| : bad)
  for (int64_t c = 0; c < ncells; ++c) {
    int32_t *v = &m->cells[c * nvc];
    std::sort(v, v + nvc);
    if (v[0] < 0 || v[nvc - 1] >= nverts) bad = true;
    for (int i = 1; i < nvc; ++i)
      if (v[i] == v[i - 1]) bad = true;
  }
  if (bad) {
    delete m;
    return fb_fail(ctx, FB_EINVAL, "fb_mesh_create: cell with out-of-range or repeated vertex");
  }

  // ---- edges: unique (min,max) pairs in lexicographic order
  const int nle = fb_num_local_edges(gdim);
  const int(*LE)[2] = gdim == 2 ? TRI_EDGES : TET_EDGES;
  std::vector<uint64_t> keys((size_t)ncells * nle);
#pragma omp parallel for
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t *v = &m->cells[c * nvc];
    for (int e = 0; e < nle; ++e)
      keys[c * nle + e] = ((uint64_t)(uint32_t)v[LE[e][0]] << 32) | (uint32_t)v[LE[e][1]];
  }
  std::vector<uint64_t> uniq(keys);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  m->ne = (int64_t)uniq.size();
  m->edges.resize(m->ne * 2);
#pragma omp parallel for
  for (int64_t e = 0; e < m->ne; ++e) {
    m->edges[2 * e] = (int32_t)(uniq[e] >> 32);
    m->edges[2 * e + 1] = (int32_t)(uniq[e] & 0xffffffffu);
  }
  m->cell_edges.resize((size_t)ncells * nle);
#pragma omp parallel for
  for (int64_t i = 0; i < ncells * nle; ++i)
    m->cell_edges[i] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), keys[i]) - uniq.begin());

  // ---- boundary facets: facet f is opposite local vertex f; boundary iff it occurs once
  std::vector<FacetKey> fk((size_t)ncells * nvc);
#pragma omp parallel for
  for (int64_t c = 0; c < ncells; ++c) {
    const int32_t *v = &m->cells[c * nvc];
    for (int f = 0; f < nvc; ++f) {
      FacetKey k;
      k.v[2] = -1;
      int n = 0;
      for (int i = 0; i < nvc; ++i)
        if (i != f) k.v[n++] = v[i];
      k.cell = (int32_t)c;
      k.local = f;
      fk[c * nvc + f] = k;
    }
  }
  std::sort(fk.begin(), fk.end());
  std::vector<std::pair<int32_t, int32_t>> bf;
  for (size_t i = 0; i < fk.size();) {
    size_t j = i + 1;
    while (j < fk.size() && fk[j].same(fk[i])) ++j;
    if (j - i == 1) bf.emplace_back(fk[i].cell, fk[i].local);
    i = j;
  }
  std::sort(bf.begin(), bf.end());  // (cell, local) ascending == numpy nonzero order
  m->bf_cell.resize(bf.size());
  m->bf_local.resize(bf.size());
  m->bvert.assign(nverts, 0);
  m->bedge.assign(m->ne, 0);
  for (size_t i = 0; i < bf.size(); ++i) {
    const int32_t c = bf[i].first, f = bf[i].second;
    m->bf_cell[i] = c;
    m->bf_local[i] = f;
    for (int k = 0; k < nvc; ++k)
      if (k != f) m->bvert[m->cells[(int64_t)c * nvc + k]] = 1;
    for (int e = 0; e < nle; ++e)
      if (LE[e][0] != f && LE[e][1] != f) m->bedge[m->cell_edges[(int64_t)c * nle + e]] = 1;
  }
  *out = m;
  return FB_OK;
}

int fb_mesh_info(fb_mesh *m, int64_t *nverts, int64_t *ncells, int64_t *nedges, int64_t *nbfacets) {
  if (!m) return FB_EINVAL;
  if (nverts) *nverts = m->nv;
  if (ncells) *ncells = m->nc;
  if (nedges) *nedges = m->ne;
  if (nbfacets) *nbfacets = (int64_t)m->bf_cell.size();
  return FB_OK;
}

int fb_mesh_cells(fb_mesh *m, const int32_t **cells) {
  if (!m || !cells) return FB_EINVAL;
  *cells = m->cells.data();
  return FB_OK;
}

int fb_mesh_edges(fb_mesh *m, const int32_t **edges) {
  if (!m || !edges) return FB_EINVAL;
  *edges = m->edges.data();
  return FB_OK;
}

int fb_mesh_boundary_facets(fb_mesh *m, const int32_t **cell, const int32_t **local_facet) {
  if (!m) return FB_EINVAL;
  if (cell) *cell = m->bf_cell.data();
  if (local_facet) *local_facet = m->bf_local.data();
  return FB_OK;
}

// Replace the boundary-facet list (and the boundary vertex / edge flags derived from it).  A rank-local
// sub-mesh of a partitioned domain must not treat its cut faces as domain boundary: the caller passes
// the facets of the GLOBAL boundary that lie in local cells.  Call before creating spaces.
int fb_mesh_set_boundary_facets(fb_mesh *m, int64_t n, const int32_t *cell, const int32_t *local_facet) {
  if (!m || n < 0 || (n > 0 && (!cell || !local_facet))) return FB_EINVAL;
  const int nvc = m->dim + 1, nle = fb_num_local_edges(m->dim);
  const int(*LE)[2] = m->dim == 2 ? TRI_EDGES : TET_EDGES;
  for (int64_t i = 0; i < n; ++i)
    if (cell[i] < 0 || cell[i] >= m->nc || local_facet[i] < 0 || local_facet[i] >= nvc)
      return fb_fail(m->ctx, FB_EINVAL, "fb_mesh_set_boundary_facets: facet out of range");
  m->bf_cell.assign(cell, cell + n);
  m->bf_local.assign(local_facet, local_facet + n);
  m->bvert.assign(m->nv, 0);
  m->bedge.assign(m->ne, 0);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t c = cell[i], f = local_facet[i];
    for (int k = 0; k < nvc; ++k)
      if (k != f) m->bvert[m->cells[(int64_t)c * nvc + k]] = 1;
    for (int e = 0; e < nle; ++e)
      if (LE[e][0] != f && LE[e][1] != f) m->bedge[m->cell_edges[(int64_t)c * nle + e]] = 1;
  }
  return FB_OK;
}

int fb_space_create(fb_mesh *m, int degree, int ncomp, fb_space **out) {
  if (!m || !out) return FB_EINVAL;
  if (degree != 1 && degree != 2) return fb_fail(m->ctx, FB_EINVAL, "fb_space_create: degree must be 1 or 2");
  if (ncomp < 1 || ncomp > 3) return fb_fail(m->ctx, FB_EINVAL, "fb_space_create: ncomp must be 1..3");
  const int d = m->dim, nvc = d + 1, nle = fb_num_local_edges(d);
  fb_space *s = new fb_space();
  s->mesh = m;
  s->degree = degree;
  s->ncomp = ncomp;
  if (degree == 1) {
    s->nl = nvc;
    s->nnodes = m->nv;
    s->cell_nodes = m->cells;
    s->coords = m->xyz;
    s->bnode = m->bvert;
    s->n_owned = s->nnodes;
  } else {
    s->nl = nvc + nle;
    s->nnodes = m->nv + m->ne;
    if (s->nnodes > INT32_MAX) {
      delete s;
      return fb_fail(m->ctx, FB_EINVAL, "fb_space_create: more than 2^31 nodes");
    }
    s->cell_nodes.resize((size_t)m->nc * s->nl);
#pragma omp parallel for
    for (int64_t c = 0; c < m->nc; ++c) {
      int32_t *dst = &s->cell_nodes[c * s->nl];
      for (int i = 0; i < nvc; ++i) dst[i] = m->cells[c * nvc + i];
      for (int e = 0; e < nle; ++e) dst[nvc + e] = (int32_t)(m->nv + m->cell_edges[c * nle + e]);
    }
    s->coords.resize((size_t)s->nnodes * d);
    std::memcpy(s->coords.data(), m->xyz.data(), sizeof(double) * m->nv * d);
#pragma omp parallel for
    for (int64_t e = 0; e < m->ne; ++e)
      for (int k = 0; k < d; ++k)
        s->coords[(m->nv + e) * d + k] = 0.5 * (m->xyz[(int64_t)m->edges[2 * e] * d + k] + m->xyz[(int64_t)m->edges[2 * e + 1] * d + k]);
    s->bnode.resize(s->nnodes);
    std::memcpy(s->bnode.data(), m->bvert.data(), m->nv);
    std::memcpy(s->bnode.data() + m->nv, m->bedge.data(), m->ne);
    s->n_owned = s->nnodes;
  }
  *out = s;
  return FB_OK;
}

// Space with a caller-chosen node numbering: perm[canonical node] = node id used by this space
// (a bijection on [0, nnodes)).  Distributed runs number the owned nodes first and the ghost
// nodes after them (grouped by owner); n_owned = number of owned nodes.
int fb_space_create_numbered(fb_mesh *m, int degree, int ncomp, const int32_t *perm, int64_t n_owned, fb_space **out) {
  fb_space *s = nullptr;
  int st = fb_space_create(m, degree, ncomp, &s);
  if (st != FB_OK) return st;
  const int64_t nn = s->nnodes;
  if (n_owned < 0 || n_owned > nn) {
    delete s;
    return fb_fail(m->ctx, FB_EINVAL, "fb_space_create_numbered: n_owned out of range");
  }
  s->n_owned = n_owned;
  if (perm) {
    std::vector<uint8_t> seen(nn, 0);
    for (int64_t i = 0; i < nn; ++i) {
      if (perm[i] < 0 || perm[i] >= nn || seen[perm[i]]) {
        delete s;
        return fb_fail(m->ctx, FB_EINVAL, "fb_space_create_numbered: perm is not a permutation");
      }
      seen[perm[i]] = 1;
    }
    const int d = m->dim;
    for (auto &v : s->cell_nodes) v = perm[v];
    std::vector<double> coords(s->coords.size());
    std::vector<uint8_t> bnode(nn);
    for (int64_t i = 0; i < nn; ++i) {
      for (int k = 0; k < d; ++k) coords[(int64_t)perm[i] * d + k] = s->coords[i * d + k];
      bnode[perm[i]] = s->bnode[i];
    }
    s->coords.swap(coords);
    s->bnode.swap(bnode);
  }
  *out = s;
  return FB_OK;
}

// Halo plan of a distributed space.  Neighbour k sends us the ghost nodes
// [n_owned + recv_ptr[k], n_owned + recv_ptr[k+1]) and receives our owned nodes
// send_nodes[send_ptr[k] .. send_ptr[k+1]) in that order.
int fb_space_set_halo(fb_space *s, int nneigh, const int32_t *ranks, const int64_t *send_ptr, const int32_t *send_nodes,
                      const int64_t *recv_ptr) {
  if (!s || nneigh < 0) return FB_EINVAL;
  if (s->dev) return fb_fail(s->mesh->ctx, FB_EINVAL, "fb_space_set_halo: must be called before the space is used on the device");
  if (nneigh > 0 && (!ranks || !send_ptr || !recv_ptr || (send_ptr[nneigh] > 0 && !send_nodes))) return FB_EINVAL;
  s->halo_ranks.assign(ranks, ranks + nneigh);
  s->halo_send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
  s->halo_recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
  s->halo_send_nodes.assign(send_nodes, send_nodes + (nneigh ? send_ptr[nneigh] : 0));
  if (nneigh && s->n_owned + recv_ptr[nneigh] != s->nnodes)
    return fb_fail(s->mesh->ctx, FB_EINVAL, "fb_space_set_halo: ghost segments do not cover [n_owned, nnodes)");
  for (int32_t v : s->halo_send_nodes)
    if (v < 0 || v >= s->n_owned) return fb_fail(s->mesh->ctx, FB_EINVAL, "fb_space_set_halo: send node is not owned");
  return FB_OK;
}

int fb_space_info(fb_space *s, int64_t *nnodes, int64_t *ndofs, int *nodes_per_cell) {
  if (!s) return FB_EINVAL;
  if (nnodes) *nnodes = s->nnodes;
  if (ndofs) *ndofs = s->nnodes * s->ncomp;
  if (nodes_per_cell) *nodes_per_cell = s->nl;
  return FB_OK;
}

int fb_space_dofmap(fb_space *s, const int32_t **cell_nodes) {
  if (!s || !cell_nodes) return FB_EINVAL;
  *cell_nodes = s->cell_nodes.data();
  return FB_OK;
}

int fb_space_node_coords(fb_space *s, const double **xyz) {
  if (!s || !xyz) return FB_EINVAL;
  *xyz = s->coords.data();
  return FB_OK;
}

int fb_space_boundary_nodes(fb_space *s, const uint8_t **flags) {
  if (!s || !flags) return FB_EINVAL;
  *flags = s->bnode.data();
  return FB_OK;
}

int fb_space_pattern(fb_space *s, int64_t *nnz, const int64_t **indptr, const int32_t **indices) {
  if (!s) return FB_EINVAL;
  int st = fb_space_build_pattern(s);
  if (st != FB_OK) return st;
  if (nnz) *nnz = (int64_t)s->indices.size();
  if (indptr) *indptr = s->indptr.data();
  if (indices) *indices = s->indices.data();
  return FB_OK;
}

}  // extern "C"

// Node-level CSR pattern: row i couples to every node of every cell containing node i.
int fb_space_build_pattern(fb_space *s) {
  if (!s->indptr.empty()) return FB_OK;
  const fb_mesh *m = s->mesh;
  const int nl = s->nl;
  const int64_t nn = s->nnodes, nc = m->nc;
  // node -> cells adjacency (CSR)
  std::vector<int64_t> nptr(nn + 1, 0);
  for (int64_t i = 0; i < nc * nl; ++i) nptr[s->cell_nodes[i] + 1]++;
  for (int64_t i = 0; i < nn; ++i) nptr[i + 1] += nptr[i];
  std::vector<int32_t> ncell(nptr[nn]);
  {
    std::vector<int64_t> fill(nptr.begin(), nptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int a = 0; a < nl; ++a) ncell[fill[s->cell_nodes[c * nl + a]]++] = (int32_t)c;
  }
  std::vector<int64_t> cnt(nn + 1, 0);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 4096)
    for (int64_t i = 0; i < nn; ++i) {
      tmp.clear();
      for (int64_t k = nptr[i]; k < nptr[i + 1]; ++k) {
        const int32_t *cn = &s->cell_nodes[(int64_t)ncell[k] * nl];
        tmp.insert(tmp.end(), cn, cn + nl);
      }
      std::sort(tmp.begin(), tmp.end());
      cnt[i + 1] = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
    }
  }
  for (int64_t i = 0; i < nn; ++i) cnt[i + 1] += cnt[i];
  s->indices.resize(cnt[nn]);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 4096)
    for (int64_t i = 0; i < nn; ++i) {
      tmp.clear();
      for (int64_t k = nptr[i]; k < nptr[i + 1]; ++k) {
        const int32_t *cn = &s->cell_nodes[(int64_t)ncell[k] * nl];
        tmp.insert(tmp.end(), cn, cn + nl);
      }
      std::sort(tmp.begin(), tmp.end());
      auto end = std::unique(tmp.begin(), tmp.end());
      std::copy(tmp.begin(), end, s->indices.begin() + cnt[i]);
    }
  }
  s->indptr.swap(cnt);
  return FB_OK;
}


// ---- device storage order of the cells ------------------------------------------------------------------------------
static inline uint64_t fb_spread_bits(uint64_t v, int dim) {
  uint64_t out = 0;
  const int bits = dim == 3 ? 21 : 31;
  for (int b = 0; b < bits; ++b) out |= ((v >> b) & 1ull) << (dim * b);
  return out;
}

const std::vector<int32_t> &fb_mesh_cell_order(fb_mesh *m) {
  if ((int64_t)m->cell_order.size() == m->nc) return m->cell_order;
  m->cell_order.resize((size_t)m->nc);
  std::iota(m->cell_order.begin(), m->cell_order.end(), 0);
  const char *e = getenv("FB_CELL_ORDER");
  if (e && atoi(e) == 0) return m->cell_order;
  const int d = m->dim, nv = d + 1;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t v = 0; v < m->nv; ++v)
    for (int k = 0; k < d; ++k) {
      lo[k] = std::min(lo[k], m->xyz[v * d + k]);
      hi[k] = std::max(hi[k], m->xyz[v * d + k]);
    }
  double ext = 0.0;
  for (int k = 0; k < d; ++k) ext = std::max(ext, hi[k] - lo[k]);
  if (!(ext > 0.0)) ext = 1.0;
  const double scale = (double)((1ull << (d == 3 ? 21 : 31)) - 1) / ext;
  std::vector<uint64_t> key((size_t)m->nc);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < m->nc; ++c) {
    uint64_t code = 0;
    for (int k = 0; k < d; ++k) {
      double x = 0.0;
      for (int v = 0; v < nv; ++v) x += m->xyz[(int64_t)m->cells[c * nv + v] * d + k];
      code |= fb_spread_bits((uint64_t)((x / nv - lo[k]) * scale), d) << k;
    }
    key[c] = code;
  }
  std::stable_sort(m->cell_order.begin(), m->cell_order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
  return m->cell_order;
}
