"""Shared helpers for the parity tests: same seeded inputs for the CUDA path and the oracle."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from oracle import fem, forms

MESHES = {
    "tri_crossed": lambda: fem.unit_square_mesh(5, 4, "crossed"),
    "tri_leftright": lambda: fem.rectangle_mesh((-1.0, 0.0), (1.0, 0.7), 4, 5, "left/right"),
    "tri_right": lambda: fem.unit_square_mesh(7, 3, "right"),
    "tet": lambda: fem.unit_cube_mesh(3, 2, 4),
    "tet_box": lambda: fem.box_mesh((0.0, -1.0, 0.5), (2.0, 0.0, 1.0), 2, 3, 2),
}


def oracle_mesh(name):
    pts, cells = MESHES[name]()
    return fem.Mesh(pts, cells)


def facade_mesh(om):
    from flow_b200 import dolfin as d

    return d.Mesh(om.points, om.cells)


def rand_state(om, seed=0):
    rng = np.random.default_rng(seed)
    W = fem.Space(om, 2, om.dim)
    P = fem.Space(om, 1, 1)
    return W, P, rng.standard_normal(W.ndofs), rng.standard_normal(W.ndofs), rng.standard_normal(P.nnodes), rng.standard_normal(P.nnodes)


def mat_to_csr(mat_handle, node_space, ncomp_block):
    """Fetch an fb_mat as scipy CSR of the interleaved system (block = ncomp_block)."""
    from flow_b200 import _lib
    from flow_b200._lib import lib

    nrows, nnzb, blk = _lib.i64(), _lib.i64(), C.c_int()
    lib.fb_mat_info(mat_handle, C.byref(nrows), C.byref(nnzb), C.byref(blk))
    b = blk.value
    vals = np.zeros(nnzb.value * b * b)
    _lib.check(lib.fb_mat_values(mat_handle, _lib.as_pd(vals)), None, "fb_mat_values")
    nnz = _lib.i64()
    ip, ix = _lib.pi64(), _lib.pi32()
    lib.fb_space_pattern(node_space.handle, C.byref(nnz), C.byref(ip), C.byref(ix))
    indptr = np.ctypeslib.as_array(ip, shape=(nrows.value + 1,)).copy()
    indices = np.ctypeslib.as_array(ix, shape=(nnz.value,)).copy()
    if b == 1:
        return sp.csr_matrix((vals, indices, indptr), shape=(nrows.value, nrows.value))
    # row-planar block CSR == scalar CSR values; expand the indices
    counts = np.diff(indptr)
    s_indptr = np.zeros(nrows.value * b + 1, dtype=np.int64)
    s_indptr[1:] = np.cumsum(np.repeat(counts * b, b))
    cols = (indices[:, None] * b + np.arange(b)[None, :]).reshape(-1)  # per block row: nb*b entries
    s_indices = np.concatenate([np.tile(cols[indptr[i] * b:indptr[i + 1] * b], b) for i in range(nrows.value)])
    return sp.csr_matrix((vals, s_indices, s_indptr), shape=(nrows.value * b, nrows.value * b))


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
