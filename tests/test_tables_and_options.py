"""Generated tables and host-side option handling (CPU tier)."""
import os
import re
import sys
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("dim,name", [(2, "FB_M3_TRI"), (3, "FB_M3_TET")])
def test_p2_triple_product_table(dim, name):
    """flow_b200/csrc/fb_p2_tables.h is what tools/gen_p2_tables.py produces, and sum_v M3[a][b][v] is the P2 mass
    matrix of the oracle (independent quadrature) on the reference simplex."""
    import gen_p2_tables as g

    from oracle import fem, forms

    t = g.table(dim)
    nl, nv = len(t), dim + 1
    text = open(os.path.join(ROOT, "flow_b200", "csrc", "fb_p2_tables.h")).read()
    body = re.search(r"%s\[\d+\] = \{(.*?)\};" % name, text, re.S).group(1)
    vals = np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()])
    ref = np.array([[[float(t[a][b][v]) for v in range(nv)] for b in range(nl)] for a in range(nl)])
    assert vals.size == nl * nl * nv and np.abs(vals - ref.ravel()).max() < 1e-17
    # mass matrix of one reference cell (volume 1/d!)
    pts = np.vstack([np.zeros(dim), np.eye(dim)])
    om = fem.Mesh(pts, np.arange(dim + 1, dtype=np.int32)[None, :])
    W = fem.Space(om, 2, 1)
    M = forms.mass_matrix(W).toarray()
    vol = 1.0 / np.prod(np.arange(1, dim + 1))
    loc = W.cell_nodes[0]
    assert np.abs(M[np.ix_(loc, loc)] / vol - ref.sum(axis=2)).max() < 1e-14
    # exact rational identities: partition of unity in both basis indices
    for v in range(nv):
        assert sum(t[a][b][v] for a in range(nl) for b in range(nl)) == Fraction(1, dim + 1)


def test_set_options_contract():
    from flow_b200 import _lib
    from flow_b200 import navier_stokes as nav
    from flow_b200.navier_stokes import pressure_correction as pc

    nav.reset_options()
    try:
        nav.set_options(pressure_precond="amg", momentum_solver="fgmres", newton_atol=1e-12, warm_start=0)
        assert pc._options == {"pressure_precond": _lib.AMG, "momentum_solver": _lib.GMRES, "newton_atol": 1e-12, "warm_start": 0}
        with pytest.raises(KeyError):
            nav.set_options(no_such_option=1)
        with pytest.raises(KeyError):
            nav.set_options(reserved=1)
        # every option name is a field of the C struct, and the struct matches the header's field order
        header = open(os.path.join(ROOT, "include", "flowb200.h")).read()
        struct = header[header.index("typedef struct fb_ns_opts {"):header.index("} fb_ns_opts;")]
        fields = re.findall(r"^\s*(?:int|double)\s+(\w+)(?:\[\d+\])?;", struct, re.M)
        assert fields == [name for name, _ in _lib.NSOpts._fields_]
        o = _lib.NSOpts()
        assert _lib.lib.fb_ns_opts_default(o) == 0
        assert o.momentum_solver == _lib.GMRES and o.pressure_precond == _lib.AMG and o.newton_atol == 1e-10
        assert o.newton_maxit == 10 and o.jacobian_fp32 == 0 and o.inner_fp32 == 0 and o.momentum_inner_its == 4
        # 0 = degree of the Chebyshev preconditioner chosen from the spectrum estimate; the exact defaults carry no mixed precision
        assert o.chebyshev_degree == 0 and o.inner_chebyshev == 1 and o.semi_implicit == 0 and o.jacobian_reuse == 0
    finally:
        nav.reset_options()
