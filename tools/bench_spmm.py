#!/usr/bin/env python
"""SpMM micro-benchmark on the benchmark mesh: scalar P2 mass matrix x 3 interleaved components, row-wise CSR kernel
(k_spmm_u) vs tile-CSR kernel (k_tile_spmm), CUDA events inside the library (fb_mat_bench_spmv).
  python tools/bench_spmm.py [n=74] [reps=50]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flow_b200 import _lib  # noqa: E402
from flow_b200 import dolfin as d  # noqa: E402
from flow_b200._lib import lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 74
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    mesh = d.UnitCubeMesh(n, n, n)
    V = d.FunctionSpace(mesh, "CG", 2)
    h = _lib.vp()
    _lib.check(lib.fb_assemble_mass(V.handle(), C.byref(h)), mesh.ctx, "mass")
    out = {"n": n, "rows": V.dim()}
    ms, byt = C.c_double(), C.c_double()
    for nc in (3, 1):
        for name, fmt in (("csr", _lib.FORMAT_CSR), ("tile", _lib.FORMAT_TILE)):
            t0 = time.perf_counter()
            _lib.check(lib.fb_mat_set_format(h, fmt), mesh.ctx, "set_format")
            t_fmt = time.perf_counter() - t0
            _lib.check(lib.fb_mat_bench_spmv(h, nc, reps, C.byref(ms), C.byref(byt)), mesh.ctx, "bench")
            rec = {"ms": ms.value, "algorithmic_GBs": byt.value / ms.value / 1e6, "set_format_s": t_fmt}
            if fmt == _lib.FORMAT_TILE:
                f, nt, ne, nu = C.c_int(), _lib.i64(), _lib.i64(), _lib.i64()
                lib.fb_mat_format_info(h, C.byref(f), C.byref(nt), C.byref(ne), C.byref(nu))
                rec.update(tiles=nt.value, entries=ne.value, union_columns=nu.value, union_per_row=nu.value / V.dim(),
                           format_bytes=ne.value * 10 + nu.value * 4 + V.dim() * 8)
            out["nc%d_%s" % (nc, name)] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
