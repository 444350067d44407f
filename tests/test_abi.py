"""The C-ABI library loads, exports every symbol include/flowb200.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "flowb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from flow_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (fb_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    unbound = [s for s in syms if s not in _lib.SIGNATURES]
    assert not unbound, unbound


def test_library_targets_sm_100a():
    from flow_b200 import _lib

    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH]).decode()
    assert "sm_100a" in out


def test_host_only_context_refuses_compute():
    from flow_b200 import _lib
    from flow_b200._lib import lib

    ctx = _lib.vp()
    assert lib.fb_ctx_create(-1, C.byref(ctx)) == _lib.FB_OK
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    cells = np.array([[0, 1, 3], [0, 2, 3]], dtype=np.int32)
    mesh = _lib.vp()
    assert lib.fb_mesh_create(ctx, 2, 4, _lib.as_pd(pts), 2, _lib.as_pi32(cells), C.byref(mesh)) == _lib.FB_OK
    sp = _lib.vp()
    assert lib.fb_space_create(mesh, 2, 1, C.byref(sp)) == _lib.FB_OK
    mat = _lib.vp()
    assert lib.fb_assemble_mass(sp, C.byref(mat)) == _lib.FB_ENODEVICE
    assert b"no CPU compute path" in lib.fb_last_error(ctx)
    W, P = _lib.vp(), _lib.vp()
    lib.fb_space_create(mesh, 2, 2, C.byref(W))
    lib.fb_space_create(mesh, 1, 1, C.byref(P))
    ns = _lib.vp()
    assert lib.fb_ns_create(W, P, None, C.byref(ns)) == _lib.FB_ENODEVICE
    # invalid input is rejected with FB_EINVAL
    bad = _lib.vp()
    assert lib.fb_space_create(mesh, 3, 1, C.byref(bad)) == _lib.FB_EINVAL
    cells_bad = np.array([[0, 1, 1], [0, 2, 9]], dtype=np.int32)
    assert lib.fb_mesh_create(ctx, 2, 4, _lib.as_pd(pts), 2, _lib.as_pi32(cells_bad), C.byref(bad)) == _lib.FB_EINVAL


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "_cbaseline" not in src and "hostsim" not in src.replace("tests/hostsim", ""), f
