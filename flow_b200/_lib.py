"""ctypes binding of libflowb200.so (the C ABI declared in include/flowb200.h).

There is deliberately no fallback: if the shared library is missing the import
fails loudly, and on a machine without a CUDA device every compute entry point
returns FB_ENODEVICE (raised as RuntimeError).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libflowb200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "flow_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C flow_b200/csrc`. There is no CPU fallback." % LIB_PATH
    )

lib = C.CDLL(LIB_PATH)

FB_OK, FB_EINVAL, FB_ENOCONV_NEWTON, FB_ENOCONV_KRYLOV, FB_ENAN, FB_ECUDA, FB_ENCCL, FB_ENOMEM, FB_ENODEVICE = range(9)
FORWARD_EULER, BACKWARD_EULER, CRANK_NICOLSON = 0, 1, 2
DEVICE_PTRS, ROTATIONAL, CHORIN = 1, 2, 4
F_NONE, F_CONSTANT, F_NODAL, F_LOAD = 0, 1, 2, 3
BICGSTAB, GMRES, CG = 0, 1, 2
JACOBI, BLOCK_JACOBI, CHEBYSHEV, AMG = 0, 1, 2, 3
FORMAT_CSR, FORMAT_TILE = 0, 1

vp = C.c_void_p
i64 = C.c_int64
dbl = C.c_double
pd = C.POINTER(C.c_double)
pi32 = C.POINTER(C.c_int32)
pi64 = C.POINTER(C.c_int64)
pu8 = C.POINTER(C.c_uint8)


class NSOpts(C.Structure):
    _fields_ = [
        ("momentum_solver", C.c_int),
        ("momentum_precond", C.c_int),
        ("pressure_precond", C.c_int),
        ("newton_maxit", C.c_int),
        ("newton_atol", C.c_double),
        ("momentum_rtol", C.c_double),
        ("momentum_maxit", C.c_int),
        ("pressure_maxit", C.c_int),
        ("correction_maxit", C.c_int),
        ("gmres_restart", C.c_int),
        ("check_every", C.c_int),
        ("chebyshev_degree", C.c_int),
        ("jacobian_reuse", C.c_int),
        ("adaptive_forcing", C.c_int),
        ("jacobian_across_steps", C.c_int),
        ("warm_start", C.c_int),
        ("jacobian_fp32", C.c_int),
        ("extrapolate_guess", C.c_int),
        ("momentum_inner_its", C.c_int),
        ("inner_fp32", C.c_int),
        ("newton_overshoot", C.c_double),
        ("inner_chebyshev", C.c_int),
        ("semi_implicit", C.c_int),
        ("inner_local", C.c_int),
        ("deterministic_assembly", C.c_int),
        ("momentum_amg_kappa", C.c_double),
        ("momentum_rtol_loose", C.c_double),
    ]


class NSStats(C.Structure):
    _fields_ = [
        ("newton_its", C.c_int),
        ("momentum_its", C.c_int),
        ("pressure_its", C.c_int),
        ("correction_its", C.c_int),
        ("newton_residual", C.c_double),
        ("ms_tentative", C.c_double),
        ("ms_pressure", C.c_double),
        ("ms_correction", C.c_double),
        ("ms_total", C.c_double),
        ("ms_assembly_J", C.c_double),
        ("ms_assembly_F", C.c_double),
        ("ms_momentum_solve", C.c_double),
        ("launches", C.c_int64),
        ("reserved", C.c_double * 8),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["newton_residuals"] = [self.reserved[k] for k in range(min(5, self.newton_its + 1))]
        d["momentum_inner_its"] = int(self.reserved[5])
        d["extrapolated_start"] = int(self.reserved[6])
        d["jacobian_assemblies"] = int(self.reserved[7])
        return d


# every exported symbol of include/flowb200.h with its signature
SIGNATURES = {
    "fb_version": (C.c_int, []),
    "fb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "fb_ctx_destroy": (C.c_int, [vp]),
    "fb_last_error": (C.c_char_p, [vp]),
    "fb_status_string": (C.c_char_p, [C.c_int]),
    "fb_ctx_launch_count": (C.c_int, [vp, pi64]),
    "fb_ctx_comm_counts": (C.c_int, [vp, pi64, pi64]),
    "fb_space_bench_comm": (C.c_int, [vp, C.c_int, C.c_int, pd, pd]),
    "fb_ctx_timer_start": (C.c_int, [vp]),
    "fb_ctx_timer_stop": (C.c_int, [vp, pd]),
    "fb_host_alloc": (C.c_int, [vp, i64, C.POINTER(vp)]),
    "fb_host_free": (C.c_int, [vp, vp]),
    "fb_mesh_create": (C.c_int, [vp, C.c_int, i64, pd, i64, pi32, C.POINTER(vp)]),
    "fb_mesh_destroy": (C.c_int, [vp]),
    "fb_mesh_info": (C.c_int, [vp, pi64, pi64, pi64, pi64]),
    "fb_mesh_cells": (C.c_int, [vp, C.POINTER(pi32)]),
    "fb_mesh_edges": (C.c_int, [vp, C.POINTER(pi32)]),
    "fb_mesh_boundary_facets": (C.c_int, [vp, C.POINTER(pi32), C.POINTER(pi32)]),
    "fb_space_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "fb_space_destroy": (C.c_int, [vp]),
    "fb_space_info": (C.c_int, [vp, pi64, pi64, C.POINTER(C.c_int)]),
    "fb_space_dofmap": (C.c_int, [vp, C.POINTER(pi32)]),
    "fb_space_node_coords": (C.c_int, [vp, C.POINTER(pd)]),
    "fb_space_boundary_nodes": (C.c_int, [vp, C.POINTER(pu8)]),
    "fb_space_pattern": (C.c_int, [vp, pi64, C.POINTER(pi64), C.POINTER(pi32)]),
    "fb_mesh_set_boundary_facets": (C.c_int, [vp, i64, pi32, pi32]),
    "fb_space_create_numbered": (C.c_int, [vp, C.c_int, C.c_int, pi32, i64, C.POINTER(vp)]),
    "fb_space_set_halo": (C.c_int, [vp, C.c_int, pi32, pi64, pi32, pi64]),
    "fb_comm_unique_id": (C.c_int, [vp]),
    "fb_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "fb_comm_destroy": (C.c_int, [vp]),
    "fb_comm_window_create": (C.c_int, [vp, i64, vp]),
    "fb_comm_window_open": (C.c_int, [vp, vp]),
    "fb_comm_window_disable": (C.c_int, [vp]),
    "fb_comm_uses_peer_memory": (C.c_int, [vp]),
    "fb_space_halo_exchange": (C.c_int, [vp, C.c_int, pd]),
    "fb_assemble_mass": (C.c_int, [vp, C.POINTER(vp)]),
    "fb_assemble_stiffness": (C.c_int, [vp, C.POINTER(vp)]),
    "fb_assemble_lumped_mass": (C.c_int, [vp, pd]),
    "fb_mat_destroy": (C.c_int, [vp]),
    "fb_mat_info": (C.c_int, [vp, pi64, pi64, C.POINTER(C.c_int)]),
    "fb_mat_values": (C.c_int, [vp, pd]),
    "fb_mat_spmv": (C.c_int, [vp, C.c_int, pd, pd]),
    "fb_mat_solve_cg": (C.c_int, [vp, C.c_int, pd, pd, i64, pi64, pd, dbl, C.c_int, C.POINTER(C.c_int)]),
    "fb_mat_set_format": (C.c_int, [vp, C.c_int]),
    "fb_space_tile_check": (C.c_int, [vp, pi64]),
    "fb_mat_format_info": (C.c_int, [vp, C.POINTER(C.c_int), pi64, pi64, pi64]),
    "fb_mat_bench_spmv": (C.c_int, [vp, C.c_int, C.c_int, pd, pd]),
    "fb_ns_opts_default": (C.c_int, [C.POINTER(NSOpts)]),
    "fb_ns_create": (C.c_int, [vp, vp, C.POINTER(NSOpts), C.POINTER(vp)]),
    "fb_ns_destroy": (C.c_int, [vp]),
    "fb_ns_set_pressure_amg_global": (C.c_int, [vp, vp, vp, pi64, i64]),
    "fb_ns_amg_info": (C.c_int, [vp, C.POINTER(C.c_int), pd, C.POINTER(C.c_int), C.c_int]),
    "fb_ns_step": (
        C.c_int,
        [vp, dbl, dbl, dbl, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, i64, pi64, pd, i64, pi64, pd, dbl, vp, vp,
         C.POINTER(NSStats)],
    ),
    "fb_ns_velocity_magnitude": (C.c_int, [vp, vp, C.c_int, C.c_int, pd, pd, dbl, vp, pd, pd]),
    "fb_ns_residual": (C.c_int, [vp, dbl, dbl, dbl, dbl, pd, pd, pd, pd, pd, C.c_int]),
    "fb_ns_matrix": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "fb_ns_pressure_rhs": (C.c_int, [vp, dbl, dbl, dbl, C.c_int, pd, pd, pd]),
    "fb_ns_correction_rhs": (C.c_int, [vp, dbl, dbl, dbl, C.c_int, pd, pd, pd, pd]),
    "fb_heat_create": (C.c_int, [vp, vp, pd, dbl, dbl, dbl, pd, C.POINTER(vp)]),
    "fb_heat_create_supg": (C.c_int, [vp, vp, pd, dbl, dbl, dbl, pd, C.c_int, dbl, C.POINTER(vp)]),
    "fb_supg_tau": (C.c_int, [vp, pd, dbl, C.c_int, pd]),
    "fb_heat_supg_mass": (C.c_int, [vp, pd]),
    "fb_heat_destroy": (C.c_int, [vp]),
    "fb_heat_eval": (C.c_int, [vp, dbl, dbl, pd, pd]),
    "fb_heat_solve": (C.c_int, [vp, dbl, dbl, pd, i64, pi64, pd, dbl, C.c_int, pd, C.POINTER(C.c_int)]),
    "fb_heat_matrix": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "fb_stokes_solve": (
        C.c_int,
        [vp, vp, dbl, C.c_int, pd, i64, pi64, pd, i64, pi64, pd, dbl, C.c_int, pd, pd, C.POINTER(C.c_int)],
    ),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here == symbol missing from the .so
    _f.restype = _res
    _f.argtypes = _args


class FlowError(RuntimeError):
    """Raised for every non-zero status; non-convergence keeps the reference's RuntimeError
    contract (error_on_nonconvergence, caught at tests/test_boussinesq.py:254)."""

    def __init__(self, status, message):
        RuntimeError.__init__(self, message)
        self.status = status


def check(status, ctx=None, what=""):
    if status == FB_OK:
        return
    msg = lib.fb_status_string(status).decode()
    if ctx:
        detail = lib.fb_last_error(ctx).decode()
        if detail:
            msg = "%s: %s" % (msg, detail)
    if what:
        msg = "%s: %s" % (what, msg)
    if status == FB_EINVAL:
        raise AssertionError(msg) if "must be > 0" in msg else ValueError(msg)
    raise FlowError(status, msg)


def as_pd(a):
    return a.ctypes.data_as(pd)


def as_pi64(a):
    return a.ctypes.data_as(pi64)


def as_pi32(a):
    return a.ctypes.data_as(pi32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_contexts = {}


def context(device=None):
    """One context per device; device None -> cuda:LOCAL_RANK if a GPU is usable, else host-only (-1)."""
    if device is None:
        device = int(os.environ.get("FLOW_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if device not in _contexts:
            h = vp()
            st = lib.fb_ctx_create(device, C.byref(h))
            if st == FB_ENODEVICE:
                device = -1
            else:
                check(st, None, "fb_ctx_create")
                _contexts[device] = h
    if device not in _contexts:
        h = vp()
        check(lib.fb_ctx_create(device, C.byref(h)), None, "fb_ctx_create")
        _contexts[device] = h
    return _contexts[device]


def has_device():
    ctx = context()
    for dev, h in _contexts.items():
        if h is ctx or h.value == ctx.value:
            return dev >= 0
    return False


# ---- page-locked host arrays (pooled: cudaMallocHost costs ~10 ms per 100 MB) ----------------
_pinned_pool = {}


class _PinnedBlock(object):
    def __init__(self, ctx, nbytes):
        p = vp()
        check(lib.fb_host_alloc(ctx, nbytes, C.byref(p)), ctx, "fb_host_alloc")
        self.ptr, self.nbytes = p, nbytes
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (p.value, False), "version": 3}


def pinned_empty(ctx, n):
    """float64 numpy array of length n in page-locked memory; the block goes back to the pool only when the LAST
    array that refers to its memory is collected.  numpy collapses the `.base` of derived views (reshape, slices) to
    the array that owns the buffer interface, so the finalizer sits on that owner (`base`), not on the view handed
    out: a view taken from the returned array keeps the block alive after the array itself is dropped."""
    import weakref

    nbytes = int(n) * 8
    free = _pinned_pool.setdefault(nbytes, [])
    blk = free.pop() if free else _PinnedBlock(ctx, nbytes)
    base = np.asarray(blk)
    weakref.finalize(base, free.append, blk)
    return base.view(np.float64)


def state_array(ctx, n, threshold=1 << 16, zero=True):
    """Dof array: pinned when a device is present and the array is large.  zero=False skips the fill for arrays
    the library overwrites completely (the outputs of a step: an 80 MB memset per step at 10 M dofs otherwise)."""
    if n >= threshold and has_device():
        a = pinned_empty(ctx, n)
        if zero:
            a[:] = 0.0
        return a
    return np.zeros(n) if zero else np.empty(n)
