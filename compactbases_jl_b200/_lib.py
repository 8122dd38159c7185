"""ctypes binding of ``libcompactbases_cuda.so`` (the C ABI in ``include/compactbases_cuda.h``).

This is the Python twin of the Julia ``ccall`` glue shown in INTEGRATION.md: it
declares the argument types of every exported entry point and turns a non-zero
status into an exception of the kind the reference throws
(``ArgumentError`` -> ``ValueError``, ``DimensionMismatch`` -> ``DimensionMismatch``).

There is no fallback: if the shared library is missing, importing the product
API raises ``ImportError`` naming the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CB_CUDA_LIB overrides the path (development: kernel-variant sweeps); the default is the in-tree build
LIB_PATH = os.environ.get("CB_CUDA_LIB") or os.path.join(_HERE, "libcompactbases_cuda.so")

CB_OK, CB_ERR_ARG, CB_ERR_CUDA, CB_ERR_UNSUPPORTED, CB_ERR_ALLOC = 0, 1, 2, 3, 4

i64 = C.c_int64
cint = C.c_int
dbl = C.c_double
vp = C.c_void_p


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (thrown before any work is done)."""


class CudaError(RuntimeError):
    pass


# name -> (restype, argtypes); pointers to device/host arrays are passed as void*
_SIGNATURES = {
    "cb_version": (C.c_char_p, []),
    "cb_last_error": (C.c_char_p, []),
    "cb_device_count": (cint, [C.POINTER(cint)]),
    "cb_bspline_create": (cint, [vp, i64, cint, cint, cint, vp, vp, cint, C.POINTER(vp)]),
    "cb_bspline_destroy": (cint, [vp]),
    "cb_bspline_numfunctions": (i64, [vp]),
    "cb_bspline_numknots": (i64, [vp]),
    "cb_bspline_num_nonempty": (i64, [vp]),
    "cb_bspline_numnodes": (i64, [vp]),
    "cb_bspline_nodes": (cint, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "cb_bspline_nodes_copy": (cint, [vp, vp, vp, vp]),
    "cb_bspline_eval": (cint, [vp, vp, i64, cint, vp, vp, vp]),
    "cb_bspline_eval_host": (cint, [vp, vp, i64, cint, vp, vp]),
    "cb_bspline_eval_csc": (cint, [vp, vp, i64, i64, i64, vp, vp, vp, C.POINTER(i64), vp]),
    "cb_bspline_eval_spline": (cint, [vp, vp, i64, cint, vp, i64, i64, vp, i64, vp]),
    "cb_bspline_quad_table": (cint, [vp, cint, vp, vp]),
    "cb_bspline_assemble": (cint, [vp, vp, cint, cint, vp, i64, i64, i64, i64, cint, cint, vp, vp]),
    "cb_bspline_assemble_coulomb": (cint, [vp, vp, dbl, i64, i64, i64, i64, cint, cint, vp, vp]),
    "cb_bspline_assemble_part": (cint, [vp, vp, cint, cint, vp, cint, dbl, i64, i64, i64, i64, cint, cint,
                                        i64, i64, i64, i64, vp, vp]),
    "cb_bspline_assemble_host": (cint, [vp, vp, cint, cint, vp, cint, dbl, i64, i64, i64, i64, cint, cint, vp, vp]),
    "cb_bspline_assemble_pair": (cint, [vp, vp, cint, dbl, i64, i64, vp, vp, vp]),
    "cb_band_to_tridiag": (cint, [vp, i64, vp, vp, vp, vp]),
    "cb_complex_split": (cint, [vp, i64, i64, i64, vp, i64, vp]),
    "cb_complex_combine": (cint, [vp, vp, i64, i64, i64, dbl, dbl, dbl, dbl, vp, i64, vp]),
    "cb_fedvr_assemble_complex": (cint, [vp, vp, vp, vp, vp, vp, vp, i64, cint, cint, cint, i64, dbl, vp, vp, vp]),
    "cb_fedvr_assemble": (cint, [vp, vp, vp, vp, vp, vp, vp, i64, cint, cint, cint, vp, vp]),
    "cb_fedvr_to_dense": (cint, [vp, vp, vp, vp, i64, i64, i64, vp, vp]),
    "cb_fedvr_to_band": (cint, [vp, vp, vp, vp, i64, i64, i64, cint, vp, vp]),
    "cb_fedvr_eval": (cint, [vp, vp, vp, vp, vp, i64, cint, vp, i64, vp, vp, vp]),
    "cb_fd_eval": (cint, [vp, vp, i64, vp, i64, vp, vp, vp]),
    "cb_fedvr_to_skyline": (cint, [vp, vp, vp, vp, i64, i64, i64, vp, i64, vp, vp, vp, vp, vp, i64, vp]),
    "cb_fedvr_to_tridiag": (cint, [vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, vp]),
    "cb_fedvr_apply": (cint, [vp, vp, vp, vp, i64, cint, i64, i64, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_fedvr_assemble_apply": (cint, [vp, vp, vp, vp, vp, vp, vp, i64, cint, cint, cint, vp, i64, i64, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_fedvr_apply_host": (cint, [vp, vp, vp, vp, i64, cint, i64, i64, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_fd_derivative": (cint, [cint, cint, vp, vp, i64, dbl, i64, i64, dbl, cint, vp, vp, vp, vp]),
    "cb_fd_derivative_banded": (cint, [cint, cint, vp, vp, i64, dbl, i64, i64, i64, i64, dbl, cint, vp, vp]),
    "cb_fd_implicit_stencils": (cint, [vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp]),
    "cb_banded_apply": (cint, [vp, i64, i64, cint, cint, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_tridiag_apply": (cint, [vp, vp, vp, i64, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_tridiag_apply_host": (cint, [vp, vp, vp, i64, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_banded_apply_host": (cint, [vp, i64, i64, cint, cint, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_host_register": (cint, [vp, C.c_size_t]),
    "cb_host_unregister": (cint, [vp]),
    "cb_diag_apply": (cint, [vp, i64, vp, i64, i64, dbl, dbl, vp, i64, vp]),
    "cb_axpby": (cint, [i64, i64, dbl, vp, i64, dbl, vp, i64, vp]),
    "cb_batched_dot": (cint, [vp, i64, vp, i64, i64, i64, dbl, vp, vp]),
    "cb_density_bspline_create": (cint, [vp, i64, i64, vp, C.POINTER(vp), vp]),
    "cb_density_destroy": (cint, [vp]),
    "cb_density_set_refine": (cint, [vp, cint]),
    "cb_density_apply": (cint, [vp, vp, i64, i64, vp, i64, i64, cint, vp, i64, vp]),
    "cb_density_apply_host": (cint, [vp, vp, i64, i64, vp, i64, i64, cint, vp, i64, vp]),
    "cb_density_nodal_product": (cint, [vp, vp, i64, i64, vp, i64, i64, cint, vp, i64, vp]),
    "cb_density_orthogonal": (cint, [vp, i64, vp, i64, i64, vp, i64, i64, cint, vp, i64, vp]),
    "cb_density_project": (cint, [vp, vp, i64, i64, vp, i64, vp]),
    "cb_bandchol_create": (cint, [vp, i64, cint, C.POINTER(vp), vp]),
    "cb_bandldl_create": (cint, [vp, i64, cint, C.POINTER(vp), vp]),
    "cb_bandchol_solve": (cint, [vp, vp, i64, i64, vp]),
    "cb_bandchol_destroy": (cint, [vp]),
    "cb_bandlu_create": (cint, [vp, i64, cint, cint, C.POINTER(vp), vp]),
    "cb_bandlu_solve": (cint, [vp, vp, i64, i64, vp]),
    "cb_bandlu_destroy": (cint, [vp]),
    "cb_batch_range": (cint, [i64, cint, cint, C.POINTER(i64), C.POINTER(i64)]),
    "cb_comm_create": (cint, [cint, cint, C.POINTER(vp)]),
    "cb_comm_export": (cint, [vp, vp]),
    "cb_comm_connect": (cint, [vp, vp]),
    "cb_comm_connect_local": (cint, [C.POINTER(vp), cint]),
    "cb_comm_status": (cint, [vp, C.POINTER(cint), vp]),
    "cb_comm_destroy": (cint, [vp]),
    "cb_interval_shard": (cint, [vp, i64, i64, cint, cint] + [C.POINTER(i64)] * 5),
    "cb_bspline_assemble_sharded": (cint, [vp, vp, vp, cint, cint, vp, cint, dbl, i64, i64, i64, i64, cint, cint, cint, vp, vp]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (once) and declares every signature."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C compactbases_jl_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().cb_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Maps a C status to the exception the reference would have thrown."""
    if status == CB_OK:
        return
    msg = last_error()
    if status == CB_ERR_ARG:
        if msg.startswith("DimensionMismatch") or "leading dimension" in msg:
            raise DimensionMismatch(msg)
        raise ValueError(msg)
    if status == CB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == CB_ERR_ALLOC:
        raise MemoryError(msg)
    raise CudaError(msg)


def call(name: str, *args):
    check(getattr(load(), name)(*args))
