"""ctypes binding of libgprn_b200.so (C ABI in include/gprn_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
calls raise.  ``build_library()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("GPRN_B200_LIB", os.path.join(CSRC, "libgprn_b200.so"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

NEXT_SET_FN = ctypes.CFUNCTYPE(ctypes.c_int64, ctypes.c_void_p)      # gprn_next_set_fn
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

# symbol -> (restype, argtypes); every symbol include/gprn_b200.h declares
SIGNATURES = {
    "gprn_last_error": (ctypes.c_char_p, []),
    "gprn_version": (ctypes.c_int, []),
    "gprn_built_for_sm": (ctypes.c_int, []),
    "gprn_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
                                   c_double_p, ctypes.POINTER(ctypes.c_void_p)]),
    "gprn_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "gprn_set_workspace_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64]),
    "gprn_set_max_slots": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "gprn_set_model": (ctypes.c_int, [ctypes.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, ctypes.c_int]),
    "gprn_elbo_batched": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int,
                                         ctypes.c_int, c_double_p, c_double_p, ctypes.c_int, c_double_p, c_int32_p,
                                         c_int32_p, ctypes.c_void_p]),
    "gprn_upload_ysub": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "gprn_elbo_batched_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "gprn_elbo_pool": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, c_double_p,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int, ctypes.c_void_p]),
    "gprn_chain_resize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "gprn_chain_set": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, c_double_p, c_double_p]),
    "gprn_chain_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, c_double_p, c_double_p,
                                      c_int32_p]),
    "gprn_chain_invalidate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]),
    "gprn_kmatrix": (ctypes.c_int, [ctypes.c_void_p, c_int32_p, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p,
                                    ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_double, c_double_p,
                                    ctypes.c_void_p]),
    "gprn_keval": (ctypes.c_int, [ctypes.c_int, c_int32_p, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p,
                                  ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_double_p]),
    "gprn_predict": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, ctypes.c_int,
                                    c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, ctypes.c_void_p]),
    "gprn_predict_batched": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p, c_double_p,
                                            c_double_p, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                            c_double_p, c_double_p, ctypes.c_void_p]),
    "gprn_sample": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_double, c_double_p,
                                   ctypes.c_void_p]),
    "gprn_debug_factor": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p, c_double_p,
                                         c_double_p]),
    "gprn_debug_panel_stress": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_int64)]),
    "gprn_graph_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "gprn_last_rounds": (ctypes.c_int64, [ctypes.c_void_p]),
    "gprn_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "gprn_reset_launch_count": (ctypes.c_int, [ctypes.c_void_p]),
    "gprn_last_elbo_ms": (ctypes.c_double, [ctypes.c_void_p]),
    "gprn_last_total_iters": (ctypes.c_int64, [ctypes.c_void_p]),
}

_lib = None


class GprnError(RuntimeError):
    pass


def build_library(verbose=False):
    """Compile gpyrn_b200/csrc/gprn_api.cu -> libgprn_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(CSRC, "gprn_api.cu")
    out = os.path.join(CSRC, "libgprn_b200.so")
    extra = os.environ.get("GPRN_NVCC_EXTRA", "").split()
    cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-o", out, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


def lib():
    """The loaded shared library (ctypes.CDLL) with argtypes set.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GprnError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a).  gpyrn_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise GprnError(lib().gprn_last_error().decode())


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
