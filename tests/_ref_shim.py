"""Import shim that lets the UNMODIFIED reference package (``gpyrn``) import in this image (test infrastructure).

The reference needs jax, emcee and matplotlib, none of which exist here, and trips over numpy >= 2 (``np.float``).
It runs from its own sources once stand-in modules are registered for those imports: ``jax.numpy`` backed by numpy,
``jax.jit`` as identity, ``jax.scipy.linalg.cho_solve`` = scipy's.  Arithmetic deviation
from real JAX is rounding-level only (LAPACK potrf / potrs either way).  ``install(path)`` puts the directory that
holds the ``gpyrn`` package first on ``sys.path``: ``/root/reference`` in the build container (golden generation),
``baseline/_ref`` (a git-ignored copy made by ``__graft_entry__.build()``) on the GPU box.
"""
import os
import sys
import types

import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "baseline", "_ref")


def available(path=REF_COPY):
    return os.path.isfile(os.path.join(path, "gpyrn", "meanfield.py"))


def install(path=REF_COPY):
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    def jit(f=None, static_argnums=None, **kw):
        return (lambda g: g) if f is None else f

    if "jax" not in sys.modules:
        jnp = mod("jax.numpy")
        jnp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
        jnp.ndarray = np.ndarray
        jsl = mod("jax.scipy.linalg", cho_solve=scipy.linalg.cho_solve)
        js = mod("jax.scipy", linalg=jsl)

        class _Cfg:
            def update(self, *a, **k):
                pass

        mod("jax", jit=jit, numpy=jnp, scipy=js, config=_Cfg())
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")
    if "emcee" not in sys.modules:
        em = mod("emcee", EnsembleSampler=object, backends=types.SimpleNamespace())
        em.utils = mod("emcee.utils", sample_ellipsoid=None)
    if not hasattr(np, "float"):
        np.float = float
    if path not in sys.path:
        sys.path.insert(0, path)
