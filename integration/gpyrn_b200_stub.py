"""Reference-side binding of libgprn_b200.so -- the file a gpyrn maintainer would add as ``gpyrn/_b200.py``.

It is written against the REFERENCE's classes (``gpyrn.meanfield.inference``, ``gpyrn.covfunc``), not against this
repository's Python package: it binds the C ABI of ``include/gprn_b200.h`` with ctypes and overrides the two hot-path
methods, ``ELBOcalc`` (gpyrn/meanfield.py:561-649) and ``_Prediction`` (:1289-1379), on an existing ``inference``
object.  Everything else of the reference class (parameter bookkeeping, ``nELBO``, ``optimize``, ``mcmc``,
``predict``) keeps working unchanged because it only calls those two.

    import gpyrn
    from gpyrn import _b200
    g = gpyrn.meanfield.inference(q, t, y1, y1err, ...); g.set_components(...)
    _b200.patch(g)            # from here on g.ELBOcalc / g._Prediction run on the B200

``tests/test_reference_integration.py`` applies exactly this file to the unmodified reference (a copy under
``baseline/_ref``) on the GPU box and compares patched against unpatched results.
"""
import ctypes
import os
import types
from itertools import chain

import numpy as np

_LIB = os.environ.get("GPRN_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpyrn_b200",
                                                    "csrc", "libgprn_b200.so"))
_L = ctypes.CDLL(_LIB)
_d = ctypes.POINTER(ctypes.c_double)
_i = ctypes.POINTER(ctypes.c_int32)
_v = ctypes.c_void_p
_L.gprn_last_error.restype = ctypes.c_char_p
_L.gprn_create.argtypes = [ctypes.c_int] * 4 + [_d, _d, _d, ctypes.POINTER(_v)]
_L.gprn_destroy.argtypes = [_v]
_L.gprn_set_model.argtypes = [_v, _i, _i, _i, _i, ctypes.c_int]
_L.gprn_elbo_batched.argtypes = [_v, ctypes.c_int, _d, _d, ctypes.c_int, ctypes.c_int, _d, _d, ctypes.c_int,
                                 _d, _i, _i, _v]
_L.gprn_predict.argtypes = [_v, _d, _d, _d, _d, ctypes.c_int, _d, _d, _d, _d, _d, _v]

OPC = {"SE": 1, "P": 2, "QP": 3, "RQ": 4, "M32": 5, "M52": 6, "WN": 7, "C": 8, "RQP": 9, "COS": 10, "EXP": 11}
DOPC = {"SE": 12, "P": 13, "QP": 14}           # Derivative(k) of the twice-differentiable kernels
# the stationary "other" kernels carry no `_tag` in the reference: by class name (CosPeriodic registers only (P, ell) as
# pars there and cannot be bound through pars)
OPC_BY_CLASS = {"GammaExp": 15, "Piecewise": 16, "Paciorek": 17, "NewPeriodic": 18, "QuasiNewPeriodic": 19,
                "QuasiCosPeriodic": 21}


def program(k):
    """gpyrn.covfunc object -> postfix opcodes (include/gprn_b200.h)."""
    name = type(k).__name__
    if name == "Sum":
        return program(k.k1) + program(k.k2) + [100]
    if name == "Multiplication":
        return program(k.k1) + program(k.k2) + [101]
    if name == "Derivative":
        return [DOPC[k.k._tag]]
    if name in OPC_BY_CLASS:
        return [OPC_BY_CLASS[name]]
    tag = getattr(k, "_tag", None)
    if tag not in OPC:
        raise NotImplementedError(f"gpyrn.covfunc.{name} has no device program in libgprn_b200")
    return [OPC[tag]]


def _p(a, t=_d):
    return a.ctypes.data_as(t)


def _chk(rc):
    if rc:
        raise RuntimeError(_L.gprn_last_error().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _bind_model(gprn, nodes, weights, jitters):
    progs = [[program(k) for k in nodes], [program(k) for k in weights]]
    flat = [np.array(sum(p, []), np.int32) for p in progs]
    offs = [np.r_[0, np.cumsum([len(x) for x in p])].astype(np.int32) for p in progs]
    hyper = _f64(np.concatenate([np.ravel(k.pars) for k in chain(nodes, weights)] + [np.ravel(np.asarray(jitters, float))]))
    _chk(_L.gprn_set_model(gprn._b200, _p(flat[0], _i), _p(offs[0], _i), _p(flat[1], _i), _p(offs[1], _i), hyper.size))
    return hyper


def ELBOcalc(gprn, nodes=None, weights=None, means=None, jitters=None, max_iter=None, mu=None, var=None):
    nodes, weights, means, jitters = gprn._get_components(nodes, weights, means, jitters)
    hyper = _bind_model(gprn, nodes, weights, jitters)
    ysub = _f64(np.concatenate(gprn.y) - gprn._mean(means))                       # meanfield.py:623
    init = 0
    m, v = np.empty(gprn.d), np.empty(gprn.d)
    if isinstance(mu, str) and mu == 'previous':                                   # meanfield.py:598-607
        if gprn._mu is not None:
            m[:] = np.ravel(gprn._mu)
            v[:] = np.ravel(gprn._var)
            init = 1
    elif mu is not None and not isinstance(mu, str):
        m[:] = np.ravel(mu)
        v[:] = np.ravel(var)
        init = 1
    elbo, it, st = np.empty(1), np.zeros(1, np.int32), np.zeros(1, np.int32)
    _chk(_L.gprn_elbo_batched(gprn._b200, 1, _p(hyper), _p(ysub), 1, init, _p(m), _p(v),
                              -1 if max_iter is None else max_iter, _p(elbo), _p(it, _i), _p(st, _i), None))
    shape = (1 + gprn.p, gprn.q, gprn.N)
    if st[0] == 2:
        print('\nMax iterations reached')                                          # meanfield.py:648
    elif st[0] == 0:
        gprn._mu, gprn._var = m.reshape(shape), v.reshape(shape)                   # meanfield.py:644-645
    return float(elbo[0]), m.reshape(shape), v.reshape(shape), int(it[0])


def _Prediction(gprn, nodes=None, weights=None, means=None, jitters=None, tstar=None, mu=None, var=None,
                separate=False):
    nodes = gprn.nodes if nodes is None else nodes
    weights = gprn.weights if weights is None else weights
    means = gprn.means if means is None else means
    jitters = gprn.jitters if jitters is None else jitters
    tstar = gprn.time if tstar is None else np.atleast_1d(np.asarray(tstar, float))
    if mu is None and var is None:                                                 # meanfield.py:1327-1331
        if gprn._mu is None and gprn._var is None:
            _, mu, var, _ = ELBOcalc(gprn, nodes, weights, means, jitters, max_iter=0)
        else:
            mu, var = gprn._mu, gprn._var
    hyper = _bind_model(gprn, nodes, weights, jitters)
    T = tstar.size
    mean_t = _f64(gprn._mean(means, tstar))
    ts, muf, varf = _f64(tstar), _f64(np.ravel(mu)), _f64(np.ravel(var))
    pm, pv = np.empty((T, gprn.p)), np.empty((T, gprn.p))
    npred, wpred = np.empty((gprn.q, T)), np.empty((gprn.q * gprn.p, T))
    _chk(_L.gprn_predict(gprn._b200, _p(hyper), _p(muf), _p(varf), _p(ts), T, _p(mean_t), _p(pm), _p(pv),
                         _p(npred), _p(wpred), None))
    if separate:
        sep = np.empty(2, dtype=object)
        sep[0], sep[1] = npred, wpred
        return pm, pv, sep
    return pm, pv


def patch(gprn, device=0):
    """Create the device handle for this ``gpyrn.meanfield.inference`` object and route its hot path to the B200."""
    h = _v()
    t, y, e = _f64(gprn.time), _f64(gprn.y), _f64(gprn.yerr)
    _chk(_L.gprn_create(device, gprn.N, gprn.p, gprn.q, _p(t), _p(y), _p(e), ctypes.byref(h)))
    gprn._b200 = h
    gprn.ELBOcalc = types.MethodType(ELBOcalc, gprn)
    gprn._Prediction = types.MethodType(_Prediction, gprn)
    return gprn


def unpatch(gprn):
    if getattr(gprn, "_b200", None) is not None:
        _L.gprn_destroy(gprn._b200)
        gprn._b200 = None
    for name in ("ELBOcalc", "_Prediction"):
        gprn.__dict__.pop(name, None)
