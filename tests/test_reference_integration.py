"""Boundary proof: the ctypes stub of INTEGRATION.md
(``integration/gpyrn_b200_stub.py``) applied to the UNMODIFIED reference class.

``baseline/_ref/gpyrn`` is a git-ignored copy of the reference package made by ``__graft_entry__.build()`` in the build
container; it travels to the GPU box with the snapshot.  The test builds a real ``gpyrn.meanfield.inference`` with real
``gpyrn.covfunc`` / ``gpyrn.meanfunc`` objects, runs the reference's own CPU ``ELBOcalc`` / ``_Prediction``, patches the
object with the stub and runs them again on the B200: ELBO 1e-10, identical iterations, prediction 1e-8 -- and the
reference's own drivers (``nELBO``, ``optimize``) keep working on top of the patched methods.
"""
import numpy as np
import pytest

from tests import _ref_shim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _ref_shim.available(), reason="baseline/_ref/gpyrn not present")]


def _reference():
    _ref_shim.install()
    from gpyrn import covfunc, meanfunc, meanfield
    return covfunc, meanfunc, meanfield


def _data(N, p, seed):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 40 * N ** 0.5, N))
    args = []
    for i in range(p):
        args += [np.sin(2 * np.pi * t / 25 + i) * (1 + 0.3 * i) + 0.1 * rng.standard_normal(N), rng.uniform(.05, .15, N)]
    return t, args


@pytest.mark.parametrize("shape", [(60, 2, 1), (90, 3, 2), (64, 2, 1, "other kernels")])
def test_stub_on_unmodified_reference(shape, capsys):
    from integration import gpyrn_b200_stub as stub
    covfunc, meanfunc, meanfield = _reference()
    N, p, q = shape[:3]
    t, args = _data(N, p, 17)
    g = meanfield.inference(q, t, *args)
    if len(shape) > 3:        # the reference's stationary "other" kernels (bound by class name: they carry no _tag)
        nodes = [covfunc.QuasiNewPeriodic(1.0, 1.5, 60.0, 25.0, 0.9)]
        weights = [covfunc.GammaExp(1.1, 1.6, 120.0) * covfunc.Piecewise(900.0) + covfunc.WhiteNoise(0.03),
                   covfunc.Paciorek(0.9, 70.0, 110.0) + covfunc.SquaredExponential(0.4, 90.0) * covfunc.NewPeriodic(1.0, 2.0, 50.0, 1.5)]
    else:
        nodes = [covfunc.QuasiPeriodic(1 + .2 * j, 60 + 5 * j, 25 + j, .7) if j == 0 else covfunc.Matern52(1.2, 35.0)
                 for j in range(q)]
        weights = [covfunc.SquaredExponential(1 + .1 * k, 80 + k) for k in range(q * p)]
    g.set_components(nodes, weights, [meanfunc.Constant(0.1 * i) for i in range(p)], [0.1] * p)
    tstar = np.linspace(t[0] - 3, t[-1] + 4, 41)
    # the reference's own CPU path
    e_ref, mu_ref, var_ref, it_ref = g.ELBOcalc()
    pm_ref, pv_ref = g._Prediction(tstar=tstar, mu=np.asarray(mu_ref), var=np.asarray(var_ref))
    val_ref = g.nELBO(g.get_parameters())              # the reference's warm start from its converged state
    g._mu = g._var = None
    # the same object, hot path on the B200
    stub.patch(g)
    try:
        e, mu, var, it = g.ELBOcalc()
        assert it == it_ref
        assert abs(e - e_ref) <= 1e-10 * abs(e_ref), (e, e_ref)
        assert np.max(np.abs(mu - np.asarray(mu_ref))) <= 1e-8 * np.max(np.abs(mu_ref))
        pm, pv = g._Prediction(tstar=tstar, mu=np.asarray(mu_ref), var=np.asarray(var_ref))
        assert np.max(np.abs(pm - pm_ref)) <= 1e-8 * np.max(np.abs(pm_ref))
        assert np.max(np.abs(pv - pv_ref)) <= 1e-8 * np.max(np.abs(pv_ref))
        # reference drivers on top of the patched methods: nELBO warm-started from the same converged state,
        # the ELBO property
        g._mu, g._var = np.asarray(mu_ref), np.asarray(var_ref)
        val = g.nELBO(g.get_parameters())
        assert abs(val - val_ref) <= 1e-10 * abs(val_ref), (val, val_ref)
        assert abs(g.ELBO - e_ref) <= 1e-10 * abs(e_ref)
    finally:
        stub.unpatch(g)
    capsys.readouterr()
    # unpatched again: the class method is back
    assert g.ELBOcalc.__func__ is meanfield.inference.ELBOcalc
