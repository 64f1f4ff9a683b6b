"""CPU tests of the host drivers above the batched device call (SURVEY.md §8 f.1 / VERDICT r1 item 6): the lock-step
multi-start optimiser and the vectorised log-posterior.  ``ELBO_batch`` -- the one device call they make -- is replaced
by an analytic stand-in, so what is checked is the coroutine scheduling, the row bookkeeping and the prior handling;
the device path itself is checked by the ``-m gpu`` tests (test_optimize_batch_equals_sequential_scipy_runs, ...)."""
import numpy as np
import pytest
from scipy import stats
from scipy.optimize import minimize

import gpyrn_b200 as gp
from gpyrn_b200 import covfunc, meanfunc


def _model():
    t = np.linspace(0.0, 10.0, 8)
    g = gp.inference(1, t, np.sin(t), 0.1 * np.ones(8))
    g.set_components(covfunc.SquaredExponential(1.0, 2.0), covfunc.SquaredExponential(1.0, 3.0), meanfunc.Constant(0.0), 0.1)
    return g


class _FakeDevice:
    """Stands in for inference.ELBO_batch(P, state='previous', work_source=...): ELBO(x) = -(|x - c|^2 + 1) for the rows
    the work source hands out; records which rows every call evaluated."""

    def __init__(self, centre):
        self.centre = np.asarray(centre, dtype=float)
        self.calls = []

    def objective(self, x):
        return float(np.sum((np.asarray(x) - self.centre) ** 2) + 1.0)

    def __call__(self, P, max_iter=None, state=None, work_source=None, **kw):
        assert state == 'previous' and work_source is not None
        P = np.atleast_2d(P)
        elbo, taken, rows = np.zeros(len(P)), np.zeros(len(P), dtype=bool), []
        while True:
            i = work_source()
            if i < 0:
                break
            rows.append(i)
            elbo[i], taken[i] = -self.objective(P[i]), True
        self.calls.append(rows)
        return elbo, taken


def _patch(g, fake, monkeypatch):
    monkeypatch.setattr(g, "ELBO_batch", fake)
    monkeypatch.setattr(g, "reset_chain_state", lambda *a, **k: None)
    monkeypatch.setattr(g, "_ensure_chain_state", lambda *a, **k: None)


def test_optimize_batch_runs_every_start_as_scipy_would(monkeypatch):
    g = _model()
    n = g.get_parameters().size
    fake = _FakeDevice(np.linspace(0.5, 2.0, n))
    _patch(g, fake, monkeypatch)
    rng = np.random.default_rng(0)
    starts = fake.centre + rng.normal(0.0, 0.5, (5, n))
    starts[3] = fake.centre                                    # a start that converges much earlier than the others
    opts = {"maxfev": 4000, "xatol": 1e-6, "fatol": 1e-9}
    res = g.optimize_batch(starts, options=opts)
    assert len(res) == 5
    for s in range(5):                                         # the same trajectory as a sequential run from that start
        seq = minimize(fake.objective, starts[s], method="Nelder-Mead", options=opts)
        assert res[s].nfev == seq.nfev and res[s].nit == seq.nit
        assert np.array_equal(res[s].x, seq.x) and res[s].fun == seq.fun
    # lock-step: a round evaluates each live start at most once; the number of rounds is the longest start's count
    assert g.n_batch_calls == len(fake.calls) == max(r.nfev for r in res)
    assert all(len(set(rows)) == len(rows) for rows in fake.calls)
    assert sum(len(rows) for rows in fake.calls) == sum(r.nfev for r in res)
    nfev = [r.nfev for r in res]
    assert len(set(nfev)) > 1                                  # the starts finish at different times ...
    assert sorted(fake.calls[0]) == [0, 1, 2, 3, 4]            # ... all are in the first round,
    assert sorted(fake.calls[-1]) == [s for s in range(5) if nfev[s] == max(nfev)]     # only the longest in the last
    best = min(res, key=lambda r: r.fun)
    assert np.array_equal(g.get_parameters(), best.x)          # the object's parameters are set to the best start's
    with pytest.raises(ValueError):
        g.optimize_batch(np.zeros((2, n + 1)))


def test_optimize_batch_reports_a_failing_device_call(monkeypatch):
    g = _model()
    n = g.get_parameters().size

    def broken(P, **kw):
        raise RuntimeError("device call failed")
    _patch(g, broken, monkeypatch)
    with pytest.raises(RuntimeError):
        g.optimize_batch(np.ones((3, n)))


def test_logposterior_batch_rows_priors_and_rejections(monkeypatch):
    g = _model()
    names = np.array(list(g.parameters_dict.keys()))[~g.frozen_mask]
    n = names.size
    fake = _FakeDevice(np.ones(n))
    _patch(g, fake, monkeypatch)
    priors = {k: stats.uniform(0.0, 5.0) for k in names}
    thetas = np.full((4, n), 1.5)
    thetas[2, 0] = -1.0                                        # outside the prior support: never sent to the device
    total, elbo = g.logposterior_batch(thetas, priors)
    assert fake.calls == [[0, 1, 3]]
    assert np.isneginf(total[2]) and np.isneginf(elbo[2])
    lp = n * np.log(1.0 / 5.0)
    for r in (0, 1, 3):
        assert elbo[r] == -fake.objective(thetas[r]) and abs(total[r] - (lp + elbo[r])) < 1e-12
    total, elbo = g.logposterior_batch(thetas, priors, rows=[1, 2])   # a half step of the sampler: only the moving rows
    assert fake.calls[-1] == [1] and np.isneginf(total[[0, 2, 3]]).all() and np.isfinite(total[1])
    thetas[:, 0] = -1.0                                        # nothing in support: no device call at all
    ncalls = len(fake.calls)
    total, _ = g.logposterior_batch(thetas, priors)
    assert np.isneginf(total).all() and len(fake.calls) == ncalls
