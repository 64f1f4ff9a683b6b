"""The ensemble sampler behind ``inference.mcmc`` (gpyrn_b200/sampler.py) on the CPU: the stretch move samples a known
Gaussian, rows-aware evaluation sees exactly the moving half, the autocorrelation-time estimator recovers an AR(1)
chain's value, the .npz backend round-trips.  (The device side -- one batched call per half step, warm starts keyed by
walker -- is exercised by the GPU test ``test_mcmc_driver_runs_on_the_batched_path``.)"""
import numpy as np
import pytest

from gpyrn_b200.sampler import EnsembleSampler, NpzBackend, State, integrated_time


def test_stretch_move_samples_a_gaussian():
    mean, sd = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    calls = []

    def logp(x):
        calls.append(x.shape[0])
        lp = -0.5 * np.sum(((x - mean) / sd) ** 2, axis=1)
        return np.column_stack([lp, 2.0 * lp])             # one blob column

    s = EnsembleSampler(12, 3, logp, seed=1)
    p0 = mean + 0.1 * np.random.default_rng(0).standard_normal((12, 3))
    s.run_mcmc(p0, 3000)
    assert set(calls[1:]) == {6}                            # every half step: ONE call with half of the walkers
    assert s.nevals == 1 + 2 * 3000
    chain = s.get_chain(flat=True, discard=500)
    assert chain.shape == (2500 * 12, 3)
    assert np.all(np.abs(chain.mean(axis=0) - mean) < 0.1 * sd)
    assert np.all(np.abs(chain.std(axis=0) / sd - 1.0) < 0.1)
    assert np.all((s.acceptance_fraction > 0.3) & (s.acceptance_fraction < 0.9))
    assert np.allclose(s.get_blobs()[..., 0], 2.0 * s.get_log_prob())
    tau = s.get_autocorr_time(tol=0)
    assert tau.shape == (3,) and np.all((tau > 1.0) & (tau < 200.0))


def test_rows_aware_log_prob_sees_the_moving_half():
    seen = []

    def logp(coords, rows):
        assert coords.shape == (8, 2)
        seen.append(np.array(rows))
        out = np.full(8, np.nan)                            # rows that do not move must never be read
        out[rows] = -0.5 * np.sum(coords[rows] ** 2, axis=1)
        return out

    s = EnsembleSampler(8, 2, logp, seed=3, rows_aware=True)
    s.run_mcmc(np.random.default_rng(1).standard_normal((8, 2)), 20)
    assert sorted(seen[0].tolist()) == list(range(8))       # the initial evaluation: everyone
    for a, b in zip(seen[1::2], seen[2::2]):
        assert a.size == b.size == 4 and sorted(np.r_[a, b].tolist()) == list(range(8))
    assert s.get_chain().shape == (20, 8, 2)


def test_integrated_time_of_an_ar1_chain():
    rng = np.random.default_rng(5)
    phi, n = 0.9, 20000
    x = np.empty((n, 4))
    x[0] = rng.standard_normal(4)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + np.sqrt(1 - phi ** 2) * rng.standard_normal(4)
    tau = integrated_time(x[:, :, None])
    assert abs(tau[0] - (1 + phi) / (1 - phi)) < 2.5           # exact value 19
    with pytest.raises(ValueError):
        integrated_time(x[:200, :, None], tol=50, quiet=False)


def test_validation_and_backend(tmp_path):
    with pytest.raises(ValueError):
        EnsembleSampler(5, 2, lambda x: x[:, 0])
    with pytest.raises(ValueError):
        EnsembleSampler(2, 2, lambda x: x[:, 0])
    s = EnsembleSampler(4, 2, lambda x: -0.5 * np.sum(x ** 2, axis=1), seed=0,
                        backend=NpzBackend(str(tmp_path / "chain.npz"), every=7))
    with pytest.raises(ValueError):
        next(s.sample(np.zeros((3, 2))))
    with pytest.raises(ValueError):
        next(s.sample(State(np.zeros((4, 2)), log_prob=np.array([0.0, -np.inf, 0.0, 0.0]))))
    s.run_mcmc(np.random.default_rng(2).standard_normal((4, 2)), 30)
    d = NpzBackend.load(str(tmp_path / "chain.npz"))
    assert int(d["iteration"]) == 30 and d["chain"].shape == (30, 4, 2) and d["log_prob"].shape == (30, 4)
    assert np.array_equal(d["chain"], s.get_chain()) and d["blobs"].shape == (30, 4, 0)
