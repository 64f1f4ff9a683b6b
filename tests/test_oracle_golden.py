"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py, reference executed under an import shim in the build container).

Tolerances: ELBO 1e-10 relative (north_star), identical iteration count, per-iteration trace 1e-9,
prediction mean/var 1e-8 (north_star).  Kernel matrices: 1e-14 (same numpy ufuncs)."""
import ast

import numpy as np
import pytest

from oracle import gprn_oracle as orc
from tests._cases import GOLDEN, golden_names, load_golden, oracle_model, relerr

SMALL = [n for n in golden_names() if not n.startswith(("c2_", "c1_"))]


@pytest.mark.parametrize("fixture", ["kernels.npz", "extra_kernels.npz"])
def test_kernel_vectors(fixture):
    z = np.load(GOLDEN + "/" + fixture)
    t, ts = z["t"], z["tstar"]
    for i, s in enumerate(z["specs"]):
        spec = ast.literal_eval(str(s))
        np.testing.assert_allclose(orc.kmatrix(spec, t), z[f"Ksq_{i}"], rtol=1e-14, atol=1e-300)
        np.testing.assert_allclose(orc.kmatrix(spec, ts, t), z[f"Krect_{i}"], rtol=1e-14, atol=1e-300)


@pytest.mark.parametrize("name", golden_names())
def test_elbo_matches_reference(name):
    d = load_golden(name)
    m = oracle_model(d)
    elbo, mu, var, it, trace = orc.elbo_calc(m, max_iter=d["max_iter"], return_trace=True)
    assert it == d["iters"]
    assert abs(elbo - d["elbo"]) <= 1e-10 * abs(d["elbo"])
    assert relerr(trace, d["trace"]) < 1e-9
    assert mu.shape == d["mu"].shape == (1 + m.p, m.q, m.N)
    scale = np.max(np.abs(d["mu"]))
    assert np.max(np.abs(mu - d["mu"])) < 1e-7 * scale
    assert np.max(np.abs(var - d["var"])) < 1e-7 * np.max(np.abs(d["var"]))


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "c3_synth_256_4_1_QP"])
def test_prediction_matches_reference(name):
    d = load_golden(name)
    m = oracle_model(d)
    T = d["tstar"].size
    mean_t = np.repeat(d["mean_consts"][:, None], T, axis=1)
    pm, pv, nP, wP = orc.prediction(m, d["tstar"], d["mu"], d["var"], mean_t)
    assert np.max(np.abs(pm - d["pred_mean"])) <= 1e-8 * np.max(np.abs(d["pred_mean"]))
    assert np.max(np.abs(pv - d["pred_var"])) <= 1e-8 * np.max(np.abs(d["pred_var"]))
    assert np.max(np.abs(nP - d["node_pred"])) <= 1e-8 * np.max(np.abs(d["node_pred"]))
    assert np.max(np.abs(wP - d["weight_pred"])) <= 1e-8 * np.max(np.abs(d["weight_pred"]))


def test_init_mu_var_scramble():
    """Q5: weights written (q,p,N) but read (p,q,N); only first p weight amplitudes used."""
    d = load_golden("synth_60_2_2_QP_means")
    m = oracle_model(d)
    mu, var = orc.init_mu_var(m)
    assert mu.shape == (m.d,) and var.shape == (m.d,)
    assert np.allclose(var[: m.q * m.N], np.mean(m.jitters))
