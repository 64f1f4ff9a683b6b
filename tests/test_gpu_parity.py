"""GPU parity tests: the CUDA path, called through the Python mirror of the reference API (ctypes ->
C ABI), against (a) golden vectors produced by the unmodified reference, (b) the CPU oracle on the
same seeded inputs, (c) size-independent properties at the full BASELINE sizes.

Bars (north_star): ELBO 1e-10 relative with identical iteration count; predictive mean / variance
1e-8 relative; kernel matrices 1e-13 relative (FP64 exp/sin/pow within a few ulp of numpy's).
"""
import ast
import os

import numpy as np
import pytest

import gpyrn_b200 as gp
from gpyrn_b200 import _lib, covfunc, meanfunc
from oracle import gprn_oracle as orc
from tests._cases import GOLDEN, golden_names, load_golden

pytestmark = pytest.mark.gpu

KCLS = {"SE": covfunc.SquaredExponential, "P": covfunc.Periodic, "QP": covfunc.QuasiPeriodic,
        "RQ": covfunc.RationalQuadratic, "M32": covfunc.Matern32, "M52": covfunc.Matern52,
        "WN": covfunc.WhiteNoise, "C": covfunc.Constant, "RQP": covfunc.RQP, "COS": covfunc.Cosine,
        "EXP": covfunc.Exponential}


def build_kernel(spec):
    if spec[0] == "sum":
        return build_kernel(spec[1]) + build_kernel(spec[2])
    if spec[0] == "mul":
        return build_kernel(spec[1]) * build_kernel(spec[2])
    if spec[0] in ("dSE", "dP", "dQP"):
        return covfunc.Derivative(KCLS[spec[0][1:]](*spec[1:]))
    return KCLS[spec[0]](*spec[1:])


def inference_from(t, ys, es, nodes, weights, mean_consts, jitters):
    args = []
    for y, e in zip(ys, es):
        args += [y, e]
    g = gp.inference(len(nodes), t, *args)
    g.set_components([build_kernel(s) for s in nodes], [build_kernel(s) for s in weights],
                     [meanfunc.Constant(c) for c in mean_consts], list(jitters))
    return g


def from_golden(d):
    return inference_from(d["t"], d["y"], d["yerr"], d["nodes"], d["weights"], d["mean_consts"], d["jitters"])


def from_oracle_model(m):
    return inference_from(m.time, m.y, m.yerr, m.nodes, m.weights, m.mean_vals[:, 0], m.jitters)


def full_parameters(m, theta):
    """oracle hyper vector [kernel pars, jitters] -> get_parameters order [kernel pars, means(=const), jitters]."""
    theta = np.atleast_2d(theta)
    return np.concatenate([theta[:, :-m.p], np.tile(m.mean_vals[:, 0], (theta.shape[0], 1)), theta[:, -m.p:]], axis=1)


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


# ---------------------------------------------------------------------------------------------
# a1 / a2: kernels and covariance-matrix assembly
# ---------------------------------------------------------------------------------------------
def test_kernel_matrices_match_reference():
    z = np.load(os.path.join(GOLDEN, "kernels.npz"))
    t, ts = z["t"], z["tstar"]
    g = gp.inference(1, t, np.zeros_like(t), np.ones_like(t))
    for i, s in enumerate(z["specs"]):
        k = build_kernel(ast.literal_eval(str(s)))
        # relative 1e-13; absolute floor 2e-15 of the amplitude for kernels that cross zero (Cosine)
        tol = dict(rtol=1e-13, atol=2e-15 * np.max(np.abs(z[f"Ksq_{i}"])))
        np.testing.assert_allclose(g._kmat(k, t, None, 0.0), z[f"Ksq_{i}"], **tol)
        np.testing.assert_allclose(g._predictKMatrix(k, ts), z[f"Krect_{i}"], **tol)
        np.testing.assert_allclose(k(t[:, None] - t[None, :]), z[f"Ksq_{i}"], **tol)
        np.testing.assert_allclose(g._KMatrix(k), z[f"Ksq_{i}"] + 1e-6 * np.eye(t.size), **tol)


def test_QP_equals_prod():
    """reference tests/test_cov_functions.py:7-14"""
    k1 = covfunc.SquaredExponential(1, 10) * covfunc.Periodic(1, 20, 0.5)
    k2 = covfunc.QuasiPeriodic(1, 10, 20, 0.5)
    t = np.sort(np.random.uniform(0, 100, size=50))
    T = t[:, None] - t[None, :]
    assert np.allclose(k1(T), k2(T))


def test_whitenoise_shape_quirk():
    """covfunc.py:144-148: identity by position for square input, constant otherwise (Q9)."""
    k = covfunc.WhiteNoise(0.5)
    assert np.allclose(k(np.ones((4, 4))), 0.25 * np.eye(4))
    assert np.allclose(k(np.zeros((3, 4))), 0.25)
    assert np.allclose(k(np.zeros(5)), 0.25)


@pytest.mark.parametrize("n", [1, 17, 64, 65, 200, 500])
def test_factorisation_kernels(n):
    rng = np.random.default_rng(n)
    tt = np.sort(rng.uniform(0, 40 * n ** 0.5 + 1, n))
    A = orc.kmatrix(("M52", 1.0, 30.0), tt, nugget=1e-6) + np.diag(rng.uniform(0.01, 1.0, n))
    g = gp.inference(1, np.arange(4.0), np.zeros(4), np.ones(4))
    L, X, ld = np.empty((n, n)), np.empty((n, n)), np.zeros(1)
    _lib.check(_lib.lib().gprn_debug_factor(g._h(), n, _lib.dptr(_lib.f64(A)), _lib.dptr(L), _lib.dptr(X), _lib.dptr(ld)))
    Lr = np.linalg.cholesky(A)
    assert rel(L, Lr) < 1e-11
    assert np.max(np.abs(X @ Lr - np.eye(n))) < 1e-9
    assert abs(ld[0] - 2 * np.sum(np.log(np.diag(Lr)))) <= 1e-12 * max(1.0, abs(ld[0]))
    with pytest.raises(_lib.GprnError):
        B = -np.eye(n)
        _lib.check(_lib.lib().gprn_debug_factor(g._h(), n, _lib.dptr(_lib.f64(B)), _lib.dptr(L), None, None))


# ---------------------------------------------------------------------------------------------
# a5-a12: ELBO against the reference golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_elbo_matches_reference(name):
    d = load_golden(name)
    g = from_golden(d)
    elbo, mu, var, it = g.ELBOcalc(max_iter=d["max_iter"])
    assert it == d["iters"]
    assert abs(elbo - d["elbo"]) <= 1e-10 * abs(d["elbo"]), (elbo, d["elbo"])
    assert mu.shape == var.shape == (1 + g.p, g.q, g.N)
    assert rel(mu, d["mu"]) < 1e-8 and rel(var, d["var"]) < 1e-8
    # converged result is cached for mu='previous' and for prediction (meanfield.py:644-645)
    assert g._mu is mu and g._var is var


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "c3_synth_256_4_1_QP"])
def test_prediction_matches_reference(name):
    d = load_golden(name)
    g = from_golden(d)
    pm, pv, sep = g._Prediction(tstar=d["tstar"], mu=d["mu"], var=d["var"], separate=True)
    assert pm.shape == pv.shape == (d["tstar"].size, g.p)
    assert rel(pm, d["pred_mean"]) < 1e-8 and rel(pv, d["pred_var"]) < 1e-8
    assert rel(sep[0], d["node_pred"]) < 1e-8 and rel(sep[1], d["weight_pred"]) < 1e-8
    pm2, pv2 = g._Prediction(tstar=d["tstar"], mu=d["mu"], var=d["var"])
    assert np.array_equal(pm, pm2) and np.array_equal(pv, pv2)


def _big(name):
    path = os.path.join(GOLDEN, "big", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("large-N anchor not generated")
    return np.load(path)


def _big_inference(z):
    """The synth(N, 4, 2, 'M52') problem of a large-N anchor with the anchor's hyper-parameter set."""
    m = orc.synth(int(z["N"]), int(z["p"]), int(z["q"]), seed=1, node="M52")
    g = from_oracle_model(m)
    if "theta" in z.files:
        g.set_parameters(full_parameters(m, z["theta"])[0])
    return g


@pytest.mark.parametrize("name", ["c5_synth_2048_4_2_M52_it6", "c4_synth_4096_4_2_M52_it2"])
def test_large_n_anchor(name):
    """C4 / C5-size anchors: the reference capped at max_iter (minutes of CPU time, generated once)."""
    z = _big(name)
    g = _big_inference(z)
    elbo, mu, var, it = g.ELBOcalc(max_iter=int(z["max_iter"]))
    assert it == int(z["iters"])
    assert abs(elbo - float(z["elbo"])) <= 1e-10 * abs(float(z["elbo"])), (elbo, float(z["elbo"]))
    assert rel(mu[:, :, :16], z["mu_head"]) < 1e-8 and rel(var[:, :, :16], z["var_head"]) < 1e-8


@pytest.mark.parametrize("name", ["c5_synth_2048_4_2_M52_conv", "c4_pool102_set0_4096_4_2_M52_conv",
                                  "c4_synth_4096_4_2_M52_conv"])
def test_large_n_converged_anchor(name):
    """The headline configurations run to the reference's OWN stopping rule: C5-size theta_0
    (64 iterations), C4 theta_0 and set 0 of the bench pool (seed 102) -- the unmodified reference under the import
    shim, about an hour of host time each, generated once (tests/golden/make_golden.py --big-converged).
    Identical iteration count, ELBO to 1e-10, full variational state to 1e-8."""
    z = _big(name)
    g = _big_inference(z)
    elbo, mu, var, it = g.ELBOcalc()
    assert it == int(z["iters"]), (it, int(z["iters"]))
    assert abs(elbo - float(z["elbo"])) <= 1e-10 * abs(float(z["elbo"])), (elbo, float(z["elbo"]))
    assert rel(mu, z["mu"]) < 1e-8 and rel(var, z["var"]) < 1e-8
    # per-iteration trace of the reference: entry k is the ELBO after k iterations (entry 0 = the discarded pre-loop call)
    tr = z["trace"]
    e6, _, _, it6 = g.ELBOcalc(max_iter=6)
    assert it6 == 6 and abs(e6 - tr[6]) <= 1e-10 * abs(tr[6])
    g.close()


def test_prediction_at_20000_epochs_matches_reference_slice():
    """C5 prediction at T = 20000 (the chunked Tc = 4096 path): every 10th epoch was predicted by the unmodified
    reference (its O(T^2 N) loop makes the full grid infeasible; the predictive is per-epoch independent,
    _gp.py:131-137).  Mean and variance to 1e-8 on the slice; chunk boundaries are covered by the slice."""
    z = _big("c5_synth_2048_4_2_M52_conv")
    if "pred_mean" not in z.files:
        pytest.skip("prediction slice not generated")
    g = _big_inference(z)
    t = g.time
    span = t[-1] - t[0]
    tstar = np.linspace(t[0] - 0.2 * span, t[-1] + 0.2 * span, int(z["T_full"]))
    pm, pv, sep = g._Prediction(tstar=tstar, mu=z["mu"], var=z["var"], separate=True)
    sl = z["slice_idx"]
    assert pm.shape == (tstar.size, g.p) and np.all(np.isfinite(pm)) and np.all(pv > 0)
    assert rel(pm[sl], z["pred_mean"]) < 1e-8 and rel(pv[sl], z["pred_var"]) < 1e-8
    assert rel(sep[0][:, sl], z["node_pred"]) < 1e-8 and rel(sep[1][:, sl], z["weight_pred"]) < 1e-8
    # the slice alone (one chunk) gives the same numbers as the full grid (four chunks + a ragged tail)
    pm2, pv2 = g._Prediction(tstar=tstar[sl], mu=z["mu"], var=z["var"])
    assert np.array_equal(pm2, pm[sl]) and np.array_equal(pv2, pv[sl])
    g.close()


# ---------------------------------------------------------------------------------------------
# API behaviour of the loop
# ---------------------------------------------------------------------------------------------
def test_reference_smoke_elbo_property():
    """reference tests/test_inference.py:39-53: unseeded random data, jitter 0.0, just has to run."""
    t, y, yerr = np.random.rand(3, 10)
    g = gp.inference(1, t, y, yerr)
    g.set_components(covfunc.SquaredExponential(1, 1), covfunc.SquaredExponential(1, 1), meanfunc.Constant(0), 0.0)
    assert np.isfinite(g.ELBO)


def test_max_iter_semantics_and_init(capsys):
    d = load_golden("synth_100_4_1_QP")
    g = from_golden(d)
    m = orc.Model(d["t"], d["y"], d["yerr"], d["nodes"], d["weights"], None, d["jitters"])
    mu0, var0 = orc.init_mu_var(m)
    mu_d, var_d = g._initMuVar(g.nodes, g.weights, g.jitters)
    assert rel(mu_d, mu0) < 1e-15 and rel(var_d, var0) < 1e-15
    for cap in (0, 1, 3):
        e_o, mu_o, var_o, it_o = orc.elbo_calc(m, max_iter=cap)
        e_g, mu_g, var_g, it_g = g.ELBOcalc(max_iter=cap)
        assert it_g == it_o == cap
        assert abs(e_g - e_o) <= 1e-10 * abs(e_o)
        assert rel(mu_g.ravel(), np.asarray(mu_o).ravel()) < 1e-9
        assert 'Max iterations reached' in capsys.readouterr().out
    assert g._mu is None          # not cached unless converged (meanfield.py:648-649)
    # explicit initial state == 'previous' semantics: continuing from iteration 3 reproduces iterations 4..
    e3, mu3, var3, _ = g.ELBOcalc(max_iter=3)
    e_g, _, _, it_g = g.ELBOcalc(mu=mu3, var=var3, max_iter=1)
    e_o, *_ = orc.elbo_calc(m, max_iter=4)
    assert abs(e_g - e_o) <= 1e-10 * abs(e_o)
    one = g.ELBOaux(mu=mu3, var=var3)
    assert abs(one[0] - e_o) <= 1e-10 * abs(e_o) and one[3] is None


def test_previous_warm_start_and_nelbo(capsys):
    d = load_golden("synth_100_4_1_QP")
    g = from_golden(d)
    e1, mu1, var1, it1 = g.ELBOcalc()
    e2, _, _, it2 = g.ELBOcalc(mu='previous', var='previous')
    assert it2 <= it1
    m = orc.Model(d["t"], d["y"], d["yerr"], d["nodes"], d["weights"], None, d["jitters"])
    e_o, *_ , it_o = orc.elbo_calc(m, mu=mu1, var=var1)
    assert it2 == it_o and abs(e2 - e_o) <= 1e-10 * abs(e_o)
    val = g.nELBO(g.get_parameters())
    assert np.isfinite(val) and 'ELBO=' in capsys.readouterr().out
    with pytest.raises(ValueError):
        g.ELBOcalc(mu='bogus', var='bogus')


def test_optimize_and_predict_drivers(capsys):
    """The host drivers above the path (reference meanfield.py:1114-1152 optimize, :1381-1400 predict) run unchanged
    against the device path: Nelder-Mead over one free parameter improves the ELBO; predict() returns the 4-tuple."""
    d = load_golden("synth_50_1_1_QP")
    g = from_golden(d)
    e0 = g.ELBO
    res = g.optimize(vars='node1.P', options={'maxfev': 12, 'xatol': 1e-3})
    assert res.x.shape == (1,) and g.frozen_mask.sum() == g.n_parameters - 1
    assert -res.fun >= e0 - 1e-9 * abs(e0)
    assert g.nodes[0].pars[2] == res.x[0]
    tstar, mean, std, sep = g.predict(nn=37)
    assert tstar.shape == (37,) and mean.shape == std.shape == (37, 1) and np.all(np.isfinite(std))
    assert sep[0].shape == (1, 37) and sep[1].shape == (1, 37)
    capsys.readouterr()


def test_not_positive_definite_reports_nan():
    t = np.linspace(0, 1, 20)
    y, e = np.sin(t), np.full(20, 0.1)
    g = gp.inference(1, t, y, e)
    # negative-definite "kernel": theta^2 * exp(..) is fine, so break it with a huge periodic sum cancelling
    g.set_components(covfunc.WhiteNoise(0.0), covfunc.SquaredExponential(1, 1), meanfunc.Constant(0), 0.1)
    # K_node = 0*I + 1e-6 I is PD; make A = K + D not PD by a NaN jitter instead
    g.jitters = np.array([np.nan])
    elbo, mu, var, it = g.ELBOcalc()
    assert np.isnan(elbo) and it <= 1


# ---------------------------------------------------------------------------------------------
# batched evaluation (ELBO_batch) against the oracle on seeded inputs, and properties at full size
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 2, 1, "QP", 12), (100, 3, 2, "M52", 6), (130, 1, 1, "QP", 8),
                                   (300, 2, 1, "QP", 3)])
def test_batch_matches_oracle(shape):
    N, p, q, node, B = shape
    m = orc.synth(N, p, q, seed=3, node=node)
    theta = orc.perturbed_hyper_sets(m, B, 11)
    g = from_oracle_model(m)
    elbo, iters, status = g.ELBO_batch(full_parameters(m, theta), return_info=True)
    assert elbo.shape == (B,)
    compared = 0
    for b in range(B):
        try:
            e, _, _, it = orc.elbo_calc(orc.model_with_hyper(m, theta[b]))
        except np.linalg.LinAlgError:
            continue                       # reference's explicit-Sigma algebra lost positive definiteness
        if not np.isfinite(e):
            continue
        assert status[b] == 0 and iters[b] == it
        assert abs(elbo[b] - e) <= 1e-10 * abs(e), (b, elbo[b], e)
        compared += 1
    assert compared >= 0.75 * B, f"only {compared} of {B} sets could be compared with the oracle"


@pytest.mark.parametrize("shape", [(60, 2, 3, "M52", 3), (50, 3, 4, "M52", 3), (40, 2, 3, "QP", 4)])
def test_more_than_two_nodes_capped(shape):
    """q = 3, 4: three / six cross-node trace pairs (quirk Q3) and the reshape pairing (Q4) beyond q = 2.  The
    reference's fixed point diverges for these inputs and its explicit-Sigma Cholesky fails after a few more
    iterations, so the comparison is capped at the first iterations (arithmetic parity, 1e-10 on the ELBO)."""
    N, p, q, node, cap = shape
    m = orc.synth(N, p, q, seed=3, node=node)
    g = from_oracle_model(m)
    e_o, mu_o, var_o, it_o = orc.elbo_calc(m, max_iter=cap)
    e_g, mu_g, var_g, it_g = g.ELBOcalc(max_iter=cap)
    assert it_g == it_o == cap
    assert abs(e_g - e_o) <= 1e-10 * abs(e_o), (e_g, e_o)
    assert rel(mu_g, mu_o) < 1e-8 and rel(var_g, var_o) < 1e-8


def test_batch_with_per_set_means():
    m = orc.synth(64, 2, 1, seed=4, node="QP")
    g = inference_from(m.time, m.y, m.yerr, m.nodes, m.weights, [0.0, 0.0], m.jitters)
    base = g.get_parameters()
    P = np.tile(base, (3, 1))
    P[1, -4:-2] = [0.3, -0.2]          # mean constants of set 1
    P[2, 0] *= 1.1
    elbo = g.ELBO_batch(P)
    for b in range(3):
        nk = P.shape[1] - 4
        mb = orc.model_with_hyper(m, np.r_[P[b, :nk], P[b, -2:]])
        mb.mean_vals = np.repeat(P[b, nk:nk + 2][:, None], m.N, axis=1)
        e, *_ = orc.elbo_calc(mb)
        assert abs(elbo[b] - e) <= 1e-10 * abs(e)


def test_batch_warm_start_state_roundtrip():
    """Row (f.1): batched warm start.  Stopping a batch after 3 iterations and resuming it from the returned
    state must land on the same fixed point as the oracle resumed from the same state (meanfield.py:598-607)."""
    m = orc.synth(64, 2, 1, seed=6, node="QP")
    theta = orc.perturbed_hyper_sets(m, 5, 21)
    g = from_oracle_model(m)
    P = full_parameters(m, theta)
    e3, it3, st3, mu3, var3 = g.ELBO_batch(P, max_iter=3, return_info=True, return_state=True)
    assert np.all(it3 == 3) and np.all(st3 == 2) and mu3.shape == (5, g.d)
    e, it, st, mu, var = g.ELBO_batch(P, return_info=True, mu=mu3, var=var3, return_state=True)
    for b in range(5):
        mb = orc.model_with_hyper(m, theta[b])
        _, mu_o, var_o, _ = orc.elbo_calc(mb, max_iter=3)
        e_o, mu_f, var_f, it_o = orc.elbo_calc(mb, mu=mu_o, var=var_o)
        assert it[b] == it_o and st[b] == 0
        assert abs(e[b] - e_o) <= 1e-10 * abs(e_o)
        assert rel(mu[b], np.asarray(mu_f).ravel()) < 1e-8
    with pytest.raises(ValueError):
        g.ELBO_batch(P, mu=mu3)


def test_c3_full_size_properties():
    """C3 (N=256, p=4, q=1) at a large batch: every set is independent, so the batch result must equal the
    single evaluation bit for bit, be invariant under permutation of the sets, and reproduce the
    reference anchor for theta_0."""
    m = orc.synth(256, 4, 1, seed=1, node="QP")
    B = 1024
    theta = orc.perturbed_hyper_sets(m, B, 101)
    theta[0] = orc.perturbed_hyper_sets(m, 1, 0)[0] * 0 + np.r_[sum((orc.spec_params(s) for s in m.nodes + m.weights), []), m.jitters]
    g = from_oracle_model(m)
    P = full_parameters(m, theta)
    elbo, iters, status = g.ELBO_batch(P, return_info=True)
    d = load_golden("c3_synth_256_4_1_QP")
    assert iters[0] == d["iters"] and abs(elbo[0] - d["elbo"]) <= 1e-10 * abs(d["elbo"])
    assert np.all(status == 0) and np.all(np.isfinite(elbo)) and iters.min() >= 4
    perm = np.random.default_rng(0).permutation(B)
    elbo_p, iters_p, _ = g.ELBO_batch(P[perm], return_info=True)
    assert np.array_equal(elbo_p, elbo[perm]) and np.array_equal(iters_p, iters[perm])
    for b in (1, 17, 1023):
        g.set_parameters(P[b])
        e1, _, _, it1 = g.ELBOcalc()
        assert e1 == elbo[b] and it1 == iters[b]
    # oracle spot check on a few perturbed sets
    for b in (1, 2, 3):
        e, _, _, it = orc.elbo_calc(orc.model_with_hyper(m, theta[b]))
        assert iters[b] == it and abs(elbo[b] - e) <= 1e-10 * abs(e)


def test_workspace_chunking_is_transparent():
    m = orc.synth(128, 2, 1, seed=2, node="QP")
    theta = orc.perturbed_hyper_sets(m, 40, 5)
    g = from_oracle_model(m)
    P = full_parameters(m, theta)
    ref = g.ELBO_batch(P)
    _lib.check(_lib.lib().gprn_set_workspace_limit(g._h(), 7 * 3 * 3 * 128 * 128 * 8 + (1 << 20)))   # ~7 sets per chunk
    got = g.ELBO_batch(P)
    assert np.array_equal(ref, got)


# ---------------------------------------------------------------------------------------------
# rows (f.2), (f.3): prediction over a chain of hyper sets, prior draws, Derivative kernels
# ---------------------------------------------------------------------------------------------
def test_prediction_batch_matches_oracle_per_set():
    """_Prediction over B hyper sets (meanfield.py:1289-1379 applied to a chain): each set equals the oracle run
    on that set alone -- converged state, then predictive mean / variance to 1e-8."""
    m = orc.synth(64, 2, 1, seed=8, node="QP")
    B = 4
    theta = orc.perturbed_hyper_sets(m, B, 31)
    g = from_oracle_model(m)
    before = g.get_parameters().copy()
    tstar = np.linspace(m.time[0] - 3, m.time[-1] + 5, 37)
    pm, pv, elbo, iters, status = g.Prediction_batch(full_parameters(m, theta), tstar=tstar, return_info=True)
    assert pm.shape == pv.shape == (B, 37, m.p)
    assert np.array_equal(g.get_parameters(), before)
    for b in range(B):
        mb = orc.model_with_hyper(m, theta[b])
        e, mu, var, it = orc.elbo_calc(mb)
        assert iters[b] == it and abs(elbo[b] - e) <= 1e-10 * abs(e)
        om, ov, _, _ = orc.prediction(mb, tstar, mu, var, np.repeat(m.mean_vals[:, :1], 37, axis=1))
        assert rel(pm[b], om) < 1e-8 and rel(pv[b], ov) < 1e-8


@pytest.mark.parametrize("N", [50, 300])
def test_prior_draws_are_cholesky_times_z(N):
    """sample() (meanfield.py:517-539): draws are chol(K + nugget I) z; checked against numpy's factor of the
    oracle's kernel matrices for the same z, and for the right covariance through unit vectors."""
    m = orc.synth(N, 2, 1, seed=12, node="QP")
    m.nodes = [("sum", m.nodes[0], ("WN", 0.05))]
    m.weights = [("sum", w, ("WN", 0.05)) for w in m.weights]
    g = from_oracle_model(m)
    z = np.random.default_rng(5).standard_normal((3, N))
    ns, ws = g.sample(z=z)
    assert ns.shape == (1, N) and ws.shape == (2, N)
    for k, (spec, got) in enumerate(zip(m.nodes + m.weights, np.vstack([ns, ws]))):
        L = np.linalg.cholesky(orc.kmatrix(spec, m.time, nugget=1.25e-12))
        assert rel(got, L @ z[k]) < 1e-10
    e3 = np.zeros((3, N)); e3[:, 3] = 1.0
    ns, _ = g.sample(z=e3)                 # column 3 of L
    K = orc.kmatrix(m.nodes[0], m.time, nugget=1.25e-12)
    assert rel(ns[0], np.linalg.cholesky(K)[:, 3]) < 1e-10
    one = g._sample_from_gp(covfunc.SquaredExponential(1.0, 5.0) + covfunc.WhiteNoise(0.1), z=z[0])
    Lse = np.linalg.cholesky(orc.kmatrix(("sum", ("SE", 1.0, 5.0), ("WN", 0.1)), m.time, nugget=1.25e-12))
    assert rel(one, Lse @ z[0]) < 1e-10
    np.random.seed(0)
    a, b = g.sample()                      # default: numpy's global generator
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))


def test_prior_draw_of_singular_kernel_raises():
    m = orc.synth(200, 1, 1, seed=2, node="QP")
    g = inference_from(m.time, m.y, m.yerr, [("SE", 1.0, 500.0)], m.weights, [0.0], m.jitters)
    with pytest.raises(_lib.GprnError):
        g.sample(nugget=-2.0)              # first pivot 1 - 2 < 0: deterministic "not positive definite"


# ---------------------------------------------------------------------------------------------
# continuous batching, dynamic work source, device-resident chain state, lock-step multi-start optimisation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 2, 1, "QP", 23), (300, 2, 2, "M52", 9), (600, 1, 2, "M52", 5)])
def test_continuous_batching_is_transparent(shape):
    """A pool evaluated with few workspace slots (sets retire at their own iteration count and the freed slot is
    refilled in the next lock-step round) gives bit-identical results to evaluating it all at once, in any slot
    count, and a set evaluated alone gives the same bits again: an evaluation never depends on its batch mates."""
    N, p, q, node, B = shape
    m = orc.synth(N, p, q, seed=2, node=node)
    theta = orc.perturbed_hyper_sets(m, B, 5)
    g = from_oracle_model(m)
    P = full_parameters(m, theta)
    ref, it_ref, st_ref = g.ELBO_batch(P, return_info=True)
    if shape[0] == 128:
        assert len(set(it_ref.tolist())) > 1               # the sets do retire at different rounds
    for slots in (1, 3, B - 1):
        got, it, st = g.ELBO_batch(P, return_info=True, slots=slots)
        assert np.array_equal(ref, got) and np.array_equal(it_ref, it) and np.array_equal(st_ref, st), slots
    g.set_parameters(P[B - 2])
    e1, _, _, it1 = g.ELBOcalc()
    assert e1 == ref[B - 2] and it1 == it_ref[B - 2]
    g.close()


def test_work_source_subset_and_order():
    """gprn_elbo_pool with an external work source: only the indices it hands out are evaluated (others stay 0 /
    not taken), in whatever order, with the same bits as the plain batch."""
    m = orc.synth(100, 2, 1, seed=3, node="QP")
    theta = orc.perturbed_hyper_sets(m, 12, 9)
    g = from_oracle_model(m)
    P = full_parameters(m, theta)
    ref, it_ref, _ = g.ELBO_batch(P, return_info=True)
    order = [7, 2, 11, 0, 5]
    src = iter(order)
    e, it, st, taken = g.ELBO_batch(P, return_info=True, slots=2, work_source=lambda: next(src, -1))
    assert sorted(np.flatnonzero(taken).tolist()) == sorted(order)
    assert np.array_equal(e[order], ref[order]) and np.array_equal(it[order], it_ref[order])
    rest = np.setdiff1d(np.arange(12), order)
    assert np.all(e[rest] == 0) and np.all(it[rest] == 0)
    with pytest.raises(_lib.GprnError):
        g.ELBO_batch(P, work_source=iter([99]).__next__)      # index outside the pool
    g.close()


def test_device_resident_chain_state_matches_sequential_warm_starts():
    """state='previous' (f.1): every row keeps its own variational state on the device; a sequence of calls equals
    the oracle run per chain with the reference's caching rule -- start from the chain's last CONVERGED state,
    keep it unchanged when an evaluation hits max_iter (meanfield.py:598-607, 643-649)."""
    m = orc.synth(80, 2, 1, seed=7, node="QP")
    B = 4
    rng = np.random.default_rng(3)
    g = from_oracle_model(m)
    base = orc.perturbed_hyper_sets(m, B, 41)
    chain_mu, chain_var = [None] * B, [None] * B
    for call, cap in enumerate([None, 2, None, None]):       # the second call cannot converge: state must survive it
        theta = base * np.exp(0.02 * rng.standard_normal(base.shape))
        P = full_parameters(m, theta)
        e, it, st = g.ELBO_batch(P, max_iter=cap, return_info=True, state='previous')
        for b in range(B):
            mb = orc.model_with_hyper(m, theta[b])
            e_o, mu_o, var_o, it_o = orc.elbo_calc(mb, max_iter=cap, mu=chain_mu[b], var=chain_var[b])
            converged = cap is None or it_o < cap
            assert it[b] == it_o and st[b] == (0 if converged else 2), (call, b)
            assert abs(e[b] - e_o) <= 1e-10 * abs(e_o), (call, b, e[b], e_o)
            if converged:
                chain_mu[b], chain_var[b] = mu_o, var_o
    mu_d, var_d, valid = g.get_chain_state()
    assert valid.all() and rel(mu_d[1], np.asarray(chain_mu[1]).ravel()) < 1e-8
    g.reset_chain_state()
    e0 = g.ELBO_batch(P, state='previous')
    assert np.array_equal(e0, g.ELBO_batch(P))               # after a reset every chain starts from 'init' again
    g.close()


def test_optimize_batch_equals_sequential_scipy_runs(capsys):
    """C5 driver at small N: S Nelder-Mead starts advanced in lock-step on the GPU return what S
    sequential ``scipy.optimize.minimize(nELBO)`` runs on the CPU oracle return -- the same simplex decisions, hence
    the same evaluation counts and optima (objective values agree to 1e-10, so no comparison flips)."""
    from scipy.optimize import minimize
    m = orc.synth(48, 2, 1, seed=9, node="QP")
    g = from_oracle_model(m)
    g.freeze_parameter(name='*')
    for name in ('node1.theta', 'node1.le', 'weight1.ell'):
        g.thaw_parameter(name=name)
    names = list(g.parameters_dict.keys())
    free = [names.index(n) for n in ('node1.theta', 'node1.le', 'weight1.ell')]
    x_full = g.get_parameters(include_frozen=True)
    S = 5
    starts = x_full[free] * np.exp(0.15 * np.random.default_rng(2).standard_normal((S, len(free))))
    opts = {"maxfev": 25, "xatol": 1e-6, "fatol": 1e-9}
    res = g.optimize_batch(starts, options=opts)
    assert len(res) == S and g.n_batch_calls <= max(r.nfev for r in res) + 1
    n_kernel = x_full.size - 2 * m.p        # [kernel pars, mean consts, jitters]
    for s in range(S):
        state = {"mu": None, "var": None}

        def nelbo(x):                        # the reference's nELBO on the oracle: warm start from the last converged state
            full = x_full.copy()
            full[free] = x
            mb = orc.model_with_hyper(m, np.r_[full[:n_kernel], full[-m.p:]])
            e, mu, var, it = orc.elbo_calc(mb, mu=state["mu"], var=state["var"])
            if it < 10000:
                state["mu"], state["var"] = mu, var
            return -e

        ref = minimize(nelbo, starts[s], method='Nelder-Mead', options=opts)
        assert res[s].nfev == ref.nfev and res[s].nit == ref.nit, (s, res[s].nfev, ref.nfev)
        assert np.allclose(res[s].x, ref.x, rtol=1e-9, atol=0) and abs(res[s].fun - ref.fun) <= 1e-9 * abs(ref.fun)
    best = min(range(S), key=lambda s: res[s].fun)
    assert np.array_equal(g.get_parameters(), res[best].x)
    capsys.readouterr()
    g.close()


def test_logposterior_batch():
    """Vectorised log-posterior (emcee vectorize=True contract; reference logposterior meanfield.py:1214-1219): prior
    + ELBO capped at 100 iterations; rows outside the prior support are -inf and are not evaluated."""
    from scipy import stats
    m = orc.synth(60, 2, 1, seed=4, node="QP")
    g = from_oracle_model(m)
    g.freeze_parameter(name='*')
    g.thaw_parameter(name='node1.theta')
    g.thaw_parameter(name='jitter1')
    priors = {'node1.theta': stats.uniform(0.1, 5.0), 'jitter1': stats.uniform(0.01, 1.0)}
    thetas = np.array([[1.0, 0.1], [1.3, 0.2], [9.0, 0.1], [0.7, 0.05]])     # row 2 is outside the prior
    total, elbo = g.logposterior_batch(thetas, priors)
    assert np.isneginf(total[2]) and np.isneginf(elbo[2])
    x_full = g.get_parameters(include_frozen=True)
    names = list(g.parameters_dict.keys())
    idx = [names.index('node1.theta'), names.index('jitter1')]
    nk = x_full.size - 2 * m.p
    for b in (0, 1, 3):
        full = x_full.copy()
        full[idx] = thetas[b]
        e_o, *_ = orc.elbo_calc(orc.model_with_hyper(m, np.r_[full[:nk], full[-m.p:]]), max_iter=100)
        lp = sum(priors[n].logpdf(v) for n, v in zip(('node1.theta', 'jitter1'), thetas[b]))
        assert abs(elbo[b] - e_o) <= 1e-10 * abs(e_o) and abs(total[b] - (lp + e_o)) <= 1e-10 * abs(e_o)
    g.close()


def test_batch_parameter_width_validation():
    """a wrong-width parameter matrix raises the ValueError of set_parameters; frozen columns of a
    full-width matrix are held at their current values."""
    m = orc.synth(40, 2, 1, seed=4, node="QP")
    g = from_oracle_model(m)
    P = np.tile(g.get_parameters(), (2, 1))
    with pytest.raises(ValueError, match='Wrong number of parameters'):
        g.ELBO_batch(P[:, :-1])
    ref = g.ELBO_batch(P)
    g.freeze_parameter(name='jitter1')
    P2 = P.copy()
    P2[:, list(g.parameters_dict.keys()).index('jitter1')] = 5.0          # frozen: must be ignored
    assert np.array_equal(g.ELBO_batch(P2), ref)
    assert np.array_equal(g.ELBO_batch(P[:, ~g.frozen_mask]), ref)
    g.close()


# ---------------------------------------------------------------------------------------------
# hygiene: ticket stress self-check, handles on two devices, chunked prediction with the WhiteNoise quirk
# ---------------------------------------------------------------------------------------------
def test_panel_ticket_stress_bit_identical():
    """Stand-in for racecheck (closed on this pool): the one-launch panel step (last-reader ticket, no fence) against
    the two-launch path, 64 matrices x 40 repetitions x 8 panel steps, bit for bit."""
    n = 512
    rng = np.random.default_rng(0)
    tt = np.sort(rng.uniform(0, 900, n))
    A = orc.kmatrix(("M52", 1.0, 30.0), tt, nugget=1e-6) + np.diag(rng.uniform(0.01, 1.0, n))
    g = gp.inference(1, np.arange(4.0), np.zeros(4), np.ones(4))
    import ctypes
    mis = ctypes.c_int64(-1)
    _lib.check(_lib.lib().gprn_debug_panel_stress(g._h(), n, _lib.dptr(_lib.f64(A)), 64, 40, ctypes.byref(mis)))
    assert mis.value == 0
    g.close()


def test_handles_on_two_devices_in_one_process():
    """the > 48 KB shared-memory opt-in is per device; a second handle on another GPU of the
    same process must launch the large-smem kernels too."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d = load_golden("synth_256_4_2_M52")
    vals = []
    for dev in (0, 1):
        args = []
        for y, e in zip(d["y"], d["yerr"]):
            args += [y, e]
        g = gp.inference(len(d["nodes"]), d["t"], *args, device=dev)
        g.set_components([build_kernel(s) for s in d["nodes"]], [build_kernel(s) for s in d["weights"]],
                         [meanfunc.Constant(c) for c in d["mean_consts"]], list(d["jitters"]))
        vals.append(g.ELBOcalc()[0])
        k = covfunc.SquaredExponential(1.0, 3.0)
        assert np.allclose(k(np.array([0.0, 1.0])), [1.0, np.exp(-0.5 / 9.0)])      # k(r) on the current device
        g.close()
    assert vals[0] == vals[1] and abs(vals[0] - d["elbo"]) <= 1e-10 * abs(d["elbo"])


def test_chunked_prediction_with_whitenoise_square_quirk(monkeypatch):
    """T == N with a WhiteNoise term (quirk Q9: Kstar gets w^2 on its diagonal BY POSITION) through the chunked
    prediction path: the chunk offset must enter the position test (round 1 refused T == N > 4096)."""
    m = orc.synth(150, 2, 1, seed=12, node="QP")
    m.nodes = [("sum", m.nodes[0], ("WN", 0.3))]
    m.weights = [("sum", w, ("WN", 0.2)) for w in m.weights]
    g = from_oracle_model(m)
    e, mu, var, it = g.ELBOcalc()
    mean_t = np.repeat(m.mean_vals[:, :1], m.N, axis=1)
    om, ov, _, _ = orc.prediction(m, m.time, mu, var, mean_t)
    pm, pv = g._Prediction(tstar=m.time, mu=mu, var=var)
    assert rel(pm, om) < 1e-8 and rel(pv, ov) < 1e-8
    monkeypatch.setenv("GPRN_PREDICT_TC", "64")               # three chunks
    pm2, pv2 = g._Prediction(tstar=m.time, mu=mu, var=var)
    assert rel(pm2, om) < 1e-8 and rel(pv2, ov) < 1e-8
    g.close()


@pytest.mark.parametrize("p,tol_oracle", [(2, 1e-10), (1, 2e-8)])
def test_mid_n_fused_path_matches_multi_kernel_path_and_oracle(p, tol_oracle):
    """256 < N <= 512, q = 1 with enough matrices in flight runs the fused single-kernel pipeline (small.cuh, nt up to
    8); a single evaluation of the same set runs the multi-kernel path.  The two device paths must agree to 1e-10 with
    identical iteration counts, and with the oracle to 1e-10 (p = 2: 5 iterations, |ELBO| ~ 2e3).
    The p = 1 case is the ill-conditioned one: 110-130 iterations and an ELBO of 4 ... 50 left over from terms of
    order 1e3, for which the ORACLE ITSELF moves by 3e-9 between two hosts (set 0: 3.9144956393 in the build container,
    3.9144956423 on the GPU box, different BLAS kernels); there the bar against the oracle is 2e-8 and the device
    paths are still held to 1e-10 against each other."""
    m = orc.synth(330, p, 1, seed=5, node="QP")           # Np = 384: 6 x 6 tiles
    B = 160                                                # 160 sets x (p + 1) matrices >= 2 x 148 SMs: fused path
    theta = orc.perturbed_hyper_sets(m, B, 13)
    g = from_oracle_model(m)
    P = full_parameters(m, theta)
    elbo, iters, status = g.ELBO_batch(P, return_info=True)
    assert np.all(status == 0) and np.all(np.isfinite(elbo))
    for b in (0, 7, B - 1):
        g.set_parameters(P[b])
        e1, _, _, it1 = g.ELBOcalc()                        # B = 1: multi-kernel path
        assert it1 == iters[b] and abs(e1 - elbo[b]) <= 1e-10 * abs(e1), (b, e1, elbo[b])
        e_o, _, _, it_o = orc.elbo_calc(orc.model_with_hyper(m, theta[b]))
        assert it_o == iters[b] and abs(elbo[b] - e_o) <= tol_oracle * abs(e_o), (b, elbo[b], e_o)
    # continuous batching through the fused path: refilled slots give the same bits
    got = g.ELBO_batch(P, slots=150)
    assert np.array_equal(got, elbo)
    g.close()
