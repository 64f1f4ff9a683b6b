"""CPU oracle for the mean-field GPRN hot path (TEST INFRASTRUCTURE -- not part of the product).

This module restates, in plain numpy/scipy, the arithmetic of the reference's
``gpyrn/meanfield.py`` ELBO evaluation and ``gpyrn/_gp.py`` GP prediction, call for call
(same LAPACK routines: LU ``solve`` for the Woodbury step, ``potrf`` for the Choleskys,
``potrs`` for the traces), so that it can serve both as the parity checker for the CUDA
path and as the timed CPU baseline (``cpu_baseline.kind == "port"`` in bench.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product (``gpyrn_b200``) never does.

Parity pinning: the reference's own tests hold NO golden vector for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the unmodified
reference executed in the build container under an import shim (``tests/golden/make_golden.py``
-> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

All citations ``file:line`` are relative to the reference checkout (``/root/reference``).

Kernel specification used throughout (a tiny expression tree, no reference classes needed):
    ("SE", theta, ell) | ("P", theta, P, ell) | ("QP", theta, elle, P, ellp) |
    ("RQ", theta, alpha, ell) | ("M32", theta, ell) | ("M52", theta, ell) | ("WN", w) |
    ("C", c) | ("RQP", theta, alpha, elle, P, ellp) | ("COS", theta, P) | ("EXP", theta, ell) |
    ("dSE", theta, ell) | ("dP", theta, P, ell) | ("dQP", theta, elle, P, ellp)   [Derivative(k)] |
    ("GammaExp", theta, gamma, ell) | ("PW", eta) | ("PAC", amp, ell_1, ell_2) | ("NP", amp, alpha2, P, ell) |
    ("QNP", amp, alpha2, ell_e, P, ell_p) | ("CP", amp, P, ell) | ("QCP", amp, ell_e, P, ell_p) |
    ("sum", spec1, spec2) | ("mul", spec1, spec2)
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

LOG2PI = np.log(2.0 * np.pi)


# --------------------------------------------------------------------------------------
# a1: covariance functions k(r)                                  gpyrn/covfunc.py
# --------------------------------------------------------------------------------------
def kernel_eval(spec, r):
    """Evaluate a kernel spec on an array of lags ``r`` with the reference's operation order."""
    tag = spec[0]
    if tag == "sum":                      # covfunc.py:65-68
        return kernel_eval(spec[1], r) + kernel_eval(spec[2], r)
    if tag == "mul":                      # covfunc.py:74-77
        return kernel_eval(spec[1], r) * kernel_eval(spec[2], r)
    a = [float(v) for v in spec[1:]]
    if tag == "SE":                       # covfunc.py:169-170
        return a[0] ** 2 * np.exp(-0.5 * r ** 2 / a[1] ** 2)
    if tag == "P":                        # covfunc.py:211-213
        return a[0] ** 2 * np.exp(-2 * np.sin(np.pi * np.abs(r) / a[1]) ** 2 / a[2] ** 2)
    if tag == "QP":                       # covfunc.py:251-255
        t1 = -2 * np.sin(np.pi * np.abs(r) / a[2]) ** 2 / a[3] ** 2
        t2 = r ** 2 / (2 * a[1] ** 2)
        return a[0] ** 2 * np.exp(t1 - t2)
    if tag == "RQ":                       # covfunc.py:286-288
        return a[0] ** 2 * (1 + 0.5 * r ** 2 / (a[1] * a[2] ** 2)) ** (-a[1])
    if tag == "M32":                      # covfunc.py:370-373
        s = np.sqrt(3.0) * np.abs(r) / a[1]
        return a[0] ** 2 * (1.0 + s) * np.exp(-s)
    if tag == "M52":                      # covfunc.py:391-396
        ar = np.abs(r)
        return a[0] ** 2 * (1.0 + (3 * np.sqrt(5) * a[1] * ar + 5 * ar ** 2) / (3 * a[1] ** 2)) \
            * np.exp(-np.sqrt(5.0) * ar / a[1])
    if tag == "C":                        # covfunc.py:122-125
        return np.full_like(r, a[0] ** 2)
    if tag == "RQP":                      # covfunc.py:307-310
        return a[0] ** 2 * np.exp(-2 * np.sin(np.pi * np.abs(r) / a[3]) ** 2 / a[4] ** 2) \
            * (1 + r ** 2 / (2 * a[1] * a[2] ** 2)) ** (-a[1])
    if tag == "COS":                      # covfunc.py:327-328
        return a[0] ** 2 * np.cos(2 * np.pi * np.abs(r) / a[1])
    if tag == "EXP":                      # covfunc.py:351-352
        return a[0] ** 2 * np.exp(-np.abs(r) / a[1])
    if tag == "dSE":                      # covfunc.py:98-100 (Derivative) -> :182-185
        term1 = a[0] ** 2 / a[1] ** 4
        term2 = a[1] ** 2 - r ** 2
        return term1 * term2 * np.exp(-0.5 * r ** 2 / a[1] ** 2)
    if tag == "dP":                       # covfunc.py:215-221
        rP = np.pi * r / a[1]
        term1 = 4 * np.pi ** 2 * a[0] ** 2
        term2 = a[2] ** 2 * np.cos(2 * rP) - 4 * np.sin(rP) ** 2 * np.cos(rP) ** 2
        term3 = np.exp(-2 * np.sin(rP) ** 2 / a[2] ** 2)
        return term1 * term2 * term3
    if tag == "dQP":                      # covfunc.py:257-266
        th, le, P, lp = a
        term1 = 2 * th ** 2 / (P ** 2 * lp ** 4 * le ** 4)
        term2 = P ** 2 * lp ** 4 * le ** 2 - \
            2 * P ** 2 * lp ** 4 * r ** 2 - \
            4 * np.pi * P * lp ** 2 * le ** 2 * r * np.sin(2 * np.pi * r / P) + \
            2 * np.pi ** 2 * lp ** 2 * le ** 4 * np.cos(2 * np.pi * r / P) - \
            8 * np.pi ** 2 * le ** 4 * np.sin(np.pi * r / P) ** 2 * np.cos(np.pi * r / P) ** 2
        term3 = np.exp(-(lp ** 2 * r ** 2 + 2 * le ** 2 * np.sin(np.pi * r / P) ** 2) / (lp ** 2 * le ** 2))
        return term1 * term2 * term3
    if tag == "GammaExp":                 # covfunc.py:431-432
        return a[0] ** 2 * np.exp(-(np.abs(r) / a[2]) ** a[1])
    if tag == "PW":                       # covfunc.py:470-474
        x = r / (0.5 * a[0])
        pw = (3 * np.abs(x) + 1) * (1 - np.abs(x)) ** 3
        return np.where(np.abs(x) > 1, 0, pw)
    if tag == "PAC":                      # covfunc.py:493-496
        aa = np.sqrt(2 * a[1] * a[2] / (a[1] ** 2 + a[2] ** 2))
        bb = np.exp(-2 * r * r / (a[1] ** 2 + a[2] ** 2))
        return a[0] ** 2 * aa * bb
    if tag == "NP":                       # covfunc.py:517-519
        aa = (1 + 2 * np.sin(np.pi * np.abs(r) / a[2]) ** 2 / (a[1] * a[3] ** 2)) ** (-a[1])
        return a[0] ** 2 * aa
    if tag == "QNP":                      # covfunc.py:543-546
        aa = (1 + 2 * np.sin(np.pi * np.abs(r) / a[3]) ** 2 / (a[1] * a[4] ** 2)) ** (-a[1])
        bb = np.exp(-0.5 * r ** 2 / a[2] ** 2)
        return a[0] ** 2 * aa * bb
    if tag == "CP":                       # covfunc.py:664-665
        return a[0] ** 2 * np.exp(-2 * np.cos(np.pi * np.abs(r) / a[1]) ** 2 / a[2] ** 2)
    if tag == "QCP":                      # covfunc.py:686-688
        return a[0] ** 2 * np.exp(- 2 * np.cos(np.pi * np.abs(r) / a[2]) ** 2 / a[3] ** 2 - r ** 2 / (2 * a[1] ** 2))
    if tag == "WN":                       # covfunc.py:144-148 (quirk Q9: decided by shape, not by r==0)
        if r.ndim == 2 and r.shape[0] == r.shape[1]:
            return a[0] ** 2 * np.eye(r.shape[0])
        return np.full_like(r, a[0] ** 2)
    raise ValueError(f"unknown kernel tag {tag!r}")


def kernel_amplitude(spec):
    """``kernel.pars[0]`` of the reference object: first leaf parameter (covfunc.py:56-62)."""
    while spec[0] in ("sum", "mul"):
        spec = spec[1]
    return float(spec[1])


def kmatrix(spec, t_rows, t_cols=None, nugget=0.0):
    """a2: ``_KMatrix`` (meanfield.py:432-433, nugget 1e-6), ``_gp._kernel_matrix`` (_gp.py:48-49,
    nugget 1.25e-12) and ``_predict_kernel_matrix`` (_gp.py:60-61, no nugget)."""
    t_cols = t_rows if t_cols is None else t_cols
    r = t_rows[:, None] - t_cols[None, :]
    K = kernel_eval(spec, r)
    if nugget:
        K = K + nugget * np.eye(t_rows.size)
    return K


# --------------------------------------------------------------------------------------
# the model container (plain data; the host API of the product has its own classes)
# --------------------------------------------------------------------------------------
class Model:
    """time[N], y[p,N] raw, yerr[p,N], q node specs, q*p weight specs (index j*p+i),
    mean values m[p,N] already evaluated on ``time`` (meanfield.py:623), p jitters."""

    def __init__(self, time, y, yerr, nodes, weights, mean_vals, jitters):
        self.time = np.asarray(time, float)
        self.y = np.atleast_2d(np.asarray(y, float))
        self.yerr = np.atleast_2d(np.asarray(yerr, float))
        self.p, self.N = self.y.shape
        self.nodes = list(nodes)
        self.q = len(self.nodes)
        self.weights = list(weights)
        assert len(self.weights) == self.q * self.p
        self.mean_vals = np.zeros_like(self.y) if mean_vals is None else \
            np.broadcast_to(np.asarray(mean_vals, float), self.y.shape).copy()
        self.jitters = np.asarray(jitters, float).reshape(self.p)
        self.yerr2 = self.yerr ** 2                     # meanfield.py:127
        self.d = self.N * self.q * (self.p + 1)          # meanfield.py:121


# --------------------------------------------------------------------------------------
# a5: _initMuVar                                                meanfield.py:491-510
# --------------------------------------------------------------------------------------
def init_mu_var(m: Model):
    a1 = [kernel_amplitude(s) for s in m.nodes]
    a2 = [kernel_amplitude(s) for s in m.weights]
    mean1, mean2, var1, var2 = [], [], [], []
    for amp in a1:
        # zip(a2, y) stops after p entries: only the first p weight amplitudes are used (Q5)
        per_out = [np.sqrt(np.abs(yi) * amp / wi) * np.sign(yi) for wi, yi in zip(a2, m.y)]
        mean1.append(np.mean(per_out, axis=0))
        mean2.append([np.sqrt(np.abs(yi) * wi / amp) for wi, yi in zip(a2, m.y)])
        var1.append([np.mean(m.jitters) * np.ones(m.N)])
        var2.append([jit * np.ones(m.N) for jit in m.jitters])     # jitter, not jitter**2 (Q5)
    mu = np.concatenate((mean1, mean2), axis=None)
    var = np.concatenate((var1, var2), axis=None)
    return mu, var


def split_state(m: Model, u):
    """a6: ``_u_to_fhatW`` (meanfield.py:473-489): flat -> f[q,N], w[p,q,N]."""
    u = np.asarray(u, float).ravel()
    return u[: m.q * m.N].reshape(m.q, m.N), u[m.q * m.N:].reshape(m.p, m.q, m.N)


# --------------------------------------------------------------------------------------
# a9: _updateSigMu                                              meanfield.py:747-771,788-792,838-865
# --------------------------------------------------------------------------------------
def update_sig_mu(m: Model, Kf, Kw, ysub, j2, muF, muW, varW):
    q, p, N = m.q, m.p, m.N
    Kw = Kw.reshape(q, p, N, N)
    variance = j2[:, None] + m.yerr2                                   # :759
    dvec = np.sum((muW * muW + varW) / variance[:, None, :], axis=0)   # :765  [q,N]
    sigF = np.empty((q, N, N))
    muF_new = np.empty((q, N))
    for j in range(q):
        A = np.diag(1.0 / dvec[j]) + Kf[j]
        sigF[j] = Kf[j] - Kf[j] @ np.linalg.solve(A, Kf[j])            # :771 (LU, N rhs)
        others = np.delete(muW * muF, j, axis=1).sum(axis=1)           # :788 old muF of other nodes
        pred = np.sum((ysub - others) * muW[:, j, :] / variance, axis=0)
        muF_new[j] = sigF[j] @ pred                                    # :792
    dv = muF_new * muF_new + np.einsum("ijj->ij", sigF)                # :838
    sigW = np.empty((q, p, N, N))
    muW_new = np.empty((p, q, N))
    for j in range(q):
        resid = ysub - np.delete(muF_new * muW, j, axis=1).sum(axis=1)  # :847 new muF, old muW
        for i in range(p):
            A = np.diag(variance[i] / dv[j]) + Kw[j, i]
            sigW[j, i] = Kw[j, i] - Kw[j, i] @ np.linalg.solve(A, Kw[j, i])   # :850
            muW_new[i, j] = sigW[j, i] @ (resid[i] * muF_new[j] / variance[i])  # :864-865
    return sigF, muF_new, sigW, muW_new


# --------------------------------------------------------------------------------------
# a10-a12: the three ELBO terms
# --------------------------------------------------------------------------------------
def expected_log_like(m: Model, j2, sigF, muF, sigW, muW):
    """meanfield.py:923-925, 939-942, 962-972.  Uses the RAW y (quirk Q1)."""
    variance = j2[:, None] + m.yerr2
    ll = -0.5 * np.sum(np.log(2 * np.pi * variance))
    omega_nu = np.einsum("pqn,qn->pn", muW, muF)
    ll += -0.5 * np.sum((m.y - omega_nu) ** 2 / variance)
    dF = np.einsum("ijj->ij", sigF)
    dW = np.einsum("ijkk->ijk", sigW)
    val = 0.0
    for i in range(m.p):
        for j in range(m.q):
            val += dF[j] @ (muW[i, j] ** 2 / variance[i])
            val += dW[j, i] @ (muF[j] ** 2 / variance[i])
            val += dF[j] @ (dW[j, i] / variance[i])
    return ll - 0.5 * val


def expected_log_prior(m: Model, Lf, Lw, sigF, muF, sigW, muW):
    """meanfield.py:1019-1065.  Cumulative sigma_f in the node trace (Q3); mu_w reshaped (Q4)."""
    q, p, N = m.q, m.p, m.N
    Lw = Lw.reshape(q, p, N, N)
    muW_r = muW.reshape(q, p, N)                     # reshape of a (p,q,N) array, not a transpose
    first = second = 0.0
    cum = np.zeros((N, N))
    for j in range(q):
        logK = np.sum(np.log(np.diag(Lf[j])))
        quad = muF[j] @ sla.cho_solve((Lf[j], True), muF[j])
        cum = cum + sigF[j]
        tr = np.trace(sla.cho_solve((Lf[j], True), cum))
        first += -logK - 0.5 * (quad + tr)
        for i in range(p):
            quad = muW_r[j, i] @ sla.cho_solve((Lw[j, i], True), muW_r[j, i])
            tr = np.trace(sla.cho_solve((Lw[j, i], True), sigW[j, i]))
            second += -np.sum(np.log(np.diag(Lw[j, i]))) - 0.5 * (quad + tr)
    return first + second - 0.5 * N * q * (p + 1) * LOG2PI


def entropy(m: Model, sigF, sigW):
    """meanfield.py:1085-1093."""
    ent = 0.0
    for j in range(m.q):
        ent += np.sum(np.log(np.diag(np.linalg.cholesky(sigF[j]))))
        for i in range(m.p):
            ent += np.sum(np.log(np.diag(np.linalg.cholesky(sigW[j, i]))))
    return ent + 0.5 * m.q * (m.p + 1) * m.N * (1 + LOG2PI)


# --------------------------------------------------------------------------------------
# a8: ELBOaux, a7: ELBOcalc
# --------------------------------------------------------------------------------------
def elbo_aux(m: Model, Kf, Kw, Lf, Lw, ysub, j2, mu, var):
    """One fixed-point iteration (meanfield.py:651-710).  Returns (ELBO, mu[1+p,q,N], var[1+p,q,N])."""
    muF, muW = split_state(m, mu)
    _, varW = split_state(m, var)
    sigF, muF, sigW, muW = update_sig_mu(m, Kf, Kw, ysub, j2, muF, muW, varW)
    varF = np.einsum("ijj->ij", sigF)
    varW = np.einsum("jikk->ijk", sigW)                                  # :695-697 -> [p,q,N]
    ent = entropy(m, sigF, sigW)
    lp = expected_log_prior(m, Lf, Lw, sigF, muF, sigW, muW)
    ll = expected_log_like(m, j2, sigF, muF, sigW, muW)
    elbo = (ll + lp + ent) / m.q                                           # :709 (Q2)
    new_mu = np.concatenate((muF[None], muW))
    new_var = np.concatenate((varF[None], varW))
    return elbo, new_mu, new_var, (ll, lp, ent)


def build_matrices(m: Model):
    Kf = np.array([kmatrix(s, m.time, nugget=1e-6) for s in m.nodes])      # :619
    Kw = np.array([kmatrix(s, m.time, nugget=1e-6) for s in m.weights])    # :620
    Lf = np.array([np.linalg.cholesky(K) for K in Kf])                     # :621 (potrf, lower)
    Lw = np.array([np.linalg.cholesky(K) for K in Kw])                     # :622
    return Kf, Kw, Lf, Lw


def elbo_calc(m: Model, max_iter=None, mu=None, var=None, return_trace=False):
    """``ELBOcalc`` (meanfield.py:561-649).  ``mu``/``var`` None -> 'init'.

    Returns (ELBO, mu, var, iterNumber[, trace]) with mu/var shaped (1+p, q, N).
    """
    if mu is None or var is None:
        mu, var = init_mu_var(m)
    if max_iter is None:
        max_iter = 10000                                                   # :615-616
    j2 = m.jitters ** 2
    Kf, Kw, Lf, Lw = build_matrices(m)
    ysub = m.y - m.mean_vals                                               # :623-624
    elbo, *_ = elbo_aux(m, Kf, Kw, Lf, Lw, ysub, j2, mu, var)              # :627 (state discarded, Q7)
    trace = [elbo]
    it = 0
    while it < max_iter:
        elbo, mu, var, _ = elbo_aux(m, Kf, Kw, Lf, Lw, ysub, j2, mu, var)
        trace.append(elbo)
        it += 1
        if it > 3:                                                         # :640-646
            last = np.array(trace[-3:])
            crit = np.abs(np.std(last) / np.mean(last))
            if crit < 1e-3 and crit != 0:
                break
    out = (elbo, mu, var, it)
    return out + (np.array(trace),) if return_trace else out


# --------------------------------------------------------------------------------------
# a14: _gp.GP.prediction, a13: inference._Prediction
# --------------------------------------------------------------------------------------
def gp_prediction(spec, time, tstar, mvec, vvec):
    """_gp.py:107-138, restated without the T x T temporary: only the diagonal of
    ``Kstarstar - Kstar A^-1 Kstar^T`` is ever used (:136-137)."""
    cov = kmatrix(spec, time, nugget=1.25e-12) + np.diag(vvec)            # :125
    cf = sla.cho_factor(cov)                                               # :126
    sol = sla.cho_solve(cf, mvec)                                          # :127
    Kstar = kmatrix(spec, tstar, time)                                     # :129
    # diag(Kstarstar): every in-scope kernel is stationary, so the diagonal is k(0) + nugget   :131
    kss = kernel_eval(spec, np.zeros((1, 1)))[0, 0] + 1.25e-12
    y_mean = Kstar @ sol                                                   # :132
    y_var = kss - np.einsum("tn,nt->t", Kstar, sla.cho_solve(cf, Kstar.T))  # :134-137
    return y_mean, y_var


def prediction(m: Model, tstar, mu, var, mean_at_tstar=None):
    """``_Prediction`` (meanfield.py:1336-1372).  Returns mean[T,p], var[T,p], nPred[q,T], wPred[q*p,T]."""
    tstar = np.asarray(tstar, float)
    T = tstar.size
    muF, muW = split_state(m, mu)
    varF, varW = split_state(m, var)
    mean_at_tstar = np.zeros((m.p, T)) if mean_at_tstar is None else np.asarray(mean_at_tstar, float)
    j2 = m.jitters ** 2
    nP, nV, wP, wV = [], [], [], []
    for j in range(m.q):
        a, b = gp_prediction(m.nodes[j], m.time, tstar, muF[j], varF[j])
        nP.append(a); nV.append(b)
        for i in range(m.p):
            a, b = gp_prediction(m.weights[j * m.p + i], m.time, tstar, muW[i, j], varW[i, j])
            wP.append(a); wV.append(b)
    nP, nV = np.array(nP), np.array(nV)
    wPm = np.array(wP).reshape(m.q, m.p, T)
    wVm = np.array(wV).reshape(m.q, m.p, T)
    pm = np.zeros((T, m.p))
    pv = np.zeros((T, m.p))
    for i in range(m.p):
        pm[:, i] += mean_at_tstar[i]
        for j in range(m.q):
            pm[:, i] += nP[j] * wPm[j, i]
            pv[:, i] += wPm[j, i] ** 2 * nV[j] + wVm[j, i] * (nV[j] + nP[j] ** 2) + j2[i]   # Q6
    return pm, pv, nP, np.array(wP)


# --------------------------------------------------------------------------------------
# synthetic workloads (SURVEY.md section 8d / Appendix B generator)
# --------------------------------------------------------------------------------------
def synth(N, p, q, seed=1, node="QP"):
    import workloads
    a = workloads.synth_arrays(N, p, q, seed=seed, node=node)
    return Model(a["t"], a["y"], a["yerr"], a["nodes"], a["weights"], None, a["jitters"])


def spec_params(spec):
    """Flat parameter list of a spec in the reference's ``.pars`` order (covfunc.py:61)."""
    if spec[0] in ("sum", "mul"):
        return spec_params(spec[1]) + spec_params(spec[2])
    return [float(v) for v in spec[1:]]


def spec_with_params(spec, vals):
    """Rebuild ``spec`` consuming parameters from the list ``vals`` (in place pop from the front)."""
    if spec[0] in ("sum", "mul"):
        a = spec_with_params(spec[1], vals)
        b = spec_with_params(spec[2], vals)
        return (spec[0], a, b)
    n = len(spec) - 1
    out = (spec[0],) + tuple(vals[:n])
    del vals[:n]
    return out


def perturbed_hyper_sets(m: Model, B, seed):
    """theta_b = theta_0 * exp(0.1 * N(0,1)) on kernel parameters and jitters.

    Returns array [B, H] in ``get_parameters`` order WITHOUT the mean parameters:
    node pars, weight pars, jitters."""
    theta0 = []
    for s in m.nodes + m.weights:
        theta0 += spec_params(s)
    theta0 += list(m.jitters)
    theta0 = np.array(theta0)
    z = np.random.default_rng(seed).standard_normal((B, theta0.size))
    return theta0[None, :] * np.exp(0.1 * z)


def model_with_hyper(m: Model, theta):
    vals = [float(v) for v in theta]
    nodes = [spec_with_params(s, vals) for s in m.nodes]
    weights = [spec_with_params(s, vals) for s in m.weights]
    jit = vals[: m.p]
    return Model(m.time, m.y, m.yerr, nodes, weights, m.mean_vals, jit)
