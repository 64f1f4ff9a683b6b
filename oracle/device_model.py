"""numpy model of the DEVICE algorithm (TEST INFRASTRUCTURE -- not part of the product).

Where ``gprn_oracle.py`` restates the reference's arithmetic literally (explicit Sigma, LU
solves), this file states the Sigma-free algorithm the CUDA path implements (SURVEY.md
Appendix A.3), with the same blocking (tile NB, identity padding, substitution-based panel TRSM,
blocked triangular inverse) so that intermediates can be compared kernel by kernel when debugging,
and so that the numerical choices (explicit triangular inverse for the solves, quadratic form
via the fixed-point identity) can be checked against the reference golden vectors on the CPU.

With A = K + D, L = chol(A), X = L^-1, g = colnorm2(X) = diag(A^-1):
    Sigma = D - D A^-1 D           diag Sigma = D - D^2 g          mu = D b - D X^T X (D b)
    logdet Sigma = logdet K - logdet A + sum log D                 tr(K^-1 Sigma) = sum D g
    mu^T K^-1 mu = mu . (b - mu / D)        (because Sigma^-1 mu = b  =>  K^-1 mu = b - D^-1 mu)
    tr(K_j^-1 Sigma_k) = sum_n D_k gK_j - || X_Kj D_k X_Ak^T ||_F^2           (k < j, quirk Q3)
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import gprn_oracle as orc

NB = 64


def pad_spd(K, nb=NB):
    n = K.shape[0]
    npad = -(-n // nb) * nb
    P = np.eye(npad)
    P[:n, :n] = K
    return P


def blocked_chol(A, nb=NB):
    """Right-looking blocked Cholesky.  The panel TRSM is a genuine forward substitution with the
    diagonal tile (multiplying by an explicit tile inverse costs 1-2 digits of ELBO parity on
    ill-conditioned K, measured on synth_256_4_2_M52: 4.8e-10 vs 1.7e-11)."""
    A = A.copy()
    n = A.shape[0]
    nt = n // nb
    for k in range(nt):
        s = slice(k * nb, (k + 1) * nb)
        Lkk = np.linalg.cholesky(A[s, s])
        A[s, s] = Lkk
        if k + 1 < nt:
            r = slice((k + 1) * nb, n)
            A[r, s] = sla.solve_triangular(Lkk, A[r, s].T, lower=True).T
            A[r, r] -= A[r, s] @ A[r, s].T
    return np.tril(A)


def blocked_trtri(L, nb=NB):
    """X = L^-1 by block rows: L_ii X_ij = -sum_{k=j}^{i-1} L_ik X_kj  (forward substitution per column)."""
    n = L.shape[0]
    nt = n // nb
    X = np.zeros_like(L)
    for i in range(nt):
        si = slice(i * nb, (i + 1) * nb)
        X[si, si] = sla.solve_triangular(L[si, si], np.eye(nb), lower=True)
        if i:
            left = slice(0, i * nb)
            X[si, left] = -sla.solve_triangular(L[si, si], L[si, left] @ X[left, left], lower=True)
    return X


class Factor:
    """chol + inverse of one SPD matrix, device style."""

    def __init__(self, A, n):
        self.n = n
        Ap = pad_spd(A)
        self.L = blocked_chol(Ap)
        self.X = blocked_trtri(self.L)
        self.logdet = 2.0 * np.sum(np.log(np.diag(self.L)[:n]))
        self.g = np.sum(self.X * self.X, axis=0)[:n]

    def ainv(self, v):
        vp = np.zeros(self.L.shape[0])
        vp[: self.n] = v
        return (self.X.T @ (self.X @ vp))[: self.n]


def elbo_calc(m: orc.Model, max_iter=None, mu=None, var=None, return_trace=False, quad_identity=True):
    q, p, N = m.q, m.p, m.N
    if mu is None or var is None:
        mu, var = orc.init_mu_var(m)
    if max_iter is None:
        max_iter = 10000
    j2 = m.jitters ** 2
    variance = j2[:, None] + m.yerr2
    ysub = m.y - m.mean_vals
    Kf = [orc.kmatrix(s, m.time, nugget=1e-6) for s in m.nodes]
    Kw = [orc.kmatrix(s, m.time, nugget=1e-6) for s in m.weights]
    FKf = [Factor(K, N) for K in Kf]
    FKw = [Factor(K, N) for K in Kw]
    need_xk = (q > 1) or not quad_identity
    muF, muW = orc.split_state(m, mu)
    varF, varW = orc.split_state(m, var)
    muF, muW, varF, varW = muF.copy(), muW.copy(), varF.copy(), varW.copy()
    const_ll = -0.5 * np.sum(np.log(2 * np.pi * variance))
    trace = []
    it = 0
    first = True
    while True:
        # ---- node phase (old weights, old nodes of the others) ----
        dvec = np.sum((muW * muW + varW) / variance[:, None, :], axis=0)
        muF_new = np.empty_like(muF)
        varF_new = np.empty_like(varF)
        ent = lp = 0.0
        FA_nodes, D_nodes = [], []
        for j in range(q):
            D = 1.0 / dvec[j]
            F = Factor(Kf[j] + np.diag(D), N)
            others = np.delete(muW * muF, j, axis=1).sum(axis=1)
            b = np.sum((ysub - others) * muW[:, j, :] / variance, axis=0)
            v = D * b
            muF_new[j] = v - D * F.ainv(v)
            varF_new[j] = D - D * D * F.g
            ent += 0.5 * (FKf[j].logdet - F.logdet + np.sum(np.log(D)))
            if need_xk:
                z = FKf[j].X[:N, :N] @ muF_new[j]
                quad = z @ z
            else:
                quad = muF_new[j] @ (b - muF_new[j] * dvec[j])
            tr = np.sum(D * F.g)
            for k in range(j):              # quirk Q3: cumulative Sigma_f
                Dk, Fk = D_nodes[k], FA_nodes[k]
                C = (FKf[j].X[:N, :N] * Dk[None, :]) @ Fk.X[:N, :N].T
                tr += np.sum(Dk * FKf[j].g) - np.sum(C * C)
            lp += -0.5 * FKf[j].logdet - 0.5 * (quad + tr)
            FA_nodes.append(F)
            D_nodes.append(D)
        # ---- weight phase (new nodes, old weights of the other nodes) ----
        dv = muF_new * muF_new + varF_new
        muW_new = np.empty_like(muW)
        varW_new = np.empty_like(varW)
        bw = np.empty_like(muW)
        dw = np.empty_like(muW)
        for j in range(q):
            resid = ysub - np.delete(muF_new * muW, j, axis=1).sum(axis=1)
            for i in range(p):
                D = variance[i] / dv[j]
                F = Factor(Kw[j * p + i] + np.diag(D), N)
                b = resid[i] * muF_new[j] / variance[i]
                v = D * b
                muW_new[i, j] = v - D * F.ainv(v)
                varW_new[i, j] = D - D * D * F.g
                bw[i, j], dw[i, j] = b, 1.0 / D
                ent += 0.5 * (FKw[j * p + i].logdet - F.logdet + np.sum(np.log(D)))
                lp += -0.5 * FKw[j * p + i].logdet - 0.5 * np.sum(D * F.g)
        muW_r = muW_new.reshape(q, p, N)        # quirk Q4
        for j in range(q):
            for i in range(p):
                if need_xk:
                    z = FKw[j * p + i].X[:N, :N] @ muW_r[j, i]
                    lp += -0.5 * (z @ z)
                else:
                    lp += -0.5 * (muW_new[i, j] @ (bw[i, j] - muW_new[i, j] * dw[i, j]))
        M = q * (p + 1)
        ent += 0.5 * M * N * (1 + orc.LOG2PI)
        lp += -0.5 * N * M * orc.LOG2PI
        # ---- likelihood (raw y: Q1) ----
        omega = np.einsum("pqn,qn->pn", muW_new, muF_new)
        ll = const_ll - 0.5 * np.sum((m.y - omega) ** 2 / variance)
        val = 0.0
        for i in range(p):
            for j in range(q):
                val += np.sum((varF_new[j] * muW_new[i, j] ** 2 + varW_new[i, j] * muF_new[j] ** 2
                               + varF_new[j] * varW_new[i, j]) / variance[i])
        ll -= 0.5 * val
        elbo = (ll + lp + ent) / q
        if first:                         # quirk Q7: the pre-loop evaluation equals iteration 1
            trace.append(elbo)
            first = False
        if max_iter == 0:
            break
        trace.append(elbo)
        muF, muW, varF, varW = muF_new, muW_new, varF_new, varW_new
        it += 1
        if it > 3:
            last = np.array(trace[-3:])
            crit = np.abs(np.std(last) / np.mean(last))
            if crit < 1e-3 and crit != 0:
                break
        if it >= max_iter:
            break
    mu_out = np.concatenate((muF[None], muW))
    var_out = np.concatenate((varF[None], varW))
    out = (elbo, mu_out, var_out, it)
    return out + (np.array(trace),) if return_trace else out
