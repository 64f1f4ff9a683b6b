"""Synthetic inputs of the benchmark / parity configurations.

Neutral helper (plain numpy, no arithmetic of the path): used by bench.py, the tests and the oracle so
that every consumer sees bit-identical inputs.  Kernel specs are tuples ("QP", theta, le, P, lp) etc.
"""
import numpy as np


def synth_arrays(N, p, q, seed=1, node="QP"):
    """t[N], y[p,N], yerr[p,N], node specs, weight specs (index j*p+i), jitters."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 4 * N ** 0.5 * 10, N))
    ys, es = [], []
    for i in range(p):
        ys.append(np.sin(2 * np.pi * t / 25 + i) * (1 + 0.3 * i) + 0.1 * rng.standard_normal(N))
        es.append(rng.uniform(.05, .15, N))
    if node == "QP":
        nodes = [("QP", 1 + .2 * j, 60 + 5 * j, 25 + j, .7) for j in range(q)]
    else:
        nodes = [("M52", 1 + .2 * j, 30 + 5 * j) for j in range(q)]
    weights = [("SE", 1 + .1 * k, 80 + k) for k in range(q * p)]
    return dict(t=t, y=np.array(ys), yerr=np.array(es), nodes=nodes, weights=weights, jitters=np.full(p, 0.1))


def spec_params(spec):
    """Flat parameter list of a spec in `.pars` order (leaf parameters left to right)."""
    if spec[0] in ("sum", "mul"):
        return spec_params(spec[1]) + spec_params(spec[2])
    return [float(v) for v in spec[1:]]


def theta0(w):
    """[node pars, weight pars, jitters] of a synth_arrays() workload."""
    out = []
    for s in list(w["nodes"]) + list(w["weights"]):
        out += spec_params(s)
    return np.array(out + list(w["jitters"]), dtype=float)


def perturbed_sets(th0, B, seed):
    """theta_b = theta_0 * exp(0.1 * N(0,1)), shape [B, H]."""
    z = np.random.default_rng(seed).standard_normal((B, th0.size))
    return th0[None, :] * np.exp(0.1 * z)
