"""Stage-by-stage check of the CUDA path against the oracle (run on the GPU box; development aid).

    python tools/gpu_selfcheck.py [--big]

Prints one line per stage; exits non-zero if a parity bar is missed.  The pytest -m gpu suite covers
the same ground as assertions; this script exists to get all diagnostics out of a single gpurun call.
"""
import ast
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gpyrn_b200 as gp  # noqa: E402
from gpyrn_b200 import _lib, covfunc, meanfunc  # noqa: E402
from oracle import gprn_oracle as orc  # noqa: E402
from tests._cases import GOLDEN, golden_names, load_golden  # noqa: E402

KCLS = {"SE": covfunc.SquaredExponential, "P": covfunc.Periodic, "QP": covfunc.QuasiPeriodic,
        "RQ": covfunc.RationalQuadratic, "M32": covfunc.Matern32, "M52": covfunc.Matern52,
        "WN": covfunc.WhiteNoise, "C": covfunc.Constant, "RQP": covfunc.RQP, "COS": covfunc.Cosine,
        "EXP": covfunc.Exponential}
fails = []


def build_kernel(spec):
    if spec[0] == "sum":
        return build_kernel(spec[1]) + build_kernel(spec[2])
    if spec[0] == "mul":
        return build_kernel(spec[1]) * build_kernel(spec[2])
    if spec[0] in ("dSE", "dP", "dQP"):
        return covfunc.Derivative(KCLS[spec[0][1:]](*spec[1:]))
    return KCLS[spec[0]](*spec[1:])


def make_inference(d):
    args = []
    for y, e in zip(d["y"], d["yerr"]):
        args += [y, e]
    g = gp.inference(len(d["nodes"]), d["t"], *args)
    g.set_components([build_kernel(s) for s in d["nodes"]], [build_kernel(s) for s in d["weights"]],
                     [meanfunc.Constant(c) for c in d["mean_consts"]], list(d["jitters"]))
    return g


def report(stage, ok, msg):
    print(f"[{'ok' if ok else 'FAIL'}] {stage}: {msg}", flush=True)
    if not ok:
        fails.append(stage)


def main():
    big = "--big" in sys.argv
    # 1. kernels
    z = np.load(os.path.join(GOLDEN, "kernels.npz"))
    t, ts = z["t"], z["tstar"]
    g0 = gp.inference(1, t, np.zeros_like(t), np.ones_like(t))
    worst = 0.0
    for i, s in enumerate(z["specs"]):
        k = build_kernel(ast.literal_eval(str(s)))
        Ksq = g0._kmat(k, t, None, 0.0)
        Kre = g0._kmat(k, ts, t, 0.0)
        Kev = k(t[:, None] - t[None, :])
        for got, ref in ((Ksq, z[f"Ksq_{i}"]), (Kre, z[f"Krect_{i}"]), (Kev, z[f"Ksq_{i}"])):
            err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300) * (np.abs(ref) > 1e-280))
            worst = max(worst, err)
    report("kernels", worst < 1e-13, f"max rel err {worst:.2e}")

    # 2. factorisation
    rng = np.random.default_rng(0)
    for n in (50, 64, 200, 256, 500):
        tt = np.sort(rng.uniform(0, 40 * n ** 0.5, n))
        A = orc.kmatrix(("QP", 1.0, 60.0, 25.0, 0.7), tt, nugget=1e-6) + np.diag(rng.uniform(0.01, 1.0, n))
        L = np.empty((n, n)); X = np.empty((n, n)); ld = np.zeros(1)
        _lib.check(_lib.lib().gprn_debug_factor(g0._h(), n, _lib.dptr(_lib.f64(A)), _lib.dptr(L), _lib.dptr(X), _lib.dptr(ld)))
        Lr = np.linalg.cholesky(A)
        eL = np.max(np.abs(L - Lr)) / np.max(np.abs(Lr))
        eX = np.max(np.abs(X @ Lr - np.eye(n)))
        eld = abs(ld[0] - 2 * np.sum(np.log(np.diag(Lr)))) / abs(2 * np.sum(np.log(np.diag(Lr))))
        report(f"factor n={n}", eL < 1e-10 and eX < 1e-8 and eld < 1e-12, f"L {eL:.1e}  |XL-I| {eX:.1e}  logdet {eld:.1e}")

    # 3. ELBO + prediction against the reference golden vectors
    for name in golden_names():
        d = load_golden(name)
        if d["t"].size > 300 and not big and name.startswith("synth_256_4_2"):
            pass
        g = make_inference(d)
        t0 = time.time()
        elbo, mu, var, it = g.ELBOcalc(max_iter=d["max_iter"])
        dt = time.time() - t0
        rel = abs(elbo - d["elbo"]) / abs(d["elbo"])
        emu = np.max(np.abs(mu - d["mu"])) / np.max(np.abs(d["mu"]))
        evar = np.max(np.abs(var - d["var"])) / np.max(np.abs(d["var"]))
        report(f"elbo {name}", it == d["iters"] and rel < 1e-10,
               f"ELBO {elbo!r} it {it}/{d['iters']} rel {rel:.2e} mu {emu:.1e} var {evar:.1e} ({dt * 1e3:.0f} ms)")
        if "tstar" in d:
            pm, pv, sep = g._Prediction(tstar=d["tstar"], mu=d["mu"], var=d["var"], separate=True)
            em = np.max(np.abs(pm - d["pred_mean"])) / np.max(np.abs(d["pred_mean"]))
            ev = np.max(np.abs(pv - d["pred_var"])) / np.max(np.abs(d["pred_var"]))
            en = np.max(np.abs(sep[0] - d["node_pred"])) / np.max(np.abs(d["node_pred"]))
            ew = np.max(np.abs(sep[1] - d["weight_pred"])) / np.max(np.abs(d["weight_pred"]))
            report(f"pred {name}", max(em, ev, en, ew) < 1e-8, f"mean {em:.1e} var {ev:.1e} node {en:.1e} weight {ew:.1e}")
        g.close()

    # 4. batched: first few perturbed sets of C3 against the oracle, and throughput
    m = orc.synth(256, 4, 1, seed=1, node="QP")
    B = 2048 if big else 256
    theta = orc.perturbed_hyper_sets(m, B, 101)
    d = load_golden("c3_synth_256_4_1_QP")
    g = make_inference(d)
    P = np.concatenate([theta[:, :-4], np.zeros((B, 4)), theta[:, -4:]], axis=1)
    g.ELBO_batch(P[:8])
    t0 = time.time()
    elbo, iters, status = g.ELBO_batch(P, return_info=True)
    dt = time.time() - t0
    ms = _lib.lib().gprn_last_elbo_ms(g._h())
    print(f"C3 batch B={B}: {dt:.3f} s wall, {ms:.1f} ms device, {B / dt:.1f} evals/s, iters mean {iters.mean():.1f} "
          f"min {iters.min()} max {iters.max()}, status {np.bincount(status)}", flush=True)
    worst = 0.0
    nchk = 6
    for b in range(nchk):
        mb = orc.model_with_hyper(m, theta[b])
        e, _, _, it = orc.elbo_calc(mb)
        rel = abs(e - elbo[b]) / abs(e)
        worst = max(worst, rel)
        if it != iters[b]:
            report(f"batch set {b}", False, f"iters {iters[b]} vs oracle {it}")
    report("batch parity", worst < 1e-10, f"max rel over {nchk} sets {worst:.2e}")
    g.close()

    if big:
        for (N, p, q, node, B) in ((4096, 4, 2, "M52", 2),):
            m = orc.synth(N, p, q, seed=1, node=node)
            theta = orc.perturbed_hyper_sets(m, B, 102)
            args = []
            for y, e in zip(m.y, m.yerr):
                args += [y, e]
            g = gp.inference(q, m.time, *args)
            g.set_components([build_kernel(s) for s in m.nodes], [build_kernel(s) for s in m.weights],
                             [meanfunc.Constant(0.0)] * p, [0.1] * p)
            P = np.concatenate([theta[:, :-p], np.zeros((B, p)), theta[:, -p:]], axis=1)
            t0 = time.time()
            elbo, iters, status = g.ELBO_batch(P, max_iter=3, return_info=True)
            dt = time.time() - t0
            print(f"C4 N={N} B={B} max_iter=3: {dt:.2f} s, elbo {elbo}, iters {iters}, status {status}", flush=True)
            g.close()
    print("FAILED: " + ", ".join(fails) if fails else "ALL OK")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
