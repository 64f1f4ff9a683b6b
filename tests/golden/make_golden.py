"""Generate golden input/output vectors from the UNMODIFIED reference (build container only).

The reference (pure Python, /root/reference/gpyrn) needs jax, emcee and matplotlib, none of
which exist in this image, and trips over numpy>=2 (``np.float``).  It runs from its own
sources once stand-in modules are registered for those imports: ``jax.numpy`` backed by numpy,
``jax.jit`` as identity, ``jax.scipy.linalg.cho_solve`` = scipy's.
Arithmetic deviation from real JAX is rounding-level only (LAPACK potrf/potrs either way).

Run once:  python tests/golden/make_golden.py     (writes tests/golden/*.npz)
The GPU box has no /root/reference; tests only read the committed .npz files.
"""
import os
import sys
import types

import numpy as np
import scipy.linalg

REF = os.environ.get("GPYRN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shim():
    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    def jit(f=None, static_argnums=None, **kw):
        return (lambda g: g) if f is None else f

    jnp = mod("jax.numpy")
    jnp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
    jnp.ndarray = np.ndarray
    jsl = mod("jax.scipy.linalg", cho_solve=scipy.linalg.cho_solve)
    js = mod("jax.scipy", linalg=jsl)

    class _Cfg:
        def update(self, *a, **k):
            pass

    mod("jax", jit=jit, numpy=jnp, scipy=js, config=_Cfg())
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")
    em = mod("emcee", EnsembleSampler=object, backends=types.SimpleNamespace())
    em.utils = mod("emcee.utils", sample_ellipsoid=None)
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)


install_shim()
from gpyrn import covfunc, meanfunc, meanfield  # noqa: E402  (the reference itself)

KCLS = {"SE": covfunc.SquaredExponential, "P": covfunc.Periodic, "QP": covfunc.QuasiPeriodic,
        "RQ": covfunc.RationalQuadratic, "M32": covfunc.Matern32, "M52": covfunc.Matern52,
        "WN": covfunc.WhiteNoise, "C": covfunc.Constant, "RQP": covfunc.RQP, "COS": covfunc.Cosine,
        "EXP": covfunc.Exponential,
        # the stationary "other" kernels (no _tag in the reference; the tags are this repo's)
        "GammaExp": covfunc.GammaExp, "PW": covfunc.Piecewise, "PAC": covfunc.Paciorek, "NP": covfunc.NewPeriodic,
        "QNP": covfunc.QuasiNewPeriodic, "CP": covfunc.CosPeriodic, "QCP": covfunc.QuasiCosPeriodic}


def build_kernel(spec):
    if spec[0] == "sum":
        return build_kernel(spec[1]) + build_kernel(spec[2])
    if spec[0] == "mul":
        return build_kernel(spec[1]) * build_kernel(spec[2])
    if spec[0] in ("dSE", "dP", "dQP"):
        return covfunc.Derivative(KCLS[spec[0][1:]](*spec[1:]))
    return KCLS[spec[0]](*spec[1:])


def spec_to_arr(spec):
    """Serialise a spec as a string (np.savez friendly)."""
    return repr(spec)


def run_case(name, t, ys, es, nodes, weights, mean_consts, jitters, max_iter=None, tstar=None):
    q = len(nodes)
    args = []
    for y, e in zip(ys, es):
        args += [y, e]
    g = meanfield.inference(q, t, *args)
    g.set_components([build_kernel(s) for s in nodes], [build_kernel(s) for s in weights],
                     [meanfunc.Constant(c) for c in mean_consts], list(jitters))
    # per-iteration trace: wrap ELBOaux
    trace = []
    orig = g.ELBOaux

    def rec(*a, **k):
        out = orig(*a, **k)
        trace.append(float(out[0]))
        return out

    g.ELBOaux = rec
    elbo, mu, var, it = g.ELBOcalc(max_iter=max_iter)
    out = dict(t=t, y=np.array(ys), yerr=np.array(es), nodes=np.array([spec_to_arr(s) for s in nodes]),
               weights=np.array([spec_to_arr(s) for s in weights]), mean_consts=np.array(mean_consts, float),
               jitters=np.array(jitters, float), elbo=float(elbo), mu=np.asarray(mu), var=np.asarray(var),
               iters=int(it), trace=np.array(trace), max_iter=-1 if max_iter is None else max_iter)
    if tstar is not None:
        pm, pv, sep = g._Prediction(tstar=tstar, mu=np.asarray(mu), var=np.asarray(var), separate=True)
        out.update(tstar=tstar, pred_mean=pm, pred_var=pv, node_pred=np.array(sep[0], float),
                   weight_pred=np.array(sep[1], float))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: ELBO={elbo!r} iters={it}")


def synth_data(N, p, seed=1):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 4 * N ** 0.5 * 10, N))
    ys, es = [], []
    for i in range(p):
        ys.append(np.sin(2 * np.pi * t / 25 + i) * (1 + 0.3 * i) + 0.1 * rng.standard_normal(N))
        es.append(rng.uniform(.05, .15, N))
    return t, ys, es


def synth_case(name, N, p, q, node="QP", means=None, max_iter=None, T=None):
    t, ys, es = synth_data(N, p)
    if node == "QP":
        nodes = [("QP", 1 + .2 * j, 60 + 5 * j, 25 + j, .7) for j in range(q)]
    else:
        nodes = [("M52", 1 + .2 * j, 30 + 5 * j) for j in range(q)]
    weights = [("SE", 1 + .1 * k, 80 + k) for k in range(q * p)]
    tstar = None if T is None else np.linspace(t[0] - 5.0, t[-1] + 5.0, T)
    run_case(name, t, ys, es, nodes, weights, means or [0.0] * p, [0.1] * p, max_iter=max_iter, tstar=tstar)


def kernel_vectors():
    """Element-wise k(r) of every in-scope kernel (+Sum, Multiplication) on a fixed lag grid."""
    rng = np.random.default_rng(7)
    t = np.sort(rng.uniform(0, 100, 40))
    ts = np.linspace(-3, 104, 23)
    specs = [("SE", 1.3, 11.0), ("P", 0.9, 17.0, 0.8), ("QP", 1.1, 35.0, 23.0, 0.6), ("RQ", 1.2, 0.7, 9.0),
             ("M32", 0.8, 14.0), ("M52", 1.4, 21.0), ("WN", 0.3),
             ("sum", ("SE", 1.0, 10.0), ("WN", 0.2)), ("mul", ("SE", 1.0, 10.0), ("P", 1.0, 20.0, 0.5)),
             ("sum", ("mul", ("M52", 1.1, 30.0), ("P", 1.0, 12.0, 0.9)), ("RQ", 0.5, 1.5, 40.0)),
             ("C", 0.7), ("RQP", 1.2, 0.8, 33.0, 19.0, 0.9), ("COS", 0.9, 14.0), ("EXP", 1.1, 25.0),
             ("sum", ("mul", ("EXP", 1.0, 50.0), ("COS", 1.0, 9.0)), ("C", 0.1)),
             ("dSE", 12.0, 11.0), ("dP", 0.2, 17.0, 0.8), ("dQP", 4.0, 35.0, 23.0, 0.6),
             ("sum", ("mul", ("dSE", 9.0, 10.0), ("M52", 1.0, 30.0)), ("WN", 0.1))]
    out = dict(t=t, tstar=ts, specs=np.array([repr(s) for s in specs]))
    for i, s in enumerate(specs):
        k = build_kernel(s)
        out[f"Ksq_{i}"] = k(t[:, None] - t[None, :])
        out[f"Krect_{i}"] = k(ts[:, None] - t[None, :])
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels: ok")


def extra_kernel_vectors():
    """k(r) of the stationary "other" kernels of the reference (covfunc.py:415-432, 458-546, 645-688) on the lag grid
    of kernel_vectors(), and one ELBO / prediction case that runs them through the whole path (without CosPeriodic /
    QuasiCosPeriodic: exp(-2 cos^2 x / l^2) has negative Fourier coefficients, the matrices are not positive definite)."""
    rng = np.random.default_rng(7)
    t = np.sort(rng.uniform(0, 100, 40))
    ts = np.linspace(-3, 104, 23)
    specs = [("GammaExp", 1.2, 1.3, 14.0), ("GammaExp", 0.7, 2.0, 9.0), ("PW", 60.0), ("PW", 250.0),
             ("PAC", 1.1, 12.0, 30.0), ("NP", 0.9, 1.4, 17.0, 0.8), ("QNP", 1.3, 0.7, 40.0, 21.0, 1.1),
             ("CP", 0.8, 19.0, 1.2), ("QCP", 1.2, 35.0, 23.0, 0.9),
             ("sum", ("mul", ("GammaExp", 1.0, 1.5, 30.0), ("PW", 300.0)), ("WN", 0.1)),
             ("mul", ("QNP", 1.0, 2.0, 80.0, 12.0, 0.9), ("PAC", 1.0, 50.0, 70.0))]
    out = dict(t=t, tstar=ts, specs=np.array([repr(s) for s in specs]))
    for i, s in enumerate(specs):
        k = build_kernel(s)
        out[f"Ksq_{i}"] = k(t[:, None] - t[None, :])
        out[f"Krect_{i}"] = k(ts[:, None] - t[None, :])
    np.savez_compressed(os.path.join(HERE, "extra_kernels.npz"), **out)
    print("extra kernels: ok")
    t, ys, es = synth_data(64, 2, seed=13)
    run_case("other_kernels_64_2_1", t, ys, es, [("QNP", 1.0, 1.5, 60.0, 25.0, 0.9)],
             [("sum", ("mul", ("GammaExp", 1.1, 1.6, 120.0), ("PW", 900.0)), ("WN", 0.03)),
              ("sum", ("PAC", 0.9, 70.0, 110.0), ("mul", ("SE", 0.4, 90.0), ("NP", 1.0, 2.0, 50.0, 1.5)))],
             [0.0, 0.1], [0.1, 0.1], tstar=np.linspace(t[0], t[-1], 31))


def big_anchors():
    """Large-N anchors (C5-size with max_iter=6, C4 with max_iter=2); minutes of CPU time.
    Only ELBO, trace and iteration count are kept (mu/var at these sizes would bloat the repo)."""
    for name, N, mi in (("c5_synth_2048_4_2_M52_it6", 2048, 6), ("c4_synth_4096_4_2_M52_it2", 4096, 2)):
        t, ys, es = synth_data(N, 4)
        nodes = [("M52", 1 + .2 * j, 30 + 5 * j) for j in range(2)]
        weights = [("SE", 1 + .1 * k, 80 + k) for k in range(8)]
        args = []
        for y, e in zip(ys, es):
            args += [y, e]
        g = meanfield.inference(2, t, *args)
        g.set_components([build_kernel(s) for s in nodes], [build_kernel(s) for s in weights],
                         [meanfunc.Constant(0.0)] * 4, [0.1] * 4)
        trace = []
        orig = g.ELBOaux

        def rec(*a, **k):
            out = orig(*a, **k)
            trace.append(float(out[0]))
            return out

        g.ELBOaux = rec
        elbo, mu, var, it = g.ELBOcalc(max_iter=mi)
        np.savez_compressed(os.path.join(HERE, "big", name + ".npz"), N=N, p=4, q=2, seed=1, node="M52",
                            max_iter=mi, elbo=float(elbo), iters=int(it), trace=np.array(trace),
                            mu_head=np.asarray(mu)[:, :, :16], var_head=np.asarray(var)[:, :, :16])
        print(f"{name}: ELBO={elbo!r} iters={it}", flush=True)


def big_converged(which):
    """Converged large-N anchors: the reference run to its own stopping rule.
    c5:  C5-size N=2048 theta_0 + _Prediction on every 10th epoch of the T=20000 grid of SURVEY.md 8d;
    c4a: C4 theta_0;  c4b: set 0 of the bench pool (perturbed_sets(theta_0, ., seed=102)[0]).
    About 1 min of host time per ELBOaux at N=4096 on 8 cores: run once, in the background."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import workloads
    N = 2048 if which == "c5" else 4096
    a = workloads.synth_arrays(N, 4, 2, seed=1, node="M52")
    th = workloads.theta0(a)
    if which == "c4b":
        th = workloads.perturbed_sets(th, 32, 102)[0]
    nodes = [("M52", th[2 * j], th[2 * j + 1]) for j in range(2)]
    weights = [("SE", th[4 + 2 * k], th[4 + 2 * k + 1]) for k in range(8)]
    jit = th[-4:]
    args = []
    for y, e in zip(a["y"], a["yerr"]):
        args += [y, e]
    g = meanfield.inference(2, a["t"], *args)
    g.set_components([build_kernel(s) for s in nodes], [build_kernel(s) for s in weights],
                     [meanfunc.Constant(0.0)] * 4, list(jit))
    trace = []
    orig = g.ELBOaux

    def rec(*aa, **k):
        out = orig(*aa, **k)
        trace.append(float(out[0]))
        print(f"  {which} ELBOaux #{len(trace)}: {out[0]!r}", flush=True)
        return out

    g.ELBOaux = rec
    elbo, mu, var, it = g.ELBOcalc()
    out = dict(N=N, p=4, q=2, seed=1, node="M52", theta=th, max_iter=-1, elbo=float(elbo), iters=int(it),
               trace=np.array(trace), mu=np.asarray(mu), var=np.asarray(var))
    name = {"c5": "c5_synth_2048_4_2_M52_conv", "c4a": "c4_synth_4096_4_2_M52_conv",
            "c4b": "c4_pool102_set0_4096_4_2_M52_conv"}[which]
    np.savez_compressed(os.path.join(HERE, "big", name + ".npz"), **out)
    print(f"{name}: ELBO={elbo!r} iters={it}", flush=True)
    if which == "c5":
        t = a["t"]
        ptp = t[-1] - t[0]
        full = np.linspace(t[0] - 0.2 * ptp, t[-1] + 0.2 * ptp, 20000)
        sl = np.arange(0, 20000, 10)
        pm, pv, sep = g._Prediction(tstar=full[sl], mu=np.asarray(mu), var=np.asarray(var), separate=True)
        out.update(T_full=20000, slice_idx=sl, pred_mean=pm, pred_var=pv, node_pred=np.array(sep[0], float),
                   weight_pred=np.array(sep[1], float))
        np.savez_compressed(os.path.join(HERE, "big", name + ".npz"), **out)
        print(f"{name}: prediction slice done", flush=True)


def derivative_case():
    """Derivative kernels (covfunc.py:80-104) through the whole ELBO / prediction path."""
    t, ys, es = synth_data(60, 2, seed=11)
    run_case("deriv_kernels_60_2_1", t, ys, es, [("dQP", 4.0, 60.0, 25.0, 0.9)],
             [("sum", ("dSE", 30.0, 40.0), ("WN", 0.05)), ("SE", 1.1, 80.0)],
             [0.0, 0.1], [0.1, 0.1], tstar=np.linspace(t[0], t[-1], 29))


def main():
    if "--big-converged" in sys.argv:
        os.makedirs(os.path.join(HERE, "big"), exist_ok=True)
        for which in sys.argv[sys.argv.index("--big-converged") + 1:]:
            big_converged(which)
        return
    if "--big" in sys.argv:
        os.makedirs(os.path.join(HERE, "big"), exist_ok=True)
        big_anchors()
        return
    if "--extra-kernels" in sys.argv:
        extra_kernel_vectors()
        return
    kernel_vectors()
    if "--kernels" in sys.argv:
        return
    derivative_case()
    if "--deriv" in sys.argv:
        return
    # notebook data (docs/examples/one_dataset.ipynb; SURVEY.md 8c anchor -267.06958539495247, 4 it)
    from scipy.stats import norm
    np.random.seed(43)
    t = np.sort(np.random.uniform(10, 60, 45))
    y = 1.5 * np.sin(2 * np.pi * t / 13.5) * np.polyval([0.01, 0.02, 2.5], t)
    yerr = np.random.uniform(2, 5, 45)
    y = y + norm(0, np.hypot(0.5, yerr)).rvs()
    run_case("notebook_45_1_1", t, [y], [yerr], [("P", 1, 13, 1)], [("SE", 1, 50)], [0.0], [0.1],
             tstar=np.linspace(5, 65, 77))
    # C1: bundled solar RV data, (497,1,1)
    d = np.loadtxt(os.path.join(REF, "gpyrn", "datasets", "Solar_observations.txt"), skiprows=1)
    t, rv, rve = d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy()
    run_case("c1_solar_497_1_1", t, [rv], [rve], [("QP", 1, 30, 27, 0.7)], [("SE", 2, 200)], [rv.mean()], [0.5],
             tstar=np.linspace(t[0], t[-1], 200))
    synth_case("synth_50_1_1_QP", 50, 1, 1, T=31)
    synth_case("synth_100_4_1_QP", 100, 4, 1, T=64)
    synth_case("synth_60_2_2_QP_means", 60, 2, 2, means=[0.3, 0.6], T=50)
    synth_case("synth_100_4_2_M52_means", 100, 4, 2, node="M52", means=[0.3, 0.6, 0.9, 1.2], T=40)
    synth_case("c3_synth_256_4_1_QP", 256, 4, 1)
    synth_case("synth_256_4_2_M52", 256, 4, 2, node="M52", T=100)
    synth_case("c2_synth_500_4_1_QP", 500, 4, 1, T=128)
    # mixed kernels: exercises every in-scope tag through the whole ELBO path.  (A bare Periodic weight is
    # avoided: K is then rank-deficient up to the nugget and the reference's own ELBO moves by 1e-9 relative
    # under a 1-ulp change of the inputs, so it cannot anchor a 1e-10 parity test.)
    t, ys, es = synth_data(80, 2, seed=5)
    run_case("mixed_80_2_2", t, ys, es,
             [("sum", ("M32", 1.0, 40.0), ("WN", 0.05)), ("mul", ("SE", 1.1, 70.0), ("P", 1.0, 25.0, 0.8))],
             [("RQ", 1.0, 0.9, 60.0), ("M52", 1.1, 90.0), ("mul", ("P", 0.9, 33.0, 1.1), ("M32", 1.0, 120.0)),
              ("SE", 1.2, 75.0)],
             [0.1, -0.2], [0.12, 0.08], tstar=np.linspace(t[0], t[-1] + 10, 45))
    # the four "next" kernels through the whole ELBO / prediction path
    t, ys, es = synth_data(70, 2, seed=9)
    run_case("next_kernels_70_2_1", t, ys, es, [("sum", ("RQP", 1.0, 1.2, 80.0, 25.0, 0.9), ("C", 0.2))],
             [("EXP", 1.1, 150.0), ("sum", ("mul", ("SE", 1.0, 90.0), ("COS", 1.0, 400.0)), ("WN", 0.02))],
             [0.0, 0.1], [0.1, 0.1], tstar=np.linspace(t[0], t[-1], 33))


if __name__ == "__main__":
    main()
