"""Host-side overhead of a single evaluation through the Python API (development aid)."""
import cProfile, pstats, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyrn_b200 as gp
from gpyrn_b200 import _lib, covfunc, meanfunc
import workloads
a = workloads.synth_arrays(500, 4, 1, seed=1, node="QP")
ya = []
for y, e in zip(a["y"], a["yerr"]):
    ya += [y, e]
g = gp.inference(1, a["t"], *ya)
g.set_components([covfunc.QuasiPeriodic(*a["nodes"][0][1:])], [covfunc.SquaredExponential(*s[1:]) for s in a["weights"]],
                 [meanfunc.Constant(0.0)] * 4, [0.1] * 4)
P = g.get_parameters()[None, :]
for _ in range(3):
    g.ELBO_batch(P); g.ELBOcalc()
for name, fn in (("ELBO_batch", lambda: g.ELBO_batch(P)), ("ELBOcalc", lambda: g.ELBOcalc())):
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    dt = (time.perf_counter() - t0) / 20
    print(f"{name}: {dt*1e3:.2f} ms per call, device {_lib.lib().gprn_last_elbo_ms(g._h()):.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    g.ELBO_batch(P)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
