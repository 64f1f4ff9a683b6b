"""Profiling driver for the prediction path: python tools/probe_predict.py N T  (p=4, q=2, Matern52 nodes)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
import gpyrn_b200 as gp
from gpyrn_b200 import covfunc, meanfunc
N, T = int(sys.argv[1]), int(sys.argv[2])
a = workloads.synth_arrays(N, 4, 2, seed=1, node="M52")
ya = []
for y, e in zip(a["y"], a["yerr"]):
    ya += [y, e]
g = gp.inference(2, a["t"], *ya)
g.set_components([covfunc.Matern52(*s[1:]) for s in a["nodes"]], [covfunc.SquaredExponential(*s[1:]) for s in a["weights"]],
                 [meanfunc.Constant(0.0)] * 4, [0.1] * 4)
_, mu, var, _ = g.ELBOcalc(max_iter=3)
t = a["t"]
tstar = np.linspace(t[0] - 0.2 * (t[-1] - t[0]), t[-1] + 0.2 * (t[-1] - t[0]), T)
g._Prediction(tstar=tstar, mu=mu, var=var)
t0 = time.time()
pm, pv = g._Prediction(tstar=tstar, mu=mu, var=var)
print(f"predict N={N} T={T}: {1e3 * (time.time() - t0):.1f} ms, finite={np.all(np.isfinite(pm))}")
