"""Accuracy of the factorisation kernels at larger N (development aid)."""
import os, sys, time
import numpy as np, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyrn_b200 as gp
from gpyrn_b200 import _lib
import workloads
from oracle import gprn_oracle as orc
g0 = gp.inference(1, np.arange(4.0), np.zeros(4), np.ones(4))
for n in [int(x) for x in (sys.argv[1:] or ["512", "1024", "2048"])]:
    a = workloads.synth_arrays(n, 4, 2, seed=1, node="M52")
    rng = np.random.default_rng(0)
    for spec in (a["nodes"][0], a["weights"][0]):
        A = orc.kmatrix(spec, a["t"], nugget=1e-6) + np.diag(rng.uniform(0.005, 0.05, n))
        L = np.empty((n, n)); X = np.empty((n, n)); ld = np.zeros(1)
        t0 = time.time()
        _lib.check(_lib.lib().gprn_debug_factor(g0._h(), n, _lib.dptr(_lib.f64(A)), _lib.dptr(L), _lib.dptr(X), _lib.dptr(ld)))
        dt = time.time() - t0
        Lr = np.linalg.cholesky(A)
        Xr = sla.solve_triangular(Lr, np.eye(n), lower=True)
        ldr = 2 * np.sum(np.log(np.diag(Lr)))
        g, gr = np.sum(X * X, axis=0), np.sum(Xr * Xr, axis=0)
        resL = np.linalg.norm(L @ L.T - A) / np.linalg.norm(A)
        resLr = np.linalg.norm(Lr @ Lr.T - A) / np.linalg.norm(A)
        print(f"n={n} {spec[0]}: |LL'-A|/|A| {resL:.1e} (lapack {resLr:.1e})  L {np.max(np.abs(L-Lr))/np.max(np.abs(Lr)):.1e}  "
              f"logdet abs {abs(ld[0]-ldr):.2e} rel {abs(ld[0]-ldr)/abs(ldr):.1e}  g rel {np.max(np.abs(g-gr)/gr):.1e}  "
              f"sum(g) rel {abs(g.sum()-gr.sum())/gr.sum():.1e}  |XL-I| {np.max(np.abs(X@Lr-np.eye(n))):.1e} (lapack {np.max(np.abs(Xr@Lr-np.eye(n))):.1e})  {dt:.2f}s", flush=True)
