"""Small driver for profiling: python tools/probe_elbo.py N p q node B [max_iter] -- one warm-up + one timed batched call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyrn_b200 as gp
from gpyrn_b200 import _lib, covfunc, meanfunc
from oracle import gprn_oracle as orc
N, p, q, node, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
max_iter = int(sys.argv[6]) if len(sys.argv) > 6 else None
m = orc.synth(N, p, q, seed=1, node=node)
theta = orc.perturbed_hyper_sets(m, B, 101)
args = []
for y, e in zip(m.y, m.yerr):
    args += [y, e]
g = gp.inference(q, m.time, *args)
K = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
g.set_components([K[s[0]](*s[1:]) for s in m.nodes], [K[s[0]](*s[1:]) for s in m.weights],
                 [meanfunc.Constant(0.0)] * p, [0.1] * p)
P = np.concatenate([theta[:, :-p], np.zeros((B, p)), theta[:, -p:]], axis=1)
if not os.environ.get("GPRN_PROBE_NOWARM"):
    g.ELBO_batch(P, max_iter=max_iter)          # warm-up (allocations)
_lib.lib().gprn_reset_launch_count(g._h())
t0 = time.time()
elbo, iters, status = g.ELBO_batch(P, max_iter=max_iter, return_info=True)
dt = time.time() - t0
ms = _lib.lib().gprn_last_elbo_ms(g._h())
print(f"N={N} p={p} q={q} B={B}: wall {dt*1e3:.1f} ms device {ms:.1f} ms launches {_lib.lib().gprn_launch_count(g._h())} "
      f"iters mean {iters.mean():.1f} max {iters.max()} evals/s {B/dt:.1f} elbo0 {elbo[0]!r}")
