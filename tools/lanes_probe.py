"""Experiment: does running two half-batches on two handles (two independent stream sets) concurrently beat one
lock-step batch?  python tools/lanes_probe.py N p q NODE B MAX_ITER [delay_ms]"""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyrn_b200 as gp
from gpyrn_b200 import covfunc, meanfunc
from oracle import gprn_oracle as orc
N, p, q, node, B, max_iter = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
delay = float(sys.argv[7]) * 1e-3 if len(sys.argv) > 7 else 0.0
m = orc.synth(N, p, q, seed=1, node=node)
theta = orc.perturbed_hyper_sets(m, B, 101)
K = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
def make():
    args = []
    for y, e in zip(m.y, m.yerr):
        args += [y, e]
    g = gp.inference(q, m.time, *args)
    g.set_components([K[s[0]](*s[1:]) for s in m.nodes], [K[s[0]](*s[1:]) for s in m.weights],
                     [meanfunc.Constant(0.0)] * p, [0.1] * p)
    return g
P = np.concatenate([theta[:, :-p], np.zeros((B, p)), theta[:, -p:]], axis=1)
g0, g1 = make(), make()
g0.ELBO_batch(P, max_iter=max_iter); g1.ELBO_batch(P[:B // 2], max_iter=max_iter)
t0 = time.time(); e_one = g0.ELBO_batch(P, max_iter=max_iter); t_one = time.time() - t0
res = [None, None]
def run(i, g, rows, d):
    time.sleep(d)
    res[i] = g.ELBO_batch(rows, max_iter=max_iter)
for rep in range(2):
    th = [threading.Thread(target=run, args=(0, g0, P[:B // 2], 0.0)), threading.Thread(target=run, args=(1, g1, P[B // 2:], delay))]
    t0 = time.time(); [t.start() for t in th]; [t.join() for t in th]; t_two = time.time() - t0
    same = np.array_equal(np.concatenate(res), e_one)
    print(f"N={N} B={B} it={max_iter}: one batch {t_one*1e3:.1f} ms, two concurrent halves (delay {delay*1e3:.0f} ms) {t_two*1e3:.1f} ms, identical {same}")
