"""C5-size check: ELBOcalc capped at 6 iterations (N=2048,p=4,q=2), prediction at T=20000 test epochs timed, and the
first 400 test epochs compared with the CPU oracle (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyrn_b200 as gp
from gpyrn_b200 import covfunc, meanfunc
from oracle import gprn_oracle as orc
N, p, q, T = 2048, 4, 2, 20000
m = orc.synth(N, p, q, seed=1, node="M52")
ya = []
for y, e in zip(m.y, m.yerr):
    ya += [y, e]
g = gp.inference(q, m.time, *ya)
g.set_components([covfunc.Matern52(*s[1:]) for s in m.nodes], [covfunc.SquaredExponential(*s[1:]) for s in m.weights],
                 [meanfunc.Constant(0.0)] * p, [0.1] * p)
t0 = time.time(); elbo, mu, var, it = g.ELBOcalc(max_iter=6); print(f"ELBOcalc(max_iter=6): {elbo!r} in {time.time()-t0:.3f} s")
span = m.time[-1] - m.time[0]
tstar = np.linspace(m.time[0] - 0.2 * span, m.time[-1] + 0.2 * span, T)
g._Prediction(tstar=tstar[:256], mu=mu, var=var)
g._Prediction(tstar=tstar, mu=mu, var=var)
t0 = time.time(); pm, pv = g._Prediction(tstar=tstar, mu=mu, var=var); dt = time.time() - t0
M = q * (p + 1)
print(f"_Prediction T={T}: {dt:.3f} s  ({M * (N**3/3 + N**2*T) / dt * 1e-12:.2f} TFLOP/s algorithmic, SURVEY 8d F_pred)")
sel = np.r_[0:200, T//2:T//2+200]
t0 = time.time(); pmo, pvo, _, _ = orc.prediction(m, tstar[sel], mu, var, None); print(f"oracle on {sel.size} epochs: {time.time()-t0:.1f} s")
print("pred mean rel", np.max(np.abs(pm[sel]-pmo))/np.max(np.abs(pmo)), "var rel", np.max(np.abs(pv[sel]-pvo))/np.max(np.abs(pvo)))
