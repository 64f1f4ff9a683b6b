import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpyrn_b200 as gp
from gpyrn_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
rng = np.random.default_rng(0)
t = np.sort(rng.uniform(0, 100, 40))
g0 = gp.inference(1, t, np.zeros_like(t), np.ones_like(t))
B = rng.standard_normal((n, n)); A = B @ B.T + n * np.eye(n)
L = np.empty((n, n)); X = np.empty((n, n)); ld = np.zeros(1)
_lib.check(_lib.lib().gprn_debug_factor(g0._h(), n, _lib.dptr(_lib.f64(A)), _lib.dptr(L), _lib.dptr(X), _lib.dptr(ld)))
Lr = np.linalg.cholesky(A)
print("ok", np.max(np.abs(L - Lr)), np.max(np.abs(X @ Lr - np.eye(n))))
