"""Latency of ONE ELBOcalc through the Python API (the C1 / C2 case: an optimiser or sampler step).

    python tools/latency_probe.py [N p reps q NODE]

Prints the wall time of inference.ELBOcalc() (host call to host result) and the device span between the events the
library records around the evaluation (gprn_last_elbo_ms): median and minimum over `reps` calls after a warm-up.
Environment switches of the library apply (GPRN_NO_LOOP, GPRN_NO_MID, GPRN_MID_WARPS, ...)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpyrn_b200 import _lib  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    q = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    node = sys.argv[5] if len(sys.argv) > 5 else "QP"
    import gpyrn_b200 as gp
    from gpyrn_b200 import covfunc, meanfunc
    import workloads
    a = workloads.synth_arrays(N, p, q, seed=1, node=node)
    ya = []
    for y, e in zip(a["y"], a["yerr"]):
        ya += [y, e]
    g = gp.inference(q, a["t"], *ya)
    K = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
    g.set_components([K[s[0]](*s[1:]) for s in a["nodes"]], [K[s[0]](*s[1:]) for s in a["weights"]],
                     [meanfunc.Constant(0.0)] * p, list(a["jitters"]))
    L = _lib.lib()
    import ctypes
    L.gprn_last_elbo_ms.restype = ctypes.c_double
    for _ in range(5):
        elbo, _, _, it = g.ELBOcalc()
    wall, dev = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        e2, _, _, it2 = g.ELBOcalc()
        wall.append(1e3 * (time.perf_counter() - t0))
        dev.append(float(L.gprn_last_elbo_ms(g._h())))
        assert e2 == elbo and it2 == it
    sw = {k: os.environ[k] for k in os.environ if k.startswith("GPRN_")}
    print(f"N={N} p={p} q={q} iterations={it} ELBO={elbo!r} switches={sw}")
    print(f"  ELBOcalc wall ms: median {np.median(wall):.3f}  min {np.min(wall):.3f}  max {np.max(wall):.3f}")
    print(f"  device span ms  : median {np.median(dev):.3f}  min {np.min(dev):.3f}")
    g.close()


if __name__ == "__main__":
    main()
