"""Phase breakdown of mid_pipeline_kernel (mid.cuh) on one C2-size ELBOcalc (development aid; run on the GPU box).

    python tools/mid_phases.py [N p]

Uses the -DGPRN_TRACE library of tools/trace_run.py: thread 0 of every CTA accumulates clock64 by phase; printed per
tile row (Cholesky) / column (inverse) in microseconds per launch at 1.965 GHz."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpyrn_b200 import _lib  # noqa: E402
import trace_run  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    _lib.LIB_PATH = trace_run.build_trace_lib()
    import gpyrn_b200 as gp
    from gpyrn_b200 import covfunc, meanfunc
    from oracle import gprn_oracle as orc
    m = orc.synth(N, p, 1, seed=1, node="QP")
    args = []
    for y, e in zip(m.y, m.yerr):
        args += [y, e]
    g = gp.inference(1, m.time, *args)
    K = {"QP": covfunc.QuasiPeriodic, "M52": covfunc.Matern52, "SE": covfunc.SquaredExponential}
    g.set_components([K[s[0]](*s[1:]) for s in m.nodes], [K[s[0]](*s[1:]) for s in m.weights],
                     [meanfunc.Constant(0.0)] * p, [0.1] * p)
    L = _lib.lib()
    L.gprn_trace_mid_phases.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    buf = (ctypes.c_ulonglong * 256)()
    g.ELBOcalc()
    _lib.check(L.gprn_trace_mid_phases(buf))       # reset after the warm-up
    _, _, _, it = g.ELBOcalc()
    _lib.check(L.gprn_trace_mid_phases(buf))
    ph = np.array(list(buf), dtype=np.float64).reshape(2, 16, 8)
    names = ["other", "flag waits", "K loads", "products", "potrf64", "solves", "stores+publish", "inverse reductions"]
    M = 1 + p
    for mode, label, launches in ((0, "set-up launch (Cholesky only)", M), (1, "iteration launches (Cholesky + inverse)", it * M)):
        print(f"{label}: thread-0 microseconds per CTA by phase ({launches} matrices)")
        print("   row " + " ".join(f"{n[:9]:>10s}" for n in names) + "      total")
        for r in range((N + 63) // 64):
            us = ph[mode, r] / launches / 1965.0
            print(f"   {r:3d} " + " ".join(f"{v:10.1f}" for v in us) + f" {us.sum():10.1f}")
    g.close()


if __name__ == "__main__":
    main()
