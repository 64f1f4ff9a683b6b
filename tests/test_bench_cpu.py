"""CPU checks of bench.py's bookkeeping: the flop model of the roofline, the workload / config dictionaries shared by
both arms, and the clock sampler's handling of samples taken before / inside the timed region (no GPU, no nvidia-smi)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
bench = importlib.import_module("bench")


def test_algorithmic_flops_follow_the_survey_formula():
    # F_eval = M N^3 / 3 + n_it [(2/3) M + q(q-1)/2] N^3
    N, p, q = 4096, 4, 2
    M = q * (p + 1)
    its = np.array([44, 61])
    want = sum(M * N ** 3 / 3 + n * ((2 / 3) * M + 1.0) * N ** 3 for n in its)
    assert np.isclose(bench.algorithmic_flops(N, p, q, its), want, rtol=1e-15)
    lo = bench.algorithmic_flops(N, p, q, its, cross=2.0 / 3.0)            # what cross_frob_kernel executes
    assert lo < want and np.isclose(want - lo, its.sum() * (1.0 / 3.0) * N ** 3, rtol=1e-12)
    assert bench.algorithmic_flops(256, 4, 1, [10]) == 5 * 256 ** 3 / 3 + 10 * (2 / 3) * 5 * 256 ** 3      # q = 1: no cross term


def test_workloads_name_the_baseline_configurations():
    w = bench.WORKLOADS
    assert (w["c4"]["N"], w["c4"]["p"], w["c4"]["q"], w["c4"]["node"]) == (4096, 4, 2, "M52")
    assert (w["c3"]["N"], w["c3"]["p"], w["c3"]["q"], w["c3"]["pool_per_gpu"]) == (256, 4, 1, 8192)
    assert (w["c2"]["N"], w["c2"]["p"], w["c2"]["q"], w["c2"]["pool_per_gpu"]) == (500, 4, 1, 1)
    assert (w["c5"]["N"], w["c5"]["p"], w["c5"]["q"]) == (2048, 4, 2)

    class A:
        scaling, pool = "weak", 64
    cfg1, cfg8 = bench.config_of(A, w["c4"], 1), bench.config_of(A, w["c4"], 8)
    assert cfg1["global_sets"] == 12 and cfg8["global_sets"] == 96 and cfg1["workload"] == cfg8["workload"]
    assert "model" not in cfg1                                             # a workload, not a neural model
    rec = bench.reference_iterations(w["c4"])                              # converged record of the unmodified reference
    assert rec is not None and rec[0] == 44


def test_clock_sampler_reports_only_the_timed_region():
    s = bench.ClockSampler(0)                      # never started: rows are filled by hand
    row = lambda mhz, cap: ["0", str(mhz), "1965", "500.0", "0x0", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [row(1200, "Not Active"), row(1300, "Active")]          # warm-up: a power cap that must not be reported
    s.mark()
    s.rows += [row(1965, "Not Active"), row(1950, "Not Active"), row(1965, "Not Active")]
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == []
    s2 = bench.ClockSampler(0)                     # a timed region shorter than the sampling period: the last sample
    s2.rows = [row(1800, "Not Active"), row(1965, "Active")]
    s2.mark()
    out2 = s2.stop()
    assert out2["samples"] == 1 and out2["sm_mhz"] == 1965.0 and out2["reasons"] == ["sw_power_cap"]
