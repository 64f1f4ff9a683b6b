"""The Sigma-free blocked algorithm the CUDA kernels implement (oracle/device_model.py) reproduces the
reference golden vectors on the CPU: validates the identities of SURVEY.md Appendix A.3 and the
numerical choices (substitution-based TRSM, explicit triangular inverse, fixed-point quadratic form)
independently of any GPU.  Tolerance: ELBO 1e-10 relative, identical iteration counts."""
import pytest

from oracle import device_model as dm
from tests._cases import load_golden, oracle_model, relerr

CASES = ["notebook_45_1_1", "synth_50_1_1_QP", "synth_100_4_1_QP", "synth_60_2_2_QP_means",
         "synth_100_4_2_M52_means", "mixed_80_2_2"]


@pytest.mark.parametrize("name", CASES)
def test_device_model_matches_reference(name):
    d = load_golden(name)
    m = oracle_model(d)
    elbo, mu, var, it, trace = dm.elbo_calc(m, max_iter=d["max_iter"], return_trace=True)
    assert it == d["iters"]
    assert abs(elbo - d["elbo"]) <= 1e-10 * abs(d["elbo"])
    assert relerr(trace, d["trace"]) < 1e-9
