"""CPU tests of the host side: the reference's own API tests (tests/test_inference.py:8-36,
tests/test_mean_functions.py:7-35, tests/test_imports.py) re-stated against gpyrn_b200, the
parameter bookkeeping (meanfield.py:180-379), kernel-program serialisation, and the C ABI
surface (library loads and exports every symbol include/gprn_b200.h declares).  No compute calls."""
import os
import re

import numpy as np
import pytest

import gpyrn_b200
from gpyrn_b200 import _lib, covfunc, meanfunc
from gpyrn_b200.meanfield import inference
from gpyrn_b200.meanfunc import Constant, Linear

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_imports():
    assert gpyrn_b200.inference is inference
    assert gpyrn_b200.SquaredExponential is covfunc.SquaredExponential
    assert gpyrn_b200.QuasiPeriodic is covfunc.QuasiPeriodic
    assert gpyrn_b200.Constant is meanfunc.Constant and gpyrn_b200.Linear is meanfunc.Linear


def test_create_inference():
    t, y, yerr = np.random.rand(3, 10)
    gprn = inference(1, t, y, yerr)
    assert gprn.time is t
    assert gprn.N == t.size and gprn.q == 1 and gprn.p == 1
    t, y1, ye1, y2, ye2 = np.random.rand(5, 10)
    gprn = inference(1, t, y1, ye1, y2, ye2)
    assert np.allclose(gprn.y, np.c_[y1, y2].T)
    assert np.allclose(gprn.yerr2, np.c_[ye1, ye2].T ** 2)
    assert gprn.q == 1 and gprn.p == 2 and gprn.qp == 2 and gprn.d == 10 * 1 * 3


def test_create_inference_exception():
    with pytest.raises(TypeError):
        _ = inference(1)
    with pytest.raises(AssertionError):
        _ = inference(1, np.random.rand(10))
    t, y1, ye1 = np.random.rand(3, 10)
    y2, ye2 = np.random.rand(2, 20)
    with pytest.raises(AssertionError):
        _ = inference(1, t, y1, ye1, y2, ye2)


def test_set_components_forms_and_errors():
    t, y, yerr = np.random.rand(3, 10)
    gprn = inference(1, t, y, yerr)
    node, weight = covfunc.SquaredExponential(1, 1), covfunc.SquaredExponential(1, 1)
    mean = meanfunc.Constant(0)
    gprn.set_components(node, weight, mean, 0.0)
    assert gprn.nodes[0] is node
    gprn.set_components([node], [weight], mean, 0.0)
    gprn.set_components([node], [weight], [mean], [0.0])
    assert gprn.jitters.dtype == float
    with pytest.raises(ValueError):
        gprn.set_components([node, node], [weight], [mean], [0.0])
    with pytest.raises(ValueError):
        gprn.set_components([node], [weight, weight], [mean], [0.0])
    with pytest.raises(ValueError):
        inference(1, t, y, yerr).get_parameters()


def _two_output():
    t = np.linspace(0, 10, 12)
    y1, e1, y2, e2 = np.random.rand(4, 12)
    g = inference(2, t, y1, e1, y2, e2)
    nodes = [covfunc.QuasiPeriodic(1, 2, 3, 4), covfunc.Matern52(5, 6)]
    weights = [covfunc.SquaredExponential(7 + k, 20 + k) for k in range(4)]
    g.set_components(nodes, weights, [meanfunc.Constant(0.5), meanfunc.Linear(0.1, 0.2)], [0.3, 0.4])
    return g


def test_parameter_vector_order_names_and_freezing():
    g = _two_output()
    assert g.n_parameters == 4 + 2 + 8 + 1 + 2 + 2
    expected = [1, 2, 3, 4, 5, 6, 7, 20, 8, 21, 9, 22, 10, 23, 0.5, 0.1, 0.2, 0.3, 0.4]
    assert np.allclose(g.get_parameters(), expected)
    names = list(g.parameters_dict.keys())
    assert names[:6] == ['node1.theta', 'node1.le', 'node1.P', 'node1.lp', 'node2.theta', 'node2.ell']
    assert names[6:8] == ['weight1.theta', 'weight1.ell'] and names[-2:] == ['jitter1', 'jitter2']
    assert names[14:17] == ['mean1.c', 'mean2.slope', 'mean2.intercept']
    new = np.arange(19, dtype=float) + 100
    g.set_parameters(new)
    assert np.allclose(g.get_parameters(), new)
    assert np.allclose(g.nodes[1].pars, [104, 105]) and np.allclose(g.jitters, [117, 118])
    g.freeze_parameter(name='node1*')
    assert g.frozen_mask.sum() == 4 and g.get_parameters().size == 15
    g.set_parameters(np.zeros(15) + 7)
    assert np.allclose(g.nodes[0].pars, [100, 101, 102, 103]) and np.allclose(g.nodes[1].pars, 7)
    g.thaw_parameter(name='node1.P')
    assert g.frozen_mask.sum() == 3
    g.freeze_parameter(index=18)
    assert g.frozen_mask[18]
    g.thaw_all_parameters()
    assert g.frozen_mask.sum() == 0
    with pytest.raises(ValueError):
        g.set_parameters(np.zeros(3))
    with pytest.raises(NotImplementedError):
        g.frozen_mask = np.zeros(19, bool)
    with pytest.raises(AssertionError):
        g.freeze_parameter(name='nope')


def test_other_kernels_of_the_reference():
    """The stationary "other" kernels have device programs; the ones the reference's own inference cannot call
    (functions of (t1, t2)) or that raise there (NewRQP) are names that fail with a clear message."""
    k = covfunc.GammaExp(1, 1.5, 10) * covfunc.Piecewise(300) + covfunc.Paciorek(1, 2, 3)
    assert k.program() == [covfunc.OP_GEXP, covfunc.OP_PIECE, covfunc.OP_MUL, covfunc.OP_PAC, covfunc.OP_ADD]
    assert np.allclose(k.pars, [1, 1.5, 10, 300, 1, 2, 3])
    for cls, n, op in ((covfunc.NewPeriodic, 4, covfunc.OP_NPER), (covfunc.QuasiNewPeriodic, 5, covfunc.OP_QNPER),
                       (covfunc.CosPeriodic, 3, covfunc.OP_COSP), (covfunc.QuasiCosPeriodic, 4, covfunc.OP_QCOSP)):
        obj = cls(*range(1, n + 1))
        assert obj.program() == [op] and obj.pars.size == n == len(obj._param_names)
        assert obj.set_parameters(np.arange(10.0, 10.0 + n + 2)).size == 2 and obj.pars[0] == 10.0
    for name in ("Linear", "Polynomial", "NewRQP", "HarmonicPeriodic", "QuasiHarmonicPeriodic"):
        with pytest.raises(NotImplementedError):
            getattr(covfunc, name)(1.0, 2.0)


def test_kernel_programs_and_composition():
    se, per = covfunc.SquaredExponential(1, 10), covfunc.Periodic(1, 20, 0.5)
    k = se * per + covfunc.WhiteNoise(0.1)
    assert k.program() == [covfunc.OP_SE, covfunc.OP_PER, covfunc.OP_MUL, covfunc.OP_WN, covfunc.OP_ADD]
    assert np.allclose(k.pars, [1, 10, 1, 20, 0.5, 0.1])
    rest = k.set_parameters(np.arange(8.0))
    assert np.allclose(rest, [6, 7])
    assert np.allclose(se.pars, [0, 1]) and np.allclose(per.pars, [2, 3, 4])      # operands follow the composite
    assert covfunc.QuasiPeriodic(1, 2, 3, 4)._param_names == ('theta', 'le', 'P', 'lp')
    assert covfunc.Matern32(1, 2)._tag == 'M32' and covfunc.RationalQuadratic(1, 2, 3)._tag == 'RQ'
    assert 'theta=1.0' in repr(covfunc.Matern52(1, 2))
    assert covfunc.SquaredExponential(1, 2).set_parameters([3, 4]) is None
    with pytest.raises(AssertionError):
        covfunc.SquaredExponential(1, 2).set_parameters([3])


def test_Derivative_kernel_objects():
    """reference covfunc.py:80-104: only twice-differentiable kernels, tag 'd'+tag, parameters shared with k."""
    qp = covfunc.QuasiPeriodic(1, 2, 3, 4)
    d = covfunc.Derivative(qp)
    assert d._tag == 'dQP' and d._param_names == qp._param_names and d.kerneltype == 'complex_unary'
    assert d.program() == [covfunc.OP_DQP] and np.allclose(d.pars, [1, 2, 3, 4])
    assert covfunc.Derivative(covfunc.SquaredExponential(1, 2)).program() == [covfunc.OP_DSE]
    assert covfunc.Derivative(covfunc.Periodic(1, 2, 3)).program() == [covfunc.OP_DPER]
    k = d * covfunc.Matern52(1, 5)
    assert k.program() == [covfunc.OP_DQP, covfunc.OP_M52, covfunc.OP_MUL]
    k.set_parameters([5, 6, 7, 8, 9, 10])
    assert np.allclose(qp.pars, [5, 6, 7, 8])
    with pytest.raises(ValueError):
        covfunc.Derivative(covfunc.Matern52(1, 2))
    with pytest.raises(ValueError):
        covfunc.Derivative(covfunc.Periodic(1, 2, 3) + covfunc.WhiteNoise(1))


def test_Constant():
    m = Constant(0.0)
    assert m.pars[0] == 0.0 and np.all(m(np.random.rand(10)) == 0.0)
    m = Constant(10.0)
    assert m.pars[0] == 10.0 and np.all(m(np.random.rand(3)) == 10.0)
    with pytest.raises(TypeError):
        m = Constant()
    assert np.all((Constant(5.0) + Constant(10.0))(np.random.rand(3)) == 15.0)
    assert np.all((Constant(2) * Constant(10.0))(np.random.rand(3)) == 20.0)
    assert (Constant(5.0) + Constant(10.0))._param_names == ('c1', 'c2')


def test_Linear_and_polynomials():
    m = Linear(0.0, 1.0)
    assert m.pars[0] == 0.0 and m.pars[1] == 1.0 and np.all(m(np.random.rand(10)) == 1.0)
    m = Linear(1.0, 2.0)
    t = np.array([0.0, 1.0, 2.0, 3.0])
    assert np.all(m(t) == np.polyval(m.pars, t - t.mean()))
    assert np.allclose(meanfunc.Parabola(1, 2, 3)(t), t ** 2 + 2 * t + 3)
    assert np.allclose(meanfunc.Cubic(1, 0, 0, 1)(t), t ** 3 + 1)
    assert np.allclose(meanfunc.Sine(2, 4, 0.5)(t), 2 * np.sin(2 * np.pi * t / 4 + 0.5))
    s = Linear(1.0, 2.0) + meanfunc.Sine(1, 2, 3)
    assert s.set_parameters(np.arange(6.0)) .tolist() == [5.0]
    assert np.allclose(s.m2.pars, [2, 3, 4])


def test_MultiConstant():
    t = np.arange(6.0)
    obs = np.array([1, 1, 1, 2, 2, 2])
    m = meanfunc.MultiConstant([1.5, 10.0], obs, t)
    assert np.allclose(m(t), [11.5, 11.5, 11.5, 10, 10, 10])
    assert m._param_names == ['off1', 'mean']


def test_mean_vector_host():
    t = np.linspace(0, 1, 5)
    y1, e1, y2, e2 = np.random.rand(4, 5)
    g = inference(1, t, y1, e1, y2, e2)
    mv = g._mean([Constant(2.0), None])
    assert np.allclose(mv, np.r_[np.full(5, 2.0), np.zeros(5)])
    f, w = g._u_to_fhatW(np.arange(g.d, dtype=float))
    assert f.shape == (1, 1, 5) and w.shape == (2, 1, 5)


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gprn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gprn_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = _lib.lib()                       # loads without a GPU (no CUDA call at load time)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gprn_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.gprn_built_for_sm() == 100


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product must raise, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    t, y, yerr = np.random.rand(3, 10)
    g = inference(1, t, y, yerr)
    g.set_components(covfunc.SquaredExponential(1, 1), covfunc.SquaredExponential(1, 1), Constant(0), 0.1)
    with pytest.raises(_lib.GprnError):
        _ = g.ELBO
    with pytest.raises(_lib.GprnError):
        covfunc.SquaredExponential(1, 1)(np.zeros((3, 3)))


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "gpyrn_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_observation_table_loader(tmp_path):
    """loader for tables in the format of the reference's bundled Solar_observations.txt."""
    from gpyrn_b200 import datasets
    f = tmp_path / "obs.txt"
    f.write_text("BJD\tRV\tRVerr\tBIS\tBISerr\tConstrast\tContrasterr\n"
                 "3.0\t1.5\t0.1\t-8.0\t0.2\t40.0\t0.5\n"
                 "1.0\t2.5\t0.3\t-9.0\t0.4\t41.0\t0.6\n")
    tab = datasets.load_table(str(f))
    assert list(tab) == ["BJD", "RV", "RVerr", "BIS", "BISerr", "Constrast", "Contrasterr"]
    t, series = datasets.load_observations(str(f), ("RV", "BIS", "Constrast"))
    assert np.allclose(t, [1.0, 3.0]) and len(series) == 6
    assert np.allclose(series[0], [2.5, 1.5]) and np.allclose(series[1], [0.3, 0.1])
    assert np.allclose(series[4], [41.0, 40.0]) and np.allclose(series[5], [0.6, 0.5])
    g = gpyrn_b200.inference(1, t, *series)            # host-side construction only
    assert g.p == 3 and g.N == 2
    with pytest.raises(KeyError):
        datasets.load_observations(str(f), ("FWHM",))


def test_trace_analysis_tool_on_synthetic_records(tmp_path, capsys):
    """tools/trace_run.py --analyse: SM-time attribution of a per-CTA trace (no GPU needed for the analysis)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("trace_run", os.path.join(ROOT, "tools", "trace_run.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    dt = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("kid", "<i4"), ("smid", "<i4")])
    rec = np.zeros(4, dtype=dt)
    rec[0] = (1000, 101000, 5, 0)        # syrk_outer on SM 0 for 100 us
    rec[1] = (1000, 51000, 3, 1)         # trsm_col on SM 1 for 50 us
    rec[2] = (51000, 101000, 8, 1)       # trtri_outer on SM 1 for 50 us
    rec[3] = (1000, 101000, 9, 0)        # trtri_inblock co-resident with the GEMM CTA on SM 0
    tr.analyse(rec, nsm=2)
    out = capsys.readouterr().out
    assert "syrk_outer" in out and "trsm_col" in out
    gemm = float(re.search(r"GEMM-class CTA .* resident: ([0-9.]+) %", out).group(1))
    assert 70.0 < gemm < 80.0            # SM 0 always, SM 1 half of the time


def test_out_of_scope_kernels_fail_loudly():
    """the reference's kernels without a device program exist as names and raise a clear
    NotImplementedError (the stationary ones have programs since round 2: test_other_kernels_of_the_reference)."""
    for name in ("Linear", "Polynomial", "NewRQP", "HarmonicPeriodic", "QuasiHarmonicPeriodic"):
        with pytest.raises(NotImplementedError, match="no device program"):
            getattr(covfunc, name)(1.0, 2.0)


def test_reference_side_stub_is_syntactically_complete():
    """integration/gpyrn_b200_stub.py (the binding INTEGRATION.md shows) binds only symbols the header declares."""
    src = open(os.path.join(ROOT, "integration", "gpyrn_b200_stub.py")).read()
    used = set(re.findall(r"_L\.(gprn_[a-z_]+)", src))
    hdr = open(os.path.join(ROOT, "include", "gprn_b200.h")).read()
    for name in used:
        assert re.search(r"\b" + name + r"\s*\(", hdr), name
    assert {"gprn_create", "gprn_set_model", "gprn_elbo_batched", "gprn_predict"} <= used


def test_stub_programs_of_reference_kernel_objects_match_the_package():
    """The reference-side stub serialises the REFERENCE's kernel objects (by `_tag`, the stationary "other" kernels by
    class name) into the same postfix programs and parameter vectors as this package's own classes.  CPU only: needs
    the reference copy under baseline/_ref (made by __graft_entry__.build() in the build container)."""
    from tests import _ref_shim
    if not _ref_shim.available():
        pytest.skip("baseline/_ref/gpyrn not present")
    _ref_shim.install()
    from gpyrn import covfunc as ref
    from integration import gpyrn_b200_stub as stub
    pairs = [
        (ref.QuasiPeriodic(1, 30, 27, 0.7) * ref.Matern52(1.0, 40.0) + ref.WhiteNoise(0.1),
         covfunc.QuasiPeriodic(1, 30, 27, 0.7) * covfunc.Matern52(1.0, 40.0) + covfunc.WhiteNoise(0.1)),
        (ref.Derivative(ref.SquaredExponential(2.0, 9.0)), covfunc.Derivative(covfunc.SquaredExponential(2.0, 9.0))),
        (ref.GammaExp(1.1, 1.6, 120.0) * ref.Piecewise(900.0) + ref.Paciorek(0.9, 70.0, 110.0),
         covfunc.GammaExp(1.1, 1.6, 120.0) * covfunc.Piecewise(900.0) + covfunc.Paciorek(0.9, 70.0, 110.0)),
        (ref.NewPeriodic(1.0, 2.0, 50.0, 1.5) + ref.QuasiNewPeriodic(1.0, 1.5, 60.0, 25.0, 0.9) * ref.QuasiCosPeriodic(1, 2, 3, 4),
         covfunc.NewPeriodic(1.0, 2.0, 50.0, 1.5) + covfunc.QuasiNewPeriodic(1.0, 1.5, 60.0, 25.0, 0.9) * covfunc.QuasiCosPeriodic(1, 2, 3, 4)),
        (ref.RQP(1.0, 1.2, 80.0, 25.0, 0.9) + ref.Constant(0.2) * ref.Cosine(1.0, 9.0) + ref.Exponential(1.0, 50.0),
         covfunc.RQP(1.0, 1.2, 80.0, 25.0, 0.9) + covfunc.Constant(0.2) * covfunc.Cosine(1.0, 9.0) + covfunc.Exponential(1.0, 50.0)),
    ]
    for kr, ko in pairs:
        assert stub.program(kr) == ko.program()
        assert np.array_equal(np.ravel(kr.pars).astype(float), ko.pars)
    with pytest.raises(NotImplementedError):           # CosPeriodic registers only (P, ell) in the reference: not bindable
        stub.program(ref.CosPeriodic(1.0, 2.0, 3.0))
