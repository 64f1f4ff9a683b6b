"""Covariance functions (kernels) for GPRN nodes and weights.

Same public surface as the reference's ``gpyrn.covfunc`` for the kernels on the hot path
(``SquaredExponential``, ``Periodic``, ``QuasiPeriodic``, ``RationalQuadratic``, ``Matern32``,
``Matern52``, ``WhiteNoise`` and their ``+`` / ``*`` compositions; reference gpyrn/covfunc.py:5-80,
128-288, 355-396), plus the "next" kernels of SURVEY.md 8(f).3 (``Constant``, ``RQP``, ``Cosine``,
``Exponential``; covfunc.py:107-125, 291-352; ``Derivative`` of SE / Periodic / QuasiPeriodic, covfunc.py:80-104) and
the stationary kernels of the reference's "other" group that work there (``GammaExp``, ``Piecewise``, ``Paciorek``,
``NewPeriodic``, ``QuasiNewPeriodic``, ``CosPeriodic``, ``QuasiCosPeriodic``; covfunc.py:415-432, 458-546, 645-688):
``.pars`` float64 array, ``_param_names``, ``_tag``, ``get_parameters`` / ``set_parameters``, ``k(r)`` on an array of
lags.

The objects hold parameters and structure only.  All arithmetic happens on the GPU: ``k(r)`` ships
``r`` through the C ABI (``gprn_keval``), and the inference engine serialises the kernel into a
postfix program (``program()``) that the assembly kernels interpret.  The remaining kernels of the
reference -- ``Linear``, ``Polynomial``, ``HarmonicPeriodic``, ``QuasiHarmonicPeriodic`` (functions of ``(t1, t2)``
that the reference's own ``inference`` cannot call) and ``NewRQP`` (raises in the reference: ``np.sine``) -- exist as
names that raise ``NotImplementedError``.
"""
import numpy as np

from . import _lib

# opcodes of include/gprn_b200.h
OP_SE, OP_PER, OP_QP, OP_RQ, OP_M32, OP_M52, OP_WN, OP_ADD, OP_MUL = 1, 2, 3, 4, 5, 6, 7, 100, 101
OP_CONST, OP_RQP, OP_COS, OP_EXP = 8, 9, 10, 11
OP_DSE, OP_DPER, OP_DQP = 12, 13, 14
OP_GEXP, OP_PIECE, OP_PAC, OP_NPER, OP_QNPER, OP_COSP, OP_QCOSP = 15, 16, 17, 18, 19, 20, 21


default_device = -1     # GPU used by k(r) on the host-side kernel objects; -1: the thread's current device


class covFunction:
    """Base class: a parameter vector plus a device program."""
    _opcode = None
    _param_names = ()

    def __init__(self, *args):
        self.pars = np.array(args, dtype=float)

    # ---- structure -------------------------------------------------------------------------
    def program(self):
        """Postfix opcode list understood by the CUDA kernel interpreter."""
        if self._opcode is None:
            raise NotImplementedError
        return [self._opcode]

    # ---- evaluation (device) ---------------------------------------------------------------
    def __call__(self, r, t1=None, t2=None):
        r = np.asarray(r, dtype=float)
        shape = r.shape
        r2 = _lib.f64(np.atleast_2d(r))
        square = int(r.ndim == 2 and shape[0] == shape[1])
        prog = np.array(self.program(), dtype=np.int32)
        pars = _lib.f64(self.pars)
        out = np.empty_like(r2)
        # device -1: the calling thread's current CUDA device (the GPU of the inference object created last)
        _lib.check(_lib.lib().gprn_keval(default_device, _lib.iptr(prog), prog.size, _lib.dptr(pars), pars.size,
                                         _lib.dptr(r2), r2.shape[0], r2.shape[1], square, _lib.dptr(out)))
        return out.reshape(shape)

    # ---- parameters ------------------------------------------------------------------------
    def get_parameters(self):
        return self.pars

    def set_parameters(self, p):
        """Take ``self.pars.size`` values from the front of ``p``; return what is left, if anything."""
        p = np.atleast_1d(np.asarray(p, dtype=float))
        n = self.pars.size
        assert p.size >= n, f'too few parameters for kernel {self.__class__.__name__}'
        self._assign(p[:n])
        if p.size > n:
            return p[n:].copy()

    def _assign(self, values):
        self.pars = np.array(values, dtype=float)

    # ---- composition -----------------------------------------------------------------------
    def __add__(self, b):
        return Sum(self, b)

    __radd__ = __add__

    def __mul__(self, b):
        return Multiplication(self, b)

    __rmul__ = __mul__

    def __repr__(self):
        if self._param_names:
            inner = ', '.join(f'{k}={v}' for k, v in zip(self._param_names, self.pars))
        else:
            inner = ', '.join(str(v) for v in self.pars)
        return f"{self.__class__.__name__}({inner})"


class _operator(covFunction):
    """Binary composition; ``pars`` is the concatenation of the operands' parameters."""
    _symbol = '?'
    _join = None

    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2
        self.kerneltype = 'complex'
        self.pars = np.r_[k1.pars, k2.pars]

    def program(self):
        return self.k1.program() + self.k2.program() + [self._join]

    def _assign(self, values):
        # unlike the reference (quirk Q11) the operands follow the composite's parameters
        self.pars = np.array(values, dtype=float)
        n1 = self.k1.pars.size
        self.k1._assign(self.pars[:n1])
        self.k2._assign(self.pars[n1:])

    def __repr__(self):
        return f"{self.k1} {self._symbol} {self.k2}"


class Sum(_operator):
    """k1(r) + k2(r)"""
    _symbol, _join = '+', OP_ADD


class Multiplication(_operator):
    """k1(r) * k2(r)"""
    _symbol, _join = '*', OP_MUL


class Derivative(covFunction):
    r""":math:`\partial^2 k / \partial x_i \partial x_j` of a twice-differentiable kernel (SquaredExponential, Periodic,
    QuasiPeriodic); shares the parameters of ``k`` (reference covfunc.py:80-104)."""

    def __init__(self, k):
        if not getattr(k, '_twice_differentiable', False) or getattr(k, '_dopcode', None) is None:
            raise ValueError(f'kernel {k} is not twice differentiable')
        self.k = k
        self.kerneltype = 'complex_unary'
        self.pars = self.k.pars
        self._param_names = self.k._param_names
        self._tag = 'd' + self.k._tag

    def program(self):
        return [self.k._dopcode]

    def _assign(self, values):
        self.pars = np.array(values, dtype=float)
        self.k._assign(self.pars)

    def __repr__(self):
        return f"d {self.k}"


class Constant(covFunction):
    r""":math:`K_{ij} = c^2`"""
    _param_names = 'c',
    _tag = 'C'
    _opcode = OP_CONST

    def __init__(self, c: float):
        super().__init__(c)


class RQP(covFunction):
    r"""Periodic times rational quadratic:
    :math:`\theta^2 \exp[-2\sin^2(\pi r/P)/\ell_p^2]\,[1 + r^2/(2\alpha\ell_e^2)]^{-\alpha}`"""
    _param_names = 'theta', 'alpha', 'elle', 'ellp', 'P'
    _tag = 'RQP'
    _opcode = OP_RQP

    def __init__(self, theta: float, alpha: float, elle: float, P: float, ellp: float):
        super().__init__(theta, alpha, elle, P, ellp)


class Cosine(covFunction):
    r""":math:`\theta^2 \cos(2\pi |r| / P)`"""
    _param_names = 'theta', 'P'
    _tag = 'COS'
    _opcode = OP_COS

    def __init__(self, theta: float, P: float):
        super().__init__(theta, P)


class Exponential(covFunction):
    r""":math:`\theta^2 \exp(-|r|/\ell)`"""
    _param_names = 'theta', 'ell'
    _tag = 'EXP'
    _opcode = OP_EXP

    def __init__(self, theta: float, ell: float):
        super().__init__(theta, ell)


class WhiteNoise(covFunction):
    r"""White noise: :math:`K_{ij} = w^2 \delta_{ij}` (identity by position for square arguments)."""
    _param_names = 'wn',
    _tag = 'WN'
    _opcode = OP_WN

    def __init__(self, w: float):
        super().__init__(w)


class SquaredExponential(covFunction):
    r""":math:`\theta^2 \exp[-(t_i-t_j)^2 / (2\ell^2)]`"""
    _param_names = 'theta', 'ell'
    _tag = 'SE'
    _opcode = OP_SE
    _dopcode = OP_DSE
    _twice_differentiable = True

    def __init__(self, theta: float, ell: float):
        super().__init__(theta, ell)


class Periodic(covFunction):
    r""":math:`\theta^2 \exp[-2 \sin^2(\pi (t_i-t_j)/P) / \ell^2]`"""
    _param_names = 'theta', 'P', 'ell'
    _tag = 'P'
    _opcode = OP_PER
    _dopcode = OP_DPER
    _twice_differentiable = True

    def __init__(self, theta: float, P: float, ell: float):
        super().__init__(theta, P, ell)


class QuasiPeriodic(covFunction):
    r""":math:`\theta^2 \exp[-(t_i-t_j)^2/(2\ell_e^2) - 2\sin^2(\pi (t_i-t_j)/P)/\ell_p^2]`"""
    _param_names = 'theta', 'le', 'P', 'lp'
    _tag = 'QP'
    _opcode = OP_QP
    _dopcode = OP_DQP
    _twice_differentiable = True

    def __init__(self, theta: float, elle: float, P: float, ellp: float):
        super().__init__(theta, elle, P, ellp)


class RationalQuadratic(covFunction):
    r""":math:`\theta^2 [1 + (t_i-t_j)^2 / (2\alpha\ell^2)]^{-\alpha}`"""
    _param_names = 'theta', 'alpha', 'ell'
    _tag = 'RQ'
    _opcode = OP_RQ

    def __init__(self, theta: float, alpha: float, ell: float):
        super().__init__(theta, alpha, ell)


class Matern32(covFunction):
    r""":math:`\theta^2 (1 + \sqrt3 |r|/\ell) \exp(-\sqrt3 |r|/\ell)`"""
    _param_names = 'theta', 'ell'
    _tag = 'M32'
    _opcode = OP_M32

    def __init__(self, theta: float, ell: float):
        super().__init__(theta, ell)


class Matern52(covFunction):
    r""":math:`\theta^2 (1 + \sqrt5 |r|/\ell + 5 r^2/(3\ell^2)) \exp(-\sqrt5 |r|/\ell)`"""
    _param_names = 'theta', 'ell'
    _tag = 'M52'
    _opcode = OP_M52

    def __init__(self, theta: float, ell: float):
        super().__init__(theta, ell)


class GammaExp(covFunction):
    r""":math:`\theta^2 \exp[-(|r|/\ell)^\gamma]`, :math:`0 < \gamma \le 2` (reference covfunc.py:415-432)"""
    _param_names = 'theta', 'gamma', 'ell'
    _tag = 'GammaExp'
    _opcode = OP_GEXP

    def __init__(self, theta: float, gamma: float, l: float):
        super().__init__(theta, gamma, l)


class Piecewise(covFunction):
    r"""Third-order piecewise polynomial with compact support: with :math:`x = |r| / (\eta/2)`,
    :math:`(3x + 1)(1 - x)^3` for :math:`x \le 1`, else 0 (reference covfunc.py:458-474)"""
    _param_names = 'eta',
    _tag = 'PW'
    _opcode = OP_PIECE

    def __init__(self, eta: float):
        super().__init__(eta)


class Paciorek(covFunction):
    r"""Stationary version of Paciorek's kernel:
    :math:`a^2 \sqrt{2\ell_1\ell_2/(\ell_1^2+\ell_2^2)}\,\exp[-2r^2/(\ell_1^2+\ell_2^2)]` (reference covfunc.py:477-496).
    The reference evaluates from the constructor arguments and ignores later ``set_parameters``; here the kernel follows
    ``pars`` like every other one."""
    _param_names = 'amplitude', 'ell_1', 'ell_2'
    _tag = 'PAC'
    _opcode = OP_PAC

    def __init__(self, amplitude: float, ell_1: float, ell_2: float):
        super().__init__(amplitude, ell_1, ell_2)


class NewPeriodic(covFunction):
    r"""Rational quadratic mapped to :math:`u(x) = (\cos x, \sin x)`:
    :math:`a^2 [1 + 2\sin^2(\pi|r|/P)/(\alpha\ell^2)]^{-\alpha}` (reference covfunc.py:499-519)"""
    _param_names = 'amplitude', 'alpha2', 'P', 'ell'
    _tag = 'NP'
    _opcode = OP_NPER

    def __init__(self, amplitude: float, alpha2: float, P: float, l: float):
        super().__init__(amplitude, alpha2, P, l)


class QuasiNewPeriodic(covFunction):
    r"""``NewPeriodic`` times a squared exponential (reference covfunc.py:522-546)"""
    _param_names = 'amplitude', 'alpha2', 'ell_e', 'P', 'ell_p'
    _tag = 'QNP'
    _opcode = OP_QNPER

    def __init__(self, amplitude: float, alpha2: float, ell_e: float, P: float, ell_p: float):
        super().__init__(amplitude, alpha2, ell_e, P, ell_p)


class CosPeriodic(covFunction):
    r""":math:`a^2 \exp[-2\cos^2(\pi|r|/P)/\ell^2]` (reference covfunc.py:645-665).  The reference registers only
    ``(P, ell)`` as ``pars`` (the amplitude is not a parameter there); here all three constructor arguments are."""
    _param_names = 'amplitude', 'P', 'ell'
    _tag = 'CP'
    _opcode = OP_COSP

    def __init__(self, amplitude: float, P: float, ell: float):
        super().__init__(amplitude, P, ell)


class QuasiCosPeriodic(covFunction):
    r""":math:`a^2 \exp[-2\cos^2(\pi|r|/P)/\ell_p^2 - r^2/(2\ell_e^2)]` (reference covfunc.py:668-688)"""
    _param_names = 'amplitude', 'ell_e', 'P', 'ell_p'
    _tag = 'QCP'
    _opcode = OP_QCOSP

    def __init__(self, amplitude: float, ell_e: float, P: float, ell_p: float):
        super().__init__(amplitude, ell_e, P, ell_p)


# ---------------------------------------------------------------------------------------------
# Kernels of the reference that cannot run on its own inference path either: Linear, Polynomial and the Harmonic
# kernels are functions of (t1, t2) while inference calls kernel(r); NewRQP raises (np.sine).  They exist as names so
# that user code fails with a clear message instead of an AttributeError.
# ---------------------------------------------------------------------------------------------
def _out_of_scope(name, ref_line):
    def __init__(self, *args):
        raise NotImplementedError(
            f"covfunc.{name} (reference gpyrn/covfunc.py:{ref_line}) has no device program in gpyrn_b200: the "
            "B200 path covers SquaredExponential, Periodic, QuasiPeriodic, RationalQuadratic, Matern32, Matern52, "
            "WhiteNoise, Constant, RQP, Cosine, Exponential, Derivative(SE/Periodic/QuasiPeriodic), GammaExp, Piecewise, "
            "Paciorek, NewPeriodic, QuasiNewPeriodic, CosPeriodic, QuasiCosPeriodic, Sum and Multiplication.  "
            "There is no CPU fallback.")
    return type(name, (covFunction,), {"__init__": __init__, "__doc__": f"Not available on the B200 path ({name})."})


for _name, _line in (("Linear", 399), ("Polynomial", 435), ("NewRQP", 549), ("HarmonicPeriodic", 579),
                     ("QuasiHarmonicPeriodic", 610)):
    globals()[_name] = _out_of_scope(_name, _line)
