"""Covariance functions (kernels) for GPRN nodes and weights.

Same public surface as the reference's ``gpyrn.covfunc`` for the kernels on the hot path
(``SquaredExponential``, ``Periodic``, ``QuasiPeriodic``, ``RationalQuadratic``, ``Matern32``,
``Matern52``, ``WhiteNoise`` and their ``+`` / ``*`` compositions; reference gpyrn/covfunc.py:5-80,
128-288, 355-396), plus the "next" kernels of SURVEY.md 8(f).3 (``Constant``, ``RQP``, ``Cosine``,
``Exponential``; covfunc.py:107-125, 291-352; ``Derivative`` of SE / Periodic / QuasiPeriodic, covfunc.py:80-104): ``.pars`` float64 array, ``_param_names``, ``_tag``, ``get_parameters`` /
``set_parameters``, ``k(r)`` on an array of lags.

The objects hold parameters and structure only.  All arithmetic happens on the GPU: ``k(r)`` ships
``r`` through the C ABI (``gprn_keval``), and the inference engine serialises the kernel into a
postfix program (``program()``) that the assembly kernels interpret.  The other kernels of the
reference (Linear, GammaExp, Polynomial, Piecewise, Paciorek, the *Periodic variants ... -- SURVEY.md section 2
row 2) are outside
the hot-path scope of this package.
"""
import numpy as np

from . import _lib

# opcodes of include/gprn_b200.h
OP_SE, OP_PER, OP_QP, OP_RQ, OP_M32, OP_M52, OP_WN, OP_ADD, OP_MUL = 1, 2, 3, 4, 5, 6, 7, 100, 101
OP_CONST, OP_RQP, OP_COS, OP_EXP = 8, 9, 10, 11
OP_DSE, OP_DPER, OP_DQP = 12, 13, 14


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


# ---------------------------------------------------------------------------------------------
# Kernels of the reference that are outside the hot-path scope (Linear, GammaExp,
# Polynomial, Piecewise, Paciorek, the *Periodic variants; several of them are broken in the reference itself).
# They exist as names so that user code fails with a clear message instead of an AttributeError.
# ---------------------------------------------------------------------------------------------
def _out_of_scope(name, ref_line):
    def __init__(self, *args):
        raise NotImplementedError(
            f"covfunc.{name} (reference gpyrn/covfunc.py:{ref_line}) has no device program in gpyrn_b200: the "
            "B200 path covers SquaredExponential, Periodic, QuasiPeriodic, RationalQuadratic, Matern32, Matern52, "
            "WhiteNoise, Constant, RQP, Cosine, Exponential, Derivative(SE/Periodic/QuasiPeriodic), Sum and "
            "Multiplication.  There is no CPU fallback.")
    return type(name, (covFunction,), {"__init__": __init__, "__doc__": f"Not available on the B200 path ({name})."})


for _name, _line in (("Linear", 399), ("GammaExp", 415), ("Polynomial", 435), ("Piecewise", 458), ("Paciorek", 477),
                     ("NewPeriodic", 499), ("QuasiNewPeriodic", 522), ("NewRQP", 549), ("HarmonicPeriodic", 579),
                     ("QuasiHarmonicPeriodic", 610), ("CosPeriodic", 645), ("QuasiCosPeriodic", 668)):
    globals()[_name] = _out_of_scope(_name, _line)
