"""Mean functions m(t) of the GPRN outputs.

Same public surface as the reference's ``gpyrn.meanfunc`` (gpyrn/meanfunc.py:9-273): ``Constant``,
``MultiConstant``, ``Linear``, ``Parabola``, ``Cubic``, ``Sine`` and their ``+`` / ``*`` compositions.
These are O(N) evaluations of arbitrary Python compositions and stay on the host by design
(SURVEY.md section 2 row 3): the inference engine evaluates them once per call and ships
``y - m(t)`` (p x N doubles) to the device.
"""
import numpy as np

__all__ = ['Constant', 'MultiConstant', 'Linear', 'Parabola', 'Cubic', 'Sine']


class meanFunction:
    """Base class: parameter vector + evaluation on an array of times."""
    _parsize = 0
    _param_names = ()

    def __init__(self, *pars):
        self.pars = np.array(pars, dtype=float)

    def _eval(self, t):
        raise NotImplementedError

    def __call__(self, t):
        return self._eval(np.atleast_1d(t))

    def get_parameters(self):
        return self.pars

    def set_parameters(self, p):
        """Consume ``pars.size`` leading values of ``p``; return the remainder when there is one."""
        p = np.atleast_1d(np.asarray(p, dtype=float))
        n = self.pars.size
        assert p.size >= n, f'too few parameters for mean {self.__class__.__name__}'
        self._assign(p[:n])
        if p.size > n:
            return p[n:].copy()

    def _assign(self, values):
        self.pars = np.array(values, dtype=float)

    def __add__(self, b):
        return Sum(self, b)

    __radd__ = __add__

    def __mul__(self, b):
        return Product(self, b)

    __rmul__ = __mul__

    def __repr__(self):
        return f"{self.__class__.__name__}({', '.join(map(str, self.pars))})"


class _binary(meanFunction):
    _symbol = '?'

    def __init__(self, m1, m2):
        self.m1, self.m2 = m1, m2
        n1, n2 = list(m1._param_names), list(m2._param_names)
        if self._number_same_class and m1.__class__ == m2.__class__:
            n1, n2 = [f'{n}1' for n in n1], [f'{n}2' for n in n2]
        self._param_names = tuple(n1 + n2)
        self._parsize = m1._parsize + m2._parsize
        self.pars = np.r_[m1.pars, m2.pars]

    def _assign(self, values):
        self.pars = np.array(values, dtype=float)
        n1 = self.m1.pars.size
        self.m1._assign(self.pars[:n1])
        self.m2._assign(self.pars[n1:])

    def __repr__(self):
        return f"{self.m1} {self._symbol} {self.m2}"


class Sum(_binary):
    """m1(t) + m2(t)"""
    _symbol = '+'
    _number_same_class = True

    def _eval(self, t):
        return self.m1(t) + self.m2(t)


class Product(_binary):
    """m1(t) * m2(t)"""
    _symbol = '*'
    _number_same_class = False

    def _eval(self, t):
        return self.m1(t) * self.m2(t)


class Constant(meanFunction):
    """m(t) = c"""
    _param_names = 'c',
    _parsize = 1

    def __init__(self, c: float):
        super().__init__(c)

    def _eval(self, t):
        return np.full(t.shape, self.pars[0])


class MultiConstant(meanFunction):
    """Per-instrument constant: offsets relative to the last instrument plus its mean value.

    Args:
        offsets: [off_1, ..., off_{n-1}, mean_n]
        obsid:   one-based instrument index of every observation
        time:    observation times (same size as ``obsid``)
    """

    def __init__(self, offsets, obsid, time):
        self.obsid = obsid
        self.time = time
        self._parsize = int((np.ediff1d(obsid) == 1).sum() + 1)
        self.ii = obsid.astype(int) - 1
        if isinstance(offsets, float):
            offsets = [offsets]
        assert len(offsets) == self._parsize, \
            f'wrong number of parameters, expected {self._parsize} got {len(offsets)}'
        super().__init__(*offsets)
        self._param_names = [f'off{i}' for i in range(1, self._parsize)] + ['mean']

    def time_bins(self):
        first = self.time[np.ediff1d(self.obsid, 0, None) != 0]
        last = self.time[np.ediff1d(self.obsid, None, 0) != 0]
        return np.sort(np.r_[self.time[0], np.mean((first, last), axis=0)])

    def _eval(self, t):
        offsets = np.pad(self.pars[:-1], (0, 1))
        ii = self.ii if t.size == self.time.size else np.digitize(t, self.time_bins()) - 1
        return np.full_like(t, self.pars[-1]) + np.take(offsets, ii)


class Linear(meanFunction):
    """m(t) = slope * (t - mean(t)) + intercept   (centred on the mean of the times it is called with)"""
    _param_names = ('slope', 'intercept')
    _parsize = 2

    def __init__(self, slope: float, intercept: float):
        super().__init__(slope, intercept)

    def _eval(self, t):
        return self.pars[0] * (t - t.mean()) + self.pars[1]


class Parabola(meanFunction):
    """m(t) = quad * t**2 + slope * t + intercept"""
    _param_names = ('slope', 'intercept', 'quadratic')
    _parsize = 3

    def __init__(self, quad: float, slope: float, intercept: float):
        super().__init__(quad, slope, intercept)

    def _eval(self, t):
        return np.polyval(self.pars, t)


class Cubic(meanFunction):
    """m(t) = cub * t**3 + quad * t**2 + slope * t + intercept"""
    _param_names = ('cub', 'quad', 'slope', 'intercept')
    _parsize = 4

    def __init__(self, cub: float, quad: float, slope: float, intercept: float):
        super().__init__(cub, quad, slope, intercept)

    def _eval(self, t):
        return np.polyval(self.pars, t)


class Sine(meanFunction):
    """m(t) = amplitude * sin(2 pi t / period + phase)"""
    _param_names = ('amplitude', 'period', 'phase')
    _parsize = 3

    def __init__(self, amplitude: float, period: float, phase: float):
        super().__init__(amplitude, period, phase)

    def _eval(self, t):
        amp, per, ph = self.pars
        return amp * np.sin((2 * np.pi * t / per) + ph)
