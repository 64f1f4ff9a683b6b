"""gpyrn_b200 -- B200-native (sm_100a, FP64) mean-field GPRN inference.

Drop-in for the hot path of iastro-pt/gpyrn: ``covfunc`` / ``meanfunc`` / ``meanfield.inference``
keep the reference's Python API; the numerics run in hand-written CUDA kernels behind the C ABI of
``include/gprn_b200.h`` (``gpyrn_b200/csrc/libgprn_b200.so``).  No CPU fallback.
"""
__version__ = '0.1.0'

from .meanfunc import Constant, Linear
from .covfunc import SquaredExponential, QuasiPeriodic
from .meanfield import inference
from . import covfunc, meanfunc, meanfield, distributed, datasets

__all__ = ['Constant', 'Linear', 'SquaredExponential', 'QuasiPeriodic', 'inference',
           'covfunc', 'meanfunc', 'meanfield', 'distributed']
