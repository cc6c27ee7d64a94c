'''B200-native batched ODE integration engine with the call surface of numba-integrators.

    import numba_integrators_b200 as ni
    solver = ni.RK45(ni.rhs.harmonic, 0.0, y0, t_bound=1, atol=1e-8, rtol=1e-8)
    while ni.step(solver): ...

Python drives a thin C ABI (include/ni_b200.h, ctypes) into hand-written CUDA kernels for
sm_100a; there is no CPU fallback.
'''
__version__ = '0.1.0'

from . import _rhs as rhs
from . import sr
from ._lib import NiError
from ._rhs import BuiltinRHS, CudaComponentRHS, CudaRHS, CudaStencilRHS, PythonRHS
from ._translate import TranslationError, translate, translate_condition
from ._solver import (ALL, MAX_FACTOR, MIN_FACTOR, RK, RK23, RK45, SAFETY, Advanced, Condition, CudaCondition, Ensemble,
                      Records,
                      Solver, Solvers, StepMask, aux_above, aux_below, convert, ff2cond, ff2t, step, y_above,
                      y_below)

__all__ = ['ALL', 'Advanced', 'BuiltinRHS', 'Condition', 'CudaCondition', 'translate_condition', 'CudaComponentRHS', 'CudaRHS', 'CudaStencilRHS', 'PythonRHS', 'TranslationError', 'translate', 'Ensemble', 'NiError', 'RK', 'RK23', 'RK45',
           'Records', 'Solver', 'Solvers', 'StepMask', 'aux_above', 'aux_below', 'convert', 'ff2cond', 'ff2t', 'rhs', 'sr',
           'step', 'y_above', 'y_below', 'SAFETY', 'MIN_FACTOR', 'MAX_FACTOR']
