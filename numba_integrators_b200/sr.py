"""`ni.sr`: the reference's StructRef flavour of RK45 (`_structref.py:19-144`): the same algorithm with the
state fields named `x` / `x_bound` and a free `step(state)` function.  Like `step_advanced` (and unlike the basic
jitclass solver) it applies `direction` to the step (`_structref.py:37`), so backward integration works."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._solver import Ensemble


class RK(Ensemble):
    x = property(lambda s: s.t)
    x_bound = property(lambda s: s.t_bound, lambda s, v: Ensemble.t_bound.fset(s, v))


def RK45(fun, x0, y0, x_bound, max_step=np.inf, rtol=1e-3, atol=1e-6, first_step=0., **kw):  # _structref.py:126-144
    """Interface for creating RK45 solver state."""
    return RK(fun, x0, y0, x_bound, max_step, rtol, atol, first_step, _lib.RK45, advanced=True, **kw)


def step(state) -> bool:  # _structref.py:19-77
    return state.step()
