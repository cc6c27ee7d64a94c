"""`ni.sr`: the reference's StructRef flavour of RK45 (`_structref.py:19-144`): the same algorithm with the
state fields named `x` / `x_bound` and a free `step(state)` function.  Like `step_advanced` (and unlike the basic
jitclass solver) it applies `direction` to the step (`_structref.py:37`), so backward integration works.

One documented deviation: the kernels behind this facade are the `step_advanced` ones, whose terminal `step()` call (the
one that finds `t == t_bound`) overwrites `step_size` with the proposed next step (`_advanced.py:151`); the reference's
StructRef `step` returns False there without touching the state (`_structref.py:21-22`).  After
`while sr.step(s): pass` every other field equals the reference's; `s.step_size` is the unsigned proposal, not the last
signed step."""
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
