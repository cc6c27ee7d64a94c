"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numba_integrators_b200 as ni  # noqa: E402
from numba_integrators_b200 import workloads as wl  # noqa: E402

w = wl.c3_lorenz(N=777, t_bound=0.5)
a = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 0.5, rtol=1e-8, atol=1e-8, record=50_000)
a.run()
print('lorenz', a.n_accepted.sum(), len(a.record_drain()))
b = ni.RK23(ni.rhs.vanderpol(np.linspace(0.1, 5, 333)), 0.0, wl.c4_vanderpol(N=333).y0, 2.0, rtol=1e-6, atol=1e-6,
            arithmetic='fast')
while ni.step(b):
    pass
print('vdp', b.n_accepted.sum())
h = wl.c5_heat(N=5, n=4096, t_bound=3.0)
c = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 3.0, rtol=1e-8, atol=1e-8, record=400)
c.run()
print('heat4096', c.n_accepted.sum(), len(c.record_drain()))
h = wl.c5_heat(N=7, n=1000, t_bound=3.0)
d = ni.RK23(ni.rhs.burgers1d(np.full(7, 0.5), np.full(7, 3.0)), 0.0, h.y0, 0.2, rtol=1e-6, atol=1e-8)
d.run()
print('burgers1000', d.n_accepted.sum(), d.status)
e = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), 10.0)
print('ff2t', ni.ff2t(e, 3.0), e.t, ni.ff2cond(e, ni.y_above(0), 100.0), e.t)
