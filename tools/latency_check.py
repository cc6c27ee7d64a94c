"""Config C1 (README sine wave, ONE trajectory): latency of the GPU path, for the record.
A single trajectory cannot fill a GPU; this is the reference's home turf (2 us/step in numba)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
_t = time.perf_counter()
import numba_integrators_b200 as ni  # noqa: E402
t_import = time.perf_counter() - _t

y0 = np.array((0., 1.))

# The reference's own start-up benchmark (tests/benchmarking.py:20-87, published in tests/benchmarking.yaml:29-32 for
# v0.3.1: import 268 ms, first initialisation 1.7 s, second 778 ms, first step 617 ms -- numba JIT).  Here: the
# built-in RHS needs no compilation; a Python RHS is translated and compiled once with NVRTC (cached per process).


def _f(t, y):
    return np.array((y[1], -y[0]))


def _startup(rhs, label):
    t0 = time.perf_counter()
    s = ni.RK45(rhs, 0.0, y0, t_bound=1, atol=1e-8, rtol=1e-8)
    t1 = time.perf_counter()
    s2 = ni.RK45(rhs, 0.0, y0, t_bound=1, atol=1e-8, rtol=1e-8)
    t2 = time.perf_counter()
    ni.step(s)
    t3 = time.perf_counter()
    ni.step(s)
    t4 = time.perf_counter()
    print(f'start-up, {label}: first initialisation {1e3 * (t1 - t0):.1f} ms, second {1e3 * (t2 - t1):.1f} ms, '
          f'first step {1e3 * (t3 - t2):.2f} ms, second step {1e3 * (t4 - t3):.3f} ms')
    del s2


print(f'import numba_integrators_b200: {1e3 * t_import:.0f} ms (reference v0.3.1: 268 ms)')
_startup(ni.rhs.harmonic, 'built-in RHS (includes CUDA context creation)')
_startup(_f, 'Python RHS (AST -> CUDA snippet -> NVRTC)')
_startup(_f, 'Python RHS again (module cache)')
print('  reference v0.3.1 (numba JIT): first initialisation 1700 ms, second 778 ms, first step 617 ms')
s = ni.RK45(ni.rhs.harmonic, 0.0, y0, t_bound=1, atol=1e-8, rtol=1e-8)
while ni.step(s):
    pass
s = ni.RK45(ni.rhs.harmonic, 0.0, y0, t_bound=1, atol=1e-8, rtol=1e-8)
t0 = time.perf_counter()
n = 0
while ni.step(s):
    n += 1
dt = time.perf_counter() - t0
print(f'README loop, lock-step ni.step from Python: {n} steps, {1e6 * dt / n:.1f} us/step (kernel launch + D2H per step)')
s = ni.RK45(ni.rhs.harmonic, 0.0, y0, t_bound=200.0, atol=1e-8, rtol=1e-8)
for _ in range(100):
    ni.step(s)
t0 = time.perf_counter()
n = 0
while ni.step(s):
    n += 1
dt = time.perf_counter() - t0
print(f'long lock-step loop (CUDA graph from the 65th step on): {n} steps, {1e6 * dt / n:.1f} us/step')
s = ni.RK45(ni.rhs.harmonic, 0.0, y0, t_bound=20000 * np.pi, atol=1e-8, rtol=1e-8)   # tests/profiling.py of the reference
t0 = time.perf_counter()
s.run()
dt = time.perf_counter() - t0
print(f'one trajectory, fast-forward on the device: {s.n_accepted} steps in {dt:.3f} s = {1e6 * dt / s.n_accepted:.2f} us/step '
      f'(reference, numba on one CPU core: 2.03 us/step, SURVEY.md section 6)')
for N in (1, 32, 1024, 32768, 1048576):
    e = ni.RK45(ni.rhs.harmonic, 0.0, np.tile(y0, (N, 1)), t_bound=200.0, atol=1e-8, rtol=1e-8)
    t0 = time.perf_counter()
    e.run()
    dt = time.perf_counter() - t0
    print(f'N = {N:8d}: {e.n_accepted.sum() / dt / 1e6:12.2f} M steps/s')
