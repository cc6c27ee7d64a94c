"""Config C1 (README sine wave, ONE trajectory): latency of the GPU path, for the record.
A single trajectory cannot fill a GPU; this is the reference's home turf (2 us/step in numba)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numba_integrators_b200 as ni  # noqa: E402

y0 = np.array((0., 1.))
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
