"""semantics='scipy' (NI_MODE_SCIPY): the kernels follow scipy.integrate.RK23/RK45._step_impl instead of the
reference's quirks (stale FSAL stage on retries, growth after a rejection, 8 ulp / max_step - eps).
Checked step by step against SciPy itself running the same Python right-hand side on the CPU."""
import numpy as np
import pytest
import scipy.integrate

import numba_integrators_b200 as ni

pytestmark = pytest.mark.gpu


def lorenz(t, y):
    return np.array((10.0 * (y[1] - y[0]), y[0] * (28.0 - y[2]) - y[1], y[0] * y[1] - 8.0 / 3.0 * y[2]))


def vdp(t, y):
    return np.array((y[1], 5.0 * (1.0 - y[0] * y[0]) * y[1] - y[0]))


def riccati(x, y):
    return 2 * y / x - x ** 2 * y ** 2


def sine(t, y):
    return np.array((y[1], -y[0]))


CASES = [(sine, 0.0, [0.0, 1.0], 10.0), (riccati, 1.0, [5 / 5.0], 20.0), (lorenz, 0.0, [1.0, 1.0, 1.0], 3.0),
         (vdp, 0.0, [2.0, 0.0], 12.0), (sine, 3.0, [0.3, -0.2], -2.0)]   # the last one integrates backward


@pytest.mark.parametrize('method', ['RK45', 'RK23'])
@pytest.mark.parametrize('tol', [1e-5, 1e-9])
@pytest.mark.parametrize('case', CASES, ids=lambda c: f'{c[0].__name__}{"_back" if c[3] < c[1] else ""}')
def test_step_sequence_matches_scipy(method, tol, case):
    f, t0, y0, tb = case
    ref = getattr(scipy.integrate, method)(f, t0, np.array(y0), tb, rtol=tol, atol=tol * 1e-2)
    ours = getattr(ni, method)(f, t0, np.array(y0), tb, rtol=tol, atol=tol * 1e-2, semantics='scipy')
    assert abs(ours.h_abs - ref.h_abs) <= 1e-13 * ref.h_abs                  # select_initial_step
    k = 0
    while ref.status == 'running':
        ref.step()
        assert ni.step(ours), (k, ref.t)
        k += 1
        assert abs(ours.t - ref.t) <= 1e-9 * max(1.0, abs(ref.t)), (k, ours.t, ref.t)
        assert np.all(np.abs(ours.y - ref.y) <= 1e-7 * np.maximum(1.0, np.abs(ref.y))), (k, ours.y, ref.y)
    assert ref.status == 'finished' and ours.t == tb and not ni.step(ours)
    assert ours.n_accepted == k
    assert np.all(np.abs(ours.y - ref.y) <= 1e-9 * np.maximum(1.0, np.abs(ref.y)))


def test_scipy_semantics_differ_from_reference_semantics_where_expected():
    """Lorenz, t in [0, 5], tol 1e-8: the reference's stale-FSAL cascade gives 405 accepted / 10 rejected steps,
    SciPy's rules 400 / 2."""
    args = (lorenz, 0.0, np.array((1.0, 1.0, 1.0)), 5.0)
    a = ni.RK45(*args, rtol=1e-8, atol=1e-8)
    b = ni.RK45(*args, rtol=1e-8, atol=1e-8, semantics='scipy')
    a.run()
    b.run()
    assert (a.n_accepted, a.n_rejected) == (405, 10)
    assert (b.n_accepted, b.n_rejected) == (400, 2)


def test_scipy_semantics_too_small_step_commits_nothing():
    s = ni.RK45(lambda t, y: y * y, 0.0, np.array((1.0,)), 2.0, rtol=1e-8, atol=1e-8, semantics='scipy')
    ref = scipy.integrate.RK45(lambda t, y: y * y, 0.0, np.array((1.0,)), 2.0, rtol=1e-8, atol=1e-8)
    k = 0
    while ref.status == 'running':
        ref.step()
        k += ref.status != 'failed'
    assert ref.status == 'failed'
    while ni.step(s):
        pass
    assert s.status == 2 and s.n_accepted == k
    assert abs(s.t - ref.t) <= 1e-9 and s.t < 2.0


def test_scipy_semantics_and_fast_arithmetic_on_the_stencil_kernels():
    """The CTA-per-trajectory kernels get their opt-in variants through NVRTC: heat equation, n = 200."""
    n, k = 200, 0.8
    rng = np.random.default_rng(4)
    x = (np.arange(n) + 1.0) / (n + 1.0)
    y0 = np.sin(np.pi * x) + 0.05 * rng.standard_normal(n)

    def heat(t, y):
        ym = np.concatenate(([0.0], y[:-1]))
        yp = np.concatenate((y[1:], [0.0]))
        return k * ((ym - 2.0 * y) + yp)

    ref = scipy.integrate.RK45(heat, 0.0, y0, 6.0, rtol=1e-7, atol=1e-9)
    ours = ni.RK45(ni.rhs.heat1d(k), 0.0, y0, 6.0, rtol=1e-7, atol=1e-9, semantics='scipy')
    steps = 0
    while ref.status == 'running':
        ref.step()
        assert ni.step(ours)
        steps += 1
        assert abs(ours.t - ref.t) <= 1e-9 * max(1.0, ref.t)
    assert ours.t == 6.0 and ours.n_accepted == steps and not ni.step(ours)
    assert np.max(np.abs(ours.y - ref.y)) < 1e-10
    strict = ni.RK45(ni.rhs.heat1d(k), 0.0, y0, 3.0, rtol=1e-8, atol=1e-8)
    fast = ni.RK45(ni.rhs.heat1d(k), 0.0, y0, 3.0, rtol=1e-8, atol=1e-8, arithmetic='fast')
    strict.run(), fast.run()
    assert fast.n_accepted == strict.n_accepted and fast.n_rejected == strict.n_rejected
    assert np.max(np.abs(fast.y - strict.y)) <= 1e-12 * np.max(np.abs(strict.y))
