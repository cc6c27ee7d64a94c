"""What the reference's functional tests pin (tests/unittests/test_public.py), checked on the GPU path with plain
Python right-hand sides going through the source translator and SciPy as the comparison solver, as upstream does:

  * accuracy against the analytic solution no worse than 1.02 x SciPy's at rtol = atol = 1e-10, and landing exactly
    on t_bound                                                                    (test_public.py:17-46)
  * every step obeys max_step*(1 - 1e-13) < dt <= max_step                         (:48-61, :127-142)
  * Advanced and basic solvers give identical t, y at every step                   (:79-125)
  * ff2t lands on t_end and restores t_bound; ff2cond stops at the first hit        (:150-175)
"""
import numpy as np
import pytest
import scipy.integrate

import numba_integrators_b200 as ni

pytestmark = pytest.mark.gpu

# the three initial-value problems of reference.py:38-57: (rhs, exact solution, x0, x_end)
IVPS = {
    'riccati': (lambda x, y: 2 * y/x - x ** 2 * y ** 2, lambda x: np.array([x**2 / (x**5 + 4) * 5]), 1.0, 20.0),
    'sine': (lambda x, y: np.array((y[1], -y[0])), lambda x: np.array([np.sin(x), np.cos(x)]), 0.0, 10.0),
    'exponential': (lambda x, y: y, lambda x: np.array([np.exp(x)]), 0.0, 10.0),
}
SCIPY = {'RK23': scipy.integrate.RK23, 'RK45': scipy.integrate.RK45}


def exp_solver(ctor=None, **kw):
    rhs, exact, x0, x_end = IVPS['exponential']
    return (ctor or ni.RK45)(rhs, x0, exact(x0), x_end, **kw), x_end


@pytest.mark.parametrize('method', ['RK23', 'RK45'])
@pytest.mark.parametrize('name', list(IVPS))
def test_accuracy_relative_to_scipy(method, name):
    rhs, exact, x0, x_end = IVPS[name]
    tol = 1e-10
    ours = getattr(ni, method)(rhs, x0, exact(x0), x_end, rtol=tol, atol=tol)
    while ni.step(ours):
        pass
    assert ours.t == x_end
    theirs = SCIPY[method](rhs, x0, exact(x0), x_end, rtol=tol, atol=tol)
    while theirs.status == 'running':
        theirs.step()
    err_ours, err_scipy = np.abs(ours.y - exact(x_end)).sum(), np.abs(theirs.y - exact(x_end)).sum()
    assert err_ours < 1.02 * err_scipy, (err_ours, err_scipy)


@pytest.mark.parametrize('first_step', [1e-4, 0.0])
def test_max_step_is_strict(first_step):
    max_step = 1e-4
    solver, _ = exp_solver(max_step=max_step, first_step=first_step)
    previous = solver.t
    for _ in range(100):
        ni.step(solver)
        assert max_step * (1 - 1e-13) < solver.t - previous <= max_step
        previous = solver.t


def test_advanced_factory_returns_dict_for_all():
    assert isinstance(ni.Advanced(None, None, ni.Solvers.ALL), dict)


@pytest.mark.parametrize('ctor', [ni.RK23, ni.RK45])
def test_advanced_is_bitwise_identical_to_basic(ctor):
    def rhs_with_parameters(t, y, p):      # upstream wraps the basic RHS and hands the parameters back as auxiliary
        return y, p
    parameters = (1.0, np.array((1.0, 1.0)))
    basic, x_end = exp_solver(ctor)
    advanced = ni.Advanced(None, None, ctor)(rhs_with_parameters, 0.0, basic.y, parameters, x_end)
    assert not isinstance(ni.Advanced(None, None, ctor), dict)
    assert np.array_equal(advanced.auxiliary, [1.0, 1.0, 1.0])          # the (float, array) tuple, flattened
    assert (basic.t, basic.y.tolist()) == (advanced.t, advanced.y.tolist())
    n = 0
    while ni.step(basic):
        assert ni.step(advanced)
        assert basic.t == advanced.t and np.array_equal(basic.y, advanced.y)
        n += 1
    assert n > 5 and not ni.step(advanced)
    assert np.array_equal(advanced.auxiliary, [1.0, 1.0, 1.0])


def test_fast_forward_to_time():
    solver, x_end = exp_solver()
    leg = x_end / 10.1
    assert ni.ff2t(solver, leg)
    assert solver.t == leg and solver.t_bound == x_end
    while ni.ff2t(solver, solver.t + leg):
        pass
    assert solver.t == solver.t_bound == x_end


def test_fast_forward_to_condition():
    def y0_above(solver, threshold):
        return solver.y[0] > threshold
    solver, _ = exp_solver()
    assert ni.ff2cond(solver, y0_above, 4.0)
    assert y0_above(solver, 4.0)
