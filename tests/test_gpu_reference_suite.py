"""The reference's own functional tests (tests/unittests/test_public.py), restated line by line against this
package: same problems (reference.py:38-57, as Python lambdas that go through the source translator), same
tolerances, same assertions, SciPy as the comparison oracle exactly as the reference uses it."""
from itertools import product

import numpy as np
import pytest
import scipy.integrate

import numba_integrators_b200 as ni

pytestmark = pytest.mark.gpu


class Problem:  # reference.py:18-33
    problems = []

    def __init__(self, name, differential, solution, x0, x_end):
        self.name, self.differential, self.solution, self.x0, self.x_end = name, differential, solution, x0, x_end
        self.y0, self.y_end = solution(x0), solution(x_end)
        Problem.problems.append(self)


riccati = Problem('riccati', lambda x, y: 2 * y/x - x ** 2 * y ** 2,
                  lambda x: np.array((x**2 / (x**5 + 4) * 5,), np.float64), 1., 20.)
sine = Problem('sine', lambda x, y: np.array((y[1], -y[0])),
               lambda x: np.array((np.sin(x), np.cos(x)), np.float64), 0., 10.)
exponential = Problem('exponential', lambda x, y: y,
                      lambda x: np.array((np.exp(x),), np.float64), 0., 10.)
scipy_integrators = {ni.RK23: scipy.integrate.RK23, ni.RK45: scipy.integrate.RK45}


def scipy_solve(solver_type, function, x0, y0, x_end, rtol, atol):  # reference.py:64-77
    solver = scipy_integrators[solver_type](function, x0, y0, x_end, atol=atol, rtol=rtol)
    while solver.status == 'running':
        solver.step()
    assert solver.t == x_end
    return solver.y


def compare_to_scipy(solver_type, rtol, atol, problem):  # test_public.py:17-38
    solver = solver_type(problem.differential, problem.x0, problem.y0, problem.x_end, rtol=rtol, atol=atol)
    while ni.step(solver):
        pass
    assert solver.t == problem.x_end
    y_ni = solver.y
    err_ni = np.sum(np.abs(y_ni - problem.y_end))
    y_scipy = scipy_solve(solver_type, problem.differential, problem.x0, problem.y0, problem.x_end, rtol, atol)
    err_scipy = np.sum(np.abs(y_scipy - problem.y_end))
    assert err_ni < err_scipy*1.02


class Test_Basic:
    @pytest.mark.parametrize(('solver', 'problem'), product(ni.ALL, Problem.problems),
                             ids=lambda v: getattr(v, 'name', getattr(v, '__name__', str(v))))
    def test_to_scipy(self, solver, problem):  # :42-46
        compare_to_scipy(solver, 1e-10, 1e-10, problem)

    def test_max_step(self):  # :48-61
        max_step = 1e-4
        low = max_step * (1 - 1e-13)
        problem = exponential
        solver = ni.RK45(problem.differential, problem.x0, problem.y0, t_bound=problem.x_end, first_step=max_step,
                         max_step=max_step)
        x_prev = solver.t
        for _ in range(100):
            ni.step(solver)
            assert low < (solver.t - x_prev) <= max_step
            x_prev = solver.t


def f_advanced(t, y, p):  # test_public.py:65-67 (the inner call inlined: a translated RHS cannot call other functions)
    return y, p


class Test_Advanced:
    parameters = (1., np.array((1., 1.), dtype=np.float64))

    def test_all_class_creation(self):  # :75-77
        Solvers = ni.Advanced(None, None, ni.Solvers.ALL)
        assert isinstance(Solvers, dict)

    @pytest.mark.parametrize('solver_type', (ni.RK23, ni.RK45))
    def test_identical_to_base(self, solver_type):  # :79-125
        problem = exponential
        solver_base = solver_type(problem.differential, problem.x0, problem.y0, problem.x_end)
        Solvers = ni.Advanced(None, None, solver_type)
        assert Solvers is not None
        assert not isinstance(Solvers, dict)
        solver_advanced = Solvers(f_advanced, problem.x0, problem.y0, self.parameters, problem.x_end)
        # the auxiliary round-trips the (float, array) parameter tuple -- flattened to (1, 1, 1) here
        assert np.array_equal(solver_advanced.auxiliary, [1., 1., 1.])
        assert solver_base.t == solver_advanced.t
        assert np.all(solver_base.y == solver_advanced.y)
        ni.step(solver_base)
        ni.step(solver_advanced)
        assert solver_base.t == solver_advanced.t
        assert np.all(solver_base.y == solver_advanced.y)
        assert np.array_equal(solver_advanced.auxiliary, [1., 1., 1.])
        x_base, y_base, x_advanced, y_advanced = [], [], [], []
        while ni.step(solver_base):
            assert ni.step(solver_advanced)
            x_base.append(solver_base.t)
            y_base.append(solver_base.y)
            x_advanced.append(solver_advanced.t)
            y_advanced.append(solver_advanced.y)
        assert x_base == x_advanced
        assert all(np.all(y_b == y_a) for y_b, y_a in zip(y_base, y_advanced))


def condition_exponential(solver, parameters):  # :145-147
    return solver.y[0] > parameters


class Test_fast_forward:
    def test_t(self):  # :150-165
        problem = exponential
        solver = ni.RK45(problem.differential, problem.x0, problem.y0, problem.x_end)
        t_range = problem.x_end - problem.x0
        t_step1 = t_range / 10.1
        ni.ff2t(solver, t_step1)
        assert solver.t == t_step1
        assert solver.t_bound == problem.x_end
        while ni.ff2t(solver, solver.t + t_step1):
            ...
        assert solver.t == solver.t_bound
        assert solver.t_bound == problem.x_end

    def test_condition(self):  # :167-175
        problem = exponential
        solver = ni.RK45(problem.differential, problem.x0, problem.y0, problem.x_end)
        condition_parameter = 4.
        assert ni.ff2cond(solver, condition_exponential, condition_parameter)
        assert condition_exponential(solver, condition_parameter)
