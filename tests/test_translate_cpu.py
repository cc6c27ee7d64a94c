"""Python right-hand side -> CUDA snippet translation (host logic; the Python function is never executed).
The emitted C is checked by compiling it with g++ and comparing against the Python function itself."""
import ctypes
import math
import subprocess

import numpy as np
import pytest

import numba_integrators_b200 as ni

K = 1.5


def f_sine(t, y):
    '''Differential equation for sine wave (README.md:54-57)'''
    return np.array((y[1], -y[0]))


def f_readme_advanced(t, y, parameters):  # README.md:80-85
    auxiliary = parameters[0] * y[1]
    dy = np.array((auxiliary, -y[0])) + parameters[1]
    return dy, auxiliary


def f_lorenz(t, y, p):
    x, yy, z = y[0], y[1], y[2]
    dy = np.array((p[0] * (yy - x), x * (p[1] - z) - yy, x * yy - p[2] * z))
    return dy, np.array(((x * x + yy * yy) + z * z,))


def f_heat(t, y, p):
    n = y.size
    dy = np.empty(n)
    for j in range(n):
        ym = y[j - 1] if j > 0 else 0.0
        yp = y[j + 1] if j + 1 < n else 0.0
        dy[j] = p[0] * K * ((ym - 2.0 * y[j]) + yp)
    return dy, (np.sum(y), math.sin(t))


def f_branchy(t, y):
    if y[0] > 0.5:
        a = y[0] ** 2 * y[0]   # (x**3 is libm pow in CPython but multiplications in numba and in the translation)
    else:
        a = -np.sqrt(abs(y[0])) + t
    b = max(a, y[1]) if t < 2 else 0.0
    return np.array([a + b, math.exp(-y[1]) * math.pi])   # math.exp == libm exp (np.exp has its own SIMD kernel)


riccati = lambda x, y: 2 * y / x - x ** 2 * y ** 2  # noqa: E731  (reference.py:39)


def _compile(body, n, P, Q, tmp_path, tag):
    src = tmp_path / f'{tag}.cpp'
    src.write_text('#include <cmath>\nstatic double ni_sign(double x){return (x>0)-(x<0);}\n'
                   'extern "C" void rhs(double t, const double* y, const double* p, double* dy, double* aux) {\n'
                   + body + '\n}\n')
    so = tmp_path / f'{tag}.so'
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', str(src), '-o', str(so)], check=True)
    fn = ctypes.CDLL(str(so)).rhs
    dp = ctypes.POINTER(ctypes.c_double)
    fn.argtypes = [ctypes.c_double, dp, dp, dp, dp]

    def call(t, y, p):
        y, p = np.ascontiguousarray(y, np.float64), np.ascontiguousarray(p, np.float64)
        dy, aux = np.zeros(n), np.zeros(max(Q, 1))
        fn(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), dy.ctypes.data_as(dp), aux.ctypes.data_as(dp))
        return dy, aux[:Q]
    return call


@pytest.mark.parametrize('fun,n,params,adv', [
    (f_sine, 2, None, False), (f_readme_advanced, 2, (2., np.array((-1., 1.))), True),
    (f_lorenz, 3, np.array((10., 28., 8 / 3)), True), (f_heat, 6, np.array((0.7,)), True),
    (f_branchy, 2, None, False), (riccati, 1, None, False)])
def test_translation_matches_python_bit_for_bit(fun, n, params, adv, tmp_path):
    body, P, Q, aux_scalar = ni.translate(fun, n, params, adv)
    call = _compile(body, n, P, Q, tmp_path, getattr(fun, '__name__', 'lam').replace('<', '').replace('>', ''))
    rng = np.random.default_rng(3)
    flat = np.zeros(0) if params is None else np.concatenate([np.ravel(np.asarray(q, float)) for q in
                                                              (params if isinstance(params, tuple) else (params,))])
    assert P == flat.size
    for _ in range(20):
        t, y = rng.uniform(0.1, 3), rng.uniform(-2, 2, n)
        dy, aux = call(t, y, flat)
        ref = fun(t, y, params) if params is not None else fun(t, y)
        ref_dy, ref_aux = (ref if adv else (ref, ()))
        assert np.array_equal(dy, np.atleast_1d(ref_dy))
        ref_aux = np.concatenate([np.ravel(np.asarray(q, float)) for q in
                                  (ref_aux if isinstance(ref_aux, tuple) else (ref_aux,))]) if adv else np.zeros(0)
        assert np.array_equal(aux, ref_aux)
    assert aux_scalar == (fun is f_readme_advanced)


def test_unsupported_constructs_are_named():
    def calls_other(t, y):
        return f_sine(t, y)

    def wrong_len(t, y):
        return np.array((y[0],))
    with pytest.raises(ni.TranslationError, match='unsupported call'):
        ni.translate(calls_other, 2)
    with pytest.raises(ni.TranslationError, match='returns 1 derivatives for 2'):
        ni.translate(wrong_len, 2)
    with pytest.raises(ni.TranslationError, match='parameters'):
        ni.translate(f_lorenz, 3, None, True)


def f_runtime_float_rebound(t, y):
    a = 2.0
    if t > 1.0:
        a = 3.0
        k = 2  # a translate-time constant first bound under the condition: usable inside the branch only
        a = a + k
    return a * y


def test_constants_and_returns_under_a_runtime_condition(tmp_path):
    """an int-valued local rebound inside a run-time `if` was folded unconditionally, and a `return`
    inside one was hoisted out of it.  Both are refused by name; the float form is a real variable."""
    def int_rebound(t, y):
        a = 2
        if t > 1.0:
            a = 3
        return a * y

    def tuple_rebound(t, y):
        c = (1.0, 2.0)
        if y[0] > 0.0:
            c = (3.0, 4.0)
        return np.array((c[0] * y[0], c[1] * y[1]))

    def conditional_return(t, y):
        if t > 1.0:
            return 2.0 * y
        return y

    def scoped_constant_escapes(t, y):
        if t > 1.0:
            k = 2
        return k * y
    with pytest.raises(ni.TranslationError, match='reassigned inside a run-time `if`'):
        ni.translate(int_rebound, 2)
    with pytest.raises(ni.TranslationError, match='`return` inside a run-time `if`'):
        ni.translate(conditional_return, 2)
    with pytest.raises(ni.TranslationError):
        ni.translate(scoped_constant_escapes, 2)
    for k, fun in enumerate((f_runtime_float_rebound, tuple_rebound)):  # variables of the generated code: fine
        body, P, Q, _ = ni.translate(fun, 2)
        call = _compile(body, 2, P, Q, tmp_path, f'rt_{k}')
        for t in (0.5, 1.5):
            for y in (np.array((0.25, -3.0)), np.array((-0.25, 3.0))):
                assert np.array_equal(call(t, y, np.zeros(0))[0], fun(t, y))


def condition_exponential(solver, parameters):  # tests/unittests/test_public.py:145-147 of the reference
    return solver.y[0] > parameters


def condition_combo(solver, p):
    r2 = solver.y[0] ** 2 + solver.y[1] ** 2
    return (r2 < p[0] and solver.t > p[1]) or solver.auxiliary[0] >= 3.0


def test_predicate_translation():
    body, cp = ni.translate_condition(condition_exponential, 1, parameters=4.0)
    assert 'return (y[0] > cp[0]);' in body and cp.tolist() == [4.0]
    body, cp = ni.translate_condition(condition_combo, 2, Q=1, parameters=np.array((0.5, 1.0)))
    assert '&&' in body and '||' in body and 'aux[0]' in body and cp.tolist() == [0.5, 1.0]
    with pytest.raises(ni.TranslationError, match='solver.h_abs'):
        ni.translate_condition(lambda solver, p: solver.h_abs > p, 1, parameters=1.0)
