"""CTA-per-trajectory kernels: auxiliary output over the whole state (_advanced.py:181, :276-278) and ni.ff2cond
(_API.py:41-49) -- against fixtures the live reference produced (oracle/gen_golden.py: heat48_aux_*), and per-trajectory
max_step / first_step on the large kernels."""
import numpy as np
import pytest

import numba_integrators_b200 as ni
from conftest import load_golden

pytestmark = pytest.mark.gpu

HEAT = 'return p[0] * ((ym - 2.0 * yc) + yp);'
HEAT_COMPONENT = ('const double ym = j > 0 ? y[j - 1] : 0.0, yp = j + 1 < n ? y[j + 1] : 0.0;\n'
                  'return p[0] * ((ym - 2.0 * y[j]) + yp);')
AUX = 'double tot = 0.0;\nfor (int j = 0; j < n; ++j) tot += y[j];\naux[0] = tot;\naux[1] = y[n / 2];'


def rhs(kind):
    if kind == 'stencil':
        return ni.CudaStencilRHS(HEAT, P=1, aux_body=AUX, Q=2)
    return ni.CudaComponentRHS(HEAT_COMPONENT, P=1, aux_body=AUX, Q=2)


@pytest.mark.parametrize('kind', ['stencil', 'component'])
def test_auxiliary_output_of_a_large_system(kind):
    meta, g = load_golden('heat48_aux_rk45')
    Solver = ni.Advanced(1, 2, ni.RK45)
    s = Solver(rhs(kind), g['t0'], g['y0'], g['params'], g['t_bound'], rtol=g['rtol'], atol=g['atol'])
    assert np.allclose(s.auxiliary, g['aux0'], rtol=1e-14)          # the FSAL seed's auxiliary output (constructor)
    # lock-step on trajectory 0: the auxiliary output after every accepted step
    one = Solver(rhs(kind), g['t0'][0], g['y0'][0], g['params'][0], g['t_bound'][0], rtol=g['rtol'], atol=g['atol'])
    k = 0
    while ni.step(one):
        assert abs(one.t - g['rec_t'][0, k]) <= 1e-9
        t, y, aux = one.state
        assert np.allclose(aux, [np.sum(y), y[len(y) // 2]], rtol=1e-13)
        k += 1
    assert k == g['n_accepted'][0]
    assert s.run() == 0
    assert np.array_equal(s.n_accepted, g['n_accepted']) and np.array_equal(s.n_rejected, g['n_rejected'])
    assert np.allclose(s.y, g['y'], rtol=1e-12, atol=1e-300) and np.allclose(s.auxiliary, g['aux'], rtol=1e-12)
    rec = s.terminal()
    assert np.array_equal(rec['aux'], s.auxiliary)


@pytest.mark.parametrize('kind', ['stencil', 'component'])
def test_ff2cond_on_a_large_system(kind):
    """Python predicate (translated), CUDA predicate and built-in threshold: all stop where the reference stops."""
    meta, g = load_golden('heat48_aux_ff2cond_rk45')
    j, thr = meta['cond_index'], meta['cond_threshold']
    Solver = ni.Advanced(1, 2, ni.RK45)

    def fresh():
        return Solver(rhs(kind), g['t0'], g['y0'], g['params'], g['t_bound'], rtol=g['rtol'], atol=g['atol'])

    def python_predicate(solver, p):
        return solver.y[24] < p
    for cond, par in ((python_predicate, thr), (ni.CudaCondition(f'return y[{j}] < cp[0];'), thr), (ni.y_below(j), thr),
                      (ni.CudaCondition('return aux[1] < cp[0] && aux[0] == aux[0];'), thr)):
        s = fresh()
        hit = ni.ff2cond(s, cond, par)
        assert np.array_equal(np.asarray(hit, bool), g['hit'].astype(bool)), cond
        assert np.array_equal(s.n_accepted, g['n_accepted']) and np.array_equal(s.n_rejected, g['n_rejected'])
        assert np.allclose(s.t, g['t'], rtol=1e-12) and np.allclose(s.y, g['y'], rtol=1e-12, atol=1e-300)
        assert np.allclose(s.auxiliary, g['aux'], rtol=1e-12)
        # the trajectories that were stopped continue to t_bound afterwards
        assert s.run() == 0 and np.all(s.t == g['t_bound'])


def test_built_in_heat_takes_predicates_too():
    meta, g = load_golden('heat48_aux_ff2cond_rk45')
    s = ni.RK45(ni.rhs.heat1d(g['params'][:, 0]), g['t0'], g['y0'], g['t_bound'], rtol=g['rtol'], atol=g['atol'])
    hit = ni.ff2cond(s, ni.y_below(meta['cond_index']), meta['cond_threshold'])
    assert np.array_equal(np.asarray(hit, bool), g['hit'].astype(bool)) and np.array_equal(s.n_accepted, g['n_accepted'])
    assert np.allclose(s.y, g['y'], rtol=1e-12, atol=1e-300)


def test_per_trajectory_step_limits_on_a_large_system():
    meta, g = load_golden('per_traj_limits_heat48_rk23')
    s = ni.Advanced(1, 0, ni.RK23)(ni.rhs.heat1d, g['t0'], g['y0'], g['params'], g['t_bound'], rtol=g['rtol'], atol=g['atol'],
                                  max_step=g['max_step_v'], first_step=g['first_step_v'])
    assert np.allclose(s.h_abs, g['h_abs0'], rtol=1e-12) and np.array_equal(s.max_step, g['max_step_v'])
    assert s.run() == 0
    assert np.array_equal(s.n_accepted, g['n_accepted']) and np.array_equal(s.n_rejected, g['n_rejected'])
    assert np.array_equal(s.nfev, g['nfev'])
    assert np.allclose(s.y, g['y'], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize('n', [200, 333])
def test_recording_on_cta_per_trajectory_kernels(n):
    """The record log of the large kernels (32 776 B per step at n = 4096): every accepted (t, y) of a fast-forward run
    equals the lock-step sequence of a single solver bit for bit; a fast-forward run reserves its record slots eight at a
    time per trajectory (the unused ones of the last chunk are tagged and skipped by the drain), lock-step calls one at a
    time; a log that is too small stops the trajectories between attempts, and draining + relaunching loses nothing."""
    from numba_integrators_b200 import workloads as wl
    h = wl.c5_heat(N=24, n=n, t_bound=1.0)
    make = lambda **kw: ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 1.0, rtol=1e-8, atol=1e-8, **kw)
    ens = make(record=20_000)
    assert ens.run() == 0
    total = int(ens.n_accepted.sum())
    assert total <= ens.record_count() <= total + 7 * ens.N    # slots in use, the tagged leftovers included
    rec = ens.record_drain().sorted()
    assert len(rec) == total
    for k in (0, 7, 23):
        one = ni.RK45(ni.rhs.heat1d(h.params[k, 0]), 0.0, h.y0[k], 1.0, rtol=1e-8, atol=1e-8)
        ts, ys = [], []
        while ni.step(one):
            ts.append(one.t)
            ys.append(one.y)
        tk, yk, _ = rec.trajectory(k)
        assert np.array_equal(tk, np.array(ts)) and np.array_equal(yk, np.array(ys))
    # a log that is too small: drain and continue
    small = make(record=300)
    got = []
    for _ in range(200):
        try:
            small.run()
            got.append(small.record_drain())
            break
        except ni.NiError as e:
            assert e.code == -6
            got.append(small.record_drain())
    assert sum(len(g) for g in got) == total
    assert np.array_equal(small.y, ens.y) and np.array_equal(small.n_accepted, ens.n_accepted)
    assert np.array_equal(small.n_rejected, ens.n_rejected)
    # lock-step recording: one slot per accepted step, nothing wasted
    lock = make(record=20_000)
    while ni.step(lock):
        pass
    assert lock.record_count() == total
    lrec = lock.record_drain().sorted()
    t7, y7, _ = lrec.trajectory(7)
    r7 = rec.trajectory(7)
    assert np.array_equal(t7, r7[0]) and np.array_equal(y7, r7[1])
