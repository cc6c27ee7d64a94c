"""Reference API behaviour on the GPU path; each test mirrors one in the reference's
tests/unittests/test_public.py (line numbers cited)."""
import numpy as np
import pytest

import numba_integrators_b200 as ni
from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_readme_example():  # README.md:47-69
    solver = ni.RK45(ni.rhs.harmonic, 0.0, np.array((0., 1.)), t_bound=1, atol=1e-8, rtol=1e-8)
    t, y = [], []
    while ni.step(solver):
        t.append(solver.t)
        y.append(solver.y)
    meta, g = load_golden('c1_readme_sine_rk45')
    assert len(t) == 11 and t[-1] == 1.0
    assert np.allclose(t, g['rec_t'][0], rtol=1e-9) and np.allclose(np.array(y), g['rec_y'][0], rtol=1e-9)
    assert isinstance(solver.t, float) and solver.y.shape == (2,)
    assert y[0] is not y[1]                                    # solver.y is a fresh array after every step


def test_readme_advanced_example():  # README.md:74-111
    Solver = ni.Advanced(3, 1, ni.RK45)
    solver = Solver(ni.rhs.readme_advanced, 0., np.array((0., 1.)), (2., np.array((-1., 1.))), t_bound=1, atol=1e-8,
                    rtol=1e-8)
    meta, g = load_golden('readme_advanced_rk45')
    assert np.allclose(solver.auxiliary, g['aux0'][0])
    k = 0
    while ni.step(solver):
        assert abs(solver.t - g['rec_t'][0, k]) < 1e-9   # step times jitter ~1e-11 once a `pow` differs by 1 ulp
        k += 1
    assert k == g['n_accepted'][0] and np.allclose(solver.auxiliary, g['aux'][0], rtol=1e-12)
    t, y, aux = solver.state
    assert t == 1.0


@pytest.mark.parametrize('method', ['RK23', 'RK45'])
@pytest.mark.parametrize('problem', ['riccati', 'sine', 'exponential'])
def test_to_reference_problems(method, problem):  # test_public.py:17-46 (rtol = atol = 1e-10)
    meta, g = load_golden(f'problem_{problem}_{method.lower()}')
    ctor = ni.RK45 if method == 'RK45' else ni.RK23
    solver = ctor(ni.BuiltinRHS(meta['rhs']), g['t0'][0], g['y0'][0], g['t_bound'][0], rtol=1e-10, atol=1e-10)
    while ni.step(solver):
        pass
    assert solver.t == g['t_bound'][0]                         # :26
    assert solver.n_accepted == g['n_accepted'][0] and solver.n_rejected == g['n_rejected'][0]
    assert np.allclose(solver.y, g['y'][0], rtol=1e-12, atol=0)


def test_max_step():  # test_public.py:48-61
    max_step = 1e-4
    low = max_step * (1 - 1e-13)
    solver = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), t_bound=10.0, first_step=max_step, max_step=max_step)
    meta, g = load_golden('max_step_exponential')
    x_prev = solver.t
    for k in range(100):
        ni.step(solver)
        assert low < (solver.t - x_prev) <= max_step
        assert solver.t == g['rec_t'][0, k]
        x_prev = solver.t
    solver = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), t_bound=10.0, max_step=max_step)   # :127-142
    x_prev = solver.t
    for _ in range(100):
        ni.step(solver)
        assert low < (solver.t - x_prev) <= max_step
        x_prev = solver.t


@pytest.mark.parametrize('method', [ni.RK23, ni.RK45])
def test_advanced_identical_to_base(method):  # test_public.py:79-125: bitwise identical t and y at every step
    base = method(ni.rhs.exponential, 0.0, np.array((1.,)), 10.0)
    adv = ni.Advanced(0, 0, method)(ni.rhs.exponential, 0.0, np.array((1.,)), None, 10.0)
    assert base.t == adv.t and np.all(base.y == adv.y)
    while ni.step(base):
        assert ni.step(adv)
        assert base.t == adv.t and np.all(base.y == adv.y)
    assert not ni.step(adv)


def test_ff2t():  # test_public.py:150-165
    solver = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), 10.0)
    t_step1 = 10.0 / 10.1
    assert ni.ff2t(solver, t_step1) is True
    assert solver.t == t_step1
    assert solver.t_bound == 10.0
    while ni.ff2t(solver, solver.t + t_step1):
        ...
    assert solver.t == solver.t_bound == 10.0


def test_ff2cond():  # test_public.py:167-175, once with a host predicate and once with the device predicate
    def condition(solver, parameters):
        return solver.y[0] > parameters
    a = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), 10.0)
    assert ni.ff2cond(a, condition, 4.)
    assert condition(a, 4.)
    b = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), 10.0)
    assert ni.ff2cond(b, ni.y_above(0), 4.)
    assert b.t == a.t and np.all(b.y == a.y)
    c = ni.RK45(ni.rhs.exponential, 0.0, np.array((1.,)), 1.0)
    assert not ni.ff2cond(c, ni.y_above(0), 4.)               # reaches t_bound first
    assert c.t == 1.0


def test_failure_exit_commits_rejected_state():  # _basic.py:179-180, SURVEY A.2-10
    meta, g = load_golden('failure_blowup_rk45_basic')
    solver = ni.RK45(ni.rhs.blowup, 0.0, np.array((1.,)), 2.0, rtol=1e-8, atol=1e-8)
    k = 0
    while ni.step(solver):
        k += 1
    assert k == 499 and solver.status == 2
    assert abs(solver.t - g['t'][0]) < 1e-12 and solver.t != 2.0
    assert not ni.step(solver)                                 # latched


def test_ensemble_semantics_and_records():
    rng = np.random.default_rng(0)
    y0 = rng.uniform(-1, 1, size=(1000, 2))
    ens = ni.RK45(ni.rhs.harmonic, 0.0, y0, 3.0, rtol=1e-8, atol=1e-8, record=100_000)
    steps = 0
    while ni.step(ens):                                        # mask with any() truthiness
        steps += 1
    assert np.all(ens.t == 3.0) and steps == ens.n_accepted.max()
    rec = ens.record_drain().sorted()
    assert len(rec) == ens.n_accepted.sum()
    t5, y5, _ = rec.trajectory(5)
    single = ni.RK45(ni.rhs.harmonic, 0.0, y0[5], 3.0, rtol=1e-8, atol=1e-8)
    ts = []
    while ni.step(single):
        ts.append(single.t)
    assert np.array_equal(t5, np.array(ts)) and np.array_equal(y5[-1], single.y)
    # fast-forward with recording and a log that is too small: drain and continue
    ens2 = ni.RK45(ni.rhs.harmonic, 0.0, y0, 3.0, rtol=1e-8, atol=1e-8, record=5000)
    got = 0
    for _ in range(100):
        try:
            ens2.run()
            got += len(ens2.record_drain())
            break
        except ni.NiError as e:
            assert e.code == -6
            got += len(ens2.record_drain())
    assert got == ens.n_accepted.sum()
    assert np.array_equal(ens2.y, ens.y) and np.array_equal(ens2.n_accepted, ens.n_accepted)


def test_user_rhs_snippet_matches_builtin():
    body = '''
        dy[0] = __dmul_rn(p[0], __dadd_rn(y[1], -y[0]));
        dy[1] = __dadd_rn(__dmul_rn(y[0], __dadd_rn(p[1], -y[2])), -y[1]);
        dy[2] = __dadd_rn(__dmul_rn(y[0], y[1]), -__dmul_rn(p[2], y[2]));
        aux[0] = __dadd_rn(__dadd_rn(__dmul_rn(y[0], y[0]), __dmul_rn(y[1], y[1])), __dmul_rn(y[2], y[2]));
    '''
    from numba_integrators_b200 import workloads as wl
    w = wl.c3_lorenz(N=512, t_bound=1.0)
    user = ni.CudaRHS(body, n=3, P=3, Q=1, name='lorenz_user')
    a = ni.Advanced(3, 1, ni.RK45)(user, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    b = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    a.run()
    b.run()
    assert np.array_equal(a.y, b.y) and np.array_equal(a.auxiliary, b.auxiliary)
    assert np.array_equal(a.n_accepted, b.n_accepted) and np.array_equal(a.n_rejected, b.n_rejected)
    with pytest.raises(ni.NiError) as ei:
        ni.RK45(ni.CudaRHS('dy[0] = undefined_symbol;', n=1), 0.0, np.ones(1), 1.0)
    assert ei.value.code == -3


def test_device_views_and_nccl_gather_single_rank():
    """The final gather runs on the ensemble's own device buffers (zero-copy torch views) over NCCL."""
    import os
    import torch
    import torch.distributed as dist
    from numba_integrators_b200 import distributed as nd, workloads as wl
    w = wl.c3_lorenz(N=1000, t_bound=1.0)
    ens = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    ens.run()
    tens, small = nd.device_tensors(ens)
    assert small and tens['y'].shape == (3, 1000) and tens['y'].is_cuda
    assert np.array_equal(tens['y'].t().cpu().numpy(), ens.y)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    h = wl.c5_heat(N=6, n=300, t_bound=1.0)                      # trajectory-major device layout
    big = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 1.0, rtol=1e-8, atol=1e-8)
    big.run()
    dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    try:
        full = nd.gather(ens)
        full_big = nd.gather(big)
    finally:
        dist.destroy_process_group()
    assert np.array_equal(full_big['y'], big.y) and np.array_equal(full_big['n_accepted'], big.n_accepted)
    assert np.array_equal(full['y'], ens.y) and np.array_equal(full['t'], ens.t)
    assert np.array_equal(full['auxiliary'], ens.auxiliary) and np.array_equal(full['n_accepted'], ens.n_accepted)
    assert np.array_equal(full['n_rejected'], ens.n_rejected) and np.array_equal(full['status'], ens.status)


def test_terminal_records_equal_the_field_getters():
    """ni_ensemble_pack_terminal / ni_ensemble_get_terminal: one packed row per trajectory (t, y, aux, counters,
    status) for both device layouts -- the payload of the final gather, and a one-copy download."""
    from numba_integrators_b200 import workloads as wl
    w = wl.c3_lorenz(N=777, t_bound=1.0)
    ens = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    ens.run()
    h = wl.c5_heat(N=5, n=300, t_bound=1.0)
    big = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 1.0, rtol=1e-8, atol=1e-8)
    big.run()
    v = wl.c4_vanderpol(N=333, t_bound=2.0)
    vdp = ni.RK23(ni.rhs.vanderpol(v.params[:, 0]), 0.0, v.y0, 2.0, rtol=1e-6, atol=1e-6)
    vdp.run()
    for e in (ens, big, vdp):
        rec = e.terminal()
        assert rec.dtype.itemsize == 8 * (1 + e.n + e.Q) + 12 and rec.shape == (e.N,)
        assert np.array_equal(rec['t'], e.t) and np.array_equal(rec['y'], e.y)
        assert np.array_equal(rec['n_accepted'], e.n_accepted) and np.array_equal(rec['n_rejected'], e.n_rejected)
        assert np.array_equal(rec['status'], e.status)
        if e.Q:
            assert np.array_equal(rec['aux'], e.auxiliary)


def test_user_rhs_n8_coupled_oscillators_against_scipy_like_accuracy():
    """A user snippet with n = 8 (stage vectors still in registers): chain of coupled oscillators,
    energy is conserved to the tolerance; ragged ensemble size (not a multiple of the warp)."""
    body = '''
        for (int i = 0; i < 4; ++i) {
            const double left = i > 0 ? y[i - 1] : 0.0, right = i < 3 ? y[i + 1] : 0.0;
            dy[i] = y[4 + i];
            dy[4 + i] = p[0] * (left - 2.0 * y[i] + right);
        }
        double e = 0.0;
        for (int i = 0; i < 4; ++i) e += 0.5 * y[4 + i] * y[4 + i];
        for (int i = 0; i <= 4; ++i) {
            const double a = i < 4 ? y[i] : 0.0, b = i > 0 ? y[i - 1] : 0.0;
            e += 0.5 * p[0] * (a - b) * (a - b);
        }
        aux[0] = e;
    '''
    rng = np.random.default_rng(5)
    N = 77
    y0 = rng.uniform(-1, 1, size=(N, 8))
    k = rng.uniform(0.5, 2.0, size=(N, 1))
    f = ni.CudaRHS(body, n=8, P=1, Q=1, name='chain8')
    s = ni.Advanced(1, 1, ni.RK45)(f, 0.0, y0, k, 10.0, rtol=1e-10, atol=1e-10)
    e0 = s.auxiliary[:, 0].copy()
    assert s.run() == 0
    assert np.all(s.t == 10.0) and np.all(s.status == 1)
    assert np.max(np.abs(s.auxiliary[:, 0] - e0) / e0) < 1e-7
    # lock-step and fast-forward give identical bits
    s2 = ni.Advanced(1, 1, ni.RK45)(f, 0.0, y0, k, 10.0, rtol=1e-10, atol=1e-10)
    while ni.step(s2):
        pass
    assert np.array_equal(s2.y, s.y) and np.array_equal(s2.n_accepted, s.n_accepted)


def test_record_consumers(tmp_path):
    """Zero-copy device views of the record log and the npz / parquet writers."""
    import pyarrow.parquet as pq
    rng = np.random.default_rng(1)
    y0 = rng.uniform(-1, 1, size=(300, 2))
    ens = ni.RK45(ni.rhs.harmonic, 0.0, y0, 2.0, rtol=1e-6, atol=1e-6, record=20_000)
    ens.run()
    v = ens.record_views()
    k = v['count']
    assert v['y'].is_cuda and v['y'].shape == (2, 20_000) and k == ens.n_accepted.sum()
    t_dev = v['t'][:k].cpu().numpy()
    rec = ens.record_drain()
    assert np.array_equal(np.sort(t_dev), np.sort(rec.t))
    rec.to_npz(tmp_path / 'r.npz')
    rec.to_parquet(tmp_path / 'r.parquet')
    z = np.load(tmp_path / 'r.npz')
    assert np.array_equal(z['y'], rec.y)
    tab = pq.read_table(tmp_path / 'r.parquet')
    assert tab.num_rows == len(rec) and tab.column_names == ['traj', 'step', 't', 'y0', 'y1']


def test_user_stencil_snippet_matches_builtin_heat():
    """A user 3-point stencil fused into the CTA-per-trajectory kernels gives the same bits as the built-in."""
    from numba_integrators_b200 import workloads as wl
    w = wl.c5_heat(N=6, n=1000, t_bound=2.0)       # ragged: 1000 components on 512 threads x 2
    user = ni.CudaStencilRHS('return __dmul_rn(p[0], __dadd_rn(__dadd_rn(ym, -__dmul_rn(2.0, yc)), yp));', P=1)
    a = ni.RK45(user, 0.0, w.y0, 2.0, rtol=1e-8, atol=1e-8, parameters=w.params)
    b = ni.RK45(ni.rhs.heat1d(w.params[:, 0]), 0.0, w.y0, 2.0, rtol=1e-8, atol=1e-8)
    a.run()
    b.run()
    assert np.array_equal(a.y, b.y) and np.array_equal(a.n_accepted, b.n_accepted)
    # a space-dependent coefficient uses j and n
    var = ni.CudaStencilRHS('const double k = p[0] * (1.0 + (double)j / n); return k * (ym - 2.0 * yc + yp);', P=1)
    c = ni.RK23(var, 0.0, w.y0, 0.5, rtol=1e-6, atol=1e-8, parameters=w.params)
    assert c.run() == 0 and np.all(np.isfinite(c.y)) and np.all(c.t == 0.5)


def _maybe_njit(*sig):
    """The reference's users hand numba-compiled functions to the solvers; decorate like the README does when
    numba is importable (the facade only reads `.py_func`'s source, it never calls the function)."""
    try:
        import numba as nb
        return nb.njit(*sig) if sig else nb.njit
    except Exception:  # pragma: no cover
        return lambda f: f


def test_readme_example_verbatim_with_a_python_rhs():  # README.md:47-69, unchanged except the import
    try:
        import numba as nb
        deco = nb.njit(nb.float64[:](nb.float64, nb.float64[:]))
    except Exception:  # pragma: no cover
        deco = lambda f: f  # noqa: E731

    @deco
    def f(t, y):
        '''Differential equation for sine wave'''
        return np.array((y[1], -y[0]))

    y0 = np.array((0., 1.))
    solver = ni.RK45(f, 0.0, y0, t_bound=1, atol=1e-8, rtol=1e-8)
    t, y = [], []
    while ni.step(solver):
        t.append(solver.t)
        y.append(solver.y)
    meta, g = load_golden('c1_readme_sine_rk45')
    assert t == g['rec_t'][0].tolist()                         # bit-identical to the reference's own run
    assert np.array_equal(np.array(y), g['rec_y'][0])


def test_readme_advanced_example_verbatim_with_a_python_rhs():  # README.md:74-111
    @_maybe_njit()
    def f(t, y, parameters):
        '''Differential equation for sine wave'''
        auxiliary = parameters[0] * y[1]
        dy = np.array((auxiliary, -y[0])) + parameters[1]
        return dy, auxiliary

    t0 = 0.
    y0 = np.array((0., 1.))
    parameters = (2., np.array((-1., 1.)))
    Solver = ni.Advanced(None, None, ni.RK45)                  # numba type signatures are not needed here
    solver = Solver(f, t0, y0, parameters, t_bound=1, atol=1e-8, rtol=1e-8)
    t, y, auxiliary = [], [], []
    while ni.step(solver):
        t.append(solver.t)
        y.append(solver.y)
        auxiliary.append(solver.auxiliary)
    meta, g = load_golden('readme_advanced_rk45')
    k = len(t)
    assert k == g['n_accepted'][0] and isinstance(auxiliary[-1], float)
    assert t == g['rec_t'][0, :k].tolist() and np.array_equal(np.array(y), g['rec_y'][0, :k])
    assert auxiliary[-1] == g['aux'][0, 0]


def test_python_rhs_ensemble_matches_builtin_lorenz():
    def lorenz(t, y, p):
        x, yy, z = y[0], y[1], y[2]
        dy = np.array((p[0] * (yy - x), x * (p[1] - z) - yy, x * yy - p[2] * z))
        return dy, np.array(((x * x + yy * yy) + z * z,))
    from numba_integrators_b200 import workloads as wl
    w = wl.c3_lorenz(N=2000, t_bound=1.0)
    a = ni.Advanced(None, None, ni.RK45)(lorenz, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    b = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    a.run()
    b.run()
    assert np.array_equal(a.y, b.y) and np.array_equal(a.auxiliary, b.auxiliary)
    with pytest.raises(ni.TranslationError):
        ni.RK45(lambda t, y: np.linalg.solve(np.eye(2), y), 0.0, np.ones(2), 1.0)


def test_structref_flavour():  # ni.sr (_structref.py): x / x_bound field names, applies direction
    def f(t, y):
        return np.array((y[1], -y[0]))
    state = ni.sr.RK45(f, 1.0, np.array((0.3, -0.7)), 0.0, rtol=1e-8, atol=1e-8)   # backward
    meta, g = load_golden('backward_harmonic_rk45')
    k = 0
    while ni.sr.step(state):
        assert state.x == g['rec_t'][0, k] and np.array_equal(state.y, g['rec_y'][0, k])
        k += 1
    assert k == g['n_accepted'][0] == 10 and state.x == state.x_bound == 0.0


def test_ff2cond_with_user_predicates_on_an_ensemble():
    """ff2cond with (a) a Python predicate translated into the kernel, (b) a CUDA predicate snippet, (c) the
    built-in threshold: all stop every trajectory at the same accepted step as the reference's host loop."""
    rng = np.random.default_rng(2)
    y0 = rng.uniform(0.5, 1.5, size=(500, 1))

    def condition(solver, parameters):                          # test_public.py:145-147
        return solver.y[0] > parameters

    T = 1.6                                                     # exp(1.6) = 4.95: only y0 > 0.81 crosses 4
    ens = {k: ni.RK45(ni.rhs.exponential, 0.0, y0, T, rtol=1e-6, atol=1e-8) for k in 'abc'}
    hit_a = ni.ff2cond(ens['a'], condition, 4.0)
    hit_b = ni.ff2cond(ens['b'], ni.CudaCondition('return y[0] > cp[0];'), 4.0)
    hit_c = ni.ff2cond(ens['c'], ni.y_above(0), 4.0)
    assert np.array_equal(hit_a, hit_b) and np.array_equal(hit_a, hit_c) and 0 < hit_a.sum() < 500
    for k in 'bc':
        assert np.array_equal(ens['a'].t, ens[k].t) and np.array_equal(ens['a'].y, ens[k].y)
    y, t = ens['a'].y[:, 0], ens['a'].t
    assert np.all(y[hit_a] > 4.0) and np.all(t[~np.asarray(hit_a)] == T) and np.all(t[hit_a] <= T)
    # the host loop of the reference gives the same stopping point for a single trajectory
    one = ni.RK45(ni.rhs.exponential, 0.0, y0[7], T, rtol=1e-6, atol=1e-8)
    while one.step():
        if one.y[0] > 4.0:
            break
    assert one.t == t[7] and one.y[0] == y[7]
    # a second ff2cond continues from the stopping points
    hit2 = ni.ff2cond(ens['a'], condition, 4.5)
    assert 0 < hit2.sum() < hit_a.sum() and np.all(ens['a'].y[:, 0][hit2] > 4.5)


def test_dense_medium_dimension_python_rhs_against_scipy():
    """n = 40 densely coupled system through the Python-source translator (stage vectors spill from registers
    to local memory above n ~ 16 -- slower, same results): compared with SciPy at tolerance level."""
    import scipy.integrate
    n = 40

    def ring(t, y):
        dy = np.zeros(40)
        m = np.sum(y) / 40
        for j in range(40):
            dy[j] = -0.5 * (y[j] - m) + 0.1 * np.sin(y[(j + 1) % 40] - y[j]) + 0.05 * np.cos(t)
        return dy

    rng = np.random.default_rng(9)
    y0 = rng.uniform(-1, 1, size=(33, n))
    ens = ni.RK45(ring, 0.0, y0, 4.0, rtol=1e-9, atol=1e-11)
    assert ens.run() == 0 and np.all(ens.t == 4.0)
    for i in (0, 17, 32):
        ref = scipy.integrate.solve_ivp(ring, (0.0, 4.0), y0[i], method='RK45', rtol=1e-11, atol=1e-13)
        assert np.max(np.abs(ens.y[i] - ref.y[:, -1])) < 1e-7


def test_attempt_budget_relaunch_and_non_finite_inputs():
    """(a) A per-launch attempt budget (bounded kernel time) parks trajectories between attempts and relaunches:
    bit-identical to the unbounded run, also mid-retry.  (b) NaN in the state latches NI_ST_NONFINITE instead of
    spinning forever like the reference's retry loop would."""
    from numba_integrators_b200 import _lib, workloads as wl
    w = wl.c3_lorenz(N=3000, t_bound=2.0)
    make = lambda: ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 2.0, rtol=1e-8, atol=1e-8)  # noqa: E731
    a, b = make(), make()
    _lib.check(_lib.lib().ni_ensemble_set_max_attempts(b._h, 37))
    assert a.run() == 0 and b.run() == 0
    assert b.timing()['launches'] > a.timing()['launches'] + 3
    assert np.array_equal(a.y, b.y) and np.array_equal(a.n_accepted, b.n_accepted)
    assert np.array_equal(a.n_rejected, b.n_rejected) and np.array_equal(a.h_abs, b.h_abs)
    y0 = w.y0[:64].copy()
    y0[5, 1] = np.nan
    c = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, y0, w.params[:64], 1.0, rtol=1e-8, atol=1e-8)
    assert c.run() == 0
    st = c.status
    assert st[5] == 3 and np.all(np.delete(st, 5) == 1)
    assert not np.any(ni.step(c))


def test_module_caches(tmp_path, monkeypatch):
    """NVRTC specialisations are shared per process, and cubins persist on disk when NI_B200_CACHE_DIR is set."""
    import time
    from numba_integrators_b200 import _lib, _rhs
    monkeypatch.setenv('NI_B200_CACHE_DIR', str(tmp_path))
    body = 'dy[0] = -y[0] * 1.2345;  /* cache test */'
    t0 = time.perf_counter()
    a = ni.RK45(ni.CudaRHS(body, n=1), 0.0, np.ones(1), 1.0)
    t_cold = time.perf_counter() - t0
    assert len(list(tmp_path.glob('*.cubin'))) == 1
    n_modules = len(_rhs._MODULE_CACHE)
    b = ni.RK45(ni.CudaRHS(body, n=1), 0.0, np.ones((5, 1)), 1.0)       # same source: in-process hit
    assert len(_rhs._MODULE_CACHE) == n_modules and a._mod.value == b._mod.value
    _rhs._MODULE_CACHE.clear()                                          # force a reload from the disk cache
    t0 = time.perf_counter()
    c = ni.RK45(ni.CudaRHS(body, n=1), 0.0, np.ones(1), 1.0)
    t_warm = time.perf_counter() - t0
    assert b'cubin cache hit' in _lib.lib().ni_module_log(c._mod)
    assert t_warm < t_cold
    a.run(), c.run()
    assert a.y == c.y and abs(a.y[0] - np.exp(-1.2345)) < 1e-3


def test_component_rhs_dense_large_system():
    """ni.CudaComponentRHS: one CTA per trajectory, the whole stage vector staged in shared memory, arbitrary coupling.
    (a) a mean-field Kuramoto model with n = 300 oscillators against SciPy; (b) the heat equation written
    component-wise against the stencil kernel (same arithmetic per component, different reduction order)."""
    import scipy.integrate
    n = 300
    body = '''
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += sin(y[k] - y[j]);
        return p[0] + 0.002 * j + p[1] / n * s;
    '''
    rng = np.random.default_rng(12)
    y0 = rng.uniform(-np.pi, np.pi, size=(5, n))
    par = np.column_stack([rng.uniform(0.5, 1.0, 5), rng.uniform(0.5, 2.0, 5)])
    ens = ni.RK45(ni.CudaComponentRHS(body, P=2), 0.0, y0, 2.0, rtol=1e-9, atol=1e-11, parameters=par)
    assert ens.run() == 0 and np.all(ens.t == 2.0) and np.all(ens.status == 1)

    def kuramoto(t, y, w, K):
        return w + 0.002 * np.arange(n) + K / n * np.sin(y[None, :] - y[:, None]).sum(axis=1)
    for i in (0, 4):
        ref = scipy.integrate.solve_ivp(kuramoto, (0, 2.0), y0[i], args=tuple(par[i]), rtol=1e-11, atol=1e-13)
        assert np.max(np.abs(ens.y[i] - ref.y[:, -1])) < 1e-7
    from numba_integrators_b200 import workloads as wl
    w = wl.c5_heat(N=4, n=777, t_bound=3.0)
    comp = ni.CudaComponentRHS('const double ym = j > 0 ? y[j - 1] : 0.0, yp = j + 1 < n ? y[j + 1] : 0.0;'
                               'return __dmul_rn(p[0], __dadd_rn(__dadd_rn(ym, -__dmul_rn(2.0, y[j])), yp));', P=1)
    a = ni.RK45(comp, 0.0, w.y0, 3.0, rtol=1e-8, atol=1e-8, parameters=w.params)
    b = ni.RK45(ni.rhs.heat1d(w.params[:, 0]), 0.0, w.y0, 3.0, rtol=1e-8, atol=1e-8)
    a.run(), b.run()
    assert np.array_equal(a.n_accepted, b.n_accepted) and np.array_equal(a.n_rejected, b.n_rejected)
    assert np.max(np.abs(a.y - b.y)) <= 1e-12 * np.max(np.abs(b.y))


def test_branch_free_tail_arithmetic_matches_the_operators():
    """csrc/ni_math.cuh: the guarded division / square root of the error-norm tail must equal CUDA's `/` and sqrt()
    bit for bit over the guarded operand range, and the controller power seeded from the mean square must equal
    the one seeded from the root (both correctly rounded but for ~1e-5 of arguments)."""
    import ctypes
    from numba_integrators_b200 import _lib
    L = _lib.lib()
    n = 400_000_000
    bad = ctypes.c_int64(-1)
    for which in (0, 1):
        _lib.check(L.ni_devmath_check(0, which, n, 1234 + which, ctypes.byref(bad)))
        assert bad.value == 0, (which, bad.value)
    for which in (2, 3):
        _lib.check(L.ni_devmath_check(0, which, n, 99 + which, ctypes.byref(bad)))
        assert 0 <= bad.value <= n * 5e-5, (which, bad.value)
    # the lean body's guard: scales up to 2^1020, any error
    _lib.check(L.ni_devmath_check(0, 5, n, 4321, ctypes.byref(bad)))
    assert bad.value == 0, bad.value
    # the reciprocal behind the quotient: |1 - b r| <= 2^-52 (so a quotient can only be misrounded when the exact value
    # lies within ~2^-104 relative of a rounding boundary: ~2^-50 of all operand pairs)
    _lib.check(L.ni_devmath_check(0, 4, n, 77, ctypes.byref(bad)))
    assert 0 < bad.value <= 2 ** 12, bad.value


def test_long_lock_step_loop_uses_the_graph_path_and_matches_fast_forward():
    """ni.step() switches to a captured CUDA graph after 64 calls on an ensemble (csrc/ni_b200.cu step_with_graph);
    the graph must be re-captured when the launch arguments change (recording switched on mid-loop) and the result
    must stay bit-identical to the fast-forward run."""
    y0 = np.array([[0.0, 1.0], [0.3, -0.2], [1.0, 1.0]])
    ref = ni.RK45(ni.rhs.harmonic, 0.0, y0, 30.0, rtol=1e-8, atol=1e-8)
    ref.run()
    s = ni.RK45(ni.rhs.harmonic, 0.0, y0, 30.0, rtol=1e-8, atol=1e-8)
    k = 0
    while ni.step(s):
        k += 1
        if k == 150:
            ni.ff2t(s, 20.0)                    # t_bound lives in device memory: same arguments, same graph
    assert k > 200
    s2 = ni.RK45(ni.rhs.harmonic, 0.0, y0, 30.0, rtol=1e-8, atol=1e-8)
    for _ in range(100):
        ni.step(s2)
    s2.record_enable(4096)                       # other kernel, other arguments: re-capture
    while ni.step(s2):
        pass
    rec = s2.record_drain()
    assert len(rec) == int(np.sum(s2.n_accepted)) - 300
    assert np.array_equal(s2.y, ref.y) and np.array_equal(s2.t, ref.t)
    assert np.array_equal(s2.n_accepted, ref.n_accepted)
    # the ff2t leg changes the step sequence, so only the end point is comparable
    assert np.all(s.t == 30.0) and np.allclose(s.y, ref.y, rtol=1e-7, atol=1e-9)


def test_state_fields_are_assignable_like_jitclass_attributes():
    """_basic.py:182-199: the solver's fields are plain jitclass attributes.  Restarting a trajectory from a stored
    (t, y, h_abs, K[-1]) -- the complete state between two steps -- continues it bit for bit, on both layouts."""
    from numba_integrators_b200 import workloads as wl
    w = wl.c3_lorenz(N=500, t_bound=2.0)
    Solver = ni.Advanced(3, 1, ni.RK45)
    a = Solver(ni.rhs.lorenz63, 0.0, w.y0, w.params, 2.0, rtol=1e-8, atol=1e-8)
    for _ in range(40):
        ni.step(a)
    snap = (a.t, a.y, a.h_abs, a.K, a.step_size)
    assert a.run() == 0
    b = Solver(ni.rhs.lorenz63, 0.0, w.y0 * 0.5 + 1.0, w.params, 2.0, rtol=1e-8, atol=1e-8)   # a different history
    b.t, b.y, b.h_abs, b.K, b.step_size = snap
    assert np.array_equal(b.t, snap[0]) and np.array_equal(b.y, snap[1]) and np.array_equal(b.K[:, -1], snap[3][:, -1])
    assert b.run() == 0
    assert np.array_equal(b.y, a.y) and np.array_equal(b.t, a.t) and np.array_equal(b.h_abs, a.h_abs)
    assert np.array_equal(b.auxiliary, a.auxiliary)
    # single solver, scalar attribute; assigning t past t_bound finishes it
    s = ni.RK45(ni.rhs.harmonic, 0.0, np.array((0., 1.)), t_bound=1, atol=1e-8, rtol=1e-8)
    s.y = np.array((0.5, 0.5))
    assert np.array_equal(s.y, [0.5, 0.5])
    s.t = 1.0
    assert not ni.step(s) and s.t == 1.0
    with pytest.raises(ni.NiError):
        s._set(7, 1.0)                      # n_accepted and the other bookkeeping fields stay read-only
    # CTA-per-trajectory layout
    h = wl.c5_heat(N=3, n=200, t_bound=1.0)
    c = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 1.0, rtol=1e-8, atol=1e-8)
    for _ in range(5):
        ni.step(c)
    snap = (c.t, c.y, c.h_abs, c.K)
    c.run()
    d = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0 * 0.3, 1.0, rtol=1e-8, atol=1e-8)
    d.t, d.y, d.h_abs, d.K = snap
    d.run()
    assert np.array_equal(d.y, c.y) and np.array_equal(d.n_accepted + 0, d.n_accepted)


def test_terminal_sink_written_by_the_run_kernel():
    """ni_ensemble_set_terminal_sink: the run kernel's retire path stores the packed terminal record of every trajectory
    that leaves for good into every sink buffer (the gather fused into the integration: distributed.FusedGather).  The
    rows must equal the records ni_ensemble_get_terminal packs afterwards -- lean body, general body (finite max_step),
    lock-step calls, a trajectory that is already finished when the sink is set, and a failing trajectory."""
    import torch
    from numba_integrators_b200 import workloads as wl
    w = wl.c3_lorenz(N=3001, t_bound=1.0)
    Solver = ni.Advanced(3, 1, ni.RK45)

    def check(ens, first_row=0, drive=lambda e: e.run()):
        dt = ens.terminal_sink_dtype()
        assert dt.itemsize % 16 == 0 and dt.itemsize >= ens.terminal_dtype().itemsize and dt.names == ens.terminal_dtype().names
        bufs = [torch.zeros((first_row + ens.N + 3) * dt.itemsize, dtype=torch.uint8, device='cuda') for _ in range(2)]
        ens.set_terminal_sink([b.data_ptr() for b in bufs], first_row)
        drive(ens)
        ens.sync()
        ref = ens.terminal()
        for b in bufs:
            rows = b.cpu().numpy().view(dt)
            got = rows[first_row:first_row + ens.N]
            for name in dt.names:
                assert np.array_equal(got[name], ref[name], equal_nan=got[name].dtype.kind == 'f'), name
            assert not rows[:first_row].view(np.uint8).any() and not rows[first_row + ens.N:].view(np.uint8).any()
        ens.set_terminal_sink([])
        return ref

    a = check(Solver(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8), first_row=5)       # lean body
    assert np.all(a['status'] == 1) and a['n_accepted'].min() > 20
    b = check(Solver(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8, max_step=0.05))       # general body
    assert np.all(b['n_accepted'] >= 20)

    def lock_step(e):
        while ni.step(e):
            pass
    c = check(Solver(ni.rhs.lorenz63, 0.0, w.y0[:200], w.params[:200], 0.2, rtol=1e-8, atol=1e-8), drive=lock_step)
    assert np.all(c['t'] == 0.2)
    done = Solver(ni.rhs.lorenz63, 0.0, w.y0[:100], w.params[:100], 0.5, rtol=1e-8, atol=1e-8)
    done.run()
    check(done)                                   # finished before the sink was set: delivered by the next launch
    blow = ni.RK45(ni.rhs.blowup, 0.0, np.linspace(1.0, 2.0, 64)[:, None], 10.0, rtol=1e-8, atol=1e-8)
    f = check(blow)
    assert np.all(f['status'] != 1)               # y' = y^2 leaves every trajectory failed / non-finite before t = 10
    # CTA-per-trajectory systems have no in-kernel sink
    h = wl.c5_heat(N=2, n=100, t_bound=0.1)
    big = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 0.1, rtol=1e-8, atol=1e-8)
    with pytest.raises(ni.NiError):
        big.set_terminal_sink([torch.zeros(1024, dtype=torch.uint8, device='cuda').data_ptr()])


def test_stiffness_probe():
    """ni_ensemble_stiffness: spectral-radius estimate of df/dy at the current state, for every kind of
    right-hand side; h_abs * rho against the method's stability interval separates the stability-limited heat equation
    from accuracy-limited problems."""
    from numba_integrators_b200 import workloads as wl
    # harmonic oscillator: J = [[0, 1], [-1, 0]] has |J v| = |v| for every v
    rng = np.random.default_rng(5)
    ho = ni.RK45(ni.rhs.harmonic, 0.0, rng.uniform(-1, 1, size=(100, 2)), 10.0, rtol=1e-8, atol=1e-8)
    ho.run()
    s = ho.stiffness()
    assert np.allclose(s['rho'], 1.0, rtol=1e-6) and not s['stiff'].any()
    # a Python right-hand side (translated), linear with eigenvalues -1 and -1000
    def f(t, y):
        return np.array((-y[0], -1000.0 * y[1]))
    lin = ni.RK45(f, 0.0, np.array((1.0, 1.0)), 5.0, rtol=1e-6, atol=1e-9)
    for _ in range(300):          # (mid-run: the clip of the last step to t_bound would leave a small h_abs behind)
        assert ni.step(lin)
    r = lin.stiffness()
    assert abs(r['rho'] - 1000.0) < 1.0 and r['stiff'] and 1.5 < r['h_rho'] < 12.0   # the proposed step hovers around the stability limit (3.3)
    # Lorenz at tol 1e-8: accuracy-limited
    w = wl.c3_lorenz(N=500, t_bound=2.0)
    lo = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 2.0, rtol=1e-8, atol=1e-8)
    lo.run()
    sl = lo.stiffness()
    assert np.all(sl['rho'] > 1.0) and np.all(sl['rho'] < 200.0) and not sl['stiff'].any()
    before = (lo.y.copy(), lo.h_abs.copy(), lo.n_accepted.copy())
    lo.stiffness(3)
    assert np.array_equal(lo.y, before[0]) and np.array_equal(lo.h_abs, before[1]) and np.array_equal(lo.n_accepted, before[2])
    # CTA-per-trajectory: 1-D heat equation, rho -> 4 k cos^2(pi / (2 (n + 1))), stability-limited steps
    for n in (200, 1000):
        h = wl.c5_heat(N=6, n=n, t_bound=50.0)
        he = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 50.0, rtol=1e-8, atol=1e-8)
        for _ in range(150):
            assert ni.step(he)
        sh = he.stiffness(40)
        exact = 4.0 * h.params[:, 0] * np.cos(np.pi / (2 * (n + 1))) ** 2
        assert np.all(sh['rho'] <= exact * 1.0001) and np.all(sh['rho'] >= 0.9 * exact), (sh['rho'], exact)
        # (an instantaneous indicator: right after a rejection the proposed step is below the boundary)
        assert sh['stiff'].mean() >= 0.5 and np.all(sh['h_rho'] > 0.5) and np.all(sh['h_rho'] < 12.0), sh['h_rho']


def test_launch_order_changes_nothing_but_the_schedule():
    """ni_ensemble_set_order: the work queue hands the trajectories out in a caller-given order (longest first against
    the tail of a run); every result stays bit-identical -- both layouts, recording included."""
    from numba_integrators_b200 import workloads as wl
    w = wl.c3_lorenz(N=5003, t_bound=1.0)
    make = lambda: ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 1.0, rtol=1e-8, atol=1e-8)
    ref = make()
    ref.run()
    rng = np.random.default_rng(9)
    a = make()
    a.set_order(rng.permutation(w.N))
    a.run()
    b = make()
    order = b.order_by_cost(ref.n_accepted + ref.n_rejected)
    assert np.array_equal(np.sort(order), np.arange(w.N)) and (ref.n_accepted + ref.n_rejected)[order[0]] == (ref.n_accepted + ref.n_rejected).max()
    b.run()
    assert np.array_equal(b.y, ref.y)
    b.set_order(None)
    assert np.array_equal(b.get_order(), np.arange(w.N))
    assert b.order_by_cost() is None   # default cost: the attempts of the ensemble's own previous run, sorted on the device
    own = b.get_order()
    cost = (ref.n_accepted + ref.n_rejected)[own]
    assert np.array_equal(np.sort(own), np.arange(w.N)) and np.all(np.diff(cost) <= 0)   # (2048 buckets of one attempt here)
    b.reset()
    b.run()
    for e in (a, b):
        assert np.array_equal(e.y, ref.y) and np.array_equal(e.t, ref.t) and np.array_equal(e.auxiliary, ref.auxiliary)
        assert np.array_equal(e.n_accepted, ref.n_accepted) and np.array_equal(e.n_rejected, ref.n_rejected)
    with pytest.raises(ni.NiError):
        a.set_order(np.zeros(w.N, dtype=np.uint32))        # not a permutation
    with pytest.raises(ValueError):
        a.set_order(np.arange(10))
    # CTA-per-trajectory layout, with recording
    h = wl.c5_heat(N=9, n=300, t_bound=1.0)
    mk = lambda: ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 1.0, rtol=1e-8, atol=1e-8, record=4000)
    c, d = mk(), mk()
    c.run()
    d.set_order(np.arange(9)[::-1].copy())
    d.run()
    assert np.array_equal(c.y, d.y) and np.array_equal(c.n_accepted, d.n_accepted)
    rc, rd = c.record_drain().sorted(), d.record_drain().sorted()
    assert np.array_equal(rc.t, rd.t) and np.array_equal(rc.y, rd.y) and np.array_equal(rc.traj, rd.traj)
