"""Generate tests/golden/*.npz by running the UNMODIFIED reference (numba) in this container.

Run (build container only; /root/reference does not exist on the GPU box):

    PYTHONPATH=oracle/_refstub:/root/reference/src NUMBA_CACHE_DIR=$PWD/.numba_cache \
        OPENBLAS_NUM_THREADS=1 python oracle/gen_golden.py

`oracle/_refstub/limedev` is a 2-line stand-in for the absent dev-tool dependency the
reference imports at _API.py:9 (it is never on the integration path).

Every case stores its inputs and the reference's outputs; rejected-step counts come from an
RHS-call counter carried in the Advanced `parameters` tuple:
nfev = 2 + attempts*S'  =>  rejected = (nfev-2)/S' - accepted.
The RHS functions below use exactly the expression order of oracle/ni_oracle.c:rhs_eval and
numba_integrators_b200/csrc/ni_rhs_builtin.cuh.
"""
import json
import sys
from pathlib import Path

import numba as nb
import numpy as np

import numba_integrators as ni
from numba_integrators import reference as ref

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from numba_integrators_b200 import workloads as wl  # noqa: E402

OUT = ROOT / 'tests' / 'golden'
OUT.mkdir(parents=True, exist_ok=True)

P_T = nb.types.Tuple((nb.float64[:], nb.int64[:]))
AUX_T = nb.float64[:]


# ---- Advanced-form RHS (p[0] = parameters, p[1][0] = call counter) -------------------------
@nb.njit
def f_harmonic(t, y, p):
    p[1][0] += 1
    return np.array((y[1], -y[0])), np.zeros(0)


@nb.njit
def f_lorenz(t, y, p):
    p[1][0] += 1
    q = p[0]
    x, yy, z = y[0], y[1], y[2]
    dy = np.array((q[0] * (yy - x), x * (q[1] - z) - yy, x * yy - q[2] * z))
    return dy, np.array(((x * x + yy * yy) + z * z,))


@nb.njit
def f_vdp(t, y, p):
    p[1][0] += 1
    mu = p[0][0]
    return np.array((y[1], (mu * (1.0 - y[0] * y[0])) * y[1] - y[0])), np.zeros(0)


@nb.njit
def f_heat(t, y, p):
    p[1][0] += 1
    k = p[0][0]
    n = y.size
    dy = np.empty(n)
    for j in range(n):
        ym = y[j - 1] if j > 0 else 0.0
        yp = y[j + 1] if j + 1 < n else 0.0
        dy[j] = k * ((ym - 2.0 * y[j]) + yp)
    return dy, np.zeros(0)


@nb.njit
def f_burgers(t, y, p):
    p[1][0] += 1
    c0, c1 = p[0][0], p[0][1]
    n = y.size
    dy = np.empty(n)
    for j in range(n):
        ym = y[j - 1] if j > 0 else 0.0
        yp = y[j + 1] if j + 1 < n else 0.0
        dy[j] = c0 * ((ym - 2.0 * y[j]) + yp) - (y[j] * c1) * (yp - ym)
    return dy, np.zeros(0)


@nb.njit
def f_exponential(t, y, p):
    p[1][0] += 1
    return y.copy(), np.zeros(0)


@nb.njit
def f_riccati(t, y, p):
    p[1][0] += 1
    return 2 * y / t - t ** 2 * y ** 2, np.zeros(0)


@nb.njit
def f_blowup(t, y, p):
    p[1][0] += 1
    return y * y, np.zeros(0)


@nb.njit
def f_readme_adv(t, y, p):
    p[1][0] += 1
    q = p[0]
    a = q[0] * y[1]
    return np.array((a + q[1], -y[0] + q[2])), np.array((a,))


@nb.njit
def f_heat_aux(t, y, p):
    """heat1d with an auxiliary output over the whole state: (sum of y in index order, the middle component)"""
    p[1][0] += 1
    k = p[0][0]
    n = y.size
    dy = np.empty(n)
    tot = 0.0
    for j in range(n):
        ym = y[j - 1] if j > 0 else 0.0
        yp = y[j + 1] if j + 1 < n else 0.0
        dy[j] = k * ((ym - 2.0 * y[j]) + yp)
        tot += y[j]
    return dy, np.array((tot, y[n // 2]))


ADV_RHS = {'harmonic': f_harmonic, 'lorenz63': f_lorenz, 'vanderpol': f_vdp, 'heat1d': f_heat,
           'heat1d_aux': f_heat_aux, 'burgers1d': f_burgers, 'exponential': f_exponential, 'riccati': f_riccati, 'blowup': f_blowup,
           'readme_advanced': f_readme_adv}

# ---- basic-form RHS (no parameters) for the basic-vs-Advanced identity cases ---------------
BASIC_SIG = nb.float64[:](nb.float64, nb.float64[:])


@nb.njit(BASIC_SIG)
def b_harmonic(t, y):
    return np.array((y[1], -y[0]))


@nb.njit(BASIC_SIG)
def b_exponential(t, y):
    return y.copy()


@nb.njit(BASIC_SIG)
def b_riccati(t, y):
    return 2 * y / t - t ** 2 * y ** 2


@nb.njit(BASIC_SIG)
def b_blowup(t, y):
    return y * y


BASIC_RHS = {'harmonic': b_harmonic, 'exponential': b_exponential, 'riccati': b_riccati, 'blowup': b_blowup}

_solvers = ni.Advanced(P_T, AUX_T, ni.Solvers.ALL)


def _direct(wrapper):
    return next(c.cell_contents for c in wrapper.__closure__ if hasattr(c.cell_contents, 'py_func'))


DIRECT = {'RK23': _direct(_solvers[ni.Solvers.RK23]), 'RK45': _direct(_solvers[ni.Solvers.RK45])}
S_PRIME = {'RK23': 3, 'RK45': 6}


def _make_runner(direct):
    @nb.njit
    def run(fun, t0, y0, p, t_bound, max_step, rtol, atol, first_step, rec_cap, max_steps):
        cnt = np.zeros(1, np.int64)
        s = direct(fun, t0, y0.copy(), (p.copy(), cnt), t_bound, max_step, rtol, atol, first_step)
        n = y0.size
        rec_t = np.full(rec_cap, np.nan)
        rec_y = np.full((rec_cap, n), np.nan)
        rec_h = np.full(rec_cap, np.nan)
        h0 = s.h_abs
        aux0 = s.auxiliary.copy()
        k = 0
        while k < max_steps and s.step():
            if k < rec_cap:
                rec_t[k] = s.t
                rec_y[k] = s.y
                rec_h[k] = s.h_abs
            k += 1
        return s.t, s.y, s.auxiliary, s.h_abs, s.step_size, k, cnt[0], h0, aux0, rec_t, rec_y, rec_h
    return run


RUN = {m: _make_runner(d) for m, d in DIRECT.items()}


ONLY = None  # names (prefixes) given on the command line: generate just those cases


def _wanted(name):
    return ONLY is None or any(name.startswith(p) for p in ONLY)


def run_case(name, rhs, method, t0, y0, params, t_bound, rtol, atol, max_step=np.inf, first_step=0.0,
             rec_cap=0, rec_traj=0, max_steps=10 ** 9, note=''):
    """Run every row of y0 through the reference's Advanced solver and save inputs + outputs.
    max_step / first_step may be arrays: one value per trajectory (the reference takes them per solver object)."""
    if not _wanted(name):
        return
    y0 = np.atleast_2d(np.asarray(y0, np.float64))
    N, n = y0.shape
    per_traj_limits = np.ndim(max_step) > 0 or np.ndim(first_step) > 0
    ms_v = np.broadcast_to(np.asarray(max_step, np.float64), (N,)).copy()
    fs_v = np.broadcast_to(np.asarray(first_step, np.float64), (N,)).copy()
    params = np.zeros((N, 0)) if params is None else np.asarray(params, np.float64).reshape(N, -1)
    t0a = np.broadcast_to(np.asarray(t0, np.float64), (N,)).copy()
    tba = np.broadcast_to(np.asarray(t_bound, np.float64), (N,)).copy()
    rt = np.broadcast_to(np.asarray(rtol, np.float64), (n,)).copy()
    at = np.broadcast_to(np.asarray(atol, np.float64), (n,)).copy()
    fun, run = ADV_RHS[rhs], RUN[method]
    o = dict(t=np.empty(N), y=np.empty((N, n)), h_abs=np.empty(N), step_size=np.empty(N), h_abs0=np.empty(N),
             n_accepted=np.zeros(N, np.int64), n_rejected=np.zeros(N, np.int64), nfev=np.zeros(N, np.int64))
    aux, aux0 = [], []
    R = min(rec_traj, N)
    rec_t, rec_y, rec_h = np.full((R, rec_cap), np.nan), np.full((R, rec_cap, n), np.nan), np.full((R, rec_cap), np.nan)
    for i in range(N):
        cap = rec_cap if i < R else 0
        t, y, a, h_abs, step_size, k, nfev, h0, a0, rt_, ry_, rh_ = run(
            fun, t0a[i], y0[i], params[i], tba[i], float(ms_v[i]), rt, at, float(fs_v[i]), cap, max_steps)
        o['t'][i], o['y'][i], o['h_abs'][i], o['step_size'][i], o['h_abs0'][i] = t, y, h_abs, step_size, h0
        o['n_accepted'][i], o['nfev'][i] = k, nfev
        n_init = 2 if fs_v[i] == 0.0 else 1
        att = (nfev - n_init) / S_PRIME[method]
        assert att == int(att), (nfev, n_init)
        o['n_rejected'][i] = int(att) - k
        aux.append(a)
        aux0.append(a0)
        if i < R:
            rec_t[i], rec_y[i], rec_h[i] = rt_, ry_, rh_
    meta = dict(name=name, rhs=rhs, method=method, advanced=True, max_step=float(ms_v[0]),
                first_step=float(fs_v[0]), rec_cap=rec_cap, rec_traj=R, max_steps=min(max_steps, 2 ** 62),
                note=note, versions=dict(numba=nb.__version__, numpy=np.__version__))
    if per_traj_limits:  # (meta carries the first trajectory's values; the arrays are authoritative)
        o.update(max_step_v=ms_v, first_step_v=fs_v)
    np.savez_compressed(OUT / f'{name}.npz', meta=json.dumps(meta), t0=t0a, y0=y0, params=params, t_bound=tba,
                        rtol=rt, atol=at, aux=np.array(aux), aux0=np.array(aux0), rec_t=rec_t, rec_y=rec_y,
                        rec_h=rec_h, **o)
    print(f'{name:34s} N={N:4d} n={n:5d} acc={o["n_accepted"].mean():9.1f} rej={o["n_rejected"].mean():8.1f}')


def run_basic_case(name, rhs, method, t0, y0, t_bound, rtol, atol, max_step=np.inf, first_step=0.0, n_steps=10 ** 9,
                   note=''):
    """One trajectory through the reference's BASIC solver, stepping from Python; records every step."""
    if not _wanted(name):
        return
    ctor = ni.RK45 if method == 'RK45' else ni.RK23
    s = ctor(BASIC_RHS[rhs], t0, np.asarray(y0, np.float64), t_bound, max_step=max_step, rtol=rtol, atol=atol,
             first_step=first_step)
    n = s.y.size
    h0 = s.h_abs
    ts, ys, hs, ss = [], [], [], []
    k = 0
    while k < n_steps and ni.step(s):
        ts.append(s.t)
        ys.append(s.y.copy())
        hs.append(s.h_abs)
        ss.append(s.step_size)
        k += 1
    meta = dict(name=name, rhs=rhs, method=method, advanced=False, max_step=float(max_step),
                first_step=float(first_step), n_steps=min(n_steps, 2 ** 62), note=note)
    np.savez_compressed(OUT / f'{name}.npz', meta=json.dumps(meta), t0=np.array([t0], np.float64),
                        y0=np.asarray(y0, np.float64).reshape(1, n), params=np.zeros((1, 0)),
                        t_bound=np.array([t_bound], np.float64), rtol=np.broadcast_to(np.float64(rtol), (n,)).copy(),
                        atol=np.broadcast_to(np.float64(atol), (n,)).copy(), t=np.array([s.t]), y=s.y.reshape(1, n),
                        h_abs=np.array([s.h_abs]), step_size=np.array([s.step_size]), h_abs0=np.array([h0]),
                        n_accepted=np.array([k]), rec_t=np.array(ts).reshape(1, -1),
                        rec_y=np.array(ys).reshape(1, len(ts), n), rec_h=np.array(hs).reshape(1, -1),
                        rec_s=np.array(ss).reshape(1, -1))
    print(f'{name:34s} basic, {k} steps, t={s.t!r}')


def _make_cond_runner(direct):
    @nb.njit
    def run(fun, t0, y0, p, t_bound, rtol, atol, j, thr):
        """ni.ff2cond (_API.py:41-49) with the predicate y[j] < thr: step until it holds after an accepted step"""
        cnt = np.zeros(1, np.int64)
        s = direct(fun, t0, y0.copy(), (p.copy(), cnt), t_bound, np.inf, rtol, atol, 0.0)
        k = 0
        hit = False
        while s.step():
            k += 1
            if s.y[j] < thr:
                hit = True
                break
        return s.t, s.y, s.auxiliary, s.h_abs, k, cnt[0], hit
    return run


COND_RUN = {m: _make_cond_runner(d) for m, d in DIRECT.items()}


def run_cond_case(name, rhs, method, t0, y0, params, t_bound, rtol, atol, j, thr, note=''):
    """ff2cond(solver, lambda s, thr: s.y[j] < thr, thr) on every row through the reference's Advanced solver."""
    if not _wanted(name):
        return
    y0 = np.atleast_2d(np.asarray(y0, np.float64))
    N, n = y0.shape
    params = np.asarray(params, np.float64).reshape(N, -1)
    rt = np.broadcast_to(np.asarray(rtol, np.float64), (n,)).copy()
    at = np.broadcast_to(np.asarray(atol, np.float64), (n,)).copy()
    o = dict(t=np.empty(N), y=np.empty((N, n)), h_abs=np.empty(N), n_accepted=np.zeros(N, np.int64),
             n_rejected=np.zeros(N, np.int64), hit=np.zeros(N, np.uint8))
    aux = []
    for i in range(N):
        t, y, a, h_abs, k, nfev, hit = COND_RUN[method](ADV_RHS[rhs], float(t0), y0[i], params[i], float(t_bound), rt, at,
                                                        int(j), float(thr))
        o['t'][i], o['y'][i], o['h_abs'][i], o['n_accepted'][i], o['hit'][i] = t, y, h_abs, k, hit
        o['n_rejected'][i] = (nfev - 2) // S_PRIME[method] - k
        aux.append(a)
    meta = dict(name=name, rhs=rhs, method=method, advanced=True, max_step=float('inf'), first_step=0.0,
                cond_index=int(j), cond_threshold=float(thr), note=note)
    np.savez_compressed(OUT / f'{name}.npz', meta=json.dumps(meta), t0=np.full(N, float(t0)), y0=y0, params=params,
                        t_bound=np.full(N, float(t_bound)), rtol=rt, atol=at, aux=np.array(aux),
                        rec_t=np.zeros((0, 0)), **o)
    print(f'{name:34s} N={N:4d} n={n:5d} acc={o["n_accepted"].mean():9.1f} hit={o["hit"].mean():.2f}')


def main():
    # -- C1: README example (README.md:47-69), basic API, every step recorded
    run_basic_case('c1_readme_sine_rk45', 'harmonic', 'RK45', 0.0, [0.0, 1.0], 1.0, 1e-8, 1e-8)
    # -- README Advanced example (README.md:74-111): p = (2, (-1, 1)), aux = p0*y1
    run_case('readme_advanced_rk45', 'readme_advanced', 'RK45', 0.0, [[0.0, 1.0]], [[2.0, -1.0, 1.0]], 1.0, 1e-8,
             1e-8, rec_cap=64, rec_traj=1)
    # -- reference.py problems at the tolerance of test_public.py:46 (1e-10), both methods, both APIs
    for prob in ref.Problem.problems:
        rhs = {'sine': 'harmonic'}.get(prob.name, prob.name)
        for method in ('RK23', 'RK45'):
            cap = 4096 if method == 'RK45' else 0
            run_case(f'problem_{prob.name}_{method.lower()}', rhs, method, prob.x0, [prob.y0], None, prob.x_end,
                     1e-10, 1e-10, rec_cap=cap, rec_traj=1, note='test_public.py:17-46 (tolerances 1e-10)')
        run_basic_case(f'problem_{prob.name}_rk45_basic', rhs, 'RK45', prob.x0, prob.y0, prob.x_end, 1e-6, 1e-6)
    # -- test_max_step (test_public.py:48-61): first_step = max_step = 1e-4, 100 steps
    run_basic_case('max_step_exponential', 'exponential', 'RK45', 0.0, [1.0], 10.0, 1e-3, 1e-6, max_step=1e-4,
                   first_step=1e-4, n_steps=100)
    # -- Advanced test_max_step variant (:127-142): select_initial_step + max_step clip
    run_basic_case('max_step_exponential_auto', 'exponential', 'RK45', 0.0, [1.0], 10.0, 1e-3, 1e-6, max_step=1e-4,
                   n_steps=100)
    # -- default tolerances (test_identical_to_base, :79-125)
    for method in ('RK23', 'RK45'):
        run_basic_case(f'default_tol_exponential_{method.lower()}', 'exponential', method, 0.0, [1.0], 10.0, 1e-3,
                       1e-6)
        run_case(f'default_tol_exponential_{method.lower()}_adv', 'exponential', method, 0.0, [[1.0]], None, 10.0,
                 1e-3, 1e-6, rec_cap=256, rec_traj=1)
    # -- failure exit: y' = y^2, y(0)=1, t_bound=2
    run_case('failure_blowup_rk45', 'blowup', 'RK45', 0.0, [[1.0]], None, 2.0, 1e-8, 1e-8, rec_cap=600, rec_traj=1)
    run_basic_case('failure_blowup_rk45_basic', 'blowup', 'RK45', 0.0, [1.0], 2.0, 1e-8, 1e-8)
    # -- backward integration through Advanced (basic ignores direction)
    run_case('backward_harmonic_rk45', 'harmonic', 'RK45', 1.0, [[0.3, -0.7]], None, 0.0, 1e-8, 1e-8, rec_cap=64,
             rec_traj=1)
    run_case('backward_lorenz_rk23', 'lorenz63', 'RK23', 0.5, [[1.0, 2.0, 20.0]], [[10.0, 28.0, 8.0 / 3.0]], 0.0,
             1e-6, 1e-9, rec_cap=512, rec_traj=1)
    # -- vector tolerances, first_step given, finite max_step
    run_case('vector_tol_lorenz_rk45', 'lorenz63', 'RK45', 0.0, [[1.0, 1.0, 20.0], [-3.0, 2.0, 25.0]],
             [[10.0, 28.0, 8.0 / 3.0]] * 2, 2.0, [1e-6, 1e-8, 1e-7], [1e-9, 1e-7, 1e-8], max_step=0.05,
             first_step=1e-3, rec_cap=256, rec_traj=2)
    # -- per-trajectory t0 / t_bound
    run_case('per_traj_time_vdp_rk23', 'vanderpol', 'RK23', [0.0, 1.0, 2.0], [[1.0, 0.0], [0.5, 0.5], [-1.0, 2.0]],
             [[0.5], [2.0], [5.0]], [3.0, 4.0, 2.5], 1e-6, 1e-6, rec_cap=128, rec_traj=3)
    # -- 1-trajectory lorenz, t = 5 (405 accepted / 10 rejected)
    run_case('lorenz_single_t5', 'lorenz63', 'RK45', 0.0, [[1.0, 1.0, 1.0]], [[10.0, 28.0, 8.0 / 3.0]], 5.0, 1e-8,
             1e-8, rec_cap=512, rec_traj=1)

    # -- ensembles: prefixes of the BASELINE workloads
    w = wl.c2_harmonic(N=192)
    run_case('c2_harmonic_t100', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol, rec_cap=1024,
             rec_traj=2)
    w = wl.c3_lorenz(N=256, t_bound=1.0)
    run_case('c3_lorenz_t1', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol, rec_cap=160,
             rec_traj=4)
    w = wl.c3_lorenz(N=192, t_bound=5.0)
    run_case('c3_lorenz_t5', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol)
    w = wl.c4_vanderpol(N=96, tol=1e-8)
    run_case('c4_vanderpol_tol1e-8', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol)
    w = wl.c4_vanderpol(N=96, tol=1e-6)
    run_case('c4_vanderpol_tol1e-6', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol)
    w = wl.c5_heat(N=3, t_bound=5.0)
    run_case('c5_heat4096_t5', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol)
    w = wl.c5_heat(N=8, n=256, t_bound=12.0)
    run_case('c5_heat256_t12', w.rhs, w.method, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol)
    w = wl.c5_heat(N=4, n=64, t_bound=0.02)
    pb = np.column_stack([np.full(4, 0.5), np.full(4, 65.0 / 2.0)])  # (nu/dx^2, 1/(2dx)), dx = 1/65
    run_case('burgers64_rk45', 'burgers1d', 'RK45', 0.0, w.y0, pb, 1.0, 1e-6, 1e-8)
    w = wl.c5_heat(N=4, n=7, t_bound=3.0)
    run_case('heat7_rk23', 'heat1d', 'RK23', 0.0, w.y0, w.params, 3.0, 1e-6, 1e-8, rec_cap=64, rec_traj=1)
    # -- per-trajectory max_step / first_step (the reference takes both per solver object, _basic.py:292-299):
    #    unlimited / binding / very tight max_step, automatic / given first_step, mixed within one ensemble
    w = wl.c3_lorenz(N=8, t_bound=1.0)
    run_case('per_traj_limits_lorenz_rk45', w.rhs, 'RK45', 0.0, w.y0, w.params, 1.0, 1e-8, 1e-8,
             max_step=[np.inf, 0.02, 0.005, np.inf, 0.011, 1.0, 0.0025, np.inf],
             first_step=[0.0, 0.0, 1e-3, 1e-4, 0.0, 5e-3, 0.0, 2e-2], rec_cap=128, rec_traj=8)
    # -- CTA-per-trajectory systems with an auxiliary output over the whole state, and ff2cond on them
    w = wl.c5_heat(N=4, n=48, t_bound=0.3)
    run_case('heat48_aux_rk45', 'heat1d_aux', 'RK45', 0.0, w.y0, w.params, 0.3, 1e-8, 1e-8, rec_cap=32, rec_traj=1)
    run_cond_case('heat48_aux_ff2cond_rk45', 'heat1d_aux', 'RK45', 0.0, w.y0, w.params, 5.0, 1e-8, 1e-8, j=24, thr=0.05,
                  note='ni.ff2cond(solver, lambda s, thr: s.y[24] < thr, 0.05)')
    w = wl.c5_heat(N=4, n=48, t_bound=0.3)
    run_case('per_traj_limits_heat48_rk23', 'heat1d', 'RK23', 0.0, w.y0, w.params, 0.3, 1e-6, 1e-8,
             max_step=[np.inf, 2e-3, 5e-4, np.inf], first_step=[0.0, 1e-4, 0.0, 1e-5])


if __name__ == '__main__':
    if len(sys.argv) > 1:
        ONLY = sys.argv[1:]
    main()
