"""ctypes front-end of the CPU oracle (oracle/ni_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product
package `numba_integrators_b200`.

Parity pin: tests/test_oracle_golden.py checks this oracle against fixtures
produced by the live reference (oracle/gen_golden.py -> tests/golden/).

Mirrors the reference call surface so parity tests read like the reference's
own: OracleSolver ~ jitclass RK (_basic.py:182-271) / _RK_Advanced
(_advanced.py:197-278); step/ff2t/ff2cond ~ _API.py:22-49.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / 'libni_oracle.so'

RK23, RK45 = 0, 1
RHS_IDS = {'harmonic': 0, 'sine': 0, 'lorenz63': 1, 'vanderpol': 2, 'heat1d': 3, 'burgers1d': 4,
           'exponential': 5, 'riccati': 6, 'readme_advanced': 7, 'blowup': 8, 'heat1d_aux': 9}
# (P, Q) of each built-in
RHS_PQ = {0: (0, 0), 1: (3, 1), 2: (1, 0), 3: (1, 0), 4: (2, 0), 5: (0, 0), 6: (0, 0), 7: (3, 1), 8: (0, 0), 9: (1, 2)}
ST_RUNNING, ST_FINISHED, ST_FAILED, ST_NONFINITE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> Path:
    """Compile libni_oracle.so with gcc (recipe: oracle/Makefile)."""
    src = _HERE / 'ni_oracle.c'
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(['make', '-C', str(_HERE), '-B', 'libni_oracle.so'], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.ni_oracle_create.restype = C.c_void_p
        L.ni_oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp,
                                       C.c_double, C.c_double, _dp, _dp, C.c_double]
        L.ni_oracle_destroy.argtypes = [C.c_void_p]
        L.ni_oracle_step.argtypes = [C.c_void_p]
        L.ni_oracle_step.restype = C.c_int
        for name in ('t', 't_bound', 'h_abs', 'step_size', 'direction'):
            f = getattr(L, f'ni_oracle_get_{name}')
            f.argtypes = [C.c_void_p]
            f.restype = C.c_double
        for name in ('n_acc', 'n_rej', 'nfev'):
            f = getattr(L, f'ni_oracle_get_{name}')
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int64
        L.ni_oracle_get_status.argtypes = [C.c_void_p]
        L.ni_oracle_get_status.restype = C.c_int
        L.ni_oracle_set_t_bound.argtypes = [C.c_void_p, C.c_double]
        L.ni_oracle_set_h_abs.argtypes = [C.c_void_p, C.c_double]
        for name in ('y', 'aux', 'K'):
            getattr(L, f'ni_oracle_get_{name}').argtypes = [C.c_void_p, _dp]
        L.ni_oracle_tableau.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.ni_oracle_max_threads.restype = C.c_int
        L.ni_oracle_run_ensemble.restype = C.c_int
        L.ni_oracle_run_ensemble.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, C.c_int, C.c_double,
            _dp, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _i64p, _i64p, _i32p, C.c_int64, _dp, _dp, _dp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _tol(v, n):
    v = np.asarray(v, dtype=np.float64)
    return np.full(n, v, dtype=np.float64) if v.ndim == 0 else np.ascontiguousarray(v, dtype=np.float64)


def _rhs_id(rhs):
    return RHS_IDS[rhs] if isinstance(rhs, str) else int(rhs)


def tableau(method):
    S = 6 if method == RK45 else 3
    A = np.zeros((S, S - 1 if method == RK45 else S))
    B, Cc, E = np.zeros(S), np.zeros(S), np.zeros(S + 1)
    lib().ni_oracle_tableau(method, _p(A), _p(B), _p(Cc), _p(E))
    return A, B, Cc, E


class OracleSolver:
    """One trajectory; attribute names follow the reference's jitclass (`_basic.py:182-199`)."""

    def __init__(self, rhs, method, t0, y0, t_bound, params=None, max_step=np.inf, rtol=1e-3, atol=1e-6,
                 first_step=0., advanced=False):
        self._L = lib()
        y0 = np.atleast_1d(np.asarray(y0, dtype=np.float64)).copy()
        self.n = len(y0)
        self.rhs_id = _rhs_id(rhs)
        self.P, self.Q = RHS_PQ[self.rhs_id]
        self.method = method
        self.n_stages = 6 if method == RK45 else 3
        self.rtol, self.atol = _tol(rtol, self.n), _tol(atol, self.n)
        par = None
        if self.P:
            par = np.ascontiguousarray(params, dtype=np.float64).reshape(self.P)
        self._h = self._L.ni_oracle_create(self.rhs_id, method, int(advanced), self.n, self.P, self.Q, float(t0),
                                           _p(y0), _p(par), float(t_bound), float(max_step), _p(self.rtol),
                                           _p(self.atol), float(first_step))
        if not self._h:
            raise ValueError('ni_oracle_create failed')

    def __del__(self):
        if getattr(self, '_h', None):
            self._L.ni_oracle_destroy(self._h)
            self._h = None

    def step(self) -> bool:
        return bool(self._L.ni_oracle_step(self._h))

    t = property(lambda s: s._L.ni_oracle_get_t(s._h))
    h_abs = property(lambda s: s._L.ni_oracle_get_h_abs(s._h), lambda s, v: s._L.ni_oracle_set_h_abs(s._h, float(v)))
    step_size = property(lambda s: s._L.ni_oracle_get_step_size(s._h))
    direction = property(lambda s: s._L.ni_oracle_get_direction(s._h))
    status = property(lambda s: s._L.ni_oracle_get_status(s._h))
    n_accepted = property(lambda s: s._L.ni_oracle_get_n_acc(s._h))
    n_rejected = property(lambda s: s._L.ni_oracle_get_n_rej(s._h))
    nfev = property(lambda s: s._L.ni_oracle_get_nfev(s._h))
    t_bound = property(lambda s: s._L.ni_oracle_get_t_bound(s._h),
                       lambda s, v: s._L.ni_oracle_set_t_bound(s._h, float(v)))

    @property
    def y(self):
        out = np.empty(self.n)
        self._L.ni_oracle_get_y(self._h, _p(out))
        return out

    @property
    def auxiliary(self):
        out = np.empty(self.Q)
        self._L.ni_oracle_get_aux(self._h, _p(out))
        return out

    @property
    def K(self):
        out = np.empty((self.n_stages + 1, self.n))
        self._L.ni_oracle_get_K(self._h, _p(out))
        return out

    @property
    def state(self):
        return (self.t, self.y, self.auxiliary) if self.Q else (self.t, self.y)


def step(solver) -> bool:  # _API.py:22-24
    return solver.step()


def ff2t(solver, t_end) -> bool:  # _API.py:27-39
    t_bound = solver.t_bound
    is_not_last = t_bound > t_end
    if is_not_last:
        solver.t_bound = t_end
    while solver.step():
        ...
    solver.t_bound = t_bound
    return is_not_last


def ff2cond(solver, condition, parameters) -> bool:  # _API.py:41-49
    while solver.step():
        if condition(solver, parameters):
            return True
    return False


def run_ensemble(rhs, method, t0, y0, t_bound, params=None, max_step=np.inf, rtol=1e-3, atol=1e-6, first_step=0.,
                 advanced=False, nthreads=0, record=0):
    """`while step(): pass` for each row of y0 (N x n) on `nthreads` host threads (0 = all cores).

    Returns a dict with t, y, aux, h_abs, n_accepted, n_rejected, status (+ rec_t/rec_y/rec_aux
    of shape (N, record, ...) when record > 0).
    """
    L = lib()
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    N, n = y0.shape
    rid = _rhs_id(rhs)
    P, Q = RHS_PQ[rid]
    par = None
    if P:
        par = np.ascontiguousarray(np.broadcast_to(np.asarray(params, dtype=np.float64).reshape(-1, P), (N, P)))
    t0a, tba = np.atleast_1d(np.asarray(t0, np.float64)), np.atleast_1d(np.asarray(t_bound, np.float64))
    per = int(t0a.size > 1 or tba.size > 1)
    if per:
        t0a = np.ascontiguousarray(np.broadcast_to(t0a, (N,)))
        tba = np.ascontiguousarray(np.broadcast_to(tba, (N,)))
    rt, at = _tol(rtol, n), _tol(atol, n)
    out = dict(t=np.empty(N), y=np.empty((N, n)), aux=np.empty((N, Q)), h_abs=np.empty(N),
               n_accepted=np.zeros(N, np.int64), n_rejected=np.zeros(N, np.int64), status=np.zeros(N, np.int32))
    rec_t = rec_y = rec_aux = None
    if record:
        rec_t = np.full((N, record), np.nan)
        rec_y = np.full((N, record, n), np.nan)
        rec_aux = np.full((N, record, Q), np.nan) if Q else None
        out.update(rec_t=rec_t, rec_y=rec_y, rec_aux=rec_aux)
    err = L.ni_oracle_run_ensemble(rid, method, int(advanced), n, P, Q, N, _p(t0a), _p(y0), _p(par), _p(tba), per,
                                   float(max_step), _p(rt), _p(at), float(first_step), int(nthreads), _p(out['t']),
                                   _p(out['y']), _p(out['aux']) if Q else None, _p(out['h_abs']),
                                   out['n_accepted'].ctypes.data_as(_i64p), out['n_rejected'].ctypes.data_as(_i64p),
                                   out['status'].ctypes.data_as(_i32p), int(record), _p(rec_t), _p(rec_y),
                                   _p(rec_aux))
    if err:
        raise RuntimeError('oracle ensemble failed')
    return out


def max_threads() -> int:
    return int(lib().ni_oracle_max_threads()) if os.cpu_count() else 1
