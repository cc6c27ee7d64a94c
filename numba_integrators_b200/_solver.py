"""Solver objects and factories: the reference's L2/L3 layers on top of the C ABI.

Mirrors (file:line under /root/reference/src/numba_integrators/):
  RK jitclass + attributes          _basic.py:182-271      -> Ensemble
  RK23 / RK45 factories, defaults   _basic.py:273-333      -> RK23 / RK45
  ALL, Solvers                      _basic.py:335-339      -> ALL, Solvers
  Advanced factory                  _advanced.py:280-368   -> Advanced
  convert / _into_1d_typearray      _aux.py:117-140        -> convert
  step / ff2t / ff2cond             _API.py:22-49          -> step / ff2t / ff2cond
`y0` of shape (n,) gives a single solver whose attributes are scalars / (n,) arrays exactly as
in the reference; `y0` of shape (N, n) gives an N-trajectory ensemble (attributes gain a
leading N axis, `step()` returns a mask whose truth value is `any()`).
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from collections.abc import Iterable

import numpy as np

from . import _lib
from ._lib import check
from ._rhs import RHS, PythonRHS, as_rhs

SAFETY = 0.9       # _aux.py:16
MIN_FACTOR = 0.2   # _aux.py:18
MAX_FACTOR = 10    # _aux.py:19

# Tableaux (_aux.py:77-115); the kernels carry the same constants (csrc/ni_core.cuh NiTab<>)
RK23_A = np.array(((0, 0, 0), (1 / 2, 0, 0), (0, 3 / 4, 0)), dtype=np.float64)
RK23_B = np.array((2 / 9, 1 / 3, 4 / 9), dtype=np.float64)
RK23_C = np.array((0, 1 / 2, 3 / 4), dtype=np.float64)
RK23_E = np.array((5 / 72, -1 / 12, -1 / 9, 1 / 8), dtype=np.float64)
RK45_A = np.array(((0., 0., 0., 0., 0.),
                   (1 / 5, 0., 0., 0., 0.),
                   (3 / 40, 9 / 40, 0., 0., 0.),
                   (44 / 45, -56 / 15, 32 / 9, 0., 0.),
                   (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0),
                   (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656)), dtype=np.float64)
RK45_B = np.array((35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84), dtype=np.float64)
RK45_C = np.array((0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1), dtype=np.float64)
RK45_E = np.array((-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40), dtype=np.float64)
_TABLEAU = {_lib.RK23: (2, 3, RK23_A, RK23_B, RK23_C, RK23_E), _lib.RK45: (4, 6, RK45_A, RK45_B, RK45_C, RK45_E)}


def _into_1d_typearray(item, length=1, dtype=np.float64, max_ndim=1):
    """_aux.py:117-132, extended: y0 may also be (N, n)."""
    if isinstance(item, np.ndarray):
        if item.ndim == 0:
            return np.full(length, item, dtype=dtype)
        elif item.ndim <= max_ndim:
            return np.ascontiguousarray(item, dtype=dtype)
        else:
            raise ValueError(f'Dimensionality of y0 is over {max_ndim}. y0 = {item}')
    elif isinstance(item, Iterable):
        return _into_1d_typearray(np.array(item, dtype=dtype), length, dtype, max_ndim)
    else:
        return np.full(length, item, dtype=dtype)


def convert(y0, rtol, atol):
    """Converts y0 and tolerances into float64 arrays; tolerances broadcast to len(y0) (_aux.py:134-140)."""
    y0 = _into_1d_typearray(y0, max_ndim=2)
    n = y0.shape[-1]
    return y0, _into_1d_typearray(rtol, n), _into_1d_typearray(atol, n)


class StepMask(np.ndarray):
    """Per-trajectory bool result of step()/ff2t() on an ensemble; truthy while ANY trajectory is."""

    def __bool__(self):
        return bool(np.ndarray.any(np.asarray(self)))


def _mask(a):
    return np.asarray(a, dtype=bool).view(StepMask)


class Records:
    """Accepted steps drained from the HBM record log."""

    def __init__(self, traj, step, t, y, aux):
        self.traj, self.step, self.t, self.y, self.aux = traj, step, t, y, aux

    def __len__(self):
        return len(self.t)

    def sorted(self):
        """Records ordered by (trajectory, step index)."""
        o = np.lexsort((self.step, self.traj))
        return Records(self.traj[o], self.step[o], self.t[o], self.y[o], self.aux[o])

    def trajectory(self, i):
        r = self.sorted()
        m = r.traj == i
        return r.t[m], r.y[m], r.aux[m]

    def to_npz(self, path):
        """Write the records to a compressed .npz (columns traj, step, t, y, aux)."""
        np.savez_compressed(path, traj=self.traj, step=self.step, t=self.t, y=self.y, aux=self.aux)

    def to_parquet(self, path):
        """Write the records to a Parquet file: one row per accepted step, columns traj, step, t, y0.., aux0.."""
        import pyarrow as pa
        import pyarrow.parquet as pq
        cols = {'traj': self.traj, 'step': self.step, 't': self.t}
        cols.update({f'y{i}': np.ascontiguousarray(self.y[:, i]) for i in range(self.y.shape[1])})
        cols.update({f'aux{i}': np.ascontiguousarray(self.aux[:, i]) for i in range(self.aux.shape[1])})
        pq.write_table(pa.table(cols), path)


class Ensemble:
    """Device-resident solver state: N trajectories of one ODE on one GPU.

    Attribute names follow the reference's jitclass spec (`_basic.py:182-199`).
    """

    def __init__(self, fun, t0, y0, t_bound, max_step, rtol, atol, first_step, method, advanced=False,
                 parameters=None, device=None, record=0, allow_reference_backward_bug=False, arithmetic='strict',
                 semantics='reference'):
        self._h = None
        self._mod = None
        self._L = _lib.lib()
        self.fun = as_rhs(fun)
        y0, rtol, atol = convert(y0, rtol, atol)
        self._single = y0.ndim == 1
        y0 = np.atleast_2d(y0)
        self.N, self.n = y0.shape
        if self.fun.n is not None and self.fun.n != self.n:
            raise ValueError(f'{self.fun!r} has {self.fun.n} components, y0 has {self.n}')
        if len(rtol) != self.n or len(atol) != self.n:
            raise ValueError('rtol / atol must be scalars or have one entry per component')
        self._method = method
        self._advanced = bool(advanced)
        order, self.n_stages, self.A, self.B, self.C, self.E = _TABLEAU[method]
        self.error_exponent = -1 / (order + 1)
        self.rtol, self.atol = rtol, atol
        # max_step / first_step: scalars as in the reference (_basic.py:292-299), or one value per trajectory
        self._max_step_v = self._first_step_v = None
        if np.ndim(max_step) > 0:
            self._max_step_v = np.ascontiguousarray(np.broadcast_to(np.asarray(max_step, dtype=np.float64), (self.N,)))
            max_step = np.inf if self.N != 1 else self._max_step_v[0]
        if np.ndim(first_step) > 0:
            self._first_step_v = np.ascontiguousarray(np.broadcast_to(np.asarray(first_step, dtype=np.float64), (self.N,)))
            first_step = 0.0 if self.N != 1 else self._first_step_v[0]
        self._max_step = float(max_step)
        self._first_step = float(first_step)
        t0a = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (self.N,)))
        tba = np.ascontiguousarray(np.broadcast_to(np.asarray(t_bound, dtype=np.float64), (self.N,)))
        if semantics not in ('reference', 'scipy'):
            raise ValueError("semantics must be 'reference' (numba-integrators, quirks included) or 'scipy'")
        self.semantics = semantics
        if not self._advanced and semantics == 'reference' and not allow_reference_backward_bug and np.any(tba < t0a):
            raise ValueError('the basic RK23/RK45 ignore `direction` when forming the step (_basic.py:148) and '
                             'never reach a t_bound < t0; build the solver through ni.Advanced for backward runs')
        if isinstance(self.fun, PythonRHS):  # translate the Python source into a CUDA snippet (never executed)
            self.fun = self.fun.specialise(self.n, parameters, self._advanced)
        if parameters is None:
            parameters = self.fun.bound_parameters
        P = self.fun.P
        if P:
            if parameters is None:
                raise ValueError(f'{self.fun!r} needs {P} parameters per trajectory')
            par = _flatten_parameters(parameters)
            if par.ndim == 1:
                par = par.reshape(1, -1)
            if par.shape[-1] != P:
                raise ValueError(f'{self.fun!r} takes {P} parameters, got {par.shape[-1]}')
            par = np.ascontiguousarray(np.broadcast_to(par, (self.N, P)))
        else:
            par = None
        self.P, self.Q = P, self.fun.Q
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0'))
        self.device = int(device)

        if arithmetic not in ('strict', 'fast'):
            raise ValueError("arithmetic must be 'strict' (the reference's unfused FP64 arithmetic, bit-faithful) or "
                             "'fast' (contracted multiply-adds, approximate reciprocal / controller root)")
        self.arithmetic = arithmetic
        self._mod = self.fun._module(self.n, method, (_lib.MODE_ADVANCED if self._advanced else 0)
                                     | (_lib.MODE_FAST if arithmetic == 'fast' else 0)
                                     | (_lib.MODE_SCIPY if semantics == 'scipy' else 0))
        kind = C.c_int(0)
        check(self._L.ni_module_layout(self._mod, kind))
        self.kind = kind.value  # 0: component-major device buffers, 1: trajectory-major
        h = _lib._vp()
        check(self._L.ni_ensemble_create(self._mod, self.device, self.N, h))
        self._h = h
        if record:
            self.record_enable(record)
        if self._max_step_v is not None or self._first_step_v is not None:
            check(self._L.ni_ensemble_set_step_limits(
                self._h, None if self._max_step_v is None else self._max_step_v.ctypes.data,
                None if self._first_step_v is None else self._first_step_v.ctypes.data))
        check(self._L.ni_ensemble_init(self._h, t0a.ctypes.data, 1, y0.ctypes.data,
                                       par.ctypes.data if par is not None else None, tba.ctypes.data, 1,
                                       rtol.ctypes.data_as(_lib._dp), atol.ctypes.data_as(_lib._dp), self._max_step,
                                       self._first_step))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, '_h', None):
            self._L.ni_ensemble_destroy(self._h)
            self._h = None
        if getattr(self, '_mod', None):
            if not getattr(self.fun, 'shared_module', False):  # NVRTC modules are cached per process (_rhs.py)
                self._L.ni_module_destroy(self._mod)
            self._mod = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ field access
    def _get(self, field, comps, dtype=np.float64):
        shape = (self.N,) if comps is None else (self.N, comps)
        out = np.empty(shape, dtype=dtype)
        if out.size:
            check(self._L.ni_ensemble_get(self._h, field, out.ctypes.data))
        return out

    def _scalar_or_array(self, a):
        if self._single:
            return a[0].item() if a.ndim == 1 else a[0]
        return a

    # The reference's jitclass fields are plain attributes (_basic.py:182-199): everything that is state can also be
    # assigned (the caller keeps t, y and K[-1] consistent, exactly as with the reference).
    def _set_rows(self, field, value, comps):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (self.N, comps)))
        check(self._L.ni_ensemble_set(self._h, field, a.ctypes.data))

    @property
    def t(self):
        return self._scalar_or_array(self._get(_lib.F_T, None))

    @t.setter
    def t(self, value):
        self._set(_lib.F_T, value)

    @property
    def y(self):
        return self._scalar_or_array(self._get(_lib.F_Y, self.n))

    @y.setter
    def y(self, value):
        self._set_rows(_lib.F_Y, value, self.n)

    @property
    def auxiliary(self):
        a = self._get(_lib.F_AUX, self.Q)
        if getattr(self.fun, 'aux_scalar', False):  # the Python RHS returned a scalar auxiliary (README.md:84)
            a = a[:, 0]
        return self._scalar_or_array(a)
    @property
    def step_size(self):
        return self._scalar_or_array(self._get(_lib.F_STEP_SIZE, None))

    @step_size.setter
    def step_size(self, value):
        self._set(_lib.F_STEP_SIZE, value)

    @property
    def max_step(self):  # _basic.py:188: a scalar field of the jitclass; one value per trajectory when given as an array
        if self._max_step_v is not None:
            return self._scalar_or_array(self._max_step_v.copy())
        return self._max_step

    direction = property(lambda s: s._scalar_or_array(s._get(_lib.F_DIRECTION, None)))
    parameters = property(lambda s: s._scalar_or_array(s._get(_lib.F_PARAMS, s.P)))
    status = property(lambda s: s._scalar_or_array(s._get(_lib.F_STATUS, None, np.int32)))
    n_accepted = property(lambda s: s._scalar_or_array(s._get(_lib.F_N_ACCEPTED, None, np.int32)))
    n_rejected = property(lambda s: s._scalar_or_array(s._get(_lib.F_N_REJECTED, None, np.int32)))

    @property
    def nfev(self):
        """RHS evaluations: 1 (FSAL seed) + 1 (select_initial_step) + n_stages per attempt."""
        if self._first_step_v is not None:
            n_init = self._scalar_or_array(np.where(self._first_step_v == 0.0, 2, 1))
        else:
            n_init = 2 if self._first_step == 0.0 else 1
        return n_init + self.n_stages * (self.n_accepted + self.n_rejected)

    def _set(self, field, value):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64), (self.N,)))
        check(self._L.ni_ensemble_set(self._h, field, a.ctypes.data))

    @property
    def t_bound(self):
        return self._scalar_or_array(self._get(_lib.F_T_BOUND, None))

    @t_bound.setter
    def t_bound(self, value):  # the one field the reference API itself writes (_API.py:33,37)
        self._set(_lib.F_T_BOUND, value)

    @property
    def h_abs(self):
        return self._scalar_or_array(self._get(_lib.F_H_ABS, None))

    @h_abs.setter
    def h_abs(self, value):
        self._set(_lib.F_H_ABS, value)

    @property
    def K(self):
        """Stage matrix (n_stages+1, n).  Only K[-1] (the FSAL stage, the one that carries state
        between steps) is kept on the device; rows 0..n_stages-1 are scratch in the reference
        (overwritten by every attempt) and read NaN here.  Assigning K writes K[-1]."""
        kl = self._get(_lib.F_KLAST, self.n)
        K = np.full((self.N, self.n_stages + 1, self.n), np.nan)
        K[:, -1, :] = kl
        return K[0] if self._single else K

    @K.setter
    def K(self, value):
        v = np.asarray(value, dtype=np.float64)
        if v.shape[-2:] == (self.n_stages + 1, self.n):  # a whole stage matrix: only the FSAL row is state
            v = v[..., -1, :]
        self._set_rows(_lib.F_KLAST, v, self.n)

    @property
    def state(self):  # _basic.py:269-271 / _advanced.py:276-278
        return (self.t, self.y, self.auxiliary) if self._advanced else (self.t, self.y)

    # ------------------------------------------------------------------ stepping
    def step(self):
        """One `solver.step()` on every trajectory (_basic.py:244-267)."""
        n_true = C.c_int64(0)
        check(self._L.ni_ensemble_step(self._h, n_true))
        if self._single:
            return bool(n_true.value)
        return _mask(self._get(_lib.F_RUNNING, None, np.uint8))

    def poll(self):
        """Synchronise and report how the last run / run_async / step ended: (unfinished, failed) trajectory counts --
        still RUNNING (attempt budget, full record log) and FAILED / NONFINITE.  Raises NiError(-6) on a full log."""
        a, b = C.c_int64(0), C.c_int64(0)
        check(self._L.ni_ensemble_poll(self._h, a, b))
        return a.value, b.value

    def run(self):
        """`while step(): pass` for every trajectory, on the device.  Returns the number of trajectories still RUNNING
        (0 unless the record log filled up); failed / non-finite ones are counted by `poll()` and named by `status`."""
        left = C.c_int64(0)
        check(self._L.ni_ensemble_run(self._h, left))
        return left.value

    def ff2t(self, t_end):
        flags = np.zeros(self.N, dtype=np.uint8)
        check(self._L.ni_ensemble_ff2t(self._h, float(t_end), flags.ctypes.data))
        return bool(flags[0]) if self._single else _mask(flags)

    def ff2cond_user(self, body, cond_params):
        """ff2cond with a CUDA predicate body `bool cond(double t, const double* y, const double* aux, const double* cp)`."""
        cp = np.ascontiguousarray(np.atleast_1d(np.asarray(cond_params, dtype=np.float64)))
        hit = np.zeros(self.N, dtype=np.uint8)
        n_hit = C.c_int64(0)
        check(self._L.ni_ensemble_ff2cond_user(self._h, str(body).encode(), cp.ctypes.data_as(_lib._dp), int(cp.size),
                                               hit.ctypes.data, n_hit))
        return bool(hit[0]) if self._single else _mask(hit)

    def ff2cond_device(self, kind, index, value):
        hit = np.zeros(self.N, dtype=np.uint8)
        n_hit = C.c_int64(0)
        check(self._L.ni_ensemble_ff2cond(self._h, int(kind), int(index), float(value), hit.ctypes.data, n_hit))
        return bool(hit[0]) if self._single else _mask(hit)

    # ------------------------------------------------------------------ bulk host<->device paths
    def set_stream(self, cuda_stream):
        """Launch on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        check(self._L.ni_ensemble_set_stream(self._h, _lib._vp(int(cuda_stream))))

    def upload(self, t0, y0, parameters, t_bound):
        """Asynchronous host -> HBM copy of new inputs: y0 (N, n), parameters (N, P) row-major float64
        buffers (numpy arrays or pinned torch tensors), t0 / t_bound scalars or (N,) buffers.  Buffers of another
        dtype / layout are converted (numpy) or refused (torch); they are kept alive until the next `sync()`."""
        t0p, t0n = _buf(t0, self.N)
        tbp, tbn = _buf(t_bound, self.N)
        y0b = _host_buffer(y0, np.float64, self.N * self.n, 'y0')
        parb = _host_buffer(parameters, np.float64, self.N * self.P, 'parameters') if self.P else None
        self._keep = (t0p, tbp, y0b, parb)  # the copies are asynchronous: the sources must outlive them
        check(self._L.ni_ensemble_upload(self._h, _ptr(t0p), int(t0n), _ptr(y0b), _ptr(parb) if self.P else None,
                                         _ptr(tbp), int(tbn)))

    def reset(self):
        """Re-run the constructor math (FSAL seed, select_initial_step) on the resident inputs (async)."""
        check(self._L.ni_ensemble_reset(self._h, self.rtol.ctypes.data_as(_lib._dp),
                                        self.atol.ctypes.data_as(_lib._dp), self._max_step, self._first_step))

    def run_async(self):
        check(self._L.ni_ensemble_run_async(self._h))

    def sync(self):
        check(self._L.ni_ensemble_sync(self._h))
        self._keep = None

    _FIELD_SHAPE = {_lib.F_T: ('f8', 1), _lib.F_H_ABS: ('f8', 1), _lib.F_STEP_SIZE: ('f8', 1), _lib.F_T_BOUND: ('f8', 1),
                    _lib.F_DIRECTION: ('f8', 1), _lib.F_STATUS: ('i4', 1), _lib.F_N_ACCEPTED: ('i4', 1),
                    _lib.F_N_REJECTED: ('i4', 1), _lib.F_RUNNING: ('u1', 1), _lib.F_COND_HIT: ('u1', 1)}

    def fetch(self, field, out):
        """Device -> host copy of one field into a caller-owned buffer (numpy array / pinned tensor), which must be
        C-contiguous and have the field's dtype and size (checked: the copy would otherwise overrun or misread it)."""
        field = int(field)
        if field in self._FIELD_SHAPE:
            dt, comps = self._FIELD_SHAPE[field]
        else:
            dt, comps = 'f8', {_lib.F_Y: self.n, _lib.F_KLAST: self.n, _lib.F_AUX: self.Q, _lib.F_PARAMS: self.P}.get(field, 1)
        buf = _host_buffer(out, np.dtype(dt), self.N * comps, f'fetch(field {field})', writable=True)
        check(self._L.ni_ensemble_get(self._h, field, _ptr(buf)))
        return out

    def compaction(self, park_below=-1):
        """Set the tail-compaction threshold (0 disables: the default; 1..32; < 0 leaves it) and return the number of
        trajectories that were parked for a later launch in the most recent run (ni_ensemble_compaction)."""
        parked = C.c_int64(0)
        check(self._L.ni_ensemble_compaction(self._h, int(park_below), parked))
        return parked.value

    # ------------------------------------------------------------------ terminal records (final gather)
    def terminal_dtype(self):
        """numpy dtype of one packed terminal record (ni_ensemble_pack_terminal): t, y, aux, counters, status."""
        return terminal_dtype(self.n, self.Q)

    def pack_terminal(self, dst_device_ptr):
        """Asynchronously pack N terminal records into a caller-owned device buffer (N * terminal_dtype().itemsize B)."""
        check(self._L.ni_ensemble_pack_terminal(self._h, _lib._vp(int(dst_device_ptr))))

    def fetch_terminal(self, out):
        """Pack + one device-to-host copy of the N terminal records into a caller-owned byte buffer (numpy uint8 array
        or pinned torch tensor of N * terminal_dtype().itemsize bytes)."""
        buf = _host_buffer(out, np.uint8, self.N * self.terminal_dtype().itemsize, 'fetch_terminal', writable=True)
        check(self._L.ni_ensemble_get_terminal(self._h, _ptr(buf)))
        return out

    def pack_terminal_peers(self, dst_device_ptrs):
        """Pack the terminal records and store them to every buffer of `dst_device_ptrs` at once (peer memory of the
        other GPUs: the fused gather of distributed.PeerGather)."""
        arr = (_lib._vp * len(dst_device_ptrs))(*[int(p) for p in dst_device_ptrs])
        check(self._L.ni_ensemble_pack_terminal_peers(self._h, arr, len(dst_device_ptrs)))

    # real stability interval of the two pairs (|R(z)| <= 1 for z in [-limit, 0]): Bogacki-Shampine 3(2), Dormand-Prince 5(4)
    STABILITY_LIMIT = {_lib.RK23: 2.51, _lib.RK45: 3.31}

    def stiffness(self, iterations=12):
        """Stiffness probe (the reference has none).  For every trajectory: `rho`, an estimate of the
        spectral radius of df/dy at its current (t, y) (nonlinear power iteration on the device, any right-hand side),
        `h_rho = h_abs * rho`, and `stiff = h_rho > 0.5 * limit` with `limit` the real stability interval of the method
        (RK45 3.31, RK23 2.51): the controller's step size sits at the stability boundary instead of being chosen by
        the tolerance -- Hairer & Wanner's test in DOPRI5 -- so an explicit pair pays in rejected steps (C5: 0.6 per
        accepted step).  Nothing is modified.  Ask between steps of a run (`ni.step`, `ff2cond`): the step that is clipped
        to `t_bound` / to the end of an `ff2t` leg leaves a small `h_abs` behind (_basic.py:153-156)."""
        rho = np.empty(self.N)
        check(self._L.ni_ensemble_stiffness(self._h, int(iterations), rho.ctypes.data))
        limit = self.STABILITY_LIMIT[self._method]
        h_rho = np.atleast_1d(self.h_abs) * rho
        out = {'rho': rho, 'h_rho': h_rho, 'limit': limit, 'stiff': h_rho > 0.5 * limit}
        return {k: (v[0] if self._single and isinstance(v, np.ndarray) else v) for k, v in out.items()}

    def set_order(self, order=None):
        """Launch order of the work queue: `order` is a permutation of range(N) (the persistent kernel pops order[0],
        order[1], ...), None restores 0, 1, 2, ...  Results are independent of it; the tail of a run is not: hand out the
        expensive trajectories first (see `order_by_cost`)."""
        if order is None:
            check(self._L.ni_ensemble_set_order(self._h, None))
            return
        o = np.ascontiguousarray(order, dtype=np.uint32)
        if o.shape != (self.N,):
            raise ValueError(f'order must have {self.N} entries')
        check(self._L.ni_ensemble_set_order(self._h, o.ctypes.data))

    def order_by_cost(self, cost=None):
        """Longest-processing-time-first launch order: trajectories sorted by decreasing `cost` (default: the attempts
        n_accepted + n_rejected they took so far -- the previous run or leg of the same ensemble, the natural predictor
        for repeated sweeps and for integration in legs; sorted on the device, nothing crosses PCIe).  Returns the order when
        it was computed here from `cost`, None for the device sort (`get_order()` fetches it)."""
        if cost is None:  # on the device: counting sort by the attempts so far (~1 ms for 10 M trajectories)
            check(self._L.ni_ensemble_order_by_attempts(self._h))
            return None
        order = np.argsort(-np.asarray(cost), kind='stable').astype(np.uint32)
        self.set_order(order)
        return order

    def get_order(self):
        """The launch order in use (the identity if none was set)."""
        o = np.empty(self.N, dtype=np.uint32)
        check(self._L.ni_ensemble_get_order(self._h, o.ctypes.data))
        return o

    def guard_check(self):
        """(guarded buffers, overwritten guard bytes) of an ensemble created with NI_B200_GUARD=1 in the environment:
        every device buffer then sits between two 64 KiB zones of 0xA5 and any out-of-bounds write of a kernel shows up
        here (compute-sanitizer is not available on the GPU pool)."""
        nb, bad = C.c_int64(0), C.c_int64(0)
        check(self._L.ni_ensemble_guard_check(self._h, nb, bad))
        return nb.value, bad.value

    def terminal_sink_dtype(self):
        """numpy dtype of one row of an in-kernel terminal sink: the fields of terminal_dtype(), rows padded to a multiple
        of 16 bytes (ni_ensemble_terminal_sink_bytes: 64 for Lorenz-63)."""
        b = C.c_int64(0)
        check(self._L.ni_ensemble_terminal_sink_bytes(self._h, b))
        dt = self.terminal_dtype()
        return np.dtype({'names': list(dt.names), 'formats': [dt.fields[k][0] for k in dt.names],
                         'offsets': [dt.fields[k][1] for k in dt.names], 'itemsize': b.value})

    def set_terminal_sink(self, dst_device_ptrs, first_row=0):
        """The gather fused into the integration: from now on every trajectory that leaves a run kernel for good stores
        its terminal record (terminal_sink_dtype()) to row first_row + i of every buffer of `dst_device_ptrs` (device
        memory of this or of peer GPUs; up to 8).  An empty list switches the sink off."""
        ptrs = [int(p) for p in dst_device_ptrs]
        arr = (_lib._vp * max(1, len(ptrs)))(*ptrs)
        check(self._L.ni_ensemble_set_terminal_sink(self._h, arr, len(ptrs), int(first_row)))

    def terminal(self):
        """The terminal records on the host: structured array (t, y, aux, n_accepted, n_rejected, status), one D2H copy."""
        out = np.empty(self.N, dtype=self.terminal_dtype())
        check(self._L.ni_ensemble_get_terminal(self._h, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ recording
    def record_enable(self, capacity):
        check(self._L.ni_ensemble_record_enable(self._h, int(capacity)))
        self._rec_cap = int(capacity)

    def record_count(self):
        c = C.c_int64(0)
        check(self._L.ni_ensemble_record_count(self._h, c))
        return c.value

    def record_drain(self) -> Records:
        c = self.record_count()
        traj, step = np.empty(c, np.uint32), np.empty(c, np.uint32)
        t, y, aux = np.empty(c), np.empty((c, self.n)), np.empty((c, self.Q))
        n_out = C.c_int64(0)
        check(self._L.ni_ensemble_record_drain(self._h, traj.ctypes.data, step.ctypes.data, t.ctypes.data,
                                               y.ctypes.data, aux.ctypes.data if self.Q else None, c, n_out))
        k = n_out.value
        return Records(traj[:k], step[:k], t[:k], y[:k], aux[:k])

    def record_views(self):
        """Zero-copy torch views of the HBM record log for device-side consumers: dict with `count` (slots in
        use), traj (uint32 viewed as int32), t, y, aux.  Layout: y is [n][capacity] for the thread-per-trajectory
        kernels, [capacity][n] for the CTA-per-trajectory ones; slots whose traj is -1 (0xffffffff) were reserved but
        not written.  The log stores no step numbers: the records of one trajectory appear in step order (slot order),
        `record_drain()` numbers them."""
        import torch
        from .distributed import _DeviceView
        cap, count = self._rec_cap, self.record_count()
        dev = torch.device('cuda', self.device)

        def view(field, shape, ts):
            return torch.as_tensor(_DeviceView(self.device_pointer(field), shape, ts), device=dev)
        small = self.kind == 0
        out = {'count': count, 'traj': view(_lib.F_REC_TRAJ, (cap,), '<i4'), 't': view(_lib.F_REC_T, (cap,), '<f8'),
               'y': view(_lib.F_REC_Y, (self.n, cap) if small else (cap, self.n), '<f8')}
        if self.Q:
            out['aux'] = view(_lib.F_REC_AUX, (self.Q, cap) if small else (cap, self.Q), '<f8')
        return out

    # ------------------------------------------------------------------ misc
    def device_pointer(self, field):
        p = _lib._vp()
        check(self._L.ni_ensemble_device_ptr(self._h, field, p))
        return p.value

    def timing(self):
        a, b, c = C.c_float(0), C.c_float(0), C.c_int64(0)
        check(self._L.ni_ensemble_timing(self._h, a, b, c))
        return dict(ms_init=a.value, ms_run=b.value, launches=c.value)


def terminal_dtype(n, Q):
    """Packed terminal record of include/ni_b200.h: double t, y[n], aux[Q]; int32 n_accepted, n_rejected, status."""
    fields = [('t', '<f8'), ('y', '<f8', (n,))] + ([('aux', '<f8', (Q,))] if Q else []) + \
             [('n_accepted', '<i4'), ('n_rejected', '<i4'), ('status', '<i4')]
    dt = np.dtype(fields)
    assert dt.itemsize == 8 * (1 + n + Q) + 12
    return dt


def _ptr(x):
    """Address of a host buffer already validated by `_host_buffer` / `_buf` (numpy array or torch tensor)."""
    if x is None:
        return None
    if hasattr(x, 'data_ptr'):  # torch tensor (e.g. pinned host memory)
        return _lib._vp(x.data_ptr())
    return _lib._vp(x.ctypes.data)


def _host_buffer(x, dtype, count, what, writable=False):
    """A C-contiguous host buffer of exactly `count` elements of `dtype` backing `x`.

    numpy input of another dtype / layout is converted (the caller keeps the result alive until the asynchronous copy
    has completed); an output buffer (`writable`) and torch tensors cannot be converted behind the caller's back and
    must already match."""
    dtype = np.dtype(dtype)
    if hasattr(x, 'data_ptr'):  # torch tensor
        import torch
        want = {np.dtype('f8'): torch.float64, np.dtype('i4'): torch.int32, np.dtype('u1'): torch.uint8}[dtype]
        if x.dtype != want or not x.is_contiguous() or x.numel() != count or x.device.type != 'cpu':
            raise ValueError(f'{what}: need a contiguous CPU tensor of {count} x {want}, got {tuple(x.shape)} {x.dtype} '
                             f'on {x.device}{"" if x.is_contiguous() else " (not contiguous)"}')
        return x
    if writable:
        if not (isinstance(x, np.ndarray) and x.dtype == dtype and x.flags.c_contiguous and x.flags.writeable
                and x.size == count):
            raise ValueError(f'{what}: need a writable C-contiguous {dtype} array of {count} elements')
        return x
    a = np.ascontiguousarray(x, dtype=dtype)
    if a.size != count:
        raise ValueError(f'{what}: expected {count} values, got {a.size}')
    return a


def _buf(v, N):
    """(buffer, per_trajectory) for a scalar or an (N,) float64 buffer."""
    if hasattr(v, 'data_ptr'):
        return v, v.numel() > 1
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    if a.size not in (1, N):
        raise ValueError(f'expected a scalar or {N} values, got {a.size}')
    return a, a.size > 1 or N == 1


def _flatten_parameters(parameters):
    """Reference parameters can be any numba-typed value (README.md:88 uses `(2., array((-1., 1.)))`);
    here they are a flat float64 vector per trajectory."""
    if isinstance(parameters, np.ndarray):
        return parameters.astype(np.float64, copy=False)
    if isinstance(parameters, (tuple, list)) and any(isinstance(p, (np.ndarray, list, tuple)) for p in parameters):
        return np.concatenate([np.ravel(np.asarray(p, dtype=np.float64)) for p in parameters])
    return np.atleast_1d(np.asarray(parameters, dtype=np.float64))


# Type aliases of the reference API (`_basic.py:183`, `_aux.py:34`)
RK = Ensemble
Solver = Ensemble


# ---------------------------------------------------------------------- factories
def RK23(fun, t0, y0, t_bound, max_step=np.inf, rtol=1e-3, atol=1e-6, first_step=0, **kw):  # _basic.py:291-302
    return Ensemble(fun, t0, y0, t_bound, max_step, rtol, atol, first_step, _lib.RK23, **kw)


def RK45(fun, t0, y0, t_bound, max_step=np.inf, rtol=1e-3, atol=1e-6, first_step=0., **kw):  # _basic.py:323-333
    return Ensemble(fun, t0, y0, t_bound, max_step, rtol, atol, first_step, _lib.RK45, **kw)


ALL = (RK23, RK45)


class Solvers(enum.Enum):  # _basic.py:336-339: RK23/RK45 are plain functions, only ALL is a member
    RK23 = RK23
    RK45 = RK45
    ALL = ALL


def Advanced(parameters_signature, auxiliary_signature, solver):
    """Factory of solvers whose RHS takes per-trajectory `parameters` and returns an `auxiliary`
    output (_advanced.py:280-368).  The numba type signatures of the reference become optional
    length checks: pass ints (P, Q) or anything else (ignored; the RHS object knows P and Q)."""

    def _check(fun):
        f = as_rhs(fun)
        if isinstance(parameters_signature, int) and parameters_signature != f.P:
            raise ValueError(f'parameters_signature says {parameters_signature} parameters, {f!r} takes {f.P}')
        if isinstance(auxiliary_signature, int) and auxiliary_signature != f.Q:
            raise ValueError(f'auxiliary_signature says {auxiliary_signature} outputs, {f!r} returns {f.Q}')
        return f

    def RK23_advanced(fun, t0, y0, parameters, t_bound, max_step=np.inf, rtol=1e-3, atol=1e-6, first_step=0., **kw):
        return Ensemble(_check(fun), t0, y0, t_bound, max_step, rtol, atol, first_step, _lib.RK23, advanced=True,
                        parameters=parameters, **kw)

    def RK45_advanced(fun, t0, y0, parameters, t_bound, max_step=np.inf, rtol=1e-3, atol=1e-6, first_step=0., **kw):
        return Ensemble(_check(fun), t0, y0, t_bound, max_step, rtol, atol, first_step, _lib.RK45, advanced=True,
                        parameters=parameters, **kw)

    if solver is RK23:
        return RK23_advanced
    if solver is RK45:
        return RK45_advanced
    if solver in (Solvers.ALL, ALL):
        return {Solvers.RK23: RK23_advanced, Solvers.RK45: RK45_advanced}
    raise ValueError('solver must be ni.RK23, ni.RK45 or ni.Solvers.ALL')


# ---------------------------------------------------------------------- drivers (_API.py:22-49)
def step(solver) -> bool:
    return solver.step()


def ff2t(solver, t_end):
    """Fast forwards to given time or t_bound."""
    return solver.ff2t(t_end)


class Condition:
    """Device-side predicate for ff2cond: `kind` compares y[index] / auxiliary[index] with the
    `parameters` value handed to ff2cond."""

    def __init__(self, kind, index):
        self.kind, self.index = kind, index


class CudaCondition:
    """User predicate for ff2cond as CUDA C++: the body of
    `bool cond(double t, const double* y, const double* aux, const double* cp)`; `cp` receives the
    `parameters` argument of ff2cond (up to 8 doubles)."""

    def __init__(self, body):
        self.body = str(body)


def y_above(index=0):
    return Condition(_lib.COND_Y_GT, index)


def y_below(index=0):
    return Condition(_lib.COND_Y_LT, index)


def aux_above(index=0):
    return Condition(_lib.COND_AUX_GT, index)


def aux_below(index=0):
    return Condition(_lib.COND_AUX_LT, index)


def ff2cond(solver, condition, parameters):
    """Fast forwards until `condition(solver, parameters)` holds after an accepted step.

    `condition` is either a device predicate (ni.y_above(i), ...) evaluated inside the kernel, or
    any Python callable `(solver, parameters) -> bool`, evaluated on the host after every
    lock-step `step()` exactly like _API.py:45-49."""
    if isinstance(condition, Condition):
        return solver.ff2cond_device(condition.kind, condition.index, parameters)
    if isinstance(condition, CudaCondition):
        return solver.ff2cond_user(condition.body, parameters)
    if callable(condition):
        # a Python predicate: translate its source and evaluate it inside the kernel; predicates outside the
        # translatable subset fall back to the reference's own host loop over lock-step step() calls
        from ._translate import TranslationError, translate_condition
        try:
            body, cp = translate_condition(condition, solver.n, solver.Q, getattr(solver.fun, 'aux_scalar', False),
                                           parameters)
        except TranslationError:
            body = None
        if body is not None:
            return solver.ff2cond_user(body, cp)
    while solver.step():
        if condition(solver, parameters):
            return True
    return False
