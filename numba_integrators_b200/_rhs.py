"""Right-hand sides accepted by the solver constructors.

The reference takes a numba first-class function `float64[:](float64, float64[:])`
(`_aux.py:64-65`) or, for Advanced, `(float64[:], aux)(float64, float64[:], params)`
(`_advanced.py:22-26`).  Here the RHS is fused into the CUDA stepper kernels, so it is either a
built-in (compiled ahead of time for sm_100a) or a CUDA C++ snippet specialised with NVRTC.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_BUILTIN_IDS = {'harmonic': 0, 'sine': 0, 'lorenz63': 1, 'vanderpol': 2, 'heat1d': 3, 'burgers1d': 4,
                'exponential': 5, 'riccati': 6, 'readme_advanced': 7, 'blowup': 8}
# name -> (n or None for "any", P, Q)
_BUILTIN_SHAPE = {'harmonic': (2, 0, 0), 'sine': (2, 0, 0), 'lorenz63': (3, 3, 1), 'vanderpol': (2, 1, 0),
                  'heat1d': (None, 1, 0), 'burgers1d': (None, 2, 0), 'exponential': (1, 0, 0), 'riccati': (1, 0, 0),
                  'readme_advanced': (2, 3, 1), 'blowup': (1, 0, 0)}


class RHS:
    """Base: `n` state dimension (None = taken from y0), `P` parameters, `Q` auxiliary outputs."""
    shared_module = False  # True: the module handle is cached for the process and must not be destroyed by a solver
    name = '?'
    n = None
    P = 0
    Q = 0
    bound_parameters = None  # parameters closed over (basic API has no `parameters` argument)

    def _module(self, n, method, mode):
        """mode: bit 0 = Advanced semantics, bit 1 = fast arithmetic (include/ni_b200.h NI_MODE_*)."""
        raise NotImplementedError


class BuiltinRHS(RHS):
    def __init__(self, name, parameters=None):
        if name not in _BUILTIN_IDS:
            raise ValueError(f'unknown built-in right-hand side {name!r}; choose from {sorted(_BUILTIN_IDS)}')
        self.name = name
        self.rhs_id = _BUILTIN_IDS[name]
        self.n, self.P, self.Q = _BUILTIN_SHAPE[name]
        self.bound_parameters = None if parameters is None else np.asarray(parameters, dtype=np.float64)

    def __call__(self, *parameters):
        """Bind parameters (scalars or per-trajectory arrays), e.g. `ni.rhs.vanderpol(mu)`."""
        if len(parameters) != self.P:
            raise ValueError(f'{self.name} takes {self.P} parameters, got {len(parameters)}')
        cols = np.broadcast_arrays(*[np.asarray(p, dtype=np.float64) for p in parameters]) if parameters else []
        par = np.stack(cols, axis=-1) if cols else None
        return BuiltinRHS(self.name, par)

    def _module(self, n, method, mode):
        out = _lib._vp()
        _lib.check(_lib.lib().ni_module_builtin(self.rhs_id, int(n), int(method), int(mode), out))
        return out

    def __repr__(self):
        return f'BuiltinRHS({self.name!r})'


_MODULE_CACHE = {}  # (kind, body, n, P, Q, method, mode, options) -> module handle; lives for the process


def _cached_module(key, make):
    """NVRTC specialisations cost 1-10 s: share them between solvers built from the same source."""
    if key not in _MODULE_CACHE:
        _MODULE_CACHE[key] = make()
    return _MODULE_CACHE[key]


class CudaRHS(RHS):
    """User right-hand side: the body of

        __device__ void rhs(double t, const double* y, const double* p, double* dy, double* aux)

    in CUDA C++.  It is fused into the RK stage loop at load time with NVRTC (sm_100a).
    """

    shared_module = True

    def __init__(self, body, n, P=0, Q=0, parameters=None, name='user', nvrtc_options=''):
        self.body, self.n, self.P, self.Q, self.name = str(body), int(n), int(P), int(Q), name
        self.nvrtc_options = nvrtc_options
        self.bound_parameters = None if parameters is None else np.asarray(parameters, dtype=np.float64)

    def _module(self, n, method, mode):
        if n != self.n:
            raise ValueError(f'y0 has {n} components, the RHS was declared with n={self.n}')
        def make():
            out = _lib._vp()
            _lib.check(_lib.lib().ni_module_compile(self.body.encode(), self.n, self.P, self.Q, int(method),
                                                    int(mode), self.nvrtc_options.encode(), out))
            return out
        return _cached_module(('dense', self.body, self.n, self.P, self.Q, int(method), int(mode), self.nvrtc_options), make)

    def __repr__(self):
        return f'CudaRHS({self.name!r}, n={self.n}, P={self.P}, Q={self.Q})'


class _CudaLargeRHS(RHS):
    """Common part of the CTA-per-trajectory snippets: an optional auxiliary function over the whole state,

        __device__ void aux(double t, const double* y, int n, const double* p, double* aux)

    writing Q outputs (the second return value of an Advanced right-hand side, _advanced.py:22-26)."""
    shared_module = True
    _component = 0

    def __init__(self, body, P=0, parameters=None, name=None, nvrtc_options='', aux_body=None, Q=0):
        self.body, self.P, self.Q, self.n = str(body), int(P), int(Q), None
        self.name = name or ('user_component' if self._component else 'user_stencil')
        if bool(self.Q) != bool(aux_body):
            raise ValueError('Q auxiliary outputs need an aux_body (and the other way round)')
        self.aux_body = None if aux_body is None else str(aux_body)
        self.nvrtc_options = nvrtc_options
        self.bound_parameters = None if parameters is None else np.asarray(parameters, dtype=np.float64)

    def _module(self, n, method, mode):
        def make():
            out = _lib._vp()
            _lib.check(_lib.lib().ni_module_compile_large(self._component, self.body.encode(),
                                                          None if self.aux_body is None else self.aux_body.encode(),
                                                          int(n), self.P, self.Q, int(method), int(mode),
                                                          self.nvrtc_options.encode(), out))
            return out
        return _cached_module((self._component, self.body, self.aux_body, int(n), self.P, self.Q, int(method), int(mode),
                               self.nvrtc_options), make)

    def __repr__(self):
        return f'{type(self).__name__}({self.name!r}, P={self.P}, Q={self.Q})'


class CudaStencilRHS(_CudaLargeRHS):
    """User right-hand side of a large 1-D method-of-lines system (n <= 4096): the body of

        __device__ double rhs_j(double t, double ym, double yc, double yp, const double* p, int j, int n)

    returning dy_j from y_{j-1}, y_j, y_{j+1} (zero outside 0..n-1).  Fused into the CTA-per-trajectory
    kernels with NVRTC; n is taken from y0."""
    _component = 0


class CudaComponentRHS(_CudaLargeRHS):
    """User right-hand side of a large system (17 <= n <= 4096 is where it pays) with arbitrary coupling, written
    component-wise: the body of

        __device__ double rhs_j(double t, const double* y, int j, int n, const double* p)

    returning dy_j.  One CTA integrates one trajectory; `y` is the stage vector in shared memory."""
    _component = 1


harmonic = BuiltinRHS('harmonic')
sine = harmonic
lorenz63 = BuiltinRHS('lorenz63')
vanderpol = BuiltinRHS('vanderpol')
heat1d = BuiltinRHS('heat1d')
burgers1d = BuiltinRHS('burgers1d')
exponential = BuiltinRHS('exponential')
riccati = BuiltinRHS('riccati')
readme_advanced = BuiltinRHS('readme_advanced')
blowup = BuiltinRHS('blowup')


class PythonRHS(RHS):
    """A reference-style Python / numba right-hand side `f(t, y)` or `f(t, y, parameters) -> (dy, aux)`.
    Its source is translated to a CUDA C++ snippet (numba_integrators_b200/_translate.py) once the state
    dimension and the parameter layout are known; the Python function itself is never executed."""

    def __init__(self, fun):
        self.fun = fun
        self.name = getattr(getattr(fun, 'py_func', fun), '__name__', 'python_rhs')
        self.n = None

    def specialise(self, n, parameters, advanced):
        from ._translate import translate
        body, P, Q, aux_scalar = translate(self.fun, n, parameters, advanced)
        rhs = CudaRHS(body, n=n, P=P, Q=Q, name=self.name)
        rhs.aux_scalar = aux_scalar
        rhs.python_source = self.fun
        return rhs

    def __repr__(self):
        return f'PythonRHS({self.name})'


def as_rhs(f) -> RHS:
    if isinstance(f, RHS):
        return f
    if isinstance(f, str):
        return BuiltinRHS(f)
    if callable(f):
        return PythonRHS(f)
    raise TypeError('the right-hand side must be a built-in name / ni.rhs.* object, a ni.CudaRHS / '
                    f'ni.CudaStencilRHS snippet or a Python function (got {type(f).__name__})')
