"""Synthetic inputs of the BASELINE.json configurations C1-C5.

Every array comes from its own `np.random.default_rng([seed, stream, shard])`, so the first
k rows of an N-row workload equal the k-row workload (sub-samples for the CPU oracle are
prefixes) and each multi-GPU shard draws an independent block.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str
    rhs: str                 # built-in RHS name
    method: str              # 'RK45' | 'RK23'
    advanced: bool           # built through ni.Advanced (h = h_abs*direction, auxiliary output)
    t0: float
    t_bound: float
    y0: np.ndarray           # (N, n)
    params: np.ndarray | None  # (N, P) or None
    rtol: float
    atol: float
    flops_per_attempt: float
    record_bytes_per_step: int  # 8*(1+n+Q)
    extra: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.y0.shape[0]

    @property
    def n(self):
        return self.y0.shape[1]


def _rng(seed, stream, shard):
    return np.random.default_rng([seed, stream, shard])


def c1_readme():
    """README sine RK45 (README.md:47-69): y0=(0,1), t_bound=1, atol=rtol=1e-8."""
    return Workload('c1_readme_sine', 'harmonic', 'RK45', False, 0.0, 1.0, np.array([[0.0, 1.0]]), None,
                    1e-8, 1e-8, 148.0, 24)


def c2_harmonic(N=1_000_000, shard=0, t_bound=100.0, tol=1e-8):
    y0 = _rng(1, 0, shard).uniform(-1.0, 1.0, size=(N, 2))
    return Workload('c2_harmonic', 'harmonic', 'RK45', False, 0.0, t_bound, y0, None, tol, tol, 148.0, 24)


def c3_lorenz(N=10_000_000, shard=0, t_bound=5.0, tol=1e-8):
    y0 = np.empty((N, 3))
    r = _rng(2, 0, shard)
    u = r.uniform(size=(N, 3))
    y0[:, 0] = -10.0 + 20.0 * u[:, 0]
    y0[:, 1] = -10.0 + 20.0 * u[:, 1]
    y0[:, 2] = 15.0 + 20.0 * u[:, 2]
    v = _rng(2, 1, shard).uniform(size=(N, 3))
    p = np.empty((N, 3))
    p[:, 0] = 9.0 + 2.0 * v[:, 0]
    p[:, 1] = 26.0 + 4.0 * v[:, 1]
    p[:, 2] = 2.5 + 0.3 * v[:, 2]
    return Workload('c3_lorenz63', 'lorenz63', 'RK45', True, 0.0, t_bound, y0, p, tol, tol, 258.0, 40)


def c4_vanderpol(N=4_000_000, shard=0, n_shards=1, t_bound=20.0, tol=1e-8, shuffled=False):
    """mu swept 0.1..20 over the GLOBAL index; shard s of n_shards takes indices s, s+n_shards, ... (interleaved)."""
    Ng = N * n_shards
    idx = np.arange(shard, Ng, n_shards, dtype=np.float64)
    mu = 0.1 + 19.9 * idx / max(Ng - 1, 1)
    if shuffled:
        mu = mu[_rng(3, 2, shard).permutation(N)]
    y0 = _rng(3, 0, shard).uniform(-2.0, 2.0, size=(N, 2))
    return Workload('c4_vanderpol', 'vanderpol', 'RK23', False, 0.0, t_bound, y0, mu.reshape(N, 1), tol, tol,
                    77.0, 24)


def c5_heat(N=16_384, n=4096, shard=0, t_bound=50.0, tol=1e-8):
    x = (np.arange(n, dtype=np.float64) + 1.0) / (n + 1.0)
    m = _rng(4, 0, shard).integers(1, 5, size=N)
    noise = _rng(4, 1, shard).standard_normal(size=(N, n))
    y0 = np.sin(np.pi * m[:, None] * x[None, :]) + 0.1 * noise
    k = _rng(4, 2, shard).uniform(0.5, 2.0, size=(N, 1))
    return Workload('c5_heat1d', 'heat1d', 'RK45', False, 0.0, t_bound, y0, k, tol, tol,
                    63.0 * n + 6 * 4.0 * n + 16.0, 8 * (1 + n))


WORKLOADS = {'c1': c1_readme, 'c2': c2_harmonic, 'c3': c3_lorenz, 'c4': c4_vanderpol, 'c5': c5_heat}
