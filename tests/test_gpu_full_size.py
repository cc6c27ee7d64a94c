"""BASELINE-size ensembles: size-independent invariants over every trajectory plus oracle parity on a
sample stratified over the whole index range (the full CPU run would take minutes)."""
import numpy as np
import pytest

import numba_integrators_b200 as ni
from numba_integrators_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _stratified(N, k, seed):
    rng = np.random.default_rng(seed)
    return np.unique(np.concatenate([np.arange(k), np.arange(N - k, N), rng.integers(0, N, size=4 * k)]))


def test_c3_lorenz_10m_parity_horizon(oracle):
    """Config C3 at its full size (10 M Lorenz trajectories, per-trajectory sigma/rho/beta, aux) at t_bound = 1."""
    N = 10_000_000
    w = wl.c3_lorenz(N=N, t_bound=1.0)
    ens = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, w.t0, w.y0, w.params, w.t_bound, rtol=w.rtol, atol=w.atol)
    assert ens.run() == 0
    t, y, aux, acc, rej, st = ens.t, ens.y, ens.auxiliary, ens.n_accepted, ens.n_rejected, ens.status
    # invariants over all 10 M trajectories
    assert np.all(st == 1) and np.all(t == 1.0)
    assert np.all(np.isfinite(y)) and np.all(np.isfinite(aux))
    assert np.array_equal(aux[:, 0], (y[:, 0] * y[:, 0] + y[:, 1] * y[:, 1]) + y[:, 2] * y[:, 2])  # aux = f(final y)
    assert 20 <= acc.min() and acc.max() <= 400 and rej.max() <= 400   # stale-FSAL rejection cascades exist
    assert abs(acc.mean() - 103.7) < 1.0 and abs(rej.mean() - 8.1) < 0.5
    # parity on ~60 k trajectories spread over the whole ensemble
    idx = _stratified(N, 10_000, 7)
    ref = oracle.run_ensemble(w.rhs, oracle.RK45, w.t0, w.y0[idx], w.t_bound, params=w.params[idx], rtol=w.rtol,
                              atol=w.atol, advanced=True)
    rel = np.max(np.abs(y[idx] - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    same = (acc[idx] == ref['n_accepted']) & (rej[idx] == ref['n_rejected'])
    divergent = idx[~(same & (rel <= 1e-12))]
    assert same.mean() >= 0.999 and np.mean(rel <= 1e-12) >= 0.999, (same.mean(), np.mean(rel <= 1e-12), divergent[:20])
    assert np.mean(np.all(y[idx] == ref['y'], axis=1)) > 0.85      # bit-identical unless glibc pow misrounded


def test_c2_harmonic_1m_recording_round_trip(oracle):
    """Config C2 at full size with step recording: every accepted step lands in the HBM log exactly once."""
    N = 1_000_000
    w = wl.c2_harmonic(N=N, t_bound=5.0)      # 1/20 of the horizon keeps the drained log at ~1.5 GB
    ens = ni.RK45(ni.rhs.harmonic, w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol, record=N * 70)
    assert ens.run() == 0
    acc = ens.n_accepted.astype(np.int64)
    rec = ens.record_drain()
    assert len(rec) == acc.sum()
    assert np.array_equal(np.bincount(rec.traj, minlength=N), acc)            # per-trajectory record counts
    last = rec.step == acc[rec.traj]                                           # the record of each final step
    assert last.sum() == N and np.all(rec.t[last] == w.t_bound)
    order = np.argsort(rec.traj[last])
    assert np.array_equal(rec.y[last][order], ens.y)
    e0 = (w.y0 ** 2).sum(axis=1)                                               # energy is conserved to the tolerance
    e = (rec.y ** 2).sum(axis=1)
    assert np.all(np.abs(e - e0[rec.traj]) <= 1e-4 * e0[rec.traj] + 1e-6)    # atol dominates for tiny amplitudes
    idx = _stratified(N, 5_000, 11)
    ref = oracle.run_ensemble(w.rhs, oracle.RK45, w.t0, w.y0[idx], w.t_bound, rtol=w.rtol, atol=w.atol)
    assert np.array_equal(acc[idx], ref['n_accepted'])
    rel = np.max(np.abs(ens.y[idx] - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    assert np.all(rel <= 1e-12)


def test_c4_vanderpol_sweep_heterogeneous_counts(oracle):
    """Config C4 (mu swept 0.1..20, RK23, tol 1e-8) on 200 k trajectories: work-queue refill under a 4x spread
    of step counts; parity on a stratified sample."""
    N = 200_000
    w = wl.c4_vanderpol(N=N, tol=1e-8)
    ens = ni.RK23(ni.rhs.vanderpol(w.params[:, 0]), w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol)
    assert ens.run() == 0
    acc, rej = ens.n_accepted, ens.n_rejected
    assert np.all(ens.status == 1) and np.all(ens.t == w.t_bound) and np.all(np.isfinite(ens.y))
    assert acc.max() > 2.5 * acc.min()                                         # heterogeneous
    idx = _stratified(N, 500, 13)
    ref = oracle.run_ensemble(w.rhs, oracle.RK23, w.t0, w.y0[idx], w.t_bound, params=w.params[idx], rtol=w.rtol,
                              atol=w.atol)
    same = (acc[idx] == ref['n_accepted']) & (rej[idx] == ref['n_rejected'])
    rel = np.max(np.abs(ens.y[idx] - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    assert same.mean() >= 0.999, same.mean()
    assert np.mean(rel <= 1e-12) >= 0.97, np.mean(rel <= 1e-12)                # appendix C floor: 98.7 %


def test_c5_heat_4096_invariants(oracle):
    """Config C5 (n = 4096 heat, one CTA per trajectory) at the parity horizon t_bound = 5 on 592 systems (4 per SM)."""
    N = 592
    w = wl.c5_heat(N=N, t_bound=5.0)
    ens = ni.RK45(ni.rhs.heat1d(w.params[:, 0]), w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol)
    assert ens.run() == 0
    y = ens.y
    assert np.all(ens.status == 1) and np.all(ens.t == 5.0) and np.all(np.isfinite(y))
    assert np.all(np.abs(y).max(axis=1) <= np.abs(w.y0).max(axis=1))           # maximum principle
    idx = np.arange(0, N, 37)
    ref = oracle.run_ensemble(w.rhs, oracle.RK45, w.t0, w.y0[idx], w.t_bound, params=w.params[idx], rtol=w.rtol,
                              atol=w.atol)
    assert np.array_equal(ens.n_accepted[idx], ref['n_accepted']) and np.array_equal(ens.n_rejected[idx], ref['n_rejected'])
    rel = np.max(np.abs(y[idx] - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    assert np.all(rel <= 1e-12), rel.max()
