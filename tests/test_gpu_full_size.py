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


@pytest.mark.parametrize('shuffled', [False, True], ids=['sorted', 'shuffled'])
def test_c4_vanderpol_4m_sweep(oracle, shuffled):
    """Config C4 at its full size: 4 M Van der Pol trajectories, mu swept 0.1..20, RK23, tol 1e-8 -- as the sorted sweep
    (neighbouring lanes have near-equal step counts) and SHUFFLED (every warp mixes 400-step and 8000-step
    trajectories: the divergence / refill stress of SURVEY.md section 8d).  Invariants over all 4 M, parity on a
    sample stratified over the index range, and the two orders must give every trajectory the same result."""
    N = 4_000_000
    w = wl.c4_vanderpol(N=N, tol=1e-8, shuffled=shuffled)
    ens = ni.RK23(ni.rhs.vanderpol(w.params[:, 0]), w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol)
    assert ens.run() == 0
    rec = ens.terminal()
    acc, rej, y = rec['n_accepted'], rec['n_rejected'], rec['y']
    assert np.all(rec['status'] == 1) and np.all(rec['t'] == w.t_bound) and np.all(np.isfinite(y))
    assert acc.max() > 4 * acc.min() and abs(acc.mean() - 4642) < 60 and abs(rej.mean() - 1076) < 40
    idx = _stratified(N, 500, 13)
    ref = oracle.run_ensemble(w.rhs, oracle.RK23, w.t0, w.y0[idx], w.t_bound, params=w.params[idx], rtol=w.rtol,
                              atol=w.atol)
    same = (acc[idx] == ref['n_accepted']) & (rej[idx] == ref['n_rejected'])
    rel = np.max(np.abs(y[idx] - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    assert same.mean() >= 0.999, (same.mean(), idx[~same][:20])
    assert np.mean(rel <= 1e-12) >= 0.97, np.mean(rel <= 1e-12)                # appendix C floor: 98.7 %
    if shuffled:
        # the same (mu, y0) pairs in sorted order: per-trajectory results do not depend on what shares the warp
        mu = w.params[:, 0]
        order = np.argsort(mu, kind='stable')
        e2 = ni.RK23(ni.rhs.vanderpol(mu[order]), w.t0, w.y0[order], w.t_bound, rtol=w.rtol, atol=w.atol)
        assert e2.run() == 0
        assert e2.terminal().tobytes() == rec[order].tobytes()


@pytest.mark.parametrize('arithmetic', ['strict', 'fast'])
def test_c5_heat_16384_systems_parity_horizon(oracle, arithmetic):
    """Config C5 at its full size: 16 384 heat systems of dimension 4096 (one CTA per trajectory) at the parity horizon
    t_bound = 5: invariants over every system, oracle parity on a stratified sample -- for the strict kernels and for
    the opt-in fast arithmetic (NVRTC variant of the same templates), which must meet the same bar."""
    N = 16_384
    w = wl.c5_heat(N=N, t_bound=5.0)
    ens = ni.RK45(ni.rhs.heat1d(w.params[:, 0]), w.t0, w.y0, w.t_bound, rtol=w.rtol, atol=w.atol, arithmetic=arithmetic)
    assert ens.run() == 0
    y = ens.y
    assert np.all(ens.status == 1) and np.all(ens.t == 5.0) and np.all(np.isfinite(y))
    assert np.all(np.abs(y).max(axis=1) <= np.abs(w.y0).max(axis=1))           # maximum principle
    acc, rej = ens.n_accepted, ens.n_rejected
    assert 20 <= acc.min() and acc.max() <= 90 and rej.max() <= 8, (acc.min(), acc.max(), rej.max())  # ~38 / ~0
    idx = np.unique(np.r_[0:8, N - 8:N, np.random.default_rng(5).integers(0, N, 48)])
    ref = oracle.run_ensemble(w.rhs, oracle.RK45, w.t0, w.y0[idx], w.t_bound, params=w.params[idx], rtol=w.rtol,
                              atol=w.atol)
    assert np.array_equal(ens.n_accepted[idx], ref['n_accepted']) and np.array_equal(ens.n_rejected[idx], ref['n_rejected'])
    rel = np.max(np.abs(y[idx] - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    assert np.all(rel <= 1e-12), rel.max()


def test_large_kernel_scipy_semantics_against_the_oracle_where_they_coincide(oracle):
    """semantics='scipy' on the CTA-per-trajectory kernels, first_step given: while no attempt is rejected and neither
    the interval nor max_step binds, SciPy's step rules and the reference's give the same step sequence (their
    differences are the stale FSAL stage on retries, the growth cap after a rejection and the initial step) -- so on
    the rejection-free parity horizon the oracle pins the scipy variant as well."""
    N = 64
    w = wl.c5_heat(N=N, n=1024, t_bound=5.0)
    kw = dict(rtol=w.rtol, atol=w.atol, first_step=1e-3)
    a = ni.RK45(ni.rhs.heat1d(w.params[:, 0]), w.t0, w.y0, w.t_bound, semantics='scipy', **kw)
    assert a.run() == 0
    ref = oracle.run_ensemble(w.rhs, oracle.RK45, w.t0, w.y0, w.t_bound, params=w.params, **kw)
    assert np.all(ref['n_rejected'] == 0)
    assert np.array_equal(a.n_accepted, ref['n_accepted']) and np.all(a.n_rejected == 0)
    rel = np.max(np.abs(a.y - ref['y']), axis=1) / np.max(np.abs(ref['y']), axis=1)
    assert np.all(rel <= 1e-12), rel.max()
