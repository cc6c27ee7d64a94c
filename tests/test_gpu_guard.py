"""Memory-safety evidence without compute-sanitizer (closed on the GPU pool): ensembles created with NI_B200_GUARD=1
keep every device buffer between two 64 KiB guard zones; after exercising every kernel family on ragged sizes no guard
byte may have changed (ni_ensemble_guard_check)."""
import numpy as np
import pytest

import numba_integrators_b200 as ni
from numba_integrators_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def clean(ens):
    n_buffers, bad = ens.guard_check()
    assert n_buffers >= 15 and bad == 0, (n_buffers, bad)


def test_guard_mode_is_off_by_default():
    e = ni.RK45(ni.rhs.harmonic, 0.0, np.ones((5, 2)), 1.0)
    assert e.guard_check() == (0, 0)


def test_no_kernel_writes_outside_its_buffers(monkeypatch):
    monkeypatch.setenv('NI_B200_GUARD', '1')
    # thread-per-trajectory kernels: ragged N, recording with a log that overflows, lock-step, fields, terminal records
    w = wl.c3_lorenz(N=777, t_bound=0.5)
    a = ni.Advanced(3, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, w.y0, w.params, 0.5, rtol=1e-8, atol=1e-8, record=3000)
    for _ in range(200):
        try:
            a.run()
            a.record_drain()
            break
        except ni.NiError as exc:
            assert exc.code == -6
            a.record_drain()
    a.terminal(), a.y, a.K, a.auxiliary, a.stiffness(3)
    a.t_bound = 0.7
    more = True
    while more:                     # lock-step against the small log: drain whenever it fills up
        try:
            more = bool(ni.step(a))
        except ni.NiError as exc:
            assert exc.code == -6
        a.record_drain()
    clean(a)
    # 70 000 trajectories: chunked record slots (lean body with recording), then a second run on finished trajectories
    rng = np.random.default_rng(1)
    big = ni.RK45(ni.rhs.harmonic, 0.0, rng.uniform(-1, 1, size=(70_001, 2)), 2.0, rtol=1e-8, atol=1e-8, record=70_001 * 40)
    big.run()
    assert len(big.record_drain()) == big.n_accepted.sum()
    big.run()
    clean(big)
    # general body: finite max_step per trajectory, fast arithmetic, RK23, ff2t / ff2cond
    b = ni.RK23(ni.rhs.vanderpol(np.linspace(0.1, 5, 333)), 0.0, wl.c4_vanderpol(N=333).y0, 2.0, rtol=1e-6, atol=1e-6,
                arithmetic='fast', max_step=np.linspace(0.01, 0.1, 333))
    ni.ff2t(b, 1.0)
    ni.ff2cond(b, ni.y_above(0), 1.5)
    b.run()
    clean(b)
    # a translated Python right-hand side (NVRTC), scipy semantics
    def f(t, y):
        return np.array((y[1], -y[0] - 0.1 * y[1], y[2] * 0.0 + 1.0))
    c = ni.RK45(f, 0.0, rng.uniform(-1, 1, size=(1001, 3)), 3.0, rtol=1e-7, atol=1e-9, semantics='scipy')
    c.run()
    clean(c)
    # CTA-per-trajectory kernels: n = 4096 with recording (transposition buffer, chunked slots), ragged n, Burgers,
    # component right-hand side with an auxiliary output
    h = wl.c5_heat(N=5, n=4096, t_bound=3.0)
    d = ni.RK45(ni.rhs.heat1d(h.params[:, 0]), 0.0, h.y0, 3.0, rtol=1e-8, atol=1e-8, record=300)
    for _ in range(50):
        try:
            d.run()
            d.record_drain()
            break
        except ni.NiError as exc:
            assert exc.code == -6
            d.record_drain()
    d.terminal(), d.stiffness(5)
    clean(d)
    h = wl.c5_heat(N=7, n=1000, t_bound=3.0)
    e = ni.RK23(ni.rhs.burgers1d(np.full(7, 0.5), np.full(7, 3.0)), 0.0, h.y0, 0.2, rtol=1e-6, atol=1e-8, record=2000)
    e.run()
    clean(e)
    h = wl.c5_heat(N=3, n=333, t_bound=1.0)
    comp = ni.CudaComponentRHS('const double ym = j > 0 ? y[j - 1] : 0.0, yp = j + 1 < n ? y[j + 1] : 0.0;\n'
                               'return p[0] * ((ym - 2.0 * y[j]) + yp);', P=1,
                               aux_body='double s = 0.0;\nfor (int j = 0; j < n; ++j) s += y[j];\naux[0] = s;', Q=1)
    g = ni.Advanced(1, 1, ni.RK45)(comp, 0.0, h.y0, h.params, 1.0, rtol=1e-8, atol=1e-8, record=500)
    while ni.step(g):
        pass
    clean(g)


def test_guard_zones_catch_an_out_of_bounds_write(monkeypatch):
    """The check itself: one double written just before and one far behind a guarded buffer (through a raw device view)
    must be reported."""
    import torch
    from numba_integrators_b200 import _lib
    from numba_integrators_b200.distributed import _DeviceView
    monkeypatch.setenv('NI_B200_GUARD', '1')
    N = 640                                           # 640 doubles = 5120 bytes: the buffer ends on its 256-byte boundary
    e = ni.RK45(ni.rhs.harmonic, 0.0, np.ones((N, 2)), 1.0)
    e.run()
    clean(e)
    ptr = e.device_pointer(_lib.F_T)
    dev = torch.device('cuda', e.device)
    before = torch.as_tensor(_DeviceView(ptr - 8, (1,), '<f8'), device=dev)
    behind = torch.as_tensor(_DeviceView(ptr + 8 * N + 40_000, (1,), '<f8'), device=dev)
    before.fill_(1.0)
    behind.fill_(1.0)
    torch.cuda.synchronize()
    n_buffers, bad = e.guard_check()
    assert n_buffers >= 15 and 1 <= bad <= 16, bad    # (bytes of 1.0 that differ from the 0xA5 fill: 8 + 8)
