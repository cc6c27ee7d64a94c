"""Host-side logic of the facade (no GPU needed): input conversion, factories, workloads."""
import numpy as np
import pytest

import numba_integrators_b200 as ni
from numba_integrators_b200 import _solver, workloads as wl


def np_array_tuple_compare(t1, t2):
    return all(np.all(a1 == a2) and a1.dtype == a2.dtype for a1, a2 in zip(t1, t2))


class Test_convert:  # mirrors tests/unittests/test_auxiliaries.py:25-62 of the reference
    y0 = np.array((1, 2, 3), dtype=np.float64)
    rtol = np.array((1, 1, 1), dtype=np.float64)
    atol = np.array((2, 2, 2), dtype=np.float64)

    def test_correct(self):
        assert np_array_tuple_compare(ni.convert(self.y0, self.rtol, self.atol), (self.y0, self.rtol, self.atol))

    def test_tol_fill(self):
        assert np_array_tuple_compare(ni.convert(self.y0, 1., 2.), (self.y0, self.rtol, self.atol))

    def test_y0_from_list(self):
        assert np_array_tuple_compare(ni.convert(list(self.y0), self.rtol, self.atol), (self.y0, self.rtol, self.atol))

    def test_single_value_y0(self):
        assert np_array_tuple_compare(ni.convert(1, 2, 3), (np.array((1.,)), np.array((2.,)), np.array((3.,))))

    def test_dtype_change(self):
        assert np_array_tuple_compare(ni.convert(self.y0.astype(np.int32), self.rtol.astype(np.int32),
                                                 self.atol.astype(np.int32)), (self.y0, self.rtol, self.atol))

    def test_ensemble_y0_and_too_many_dims(self):
        y0, rtol, atol = ni.convert(np.ones((5, 3)), 1e-3, 1e-6)
        assert y0.shape == (5, 3) and rtol.shape == (3,) and atol.shape == (3,)
        with pytest.raises(ValueError):  # _aux.py:125-126 raises for ndim > 1; the ensemble axis moves that to > 2
            ni.convert(np.ones((2, 2, 2)), 1e-3, 1e-6)


def test_solvers_enum_quirk():
    """_basic.py:336-339: RK23/RK45 are plain functions, only ALL is an Enum member."""
    assert ni.Solvers.RK45 is ni.RK45 and ni.Solvers.RK23 is ni.RK23
    assert list(ni.Solvers) == [ni.Solvers.ALL]
    assert ni.ALL == (ni.RK23, ni.RK45)


def test_advanced_factory_shapes():
    both = ni.Advanced(3, 1, ni.Solvers.ALL)  # test_public.py:75-77
    assert isinstance(both, dict) and set(both) == {ni.Solvers.RK23, ni.Solvers.RK45}
    one = ni.Advanced(3, 1, ni.RK45)
    assert callable(one) and not isinstance(one, dict)
    with pytest.raises(ValueError):
        ni.Advanced(3, 1, 'RK99')
    with pytest.raises(ValueError):  # signature length check happens before any GPU work
        ni.Advanced(2, 1, ni.RK45)(ni.rhs.lorenz63, 0.0, np.ones(3), np.ones(3), 1.0)


def test_rhs_validation():
    with pytest.raises(TypeError):
        ni.RK45(42, 0.0, np.ones(2), 1.0)                       # not an RHS at all
    with pytest.raises(ni.TranslationError):                    # Python RHS outside the translatable subset
        ni.RK45(lambda t, y: np.linalg.inv(y), 0.0, np.ones(2), 1.0)
    with pytest.raises(ValueError):
        ni.RK45(ni.rhs.harmonic, 0.0, np.ones(3), 1.0)          # wrong dimension
    with pytest.raises(ValueError):
        ni.RK45(ni.rhs.vanderpol, 0.0, np.ones(2), 1.0)         # parameters missing
    with pytest.raises(ValueError):
        ni.RK45(ni.rhs.harmonic, 1.0, np.ones(2), 0.0)          # basic backward never terminates
    bound = ni.rhs.vanderpol(np.linspace(0.1, 2, 5))
    assert bound.bound_parameters.shape == (5, 1)
    assert ni.rhs.lorenz63(10.0, 28.0, 8 / 3).bound_parameters.shape == (3,)


def test_tableaux_and_constants():
    assert (ni.SAFETY, ni.MIN_FACTOR, ni.MAX_FACTOR) == (0.9, 0.2, 10)
    for A, B, C, E in ((_solver.RK23_A, _solver.RK23_B, _solver.RK23_C, _solver.RK23_E),
                       (_solver.RK45_A, _solver.RK45_B, _solver.RK45_C, _solver.RK45_E)):
        assert np.allclose(A.sum(axis=1), C, atol=1e-15)      # row-sum condition
        assert abs(B.sum() - 1) < 1e-15 and abs(E.sum()) < 1e-15


def test_step_mask_truthiness():
    m = _solver._mask([0, 0, 1])
    assert bool(m) and not bool(_solver._mask([0, 0]))


def test_workload_prefix_property_and_shards():
    a, b = wl.c3_lorenz(N=100), wl.c3_lorenz(N=1000)
    assert np.array_equal(a.y0, b.y0[:100]) and np.array_equal(a.params, b.params[:100])
    s0, s1 = wl.c3_lorenz(N=50, shard=0), wl.c3_lorenz(N=50, shard=1)
    assert not np.array_equal(s0.y0, s1.y0)
    v0, v1 = wl.c4_vanderpol(N=4, shard=0, n_shards=2), wl.c4_vanderpol(N=4, shard=1, n_shards=2)
    mu = np.sort(np.concatenate([v0.params[:, 0], v1.params[:, 0]]))
    assert np.allclose(mu, 0.1 + 19.9 * np.arange(8) / 7)     # interleaved shards cover the sweep
    h = wl.c5_heat(N=2, n=64)
    assert h.y0.shape == (2, 64) and h.params.shape == (2, 1)


def test_host_buffer_validation():
    """upload() / fetch() hand raw addresses to asynchronous copies: every buffer is checked (dtype, contiguity, size)
    and converted inputs are returned so that the caller can keep them alive (a temporary's address
    used to reach cudaMemcpyAsync)."""
    hb = _solver._host_buffer
    a = np.arange(12, dtype=np.float64).reshape(4, 3)
    assert hb(a, np.float64, 12, 'y0') is a  # already right: no copy
    for bad in (a.astype(np.float32), np.asfortranarray(a), a[:, ::-1], a.tolist()):
        b = hb(bad, np.float64, 12, 'y0')
        assert b.dtype == np.float64 and b.flags.c_contiguous and np.array_equal(b.reshape(4, 3), np.asarray(bad))
    with pytest.raises(ValueError, match='expected 12 values'):
        hb(a[:3], np.float64, 12, 'y0')
    out = np.empty(4, dtype=np.int32)
    assert hb(out, np.int32, 4, 'fetch', writable=True) is out
    for bad in (np.empty(4, dtype=np.int64), np.empty(5, dtype=np.int32), np.empty(8, dtype=np.int32)[::2], [0, 0, 0, 0]):
        with pytest.raises(ValueError, match='writable C-contiguous'):
            hb(bad, np.int32, 4, 'fetch', writable=True)
    torch = pytest.importorskip('torch')
    t = torch.zeros(4, 3, dtype=torch.float64)
    assert hb(t, np.float64, 12, 'y0') is t
    for bad in (t.float(), t.t(), t[:3]):
        with pytest.raises(ValueError, match='contiguous CPU tensor'):
            hb(bad, np.float64, 12, 'y0')
