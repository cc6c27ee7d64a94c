"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/ni_b200.h declares, and fails loudly (no CPU fallback) when no device is usable."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def declared_symbols():
    text = (ROOT / 'include' / 'ni_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ni_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from numba_integrators_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 28
    for name in names:
        assert hasattr(L, name), f'{name} declared in include/ni_b200.h but not exported'
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert L.ni_version() >= 100


def test_header_compiles_as_c(tmp_path):
    import subprocess
    src = tmp_path / 't.c'
    src.write_text('#include "ni_b200.h"\nint main(void){return NI_OK;}\n')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', str(ROOT / 'include'), '-c', str(src), '-o',
                    str(tmp_path / 't.o')], check=True)


@pytest.mark.skipif(has_gpu(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback_without_gpu():
    import numba_integrators_b200 as ni
    with pytest.raises(ni.NiError) as ei:
        ni.RK45(ni.rhs.harmonic, 0.0, np.array((0., 1.)), t_bound=1)
    assert ei.value.code == -2  # NI_ERR_CUDA


def test_argument_errors_do_not_need_a_gpu():
    from numba_integrators_b200 import _lib
    L = _lib.lib()
    out = ctypes.c_void_p()
    assert L.ni_module_builtin(999, 2, 1, 0, out) == _lib.NI_ERR_UNSUPPORTED
    assert b'no built-in kernel' in L.ni_last_error()
    assert L.ni_module_builtin(0, 2, 7, 0, out) == _lib.NI_ERR_INVALID
    assert L.ni_module_builtin(0, 2, 1, 0, None) == _lib.NI_ERR_INVALID
    m = ctypes.c_void_p()
    assert L.ni_module_builtin(1, 3, 1, 1, m) == 0  # lorenz63 / RK45 / advanced
    vals = [ctypes.c_int() for _ in range(6)]
    assert L.ni_module_info(m, *vals) == 0
    assert [v.value for v in vals] == [3, 3, 1, 1, 1, 6]
    assert L.ni_module_destroy(m) == 0
    assert L.ni_module_builtin(1, 3, 1, 3, m) == 0  # + NI_MODE_FAST
    assert L.ni_module_info(m, *vals) == 0 and vals[4].value == 3
    assert L.ni_module_destroy(m) == 0
    # the opt-in variants of the stencil kernels are specialised with NVRTC: without a GPU that ends in NI_ERR_CUDA
    # (or NI_ERR_NVRTC where libnvrtc is absent), never in a silent fallback
    assert L.ni_module_builtin(3, 64, 1, 2, m) in (0, _lib.NI_ERR_CUDA, _lib.NI_ERR_NVRTC)
    L.ni_module_destroy(m)
    assert L.ni_module_builtin(1, 3, 1, 8, m) == _lib.NI_ERR_INVALID
    assert L.ni_module_destroy(m) == 0


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under numba_integrators_b200/ may import or load it."""
    for p in (ROOT / 'numba_integrators_b200').rglob('*'):
        if p.suffix in ('.py', '.cu', '.cuh', '.h', '.cpp') and 'build' not in p.parts[-2:-1]:
            text = p.read_text()
            assert 'libni_oracle' not in text and 'from oracle' not in text and 'import oracle' not in text, p
