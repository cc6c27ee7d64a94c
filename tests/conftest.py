import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # the CUDA extension and the CPU oracle are build artefacts (git-ignored): compile them when absent
    from numba_integrators_b200 import build as _b
    if not _b.LIB.exists():
        _b.build()


def golden_names():
    return sorted(p.stem for p in GOLDEN.glob('*.npz'))


def load_golden(name):
    g = np.load(GOLDEN / f'{name}.npz')
    meta = json.loads(str(g['meta']))
    return meta, {k: g[k] for k in g.files if k != 'meta'}


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def has_gpu():
    try:
        import ctypes
        from numba_integrators_b200 import _lib
        c = ctypes.c_int(0)
        return _lib.lib().ni_device_count(ctypes.byref(c)) == 0 and c.value > 0
    except Exception:
        return False
