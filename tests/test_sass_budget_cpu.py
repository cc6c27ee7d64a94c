"""Instruction budget of the run kernels' hot loops (static SASS counts, tools/sass_count.py).

The thread-per-trajectory run kernel is bound by issue slots (one attempt costs about 2 D + I
scheduler cycles, D = FP64-pipe instructions, I = the others), so an accidental extra select, a value pushed through
local memory or a sign mask that nvcc turns back into an FP64-pipe |x| is a measurable regression.  This test pins the
counts of the benchmark kernels a little above what the committed sources compile to with nvcc 12.9."""
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tools'))

# label -> (D budget, I budget) for the max_step = inf body (the first loop of the kernel)
BUDGET = {
    'c3 lorenz63 RK45 Advanced': (250, 165),
    'c4 vanderpol RK23 basic': (105, 140),
    'c1 harmonic RK45 basic': (146, 150),
}


@pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='cuobjdump not on PATH')
@pytest.mark.parametrize('label', sorted(BUDGET))
def test_hot_loop_instruction_budget(label):
    import sass_count
    from numba_integrators_b200 import build
    build.build()
    obj, kernel = sass_count.DEFAULTS[label]
    loops = sass_count.hot_loops(sass_count.BUILD / obj, kernel)
    assert len(loops) == 2, 'expected the max_step = inf body and the general body'
    d_max, i_max = BUDGET[label]
    for k, (hist, d, i) in enumerate(loops):
        assert hist['STL'] + hist['LDL'] == 0, (label, 'local memory in the hot loop', dict(hist))
        assert hist['DSETP'] == 0, (label, 'FP64-pipe compare in the hot loop', dict(hist))
        assert d <= d_max + (1 if k else 0) and i <= i_max + (6 if k else 0), (label, k, d, i, dict(hist))
    # the general body only adds the max_step clamp
    assert loops[1][1] - loops[0][1] <= 1 and loops[1][2] - loops[0][2] <= 6
