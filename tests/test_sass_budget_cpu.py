"""Instruction and register-operand budget of the run kernels' hot loops (static SASS counts, tools/sass_count.py).

The thread-per-trajectory run kernel is bound by register-file read bandwidth and the FP64 pipe (DESIGN.md section 3.2:
one attempt costs about max(2 D, R / 2) / 0.85 scheduler cycles, D = FP64-pipe instructions, R = 32-bit register source
operands), so an accidental extra select, a value pushed through local memory or a sign mask that nvcc turns back into
an FP64-pipe |x| is a measurable regression.  This test pins the counts of the benchmark kernels a little above what
the committed sources compile to with nvcc 12.9: the first loop of a kernel is the lean body (fast-forward launches),
the second the general one (lock-step, budgets, finite max_step)."""
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tools'))

# label -> ((D, I, R) budget of the lean body, (D, I, R) budget of the general body)
BUDGET = {
    'c3 lorenz63 RK45 Advanced': ((246, 132, 1040), (250, 172, 1110)),
    'c4 vanderpol RK23 basic': ((101, 102, 485), (105, 144, 560)),
    'c1 harmonic RK45 basic': ((142, 110, 625), (146, 154, 705)),
}


@pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='cuobjdump not on PATH')
@pytest.mark.parametrize('label', sorted(BUDGET))
def test_hot_loop_instruction_budget(label):
    import sass_count
    from numba_integrators_b200 import build
    build.build()
    obj, kernel = sass_count.DEFAULTS[label]
    loops = sass_count.hot_loops(sass_count.BUILD / obj, kernel)
    assert len(loops) == 2, 'expected the lean body and the general body'
    for k, (hist, d, i) in enumerate(loops):
        d_max, i_max, r_max = BUDGET[label][k]
        assert hist['STL'] + hist['LDL'] == 0, (label, 'local memory in the hot loop', dict(hist))
        assert hist['DSETP'] == 0, (label, 'FP64-pipe compare in the hot loop', dict(hist))
        assert d <= d_max and i <= i_max and hist['_reads'] <= r_max, (label, k, d, i, dict(hist))
