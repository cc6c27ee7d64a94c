"""Instruction and register-port budget of the run kernels' hot loops (static SASS counts, tools/sass_count.py).

The thread-per-trajectory run kernel is bound by the register file's read ports (a scheduler's
register file has an even and an odd bank, each delivers one 32-bit register per lane per cycle; an attempt costs
1.03 - 1.06 x the sum over its instructions of max(1, distinct even sources, distinct odd sources) -- measured on C2,
C3 and C4), so an accidental extra select, a value pushed through local memory or a sign mask that nvcc turns back into
an FP64-pipe |x| is a measurable regression.  This test pins D (FP64-pipe instructions), I (all others), R (32-bit
register source operands) and the banked port cycles of the benchmark kernels a little above what the committed
sources compile to with nvcc 12.9: the first loop of a kernel is the lean body (fast-forward launches), the second the
general one (lock-step, budgets, finite max_step)."""
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tools'))

# label -> ((D, I, R, port cycles) budget of the lean body, the same of the general body)
BUDGET = {
    'c3 lorenz63 RK45 Advanced': ((243, 90, 975, 538), (244, 168, 1085, 642)),
    'c4 vanderpol RK23 basic': ((99, 74, 436, 256), (101, 142, 535, 348)),
    'c1 harmonic RK45 basic': ((140, 82, 584, 340), (142, 150, 682, 432)),
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
        d_max, i_max, r_max, p_max = BUDGET[label][k]
        assert hist['STL'] + hist['LDL'] == 0, (label, 'local memory in the hot loop', dict(hist))
        # FP64-pipe compares: the lean body spends one per component on max(|y|, |y_new|) and one on the MIN_FACTOR clamp
        # (cheaper in register-port cycles than the two-word integer compares, and the pipe has the slack); the
        # general body has none
        n_state = {'c3 lorenz63 RK45 Advanced': 3, 'c4 vanderpol RK23 basic': 2, 'c1 harmonic RK45 basic': 2}[label]
        assert hist['DSETP'] <= (n_state + 1 if k == 0 else 0), (label, 'FP64-pipe compares in the hot loop', dict(hist))
        assert d <= d_max and i <= i_max and hist['_reads'] <= r_max and hist['_ports'] <= p_max, (label, k, d, i, dict(hist))
