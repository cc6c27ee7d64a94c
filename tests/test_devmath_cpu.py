"""Controller math of csrc/ni_math.cuh compiled for the host: ni_pow_ctrl must be the correctly rounded
x**error_exponent (glibc pow, which the reference calls, is itself off by an ulp in ~5e-4 of calls), ni_div3 the
IEEE quotient, and the fast-mode helpers accurate to a few ulp."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / 'tools'))


def test_controller_math_is_correctly_rounded():
    import check_devmath
    res = check_devmath.main(4000)
    for name in ('RK45 e=-0.2', 'RK23 e=RN(-1/3)'):
        bad_mine, bad_libm, differ, n = res[name]
        assert bad_mine == 0, (name, bad_mine)
        assert res[name + ' ms'] == 0, (name, res[name + ' ms'])
        assert res[name + ' ms seed'] == 0, (name, res[name + ' ms seed'])  # seed error of the MUFU unit emulated
        assert differ == bad_libm and differ <= 0.002 * n       # the only disagreements are glibc's misroundings
    assert res['div3'] == 0
    assert max(res['fast']) < 5e-15
