"""GPU-vs-reference parity table over every golden fixture (run on the B200 box).

    python tools/parity_report.py [--oracle-n 20000] [--divergent gpurun_out/divergent.json] > gpurun_out/parity.txt

For each fixture: fraction of trajectories with identical accepted AND rejected counts,
fraction with max|dy|/max|y| <= 1e-12, fraction bit-identical, worst relative deviation.
With --oracle-n it also runs bigger samples of C2/C3/C4 through the CPU oracle on the box.
--divergent lists every trajectory that misses the bar (counts differ or relative deviation > 1e-12) with both
terminal states (BASELINE.json north_star: "with any divergent ones listed").
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numba_integrators_b200 as ni  # noqa: E402
from numba_integrators_b200 import workloads as wl  # noqa: E402


def solve_gpu(rhs, method, advanced, t0, y0, params, t_bound, rtol, atol, max_step=np.inf, first_step=0.0):
    f = ni.BuiltinRHS(rhs)
    if advanced:
        ctor = ni.Advanced(f.P, f.Q, ni.RK45 if method == 'RK45' else ni.RK23)
        s = ctor(f, t0, y0, params if f.P else None, t_bound, max_step=max_step, rtol=rtol, atol=atol,
                 first_step=first_step)
    else:
        ctor = ni.RK45 if method == 'RK45' else ni.RK23
        s = ctor(f(*params.T) if f.P else f, t0, y0, t_bound, max_step=max_step, rtol=rtol, atol=atol,
                 first_step=first_step)
    s.run()
    return s


DIVERGENT = {}
MAX_LISTED = 2000  # per configuration (the count is always complete)


def compare(name, s, ref_t, ref_y, ref_acc, ref_rej):
    y, t = np.atleast_2d(s.y), np.atleast_1d(s.t)
    acc, rej = np.atleast_1d(s.n_accepted), np.atleast_1d(s.n_rejected)
    scale = np.maximum(np.max(np.abs(ref_y), axis=1), 1e-300)
    rel = np.max(np.abs(y - ref_y), axis=1) / scale
    same_cnt = (acc == ref_acc) & ((rej == ref_rej) | (ref_rej < 0))
    bit = np.all(y == ref_y, axis=1) & (t == ref_t)
    print(f'{name:34s} N={len(t):6d} counts_identical={same_cnt.mean():8.4%} rel<=1e-12={np.mean(rel <= 1e-12):8.4%} '
          f'bit_identical={bit.mean():8.4%} median_rel={np.median(rel):.1e} max_rel={rel.max():.1e} '
          f't_exact={np.mean(t == ref_t):.4f}')
    bad = np.flatnonzero(~same_cnt | ~(rel <= 1e-12))
    DIVERGENT[name] = {
        'trajectories': int(len(t)), 'divergent': int(len(bad)), 'counts_differ': int(np.sum(~same_cnt)),
        'listed': [{'index': int(i), 'rel': float(rel[i]),
                    'gpu': {'t': float(t[i]), 'y': [float(v) for v in y[i]], 'accepted': int(acc[i]), 'rejected': int(rej[i])},
                    'reference': {'t': float(ref_t[i]), 'y': [float(v) for v in ref_y[i]], 'accepted': int(ref_acc[i]),
                                  'rejected': int(ref_rej[i])}} for i in bad[:MAX_LISTED] if y.shape[1] <= 16]}
    return same_cnt, rel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--oracle-n', type=int, default=0)
    ap.add_argument('--divergent', default='', help='write the divergent trajectories (index + both states) to this JSON file')
    args = ap.parse_args()
    for p in sorted((ROOT / 'tests' / 'golden').glob('*.npz')):
        g = np.load(p)
        meta = json.loads(str(g['meta']))
        if 'n_steps' in meta and meta['n_steps'] < 10 ** 6:
            continue  # partial runs are covered by the lock-step tests
        if meta['rhs'] not in ni.rhs._BUILTIN_IDS or 'cond_index' in meta:
            continue  # user-snippet right-hand sides and ff2cond runs: tests/test_gpu_large_features.py
        try:
            s = solve_gpu(meta['rhs'], meta['method'], meta['advanced'], g['t0'], g['y0'], g['params'], g['t_bound'],
                          g['rtol'], g['atol'], g['max_step_v'] if 'max_step_v' in g.files else meta['max_step'],
                          g['first_step_v'] if 'first_step_v' in g.files else meta['first_step'])
        except ni.NiError as e:
            print(f'{meta["name"]:34s} SKIPPED: {e}')
            continue
        rej = g['n_rejected'] if 'n_rejected' in g.files else np.full(len(g['t']), -1)
        compare(meta['name'], s, g['t'], g['y'], g['n_accepted'], rej)
    if args.oracle_n:
        from oracle import oracle as orc
        for w in (wl.c2_harmonic(N=args.oracle_n), wl.c3_lorenz(N=args.oracle_n, t_bound=1.0),
                  wl.c3_lorenz(N=args.oracle_n, t_bound=5.0), wl.c4_vanderpol(N=args.oracle_n // 4, tol=1e-8),
                  wl.c4_vanderpol(N=args.oracle_n // 4, tol=1e-6)):
            m = orc.RK45 if w.method == 'RK45' else orc.RK23
            t0 = time.time()
            o = orc.run_ensemble(w.rhs, m, w.t0, w.y0, w.t_bound, params=w.params, rtol=w.rtol, atol=w.atol,
                                 advanced=w.advanced)
            dt = time.time() - t0
            s = solve_gpu(w.rhs, w.method, w.advanced, w.t0, w.y0, w.params, w.t_bound, w.rtol, w.atol)
            compare(f'{w.name}(t={w.t_bound:g},tol={w.rtol:g}) vs oracle', s, o['t'], o['y'], o['n_accepted'],
                    o['n_rejected'])
            print(f'    oracle: {o["n_accepted"].sum() / dt / 1e6:.2f} M accepted steps/s on {orc.max_threads()} threads')
    if args.divergent:
        Path(args.divergent).write_text(json.dumps(
            {'_note': 'trajectories whose accepted / rejected counts differ from the reference (golden fixtures) or the CPU '
                      'oracle (samples), or whose terminal state deviates by more than 1e-12 relative (max-norm); at most '
                      f'{MAX_LISTED} listed per configuration, counts complete; systems with n > 16 are counted, not listed',
             'configurations': DIVERGENT}, indent=1) + '\n')


if __name__ == '__main__':
    main()
