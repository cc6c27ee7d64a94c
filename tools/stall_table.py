"""Per-SASS-line warp-stall table of a run kernel's hot loop from an `ncu --set full --import-source on` capture.

    python tools/stall_table.py gpurun_out/prof_c3.ncu-rep [out.txt] [--top 25]

Reads `ncu -i <rep> --page source --csv`, keeps the instructions executed (almost) once per loop pass, and prints
 * the totals per stall reason in scheduler cycles per warp per pass (samples / samples-per-issue: an instruction that
   issues once per pass collects one `selected` sample per sampling period it is issued in, so `selected` of the hot
   loop divided by its instruction count calibrates samples -> cycles),
 * the same per opcode class,
 * the top lines by stall samples, each with its three largest reasons.
`dispatch` stalls hold the scheduler's issue port (register-file port conflicts): unlike every other reason they are
cycles in which NO warp of the scheduler issues, so they add to the kernel time one for one.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def load(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    k = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    return rows[0][1] if rows[0][0] == 'Kernel Name' else '?', rows[k], rows[k + 1:]


def main(argv):
    top = 25
    if '--top' in argv:
        i = argv.index('--top')
        top = int(argv[i + 1])
        del argv[i:i + 2]
    rep = argv[0]
    out = open(argv[1], 'w') if len(argv) > 1 else sys.stdout
    kernel, hdr, data = load(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith('stall_') and '(' not in h]
    ex = [int(r[ix['Instructions Executed']]) for r in data]
    # the hot loop's instructions all execute once per pass: the most common (large) execution count
    from collections import Counter
    mx = Counter(e for e in ex if e > 0.05 * max(ex)).most_common(1)[0][0]
    hot = [r for r, e in zip(data, ex) if 0.97 * mx <= e <= 1.03 * mx]
    per_issue = sum(int(r[ix['stall_selected']]) for r in hot) / len(hot)
    tot_all = sum(int(r[ix['# Samples']]) for r in data)
    tot_hot = sum(int(r[ix['# Samples']]) for r in hot)
    p = lambda *a: print(*a, file=out)  # noqa: E731
    p(f'# {kernel}')
    p(f'# source: ncu -i {rep} --page source --csv   (tools/stall_table.py)')
    p(f'# hot loop = {len(hot)} SASS instructions executed once per pass ({mx} warp-passes); it holds '
      f'{100 * tot_hot / tot_all:.1f} % of all samples; {per_issue:.0f} samples = 1 cycle of one warp per pass')
    p('#\n# stall reason        cycles per warp per pass   share')
    tot = {h: sum(int(r[ix[h]]) for r in hot) for h in reasons}
    for h, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v:
            p(f'  {h[6:]:20s} {v / per_issue:10.1f}              {100 * v / tot_hot:5.1f} %')
    p(f'  {"total":20s} {tot_hot / per_issue:10.1f}')
    cls = defaultdict(lambda: [0, 0, defaultdict(int)])
    for r in hot:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
        c = cls[m.group(2).split('.')[0]]
        c[0] += 1
        c[1] += int(r[ix['# Samples']])
        for h in reasons:
            c[2][h] += int(r[ix[h]])
    p('#\n# opcode    count   cycles/pass  per instr   dispatch  wait   (cycles per warp per pass)')
    for op, c in sorted(cls.items(), key=lambda x: -x[1][1]):
        p(f'  {op:8s} {c[0]:5d}  {c[1] / per_issue:10.1f}  {c[1] / per_issue / c[0]:8.2f}  '
          f'{c[2]["stall_dispatch"] / per_issue:8.1f} {c[2]["stall_wait"] / per_issue:6.1f}')
    p(f'#\n# top {top} lines by stall samples (cycles per warp per pass; three largest reasons)')
    ranked = sorted(hot, key=lambda r: -int(r[ix['# Samples']]))[:top]
    for r in ranked:
        s = int(r[ix['# Samples']])
        big = sorted(((int(r[ix[h]]), h[6:]) for h in reasons), reverse=True)[:3]
        p(f'  {r[0][-5:]}  {s / per_issue:7.1f}  {"  ".join(f"{n} {v / per_issue:.1f}" for v, n in big if v):48s} '
          f'{r[ix["Source"]].strip()}')


if __name__ == '__main__':
    main(sys.argv[1:])
