"""Static instruction mix of the innermost loop of every kernel in a `cuobjdump -sass` dump (stdin), one line per
kernel: used to check that tools/micro/fp64_shadow.cu really contains the instruction classes its table claims.

    cuobjdump -sass /tmp/fp64_shadow | python tools/micro/shadow_mix.py
"""
import collections
import re
import subprocess
import sys


def demangle(name):
    try:
        return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    name, ins = None, []
    kernels = []
    for line in sys.stdin:
        m = re.match(r'\s+Function : (\S+)', line)
        if m:
            if name:
                kernels.append((name, ins))
            name, ins = m.group(1), []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r'\s+', ' ', m.group(2)).strip()))
    if name:
        kernels.append((name, ins))
    for name, ins in kernels:
        best = None
        for addr, text in ins:
            m = re.search(r'BRA(?:\.\w+)* (?:!?U?P\d, )?0x([0-9a-f]+)$', text)
            if m and int(m.group(1), 16) < addr:
                span = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
                if best is None or len(span) > len(best):
                    best = span
        if not best:
            continue
        hist = collections.Counter(re.sub(r'^@!?U?P\d\s+', '', t).split()[0].split('.')[0] for t in best)
        d = sum(hist[k] for k in ('DFMA', 'DMUL', 'DADD'))
        print(f'{demangle(name):60s} loop {len(best):4d}: FP64 {d:3d} | ' +
              ' '.join(f'{k}:{v}' for k, v in hist.most_common() if k not in ('DFMA', 'DMUL', 'DADD')))


if __name__ == '__main__':
    main()
