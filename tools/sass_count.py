"""Static instruction counts of the run kernels' hot loops, read off the SASS with cuobjdump.

The thread-per-trajectory run kernel is bound by issue slots: one attempt costs about 2 D + I scheduler cycles, D =
FP64-pipe instructions, I = all others (profiles/r01_ablation_issue_model.txt).  This script
prints D and I of one pass through the hot loop -- everything between the loop's EXIT test and its back edge, minus the
cold regions (forward-skipped spans that contain a CALL: the operator slow paths) -- for every loop body of a kernel
(ni_small_run_body carries two: max_step = inf and the general one).

    python tools/sass_count.py [object [mangled kernel name]]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / 'numba_integrators_b200' / 'csrc' / 'build'
FP64 = ('DMUL', 'DFMA', 'DADD', 'DSETP')
DEFAULTS = {  # label -> (object, kernel): the kernels of the benchmark configurations
    'c3 lorenz63 RK45 Advanced': ('ni_builtin_lorenz63.o', '_Z8ni_k_runI13NiRhsLorenz63Li1ELi1ELb0ELb0EEv6NiArgs'),
    'c4 vanderpol RK23 basic': ('ni_builtin_vanderpol.o', '_Z8ni_k_runI14NiRhsVanDerPolLi0ELi0ELb0ELb0EEv6NiArgs'),
    'c1 harmonic RK45 basic': ('ni_builtin_harmonic.o', '_Z8ni_k_runI13NiRhsHarmonicLi1ELi0ELb0ELb0EEv6NiArgs'),
}


def disassemble(obj, kernel):
    out = subprocess.run(['cuobjdump', '-sass', '-fun', kernel, str(obj)], capture_output=True, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r'\s+', ' ', m.group(2)).strip()))
    return ins


def _opcode(text):
    return re.sub(r'^@!?U?P\d\s+', '', text).split()[0].split('.')[0]


def hot_loops(obj, kernel):
    """[(histogram, n_fp64, n_other)] for every attempt loop of the kernel, in address order."""
    ins = disassemble(obj, kernel)
    loops = []
    for addr, text in ins:
        m = re.match(r'BRA 0x([0-9a-f]+)$', text)
        if not m or int(m.group(1), 16) >= addr:
            continue
        span = [(a, t) for a, t in ins if int(m.group(1), 16) <= a <= addr]
        exits = [k for k, (_, t) in enumerate(span) if t.endswith('EXIT')]
        if len(span) < 150 or not exits:
            continue
        if any(re.match(r'BRA 0x([0-9a-f]+)$', t) and int(t.split('0x')[1], 16) < a and (a, t) != (addr, text)
               for a, t in span[exits[0]:]):
            continue  # an outer span that contains another loop's back edge
        body = span[exits[0] + 1:]
        cold = set()
        for a, t in body:
            mm = re.match(r'@!?P\d BRA (?:P\d, )?0x([0-9a-f]+)$', t)
            if mm:
                skipped = [b for b, u in body if a < b < int(mm.group(1), 16)]
                if any('CALL' in u for b, u in body if b in skipped):
                    cold.update(skipped)
        hist = collections.Counter(_opcode(t) for a, t in body if a not in cold)
        d = sum(hist[k] for k in FP64)
        loops.append((hist, d, sum(hist.values()) - d))
    return loops


def main(argv):
    targets = {'kernel': (argv[0], argv[1])} if len(argv) >= 2 else DEFAULTS
    for label, (obj, kernel) in targets.items():
        path = Path(obj) if Path(obj).exists() else BUILD / obj
        for hist, d, i in hot_loops(path, kernel):
            local = hist['STL'] + hist['LDL']
            print(f'{label:28s} D = {d:4d}  I = {i:4d}  2D+I = {2 * d + i:4d}  local-memory accesses {local}')
            print('    ' + ' '.join(f'{k}:{v}' for k, v in hist.most_common()))


if __name__ == '__main__':
    main(sys.argv[1:])
