"""Static instruction, register-operand and register-port counts of the run kernels' hot loops, read off the SASS.

The thread-per-trajectory run kernel is bound by the register file's read ports (DESIGN.md section 3.2,
profiles/r02_port_model.txt).  This script prints, for one pass through the hot loop -- everything between the loop's
EXIT test and its back edge, minus the cold regions (forward-skipped spans that contain a CALL, a shared-memory store or
an FP64 compare: the operator slow paths, the cold block, the rare clip to t_bound) --
  D = FP64-pipe instructions, I = all others,
  R = 32-bit register source operands read (64-bit operands count twice, reuse-cache hits not at all),
  port cycles = sum over the instructions of max(1, distinct even-numbered sources, distinct odd-numbered sources):
      a scheduler's register file has two banks (register number mod 2), each delivers one 32-bit register per lane
      per cycle; a 64-bit operand is one read in each bank.  Measured kernel time per warp-pass per scheduler is
      1.00 - 1.06 x this number for C2, C3, C4 and four variants of C3 (it replaces round 1's `2 D + I` and the bank-blind
      `1.18 x R / 2`).
Every loop body of a kernel is listed (ni_small_run_body carries two: the lean one and the general one).

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


def register_reads(body):
    """32-bit register source operands of a list of (address, text) instructions; reuse-cache hits excluded."""
    total = 0
    prev_reuse = set()
    for _, t in body:
        tt = re.sub(r'^@!?U?P\d\s+', '', t)
        head = tt.split()[0]
        op = head.split('.')[0]
        parts = [x.strip() for x in tt.split(None, 1)[1].split(',')] if ' ' in tt else []
        if op in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'NOP', 'WARPSYNC', 'LDCU', 'LDC', 'S2R', 'CS2R', 'BREAK'):
            srcs = []
        elif op in ('STG', 'STL', 'STS', 'RED', 'REDG'):
            srcs = parts
        elif op in ('ISETP', 'DSETP', 'FSETP'):
            srcs = parts[2:]
        else:
            srcs = parts[1:]
        wide = op in FP64 or '.64' in head or op == 'DSETP'
        cur_reuse = set()
        for k, src in enumerate(srcs):
            m = re.match(r'[-|~!]*(R\d+)(\.reuse)?', src)
            if not m:
                continue
            w = 2 if wide else 1
            if (k, m.group(1)) not in prev_reuse:
                total += w
            if m.group(2):
                cur_reuse.add((k, m.group(1)))
        prev_reuse = cur_reuse
    return total


def port_cycles(body):
    """Register-port cycles of a list of (address, text) instructions with the banks taken into account: a scheduler's
    register file has an even and an odd bank, each delivers one 32-bit register per lane per cycle, so an instruction
    holds the read ports for max(1, distinct even sources, distinct odd sources) cycles (a 64-bit operand is one of
    each; reuse-cache hits are free).  An instruction without register sources still takes its issue cycle."""
    total = 0
    prev_reuse = set()
    for _, t in body:
        tt = re.sub(r'^@!?U?P\d\s+', '', t)
        head = tt.split()[0]
        op = head.split('.')[0]
        parts = [x.strip() for x in tt.split(None, 1)[1].split(',')] if ' ' in tt else []
        if op in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'NOP', 'WARPSYNC', 'LDCU', 'LDC', 'S2R', 'CS2R', 'BREAK'):
            srcs = []
        elif op in ('STG', 'STL', 'STS', 'RED', 'REDG'):
            srcs = parts
        elif op in ('ISETP', 'DSETP', 'FSETP'):
            srcs = parts[2:]
        else:
            srcs = parts[1:]
        wide = op in FP64 or '.64' in head
        banks = (set(), set())
        cur_reuse = set()
        for k, src in enumerate(srcs):
            m = re.match(r'[-|~!]*R(\d+)(\.reuse)?', src)
            if not m:
                continue
            r = int(m.group(1))
            if (k, r) not in prev_reuse:
                for x in ((r, r + 1) if wide else (r,)):
                    banks[x & 1].add(x)
            if m.group(2):
                cur_reuse.add((k, r))
        prev_reuse = cur_reuse
        total += max(1, len(banks[0]), len(banks[1]))
    return total


def hot_loops(obj, kernel):
    """[(histogram, n_fp64, n_other)] for every attempt loop of the kernel, in address order; the histogram carries
    the register-read count under the key '_reads'."""
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
                # cold: the operator slow paths and the general body's cold block (a CALL), the lean body's stash of
                # an attempt for the deferred cold block (shared-memory stores), the rare exact clip to t_bound (the
                # only FP64 compares of the loop)
                if any('CALL' in u or u.startswith('DSETP') or u.startswith('STS') for b, u in body if b in skipped):
                    cold.update(skipped)
        hot = [(a, t) for a, t in body if a not in cold]
        hist = collections.Counter(_opcode(t) for a, t in hot)
        d = sum(hist[k] for k in FP64)
        i = sum(hist.values()) - d
        hist['_reads'] = register_reads(hot)
        hist['_ports'] = port_cycles(hot)
        loops.append((hist, d, i))
    return loops


def main(argv):
    targets = {'kernel': (argv[0], argv[1])} if len(argv) >= 2 else DEFAULTS
    for label, (obj, kernel) in targets.items():
        path = Path(obj) if Path(obj).exists() else BUILD / obj
        for hist, d, i in hot_loops(path, kernel):
            local = hist['STL'] + hist['LDL']
            r = hist.pop('_reads')
            pc = hist.pop('_ports')
            print(f'{label:28s} D = {d:4d}  I = {i:4d}  R = {r:4d}  max(2D, R/2) = {max(2 * d, r / 2):6.1f}  '
                  f'port cycles (banked) = {pc:4d}  local-memory accesses {local}')
            print('    ' + ' '.join(f'{k}:{v}' for k, v in hist.most_common()))


if __name__ == '__main__':
    main(sys.argv[1:])
