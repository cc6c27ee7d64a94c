"""Summarise an .ncu-rep (raw page + per-opcode executed-instruction histogram) into a text file for profiles/."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def main(rep, out, note=''):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f'# {note}', f'# source: ncu -i {rep} --page raw --csv  (ncu --set full --clock-control none --import-source on)']
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            lines.append(f'{h:95s} {v:>24s} {u}')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, thr = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) < 10:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ix['Source']].strip())
        op = m.group(2) if m else '?'
        ops[op] += int(r[ix['Instructions Executed']])
        thr[op] += int(r[ix['Thread Instructions Executed']])
    tot = sum(ops.values())
    lines.append('')
    lines.append(f'# executed warp-instructions by opcode (SASS source page), total {tot}')
    for op, c in ops.most_common(24):
        lines.append(f'{op:12s} {c:14d}  {100 * c / tot:5.1f} %   avg active threads {thr[op] / c:5.1f}')
    fp64 = sum(ops[o] for o in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
    lines.append(f'# FP64-pipe instructions (DFMA+DMUL+DADD+DSETP): {fp64} = {100 * fp64 / tot:.1f} % of all')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], ' '.join(sys.argv[3:]))
