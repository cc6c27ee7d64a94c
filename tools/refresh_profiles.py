"""Rebuild the tracked profile artefacts under profiles/ from the ncu captures that tools/gpu_prof_final.sh (round 1) /
tools/gpu_r2_prof.sh (round 2) left in gpurun_out/: one text summary per run kernel (tools/ncu_summary.py), the launch
list of the default bench, and profiles/rNN_dram_traffic.json (DRAM bytes per trajectory and FP64-pipe utilisation,
which bench.py scales to its own launch for roofline.traffic).

    python tools/refresh_profiles.py [r01|r02]
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tools'))
import ncu_summary  # noqa: E402

OUT, PROF = ROOT / 'gpurun_out', ROOT / 'profiles'
ROUND = sys.argv[1] if len(sys.argv) > 1 else 'r02'
PRE = 'prof_' if ROUND == 'r01' else 'prof_r2_'
LAUNCHES = 'launches_c3.csv' if ROUND == 'r01' else 'r02_launches_c3.csv'
CAPTURES = {  # key -> (report, trajectories of the captured launch, note)
    'c3': ('prof_c3', 2_000_000, 'C3 Lorenz-63 RK45 via Advanced, strict arithmetic, 2M trajectories (bench.py --traj 2000000)'),
    'c3_fast': ('prof_c3fast', 2_000_000, 'C3 Lorenz-63 RK45, arithmetic=fast, 2M trajectories'),
    'c2': ('prof_c2', 300_000, 'C2 harmonic RK45 with step recording (REC kernel), 300k trajectories'),
    'c4': ('prof_c4', 400_000, 'C4 Van der Pol RK23 mu sweep, 400k trajectories'),
    'c5': ('prof_c5', 2_048, 'C5 heat n=4096 RK45, CTA per trajectory, 2048 trajectories'),
    'c5_rec': ('prof_c5rec', 1_184, 'C5 heat n=4096 RK45 with every accepted step recorded (32 776 B each), 1184 trajectories'),
}


def raw(rep):
    txt = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def to_bytes(unit, val):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]
    return float(val.replace(',', '')) * scale


def main():
    traffic = {'_note': f'dram__bytes_read.sum + dram__bytes_write.sum of the run kernel from the {ROUND} ncu --set full '
                        f'captures (profiles/{ROUND}_*_run_kernel_ncu_full_summary.txt), divided by the trajectories of that '
                        'launch; bench.py scales it to its own launch size for roofline.traffic; fp64_pipe_busy = '
                        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active of the same capture (the north-star '
                        'evidence: DFMA-pipe utilisation)'}
    for key, (name, ntraj, note) in CAPTURES.items():
        rep = OUT / f"{name.replace('prof_', PRE)}.ncu-rep"
        if not rep.exists():
            print('missing', rep)
            continue
        out = PROF / f"{ROUND}_{key.replace('_', '')}_run_kernel_ncu_full_summary.txt"
        ncu_summary.main(str(rep), str(out), f'{ROUND} final kernels: {note}; --steps 1 --warmup 1 --no-e2e --no-cpu-baseline')
        m = raw(rep)
        rd, wr = to_bytes(*m['dram__bytes_read.sum']), to_bytes(*m['dram__bytes_write.sum'])
        traffic[key] = {
            'bytes_per_trajectory': round((rd + wr) / ntraj, 1),
            'capture': f'profiles/{out.name} ({ntraj:,} trajectories: {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written)',
            'fp64_pipe_busy': round(float(m['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'][1]) / 100, 3),
        }
    (PROF / f'{ROUND}_dram_traffic.json').write_text(json.dumps(traffic, indent=1) + '\n')
    launches = OUT / LAUNCHES
    if launches.exists():
        body = [l for l in launches.read_text().splitlines() if not l.startswith('==')]
        head = [f'# {ROUND} (final kernels): ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv',
                '# command: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e [--no-fast]   (C3, 10 M Lorenz trajectories)',
                '# cold-cache serialised per-launch times: compare shares, not absolutes']
        (PROF / f'{ROUND}_c3_bench_launch_list.csv').write_text('\n'.join(head + body) + '\n')
    print(json.dumps(traffic, indent=1))


if __name__ == '__main__':
    main()
