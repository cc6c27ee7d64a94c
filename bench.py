"""Headline benchmark: accepted trajectory-steps/s of the batched adaptive RK integrator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--traj M] [--impl reference]

One "step" integrates the whole ensemble shard of this GPU from t0 to t_bound: the init kernel
(FSAL seed + select_initial_step, inputs already resident in HBM) followed by ONE launch of the
persistent fast-forward kernel.  `value` is the whole-job aggregate (all ranks' accepted steps /
max-over-ranks device time); `e2e` repeats the measurement through the C ABI with pinned HOST
buffers (upload -> reset -> run -> download of terminal t, y, aux, counters) inside the timed
region.  Weak scaling: every rank integrates its own --traj trajectories (`value`).  With more than one GPU the
line also carries the one collective of the path -- the final NCCL gather of terminal records -- as
`final_gather_ms` / `value_incl_gather`, and a `strong` record: the SAME --traj trajectories IN TOTAL sharded over
the GPUs (the north-star configuration: 10 M Lorenz trajectories on 8 GPUs), gather included.

The CPU oracle (oracle/) is used only as the reported baseline (`cpu_baseline`, `--impl reference`)
-- never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'accepted_trajectory_steps_per_s'
UNIT = 'steps/s'
DEFAULT_TRAJ = {'c2': 1_000_000, 'c3': 10_000_000, 'c4': 4_000_000, 'c5': 16_384}


def make_workload(name, traj, shard, n_shards, t_bound):
    from numba_integrators_b200 import workloads as wl
    kw = {} if t_bound is None else {'t_bound': t_bound}
    if name == 'c2':
        return wl.c2_harmonic(N=traj, shard=shard, **kw)
    if name == 'c3':
        return wl.c3_lorenz(N=traj, shard=shard, **kw)
    if name == 'c4':
        return wl.c4_vanderpol(N=traj, shard=shard, n_shards=n_shards, **kw)
    if name == 'c5':
        return wl.c5_heat(N=traj, shard=shard, **kw)
    raise SystemExit(f'unknown workload {name}')


def workload_label(w, traj):
    return (f'{w.name}: {traj} trajectories/GPU, {w.method}, n={w.n}, t_bound={w.t_bound:g}, atol=rtol={w.rtol:g}'
            + (', via ni.Advanced (per-trajectory parameters, auxiliary output)' if w.advanced else ''))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_oracle_rate(w, seconds, threads=0):
    """Oracle port on the host cores over a bounded prefix sample of the workload; ~`seconds` of work."""
    from oracle import oracle as orc
    method = orc.RK45 if w.method == 'RK45' else orc.RK23
    threads = threads or orc.max_threads()
    n = min(w.N, max(threads * 8, 256 if w.n < 64 else threads))

    def run(k):
        t0 = time.perf_counter()
        o = orc.run_ensemble(w.rhs, method, w.t0, w.y0[:k], w.t_bound, params=None if w.params is None else w.params[:k],
                             rtol=w.rtol, atol=w.atol, advanced=w.advanced, nthreads=threads)
        return time.perf_counter() - t0, int(o['n_accepted'].sum())

    dt, acc = run(n)  # calibration (also warms caches): grow the sample until it runs for >= 0.5 s
    while dt < 0.5 and n < w.N:
        n = min(w.N, n * 4)
        dt, acc = run(n)
    k = int(min(w.N, max(n, n * seconds / max(dt, 1e-3))))
    if k != n:
        dt, acc = run(k)
    return acc / dt, k, threads, dt


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's algorithm on the host cores (oracle port; the reference itself is
    Python+numba and cannot travel to the GPU box).  Rank 0 only."""
    if rank != 0:
        return
    w = make_workload(args.workload, min(args.traj, 2_000_000 if args.workload != 'c5' else 512), 0, 1, args.t_bound)
    # size one step at ~args.cpu_seconds / steps of CPU work
    per_step = max(1.0, args.cpu_seconds / max(args.steps, 1))
    rate, k, threads, _ = cpu_oracle_rate(w, per_step)
    from oracle import oracle as orc
    method = orc.RK45 if w.method == 'RK45' else orc.RK23

    def one():
        t0 = time.perf_counter()
        o = orc.run_ensemble(w.rhs, method, w.t0, w.y0[:k], w.t_bound, params=None if w.params is None else w.params[:k],
                             rtol=w.rtol, atol=w.atol, advanced=w.advanced, nthreads=threads)
        return time.perf_counter() - t0, int(o['n_accepted'].sum())

    for _ in range(args.warmup):
        one()
    tot_t, tot_acc = 0.0, 0
    for _ in range(args.steps):
        dt, acc = one()
        tot_t += dt
        tot_acc += acc
    value = tot_acc / tot_t
    sample = f'first {k} trajectories of the workload per step, oracle C port, {threads} host threads'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_label(w, args.traj), 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(DEFAULT_TRAJ))
    ap.add_argument('--traj', type=int, default=0, help='trajectories per GPU (default: the BASELINE config size)')
    ap.add_argument('--t-bound', type=float, default=None)
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='CPU work budget of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--record', action='store_true', help='record every accepted step to the HBM log (always on for c2)')
    ap.add_argument('--no-fast', action='store_true', help='skip the extra fast-arithmetic launches (roofline.fast)')
    ap.add_argument('--arithmetic', default='strict', choices=['strict', 'fast'],
                    help="'strict' = the reference's unfused FP64 arithmetic (default, bit-faithful); 'fast' = opt-in mode")
    args = ap.parse_args()
    args.traj = args.traj or DEFAULT_TRAJ[args.workload]
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from numba_integrators_b200 import build as _build
    if not _build.LIB.exists() and rank == 0:
        _build.build()  # build artefacts are git-ignored; normally they travel with the snapshot
    import numba_integrators_b200 as ni
    from numba_integrators_b200 import _lib

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = make_workload(args.workload, args.traj, rank, world, args.t_bound)
    N, n = w.N, w.n
    f = ni.BuiltinRHS(w.rhs)
    record = args.workload == 'c2' or args.record
    rec_per_traj = {'c2': 1000, 'c3': 620, 'c4': 5200, 'c5': 160}[args.workload]  # accepted steps per trajectory, with headroom

    def build_ensemble(wk=None, arithmetic=None):
        wk = wk or w
        kw = dict(rtol=wk.rtol, atol=wk.atol, device=local, arithmetic=arithmetic or args.arithmetic)
        if wk.advanced:
            Solver = ni.Advanced(f.P, f.Q, ni.RK45 if wk.method == 'RK45' else ni.RK23)
            e = Solver(f, wk.t0, wk.y0, wk.params, wk.t_bound, **kw)
        else:
            ctor = ni.RK45 if wk.method == 'RK45' else ni.RK23
            e = ctor(f(*wk.params.T) if f.P else f, wk.t0, wk.y0, wk.t_bound, **kw)
        if record:
            e.record_enable(int(wk.N * rec_per_traj))  # the whole ensemble's accepted steps stay in HBM
        return e

    def timed_steps(e, steps, warmup):
        """`steps` x (init kernel + run kernel) on resident inputs: (total ms, mean run-kernel ms, accepted, attempts)."""
        for _ in range(warmup):
            e.reset()
            e.run_async()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        a0.record()
        for i in range(steps):
            e.reset()
            kv[i][0].record()
            e.run_async()
            kv[i][1].record()
        a1.record()
        barrier()
        acc_ = e.n_accepted.astype(np.int64)
        rej_ = e.n_rejected.astype(np.int64)
        st = e.status
        assert np.all(st == _lib.ST_FINISHED), f'{np.sum(st != _lib.ST_FINISHED)} trajectories did not finish'
        return (a0.elapsed_time(a1), float(np.mean([x.elapsed_time(y) for x, y in kv])), int(acc_.sum()),
                int(acc_.sum() + rej_.sum()))

    def reduce_max(x):
        tt = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def reduce_sum(*xs):
        tt = torch.tensor(xs, dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return [float(v) for v in tt.tolist()]

    ens = build_ensemble()
    stream = torch.cuda.current_stream()
    ens.set_stream(stream.cuda_stream)

    # measured FP64 peak (MEASURED_PEAKS.json has HBM and bf16 only)
    pk = _lib.C.c_double(0)
    _lib.check(_lib.lib().ni_fp64_peak(local, 2000, 3, pk))
    fp64_peak = pk.value
    peaks = {}
    try:
        peaks = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())
    except OSError:
        pass

    launches0 = ens.timing()['launches']
    sampler = ClockSampler(local)
    sampler.start()
    ms, run_ms, acc_rank, att_rank = timed_steps(ens, args.steps, args.warmup)
    clocks = sampler.stop()
    launches = ens.timing()['launches'] - launches0 - 2 * args.warmup
    ms_max = reduce_max(ms)

    def timed_ordered(e, acc_total):
        """The same steps with the work queue in longest-first order, the cost being the attempts of the run just
        timed (ni_ensemble_set_order): what a repeated sweep / an integration in legs can do against the tail."""
        e.order_by_cost()
        o_ms, o_run, o_acc, _ = timed_steps(e, 3, 1)
        e.set_order(None)
        o_ms_max = reduce_max(o_ms) / 3
        return {'ms_per_step': o_ms_max, 'run_kernel_ms_max_over_ranks': reduce_max(o_run),
                'value': acc_total / (o_ms_max * 1e-3),
                'note': 'launch order = decreasing attempts of the previous run of the same ensemble (longest first); '
                        'not part of `value`'}

    ordered = None
    if not record:
        ordered = timed_ordered(ens, reduce_sum(acc_rank)[0])
    acc_all, = reduce_sum(acc_rank)
    value = acc_all * args.steps / (ms_max * 1e-3)

    # ---- the one collective of the path: final NCCL gather of terminal records (t, y, aux, counters, status)
    def timed_gather(e, n_global, interleaved, peer=False):
        from numba_integrators_b200 import distributed as nd
        g = nd.PeerGather(e, n_global, interleaved) if peer else nd.TerminalGather(e, n_global, interleaved)
        if peer and g.hdl is None:
            return None, g.dtype.itemsize
        g.run()  # warm-up (NCCL channel setup, buffer registration)
        barrier()
        best = None
        for _ in range(3):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            g0.record()
            g.run()
            g1.record()
            barrier()
            tmax = reduce_max(g0.elapsed_time(g1))
            best = tmax if best is None else min(best, tmax)
        return best, g.dtype.itemsize

    def timed_fused(e, n_global, interleaved, acc_total):
        """Steps with the gather fused into the integration (distributed.FusedGather): the run kernel stores every
        finished trajectory's terminal record to all GPUs' receive buffers while it integrates; the step ends with the
        barrier that makes them visible.  -> whole step incl. gather, ms (max over ranks), or None without peer memory."""
        from numba_integrators_b200 import distributed as nd
        if e.kind:  # CTA-per-trajectory kernels have no in-kernel sink
            return None
        g = nd.FusedGather(e, n_global, interleaved)
        if g.hdl is None:
            g.close()
            return None
        for _ in range(max(1, args.warmup)):
            e.reset()
            g.begin()
            e.run_async()
            g.finish()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record()
        for _ in range(args.steps):
            e.reset()
            g.begin()
            e.run_async()
            g.finish()
        f1.record()
        barrier()
        ms_f = reduce_max(f0.elapsed_time(f1)) / args.steps
        st = e.status
        assert np.all(st == _lib.ST_FINISHED)
        rows = g.recv.view(world, -1)[rank][:e.N * g.dtype.itemsize].cpu().numpy().view(g.dtype)
        assert np.array_equal(rows['n_accepted'], e.n_accepted) and np.array_equal(rows['t'], e.t)  # own slot, spot check
        impl = g.impl
        g.close()
        return {'ms_per_step_incl_gather': ms_f, 'value_incl_gather': acc_total / (ms_f * 1e-3), 'impl': impl,
                'record_bytes': g.dtype.itemsize}

    gather_ms, strong, fused = None, None, None
    if world > 1:
        gather_nccl_ms, rec_bytes = timed_gather(ens, N * world, False)
        gather_peer_ms, _ = timed_gather(ens, N * world, False, peer=True)
        gather_ms = min(x for x in (gather_nccl_ms, gather_peer_ms) if x is not None)
        fused = timed_fused(ens, N * world, False, acc_all)
        # ---- strong scaling: the single-GPU problem (--traj trajectories IN TOTAL) sharded over the GPUs
        from numba_integrators_b200 import distributed as nd
        from numba_integrators_b200.workloads import Workload
        wg = make_workload(args.workload, args.traj, 0, 1, args.t_bound)
        inter = args.workload == 'c4'  # sorted parameter sweep: interleave so that every GPU gets the same cost
        idx = nd.shard_indices(wg.N, rank, world, inter)
        ws = Workload(wg.name, wg.rhs, wg.method, wg.advanced, wg.t0, wg.t_bound, np.ascontiguousarray(wg.y0[idx]),
                      None if wg.params is None else np.ascontiguousarray(wg.params[idx]), wg.rtol, wg.atol,
                      wg.flops_per_attempt, wg.record_bytes_per_step)
        del wg
        ens_s = build_ensemble(ws)
        ens_s.set_stream(stream.cuda_stream)
        s_ms, s_run, s_acc, s_att = timed_steps(ens_s, args.steps, args.warmup)
        s_ms_max = reduce_max(s_ms)
        s_run_max = reduce_max(s_run)
        s_acc_all, s_att_all = reduce_sum(s_acc, s_att)
        s_gather_nccl, _ = timed_gather(ens_s, args.traj, inter)
        s_gather_peer, _ = timed_gather(ens_s, args.traj, inter, peer=True)
        s_gather = min(x for x in (s_gather_nccl, s_gather_peer) if x is not None)
        step_ms = s_ms_max / args.steps
        strong = {'trajectories_total': args.traj, 'trajectories_per_gpu': len(idx), 'ms_per_step': step_ms,
                  'value': s_acc_all / (step_ms * 1e-3), 'final_gather_ms': s_gather,
                  'final_gather_nccl_ms': s_gather_nccl, 'final_gather_peer_store_ms': s_gather_peer,
                  'ms_incl_gather': step_ms + s_gather, 'value_incl_gather': s_acc_all / ((step_ms + s_gather) * 1e-3),
                  'run_kernel_ms_max_over_ranks': s_run_max,
                  'gather_GBps_per_rank': rec_bytes * args.traj * (world - 1) / world / (s_gather * 1e-3) / 1e9,
                  'efficiency_vs_weak_per_gpu_rate': (s_acc_all / (step_ms * 1e-3)) / value,
                  'efficiency_vs_weak_incl_gather': (s_acc_all / ((step_ms + s_gather) * 1e-3)) / value,
                  'note': 'the same trajectories a 1-GPU run integrates, sharded by index; efficiency = strong rate / '
                          '(weak whole-job rate, i.e. n_gpus x the per-GPU rate at full occupancy)'}
        strong['_fp64_flops'] = ws.flops_per_attempt * s_att_all  # -> frac below, once the peak is known
        s_ordered = timed_ordered(ens_s, s_acc_all)
        s_ordered['efficiency_vs_weak_per_gpu_rate'] = s_ordered['value'] / value
        strong['cost_ordered'] = s_ordered
        s_fused = timed_fused(ens_s, args.traj, inter, s_acc_all)
        if s_fused is not None:
            s_fused['efficiency_vs_weak_incl_gather'] = s_fused['value_incl_gather'] / value
            strong['fused_gather'] = s_fused
        del ens_s

    # ---- end to end through the C ABI with pinned host buffers
    e2e = None
    if not args.no_e2e:
        P, Q = ens.P, ens.Q
        y0_h = torch.from_numpy(w.y0).pin_memory()
        par_h = torch.from_numpy(w.params).pin_memory() if P else None
        # the step's result: ONE packed terminal record per trajectory (t, y, aux, n_accepted, n_rejected, status)
        rec_dt = ens.terminal_dtype()
        out = torch.empty(N * rec_dt.itemsize, dtype=torch.uint8).pin_memory()
        h2d = y0_h.numel() * 8 + (par_h.numel() * 8 if P else 0) + 16
        d2h = out.numel()

        # Two ensembles on two streams (double buffering): while one integrates, the other one's inputs are
        # uploaded and the previous results are downloaded.  Every step still uploads its own inputs and downloads
        # its own results inside the timed region.
        ens_b = build_ensemble()
        out_b = torch.empty_like(out).pin_memory()
        s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()
        ens.set_stream(s_a.cuda_stream)
        ens_b.set_stream(s_b.cuda_stream)
        lanes = [(ens, out), (ens_b, out_b)]

        def launch(i):
            e, _ = lanes[i % 2]
            e.upload(w.t0, y0_h, par_h, w.t_bound)
            e.reset()
            e.run_async()

        def collect(i):
            e, o = lanes[i % 2]
            e.fetch_terminal(o)  # pack kernel + one device-to-host copy + stream synchronise

        launch(0)
        launch(1)
        collect(0)
        collect(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_a)
        launch(0)
        for i in range(1, args.steps):
            launch(i)
            collect(i - 1)
        collect(args.steps - 1)
        s_a.wait_stream(s_b)
        e1.record(s_a)
        barrier()
        tms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tms2, op=dist.ReduceOp.MAX)
        # single-shot latency: ONE ensemble, nothing overlapped -- upload, init, run, packed download, synchronise
        barrier()
        t_ss = time.perf_counter()
        e0.record(s_a)
        launch(0)
        collect(0)
        e1.record(s_a)
        torch.cuda.synchronize()
        single_wall_ms = (time.perf_counter() - t_ss) * 1e3
        single_ms = e0.elapsed_time(e1)
        for o in (out, out_b):
            rec = o.numpy().view(rec_dt)
            acc_e2e = int(rec['n_accepted'].astype(np.int64).sum())
            assert acc_e2e == acc_rank and np.all(rec['status'] == _lib.ST_FINISHED), (acc_e2e, acc_rank)
        e2e = {'value': acc_all * args.steps / (float(tms2.item()) * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'ms_per_step': float(tms2.item()) / args.steps,
               'single_shot_ms': single_ms, 'single_shot_wall_ms': single_wall_ms,
               'path': 'ni_ensemble_upload(pinned) -> ni_ensemble_reset -> ni_ensemble_run_async -> '
                       'ni_ensemble_get_terminal (packed t, y, aux, counters, status: one copy), two ensembles '
                       'double-buffered on two CUDA streams'}

    # ---- the opt-in fast arithmetic on the same workload: one extra ensemble, three timed launches (rank 0's GPU)
    fast = None
    if args.arithmetic == 'strict' and not args.no_fast and not record:
        ens.set_stream(stream.cuda_stream)
        try:
            ens_f = build_ensemble(arithmetic='fast')
            ens_f.set_stream(stream.cuda_stream)
            f_ms, f_run, f_acc, f_att = timed_steps(ens_f, 3, 2)
            fast = {'value': f_acc * 3 / (f_ms * 1e-3), 'kernel_ms': f_run,
                    'achieved': w.flops_per_attempt * f_att / (f_run * 1e-3) / 1e12,
                    'note': "arithmetic='fast' (contracted multiply-adds, MUFU-seeded reciprocal and controller root): "
                            'meets the 1e-12 / identical-counts bar, not bit-faithful; this GPU only, 3 launches'}
            fast['frac'] = fast['achieved'] / fp64_peak if fp64_peak else None
            del ens_f
        except ni.NiError as exc:  # no fast variant for this kernel family
            fast = {'unavailable': str(exc)}

    if rank == 0:
        flops = w.flops_per_attempt * att_rank  # algorithmic flops of one launch
        achieved = flops / (run_ms * 1e-3) / 1e12
        kname = 'ni_k_large_run (CTA per trajectory)' if ens.kind else 'ni_k_run (thread per trajectory, persistent)'
        roofline = {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                    'frac': achieved / fp64_peak if fp64_peak else None, 'traffic': None,
                    'kernel': kname, 'kernel_ms': run_ms,
                    'flops_per_attempt': w.flops_per_attempt, 'attempts_per_launch': att_rank,
                    'peak_source': 'live DFMA-chain microbenchmark ni_fp64_peak, operand form DFMA R,R,R,UR (MEASURED_PEAKS.json '
                                   'has no FP64 entry; nominal 37.2 TFLOP/s; ncu of the same kernel: FP64 pipe 99.97 % busy, '
                                   'profiles/r02_fp64_peak_kernel_ncu.csv)',
                    'hbm_gbs_measured': peaks.get('hbm_gbs')}
        try:  # DRAM traffic of this kernel measured by ncu (profiles/), scaled to this launch's trajectory count
            key = args.workload + ('_fast' if args.arithmetic == 'fast' else '')
            tr = json.loads((ROOT / 'profiles' / 'r02_dram_traffic.json').read_text()).get(key)
            if tr:
                roofline['traffic'] = tr['bytes_per_trajectory'] * N
                roofline['traffic_source'] = tr['capture']
                if 'fp64_pipe_busy' in tr:  # ncu: fraction of cycles the FP64 pipe was busy (same capture)
                    roofline['fp64_pipe_busy_ncu'] = tr['fp64_pipe_busy']
                rows = 8 * n if (not ens.kind and n > 1) else 0  # row-major mirror of y written at retirement
                roofline['algorithmic_bytes'] = (N * (8 * (2 * n + ens.P + ens.Q + 5) + 13 + 8 * (2 * n + ens.Q + 3) + 13 + rows)
                                                 + (w.record_bytes_per_step * acc_rank if record else 0))
        except (OSError, ValueError):
            pass
        if record:  # C2: recording-heavy -- report the HBM side of the ridge as well
            rec_bytes = w.record_bytes_per_step * acc_rank
            hbm = peaks.get('hbm_gbs') or 6650.0
            roofline['hbm'] = {'bound': 'hbm', 'achieved': rec_bytes / (run_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                               'frac': rec_bytes / (run_ms * 1e-3) / 1e9 / hbm, 'bytes_per_accepted_step':
                               w.record_bytes_per_step, 'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)'
                               if peaks.get('hbm_gbs') else 'fallback 6650 GB/s'}
        if ens.kind and not record:  # CTA per trajectory (C5): the HBM side, compulsory traffic and what streaming would cost
            hbm = peaks.get('hbm_gbs') or 6650.0
            comp = N * (8 * (2 * n + ens.P + ens.Q + 5) + 13 + 8 * (2 * n + ens.Q + 3) + 13)
            stream = 264.0 * n * att_rank  # stage s reads nnz_s + 1 vectors and writes 1, y_new + error 8
            roofline['hbm'] = {'bound': 'hbm', 'achieved': comp / (run_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                               'frac': comp / (run_ms * 1e-3) / 1e9 / hbm,
                               'note': 'compulsory traffic only (state in, state out): every stage vector stays in registers / '
                                       'shared memory',
                               'streaming_design_gbs': stream / (run_ms * 1e-3) / 1e9,
                               'streaming_design_note': 'a kernel that streams the stage vectors through HBM moves 264 n bytes '
                                                        'per attempt; at this throughput it would need this bandwidth'}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': workload_label(w, args.traj), 'arithmetic': args.arithmetic, 'parallelism': f'{world} GPU(s), trajectories '
                           'sharded by index, no data-path collective',
                           'l2': f'inputs {N * (n + ens.P) * 8 / 1e6:.0f} MB + state per launch exceed the 126 MB L2',
                           'accepted_per_traj': acc_rank / N, 'rejected_per_traj': (att_rank - acc_rank) / N},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline}
        if fast is not None:
            roofline['fast'] = fast
        if ordered is not None:
            if fp64_peak:  # (rank 0's own launch: same attempts, the ordered launch's kernel time)
                ordered['frac'] = w.flops_per_attempt * att_rank / (ordered['run_kernel_ms_max_over_ranks'] * 1e-3) / 1e12 / fp64_peak
            line['cost_ordered'] = ordered
        if gather_ms is not None:
            # the one collective of the path, once per job after the last step: pack kernel + ONE NCCL
            # all_gather_into_tensor of equal-sized record slots (not part of `value`, which is the driver's weak-scaling
            # figure; `value_incl_gather` charges it to every step)
            step_ms = ms_max / args.steps
            line['final_gather_ms'] = gather_ms
            line['final_gather_nccl_ms'] = gather_nccl_ms  # pack kernel + ONE all_gather_into_tensor
            line['final_gather_peer_store_ms'] = gather_peer_ms  # ONE kernel: pack + stores to every GPU over NVLink
            line['ms_incl_gather'] = step_ms + gather_ms
            line['value_incl_gather'] = acc_all / ((step_ms + gather_ms) * 1e-3)
            line['gather_GBps_per_rank'] = rec_bytes * N * (world - 1) / (gather_ms * 1e-3) / 1e9
            if fused is not None:
                # the gather fused into the integration: the run kernel stores the terminal records to every GPU while
                # it integrates; the figure is a whole step INCLUDING the gather
                fused['efficiency_vs_value'] = fused['value_incl_gather'] / value
                line['fused_gather'] = fused
        if strong is not None:
            strong['frac'] = strong.pop('_fp64_flops') / world / (strong['run_kernel_ms_max_over_ranks'] * 1e-3) / 1e12 / fp64_peak
            line['strong'] = strong
        if world == 1 and not args.no_cpu_baseline:
            rate, k, threads, dt = cpu_oracle_rate(w, args.cpu_seconds)
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                    'note': 'C port of the reference algorithm (the reference itself is Python + numba and '
                                            'cannot travel to the GPU box; it measured 0.27-0.44 M steps/s/thread in '
                                            'BASELINE.md section 2, the port runs ~7 M steps/s/thread)',
                                    'sample': f'first {k} trajectories of the same workload, {dt:.1f} s, oracle C port '
                                              f'of the reference algorithm on {threads} host threads'}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
