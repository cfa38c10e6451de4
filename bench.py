#!/usr/bin/env python
"""Benchmark of the extremal-field hot path (BASELINE.json: extremal-RK-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload era5|planar|gyre] [--impl reference]

One step = one full `SolverEFResampling` solve (setup, integration of every extremal over the whole horizon with
events, resampling scan, trim_distance, arrival-time argmin) of the named workload on synthetic data:

  era5   (default) BASELINE config 4: ERA5-shaped spherical grid 1440 x 721 x 24, 1.25 M root extremals PER GPU
         (10 M over 8 GPUs), max_depth = 1, DOPRI5 restated from scipy, fp64 grid + fp64 state.  Weak scaling.
  planar BASELINE config 3: planar unsteady DiscreteFF 1024 x 1024 x 48, 2^19 roots per GPU, max_depth = 4.
  gyre   BASELINE config 2 flow: analytic Gyre, 100 k roots per GPU, max_depth = 1 (fp64 pipe bound, no gather).

`value`  : whole-job RK steps / device time, flow grid resident in HBM, roots generated on the device.
`e2e`    : the same solve through the public API from (page-locked) HOST arrays, everything inside the timed region: a
           new solver per step, grid upload + repack (streamed in four groups of frames behind the free arcs), root
           costates upload, solve, then the optimum, the slots of the solution sites and the optimal trajectory read back.
`--impl reference`: the CPU oracle (oracle/solver_port.py: numpy + scipy solve_ivp like the reference) on a bounded
           sample of the same workload, all host cores (one process per core, independent theta_0 shards).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# RHS per step x 8 nodes x 6 scalars x bytes per scalar + 32 B of trajectory output
BYTES_PER_STEP = {('dopri5', 'f64'): 2336, ('dopri5', 'f32'): 1184, ('rk4', 'f64'): 1568, ('rk4', 'f32'): 800}
BYTES_PER_STEP_DOPRI5_F64 = 2336
WORKLOADS = ('era5', 'planar', 'gyre', 'regional')


def build_problem(workload, small=False):
    """(problem, total_duration, solver kwargs, description) - identical on every rank."""
    from dabry_b200 import synthetic
    from dabry_b200.flowfield import DiscreteFF
    from dabry_b200.misc import Coords
    from dabry_b200.problem import NavigationProblem
    if workload == 'era5':
        shape = (6, 180, 91) if small else (24, 1440, 721)
        sc = synthetic.era5_like(*shape)
        ff = DiscreteFF(sc.values, sc.bounds, Coords.GCS, force_no_diff=True)
        pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, name='era5_like').rescale()
        # fixed horizon: 1.1 x the great-circle time at srf, in the rescaled clock (skips the host heuristics)
        from dabry_b200.misc import Utils
        t_gc = Utils.distance(sc.x_init, sc.x_target, Coords.GCS) / sc.srf
        total = 1.1 * t_gc / (1. / (sc.srf / Utils.EARTH_RADIUS))
        return pb, total, dict(max_depth=1), 'era5_like %dx%dx%d gcs' % shape
    if workload == 'planar':
        shape = (6, 128, 128) if small else (48, 1024, 1024)
        sc = synthetic.planar_waves(*shape)
        ff = DiscreteFF(sc.values, sc.bounds, Coords.CARTESIAN, force_no_diff=True)
        pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, name='planar_waves').rescale()
        # max_dist: 2^19 roots refine to ~1.0 M extremals at full size (SURVEY 8d config 3 asks for 0.9 - 1.1 M;
        # profiles/tune_planar_max_dist.py: 5e-5 -> 918 k, 3e-5 -> 1.077 M)
        return pb, sc.total_duration, dict(max_depth=4, max_dist=0.002 if small else 4e-5), 'planar_waves %dx%dx%d' % shape
    if workload == 'gyre':
        pb = NavigationProblem.from_name('gyre').rescale()
        return pb, 1.0, dict(max_depth=1), 'gyre analytic'
    if workload == 'regional':
        # BASELINE config 5: what NavigationProblem.from_database builds (0.5 deg box, 3-hourly frames over 48 h) plus
        # one circular obstacle and one circular penalty; SolverEFTrimming
        from dabry_b200.obstacle import CircleObs
        from dabry_b200.penalty import CirclePenalty
        from dabry_b200.misc import Utils
        shape = (5, 41, 23) if small else (17, 181, 101)
        sc = synthetic.regional_constrained(*shape)
        ff = DiscreteFF(sc.values, sc.bounds, Coords.GCS, force_no_diff=True)
        obstacles = [CircleObs(np.array(c), float(r)) for c, r in sc.circle_obstacles]
        c, r, a = sc.circle_penalty
        pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, obstacles=obstacles,
                               penalty=CirclePenalty(np.array(c), float(r), float(a)), name='regional').rescale()
        t_gc = Utils.distance(sc.x_init, sc.x_target, Coords.GCS) / sc.srf
        total = 1.1 * t_gc * (sc.srf / Utils.EARTH_RADIUS)
        return pb, total, dict(max_depth=3, trimming=True), 'regional_constrained %dx%dx%d gcs + obstacle + penalty' % shape
    raise ValueError(workload)


def default_roots(workload):
    return {'era5': 1_250_000, 'planar': 1 << 19, 'gyre': 100_000, 'regional': 500_000}[workload]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(',')]
                if len(parts) >= 9:
                    self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        mx = [float(s[2]) for s in self.samples if s[2].replace('.', '').isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[5:9]):
                if v.lower() == 'active':
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---- CPU oracle legs ---------------------------------------------------------------------------------------
_ORACLE_PROBLEM = {}


def _oracle_problem(workload, small):
    """Per-process cache: the problem (grid + host derivatives) is built once per worker."""
    key = (workload, small)
    if key not in _ORACLE_PROBLEM:
        import warnings
        warnings.filterwarnings('ignore')
        pb, total, kw, _ = build_problem(workload, small)
        inner = getattr(pb.model.ff, 'ff', None)
        if inner is not None and getattr(inner, 'grad_values', 1) is None:
            inner.compute_derivatives()  # the GPU arm differentiates on the device; the CPU oracle needs the host copy
        _ORACLE_PROBLEM[key] = (pb, total)
    return _ORACLE_PROBLEM[key]


def _oracle_warm(args):
    _oracle_problem(*args)
    return 0


def _oracle_shard(args):
    workload, small, n_roots, lo, count = args
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from solver_port import Port
    pb, total = _oracle_problem(workload, small)
    port = Port(pb, total, n_costate_sectors=n_roots, root_range=(lo, count), max_depth=1)
    t0 = time.perf_counter()
    port.run_resampling()
    return port.n_rk, time.perf_counter() - t0, port.solution_cost


class OraclePool:
    """`cores` worker processes, each holding its own copy of the problem; a call times `sample_roots` extremals of
    the workload's fan split into independent theta_0 shards, one per core."""

    def __init__(self, workload, small, cores):
        from multiprocessing import get_context
        self.workload, self.small, self.cores = workload, small, cores
        self.pool = get_context('spawn').Pool(cores) if cores > 1 else None
        if self.pool:
            self.pool.map(_oracle_warm, [(workload, small)] * cores, chunksize=1)
        else:
            _oracle_problem(workload, small)

    def run(self, n_roots, sample_roots):
        per = max(1, sample_roots // self.cores)
        stride = max(1, n_roots // self.cores)
        jobs = [(self.workload, self.small, n_roots, min(i * stride, n_roots - per), per) for i in range(self.cores)]
        results = self.pool.map(_oracle_shard, jobs, chunksize=1) if self.pool else [_oracle_shard(jobs[0])]
        steps = sum(r[0] for r in results)
        busy = max(r[1] for r in results)
        return steps / busy, steps, busy

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def oracle_throughput(workload, small, n_roots, sample_roots, cores):
    pool = OraclePool(workload, small, cores)
    try:
        rate, steps, busy = pool.run(n_roots, sample_roots)
    finally:
        pool.close()
    return rate, steps, busy, busy


def run_reference(args, rank, world):
    """`--impl reference`: the reference's numpy/scipy CPU algorithm (oracle port) on all host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_roots = args.roots_per_gpu * world
    n_runs = args.warmup + args.steps
    # bounded sample: ~180 s of CPU work for the whole run at ~1 300 steps/s/core and ~109 steps per extremal
    per_core = args.cpu_sample_roots // cores if args.cpu_sample_roots else int(max(2, min(100, 180. / n_runs * 1300 / 109)))
    sample = per_core * cores
    pool = OraclePool(args.workload, args.small, cores)
    times, steps_tot = [], 0
    try:
        for i in range(n_runs):
            rate, steps, busy = pool.run(n_roots, sample)
            if i >= args.warmup:
                times.append(busy)
                steps_tot += steps
    finally:
        pool.close()
    value = steps_tot / sum(times)
    line = {
        'impl': 'reference', 'metric': 'extremal_rk_steps_per_sec', 'value': value, 'unit': 'steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(times) / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload + ('-small' if args.small else ''), 'roots_total': n_roots,
                   'parallelism': 'cpu x%d processes' % cores},
        'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d of %d root extremals (theta_0 shards of %d per core), full horizon, per step'
                                   % (sample, n_roots, per_core)},
        'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -----------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', choices=WORKLOADS, default='era5')
    ap.add_argument('--roots-per-gpu', type=int, default=0)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--small', action='store_true', help='reduced grid (debugging only; not a valid bench number)')
    ap.add_argument('--cpu-sample-roots', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--shard-block', type=int, default=5000,
                    help='N > 1: block-cyclic theta_0 sharding with this many roots per block (0: contiguous ranges)')
    ap.add_argument('--dtype', choices=('f64', 'f32'), default='f64', help="f32 = flow grid stored in fp32 (tolerance 1e-4)")
    ap.add_argument('--integrator', choices=('dopri5', 'rk4'), default='dopri5')
    args = ap.parse_args()
    if args.roots_per_gpu <= 0:
        args.roots_per_gpu = default_roots(args.workload)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        print('note: fewer than 3 warm-up steps requested', file=sys.stderr)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: dabry_b200 has no CPU execution path'
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from dabry_b200 import _native as nat
    from dabry_b200.distributed import shard_range
    from dabry_b200.solver_ef import SolverEFResampling, SolverEFTrimming

    pb, total, kw, descr = build_problem(args.workload, args.small)
    trimming = bool(kw.pop('trimming', False))
    solver_cls = SolverEFTrimming if trimming else SolverEFResampling
    variant = nat.TRIMMING if trimming else nat.RESAMPLING
    n_roots = args.roots_per_gpu * world
    shard = (rank, world, args.shard_block) if world > 1 and args.shard_block > 0 else None
    root_range = shard_range(n_roots, rank, world) if world > 1 and shard is None else None
    count = n_roots // world if shard else (root_range[1] if root_range else n_roots)
    cap_factor = 1.02 if kw.get('max_depth', 1) == 1 else 2.5
    max_sites = int(count * cap_factor) + 4096

    def make_solver(device_roots, **extra):
        args_ = dict(n_costate_sectors=n_roots, root_range=root_range, shard=shard, device=local_rank,
                     max_sites=max_sites, device_roots=device_roots, dtype=args.dtype, integrator=args.integrator)
        args_.update(kw)
        args_.update(extra)
        return solver_cls(pb, total, **args_)

    # ---- resident arm: grid uploaded once, K timed solves ----
    from dabry_b200.distributed import ShardedResampling, ShardedTrimming
    if world == 1:
        # single process: ShardedResampling needs a process group only for its collectives
        class _Solo:
            def __init__(self, solver):
                self.solver = solver

            def solve(self, fetch=True):
                h = self.solver._ensure_handle()
                h.call('dabry_solve', variant, nat.dptr(self.solver._root_costates()))
                return self
        runner = _Solo(make_solver(device_roots=True))
    else:
        runner = (ShardedTrimming if trimming else ShardedResampling)(
            lambda **k: make_solver(device_roots=True), n_roots, device=local_rank, block=args.shard_block or None)
    solver = runner.solver
    handle = solver._ensure_handle()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')  # > 126 MB L2

    def one_step():
        flush.zero_()  # L2 flush between iterations
        torch.cuda.synchronize()
        runner.solve(fetch=False)
        after = handle.stats()
        # dabry_solve times itself (ms_total); the split-step path of the sharded run is timed phase by phase
        if world > 1:
            after['ms_total'] = sum(after[k] for k in ('ms_integrate', 'ms_obstacle', 'ms_resample', 'ms_trim',
                                                       'ms_arrival'))
        return after

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    base = handle.stats()
    dev_ms, integ_ms, steps_rk = [], [], []
    phases = {}
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        st = one_step()
        dev_ms.append(st['ms_total'])
        integ_ms.append(st['ms_integrate'])
        steps_rk.append(st['rk_steps'])
        for k in ('ms_integrate', 'ms_obstacle', 'ms_resample', 'ms_trim', 'ms_arrival', 'ms_total'):
            phases[k] = phases.get(k, 0.) + st[k] / args.steps
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop()
    last = handle.stats()
    launches = int(last['kernel_launches'] - base['kernel_launches'])
    n_sites = int(last['n_sites'])
    cost_resident = handle.solution()[1]
    cost_global = getattr(runner, 'global_cost', cost_resident) if world > 1 else cost_resident

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    total_wall_ms = reduce_max(wall_ms)
    # N = 1: CUDA events around dabry_solve on the library's stream.  N > 1: the bracket between the two barriers
    # (it contains the ring-boundary exchange and the arrival all_reduce, which no single-stream event pair sees)
    total_dev_ms = reduce_max(float(sum(dev_ms))) if world == 1 else total_wall_ms
    total_steps = reduce_sum(float(sum(steps_rk)))
    value = total_steps / (total_dev_ms * 1e-3)
    integ_total_ms = float(sum(integ_ms))
    local_steps = float(sum(steps_rk))
    solver.close()
    del flush

    # ---- end-to-end arm: public API from host arrays, every step ----
    e2e = None
    if not args.no_e2e:
        inner = getattr(pb.model.ff, 'ff', pb.model.ff)
        h2d = int((inner.values.nbytes if hasattr(inner, 'values') else 0) + 16 * count)
        if hasattr(inner, 'pin_memory'):
            inner.pin_memory()  # the step's input (the flow samples) lives in pinned host memory, as the contract asks

        def e2e_step():
            s = make_solver(device_roots=False)
            s.solve()
            cost = s.solution_cost
            # what is read back: the optimum (16 B), the slots of the solution sites (compacted on the device) and the
            # optimal trajectory; the per-site arrival arrays stay on the device until someone asks for them
            d2h = 16 + 4 * int(s._solution_slots.shape[0])
            if s.solution_site is not None:
                traj = s.solution_site.traj
                d2h += int(traj.states.nbytes + traj.times.nbytes + traj.costates.nbytes)
            rk = s.native_stats()['rk_steps']
            s.close()
            return cost, rk, d2h

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        rk_sum, d2h = 0, 0
        n_e2e = max(1, min(args.steps, 3))
        step_ms, step_calls = [], []
        for _ in range(n_e2e):
            nat.CALL_TIMES = {}
            t1 = time.perf_counter()
            cost_e2e, rk, d2h = e2e_step()
            rk_sum += rk
            step_ms.append(round(1e3 * (time.perf_counter() - t1), 2))
            step_calls.append({k[6:]: round(1e3 * v, 2) for k, v in nat.CALL_TIMES.items() if v > 2e-4})
        nat.CALL_TIMES = None
        barrier()
        e2e_s = reduce_max(time.perf_counter() - t0)
        e2e = {'value': reduce_sum(float(rk_sum)) / e2e_s, 'unit': 'steps/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': int(d2h), 'steps': n_e2e, 'ms_per_step': 1e3 * e2e_s / n_e2e, 'ms_each_step_rank0': step_ms, 'native_ms_slowest_step_rank0': step_calls[int(np.argmax(step_ms))],
               'arrival_cost': None if np.isnan(cost_e2e) else cost_e2e}

    # ---- CPU baseline (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = 1
        sample = args.cpu_sample_roots if args.cpu_sample_roots else 160
        rate, steps, busy, _ = oracle_throughput(args.workload, args.small, n_roots, sample, cores)
        cpu = {'value': rate, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
               'sample': '%d of %d root extremals, full horizon (%d RK steps in %.1f s); %d host cores present'
                         % (sample, n_roots, steps, busy, os.cpu_count() or 1)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bytes_per_step = BYTES_PER_STEP[(args.integrator, args.dtype)]
        ach = local_steps * bytes_per_step / (integ_total_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the hot kernel at exactly this workload, from the
        # ncu --set full capture summarised in profiles/r01_ncu_full_freearc_final.txt (0.403 GB read + 4.970 GB written)
        traffic = 5.373e9 if (args.workload, args.roots_per_gpu, args.dtype, args.integrator, args.small) == \
            ('era5', 1_250_000, 'f64', 'dopri5', False) else None
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': traffic, 'kernel': 'phase_kernel<FreeArcPhase<GRID>>', 'peak_source': peak_src,
                    'algorithmic_bytes_per_step': bytes_per_step,
                    'kernel_ms_per_launch': integ_total_ms / args.steps,
                    'ncu_at_this_workload': {'l1_data_pipe_lsu_wavefronts_pct': 66.0, 'l1_lsu_writeback_pct': 45.7,
                                             'fp64_pipe_active_pct': 44.3, 'issue_active_pct': 50.1,
                                             'l1_hit_pct_global_loads': 96.4, 'l2_hit_pct': 81.0,
                                             'active_lanes_per_warp_instruction': 23.2},
                    'note': 'algorithmic bytes (8 nodes x 6 scalars per RHS, zero reuse credit, + 32 B output) / measured '
                            'free-arc kernel time.  frac > 1: theta-sorted neighbours share grid cells, so the gather is served '
                            'by L1 (96 % of the node sectors) and DRAM traffic is 60x below the algorithmic bytes.  The unit that '
                            'binds is the L1 data stage: every 16-byte node piece is delivered to all 32 lanes (4 wavefronts per '
                            'LDG.128, broadcast or not) and the Runge-Kutta stage values cycle through local memory; ncu: L1 '
                            'LSU data-pipe wavefronts 66 % of peak, fp64 pipe 44.3 %, issue slots 50.1 % - DESIGN.md 3 and 6'}
        if args.workload == 'gyre':
            roofline['note'] = 'analytic field: no gather, fp64-pipe bound; fraction of the gather roofline is nominal'
        line = {
            'metric': 'extremal_rk_steps_per_sec', 'value': value, 'unit': 'steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_dev_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64' if args.dtype == 'f64' else 'f64 state / f32 grid', 'data': 'synthetic',
            'config': {'workload': args.workload + ('-small' if args.small else ''), 'flow': descr,
                       'roots_per_gpu': args.roots_per_gpu, 'roots_total': n_roots, 'sites_rank0': n_sites,
                       'integrator': 'dopri5_scipy' if args.integrator == 'dopri5' else 'rk4_fixed', 'n_time': 100, 'max_depth': kw.get('max_depth'), 'solver': solver_cls.__name__,
                       'parallelism': 'theta0-shard x%d (%s)' % (world, 'block-cyclic, %d roots per block' % args.shard_block
                                                                 if shard else 'contiguous'), 'l2': 'flushed between steps (256 MiB write)',
                       'timing': 'cuda events on the library stream' if world == 1 else 'barrier-bracketed wall clock, max over ranks',
                       'wall_ms_per_step': total_wall_ms / args.steps,
                       'phase_ms_rank0': {k: round(v, 3) for k, v in phases.items()},
                       'arrival_cost_rank0': None if np.isnan(cost_resident) else cost_resident,
                       'arrival_cost_global': None if np.isnan(cost_global) else cost_global},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
