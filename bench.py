#!/usr/bin/env python
"""Benchmark of the extremal-field hot path (BASELINE.json: extremal-RK-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload era5|planar|gyre|regional] [--impl reference]

One step = one full `SolverEFResampling` / `SolverEFTrimming` solve (setup, integration of every extremal over the whole
horizon with events, resampling scan, trimming, arrival-time argmin) of the named workload on synthetic data:

  era5     (default) BASELINE config 4: ERA5-shaped spherical grid 1440 x 721 x 24, 1.25 M root extremals PER GPU
           (10 M over 8 GPUs), max_depth = 1, DOPRI5 restated from scipy, fp64 grid + fp64 state.  Weak scaling.
  planar   BASELINE config 3: planar unsteady DiscreteFF 1024 x 1024 x 48, 2^19 roots per GPU, max_depth = 4.
  gyre     BASELINE config 2 flow: analytic Gyre, 100 k roots per GPU, max_depth = 1 (fp64 pipe bound, no gather).
  regional BASELINE config 5: regional spherical grid + obstacle + penalty, SolverEFTrimming, 500 k roots per GPU.

`value`    : whole-job RK steps / device time (CUDA events on the solver's stream around the solve - they bracket the
             in-library NCCL exchanges too; max over ranks), flow grid resident in HBM, roots generated on the device.
`e2e`      : the same solve through the public API from (page-locked) HOST arrays, everything inside the timed region: a
             new solver per step, grid upload + repack (N = 1: streamed in groups of frames behind the free arcs;
             N > 1: 1 / N of the grid per rank over PCIe, the rest from the peers by all-gather), root costates upload,
             the sharded solve with its boundary exchange and global argmin, then the optimum, the slots of the solution
             sites and the optimal trajectory read back.
`roofline` : the unit that binds the hot kernel is the fp64 pipe (the gather is served by L1: DRAM traffic is 60x below
             the algorithmic bytes), so the fraction is fp64 work / measured fp64 peak; SURVEY 8(d)'s byte figure is
             kept beside it as `gather_equiv`.
`secondary`: one short line per other BASELINE config (N = 1: configs 2, 3, 5; N > 1: config 5).
`--impl reference`: the UNMODIFIED reference (baseline/_ref, dabry v1.1 through its own public API) on a bounded
             sample of the same workload, one process per host core; the oracle port when baseline/_ref is absent.
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
WORKLOADS = ('era5', 'planar', 'gyre', 'regional')
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def build_problem(workload, small=False, api=None):
    """(problem, total_duration, solver kwargs, description) - identical on every rank.  `api`: the module set to build
    the problem with (default: this package's drop-in classes; the reference arm passes the reference's own)."""
    from dabry_b200 import synthetic
    if api is None:
        from dabry_b200.flowfield import DiscreteFF
        from dabry_b200.misc import Coords, Utils
        from dabry_b200.problem import NavigationProblem
        from dabry_b200.obstacle import CircleObs
        from dabry_b200.penalty import CirclePenalty
        grid_kw = dict(force_no_diff=True)  # derivatives are computed on the device
    else:
        DiscreteFF, Coords, Utils, NavigationProblem = api.DiscreteFF, api.Coords, api.Utils, api.NavigationProblem
        CircleObs, CirclePenalty = api.CircleObs, api.CirclePenalty
        grid_kw = {}
    if workload == 'era5':
        shape = (6, 180, 91) if small else (24, 1440, 721)
        sc = synthetic.era5_like(*shape)
        ff = DiscreteFF(sc.values, sc.bounds, Coords.GCS, **grid_kw)
        pb = NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr, name='era5_like').rescale()
        # fixed horizon: 1.1 x the great-circle time at srf, in the rescaled clock (skips the host heuristics)
        t_gc = Utils.distance(sc.x_init, sc.x_target, Coords.GCS) / sc.srf
        total = 1.1 * t_gc / (1. / (sc.srf / Utils.EARTH_RADIUS))
        return pb, total, dict(max_depth=1), 'era5_like %dx%dx%d gcs' % shape
    if workload == 'planar':
        shape = (6, 128, 128) if small else (48, 1024, 1024)
        sc = synthetic.planar_waves(*shape)
        ff = DiscreteFF(sc.values, sc.bounds, Coords.CARTESIAN, **grid_kw)
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
        shape = (5, 41, 23) if small else (17, 181, 101)
        sc = synthetic.regional_constrained(*shape)
        ff = DiscreteFF(sc.values, sc.bounds, Coords.GCS, **grid_kw)
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


def _load_json(name):
    try:
        with open(os.path.join(ROOT, 'profiles', name)) as f:
            return json.load(f)
    except Exception:
        return None


def measured_hbm_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def fp64_peak_tflops():
    """fp64 pipe peak of a B200 (one DFMA = 2 flop), measured by profiles/microbench_fp64.cu - MEASURED_PEAKS.json has
    no fp64 entry.  Fallback: NVIDIA's nominal 40 TFLOP/s."""
    d = _load_json('r02_microbench_fp64.json')
    if d and d.get('fp64_tflops_sustained'):
        return float(d['fp64_tflops_sustained']), 'measured (profiles/microbench_fp64.cu -> profiles/r02_microbench_fp64.json)'
    return 40.0, 'fallback (nominal B200 fp64)'


def workload_signature(workload, roots_per_gpu, dtype, integrator, small):
    return '%s|%d|%s|%s%s' % (workload, roots_per_gpu, dtype, integrator, '|small' if small else '')


def roofline_for(sig, integrator, dtype, rk_steps_free, kernel_ms_total, n_launches, field_is_grid):
    """fp64-pipe roofline of the free-arc kernel.  `achieved` = fp64 work of the timed launches / their measured
    duration; the work per RK step is the ncu-derived constant of the capture of THIS workload
    (profiles/r02_hot_kernel_counters.json, thread-level DFMA / DMUL / DADD of one launch / its RK steps) when there is
    one, else the algorithmic estimate of SURVEY 8(d) - the source is stated.  Counters and DRAM traffic of a capture are
    printed only for the workload they were captured on."""
    caps = _load_json('r02_hot_kernel_counters.json') or {}
    cap = caps.get(sig)
    peak, peak_src = fp64_peak_tflops()
    if cap:
        flop_per_step = float(cap['fp64_flop_per_rk_step'])
        ops_src = 'ncu capture of this workload (profiles/r02_hot_kernel_counters.json)'
    else:
        rhs = 6.02 if integrator == 'dopri5' else 4.
        # per RHS ~170 mul/add (48 interpolation FMAs, weights, dynamics; sin, cos, sqrt, divisions expanded ~60) +
        # ~200 per step of stage combinations, error norm and dense output; FMA-dominated: 1.6 flop per instruction
        flop_per_step = (rhs * 170 + 200) * 1.6
        ops_src = 'estimate (SURVEY 8d flop count); no ncu capture of this workload'
    secs = kernel_ms_total * 1e-3
    ach = rk_steps_free * flop_per_step / secs / 1e12
    out = {'bound': 'fp64', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
           'traffic': cap.get('dram_bytes_per_launch') if cap else None,
           'kernel': 'phase_kernel<FreeArcPhase>', 'peak_source': peak_src,
           'fp64_flop_per_rk_step': flop_per_step, 'flop_per_step_source': ops_src,
           'rk_steps_free_per_solve': rk_steps_free / max(n_launches, 1),
           'kernel_ms_per_solve': kernel_ms_total / max(n_launches, 1)}
    if field_is_grid:
        hbm, hbm_src = measured_hbm_gbs()
        bps = BYTES_PER_STEP[(integrator, dtype)]
        gbs = rk_steps_free * bps / secs / 1e9
        out['gather_equiv'] = {'algorithmic_bytes_per_step': bps, 'gbs': gbs, 'hbm_peak_gbs': hbm, 'hbm_peak_source': hbm_src,
                               'frac_of_hbm': gbs / hbm,
                               'note': 'SURVEY 8(d) bytes with zero reuse credit; > 1 because theta-sorted neighbours share '
                                       'grid cells and the gather is served by L1 - not the binding unit'}
    if cap:
        out['ncu_at_this_workload'] = cap.get('counters')
    return out


# ---- CPU legs: the unmodified reference (baseline/_ref) or the oracle port -----------------------------------------
_CPU_PROBLEM = {}


def reference_installed():
    return os.path.isdir(os.path.join(REF_DIR, 'dabry'))


def _reference_api():
    """The reference's own modules from baseline/_ref, under the three import shims of SURVEY 8(c) (stub h5py / pygrib,
    numpy.infty); its source is not touched."""
    import types
    import warnings
    from types import SimpleNamespace
    warnings.filterwarnings('ignore')
    if not hasattr(np, 'infty'):
        np.infty = np.inf
    for missing in ('h5py', 'pygrib'):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                sys.modules[missing] = types.ModuleType(missing)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import dabry.solver_ef as sef
    from dabry.problem import NavigationProblem
    from dabry.flowfield import DiscreteFF
    from dabry.misc import Coords, Utils
    from dabry.obstacle import CircleObs
    from dabry.penalty import CirclePenalty
    return SimpleNamespace(NavigationProblem=NavigationProblem, DiscreteFF=DiscreteFF, Coords=Coords, Utils=Utils,
                           CircleObs=CircleObs, CirclePenalty=CirclePenalty, sef=sef)


def _cpu_problem(workload, small, kind):
    """Per-process cache: the problem (grid + host derivatives) is built once per worker."""
    key = (workload, small, kind)
    if key not in _CPU_PROBLEM:
        import warnings
        warnings.filterwarnings('ignore')
        if kind == 'reference':
            api = _reference_api()
            pb, total, kw, _ = build_problem(workload, small, api)
            _CPU_PROBLEM[key] = (pb, total, kw, api)
        else:
            pb, total, kw, _ = build_problem(workload, small)
            inner = getattr(pb.model.ff, 'ff', None)
            if inner is not None and getattr(inner, 'grad_values', 1) is None:
                inner.compute_derivatives()  # the GPU arm differentiates on the device; the CPU port needs the host copy
            _CPU_PROBLEM[key] = (pb, total, kw, None)
    return _CPU_PROBLEM[key]


def _cpu_warm(args):
    _cpu_problem(*args)
    return 0


class _IvpCounter:
    def __init__(self, fn):
        self.fn, self.steps = fn, 0

    def __call__(self, *a, **k):
        res = self.fn(*a, **k)
        self.steps += (res.nfev - 2) // 6  # canonical count
        return res


def _cpu_shard(args):
    workload, small, kind, n_roots, lo, count = args
    pb, total, kw, api = _cpu_problem(workload, small, kind)
    if kind == 'reference':
        # the reference has no root_range: its sample is a fan of `count` roots over the whole circle, its own solver class
        import contextlib
        import io
        sef = api.sef
        kw = dict(kw)
        cls = sef.SolverEFTrimming if kw.pop('trimming', False) else sef.SolverEFResampling
        counter = _IvpCounter(sef.scitg.solve_ivp)
        sef.scitg.solve_ivp = counter
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                solver = cls(pb, total, n_costate_sectors=count, **kw)
                t0 = time.perf_counter()
                solver.solve()
                dt = time.perf_counter() - t0
        finally:
            sef.scitg.solve_ivp = counter.fn
        return counter.steps, dt, float('nan')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from solver_port import Port
    kw = dict(kw)
    trimming = kw.pop('trimming', False)
    port = Port(pb, total, n_costate_sectors=n_roots, root_range=(lo, count), **kw)
    t0 = time.perf_counter()
    port.run_trimming() if trimming else port.run_resampling()
    return port.n_rk, time.perf_counter() - t0, port.solution_cost


class CpuPool:
    """`cores` worker processes, each holding its own copy of the problem; a call times `sample_roots` extremals of
    the workload split evenly over the cores."""

    def __init__(self, workload, small, cores, kind):
        from multiprocessing import get_context
        self.workload, self.small, self.cores, self.kind = workload, small, cores, kind
        self.pool = get_context('spawn').Pool(cores) if cores > 1 else None
        if self.pool:
            self.pool.map(_cpu_warm, [(workload, small, kind)] * cores, chunksize=1)
        else:
            _cpu_problem(workload, small, kind)

    def run(self, n_roots, sample_roots):
        per = max(2, sample_roots // self.cores)
        stride = max(1, n_roots // self.cores)
        jobs = [(self.workload, self.small, self.kind, n_roots, min(i * stride, n_roots - per), per)
                for i in range(self.cores)]
        results = self.pool.map(_cpu_shard, jobs, chunksize=1) if self.pool else [_cpu_shard(jobs[0])]
        steps = sum(r[0] for r in results)
        busy = max(r[1] for r in results)
        return steps / busy, steps, busy, per

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def cpu_cores():
    """Host threads the CPU leg uses: every core, bounded by memory (a worker of the era5 workload holds ~3 GB)."""
    cores = os.cpu_count() or 1
    try:
        with open('/proc/meminfo') as f:
            kb = int([l for l in f if l.startswith('MemAvailable')][0].split()[1])
        cores = max(1, min(cores, int(kb / 1e6 / 4)))
    except Exception:
        pass
    return cores


def cpu_sample_description(kind, sample, per, cores, n_roots):
    if kind == 'reference':
        return ('%d root extremals per step: %d solver instances of the unmodified reference (baseline/_ref, dabry v1.1), one '
                'per core, each a fan of %d roots over the full circle of the same problem, full horizon' % (per * cores, cores, per))
    return ('%d of %d root extremals (theta_0 shards of %d per core) through the oracle port, full horizon, per step'
            % (per * cores, n_roots, per))


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores."""
    if rank != 0:
        return
    kind = 'reference' if reference_installed() else 'port'
    cores = cpu_cores()
    n_roots = args.roots_per_gpu * world
    n_runs = args.warmup + args.steps
    # bounded sample: ~150 s of CPU work for the whole run; the reference does ~450 steps/s/core on a gridded spherical
    # problem, the port ~1 300, an extremal takes ~110 steps
    rate = 450. if kind == 'reference' else 1300.
    per_core = args.cpu_sample_roots // cores if args.cpu_sample_roots else int(max(2, min(100, 150. / n_runs * rate / 110)))
    pool = CpuPool(args.workload, args.small, cores, kind)
    times, steps_tot, per = [], 0, per_core
    try:
        for i in range(n_runs):
            _, steps, busy, per = pool.run(n_roots, per_core * cores)
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
        'cpu_baseline': {'value': value, 'unit': 'steps/s', 'cores': cores, 'kind': kind,
                         'sample': cpu_sample_description(kind, per * cores, per, cores, n_roots)},
        'e2e': {'value': value, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -----------------------------------------------------------------------------------------------

class Ctx:
    """Process-group plumbing of one bench process."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        assert torch.cuda.is_available(), 'bench.py needs a CUDA device: dabry_b200 has no CPU execution path'
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def reduce_max(self, x):
        return self._reduce(x, 'MAX')

    def reduce_min(self, x):
        return self._reduce(x, 'MIN')

    def reduce_sum(self, x):
        return self._reduce(x, 'SUM')


class Workload:
    """One BASELINE config on this rank: problem, sharding, solver factory, the two timed arms."""

    def __init__(self, ctx, name, roots_per_gpu, dtype, integrator, small, shard_block):
        from dabry_b200 import _native as nat
        from dabry_b200.distributed import shard_range, ShardedResampling, ShardedTrimming
        from dabry_b200.solver_ef import SolverEFResampling, SolverEFTrimming
        self.ctx, self.name, self.dtype, self.integrator, self.small = ctx, name, dtype, integrator, small
        self.nat = nat
        self.pb, self.total, kw, self.descr = build_problem(name, small)
        self.trimming = bool(kw.pop('trimming', False))
        self.kw = kw
        self.solver_cls = SolverEFTrimming if self.trimming else SolverEFResampling
        self.runner_cls = ShardedTrimming if self.trimming else ShardedResampling
        self.variant = nat.TRIMMING if self.trimming else nat.RESAMPLING
        self.roots_per_gpu = roots_per_gpu
        self.n_roots = roots_per_gpu * ctx.world
        self.shard_block = shard_block if ctx.world > 1 else 0
        if self.shard_block and roots_per_gpu % self.shard_block != 0:
            self.shard_block = 0
        self.shard = (ctx.rank, ctx.world, self.shard_block) if self.shard_block else None
        self.root_range = shard_range(self.n_roots, ctx.rank, ctx.world) if ctx.world > 1 and self.shard is None else None
        self.count = self.n_roots // ctx.world if self.shard else (self.root_range[1] if self.root_range else self.n_roots)
        cap_factor = 1.02 if kw.get('max_depth', 1) == 1 else 2.5
        self.max_sites = int(self.count * cap_factor) + 4096
        inner = getattr(self.pb.model.ff, 'ff', self.pb.model.ff)
        self.grid = inner if hasattr(inner, 'values') else None

    def signature(self):
        return workload_signature(self.name, self.roots_per_gpu, self.dtype, self.integrator, self.small)

    def make_solver(self, device_roots, **extra):
        args_ = dict(n_costate_sectors=self.n_roots, device=self.ctx.local_rank, max_sites=self.max_sites,
                     device_roots=device_roots, dtype=self.dtype, integrator=self.integrator)
        args_.update(self.kw)
        args_.update(extra)
        return self.solver_cls(self.pb, self.total, **args_)

    def make_runner(self, device_roots):
        """The sharded runner (every N): its solver calls dabry_solve_sharded, which is dabry_solve on one GPU."""
        return self.runner_cls(lambda **k: self.make_solver(device_roots, **k), self.n_roots, device=self.ctx.local_rank,
                               block=self.shard_block or None)

    # ---- resident arm: grid uploaded once, K timed solves ----
    def resident(self, steps, warmup, keep=False):
        ctx, torch = self.ctx, self.ctx.torch
        runner = self.make_runner(device_roots=True)
        solver = runner.solver
        handle = solver._ensure_handle()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')  # > 126 MB L2

        def one_step():
            flush.zero_()  # L2 flush between iterations
            torch.cuda.synchronize()
            runner.solve(fetch=False)
            return handle.stats()

        for _ in range(warmup):
            one_step()
        ctx.barrier()
        sampler = ClockSampler(ctx.local_rank)
        sampler.start()
        base = handle.stats()
        dev_ms, integ_ms, steps_rk, steps_free = [], [], [], []
        phases = {}
        t_wall = time.perf_counter()
        for _ in range(steps):
            st = one_step()
            dev_ms.append(st['ms_total'])
            integ_ms.append(st['ms_integrate'])
            steps_rk.append(st['rk_steps'])
            steps_free.append(st['rk_steps_free'])
            for k in ('ms_integrate', 'ms_obstacle', 'ms_resample', 'ms_trim', 'ms_arrival', 'ms_comm', 'ms_total'):
                phases[k] = phases.get(k, 0.) + st[k] / steps
        ctx.barrier()
        wall_ms = 1e3 * (time.perf_counter() - t_wall)
        clocks = sampler.stop()
        last = handle.stats()
        out = {
            'launches': int(last['kernel_launches'] - base['kernel_launches']),
            'n_sites': int(last['n_sites']),
            'cost_local': handle.solution()[1],
            'cost_global': runner.global_cost,
            'owner_rank': runner.owner_rank,
            # CUDA events on the solver's stream around dabry_solve_sharded (they bracket the NCCL exchanges, which are
            # enqueued on the same stream); the slowest rank counts
            'dev_ms_total': ctx.reduce_max(float(sum(dev_ms))),
            'wall_ms_total': ctx.reduce_max(wall_ms),
            'total_steps': ctx.reduce_sum(float(sum(steps_rk))),
            'local_steps': float(sum(steps_rk)), 'local_steps_free': float(sum(steps_free)),
            'integ_ms_total': float(sum(integ_ms)),
            'phases': {k: round(v, 3) for k, v in phases.items()}, 'clocks': clocks,
        }
        del flush
        if keep:
            out['runner'] = runner
        else:
            solver.close()
        return out

    # ---- end-to-end arm: public API from host arrays, every step ----
    def e2e(self, steps):
        ctx, nat = self.ctx, self.nat
        h2d = int((self.grid.values.nbytes // ctx.world if self.grid is not None else 0) + 16 * self.count)
        if self.grid is not None and hasattr(self.grid, 'pin_memory'):
            self.grid.pin_memory()  # the step's input (the flow samples) lives in pinned host memory

        def e2e_step():
            runner = self.make_runner(device_roots=False)
            runner.solve(fetch=True)
            s = runner.solver
            cost = runner.global_cost
            # what is read back: the optimum (16 B), the slots of the solution sites (compacted on the device) and the
            # optimal trajectory; the per-site arrival arrays stay on the device until someone asks for them
            d2h = 16 + 4 * int(s._solution_slots.shape[0])
            if s.solution_site is not None:
                traj = s.solution_site.traj
                d2h += int(traj.states.nbytes + traj.times.nbytes + traj.costates.nbytes)
            rk = s.native_stats()['rk_steps']
            s.close()
            return cost, rk, d2h

        for _ in range(2):  # untimed: the first solves of a new shape pay pool growth and page-locking once
            e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        rk_sum, d2h, cost = 0, 0, float('nan')
        step_ms, step_calls = [], []
        for _ in range(steps):
            nat.CALL_TIMES = {}
            t1 = time.perf_counter()
            cost, rk, d2h = e2e_step()
            rk_sum += rk
            step_ms.append(round(1e3 * (time.perf_counter() - t1), 2))
            step_calls.append({k[6:]: round(1e3 * v, 2) for k, v in nat.CALL_TIMES.items() if v > 2e-4})
        nat.CALL_TIMES = None
        ctx.barrier()
        secs = ctx.reduce_max(time.perf_counter() - t0)
        return {'value': ctx.reduce_sum(float(rk_sum)) / secs, 'unit': 'steps/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': int(d2h), 'steps': steps, 'ms_per_step': 1e3 * secs / steps,
                'ms_each_step_rank0': step_ms, 'native_ms_slowest_step_rank0': step_calls[int(np.argmax(step_ms))],
                'path': 'ShardedResampling/ShardedTrimming -> dabry_solve_sharded (in-library NCCL)' if ctx.world > 1
                        else 'SolverEF*.solve -> dabry_solve',
                'arrival_cost': None if np.isnan(cost) else cost}

    # ---- driver-visible correctness of the sharding ----
    def shard_check(self, runner, sample=1000):
        """Every rank re-solves a sample of its NEIGHBOUR's roots on its own GPU (a plain single-handle solve of the ring
        range) and compares them bit for bit with what the neighbour holds from the timed sharded solve."""
        ctx, torch, dist = self.ctx, self.ctx.torch, self.ctx.dist
        if ctx.world == 1:
            return 'n/a (1 GPU)'
        if self.trimming or self.kw.get('max_depth', 1) != 1:
            return 'skipped (children and trimming depend on the whole front)'
        n = min(sample, self.shard_block or self.count)
        mine = runner.solver._handle
        rec = mine.sites(0, n)
        own = np.concatenate((mine.trajs(0, n, want=('states',))['states'].reshape(n, -1),
                              rec['n_samples'][:, None].astype(float), rec['closure'][:, None].astype(float)), axis=1)
        t_own = torch.from_numpy(np.ascontiguousarray(own)).cuda()
        gathered = [torch.empty_like(t_own) for _ in range(ctx.world)]
        dist.all_gather(gathered, t_own)
        nb = (ctx.rank + 1) % ctx.world
        first = nb * self.shard_block if self.shard_block else self.root_range_of(nb)[0]
        # same build of the arc kernels as the big fan: small batches normally run the latency build, in which ptxas
        # contracts multiply-adds differently - DABRY_B200_WIDE_BATCH=0 pins the throughput build
        prev = os.environ.get('DABRY_B200_WIDE_BATCH')
        os.environ['DABRY_B200_WIDE_BATCH'] = '0'
        try:
            s = self.make_solver(device_roots=True, root_range=(first, n), max_sites=n + 4096)
            s.solve()
        finally:
            if prev is None:
                del os.environ['DABRY_B200_WIDE_BATCH']
            else:
                os.environ['DABRY_B200_WIDE_BATCH'] = prev
        h = s._handle
        rec2 = h.sites(0, n)
        again = np.concatenate((h.trajs(0, n, want=('states',))['states'].reshape(n, -1),
                                rec2['n_samples'][:, None].astype(float), rec2['closure'][:, None].astype(float)), axis=1)
        s.close()
        theirs = gathered[nb].cpu().numpy()
        same = np.array_equal(again.view(np.uint64), theirs.view(np.uint64))  # bit for bit, NaN padding included
        ok = ctx.reduce_min(1. if same else 0.)
        return 'ok' if ok == 1. else 'MISMATCH'

    def root_range_of(self, rank):
        from dabry_b200.distributed import shard_range
        return shard_range(self.n_roots, rank, self.ctx.world)


def short_line(w, r):
    return {'workload': w.name, 'flow': w.descr, 'solver': w.solver_cls.__name__, 'roots_per_gpu': w.roots_per_gpu,
            'integrator': w.integrator, 'sites_rank0': r['n_sites'],
            'value': r['total_steps'] / (r['dev_ms_total'] * 1e-3), 'unit': 'steps/s',
            'ms_per_step': None, 'phase_ms_rank0': r['phases'], 'gpu_launches': r['launches'],
            'arrival_cost_global': None if np.isnan(r['cost_global']) else r['cost_global']}


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
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--shard-block', type=int, default=5000,
                    help='N > 1: block-cyclic theta_0 sharding with this many roots per block (0: contiguous ranges)')
    ap.add_argument('--dtype', choices=('f64', 'f32'), default='f64', help="f32 = flow grid stored in fp32 (tolerance 1e-4)")
    ap.add_argument('--integrator', choices=('dopri5', 'rk4'), default='dopri5')
    args = ap.parse_args()
    if args.roots_per_gpu <= 0:
        args.roots_per_gpu = default_roots(args.workload)
    if args.impl == 'reference':
        run_reference(args, int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')))
        return
    if args.warmup < 3:
        print('note: fewer than 3 warm-up steps requested', file=sys.stderr)

    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    w = Workload(ctx, args.workload, args.roots_per_gpu, args.dtype, args.integrator, args.small, args.shard_block)
    r = w.resident(args.steps, args.warmup, keep=True)
    check = w.shard_check(r['runner'])
    r.pop('runner').solver.close()
    value = r['total_steps'] / (r['dev_ms_total'] * 1e-3)

    e2e = None
    if not args.no_e2e:
        e2e = w.e2e(max(1, min(args.steps, 8)))

    # ---- the other BASELINE configs, short (same build, same process) ----
    secondary = None
    if not args.no_secondary and not args.small:
        secondary = {}
        others = [('gyre', 'rk4'), ('planar', 'dopri5'), ('regional', 'dopri5')] if world == 1 else [('regional', 'dopri5')]
        for name, integ in others:
            if name == args.workload:
                continue
            try:
                w2 = Workload(ctx, name, default_roots(name), 'f64', integ, False, args.shard_block)
                r2 = w2.resident(3, 2)
                line2 = short_line(w2, r2)
                line2['ms_per_step'] = r2['dev_ms_total'] / 3
                line2['wall_ms_per_step'] = r2['wall_ms_total'] / 3
                secondary[name + ('_rk4' if integ == 'rk4' else '')] = line2
                del w2
            except Exception as exc:  # a secondary line must not take the headline down
                secondary[name] = {'error': repr(exc)[:300]}

    # ---- CPU baseline (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = 'reference' if reference_installed() else 'port'
        cores = cpu_cores()
        per_core = (args.cpu_sample_roots // cores) if args.cpu_sample_roots else (64 if kind == 'reference' else 160)  # ~10 s of CPU work
        pool = CpuPool(args.workload, args.small, cores, kind)
        try:
            rate, steps, busy, per = pool.run(w.n_roots, per_core * cores)
        finally:
            pool.close()
        cpu = {'value': rate, 'unit': 'steps/s', 'cores': cores, 'kind': kind,
               'sample': cpu_sample_description(kind, per * cores, per, cores, w.n_roots)
                         + ' (%d RK steps in %.1f s; %d host cores present)' % (steps, busy, os.cpu_count() or 1)}

    if rank == 0:
        roofline = roofline_for(w.signature(), args.integrator, args.dtype, r['local_steps_free'], r['integ_ms_total'],
                                args.steps, w.grid is not None)
        cost_local, cost_global = r['cost_local'], r['cost_global']
        line = {
            'metric': 'extremal_rk_steps_per_sec', 'value': value, 'unit': 'steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['dev_ms_total'] / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64' if args.dtype == 'f64' else 'f64 state / f32 grid', 'data': 'synthetic',
            'config': {'workload': args.workload + ('-small' if args.small else ''), 'flow': w.descr,
                       'roots_per_gpu': args.roots_per_gpu, 'roots_total': w.n_roots, 'sites_rank0': r['n_sites'],
                       'integrator': 'dopri5_scipy' if args.integrator == 'dopri5' else 'rk4_fixed', 'n_time': 100,
                       'max_depth': w.kw.get('max_depth'), 'solver': w.solver_cls.__name__,
                       'parallelism': 'theta0-shard x%d (%s)' % (world, 'block-cyclic, %d roots per block' % w.shard_block
                                                                 if w.shard else 'contiguous'),
                       'collectives': 'in-library NCCL on the solver stream (ring send/recv, all-gather argmin)' if world > 1 else None,
                       'l2': 'flushed between steps (256 MiB write)',
                       'timing': 'cuda events on the solver stream around the (sharded) solve, max over ranks',
                       'wall_ms_per_step': r['wall_ms_total'] / args.steps,
                       'phase_ms_rank0': r['phases'],
                       'arrival_cost_rank0': None if np.isnan(cost_local) else cost_local,
                       'arrival_cost_global': None if np.isnan(cost_global) else cost_global,
                       'owner_rank': r['owner_rank']},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': r['launches'], 'clocks': r['clocks'],
            'shard_check': check, 'secondary': secondary,
        }
        print(json.dumps(line), flush=True)
    if ctx.dist:
        from dabry_b200.distributed import close_comms
        close_comms()
        ctx.dist.destroy_process_group()


if __name__ == '__main__':
    main()
