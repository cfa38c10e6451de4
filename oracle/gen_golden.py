"""Test infrastructure: generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (dabry v1.1 at
/root/reference, scipy 1.18.1 / numpy 2.3.5 in this container) on the registered parity cases.

Usage:  python -W ignore oracle/gen_golden.py [-j 8] [case_key ...]

The reference pins nothing numeric (its tests assert only ``solver.success``), so these vectors are what pins
parity for this build.  Compact records are committed for every case; full ragged trajectories
are committed only for the cases flagged ``full`` and written for all cases to ``tests/golden_full/``
(git-ignored, local debugging only).
"""
import os
import sys
import time
from multiprocessing import Pool
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402


def reference_api():
    import refshim
    refshim.load_reference()
    import dabry.solver_ef as sef
    from dabry.problem import NavigationProblem
    from dabry.flowfield import DiscreteFF
    from dabry.misc import Coords
    from dabry.obstacle import CircleObs
    from dabry.penalty import CirclePenalty, DiscretePenalty
    import dabry.flowfield as flowfield
    return SimpleNamespace(flowfield=flowfield, NavigationProblem=NavigationProblem, DiscreteFF=DiscreteFF, Coords=Coords,
                           CircleObs=CircleObs, CirclePenalty=CirclePenalty, DiscretePenalty=DiscretePenalty,
                           SolverEFResampling=sef.SolverEFResampling, SolverEFTrimming=sef.SolverEFTrimming,
                           SolverEFSimple=sef.SolverEFSimple, sef=sef)


class IvpCounter:
    """Wraps scipy.integrate.solve_ivp as seen by dabry.solver_ef to count calls, nfev and exit statuses."""

    def __init__(self, fn):
        self.fn = fn
        self.reset()

    def reset(self):
        self.calls = 0
        self.nfev = 0
        self.steps = 0
        self.status = {0: 0, 1: 0, -1: 0}

    def __call__(self, *a, **k):
        res = self.fn(*a, **k)
        self.calls += 1
        self.nfev += res.nfev
        self.steps += (res.nfev - 2) // 6
        self.status[res.status] += 1
        return res


def generate(key):
    import io
    import contextlib
    from cases import CASES, build_problem, run_solver
    from dump import solver_record
    api = reference_api()
    case = CASES[key]
    pb = build_problem(case, api)
    import scipy.integrate
    orig = scipy.integrate.solve_ivp
    counter = IvpCounter(orig)
    # count only the solver's own calls: install after the problem (and its feedback solves) exist
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        # auto_time_upper_bound runs in the constructor; install counter lazily by patching, then reset in solve
        api.sef.scitg.solve_ivp = counter
        try:
            from cases import SOLVERS, solver_args
            cls = getattr(api, SOLVERS[case.solver])
            args, kwargs = solver_args(case, pb)
            solver = cls(*args, **kwargs)
            counter.reset()
            t1 = time.perf_counter()
            if case.solver == 'S':
                solver.setup()
                for _ in range(case.n_steps):
                    solver.step()
            else:
                solver.solve()
            t_solve = time.perf_counter() - t1
        finally:
            api.sef.scitg.solve_ivp = orig
    if case.solver == 'S':
        rec = simple_record(solver)
    else:
        rec = solver_record(solver, full=True, next_nb=True)
    rec['ref_solve_seconds'] = np.float64(t_solve)
    rec['ref_ivp_calls'] = np.int64(counter.calls)
    rec['ref_nfev'] = np.int64(counter.nfev)
    rec['ref_rk_steps'] = np.int64(counter.steps)
    rec['ref_ivp_status'] = np.array([counter.status[0], counter.status[1], counter.status[-1]])
    os.makedirs(os.path.join(ROOT, 'tests', 'golden_full'), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden_full', key + '.npz'), **rec)
    if not case.full:
        for k in list(rec.keys()):
            if k.startswith('traj_') and k != 'traj_len':
                del rec[k]
        rec.pop('next_nb', None)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', key + '.npz'), **rec)
    return key, float(t_solve), int(rec['n_sites']) if 'n_sites' in rec else -1, \
        float(rec.get('solution_cost', np.nan)), int(counter.steps), time.perf_counter() - t0


def simple_record(solver):
    """SolverEFSimple keeps plain Trajectory groups (solver_ef.py:430-436)."""
    trajs = solver.trajs
    ev_names = list(solver.events.keys())
    n = len(trajs)
    lens = np.array([len(t) for t in trajs], dtype=np.int32)
    offs = np.zeros(n + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    rec = {
        'n_trajs': np.int64(n),
        'event_names': np.array(ev_names),
        'traj_len': lens,
        'traj_offsets': offs,
        'traj_times': np.concatenate([t.times for t in trajs]),
        'traj_states': np.concatenate([t.states.reshape(-1, 2) for t in trajs]),
        'traj_costates': np.concatenate([t.costates.reshape(-1, 2) for t in trajs]),
        'traj_cost': np.concatenate([t.cost for t in trajs]),
        'times': np.asarray(solver.times),
        'total_duration': np.float64(solver.total_duration),
        'max_int_step': np.float64(solver.max_int_step),
        'success': np.bool_(solver.success),
    }
    for e in ev_names:
        rec['n_events_' + e] = np.array([t.events[e].shape[0] for t in trajs], dtype=np.int32)
        rec['t_events_' + e] = np.concatenate([t.events[e].reshape(-1) for t in trajs] + [np.zeros(0)])
    return rec


if __name__ == '__main__':
    from cases import CASES
    argv = sys.argv[1:]
    jobs = 8
    if argv and argv[0] == '-j':
        jobs = int(argv[1])
        argv = argv[2:]
    keys = argv if argv else list(CASES.keys())
    with Pool(jobs) as pool:
        for out in pool.imap_unordered(generate, keys):
            print('%-32s solve %.1fs sites %d cost %.15g steps %d (total %.1fs)' % out, flush=True)
