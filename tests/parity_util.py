"""Shared helpers of the parity tests: run a registered case through this package's drop-in solvers and compare
the flattened result (oracle/dump.py layout) with a golden record produced by the unmodified reference
(oracle/gen_golden.py).  Bars (BASELINE.json north_star): extremal counts, names, indices and termination flags
bit-exact; endpoints, entry times and arrival cost within 1e-9 (relative to the column scale, >= 1)."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
HOSTSIM_DIR = os.path.join(ROOT, 'tests', 'hostsim')
HOSTSIM_LIB = os.path.join(HOSTSIM_DIR, 'libdabry_hostsim.so')

EXACT_FIELDS = ['index_t_init', 'traj_len', 'closure', 'neuter', 'obstacle', 'index_t_obs', 'obs_trigo',
                'index_t_check_next']
FLOAT_FIELDS = ['first', 'last', 'state_sum', 't_enter_obs']
FP64_TOL = 1e-9


def package_api():
    import dabry_b200.solver_ef as sef
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.flowfield import DiscreteFF
    from dabry_b200.misc import Coords
    from dabry_b200.obstacle import CircleObs
    from dabry_b200.penalty import CirclePenalty, DiscretePenalty
    import dabry_b200.flowfield as flowfield
    return SimpleNamespace(flowfield=flowfield, NavigationProblem=NavigationProblem, DiscreteFF=DiscreteFF, Coords=Coords,
                           CircleObs=CircleObs, CirclePenalty=CirclePenalty, DiscretePenalty=DiscretePenalty,
                           SolverEFResampling=sef.SolverEFResampling, SolverEFTrimming=sef.SolverEFTrimming,
                           SolverEFSimple=sef.SolverEFSimple)


def build_hostsim():
    """g++ build of the device headers behind the C ABI (tests/hostsim/hostsim.cpp) - test double for CPU-only CI."""
    import subprocess
    src = os.path.join(HOSTSIM_DIR, 'hostsim.cpp')
    deps = [src] + [os.path.join(ROOT, 'dabry_b200', 'csrc', f) for f in os.listdir(os.path.join(ROOT, 'dabry_b200', 'csrc'))]
    deps.append(os.path.join(ROOT, 'include', 'dabry_b200.h'))
    if os.path.exists(HOSTSIM_LIB) and all(os.path.getmtime(HOSTSIM_LIB) >= os.path.getmtime(d) for d in deps):
        return HOSTSIM_LIB
    subprocess.run(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off', '-x', 'c++', src, '-o',
                    HOSTSIM_LIB], check=True)
    return HOSTSIM_LIB


def golden(key):
    return np.load(os.path.join(GOLDEN_DIR, key + '.npz'))


def golden_keys(kind=None):
    keys = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz'))
    return [k for k in keys if kind is None or k.endswith('_' + kind)]


def solve_case(key, horizon=None):
    """horizon: `total_duration` to pass explicitly (the golden record's, bit for bit) instead of letting the solver
    derive it from the feedback flights - step counts are compared exactly, and one more step is attempted per arc when
    the horizon moves by an ulp (100 steps of 0.01 T either reach T or fall one ulp short of it)."""
    from cases import CASES, SOLVERS, build_problem, run_solver, solver_args
    api = package_api()
    case = CASES[key]
    pb = build_problem(case, api)
    if horizon is None or case.solver == 'S':
        return case, run_solver(case, pb, api)
    args, kwargs = solver_args(case, pb)
    solver = getattr(api, SOLVERS[case.solver])(pb, float(horizon), **kwargs)
    solver.solve()
    return case, solver


def scaled_diff(a, b):
    """max |a - b| / max(1, column scale of b); NaN must match NaN."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.
    assert np.array_equal(np.isnan(a), np.isnan(b)), 'NaN pattern differs'
    d = np.abs(a - b)
    d[np.isnan(d)] = 0.
    if a.ndim >= 2:
        scale = np.maximum(1., np.nanmax(np.abs(np.where(np.isnan(b), 0., b)), axis=0))
    else:
        scale = max(1., float(np.nanmax(np.abs(np.where(np.isnan(b), 0., b)))))
    return float(np.max(d / scale))


def compare_sites_record(rec, g, tol=FP64_TOL):
    """Returns a dict of diagnostics; raises AssertionError on any violation."""
    assert int(rec['n_sites']) == int(g['n_sites']), ('site count', int(rec['n_sites']), int(g['n_sites']))
    assert list(rec['event_names']) == list(g['event_names'])
    assert np.array_equal(rec['names'], g['names']), 'site names / insertion order differ'
    for f in EXACT_FIELDS:
        bad = np.nonzero(np.asarray(rec[f]) != np.asarray(g[f]))[0]
        assert bad.size == 0, (f, bad[:8], np.asarray(rec[f])[bad[:8]], np.asarray(g[f])[bad[:8]])
    out = {}
    for f in FLOAT_FIELDS:
        out[f] = scaled_diff(rec[f], g[f])
        assert out[f] <= tol, (f, out[f])
    assert bool(rec['success']) == bool(g['success'])
    for f in ('total_duration', 'max_int_step'):
        assert abs(float(rec[f]) - float(g[f])) <= tol * max(1., abs(float(g[f]))), f
    assert scaled_diff(rec['times'], g['times']) <= tol
    gc, rc = float(g['solution_cost']), float(rec['solution_cost'])
    if np.isnan(gc):
        assert np.isnan(rc)
    else:
        assert abs(rc - gc) <= tol * max(1., abs(gc)), ('arrival cost', rc, gc)
    out['solution_cost'] = 0. if np.isnan(gc) else abs(rc - gc)
    # the reference breaks cost ties by Python set order: compare the SET of solution sites, and the winner only
    # when the optimum is unique among them
    assert sorted(rec['solution_names']) == sorted(g['solution_names']), 'solution site sets differ'
    return out


def compare_full_trajectories(rec, g, tol=FP64_TOL):
    assert np.array_equal(rec['traj_offsets'], g['traj_offsets'])
    out = {}
    for f in ('traj_times', 'traj_states', 'traj_costates', 'traj_controls', 'traj_cost'):
        out[f] = scaled_diff(rec[f], g[f])
        assert out[f] <= tol, (f, out[f])
    return out


def simple_record(solver):
    """Flatten a SolverEFSimple (list of Trajectory groups) like oracle/gen_golden.py does."""
    trajs = solver.trajs
    ev_names = list(solver.events.keys())
    lens = np.array([len(t) for t in trajs], dtype=np.int32)
    offs = np.zeros(len(trajs) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    rec = {'n_trajs': np.int64(len(trajs)), 'event_names': np.array(ev_names), 'traj_len': lens, 'traj_offsets': offs,
           'traj_times': np.concatenate([t.times for t in trajs]),
           'traj_states': np.concatenate([t.states.reshape(-1, 2) for t in trajs]),
           'traj_costates': np.concatenate([t.costates.reshape(-1, 2) for t in trajs]),
           'traj_cost': np.concatenate([t.cost for t in trajs]),
           'times': np.asarray(solver.times), 'total_duration': np.float64(solver.total_duration),
           'max_int_step': np.float64(solver.max_int_step), 'success': np.bool_(solver.success)}
    for e in ev_names:
        rec['n_events_' + e] = np.array([t.events[e].shape[0] for t in trajs], dtype=np.int32)
        rec['t_events_' + e] = np.concatenate([t.events[e].reshape(-1) for t in trajs] + [np.zeros(0)])
    return rec


def compare_simple_record(rec, g, tol=FP64_TOL):
    assert int(rec['n_trajs']) == int(g['n_trajs'])
    assert list(rec['event_names']) == list(g['event_names'])
    assert np.array_equal(rec['traj_len'], g['traj_len']), 'trajectory lengths differ'
    out = {}
    for f in ('traj_times', 'traj_states', 'traj_costates', 'traj_cost', 'times'):
        out[f] = scaled_diff(rec[f], g[f])
        assert out[f] <= tol, (f, out[f])
    for e in rec['event_names']:
        assert np.array_equal(rec['n_events_' + e], g['n_events_' + e]), ('event counts', e)
        out['t_events_' + e] = scaled_diff(rec['t_events_' + e], g['t_events_' + e])
        assert out['t_events_' + e] <= tol
    assert bool(rec['success']) == bool(g['success'])
    return out


def tie_set(key):
    """Sites of `key` whose record in the UNMODIFIED reference changes when its right-hand side / obstacle entry state
    moves by <= 1 ulp: tests/golden/tie_sets.json, measured by oracle/tie_sensitivity.py on 64 seeded runs of the
    reference.  Empty for most cases."""
    import json
    path = os.path.join(GOLDEN_DIR, 'tie_sets.json')
    with open(path) as f:
        cases = json.load(f)['cases']
    return set(cases.get(key, {}).get('sites', ()))


def border_sensitive(rec, bl, tr, n_map=100):
    """Trimming: the sites of a flattened run whose fate the outermost line of the cost map decides - captured by the
    domain FRAME (they slide at x = +-1e-17 of that line, whose cells flip between inf and finite with the sign of the
    residue, misc.py:138 `a > 0`), or ending within one map cell of the border (their cost-map lookup, solver_ef.py:832,
    reads those cells)."""
    names = [str(n) for n in rec['names']]
    frame = [i for i, e in enumerate(rec['event_names']) if 'FrameObs' in str(e)]
    cell = (np.asarray(tr, dtype=float) - np.asarray(bl, dtype=float)) / (n_map - 1)
    out = set()
    for n, o, last in zip(names, rec['obstacle'], rec['last']):
        x = np.asarray(last[:2], dtype=float)
        if int(o) in frame or np.any(x - bl < cell) or np.any(tr - x < cell):
            out.add(n)
    return out


def compare_except_ties(rec, g, ties, mngr, tol=FP64_TOL):
    """Everything `compare_sites_record` checks, site by site.  A site may differ only if

      (a) it is in `ties` - tie-sensitive in the reference itself (measured, or border-sensitive in either run), or
      (b) one of its parents differs or is missing (a child exists, and is born where it is, only as long as both parents
          are open: a difference propagates down a family, never sideways);

    every other site - flags, indices, lengths, endpoints, checksums - must match, and so must the global outcome
    (success, arrival cost, site count, time grid) and, for the tie-sensitive sites present in both runs, everything that
    precedes their capture."""
    from dump import differing_sites
    assert int(rec['n_sites']) == int(g['n_sites']), ('site count', int(rec['n_sites']), int(g['n_sites']))
    assert list(rec['event_names']) == list(g['event_names'])
    assert bool(rec['success']) == bool(g['success'])
    for f in ('total_duration', 'max_int_step'):
        assert abs(float(rec[f]) - float(g[f])) <= tol * max(1., abs(float(g[f]))), f
    assert scaled_diff(rec['times'], g['times']) <= tol
    gc, rc = float(g['solution_cost']), float(rec['solution_cost'])
    assert (np.isnan(gc) and np.isnan(rc)) or abs(rc - gc) <= tol * max(1., abs(gc)), ('arrival cost', rc, gc)
    differing = differing_sites(rec, g, tol)
    unexplained = []
    for n in differing:
        if n in ties:
            continue
        parents = mngr.parents_name_from_name(n) if '-' in n else (None, None)
        if not any(p in differing for p in parents):
            unexplained.append(n)
    assert not unexplained, ('sites differing with no tie-sensitive ancestor', sorted(unexplained)[:10], len(unexplained))
    # insertion order of everything that does not differ
    ours = [n for n in rec['names'] if n not in differing]
    theirs = [n for n in g['names'] if n not in differing]
    assert ours == theirs, 'site names / insertion order differ'
    sol_diff = set(map(str, rec['solution_names'])) ^ set(map(str, g['solution_names']))
    assert sol_diff <= differing | ties, ('solution site sets differ', sorted(sol_diff))
    # differing sites present in both runs: their birth must still agree unless a parent differs (index and state)
    a = {str(n): i for i, n in enumerate(rec['names'])}
    b = {str(n): i for i, n in enumerate(g['names'])}
    for n in differing & ties & set(a) & set(b):
        parents = mngr.parents_name_from_name(n) if '-' in n else (None, None)
        if any(p in differing for p in parents):
            continue
        i, j = a[n], b[n]
        assert rec['index_t_init'][i] == g['index_t_init'][j], (n, 'index_t_init')
        x, y = np.asarray(rec['first'][i], dtype=float), np.asarray(g['first'][j], dtype=float)
        assert np.array_equal(np.isnan(x), np.isnan(y)), (n, 'first')
        assert np.all(np.abs(x - y)[~np.isnan(y)] <= 1e-6 * np.maximum(1., np.abs(y[~np.isnan(y)]))), (n, 'first')
    return {'differing': len(differing), 'rooted_in_ties': len(differing & ties), 'tie_set': len(ties),
            'sites': int(g['n_sites'])}


def compare_solution_full(rec, g, tol=FP64_TOL):
    """extrapolate_back_traj (solver_ef.py:701-749): the optimal trajectory extended back to t_init through its
    ancestors, when both runs crown the same site (cost ties are broken by Python set order in the reference)."""
    if 'solution_full_states' not in g or str(rec['solution_name']) != str(g['solution_name']):
        return None
    assert 'solution_full_states' in rec, 'traj_full of the solution site missing'
    out = {}
    for f in ('solution_full_times', 'solution_full_states'):
        assert np.asarray(rec[f]).shape == np.asarray(g[f]).shape, (f, np.asarray(rec[f]).shape, np.asarray(g[f]).shape)
        out[f] = scaled_diff(rec[f], g[f])
        assert out[f] <= tol, (f, out[f])
    return out


def port_record(port):
    """Flatten an oracle `Port` (oracle/solver_port.py) into the golden record layout."""
    sites = list(port.sites.values())
    ev = port.event_names
    n = len(sites)
    rec = {
        'n_sites': np.int64(n), 'names': np.array([s.name for s in sites]), 'event_names': np.array(ev),
        'index_t_init': np.array([s.index_t_init for s in sites], dtype=np.int32),
        'traj_len': np.array([len(s.t) for s in sites], dtype=np.int32),
        'closure': np.array([-1 if s.closure is None else s.closure for s in sites], dtype=np.int32),
        'neuter': np.array([-1 if s.neuter is None else s.neuter for s in sites], dtype=np.int32),
        'obstacle': np.array([ev.index(s.obstacle_name) if s.obstacle_name else -1 for s in sites], dtype=np.int32),
        'index_t_obs': np.array([s.index_t_obs for s in sites], dtype=np.int32),
        'obs_trigo': np.array([bool(s.obs_trigo) for s in sites]),
        't_enter_obs': np.array([np.nan if s.t_enter_obs is None else s.t_enter_obs for s in sites]),
        'index_t_check_next': np.array([s.check_next for s in sites], dtype=np.int32),
        'times': port.times, 'total_duration': np.float64(port.total_duration),
        'max_int_step': np.float64(port.max_int_step), 'success': np.bool_(port.success),
        'first': np.array([np.concatenate((s.x[0], s.p[0], s.u[0])) for s in sites]),
        'last': np.array([np.concatenate((s.x[-1], s.p[-1], s.u[-1])) for s in sites]),
        'state_sum': np.array([np.sum(np.array(s.x), axis=0) for s in sites]),
        'solution_name': np.array(port.solution_site.name if port.solution_site is not None else ''),
        'solution_cost': np.float64(port.solution_cost),
        'solution_names': np.array(sorted(s.name for s in port.solution_sites)),
    }
    offs = np.zeros(n + 1, dtype=np.int64)
    offs[1:] = np.cumsum(rec['traj_len'])
    rec['traj_offsets'] = offs
    rec['traj_times'] = np.concatenate([np.array(s.t) for s in sites])
    rec['traj_states'] = np.concatenate([np.array(s.x).reshape(-1, 2) for s in sites])
    rec['traj_costates'] = np.concatenate([np.array(s.p).reshape(-1, 2) for s in sites])
    rec['traj_controls'] = np.concatenate([np.array(s.u).reshape(-1, 2) for s in sites])
    rec['traj_cost'] = np.concatenate([np.array(s.cost) for s in sites])
    return rec


def solve_case_port(key, api=None):
    """Run a registered case through the CPU oracle (oracle/solver_port.py) on `api`'s problem classes
    (default: this package's)."""
    from cases import CASES, build_problem, solver_args
    from solver_port import Port
    api = package_api() if api is None else api
    case = CASES[key]
    pb = build_problem(case, api)
    args, kwargs = solver_args(case, pb)
    port = Port(*args, **kwargs)
    if case.solver == 'T':
        port.run_trimming()
    elif case.solver == 'R':
        port.run_resampling()
    return case, port
