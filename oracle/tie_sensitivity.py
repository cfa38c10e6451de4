"""Test infrastructure (NOT product code): which outcomes of the UNMODIFIED reference hinge on the last bit of its
own arithmetic?

The parity bar is bit-exact extremal counts and termination flags (BASELINE.json north_star).  That bar is only
meaningful where the reference's result is a function of the problem and not of the rounding of its libm: an
extremal that slides along the domain frame sits at x = +-1e-17 of the outermost line of the trimming cost map
(solver_ef.py:819-841, misc.py:115-139 test `a > 0` strictly), and the sign of that residue is decided by the last
bit of `np.arctan2 / np.cos / np.sin` in `directional_timeopt_control` (misc.py:72-84).  No implementation with a
different libm (CUDA's, or another numpy build: numpy dispatches sin / cos to SIMD kernels by CPU) can reproduce it.

This script PROVES the claim on the reference itself instead of asserting it.  It runs the reference from
/root/reference on every registered Resampling / Trimming case with ONE change: the value returned by the
right-hand side is moved by at most one unit in the last place (`np.nextafter`, random direction per call and
component, seeded) -

  * model `slide`: the control returned by `directional_timeopt_control` as `SolverEF.dyn_constr` calls it
                   (solver_ef.py:361-365, misc.py:72-84) - exactly what another arctan2 / cos / sin changes;
  * model `free` : `SolverEF.dyn_augsys_*` (solver_ef.py:355-359), the free augmented dynamics - what another
                   summation order or a fused multiply-add changes; it moves the obstacle entry point (the root of
                   the event on the dense output) by an ulp, and in `trap`, whose saddle flow amplifies one ulp to
                   1e-6 within the horizon, the free extremals themselves

  * model `entry`: the state `res.sol(t_enter_obs)` at which an extremal enters an obstacle (solver_ef.py:518): the
                   dense-output polynomial evaluated at the event root - on the frame it is a cancellation residue of
                   1e-17 whose sign another evaluation order of the same polynomial changes

(all three models in every run) - and records, per case, the sites whose record (tests/parity_util.py fields) differs from the unperturbed golden run
in at least one of the seeded runs: the *tie set*.  `tests/golden/tie_sets.json` holds these sets; the parity tests
compare every site OUTSIDE its case's tie set exactly and require the sites on which the CUDA path differs from the
golden to be a subset of the tie set.  A case whose tie set is empty is compared exactly everywhere.

Usage (container only, /root/reference must be mounted):

    python -W ignore oracle/tie_sensitivity.py [-j 6] [--seeds 8] [--first-seed 0] [case_key ...]
"""
import json
import os
import sys
import time
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'tie_sets.json')


def _nudger(seed):
    rng = np.random.default_rng(seed)

    def nudge(v):
        v = np.asarray(v, dtype=float)
        k = rng.integers(-1, 2, size=v.shape)
        out = np.where(k > 0, np.nextafter(v, np.inf), np.where(k < 0, np.nextafter(v, -np.inf), v))
        return out
    return nudge


def run_perturbed(job):
    key, seed = job
    import contextlib
    import io
    from cases import CASES, build_problem, SOLVERS, solver_args
    from dump import solver_record, differing_sites
    from gen_golden import reference_api
    api = reference_api()
    sef = api.sef
    case = CASES[key]
    t0 = time.perf_counter()
    nudge = _nudger(max(seed, 0))
    orig_constr = sef.directional_timeopt_control
    orig_cart, orig_gcs = sef.SolverEF.dyn_augsys_cartesian, sef.SolverEF.dyn_augsys_gcs

    def dyn_constr(*a, **k):
        return nudge(orig_constr(*a, **k))

    def aug_cart(self, *a, **k):
        return nudge(orig_cart(self, *a, **k))

    def aug_gcs(self, *a, **k):
        return nudge(orig_gcs(self, *a, **k))

    orig_ivp = sef.scitg.solve_ivp

    class NudgedSol:
        def __init__(self, sol):
            self.sol = sol

        def __call__(self, t):
            return nudge(self.sol(t))

    def solve_ivp(*a, **k):
        res = orig_ivp(*a, **k)
        if getattr(res, 'sol', None) is not None:
            res.sol = NudgedSol(res.sol)
        return res

    try:
        if seed >= 0:  # seed < 0: the unperturbed reference (must reproduce the golden at 0.0)
            sef.directional_timeopt_control = dyn_constr
            sef.SolverEF.dyn_augsys_cartesian = aug_cart
            sef.SolverEF.dyn_augsys_gcs = aug_gcs
            sef.scitg.solve_ivp = solve_ivp
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            pb = build_problem(case, api)
            cls = getattr(api, SOLVERS[case.solver])
            args, kwargs = solver_args(case, pb)
            solver = cls(*args, **kwargs)
            solver.solve()
        rec = solver_record(solver)
    finally:
        sef.directional_timeopt_control = orig_constr
        sef.SolverEF.dyn_augsys_cartesian, sef.SolverEF.dyn_augsys_gcs = orig_cart, orig_gcs
        sef.scitg.solve_ivp = orig_ivp
    g = np.load(os.path.join(ROOT, 'tests', 'golden', key + '.npz'))
    diff = sorted(differing_sites(rec, g))
    gc, rc = float(g['solution_cost']), float(rec['solution_cost'])
    cost_same = (np.isnan(gc) and np.isnan(rc)) or abs(gc - rc) <= 1e-9 * max(1., abs(gc))
    return key, seed, diff, int(rec['n_sites']), int(g['n_sites']), bool(cost_same), time.perf_counter() - t0


def main():
    from cases import CASES
    argv = sys.argv[1:]
    jobs, seeds, first_seed = 6, 8, 0
    while argv and argv[0] in ('-j', '--seeds', '--first-seed'):
        if argv[0] == '-j':
            jobs = int(argv[1])
        elif argv[0] == '--first-seed':  # more seeds for cases already in the file: the sets are merged
            first_seed = int(argv[1])
        else:
            seeds = int(argv[1])
        argv = argv[2:]
    keys = argv if argv else [k for k, c in CASES.items() if c.solver in ('R', 'T')]
    work = [(k, s) for s in ([-1] if first_seed == 0 else []) + list(range(first_seed, first_seed + seeds)) for k in keys]
    result = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            result = json.load(f).get('cases', {})
    fresh = {k: {'sites': set(), 'runs': 0, 'runs_differing': 0, 'unperturbed_reproduces_golden': None,
                 'arrival_cost_stable': True, 'site_count_stable': True,
                 'model': 'free+slide+entry'} for k in keys}
    with Pool(jobs) as pool:
        for key, seed, diff, n, ng, cost_same, dt in pool.imap_unordered(run_perturbed, work):
            r = fresh[key]
            if seed < 0:
                r['unperturbed_reproduces_golden'] = len(diff) == 0 and cost_same
            else:
                r['runs'] += 1
                r['runs_differing'] += 1 if diff else 0
                r['sites'].update(diff)
                r['arrival_cost_stable'] = r['arrival_cost_stable'] and cost_same
                r['site_count_stable'] = r['site_count_stable'] and n == ng
            print('%-32s seed %2d: %3d of %3d sites differ, cost %s (%.0f s)' % (
                key, seed, len(diff), ng, 'same' if cost_same else 'DIFFERS', dt), flush=True)
    for k, r in fresh.items():
        old = result.get(k)
        if first_seed > 0 and old:  # merge with the seeds already recorded
            r['sites'].update(old['sites'])
            r['runs'] += old['runs']
            r['runs_differing'] += old['runs_differing']
            r['unperturbed_reproduces_golden'] = old['unperturbed_reproduces_golden']
            r['arrival_cost_stable'] = r['arrival_cost_stable'] and old['arrival_cost_stable']
            r['site_count_stable'] = r['site_count_stable'] and old['site_count_stable']
        r['sites'] = sorted(r['sites'])
        result[k] = r
    doc = {
        'what': 'per case: sites of the UNMODIFIED reference whose record changes when its right-hand side is moved by '
                '<= 1 ulp (oracle/tie_sensitivity.py); parity tests compare every other site exactly',
        'perturbation': 'np.nextafter on every component returned by directional_timeopt_control inside SolverEF.dyn_constr (model slide) and by '
                        'SolverEF.dyn_augsys_* (model free) and by res.sol(t_enter_obs) (model entry); random direction in '
                        '{-1, 0, +1} ulp, numpy default_rng(seed)',
        'numpy': np.__version__,
        'cases': {k: result[k] for k in sorted(result)},
    }
    with open(OUT, 'w') as f:
        json.dump(doc, f, indent=1)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
