"""Test infrastructure: the registry of parity cases shared by ``oracle/gen_golden.py`` (which runs the
unmodified reference on them, in the container) and ``tests/`` (which run this package on them).

A case is (problem builder, solver class name, solver kwargs).  ``build_problem(case, api)`` builds the
*unscaled-then-rescaled* problem with whichever API module set is passed: the reference's
(``dabry.problem`` etc.) or this package's drop-in classes - they are source compatible.
"""
import csv
import os
from types import SimpleNamespace

import numpy as np

# the 11 cases the reference's CI runs (problems.csv, gh_wf == True; tests/test_solverefr.py:8-19)
CI_CASES = ['linear', 'gyre', 'gyre_rhoads2010', 'three_vortices', 'chertovskih', 'moving_vortex',
            'big_rankine', 'obstacle', 'spherical_geodesic', 'trap', 'point_symmetric_techy2011']
# time_bound column of problems.csv for those cases (passed as total_duration, tests/test_solverefr.py:12-15)
TIME_BOUND = {'gyre_rhoads2010': 1., 'big_rankine': 1., 'trap': 2.}

SOLVERS = {'R': 'SolverEFResampling', 'T': 'SolverEFTrimming', 'S': 'SolverEFSimple'}


def _case(key, kind, solver, full=False, **kw):
    return SimpleNamespace(key=key, kind=kind, solver=solver, full=full, **kw)


CASES = {}
for _n in CI_CASES:
    CASES[f'{_n}_R'] = _case(f'{_n}_R', 'catalog', 'R', name=_n, solver_kwargs={},
                             full=_n in ('obstacle', 'linear', 'spherical_geodesic'))
    CASES[f'{_n}_T'] = _case(f'{_n}_T', 'catalog', 'T', name=_n, solver_kwargs={},
                             full=_n in ('obstacle',))
CASES['moving_vortex_S'] = _case('moving_vortex_S', 'catalog', 'S', name='moving_vortex', solver_kwargs={},
                                 n_steps=2, full=True)
CASES['planar_waves_S'] = _case('planar_waves_S', 'synthetic', 'S', gen='planar_waves', gen_args=(12, 90, 46),
                                solver_kwargs={}, n_steps=1, full=True)
CASES['planar_waves_R'] = _case('planar_waves_R', 'synthetic', 'R', gen='planar_waves', gen_args=(12, 90, 46),
                                solver_kwargs=dict(max_depth=8))
CASES['era5_like_S'] = _case('era5_like_S', 'synthetic', 'S', gen='era5_like', gen_args=(12, 90, 46),
                             solver_kwargs={}, n_steps=1, full=True)
CASES['era5_like_R'] = _case('era5_like_R', 'synthetic', 'R', gen='era5_like', gen_args=(12, 90, 46),
                             solver_kwargs=dict(max_depth=6))
CASES['regional_constrained_R'] = _case('regional_constrained_R', 'synthetic', 'R', gen='regional_constrained',
                                        gen_args=(9, 91, 51), solver_kwargs=dict(max_depth=6))
CASES['regional_discrete_penalty_R'] = _case('regional_discrete_penalty_R', 'synthetic', 'R',
                                            gen='regional_discrete_penalty', gen_args=(9, 91, 51),
                                            solver_kwargs=dict(max_depth=6))
CASES['regional_constrained_T'] = _case('regional_constrained_T', 'synthetic', 'T', gen='regional_constrained',
                                        gen_args=(9, 91, 51), solver_kwargs=dict(max_depth=8))


# Fields of the catalogue that no CI case of the reference uses (flowfield.py:906-973): a radial Gauss outflow whose
# centre, radius and strength are tabulated in time, over a uniform cross flow.  `api.flowfield` is the flowfield module
# of the reference or of this package.
def _radial_gauss_t_problem(api):
    F = api.flowfield
    nt = 6
    center = np.column_stack((np.linspace(0.35, 0.65, nt), np.linspace(0.7, 0.3, nt)))
    ff = F.RadialGaussFFT(center, np.linspace(0.15, 0.3, nt), np.linspace(0.5, 0.7, nt), np.linspace(0.3, 0.6, nt),
                          t_end=1.2, t_start=0.1) + F.UniformFF(np.array((0.1, -0.15)))
    return api.NavigationProblem(ff, np.array((0.1, 0.5)), np.array((0.9, 0.5)), 1., bl=np.array((0., 0.)),
                                 tr=np.array((1., 1.)), name='radial_gauss_t')


ANALYTIC = {'radial_gauss_t': _radial_gauss_t_problem}
CASES['radial_gauss_t_R'] = _case('radial_gauss_t_R', 'analytic', 'R', build='radial_gauss_t',
                                  solver_kwargs=dict(max_depth=6), total_duration=1.5)
CASES['radial_gauss_t_T'] = _case('radial_gauss_t_T', 'analytic', 'T', build='radial_gauss_t',
                                  solver_kwargs=dict(max_depth=6), total_duration=1.5)


def build_problem(case, api):
    """``api`` needs: NavigationProblem, DiscreteFF, Coords, CircleObs, CirclePenalty (DiscretePenalty for the case that
    uses one)."""
    if case.kind == 'catalog':
        return api.NavigationProblem.from_name(case.name).rescale()
    if case.kind == 'analytic':
        return ANALYTIC[case.build](api).rescale()
    from dabry_b200 import synthetic
    sc = getattr(synthetic, case.gen)(*case.gen_args)
    coords = api.Coords.GCS if sc.coords == 'gcs' else api.Coords.CARTESIAN
    ff = api.DiscreteFF(sc.values, sc.bounds, coords)
    obstacles = [api.CircleObs(np.array(c), float(r)) for c, r in sc.circle_obstacles]
    penalty = None
    if sc.circle_penalty is not None:
        c, r, a = sc.circle_penalty
        penalty = api.CirclePenalty(np.array(c), float(r), float(a))
    if getattr(sc, 'discrete_penalty', None) is not None:
        penalty = api.DiscretePenalty(*sc.discrete_penalty)
    pb = api.NavigationProblem(ff, sc.x_init, sc.x_target, sc.srf, bl=sc.bl, tr=sc.tr,
                               obstacles=obstacles, penalty=penalty, name=sc.name)
    return pb.rescale()


def solver_args(case, pb):
    """Positional/keyword args for the solver constructor (total_duration second, as the reference tests do)."""
    kwargs = dict(case.solver_kwargs)
    if case.kind == 'catalog' and case.name in TIME_BOUND:
        return (pb, TIME_BOUND[case.name]), kwargs
    if case.kind == 'analytic':
        # the horizon in the rescaled clock of the problem (rescale() divides times by length_reference / srf)
        return (pb, case.total_duration), kwargs
    if case.kind == 'synthetic':
        from dabry_b200 import synthetic
        sc_td = getattr(synthetic, case.gen)(2, 4, 4).total_duration
        if sc_td is not None:
            return (pb, sc_td), kwargs
    return (pb,), kwargs


def run_solver(case, pb, api):
    """Instantiate + run.  SolverEFSimple has no solve(): setup() then n_steps x step() (solver_ef.py:419-437)."""
    cls = getattr(api, SOLVERS[case.solver])
    args, kwargs = solver_args(case, pb)
    solver = cls(*args, **kwargs)
    if case.solver == 'S':
        solver.setup()
        for _ in range(case.n_steps):
            solver.step()
    else:
        solver.solve()
    return solver
