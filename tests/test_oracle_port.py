"""Pins the CPU oracle (oracle/solver_port.py, numpy + scipy restatement of dabry/solver_ef.py) on the golden
vectors written by the UNMODIFIED reference (oracle/gen_golden.py): counts, names, flags bit-exact and every float
within 1e-12 (observed: exactly 0.0 on all 26 cases with scipy 1.18.1 / numpy 2.3.5).

The full sweep takes ~8 CPU-minutes; by default a subset covering planar / spherical, analytic / discrete, obstacle,
penalty, resampling and trimming runs; DABRY_ORACLE_ALL=1 runs everything.  When /root/reference is mounted
(build container) the port is also run on the REFERENCE's own problem classes."""
import os

import numpy as np
import pytest

import parity_util as pu

FAST = ['linear_R', 'obstacle_R', 'chertovskih_T', 'big_rankine_T', 'linear_T', 'era5_like_R', 'spherical_geodesic_R']
ALL = pu.golden_keys('R') + pu.golden_keys('T')
CASES_TO_RUN = ALL if os.environ.get('DABRY_ORACLE_ALL') else FAST


@pytest.mark.parametrize('key', CASES_TO_RUN)
def test_port_matches_reference_golden(key):
    case, port = pu.solve_case_port(key)
    g = pu.golden(key)
    rec = pu.port_record(port)
    pu.compare_sites_record(rec, g, tol=1e-12)
    if 'traj_states' in g.files:
        pu.compare_full_trajectories(rec, g, tol=1e-12)
    # the metric's unit: RK step attempts = (nfev - 2) / 6 per solve_ivp call, as counted on the reference
    assert port.n_rk == int(g['ref_rk_steps'])
    assert port.n_ivp == int(g['ref_ivp_calls'])


def test_port_on_reference_classes():
    """Container only: the port driven by the reference's own NavigationProblem gives the golden result too, i.e.
    the restated solver loops - not just this package's field classes - reproduce the reference."""
    import refshim
    if not refshim.reference_available():
        pytest.skip('reference not mounted (GPU box)')
    from types import SimpleNamespace
    refshim.load_reference()
    from dabry.problem import NavigationProblem
    from dabry.flowfield import DiscreteFF
    from dabry.misc import Coords
    from dabry.obstacle import CircleObs
    from dabry.penalty import CirclePenalty
    api = SimpleNamespace(NavigationProblem=NavigationProblem, DiscreteFF=DiscreteFF, Coords=Coords,
                          CircleObs=CircleObs, CirclePenalty=CirclePenalty)
    case, port = pu.solve_case_port('obstacle_T', api)
    pu.compare_sites_record(pu.port_record(port), pu.golden('obstacle_T'), tol=1e-12)


def test_simple_fan_port():
    """SolverEFSimple fan (solver_ef.py:419-437) through the port vs golden."""
    from cases import CASES, build_problem, solver_args
    from solver_port import Port
    key = 'planar_waves_S'
    case = CASES[key]
    pb = build_problem(case, pu.package_api())
    args, kwargs = solver_args(case, pb)
    port = Port(*args, **kwargs)
    results = port.run_simple(0)
    g = pu.golden(key)
    assert np.array_equal(np.array([len(r.t) for r in results]), g['traj_len'])
    states = np.concatenate([r.y[:2].T for r in results])
    assert pu.scaled_diff(states, g['traj_states']) <= 1e-12
