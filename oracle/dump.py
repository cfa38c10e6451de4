"""Test infrastructure: flatten a solved extremal-field solver (the reference's, the oracle's, or this
package's drop-in classes - they expose the same attributes) into a dict of numpy arrays.

The layout is what ``tests/golden/*.npz`` hold and what the parity tests compare:

* per site, in ``solver.sites`` insertion order (solver_ef.py:470, 663): name, index_t_init, trajectory length,
  closure / neutering flags (-1 = None), obstacle event index (-1 = free), index_t_obs, obs_trigo, t_enter_obs,
  index_t_check_next, first and last (state, costate/control) samples, a position checksum;
* optional ragged trajectories (``full=True``): times, states, costates, controls, cost with offsets;
* the arrival record: solution site name, its in-target minimum cost (solver_ef.py:675-677), success flag,
  the set of solution sites.
"""
import numpy as np


def _enum_val(e):
    return -1 if e is None else int(e.value)


def _arr2(a, n):
    if a is None:
        return np.full((n, 2), np.nan)
    a = np.asarray(a, dtype=float)
    if a.size == 0:
        return np.full((n, 2), np.nan)
    return a.reshape(n, 2)


def solver_record(solver, full: bool = False, next_nb: bool = False) -> dict:
    sites = list(solver.sites.values())
    ev_names = list(solver.events.keys())
    n = len(sites)
    rec = {
        'n_sites': np.int64(n),
        'names': np.array([s.name for s in sites]),
        'event_names': np.array(ev_names),
        'index_t_init': np.array([s.index_t_init for s in sites], dtype=np.int32),
        'traj_len': np.array([len(s.traj) for s in sites], dtype=np.int32),
        'closure': np.array([_enum_val(s.closure_reason) for s in sites], dtype=np.int32),
        'neuter': np.array([_enum_val(s.neutering_reason) for s in sites], dtype=np.int32),
        'obstacle': np.array([ev_names.index(s.obstacle_name) if s.obstacle_name else -1 for s in sites],
                             dtype=np.int32),
        'index_t_obs': np.array([s.index_t_obs for s in sites], dtype=np.int32),
        'obs_trigo': np.array([bool(s.obs_trigo) for s in sites]),
        't_enter_obs': np.array([np.nan if s.t_enter_obs is None else s.t_enter_obs for s in sites]),
        'index_t_check_next': np.array([s.index_t_check_next for s in sites], dtype=np.int32),
        'times': np.asarray(solver.times, dtype=float),
        'total_duration': np.float64(solver.total_duration),
        'max_int_step': np.float64(solver.max_int_step),
        'success': np.bool_(solver.success),
    }
    first = np.zeros((n, 6))
    last = np.zeros((n, 6))
    csum = np.zeros((n, 2))
    for i, s in enumerate(sites):
        m = len(s.traj)
        co = _arr2(s.traj.costates, m)
        cu = _arr2(s.traj.controls, m)
        st = np.asarray(s.traj.states, dtype=float).reshape(m, 2)
        first[i] = np.concatenate((st[0], co[0], cu[0]))
        last[i] = np.concatenate((st[-1], co[-1], cu[-1]))
        csum[i] = st.sum(axis=0)
    rec['first'] = first
    rec['last'] = last
    rec['state_sum'] = csum
    sol = solver.solution_site
    if sol is not None:
        tr2 = solver._target_radius_sq if hasattr(solver, '_target_radius_sq') else solver.target_radius ** 2
        mask = np.sum(np.square(sol.traj.states - solver.pb.x_target), axis=-1) < tr2
        rec['solution_name'] = np.array(sol.name)
        rec['solution_cost'] = np.float64(np.min(sol.traj.cost[mask]))
        rec['solution_names'] = np.array(sorted(s.name for s in solver.solution_sites))
        if getattr(sol, 'traj_full', None) is not None:
            tf = sol.traj_full
            rec['solution_full_times'] = np.asarray(tf.times)
            rec['solution_full_states'] = np.asarray(tf.states)
    else:
        rec['solution_name'] = np.array('')
        rec['solution_cost'] = np.float64(np.nan)
        rec['solution_names'] = np.array([], dtype='U1')
    if full:
        offs = np.zeros(n + 1, dtype=np.int64)
        offs[1:] = np.cumsum(rec['traj_len'])
        rec['traj_offsets'] = offs
        rec['traj_times'] = np.concatenate([np.asarray(s.traj.times, dtype=float) for s in sites])
        rec['traj_states'] = np.concatenate([np.asarray(s.traj.states, dtype=float).reshape(-1, 2) for s in sites])
        rec['traj_costates'] = np.concatenate([_arr2(s.traj.costates, len(s.traj)) for s in sites])
        rec['traj_controls'] = np.concatenate([_arr2(s.traj.controls, len(s.traj)) for s in sites])
        rec['traj_cost'] = np.concatenate([np.asarray(s.traj.cost, dtype=float) for s in sites])
    if next_nb:
        ids = {id(s): i for i, s in enumerate(sites)}
        nb = np.full((n, len(solver.times)), -1, dtype=np.int32)
        for i, s in enumerate(sites):
            for k, o in enumerate(s.next_nb):
                if o is not None:
                    nb[i, k] = ids[id(o)]
        rec['next_nb'] = nb
    return rec


EXACT_SITE_FIELDS = ['index_t_init', 'traj_len', 'closure', 'neuter', 'obstacle', 'index_t_obs', 'obs_trigo',
                     'index_t_check_next']
FLOAT_SITE_FIELDS = ['first', 'last', 'state_sum', 't_enter_obs']


def differing_sites(rec, g, tol=1e-9):
    """Names of the sites whose record differs between two flattened runs, or that exist in only one of them.
    Integer fields and flags must be equal; floating-point fields may differ by tol * max(1, |reference value|) per
    element (costates reach 1e17 on captured extremals, so the bar is relative to the element, not to the column)."""
    a = {n: i for i, n in enumerate(rec['names'])}
    b = {n: i for i, n in enumerate(g['names'])}
    out = set(a) ^ set(b)
    for n in set(a) & set(b):
        i, j = a[n], b[n]
        bad = any(rec[f][i] != g[f][j] for f in EXACT_SITE_FIELDS)
        for f in FLOAT_SITE_FIELDS:
            x = np.atleast_1d(np.asarray(rec[f][i], dtype=float))
            y = np.atleast_1d(np.asarray(g[f][j], dtype=float))
            if not np.array_equal(np.isnan(x), np.isnan(y)):
                bad = True
                continue
            d = np.abs(x - y)
            d[np.isnan(d)] = 0.
            if np.any(d > tol * np.maximum(1., np.abs(np.where(np.isnan(y), 0., y)))):
                bad = True
        if bad:
            out.add(str(n))
    return {str(n) for n in out}
