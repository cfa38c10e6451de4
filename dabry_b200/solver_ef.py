"""Extremal-field solvers of the drop-in API (reference: dabry/solver_ef.py).

`SolverEFSimple`, `SolverEFResampling` and `SolverEFTrimming` keep the reference's constructor signatures,
methods (`setup / step / solve / save_results / cost_map_triangle`) and result attributes (`success`, `sites`,
`solution_site`, `solution_sites`, `trajs`, `times`, ...).  All numerical work - integration of the augmented
state-costate system, events, obstacle following, front resampling, trimming, arrival-time selection - runs on the
GPU through libdabry_b200.so; this module prepares the constants (`SolverEF.__init__`, solver_ef.py:284-339), lowers
the problem, and turns the device arrays back into `Site` / `Trajectory` objects.
"""
import ctypes as C
import math
from abc import ABC
from enum import Enum
from typing import Optional, Dict

import numpy as np
from numpy import ndarray

from dabry_b200 import _native as nat
from dabry_b200 import _lowering as low
from dabry_b200.misc import Coords, to_alpha, alpha_to_int, diadic_valuation, non_terminal, terminal, \
    directional_timeopt_control, triangle_mask_and_cost
from dabry_b200.problem import NavigationProblem
from dabry_b200.trajectory import Trajectory


class ClosureReason(Enum):  # solver_ef.py:42-48
    TERMINAL_INTEGRATION = 0
    WITHIN_OBS = 1
    SUBOPTIMAL = 2
    IMPOSSIBLE_OBS_TRACKING = 3


class NeuteringReason(Enum):  # solver_ef.py:51-53
    SELF_AND_NB_IN_OBS = 0


class SiteManager:
    """Dyadic addressing of extremals (solver_ef.py:167-259): root `prefix` of the fan sits at index
    prefix * 2^max_depth; a child of depth d and location l at prefix * 2^max_depth + 2^(max_depth - d) (2 l + 1).
    Names are `PREFIX[-depth-location]` with the prefix spelled in base 26."""

    def __init__(self, n_sectors: int, max_depth: int, looping_sectors=True):
        self.n_sectors = n_sectors
        self.max_depth = max_depth
        self.looping_sectors = looping_sectors
        self._span = 1 << max_depth
        self._n_prefix_chars = int(np.ceil(math.log(n_sectors) / math.log(26))) if n_sectors > 1 else 0

    @property
    def n_total_sites(self):
        return self.n_sectors * self._span + (0 if self.looping_sectors else 1)

    def check_pdl(self, prefix: int, depth: int, location: int):
        n_prefix = self.n_sectors if self.looping_sectors else self.n_sectors + 1
        if not 0 <= prefix < n_prefix:
            raise IndexError(f'"prefix" out of range. (val: {prefix}, range: (0, {n_prefix}))')
        if not 0 <= depth <= self.max_depth:
            raise IndexError(f'"depth" out of range. (val: {depth}, range: (0, {self.max_depth + 1}))')
        if not 0 <= location < max(self._span // 2, 1):
            raise IndexError(f'"location" out of range. (val: {location}, range: (0, {self._span // 2}))')

    def index_from_pdl(self, prefix: int, depth: int, location: int):
        self.check_pdl(prefix, depth, location)
        if depth == 0 and location == 0:
            return self._span * prefix
        return self._span * prefix + (1 << (self.max_depth - depth)) * (2 * location + 1)

    def pdl_from_index(self, index: int):
        prefix, u = divmod(int(index), self._span)
        if u == 0:
            return prefix, 0, 0
        depth = self.max_depth - diadic_valuation(u)
        return prefix, depth, ((u >> (self.max_depth - depth)) - 1) // 2

    def name_from_pdl(self, prefix: int, depth: int, location: int):
        self.check_pdl(prefix, depth, location)
        head = to_alpha(prefix).rjust(self._n_prefix_chars, 'A')
        if depth == 0 and location == 0:
            return head
        return f'{head}-{depth}-{location}'

    def name_from_index(self, index: int):
        return self.name_from_pdl(*self.pdl_from_index(index))

    @staticmethod
    def pdl_from_name(name: str):
        if '-' not in name:
            return alpha_to_int(name), 0, 0
        head, depth, location = name.split('-')
        return alpha_to_int(head), int(depth), int(location)

    def index_from_name(self, name: str):
        return self.index_from_pdl(*self.pdl_from_name(name))

    def name_from_parents_name(self, name_prev: str, name_next: str):
        i_prev, i_next = self.index_from_name(name_prev), self.index_from_name(name_next)
        if i_next < i_prev:
            if not self.looping_sectors:
                raise ValueError('Next parent index is inferior to previous parent in non-looping mode')
            i_next += self.n_total_sites
        if self.max_depth - 1 in (self.pdl_from_name(name_prev)[1], self.pdl_from_name(name_next)[1]):
            raise Exception('Maximum depth reached: cannot create child site')
        if (i_prev + i_next) % 2:
            raise AssertionError('parents are not dyadic neighbours')
        return self.name_from_index((i_prev + i_next) // 2)

    def parents_name_from_name(self, name: str):
        index = self.index_from_name(name)
        _, depth, location = self.pdl_from_name(name)
        if depth == 0 and location == 0:
            return None, None
        half = 1 << (self.max_depth - depth)
        return self.name_from_index(index - half), self.name_from_index((index + half) % self.n_total_sites)


class Site:
    """One extremal, read back from the device (fields of solver_ef.py:56-164).  The trajectory is fetched
    on first access."""

    def __init__(self, solver, slot: int, rec, name: str):
        self._solver = solver
        self.slot = slot
        self.name = name
        self.index_t_init = int(rec['index_t_init'])
        self.t_init = float(solver.times[self.index_t_init])
        self.index_t_check_next = int(rec['index_t_check_next'])
        obstacle = int(rec['obstacle'])
        self.obstacle_name = '' if obstacle < 0 else solver._event_names[obstacle]
        self.index_t_obs = int(rec['index_t_obs'])
        t_enter = float(rec['t_enter_obs'])
        self.t_enter_obs = None if np.isnan(t_enter) else t_enter
        self.obs_trigo = bool(rec['obs_trigo'])
        closure, neuter = int(rec['closure']), int(rec['neuter'])
        self.closure_reason = None if closure < 0 else ClosureReason(closure)
        self.neutering_reason = None if neuter < 0 else NeuteringReason(neuter)
        self._index_neutered = int(rec['index_neutered'])
        self._n_samples = int(rec['n_samples'])
        self._depth = int(rec['depth'])
        self._parents = (int(rec['parent_prev']), int(rec['parent_next']))
        n_ev = min(int(rec['n_target_events']), nat.MAX_TARGET_EVENTS)
        self._target_events = np.array(rec['t_target_events'][:n_ev], dtype=float)
        self._traj = None
        self._next_nb = None
        self._traj_full: Optional[Trajectory] = None

    @property
    def traj(self) -> Trajectory:
        if self._traj is None:
            self._traj = self._solver._fetch_traj(self)
        return self._traj

    @property
    def traj_full(self) -> Optional[Trajectory]:
        """Trajectory extended back to t_init through the ancestors (solver_ef.py:701-749).  The reference builds it
        for every solution site at the end of solve(); here it is built on first access (a 10^6 fan has 10^4+
        solution sites), and stays None for the other sites like in the reference."""
        if self._traj_full is None and self._solver._is_solution_slot(self.slot):
            self._solver.extrapolate_back_traj(self)
        return self._traj_full

    @traj_full.setter
    def traj_full(self, value):
        self._traj_full = value

    @property
    def next_nb(self):
        if self._next_nb is None:
            row = self._solver._handle.next_nb(self.slot, 1)[0]
            table = self._solver._site_by_slot
            self._next_nb = [None if j < 0 else table(int(j)) for j in row]
        return self._next_nb

    @property
    def depth(self):
        return self._depth

    @property
    def closed(self):
        return self.closure_reason is not None

    @property
    def neutered(self):
        return self.neutering_reason is not None

    @property
    def has_neighbours(self):
        return any(nb is not None for nb in self.next_nb)

    @property
    def index_t(self):
        return self.index_t_init + self._n_samples - 1

    @property
    def n_time(self):
        return self._solver.n_time

    def in_obs_at(self, index: int):
        return len(self.obstacle_name) > 0 and index >= self.index_t_obs

    def time_at_index(self, index: int):
        return self.traj.times[index - self.index_t_init]

    def state_at_index(self, index: int):
        return self.traj.states[index - self.index_t_init]

    def costate_at_index(self, index: int):
        return self.traj.costates[index - self.index_t_init]

    def state_aug_at_index(self, index: int):
        return np.concatenate((self.state_at_index(index), self.costate_at_index(index)))

    def control_at_index(self, index: int):
        return self.traj.controls[index - self.index_t_init]

    def cost_at_index(self, index: int):
        return self.traj.cost[index - self.index_t_init]

    def is_root(self):
        return '-' not in self.name

    def __repr__(self):
        return f"<Site {self.name}>"

    __str__ = __repr__


_ROOT_FAN_CACHE = {}


def _root_fan(n, root_range, shard):
    """[count][2] initial costates 1000 (cos, sin)(theta_i) of the local roots (solver_ef.py:465-467) with numpy's own
    cos / sin.  The fan is a constant of (n, shard): large fans are computed by a few threads (numpy releases the GIL)
    and the last one is kept, so that a sequence of solves with the same fan pays for it once."""
    key = (n, root_range, shard)
    hit = _ROOT_FAN_CACHE.get(key)
    if hit is not None:
        return hit
    offset, count = (0, n) if root_range is None else root_range
    j = np.arange(count)
    if shard is not None:
        _, world, block = shard
        j = (j // block) * (world * block) + j % block
    theta = np.linspace(0, 2 * np.pi, n, endpoint=False)[(offset + j) % n]  # the library validates the range
    out = np.empty((count, 2))

    def fill(lo, hi):
        np.cos(theta[lo:hi], out=out[lo:hi, 0])
        np.sin(theta[lo:hi], out=out[lo:hi, 1])
        out[lo:hi] *= 1000

    if count >= 1 << 16:
        from concurrent.futures import ThreadPoolExecutor
        import os
        workers = max(1, min(8, os.cpu_count() or 1))
        edges = np.linspace(0, count, workers + 1).astype(int)
        with ThreadPoolExecutor(workers) as pool:
            list(pool.map(lambda k: fill(edges[k], edges[k + 1]), range(workers)))
    else:
        fill(0, count)
    if count >= 1 << 16:  # keep only the last large fan (16 B per root), page-locked for its uploads
        _ROOT_FAN_CACHE.clear()
        _ROOT_FAN_CACHE[key] = out
        try:
            _ROOT_FAN_CACHE['pin'] = nat.PinnedArray(out)
        except Exception:
            pass
    return out


class SolverEF(ABC):
    _ALL_MODES = ['time', 'energy']
    _VARIANT = nat.RESAMPLING

    def __init__(self, pb: NavigationProblem, total_duration: Optional[float] = None, mode: str = 'time',
                 max_depth: int = 20, n_time: int = 100, n_costate_sectors: int = 30,
                 t_init: Optional[float] = None, n_costate_norm: int = 10, costate_norm_bounds=(0., 1.),
                 target_radius: Optional[float] = None, abs_max_step: Optional[float] = None,
                 rel_max_step: Optional[float] = 0.01, *, device: int = 0, max_sites: int = 0,
                 root_range: Optional[tuple] = None, shard: Optional[tuple] = None, integrator: str = 'dopri5',
                 dtype: str = 'f64', comm=None):
        if mode not in self._ALL_MODES:
            raise ValueError('Mode %s is not defined' % mode)
        if mode == 'energy':
            raise ValueError('Mode "energy" is not implemented yet')
        self.pb = pb
        if total_duration is None:
            # 1.1 x the faster of the two feedback heuristics; 2 L / v when neither reaches the target (:302-309)
            self.total_duration = 1.1 * self.pb.auto_time_upper_bound(device)
            if self.total_duration == np.inf:
                self.total_duration = 2 * self.pb.length_reference / self.pb.srf_max
        else:
            self.total_duration = total_duration
        self.t_init = t_init if t_init is not None else self.pb.model.ff.t_start
        self.target_radius = target_radius if target_radius is not None else pb.target_radius
        self._target_radius_sq = self.target_radius ** 2
        self.times: ndarray = np.linspace(self.t_init, self.t_upper_bound, n_time)
        self.costate_norm_bounds = costate_norm_bounds
        self.n_costate_sectors = n_costate_sectors
        self.n_costate_norm = n_costate_norm
        self.max_depth = max_depth
        self.success = False
        self.depth = 0
        self.traj_groups = []
        # event table: target first, then the obstacles in problem order, the domain frame last (:322-334)
        self.events = {'target': self._event_target}
        self.obstacles = {}
        seen = {}
        for obs in self.pb.obstacles:
            cls_name = type(obs).__name__
            full_name = 'obs_%s_%d' % (cls_name, seen.get(cls_name, 0))
            seen[cls_name] = seen.get(cls_name, 0) + 1
            self.events[full_name] = obs.event
            self.obstacles[full_name] = obs
        abs_max_step = self.total_duration / 2 if abs_max_step is None else abs_max_step
        self.max_int_step = abs_max_step if rel_max_step is None else min(abs_max_step,
                                                                          rel_max_step * self.total_duration)
        self.dyn_augsys = self.dyn_augsys_cartesian if self.pb.coords == Coords.CARTESIAN else self.dyn_augsys_gcs
        # native side.  integrator: 'dopri5' = scipy's RK45 restated (the reference's integrator), 'rk4' = classic
        # fixed-step RK4 with h = max_int_step.  dtype: 'f64', or 'f32' = flow grid stored in fp32 (tolerance 1e-4)
        if integrator not in ('dopri5', 'rk4'):
            raise ValueError('integrator must be "dopri5" or "rk4"')
        if dtype not in ('f64', 'f32'):
            raise ValueError('dtype must be "f64" or "f32"')
        self.integrator = integrator
        self.dtype = dtype
        self.device = device
        self.max_sites = max_sites
        self.root_range = root_range
        # shard = (rank, world, block): block-cyclic theta_0 sharding - this solver owns blocks of `block` roots, every
        # world-th block of the ring starting with block `rank` (balances the load between ranks); root_range =
        # (offset, count) is the contiguous alternative
        self.shard = shard
        if shard is not None:
            rank, world, block = shard
            if n_costate_sectors % (world * block) != 0:
                raise ValueError('n_costate_sectors must be a multiple of world * block')
            self.root_range = (rank * block, n_costate_sectors // world)
        # comm: the library's communicator over the GPUs of a sharded job (dabry_b200.distributed.library_comm); with it
        # solve() runs dabry_solve_sharded and the flow grid is uploaded cooperatively (1 / world per rank + all-gather)
        self._comm = comm
        self._event_names = list(self.events.keys())
        self._handle: Optional[nat.Handle] = None
        self._stale = True

    # ---- reference-compatible helpers (host numpy; the device runs csrc/model.cuh) -----------------------------
    @property
    def t_upper_bound(self):
        return self.t_init + self.total_duration

    @property
    def n_time(self):
        return self.times.shape[0]

    def dyn_augsys_cartesian(self, t: float, y: ndarray):
        return self.pb.augsys_dyn_timeopt_cartesian(t, y[:2], y[2:])

    def dyn_augsys_gcs(self, t: float, y: ndarray):
        return self.pb.augsys_dyn_timeopt_gcs(t, y[:2], y[2:])

    def dyn_constr(self, t: float, x: ndarray, obstacle: str, trigo: bool):
        """Ground velocity when sliding along an obstacle boundary (solver_ef.py:361-365)."""
        sign = 2. * trigo - 1.
        d = np.array(((0., -sign), (sign, 0.))) @ self.obstacles[obstacle].d_value(x)
        ff_val = self.pb.model.ff.value(t, x)
        return ff_val + directional_timeopt_control(ff_val, d, self.pb.srf_max)

    @non_terminal
    def _event_target(self, _, x):
        return np.sum(np.square(x[:2] - self.pb.x_target)) - self._target_radius_sq

    @terminal
    def _event_quit_obs(self, t, x, obstacle: str, trigo: bool):
        g = self.obstacles[obstacle].d_value(x)
        return self.pb.srf_max - np.abs(self.pb.model.ff.value(t, x) @ (g / np.linalg.norm(g)))

    def t_events_to_dict(self, t_events) -> Dict[str, ndarray]:
        return {name: t_events[i] for i, name in enumerate(self.events.keys())}

    def dyn_augsys_device(self, t, y):
        """`dyn_augsys` evaluated by the CUDA kernels, for any number of points: t (n,), y (n, 4) -> (n, 4)."""
        self._ensure_handle()
        return self._handle.eval_rhs(t, y)

    @property
    def trajs(self):
        return []

    def save_results(self):
        self.pb.io.save_trajs(self.trajs, group_name='ef_01')
        self.pb.io.save_ff(self.pb.model.ff, bl=self.pb.bl, tr=self.pb.tr)

    # ---- native plumbing -------------------------------------------------------------------------------------
    def _native_config(self, n_roots, max_depth, max_dist=0., mode_origin=True, ff_max_norm=0.,
                       n_index_per_subframe=10, trimming_band_width=3):
        cfg = nat.Config()
        cfg.abi_version = nat.ABI_VERSION
        cfg.coords = low.coords_code(self.pb.coords)
        cfg.n_time = self.n_time
        cfg.max_depth = max_depth
        cfg.mode_origin = 1 if mode_origin else 0
        cfg.integrator = 1 if self.integrator == 'rk4' else 0
        cfg.dtype = 1 if self.dtype == 'f32' else 0
        cfg.device = self.device
        cfg.n_index_per_subframe = n_index_per_subframe
        cfg.trimming_band_width = trimming_band_width
        cfg.cost_map_nx = cfg.cost_map_ny = 100
        cfg.n_roots = n_roots
        offset, count = (0, n_roots) if self.root_range is None else self.root_range
        cfg.root_offset, cfg.root_count = offset, count
        cfg.max_sites = self.max_sites
        if self.shard is not None:
            cfg.shard_block, cfg.shard_world = self.shard[2], self.shard[1]
        cfg.t_init = self.t_init
        cfg.max_step = self.max_int_step
        cfg.rtol, cfg.atol = 1e-3, 1e-6
        cfg.srf = self.pb.srf_max
        cfg.x_init[0], cfg.x_init[1] = self.pb.x_init
        cfg.x_target[0], cfg.x_target[1] = self.pb.x_target
        cfg.target_radius = self.target_radius
        cfg.max_dist = max_dist
        cfg.bl[0], cfg.bl[1] = self.pb.bl
        cfg.tr[0], cfg.tr[1] = self.pb.tr
        cfg.ff_max_norm = ff_max_norm
        self._times_c = nat.f64(self.times)
        cfg.times = nat.dptr(self._times_c)
        return cfg

    def _make_handle(self, cfg):
        handle = nat.Handle(cfg)
        if self._comm is not None:
            handle.attach_comm(self._comm)
        low.apply_flow(handle, self.pb.model.ff, shared=self._comm is not None)
        low.apply_obstacles(handle, self.pb.obstacles)
        low.apply_penalty(handle, self.pb.penalty)
        return handle

    def _ensure_handle(self):
        if self._handle is None:
            self._handle = self._make_handle(self._default_config())
        return self._handle

    def _default_config(self):
        return self._native_config(self.n_costate_sectors, self.max_depth)

    def native_stats(self) -> dict:
        """Device counters and per-phase device times of the last solve (`dabry_stats`)."""
        return self._ensure_handle().stats()

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


class SolverEFSimple(SolverEF):
    """Plain fans of extremals, one finer fan per `step()`, no resampling (solver_ef.py:406-444)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.layers = []

    def setup(self):
        self.traj_groups = []
        self.layers = []
        self.depth = 0

    def step(self):
        angles = np.linspace(0., 2 * np.pi, 2 ** self.depth * self.n_costate_sectors + 1)
        if self.depth != 0:
            angles = angles[1::2]
        costates = nat.f64(np.stack((np.cos(angles), np.sin(angles)), -1))
        cfg = self._native_config(len(angles), 0)
        cfg.root_offset, cfg.root_count = 0, len(angles)
        handle = self._make_handle(cfg)
        try:
            handle.call('dabry_solve', nat.SIMPLE, nat.dptr(costates))
            rec = handle.sites()
            data = handle.trajs(0, len(angles), want=('times', 'states', 'costates', 'cost'))
            self._last_stats = handle.stats()
        finally:
            handle.close()
        group = []
        for i in range(len(angles)):
            n = int(rec['n_samples'][i])
            events = {name: np.array(()) for name in self._event_names}
            n_ev = min(int(rec['n_target_events'][i]), nat.MAX_TARGET_EVENTS)
            events['target'] = np.array(rec['t_target_events'][i][:n_ev])
            if rec['obstacle'][i] >= 0:
                events[self._event_names[rec['obstacle'][i]]] = np.array((rec['t_enter_obs'][i],))
            traj = Trajectory.cartesian(data['times'][i, :n], data['states'][i, :n],
                                        costates=data['costates'][i, :n], events=events,
                                        cost=data['times'][i, :n] - self.t_init)
            if events['target'].shape[0] > 0:
                self.success = True
            group.append(traj)
        self.traj_groups.append(tuple(group))
        self.layers.append(tuple(angles))
        self.depth += 1

    def native_stats(self):
        return getattr(self, '_last_stats', {})

    @property
    def trajs(self):
        return [traj for group in self.traj_groups for traj in group]


class SolverEFResampling(SolverEF):
    """Extremal field with front resampling: a child extremal is inserted wherever two neighbours drift further
    apart than `max_dist` (solver_ef.py:447-754)."""
    _VARIANT = nat.RESAMPLING

    def __init__(self, *args, max_dist: Optional[float] = None, mode_origin: bool = True,
                 device_roots: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        # device_roots: generate the root costates on the GPU (CUDA sincos, within 1 ulp of numpy's) instead of
        # uploading numpy's - for very large fans
        self.device_roots = device_roots
        self.max_dist = max_dist if max_dist is not None else self.target_radius
        self._max_dist_sq = self.max_dist ** 2
        self.mode_origin = mode_origin
        self.site_mngr = SiteManager(self.n_costate_sectors, self.max_depth)
        # max |w| over the grid nodes (solver_ef.py:459-460): computed by the device while it repacks the grid
        self._flow_is_grid = isinstance(low.unwrap_flow(self.pb.model.ff)[1], low.DiscreteFF)
        self._solution_sites: Optional[set] = None
        self._solution_slots = np.zeros(0, dtype=np.int64)
        self.solution_site: Optional[Site] = None
        self._n_sites_cache = None
        self._sites: Optional[dict] = None
        self._records = None
        self._slot_cache = {}
        self._bulk = None

    def _default_config(self):
        return self._native_config(self.n_costate_sectors, self.max_depth, max_dist=self.max_dist,
                                   mode_origin=self.mode_origin,
                                   ff_max_norm=-1. if self._flow_is_grid else 0.,
                                   n_index_per_subframe=getattr(self, 'n_index_per_subframe', 10),
                                   trimming_band_width=getattr(self, '_trimming_band_width', 3))

    def _root_costates(self):
        """1000 (cos, sin)(theta), theta = linspace(0, 2 pi, n, endpoint=False) (solver_ef.py:465-467): computed
        with numpy so that the device starts from the same bits as the reference."""
        if self.device_roots:
            return None
        return _root_fan(self.n_costate_sectors, self.root_range, self.shard)

    def _ring_indices(self):
        """Ring (global) index of every local root, in slot order."""
        offset, count = (0, self.n_costate_sectors) if self.root_range is None else self.root_range
        j = np.arange(count)
        if self.shard is None:
            return offset + j
        _, world, block = self.shard
        return offset + (j // block) * (world * block) + j % block

    @property
    def _ff_max_norm(self):
        return self._ensure_handle().flow_max_norm() if self._flow_is_grid else None

    @property
    def solution_sites(self) -> set:
        if self._solution_sites is None:
            self._solution_sites = {self._site_by_slot(int(s)) for s in self._solution_slots}
        return self._solution_sites

    @solution_sites.setter
    def solution_sites(self, value):
        self._solution_sites = value

    def _is_solution_slot(self, slot):
        i = np.searchsorted(self._solution_slots, slot)
        return i < len(self._solution_slots) and self._solution_slots[i] == slot

    # ---- reference API ---------------------------------------------------------------------------------------
    def setup(self):
        self.traj_groups = []
        handle = self._ensure_handle()
        costates = self._root_costates()
        handle.call('dabry_setup', self._VARIANT, nat.dptr(costates))
        self._invalidate()
        self.depth = 0

    def step(self):
        self._ensure_handle().call('dabry_step')
        self._invalidate()
        self.depth += 1

    def solve(self):
        handle = self._ensure_handle()
        costates = self._root_costates()
        handle.call('dabry_solve_sharded' if self._comm is not None else 'dabry_solve', self._VARIANT,
                    nat.dptr(costates))
        self._invalidate()
        self._after_solve()
        self.check_solutions(_finished=True)
        if self.solution_site is not None:
            self.extrapolate_back_traj(self.solution_site)  # the other solution sites: on first access of traj_full

    def _after_solve(self):
        self.depth = self.max_depth

    def check_solutions(self, _finished=False):
        """Sites with a grid sample strictly inside the target disc; the cheapest is `solution_site`, those within
        1.15 x its cost stay in `solution_sites` (solver_ef.py:667-685).  Cost ties -> lowest insertion rank."""
        handle = self._ensure_handle()
        if not _finished:
            handle.call('dabry_finish')
        self._arrivals = None  # per-site arrays: read back on first access (12 B per site)
        slot, best = handle.solution()
        self._solution_sites = None
        self._solution_slots = np.zeros(0, dtype=np.int64)
        self.solution_site = None
        if slot >= 0:
            self.success = True
            # sites with `not cost > 1.15 * best` (solver_ef.py:681-684), compacted on the device: a few kB instead of
            # the arrival arrays of the whole front
            self._solution_slots = handle.solution_sites(1.15 * best)
            self.solution_site = self._site_by_slot(slot)

    def _fetch_arrivals(self):
        if getattr(self, '_arrivals', None) is None:
            self._arrivals = self._ensure_handle().arrivals(0, self._n_sites())
        return self._arrivals

    @property
    def _arrival_index(self):
        """Per site: grid index of the cheapest in-target sample, -1 if the site never enters the target disc."""
        return self._fetch_arrivals()[0]

    @property
    def _arrival_cost(self):
        return self._fetch_arrivals()[1]

    @property
    def solution_cost(self):
        """Arrival cost of `solution_site` (min cost over its in-target samples), NaN if none."""
        return float(self._ensure_handle().solution()[1])

    @property
    def sites(self) -> dict:
        if self._sites is None:
            n = self._n_sites()
            self._sites = {}
            records = self._all_records()
            for slot in range(n):
                if records['flags'][slot] & 1:
                    continue  # ghost copy of the next rank's first root (sharded ring)
                site = self._site_by_slot(slot)
                self._sites[site.name] = site
        return self._sites

    @property
    def trajs(self):
        return [site.traj for site in self.sites.values()]

    @property
    def suboptimal_sites(self):
        return self.solution_sites.difference({self.solution_site})

    @property
    def _closed_sites(self):
        return [site for site in self.sites.values() if site.closed]

    def cost_map_triangle(self, nx: int = 100, ny: int = 100):
        if (nx, ny) != (100, 100):
            raise ValueError('the device cost map is 100 x 100 (solver_ef.py:820)')
        return self._ensure_handle().cost_map(nx, ny, full=True)

    def extrapolate_back_traj(self, site: Site):
        """Trajectory of a child extended back to t_init: before its birth index it is the convex combination of
        its ancestors, the weight of the more recent ancestor halving at each generation (solver_ef.py:701-749)."""
        def padded(arr, n):
            return np.full((n, 2), np.nan) if arr is None else arr

        times, states, cost = site.traj.times.copy(), site.traj.states.copy(), site.traj.cost.copy()
        costates = padded(site.traj.costates, len(times)).copy()
        controls = padded(site.traj.controls, len(times)).copy()
        prev_slot, next_slot = site._parents
        w_prev = w_next = 0.5
        index_hi = site.index_t_init - 1
        while prev_slot >= 0 and next_slot >= 0:
            a, b = self._site_by_slot(prev_slot), self._site_by_slot(next_slot)
            index_lo = max(a.index_t_init, b.index_t_init)
            a_is_younger = a.index_t_init > b.index_t_init
            sa = slice(index_lo - a.index_t_init, index_hi - a.index_t_init + 1)
            sb = slice(index_lo - b.index_t_init, index_hi - b.index_t_init + 1)
            m = sa.stop - sa.start
            times = np.concatenate((a.traj.times[sa], times))
            states = np.concatenate((w_prev * a.traj.states[sa] + w_next * b.traj.states[sb], states))
            costates = np.concatenate((w_prev * padded(a.traj.costates, len(a.traj))[sa]
                                       + w_next * padded(b.traj.costates, len(b.traj))[sb], costates))
            controls = np.concatenate((w_prev * padded(a.traj.controls, len(a.traj))[sa]
                                       + w_next * padded(b.traj.controls, len(b.traj))[sb], controls))
            cost = np.concatenate((w_prev * a.traj.cost[sa] + w_next * b.traj.cost[sb], cost))
            if index_lo == 0:
                break
            index_hi = index_lo - 1
            if a_is_younger:
                prev_slot = a._parents[0]
                w_prev = w_prev / 2
                w_next = 1 - w_prev
            else:
                next_slot = b._parents[1]
                w_next = w_next / 2
                w_prev = 1 - w_next
            del m
        site._traj_full = Trajectory(times, states, site.traj.coords, controls=controls, costates=costates,
                                     cost=cost, events=site.traj.events)

    def save_results(self):
        super().save_results()
        if self.solution_site is not None and self.solution_site.traj_full is not None:
            self.pb.io.save_traj(self.solution_site.traj_full, name='solution')

    def save_rff(self, nt: Optional[int] = None, fmt: Optional[str] = None):
        """Reachability-front function of the solved front in the legacy `rff` layout the reference's viewer reads
        (`docs/rff_data_h5_format.txt`, `display/display.py:823-855`): data[k] = V - ts[k] with V the minimum-time map of
        `cost_map_triangle()` (the device rasteriser, K4) shifted to absolute time, so that {data[k] <= 0} is the set
        reached by ts[k]; cells no triangle covers stay +inf.  h5 when h5py is available, else npz (same names)."""
        value = self.cost_map_triangle() + self.t_init
        ts = self.times if nt is None else np.linspace(self.times[0], self.times[-1], nt)
        nx, ny = value.shape
        grid = np.stack(np.meshgrid(np.linspace(self.pb.bl[0], self.pb.tr[0], nx),
                                    np.linspace(self.pb.bl[1], self.pb.tr[1], ny), indexing='ij'), -1)
        return self.pb.io.save_rff(value[None, :, :] - ts[:, None, None], ts, grid, self.pb.coords, fmt=fmt)

    def save_trajs_h5(self, filename: str = 'trajectories.h5'):
        """Every site trajectory in the legacy h5 layout (`docs/traj_data_h5_format.txt`); needs h5py."""
        sites = self.sites
        return self.pb.io.save_trajs_h5([s.traj for s in sites.values()], filename, names=list(sites.keys()))

    def save_front_bundle(self, name='ef_front'):
        """The whole front in ONE npz (padded trajectories + site records): the usable output at 10^5+ extremals."""
        n = self._n_sites()
        data = self._handle.trajs(0, n)
        self.pb.io.save_trajs_bundle(name, records=self._all_records(), times_grid=self.times,
                                     names=np.array([self.site_mngr.name_from_index(d)
                                                     for d in self._all_records()['dyadic']]), **data)

    # ---- device -> host --------------------------------------------------------------------------------------
    def _invalidate(self):
        self._arrivals = None
        self._n_sites_cache = None
        self._sites = None
        self._records = None
        self._slot_cache = {}
        self._bulk = None

    def _n_sites(self):
        if self._n_sites_cache is None:
            self._n_sites_cache = int(self._handle.stats()['n_sites'])
        return self._n_sites_cache

    def _all_records(self):
        if self._records is None:
            self._records = self._handle.sites(0, self._n_sites())
        return self._records

    def _site_by_slot(self, slot: int) -> Site:
        site = self._slot_cache.get(slot)
        if site is None:
            # one record for a single site of a big front, the whole table once `sites` has been asked for
            rec = self._records[slot] if self._records is not None else self._handle.sites(slot, 1)[0]
            site = Site(self, slot, rec, self.site_mngr.name_from_index(int(rec['dyadic'])))
            self._slot_cache[slot] = site
        return site

    _BULK_LIMIT = 20000

    def _fetch_traj(self, site: Site) -> Trajectory:
        n_sites = self._n_sites()
        if n_sites <= self._BULK_LIMIT:
            if self._bulk is None:
                self._bulk = self._handle.trajs(0, n_sites)
            data, row = self._bulk, site.slot
        else:
            data, row = self._handle.trajs(site.slot, 1), 0
        k0, n = site.index_t_init, site._n_samples
        sl = slice(k0, k0 + n)
        events = {name: np.array(()) for name in self._event_names}
        events['target'] = site._target_events
        if site.obstacle_name and site.t_enter_obs is not None:
            events[site.obstacle_name] = np.array((site.t_enter_obs,))
        controls = data['controls'][row, sl] if n > 1 else None
        return Trajectory.cartesian(data['times'][row, sl], data['states'][row, sl], controls=controls,
                                    costates=data['costates'][row, sl], cost=data['cost'][row, sl], events=events)


class SolverEFBisection(SolverEFResampling):
    pass


class SolverEFTrimming(SolverEFResampling):
    """Resampling solver advancing by sub-frames of `n_index_per_subframe` grid indices; after each sub-frame the
    extremals dominated by the triangle-interpolated cost map of the recent front are closed as SUBOPTIMAL
    (solver_ef.py:762-857)."""
    _VARIANT = nat.TRIMMING

    def __init__(self, *args, n_index_per_subframe: int = 10, trimming_band_width: int = 3, **kwargs):
        self.n_index_per_subframe = n_index_per_subframe
        self._trimming_band_width = trimming_band_width
        super().__init__(*args, **kwargs)
        self.i_subframe = 0
        self.depth = 1

    @property
    def n_subframes(self):
        q = self.n_index_per_subframe
        return self.n_time // q + (1 if self.n_time % q > 0 else 0)

    def setup(self):
        super().setup()
        self.i_subframe = 0

    def step(self):
        self._ensure_handle().call('dabry_step')
        self._invalidate()
        self.i_subframe += 1

    def _after_solve(self):
        self.i_subframe = self.n_subframes

    @property
    def _partial_cost_map(self):
        return self._ensure_handle().cost_map(100, 100, full=False)


def cost_map_triangle(sites, index_t_lo: int, index_t_hi: int, bl: ndarray, tr: ndarray, nx: int, ny: int):
    """Host restatement of the triangle cost map over `Site` objects (solver_ef.py:860-889); the solvers use the
    device rasteriser (csrc/site.cuh `raster_triangle`), this function serves post-processing of read-back sites."""
    res = np.full((nx, ny), np.inf)
    grid = np.stack(np.meshgrid(np.linspace(bl[0], tr[0], nx), np.linspace(bl[1], tr[1], ny), indexing='ij'), -1)
    for site in sites:
        for k in range(max(index_t_lo, site.index_t_init), min(index_t_hi + 1, site.index_t_check_next)):
            nb = site.next_nb[k]
            p1, p3 = site.state_at_index(k), site.state_at_index(k + 1)
            p2, p4 = nb.state_at_index(k), nb.state_at_index(k + 1)
            c1, c3 = site.cost_at_index(k), site.cost_at_index(k + 1)
            c2, c4 = nb.cost_at_index(k), nb.cost_at_index(k + 1)
            np.minimum(res, triangle_mask_and_cost(grid, p1, p3, p4, c1, c3, c4), out=res)
            if k > 0:
                np.minimum(res, triangle_mask_and_cost(grid, p1, p4, p2, c1, c4, c2), out=res)
    return res
