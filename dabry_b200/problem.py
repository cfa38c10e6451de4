"""`NavigationProblem` of the drop-in API (reference: dabry/problem.py).

Same constructor, attributes and catalogue (`from_name`) as the reference.  The methods on the solver's hot path
(`augsys_dyn_timeopt*`, `timeopt_control_*`, `in_obs`; problem.py:172-191, 664-670) are kept as host numpy code for
the public API and the parity tests; `SolverEF*` runs their device counterparts (csrc/model.cuh).
"""
import json
import math
import os
import warnings
from typing import List, Optional

import numpy as np
import scipy.integrate as scitg
from numpy import ndarray

from dabry_b200.feedback import Feedback, GSTargetFB, HTargetFB
from dabry_b200.flowfield import (FlowField, DiscreteFF, WrapperFF, StateLinearFF, VortexFF, ZeroFF, GyreFF,
                                  PointSymFF, UniformFF, RadialGaussFF, RankineVortexFF, LinearFFT, ChertovskihFF,
                                  TrapFF, BandFF, GyreMSEASFF)
from dabry_b200.io_manager import IOManager
from dabry_b200.misc import Utils, Coords
from dabry_b200.model import Model
from dabry_b200.obstacle import Obstacle, CircleObs, FrameObs, WrapperObs
from dabry_b200.penalty import Penalty, NullPenalty, WrapperPen
from dabry_b200.trajectory import Trajectory

# name -> (short name, time_bound or None, in the reference's CI set); the columns of dabry/problems.csv that code reads
_CATALOGUE = {
    'linear': ('linear', None, True), 'gyre': ('gyre', None, True), 'gyre_rhoads2010': ('gyre_rhoads', 1., True),
    'three_vortices': ('3vor', None, True), 'chertovskih': ('chertovskih', None, True),
    'moving_vortex': ('movor', None, True), 'montreal_reykjavik': ('montreal_reykjavik', None, False),
    'big_rankine': ('big_rankine', 1., True), 'obstacle': ('obs', None, True),
    'spherical_geodesic': ('spherical_geodesic', None, True), 'trap': ('trap', 2., True),
    'stream': ('stream', None, False), 'point_symmetric_techy2011': ('pointsym', None, True),
    'dakar_natal_constr': ('dakar_natal_constr', None, False), 'four_vortices': ('4vor', None, False),
    'linear_time_varying': ('tv_linear', None, False), 'moving_vortices': ('movors', None, False),
    'three_obstacles': ('3obs', None, False), 'gyre_li2020': ('gyre_li2020', None, False),
    'atlantic': ('atlantic', None, False), 'reykjavik_dublin': ('reykjavik_dublin', None, False),
    'double_gyre_time_dependent': ('dg_td', None, False),
}


def timeopt_control_cartesian(costate: ndarray, srf_max: float):
    """u = -srf p / |p| (problem.py:664-665)."""
    return -srf_max * costate / np.linalg.norm(costate)


def timeopt_control_gcs(state: ndarray, costate: ndarray, srf_max: float):
    """u = -srf q / |q| with q = (p_lon / cos(lat), p_lat) (problem.py:668-670)."""
    q = costate.dot(np.array(((1 / np.cos(state[..., 1]), 0.), (0., 1.))))
    return -srf_max * q / np.linalg.norm(q)


class NavigationProblem:
    ALL = {name: {'s_name': s, 'time_bound': '' if tb is None else str(tb), 'gh_wf': str(ci)}
           for name, (s, tb, ci) in _CATALOGUE.items()}

    def __init__(self, ff: FlowField, x_init: ndarray, x_target: ndarray, srf_max: float,
                 obstacles: Optional[List[Obstacle]] = None, bl: Optional[ndarray] = None,
                 tr: Optional[ndarray] = None, name: Optional[str] = None, target_radius: Optional[float] = None,
                 aero=None, penalty: Optional[Penalty] = None, autoframe=True):
        self.model = Model.zermelo(ff)
        self.x_init = x_init.copy()
        self.x_target = x_target.copy()
        self.srf_max = srf_max
        self.aero = aero
        self.bl = bl.copy() if bl is not None else np.array(())
        self.tr = tr.copy() if tr is not None else np.array(())
        self.length_reference = np.min(tr - bl) if bl is not None and tr is not None else \
            np.linalg.norm(x_target - x_init)
        self.name = 'Unnamed problem' if name is None else name
        self.obstacles = obstacles if obstacles is not None else []
        self.penalty = NullPenalty() if penalty is None else penalty
        self.io = IOManager(self.name)
        if self.bl.size == 0 or self.tr.size == 0:
            # default box: the grid extent for a bounded field, else a square 1.15 L around the end points
            # (problem.py:79-94)
            flow = self.model.ff
            if hasattr(flow, 'bounds'):
                self.bl = np.array((flow.bounds[-2, 0], flow.bounds[-1, 0]))
                self.tr = np.array((flow.bounds[-2, 1], flow.bounds[-1, 1]))
                out = [label for label, p in (('x_init', self.x_init), ('x_target', self.x_target))
                       if np.any(p < self.bl) or np.any(p > self.tr)]
                if out:
                    warnings.warn('When setting problem bounds to ff bounds, the following were out of bounds: '
                                  + ' '.join(out), category=UserWarning)
            else:
                half = 0.5 * 1.15 * self.length_reference
                mid = (self.x_init + self.x_target) / 2.
                self.bl = mid - np.array((half, half))
                self.tr = mid + np.array((half, half))
        self.target_radius = target_radius if target_radius is not None else \
            0.025 * np.linalg.norm(self.tr - self.bl)
        if autoframe:
            # the domain frame is always the LAST obstacle (problem.py:99-104)
            self.obs_frame = FrameObs(self.bl, self.tr)
            self.obstacles.append(self.obs_frame)

    def get_grid_params(self, nx: int, ny: int):
        return self.bl, (self.tr - self.bl) / (np.array((nx, ny)) - np.ones(2).astype(np.int32))

    def update_ff(self, ff: FlowField):
        self.model = Model.zermelo(ff)

    def save_ff(self):
        self.io.save_ff(self.model.ff, bl=self.bl, tr=self.tr)

    def save_info(self):
        info = {'x_init': tuple(self.x_init), 'x_target': tuple(self.x_target), 'srf_max': self.srf_max,
                'target_radius': self.target_radius, 'bl': self.bl.tolist(), 'tr': self.tr.tolist(),
                'time_orthodromic': self.time_orthodromic(), 'time_htarget': self.time_htarget()}
        self.io.setup_dir()
        with open(os.path.join(self.io.case_dir, f'{self.name}.json'), 'w') as f:
            json.dump(info, f, indent=4)

    @property
    def coords(self):
        return self.model.coords

    def __str__(self):
        return str(self.name)

    def dualize(self):
        """The problem from target to start in the reversed flow (problem.py:144-158)."""
        return NavigationProblem(self.model.ff.dualize(), self.x_target.copy(), self.x_init.copy(), self.srf_max,
                                 obstacles=[o for o in self.obstacles if o is not getattr(self, 'obs_frame', None)],
                                 bl=self.bl, tr=self.tr, target_radius=self.target_radius, aero=self.aero,
                                 penalty=self.penalty)

    def distance(self, x1, x2):
        return Utils.distance(x1, x2, self.coords)

    # ---- hot-path functions, host restatement (device: csrc/model.cuh rhs_aug) -----------------------------
    def timeopt_control_cartesian(self, costate: ndarray):
        return timeopt_control_cartesian(costate, self.srf_max)

    def timeopt_control_gcs(self, state: ndarray, costate: ndarray):
        return timeopt_control_gcs(state, costate, self.srf_max)

    def augsys_dyn_timeopt(self, t: float, state: ndarray, costate: ndarray, control: ndarray):
        """(x', p') with p' = -(df/dx)^T p - grad(penalty) (problem.py:172-175)."""
        dyn = self.model.dyn
        return np.hstack((dyn.value(t, state, control),
                          -dyn.d_value__d_state(t, state, control).transpose() @ costate
                          - self.penalty.d_value(t, state)))

    def augsys_dyn_timeopt_cartesian(self, t, state, costate):
        return self.augsys_dyn_timeopt(t, state, costate, self.timeopt_control_cartesian(costate))

    def augsys_dyn_timeopt_gcs(self, t, state, costate):
        return self.augsys_dyn_timeopt(t, state, costate, self.timeopt_control_gcs(state, costate))

    def hamiltonian(self, t, state, costate, control):
        return costate @ (control + self.model.ff.value(t, state)) + 1

    def hamiltonian_reduced(self, t, state, costate):
        return self.hamiltonian(t, state, costate, self.timeopt_control_cartesian(costate))

    def in_obs(self, state):
        return [obs for obs in self.obstacles if obs.value(state) < 0.]

    # ---- time-horizon heuristics ------------------------------------------------------------
    # The two arrival times that fix `total_duration` (and with it the time grid of a solve) come from the device:
    # dabry_time_upper_bound integrates both feedback flights with the RK45 / event machinery of the extremals.  The
    # host `apply_feedback` below only serves callers that want the whole feedback TRAJECTORY (display); no solver uses it.
    def apply_feedback(self, fb: Feedback, rel_timeout=10, n_time=1000) -> Trajectory:
        """Integrate the closed loop x' = f(t, x, fb(t, x)) until the target disc is entered (problem.py:193-214)."""
        t0 = self.model.ff.t_start
        time_scale = self.length_reference / self.srf_max
        times = np.linspace(t0, t0 + rel_timeout * time_scale, n_time)
        radius_sq = self.target_radius ** 2

        def event_target(t, x):
            return np.sum(np.square(x - self.x_target)) - radius_sq

        event_target.terminal = True
        res = scitg.solve_ivp(lambda t, x: self.model.dyn(t, x, fb(t, x)), (times[0], times[-1]), self.x_init,
                              t_eval=times, events=[event_target])
        return Trajectory(res.t, res.y.transpose(), self.model.coords, events={'target': res.t_events[0]})

    def auto_time_upper_bound(self, device=0):
        return min(self._feedback_times(device))

    def _feedback_times(self, device=0):
        """(time_orthodromic, time_htarget) from the device, once per problem."""
        cached = getattr(self, '_feedback_times_cache', None)
        if cached is None:
            from dabry_b200 import _lowering
            cached = self._feedback_times_cache = _lowering.time_upper_bound(self, device=device)
        return cached

    def orthodromic(self):
        return self.apply_feedback(GSTargetFB(self.model.ff, self.srf_max, self.x_target))

    def htarget(self):
        return self.apply_feedback(HTargetFB(self.x_target, self.model.coords))

    @staticmethod
    def _arrival_time(traj):
        hits = traj.events['target']
        return hits[0] if hits.size > 0 else np.inf

    def time_orthodromic(self):
        return self._feedback_times()[0]

    def time_htarget(self):
        return self._feedback_times()[1]

    # ---- non-dimensionalisation ------------------------------------------------------------------------------
    def rescale(self):
        """Problem in units where the vehicle speed is 1 and (planar case) the box is O(1) (problem.py:241-281).
        On the sphere lon/lat stay in radians and only time is rescaled."""
        if self.coords == Coords.CARTESIAN:
            scale_length = self.length_reference
            x_init = (self.x_init - self.bl) / scale_length
            x_target = (self.x_target - self.bl) / scale_length
            bl_new, tr_new = np.zeros(2), (self.tr - self.bl) / scale_length
            bl_wrapper = self.bl.copy()
            srf = self.srf_max
        else:
            scale_length = 1.
            x_init, x_target = self.x_init, self.x_target
            bl_new, tr_new = self.bl, self.tr
            bl_wrapper = np.zeros(2)
            srf = self.srf_max / Utils.EARTH_RADIUS  # metres per second -> radians per second
        scale_time = scale_length / srf
        t0 = self.model.ff.t_start
        flow = WrapperFF(self.model.ff, scale_length, bl_wrapper, scale_time, t0)
        user_obstacles = [obs for obs in self.obstacles if obs is not self.obs_frame]
        obstacles = [WrapperObs(obs, scale_length, bl_wrapper) for obs in user_obstacles]
        penalty = WrapperPen(self.penalty, scale_length, bl_wrapper, scale_time, t0)
        return NavigationProblem(flow, x_init, x_target, 1., bl=bl_new, tr=tr_new, obstacles=obstacles,
                                 penalty=penalty, name=self.name + ' (scaled)')

    # ---- catalogue ---------------------------------------------------------------------------------------------
    @classmethod
    def base_name(cls, name: str):
        for b_name, (s_name, _, _) in _CATALOGUE.items():
            if name in (b_name, s_name):
                return b_name
        raise ValueError('Cannot find problem "%s"' % name)

    @classmethod
    def time_bound(cls, name: str):
        """`time_bound` column of problems.csv, passed as total_duration by the reference tests."""
        return _CATALOGUE[cls.base_name(name)][1]

    @classmethod
    def from_database(cls, *args, **kwargs):
        raise NotImplementedError('ERA5 download (CDS + pygrib) is out of scope: build a DiscreteFF from arrays '
                                  'or DiscreteFF.from_npz and pass it to NavigationProblem')

    @classmethod
    def from_name(cls, name: str):
        """The reference's problem catalogue (problem.py:324-661), same parameters."""
        b_name = cls.base_name(name)
        arr = np.array
        unit_start, unit_target = arr([0., 0.]), arr([1., 0.])
        if b_name == 'linear':
            ff = StateLinearFF(arr([[0., 1.], [0., 0.]]), arr([0., 0.]), arr([0., 0.]))
            return cls(ff, unit_start, unit_target, 1., bl=arr([-0.2, -1.]), tr=arr([1.2, 1.]), name=b_name)
        if b_name == 'three_vortices':
            centers = [(0.5, 0.5), (0.5, 0.), (0.5, -0.5)]
            ff = sum([VortexFF(arr(c), g) for c, g in zip(centers, (1., -1, 1))], ZeroFF())
            obs = [CircleObs(arr(c), 0.1) for c in centers]
            return cls(ff, unit_start, unit_target, 1., bl=arr([-0.1, -1]), tr=arr([1.1, 1]), obstacles=obs,
                       name=b_name)
        if b_name == 'gyre':
            return cls(GyreFF(0.5, 0.5, 2., 2., 1), arr((0.6, 0.6)), arr((2.4, 2.4)), 1, name=b_name)
        if b_name == 'gyre_li2020':
            return cls(GyreFF(0., 0., 500., 500., 1.), arr((125., 125.)), arr((375., 375.)), 0.6,
                       bl=arr((-10, -10)), tr=arr((510, 510)), name=b_name)
        if b_name == 'point_symmetric_techy2011':
            return cls(PointSymFF(arr((0.5, 0.3)), 1., 0.), arr((0., 0.)), arr((1., 0.)), 1.,
                       bl=arr((-0.1, -1.)), tr=arr((1.1, 1.)), name=b_name)
        if b_name == 'three_obstacles':
            blowers = [RadialGaussFF(arr(c), 0.1, 0.5 * 0.2, 5.) for c in ((0.5, 0.), (0.5, 0.1), (0.5, -0.1))]
            ff = blowers[0] + blowers[1] + blowers[2] + UniformFF(arr([.1, .1]))
            return cls(ff, arr((0., 0.)), arr((1., 0.)), 1., bl=arr((-0.15, -0.65)), tr=arr((1.15, 0.65)),
                       name=b_name)
        if b_name == 'big_rankine':
            return cls(RankineVortexFF(arr((0.5, 0.5)), -1., 1.), arr([0.2, 0.5]), arr([1.3, 0.5]), 0.01,
                       bl=arr((-0.3, -0.3)), tr=arr((1.6, 1.3)), name=b_name)
        if b_name == 'four_vortices':
            centers = arr(((0.5, 0.5), (0.5, 0.2), (0.5, -0.2), (0.5, -0.5)))
            strength = arr([1., -1., 1.5, -1.5])
            vortices = [RankineVortexFF(centers[i], strength[i], 1e-1) for i in range(4)]
            return cls(sum(vortices, UniformFF(arr([0., 0.]))), unit_start, unit_target, 1.,
                       bl=arr([-0.2, -1.]), tr=arr([1.2, 1.]), name=b_name)
        if b_name == 'moving_vortex':
            nt = 10
            track = np.stack((np.linspace(0.5, 0.5, nt), np.linspace(-0.4, 0.4, nt)), axis=1)
            ff = RankineVortexFF(track, -1. * np.ones(nt), 1e-1 * np.ones(nt), t_end=1.)
            return cls(ff, unit_start, unit_target, 1., name=b_name)
        if b_name == 'linear_time_varying':
            ff = LinearFFT(arr(((0., -3.), (0., 0.))), arr(((0., 3.), (0., 0.))), arr([0., 0.]), arr([0., 0.]))
            return cls(ff, unit_start, unit_target, 1., bl=arr([-0.2, -1.]), tr=arr([1.2, 1.]), name=b_name)
        if b_name == 'moving_vortices':
            nt = 10
            y_tracks = (np.linspace(-0.3, 0.1, nt), np.linspace(-0.1, 0.5, nt), np.linspace(-0.2, 0.4, nt)[::-1],
                        np.linspace(-0.4, 0.2, nt)[::-1])
            fields: list = [UniformFF(arr((0., 0.)))]
            for x_vortex, ys in zip((0.2, 0.4, 0.6, 0.8), y_tracks):
                track = np.stack((np.linspace(x_vortex, x_vortex, nt), ys), axis=1)
                fields.append(RankineVortexFF(track, -1. * np.ones(nt), 1e-3 * np.ones(nt), t_end=1.))
            weight = 1 / len(fields)
            return cls(sum([weight * f for f in fields], ZeroFF()), unit_start, unit_target, 1., name=b_name)
        if b_name == 'gyre_rhoads2010':
            return cls(GyreFF(0, -0.5, 2, 2, 2), arr((0, 0.25)), arr((0.03, -0.25)), 1.,
                       bl=arr((-1, -0.5)), tr=arr((1., 0.5)))
        if b_name == 'spherical_geodesic':
            x_init, x_target = arr((0, np.pi / 6)), arr((np.pi / 4, np.pi / 6))
            margin = arr((0.15, 0.15))
            # only a DiscreteFF can carry GCS coordinates: sample the null field (problem.py:577-584)
            ff = DiscreteFF.from_ff(ZeroFF(), (x_init - margin, x_target + margin), coords=Coords.GCS)
            return cls(ff, x_init, x_target, 1, name=b_name)
        if b_name == 'chertovskih':
            return cls(ChertovskihFF(), arr((0.5, 0.)), arr((-0.7, -6)), 1., bl=arr((-1.5, -6.2)),
                       tr=arr((1.5, 0.2)), name=b_name)
        if b_name == 'obstacle':
            return cls(ZeroFF(), arr((0., 0.)), arr((1., 0.)), 1., obstacles=[CircleObs((0.5, 0.1), 0.2)],
                       name=b_name)
        if b_name == 'trap':
            nt = 40
            center = np.zeros((nt, 2))
            center[5:35, 0] = 0.05 * np.arange(30)
            ff = TrapFF(2 * np.ones(nt), center, 0.2 * np.ones(nt), t_end=4)
            return cls(ff, arr((0., 0.)), arr((1., 0.)), 1., bl=arr((-0.2, -0.6)), tr=arr((1.2, 0.6)), name=b_name)
        if b_name == 'stream':
            ff = BandFF(arr((0., 2.5)), arr((1., 0.)), arr((-1., 0.)), 1)
            return cls(ff, arr((1, 1)), arr((4, 4)), 1, bl=arr((0.75, 0.75)), tr=arr((4.25, 4.25)), name=b_name)
        if b_name == 'double_gyre_time_dependent':
            return cls(GyreMSEASFF(), arr((0.25, 0.25)), arr((1.75, 0.75)), 0.5, name=b_name, bl=arr((0, 0)),
                       tr=arr((2, 1)))
        if b_name in ('atlantic', 'montreal_reykjavik', 'reykjavik_dublin', 'dakar_natal_constr'):
            return cls._from_data_file(b_name)
        raise ValueError('No corresponding problem for name "%s"' % b_name)

    @classmethod
    def _from_data_file(cls, b_name):
        """Catalogue entries backed by data files (problem.py:568-575, 595-601, 638-650).  The files are not part of
        this package: point DABRY_DATA_DIR at a checkout of the reference's data/ directory."""
        root = os.environ.get('DABRY_DATA_DIR', os.path.join(os.getcwd(), 'data'))
        if b_name == 'atlantic':
            ff = DiscreteFF.from_npz(os.path.join(root, 'demo', 'atlantic', 'flow.npz'))
            return cls(ff, np.array((-0.7, 0.6)), np.array((-0.15, 1.1)), 15., name=b_name)
        if b_name == 'dakar_natal_constr':
            ff = DiscreteFF.from_h5(os.path.join(root, 'demo', 'dakar_natal_constr', 'wind.h5'))
            return cls(ff, Utils.DEG_TO_RAD * np.array([-17.447938, 14.693425]),
                       Utils.DEG_TO_RAD * np.array([-35.2080905, -5.805398]), 23., name=b_name)
        sub = ('demo', 'montreal_reykjavik') if b_name == 'montreal_reykjavik' else ('cds_omerc', 'reykjavik_dublin')
        ff = DiscreteFF.from_npz(os.path.join(root, *sub, 'ff.npz'))
        extent = ff.bounds[1:, 1] - ff.bounds[1:, 0]
        return cls(ff, np.diag((5 / 6, 1 / 2)) @ extent, np.diag((1 / 6, 1 / 2)) @ extent, 10, name=b_name)
