"""TEST INFRASTRUCTURE - CPU oracle of the extremal-field hot path.  Not imported by the dabry_b200 package; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may use it.

A numpy/scipy restatement of the reference's solver loops (dabry v1.1, dabry/solver_ef.py), written against the
*public* problem interface (`pb.model.dyn`, `pb.model.ff`, `pb.obstacles`, `pb.penalty`, `pb.in_obs`, ...) so that it
runs both on the reference's own `NavigationProblem` (in the build container: validates the restatement against the
unmodified reference, tests/test_oracle_port.py) and on dabry_b200's drop-in classes (on the GPU box, where
/root/reference does not exist).  Like the reference it integrates every extremal with
`scipy.integrate.solve_ivp` (RK45, rtol 1e-3, atol 1e-6, max_step) - scipy is a third-party dependency of the
reference, unpinned in its requirements.txt; the goldens were produced with scipy 1.18.1.

Parity pin: tests/test_oracle_port.py checks this port against every golden vector in tests/golden (written by
oracle/gen_golden.py from the unmodified reference) - bit-for-bit on counts, names and flags, 1e-12 on floats.

Reference map
  Port.__init__            SolverEF.__init__ (solver_ef.py:284-339) + SolverEFResampling.__init__ (449-460)
  Port.rhs / rhs_constr    dyn_augsys_* (355-359), dyn_constr (361-365), problem.py:172-181, 664-670
  Port.shoot               integrate_site_to_target_index (479-580)
  Port.link                connect_to_parents (690-699)
  Port.refine              compute_new_sites (630-665) + SiteManager.site_from_parents (261-278)
  Port.prune_far           trim_distance (594-600)
  Port.arrivals            check_solutions (667-685)
  Port.run_resampling      SolverEFResampling.solve (602-608)
  Port.run_trimming        SolverEFTrimming.step / trim / solve (802-857), cost_map_triangle (860-889)
"""
import math
from types import SimpleNamespace

import numpy as np
import scipy.integrate as scitg

WITHIN_OBS, SUBOPTIMAL, IMPOSSIBLE_OBS_TRACKING = 1, 2, 3
SELF_AND_NB_IN_OBS = 0


class _HermiteDense(scitg.DenseOutput):
    """Cubic Hermite interpolant of one RK4 step, evaluated in the same operation order as the device
    (csrc/rk45.cuh rk4_dense + dense_eval)."""

    def __init__(self, t_old, t, y_old, y, f_old, f):
        super().__init__(t_old, t)
        self.h = t - t_old
        self.y_old = y_old
        slope = (y - y_old) / self.h
        self.q0 = f_old
        self.q1 = (3. * slope - 2. * f_old) - f
        self.q2 = (f_old + f) - 2. * slope

    def _call_impl(self, t):
        x = (np.asarray(t) - self.t_old) / self.h
        x2 = x * x
        x3 = x2 * x
        if x.ndim == 0:
            return self.h * ((self.q0 * x + self.q1 * x2) + self.q2 * x3) + self.y_old
        q0, q1, q2, y0 = (a[:, None] for a in (self.q0, self.q1, self.q2, self.y_old))
        return self.h * ((q0 * x + q1 * x2) + q2 * x3) + y0


class RK4Fixed(scitg.OdeSolver):
    """Integrator RK4_FIXED as a scipy OdeSolver, so that `solve_ivp(method=RK4Fixed)` applies scipy's own driver
    (events, brentq roots, t_eval sampling) around it: classic RK4 with steps of exactly max_step (the last clipped
    to t_bound), 4 right-hand sides per step, Hermite dense output.  The reference has no counterpart (it always
    runs RK45); this defines the semantics of the device's fixed-step variant (csrc/rk45.cuh rk4_step)."""

    def __init__(self, fun, t0, y0, t_bound, max_step=np.inf, vectorized=False, **extraneous):
        super().__init__(fun, t0, y0, t_bound, vectorized, support_complex=True)
        self.max_step = max_step
        self.f = self.fun(self.t, self.y)
        self.y_old = None
        self.f_old = None

    def _step_impl(self):
        t, y = self.t, self.y
        t_new = t + self.max_step
        if t_new - self.t_bound > 0:
            t_new = self.t_bound
        h = t_new - t
        k1 = self.f
        k2 = self.fun(t + 0.5 * h, y + (0.5 * h) * k1)
        k3 = self.fun(t + 0.5 * h, y + (0.5 * h) * k2)
        k4 = self.fun(t_new, y + h * k3)
        y_new = y + (h / 6.) * (((k1 + 2. * k2) + 2. * k3) + k4)
        self.y_old, self.f_old = y, k1
        self.t, self.y = t_new, y_new
        self.f = self.fun(t_new, y_new)
        return True, None

    def _dense_output_impl(self):
        return _HermiteDense(self.t_old, self.t, self.y_old, self.y, self.f_old, self.f)


def _is_gcs(pb):
    return getattr(pb.coords, 'value', pb.coords) == 'gcs'


def _alpha(i):
    if i == 0:
        return 'A'
    out = ''
    while i > 0:
        out = chr(65 + i % 26) + out
        i //= 26
    return out


class PortSite:
    """Plain record of one extremal: arrays grow as integration proceeds."""

    def __init__(self, index, name, depth, k0, t0, x0, p0, c0, n_time):
        self.index, self.name, self.depth = index, name, depth
        self.index_t_init = k0
        self.t = [t0]
        self.x = [np.array(x0, dtype=float)]
        self.p = [np.array(p0, dtype=float)]
        self.u = [np.full(2, np.nan)]
        self.cost = [c0]
        self.check_next = k0
        self.obstacle_name = ''
        self.index_t_obs = 0
        self.t_enter_obs = None
        self.obs_trigo = True
        self.closure = None
        self.neuter = None
        self.next_nb = [None] * n_time
        self.parents = (None, None)
        self.linked = False
        self.has_controls = False
        self.target_events = []

    @property
    def index_t(self):
        return self.index_t_init + len(self.t) - 1

    def in_obs_at(self, k):
        return bool(self.obstacle_name) and k >= self.index_t_obs

    def at(self, k):
        return k - self.index_t_init

    def traj_view(self):
        n = len(self.t)
        return SimpleNamespace(times=np.array(self.t), states=np.array(self.x).reshape(n, 2),
                               costates=np.array(self.p).reshape(n, 2),
                               controls=np.array(self.u).reshape(n, 2) if n > 1 else None,
                               cost=np.array(self.cost))


class Port:
    def __init__(self, pb, total_duration=None, max_depth=20, n_time=100, n_costate_sectors=30, t_init=None,
                 target_radius=None, abs_max_step=None, rel_max_step=0.01, max_dist=None, mode_origin=True,
                 n_index_per_subframe=10, trimming_band_width=3, root_range=None, integrator='dopri5',
                 open_window=False, external_cost_maps=None):
        self.pb = pb
        if total_duration is None:
            # problem.py:216-239 through the problem's own host feedback flights (scipy), never through the product's
            # device routine
            def arrival(traj):
                hits = traj.events['target']
                return hits[0] if hits.size > 0 else np.inf
            total_duration = 1.1 * min(arrival(pb.orthodromic()), arrival(pb.htarget()))
            if total_duration == np.inf:
                total_duration = 2 * pb.length_reference / pb.srf_max
        self.total_duration = total_duration
        self.t_init = pb.model.ff.t_start if t_init is None else t_init
        self.target_radius = pb.target_radius if target_radius is None else target_radius
        self.r2 = self.target_radius ** 2
        self.times = np.linspace(self.t_init, self.t_init + total_duration, n_time)
        self.n_time = n_time
        self.n_roots = n_costate_sectors
        self.max_depth = max_depth
        cap = total_duration / 2 if abs_max_step is None else abs_max_step
        self.max_int_step = cap if rel_max_step is None else min(cap, rel_max_step * total_duration)
        self.max_dist = self.target_radius if max_dist is None else max_dist
        self.d2 = self.max_dist ** 2
        self.mode_origin = mode_origin
        self.q = n_index_per_subframe
        self.band = trimming_band_width
        self.root_range = root_range
        # open_window: `root_range` is a WINDOW of a larger fan that is solved elsewhere (the full-size GPU run): the last
        # root of the window has no neighbour instead of wrapping to the first, so every site of the window is what the
        # full fan computes for it (a site only ever looks at its next neighbour).  external_cost_maps: the trimming
        # cost map of the WHOLE front per sub-frame, from that other run (the map is a global min over all triangles; the
        # window's own triangles are checked against it and the dominated-extremal test uses it).
        self.open_window = open_window
        self.external_cost_maps = external_cost_maps
        self.integrator = integrator
        self.gcs = _is_gcs(pb)
        # event table (solver_ef.py:322-334)
        self.event_names = ['target']
        self.obstacle_of = {}
        counts = {}
        for obs in pb.obstacles:
            cls = type(obs).__name__
            label = 'obs_%s_%d' % (cls, counts.get(cls, 0))
            counts[cls] = counts.get(cls, 0) + 1
            self.event_names.append(label)
            self.obstacle_of[label] = obs
        ff = pb.model.ff
        self.ff_max_norm = np.max(np.linalg.norm(ff.values, axis=-1)) if hasattr(ff, 'values') else None
        self.span = 1 << max_depth
        self.prefix_chars = int(np.ceil(math.log(self.n_roots) / math.log(26))) if self.n_roots > 1 else 0
        self.sites = {}
        self.success = False
        self.solution_site = None
        self.solution_sites = []
        self.solution_cost = float('nan')
        self.n_ivp = 0
        self.n_rk = 0
        self.cost_maps = []

    # ---- naming -------------------------------------------------------------------------------------------
    def name_of(self, index):
        prefix, u = divmod(index, self.span)
        head = _alpha(prefix).rjust(self.prefix_chars, 'A')
        if u == 0:
            return head, 0
        val = (u & -u).bit_length() - 1
        depth = self.max_depth - val
        return '%s-%d-%d' % (head, depth, ((u >> val) - 1) // 2), depth

    # ---- dynamics -----------------------------------------------------------------------------------------
    def rhs(self, t, y):
        pb = self.pb
        x, p = y[:2], y[2:]
        if self.gcs:
            q = p.dot(np.array(((1 / np.cos(x[1]), 0.), (0., 1.))))
            u = -pb.srf_max * q / np.linalg.norm(q)
        else:
            u = -pb.srf_max * p / np.linalg.norm(p)
        dyn = pb.model.dyn
        return np.hstack((dyn.value(t, x, u), -dyn.d_value__d_state(t, x, u).transpose() @ p
                          - pb.penalty.d_value(t, x)))

    def _slide_control(self, w, d):
        e = d / np.linalg.norm(d)
        n = np.array(((0., -1.), (1., 0.))) @ e
        c = w @ n
        ang = np.arctan2(np.sqrt(np.max((self.pb.srf_max ** 2 - c ** 2, 0.))), c)
        return self.pb.srf_max * np.array(((np.cos(ang), -np.sin(ang)), (np.sin(ang), np.cos(ang)))) @ (-n)

    def rhs_constr(self, t, x, label, trigo):
        sign = 2. * trigo - 1.
        d = np.array(((0., -sign), (sign, 0.))) @ self.obstacle_of[label].d_value(x)
        w = self.pb.model.ff.value(t, x)
        return w + self._slide_control(w, d)

    def _events(self):
        def target(_, y):
            return np.sum(np.square(y[:2] - self.pb.x_target)) - self.r2
        target.terminal, target.direction = False, -1.
        evs = [target]
        for label in self.event_names[1:]:
            obs = self.obstacle_of[label]

            def ev(_, y, _obs=obs):
                return _obs.value(y[:2])
            ev.terminal, ev.direction = True, -1.
            evs.append(ev)
        return evs

    def _ivp(self, *a, **k):
        if self.integrator == 'rk4':
            res = scitg.solve_ivp(*a, method=RK4Fixed, **k)
            self.n_rk += (res.nfev - 1) // 4
        else:
            res = scitg.solve_ivp(*a, **k)
            self.n_rk += (res.nfev - 2) // 6
        self.n_ivp += 1
        return res

    # ---- one extremal ---------------------------------------------------------------------------------------
    def shoot(self, s, index_t):
        if s.index_t >= index_t:
            return
        k_from = s.index_t
        t_eval = np.linspace(self.times[k_from], self.times[index_t], index_t - k_from + 1)
        if t_eval.shape[0] <= 1:
            return
        if not s.in_obs_at(k_from):
            y0 = np.hstack((s.x[-1], s.p[-1]))
            res = self._ivp(self.rhs, (t_eval[0], t_eval[-1]), y0, t_eval=t_eval[1:], events=self._events(),
                            dense_output=True, max_step=self.max_int_step)
            n_free = len(res.t)
            for j in range(n_free):
                self._append(s, res.t[j], res.y[:2, j], res.y[2:, j], None)
            s.target_events = list(res.t_events[0])
            hit = [lab for lab, te in zip(self.event_names[1:], res.t_events[1:]) if te.shape[0] > 0]
            if hit:
                label = hit[0]
                t_in = res.t_events[1 + self.event_names[1:].index(label)][0]
                s.t_enter_obs, s.obstacle_name = t_in, label
                s.index_t_obs = k_from + n_free + 1
                y_in = res.sol(t_in)
                obs = self.obstacle_of[label]
                heading = -y_in[2:] / np.linalg.norm(y_in[2:])
                g = obs.d_value(y_in[:2])
                v = self.pb.model.dyn.value(t_in, y_in[:2], heading)
                s.obs_trigo = bool(g[0] * v[1] - g[1] * v[0] >= 0.)
                sign = 2 * s.obs_trigo - 1
                direction = np.array(((0, -sign), (sign, 0))) @ g
                w = self.pb.model.ff.value(t_in, y_in[:2])
                e = direction / np.linalg.norm(direction)
                feasible = np.abs(w @ (np.array(((0., -1.), (1., 0.))) @ e)) < self.pb.srf_max
                if feasible:
                    t_next = t_eval[t_eval > t_in][0]
                    r2 = self._ivp(self.rhs_constr, (t_in, t_next), y_in[:2], t_eval=np.array((t_next,)),
                                   max_step=self.max_int_step, args=[label, s.obs_trigo])
                    if len(r2.t) == 0:
                        s.closure = IMPOSSIBLE_OBS_TRACKING
                    for j in range(len(r2.t)):
                        x = r2.y[:, j]
                        self._append(s, r2.t[j], x, None,
                                     self.rhs_constr(r2.t[j], x, label, s.obs_trigo) - self.pb.model.ff.value(r2.t[j], x))
                else:
                    s.closure = IMPOSSIBLE_OBS_TRACKING
        if s.index_t < index_t and s.obstacle_name:
            j0 = s.index_t - k_from
            r3 = self._ivp(self.rhs_constr, (t_eval[j0], t_eval[-1]), s.x[-1], t_eval=t_eval[j0 + 1:],
                           events=[self._quit_event()], max_step=self.max_int_step,
                           args=[s.obstacle_name, s.obs_trigo])
            for j in range(len(r3.t)):
                x = r3.y[:, j]
                self._append(s, r3.t[j], x, None, self.rhs_constr(r3.t[j], x, s.obstacle_name, s.obs_trigo)
                             - self.pb.model.ff.value(r3.t[j], x))
            if r3.t_events[0].size > 0:
                s.closure = IMPOSSIBLE_OBS_TRACKING

    def _quit_event(self):
        def ev(t, x, label, trigo):
            g = self.obstacle_of[label].d_value(x)
            return self.pb.srf_max - np.abs(self.pb.model.ff.value(t, x) @ (g / np.linalg.norm(g)))
        ev.terminal, ev.direction = True, -1.
        return ev

    def _append(self, s, t, x, p, u):
        s.t.append(float(t))
        s.x.append(np.array(x, dtype=float))
        s.p.append(np.full(2, np.nan) if p is None else np.array(p, dtype=float))
        s.u.append(np.full(2, np.nan) if u is None else np.array(u, dtype=float))
        s.cost.append(float(t) - self.t_init)

    def link(self, s):
        a, b = s.parents
        if a is None or s.linked:
            return
        s.linked = True
        k = s.index_t_init
        a.next_nb[k] = s
        s.next_nb[k] = b
        a.check_next = k

    # ---- front ------------------------------------------------------------------------------------------------
    def make_roots(self):
        n = self.n_roots
        lo, cnt = (0, n) if self.root_range is None else self.root_range
        thetas = np.linspace(0, 2 * np.pi, n, endpoint=False)
        roots = []
        for i in range(lo, lo + cnt):
            name, _ = self.name_of(i * self.span)
            roots.append(PortSite(i * self.span, name, 0, 0, self.t_init, self.pb.x_init,
                                  1000 * np.array((np.cos(thetas[i]), np.sin(thetas[i]))), 0., self.n_time))
        for i, s in enumerate(roots):
            s.next_nb[0] = roots[(i + 1) % len(roots)]
        if self.open_window:
            roots[-1].next_nb[0] = None
        self.sites = {s.name: s for s in roots}
        return roots

    def refine(self, index_t_hi=None):
        hi = self.n_time - 1 if index_t_hi is None else index_t_hi
        born = []
        for s in list(self.sites.values()):
            nb = s.next_nb[s.check_next]
            if nb is None:  # last root of an open window
                continue
            reached = s.check_next
            child = None
            for k in range(s.check_next, min(s.index_t + 1, nb.index_t + 1, hi)):
                if s.in_obs_at(k) and nb.in_obs_at(k):
                    s.neuter = SELF_AND_NB_IN_OBS
                    break
                if np.sum(np.square(s.x[s.at(k)] - nb.x[nb.at(k)])) > self.d2:
                    from_origin = self.mode_origin and not s.in_obs_at(k) and not nb.in_obs_at(k) \
                        and s.index_t_init == 0 and nb.index_t_init == 0
                    kc = 0 if from_origin else k
                    if s.depth < self.max_depth - 1 and nb.depth < self.max_depth - 1:
                        child = self._midpoint(s, nb, kc)
                        if len(self.pb.in_obs(child.x[0])) > 0:
                            s.closure = WITHIN_OBS
                            child = None
                    break
                reached = k
            for k in range(s.check_next, reached + 1):
                s.next_nb[k] = nb
            s.check_next = reached
            if child is not None:
                self.sites[child.name] = child
                born.append(child)
        return born

    def _midpoint(self, a, b, k):
        pa, pb_ = a.p[a.at(k)], b.p[b.at(k)]
        if a.in_obs_at(k):
            pa = -a.u[a.at(k)] * np.linalg.norm(pb_)
        if b.in_obs_at(k):
            pb_ = -b.u[b.at(k)] * np.linalg.norm(pa)
        ia, ib = a.index, b.index
        if ib < ia:
            ib += self.n_roots * self.span
        index = (ia + ib) // 2
        name, depth = self.name_of(index)
        child = PortSite(index, name, depth, k, a.t[a.at(k)], 0.5 * (a.x[a.at(k)] + b.x[b.at(k)]),
                         0.5 * (pa + pb_), 0.5 * (a.cost[a.at(k)] + b.cost[b.at(k)]), self.n_time)
        child.parents = (a, b)
        return child

    def prune_far(self):
        if self.ff_max_norm is None or np.isclose(self.ff_max_norm, 0):
            return
        for s in self.sites.values():
            if self.times[-1] - self.times[s.index_t] > np.linalg.norm(s.x[-1] - self.pb.x_target) / self.ff_max_norm:
                s.closure = SUBOPTIMAL

    def arrivals(self):
        best = None
        costs = {}
        for s in self.sites.values():
            inside = np.sum(np.square(np.array(s.x) - self.pb.x_target), axis=-1) < self.r2
            if np.any(inside):
                self.success = True
                costs[s.name] = np.min(np.array(s.cost)[inside])
                if best is None or costs[s.name] < costs[best]:
                    best = s.name
        if best is not None:
            self.solution_site = self.sites[best]
            self.solution_cost = float(costs[best])
            self.solution_sites = [self.sites[n] for n, c in costs.items() if not c > 1.15 * costs[best]]
        self.arrival_costs = costs

    # ---- solvers ------------------------------------------------------------------------------------------------
    def run_resampling(self):
        to_shoot = self.make_roots()
        for _ in range(self.max_depth):
            for s in to_shoot:
                self.shoot(s, self.n_time - 1)
                self.link(s)
            to_shoot = self.refine()
            self.prune_far()
        self.arrivals()
        return self

    def run_simple(self, depth=0):
        """One fan of SolverEFSimple.step (solver_ef.py:419-437); returns the list of solve_ivp results."""
        angles = np.linspace(0., 2 * np.pi, 2 ** depth * self.n_roots + 1)
        if depth != 0:
            angles = angles[1::2]
        out = []
        for th in angles:
            y0 = np.array(tuple(self.pb.x_init) + (np.cos(th), np.sin(th)))
            out.append(self._ivp(self.rhs, (self.t_init, self.t_init + self.total_duration), y0, t_eval=self.times,
                                 events=self._events(), max_step=self.max_int_step))
        return out

    def run_trimming(self):
        roots = self.make_roots()
        n_sub = self.n_time // self.q + (1 if self.n_time % self.q > 0 else 0)
        valid = [[] for _ in range(n_sub + 1)]
        created = [[] for _ in range(n_sub)]
        valid[0] = list(roots)
        created[0] = list(roots)
        for i in range(n_sub):
            target = min((i + 1) * self.q, self.n_time) - 1
            pending = list(valid[i])
            while pending:
                for s in pending:
                    self.shoot(s, target)
                    self.link(s)
                pending = self.refine(index_t_hi=target)
                created[i].extend(pending)
            # trim (solver_ef.py:819-841); the valid-set slice keeps Python's negative-start semantics
            start = i + 1 - self.band
            pool = {id(s): s for s in created[i]}
            for group in valid[start:i + 1]:
                pool.update({id(s): s for s in group})
            cmap = self.cost_map(pool.values(), start * self.q, target)
            if self.external_cost_maps is not None:
                ext = np.asarray(self.external_cost_maps[i])
                # the global map is a min over a superset of this window's triangles
                own = np.isfinite(cmap)
                assert np.all(ext[own] <= cmap[own] * (1 + 1e-9) + 1e-12), 'window triangles below the global cost map'
                self.window_cells_equal = getattr(self, 'window_cells_equal', 0) + int(np.sum(np.isclose(ext[own], cmap[own], rtol=1e-9, atol=1e-12)))
                self.window_cells = getattr(self, 'window_cells', 0) + int(own.sum())
                cmap = ext
            self.last_cost_map = cmap
            self.cost_maps.append(cmap)
            considered = {id(s): s for s in valid[i]}
            considered.update({id(s): s for s in created[i]})
            bl = np.asarray(self.pb.bl, dtype=float)
            spacing = (np.asarray(self.pb.tr) - bl) / 99.
            for s in considered.values():
                if s.closure is not None:
                    continue
                v = self._bilinear(cmap, bl, spacing, s.x[s.at(target)])
                if not np.isnan(v) and s.cost[s.at(target)] > 1.02 * v:
                    s.closure = SUBOPTIMAL
                else:
                    valid[i + 1].append(s)
        self.arrivals()
        return self

    @staticmethod
    def _bilinear(grid, bl, spacing, x):
        pos = (x - bl) / spacing
        lo = np.floor(pos).astype(np.int32)
        wh = pos - lo
        wl = 1 - wh
        top = grid.shape[0] - 1
        i0, i1 = np.clip(lo, 0, top), np.clip(lo + 1, 0, top)
        return ((wl[0] * wl[1]) * grid[i0[0], i0[1]] + (wl[0] * wh[1]) * grid[i0[0], i1[1]]
                + (wh[0] * wl[1]) * grid[i1[0], i0[1]]) + (wh[0] * wh[1]) * grid[i1[0], i1[1]]

    def cost_map(self, sites, k_lo, k_hi, nx=100, ny=100):
        out = np.full((nx, ny), np.inf)
        bl, tr = self.pb.bl, self.pb.tr
        gx = np.linspace(bl[0], tr[0], nx)[:, None]
        gy = np.linspace(bl[1], tr[1], ny)[None, :]
        for s in sites:
            for k in range(max(k_lo, s.index_t_init), min(k_hi + 1, s.check_next)):
                nb = s.next_nb[k]
                p1, p3 = s.x[s.at(k)], s.x[s.at(k + 1)]
                p2, p4 = nb.x[nb.at(k)], nb.x[nb.at(k + 1)]
                c1, c3 = s.cost[s.at(k)], s.cost[s.at(k + 1)]
                c2, c4 = nb.cost[nb.at(k)], nb.cost[nb.at(k + 1)]
                self._raster(out, gx, gy, p1, p3, p4, c1, c3, c4)
                if k > 0:
                    self._raster(out, gx, gy, p1, p4, p2, c1, c4, c2)
        return out

    @staticmethod
    def _raster(out, gx, gy, p1, p2, p3, c1, c2, c3):
        v1, v2 = p2 - p1, p3 - p1
        den = v1[0] * v2[1] - v1[1] * v2[0]
        if np.isclose(den, 0.):
            return
        a = ((gx * v2[1] - gy * v2[0]) - (p1[0] * v2[1] - p1[1] * v2[0])) / den
        b = -((gx * v1[1] - gy * v1[0]) - (p1[0] * v1[1] - p1[1] * v1[0])) / den
        inside = (a > 0) & (b > 0) & (a + b < 1)
        if np.any(inside):
            val = (1 - (a + b)) * c1 + a * c2 + b * c3
            out[inside] = np.minimum(out[inside], val[inside])


def solve_port(pb, variant='R', **kwargs):
    """Run the oracle on a problem: variant 'R' (resampling) or 'T' (trimming).  Returns the Port; its `sites`
    dict maps names to PortSite (use `.traj_view()` for arrays; `.traj` is set to the same view)."""
    port = Port(pb, **kwargs)
    port.run_trimming() if variant == 'T' else port.run_resampling()
    for s in port.sites.values():
        s.traj = s.traj_view()
    return port
