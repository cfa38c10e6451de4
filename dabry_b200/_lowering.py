"""Lowering pass: a `NavigationProblem` (flow-field expression tree, obstacles, penalty, wrappers) -> the flat
structures of include/dabry_b200.h.  Arbitrary Python subclasses cannot run on the device: anything outside the
supported set raises `LoweringError` (no silent CPU evaluation)."""
import ctypes as C

import numpy as np

from dabry_b200 import _native as nat
from dabry_b200.flowfield import DiscreteFF, WrapperFF, LoweringError, FlowField
from dabry_b200.misc import Coords


def unwrap_flow(ff):
    """(wrapper struct, inner field) of `ff`; identity wrapper when `ff` is not a WrapperFF."""
    w = nat.Wrapper(0., 1., 1., (0., 0.), 1., 1.)
    if isinstance(ff, WrapperFF):
        w = nat.Wrapper(float(ff.time_origin), float(ff.scale_time), float(ff.scale_length),
                        (float(ff.bl[0]), float(ff.bl[1])), float(ff._scaler_speed), float(ff._scaler_dspeed))
        ff = ff.ff
        if isinstance(ff, WrapperFF):
            raise LoweringError('nested WrapperFF is not supported on the device')
    return w, ff


def apply_flow(handle, ff, shared=False):
    """Upload the flow field of a problem into a native handle.  shared: every rank of the handle's communicator holds
    the same grid - each uploads its share and the peers exchange the rest (dabry_set_flow_grid_shared)."""
    wrapper, inner = unwrap_flow(ff)
    if isinstance(inner, DiscreteFF):
        values = nat.f64(inner.values)
        unsteady = values.ndim == 4
        nt = values.shape[0] if unsteady else 1
        nx, ny = values.shape[-3], values.shape[-2]
        # host-supplied Jacobian samples (DiscreteFF.from_ff) are uploaded; the default central differences
        # (flowfield.py:474-504) are recomputed on the device instead of being shipped
        grad = None
        if inner.has_host_jacobian():
            grad = nat.f64(inner.grad_values)
        b = np.asarray(inner.bounds, dtype=float)
        lo = nat.f64(np.concatenate(((0.,), b[:, 0])) if not unsteady else b[:, 0])
        hi = nat.f64(np.concatenate(((0.,), b[:, 1])) if not unsteady else b[:, 1])
        if shared and grad is None:
            handle.call('dabry_set_flow_grid_shared', nat.dptr(values), nt, nx, ny, 1 if unsteady else 0,
                        nat.dptr(lo), nat.dptr(hi), C.byref(wrapper))
        else:
            handle.call('dabry_set_flow_grid', nat.dptr(values), nat.dptr(grad), nt, nx, ny, 1 if unsteady else 0,
                        nat.dptr(lo), nat.dptr(hi), C.byref(wrapper))
        # a page-locked array is uploaded in the background (streamed upload): it must outlive the first solve
        handle._keep = tuple(handle._keep) + (values,)
        return
    if not isinstance(inner, FlowField):
        raise LoweringError(f'{type(inner).__name__} is not a FlowField')
    leaves = inner.lower()
    arr = (nat.Leaf * len(leaves))()
    tables = []
    offset = 0
    for i, leaf in enumerate(leaves):
        arr[i].kind = leaf.kind
        arr[i].coeff = leaf.coeff
        arr[i].t_start = leaf.t_start
        arr[i].t_end = leaf.t_end
        for k, v in enumerate(leaf.params):
            arr[i].p[k] = v
        if leaf.table is not None:
            if leaf.table.shape[1] != 4:
                raise LoweringError('leaf time tables have 4 columns')
            arr[i].nt = getattr(leaf, 'nt', None) or leaf.table.shape[0]  # RadialGaussFFT: two stacked tables of nt rows
            arr[i].table_off = offset
            tables.append(leaf.table.reshape(-1))
            offset += leaf.table.size
    table = nat.f64(np.concatenate(tables)) if tables else None
    handle.call('dabry_set_flow_program', arr, len(leaves), nat.dptr(table), offset, C.byref(wrapper))


def apply_obstacles(handle, obstacles):
    arr = (nat.Obstacle * max(len(obstacles), 1))()
    for i, obs in enumerate(obstacles):
        kind, scale, bl, params = obs.lower()
        arr[i].kind = kind
        arr[i].scale_length = scale
        arr[i].bl[0], arr[i].bl[1] = bl
        for k, v in enumerate(params):
            arr[i].p[k] = v
    handle.call('dabry_set_obstacles', arr, len(obstacles))


def apply_penalty(handle, penalty):
    d = penalty.lower()
    pen = nat.Penalty()
    pen.kind = d['kind']
    pen.time_origin, pen.scale_time, pen.scale_length = d['time_origin'], d['scale_time'], d['scale_length']
    pen.bl[0], pen.bl[1] = d['bl']
    pen.scaler, pen.dx = d['scaler'], d['dx']
    for k, v in enumerate(d['params']):
        pen.p[k] = v
    grid = d.get('grid')
    if grid is not None:  # DiscretePenalty: a second gathered grid (data, ts, xs, ys)
        data, ts, xs, ys = (nat.f64(a) for a in grid)
        handle.call('dabry_set_penalty_grid', C.byref(pen), nat.dptr(data), nat.dptr(ts), nat.dptr(xs), nat.dptr(ys),
                    data.shape[0], data.shape[1], data.shape[2])
        return
    handle.call('dabry_set_penalty', C.byref(pen))


def coords_code(coords):
    return nat.GCS if coords == Coords.GCS else nat.CARTESIAN


def flow_evaluator(ff, device=0):
    """A native handle that holds only the flow field `ff` (analytic program or grid), for `Handle.eval_flow`:
    w and the Jacobian at arbitrary (t, x) on the device (dabry_eval_flow).  Raises LoweringError for fields outside
    the supported set and NativeError without a device - there is no host evaluation behind it."""
    cfg = nat.Config()
    cfg.abi_version = nat.ABI_VERSION
    cfg.coords = coords_code(getattr(ff, 'coords', Coords.CARTESIAN))
    cfg.n_time, cfg.max_depth, cfg.mode_origin = 2, 1, 1
    cfg.device = device
    cfg.n_index_per_subframe, cfg.trimming_band_width = 10, 3
    cfg.cost_map_nx = cfg.cost_map_ny = 2
    cfg.n_roots = cfg.root_count = 1
    cfg.max_step, cfg.rtol, cfg.atol, cfg.srf = 1., 1e-3, 1e-6, 1.
    cfg.target_radius = 1.
    times = nat.f64(np.array((0., 1.)))
    cfg.times = nat.dptr(times)
    handle = nat.Handle(cfg, keepalive=(times,))
    apply_flow(handle, ff)
    return handle


def time_upper_bound(pb, rel_timeout=10, device=0):
    """(time_orthodromic, time_htarget) of `pb` (problem.py:193-239): the two feedback flights of the time-horizon
    heuristic, integrated on the device (dabry_time_upper_bound) over the window the reference uses,
    [ff.t_start, ff.t_start + rel_timeout * length_reference / srf_max]."""
    cfg = nat.Config()
    cfg.abi_version = nat.ABI_VERSION
    cfg.coords = coords_code(pb.coords)
    cfg.n_time, cfg.max_depth, cfg.mode_origin = 2, 1, 1
    cfg.device = device
    cfg.n_index_per_subframe, cfg.trimming_band_width = 10, 3
    cfg.cost_map_nx = cfg.cost_map_ny = 2
    cfg.n_roots = cfg.root_count = 1
    cfg.max_step, cfg.rtol, cfg.atol = 1., 1e-3, 1e-6  # solve_ivp defaults (rk.py:86)
    cfg.srf = pb.srf_max
    cfg.x_target[0], cfg.x_target[1] = pb.x_target
    cfg.target_radius = pb.target_radius
    t0 = float(pb.model.ff.t_start)
    t_end = float(np.linspace(t0, t0 + rel_timeout * pb.length_reference / pb.srf_max, 1000)[-1])
    times = nat.f64(np.array((t0, t_end)))
    cfg.times = nat.dptr(times)
    handle = nat.Handle(cfg, keepalive=(times,))
    try:
        apply_flow(handle, pb.model.ff)
        return handle.time_upper_bound(pb.x_init, t0, t_end)
    finally:
        handle.close()
