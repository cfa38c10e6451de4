"""Flow fields of the drop-in API (reference: dabry/flowfield.py).

Every class keeps the reference's constructor signature and its `value(t, x)` / `d_value(t, x)` semantics, evaluated
with numpy on the host for problem set-up and for the public API.  The solver hot path does NOT call these: a field
is *lowered* (``lower()``) to the flat description `libdabry_b200.so` evaluates on the device - a packed grid for
`DiscreteFF` (csrc/fields.cuh `grid_eval`) or a list of analytic leaves for an expression tree
(csrc/fields.cuh `leaf_eval`).  Fields with no device counterpart raise `LoweringError` with the way out
(`DiscreteFF.from_ff`), they are never evaluated on the CPU behind the solver's back.
"""
import sys
from math import exp, log
from typing import Optional, Union

import numpy as np
from numpy import ndarray

from dabry_b200.misc import Utils, Coords

# leaf kinds of include/dabry_b200.h (enum dabry_leaf_kind)
LEAF_UNIFORM, LEAF_STATE_LINEAR, LEAF_GYRE, LEAF_VORTEX, LEAF_RANKINE, LEAF_CHERTOVSKIH, LEAF_TRAP, \
    LEAF_RADIAL_GAUSS, LEAF_TWO_SECTORS, LEAF_LINEAR_T, LEAF_RADIAL_GAUSS_T, LEAF_BAND_GAUSS, LEAF_BAND, \
    LEAF_GYRE_MSEAS = range(14)


class LoweringError(TypeError):
    """The flow field has no device implementation."""


class Leaf:
    """One analytic term `coeff * field` of a flattened expression tree."""

    def __init__(self, kind, params=(), table=None, t_start=0., t_end=0.):
        self.kind = kind
        self.params = [float(v) for v in params]
        self.table = None if table is None else np.ascontiguousarray(table, dtype=np.float64)
        self.t_start = float(t_start)
        self.t_end = float(t_end)
        self.coeff = 1.


class FlowField:
    """Base class and node of the `+ - * float` expression tree (flowfield.py:39-150)."""

    def __init__(self, _lch=None, _rch=None, _op=None, t_start=0., t_end=None, nt_int=None,
                 coords=Coords.CARTESIAN):
        self._lch = _lch
        self._rch = _rch
        self._op = _op
        self.t_start = t_start
        self.t_end = t_end
        self.nt_int = nt_int
        self.dualizable = True
        self.coords = coords

    @property
    def is_unsteady(self):
        return self.t_end is not None

    # -- evaluation --------------------------------------------------------------------------------------
    def _combine(self, method, t, x):
        left, right, op = self._lch, self._rch, self._op
        if op == '+':
            return getattr(left, method)(t, x) + getattr(right, method)(t, x)
        if op == '-':
            return getattr(left, method)(t, x) - getattr(right, method)(t, x)
        if op == '*':
            if isinstance(left, float):
                return left * getattr(right, method)(t, x)
            if isinstance(right, float):
                return getattr(left, method)(t, x) * right
        raise Exception('Only scaling by float implemented for flow fields')

    def value(self, t, x):
        if self._lch is None:
            return self._value(t, x)
        return self._combine('value', t, x)

    def d_value(self, t, x):
        if self._lch is None or self._rch is None:
            return self._d_value(t, x)
        return self._combine('d_value', t, x)

    def _value(self, t, x):
        return None

    def _d_value(self, t, x):
        return None

    # -- algebra -----------------------------------------------------------------------------------------
    def __add__(self, other):
        if not isinstance(other, FlowField):
            raise TypeError(f"Unsupported type for addition : {type(other)}")
        return FlowField(_lch=self, _rch=other, _op='+')

    def __sub__(self, other):
        if not isinstance(other, FlowField):
            raise TypeError(f"Unsupported type for substraction : {type(other)}")
        return FlowField(_lch=self, _rch=other, _op='-')

    def __mul__(self, other):
        if isinstance(other, int):
            other = float(other)
        if not isinstance(other, float):
            raise TypeError(f"Unsupported type for multiplication : {type(other)}")
        return FlowField(_lch=self, _rch=other, _op='*', t_start=self.t_start, t_end=self.t_end)

    def __rmul__(self, other):
        return self.__mul__(other)

    def __getattr__(self, item):
        # a scaled field exposes the attributes of the field it scales (flowfield.py:145-150)
        if item in ('_lch', '_rch', '_op'):
            raise AttributeError(item)
        if type(self._lch) == float:
            return self._rch.__getattribute__(item)
        if type(self._rch) == float:
            return self._lch.__getattribute__(item)
        raise AttributeError(f'{item}')

    def dualize(self):
        """The field of the time-reversed problem (flowfield.py:100-108)."""
        if self.t_end is not None:
            self.t_start, self.t_end = self.t_end, self.t_start
        return (-1. if self.dualizable else 1.) * self

    def _index(self, t):
        """Lower frame index and interpolation coefficient of time-tabulated parameters (flowfield.py:152-175)."""
        if self.t_end is None:
            return 0, 0.
        nt = self.nt_int
        if nt == 1:
            return 0, 0.
        t_lo, t_hi = min(self.t_start, self.t_end), max(self.t_start, self.t_end)
        if t <= t_lo:
            return 0, 0.
        if t > t_hi:
            return nt - 2, 1.
        pos = (t - t_lo) / (t_hi - t_lo) * (nt - 1)
        i = int(pos)
        alpha = pos - int(pos)
        if i == nt - 1:
            return nt - 2, 1.
        return i, alpha

    # -- lowering ----------------------------------------------------------------------------------------
    def _leaf(self) -> Leaf:
        raise LoweringError(
            f'{type(self).__name__} has no device implementation; discretise it with DiscreteFF.from_ff(...)')

    def lower(self):
        """Flatten the expression tree into `[Leaf]` with signed coefficients (sum of leaves, in tree order)."""
        if self._lch is None or self._rch is None or self._op is None:
            return [self._leaf()]
        left, right, op = self._lch, self._rch, self._op
        if op == '*':
            scale, field = (left, right) if isinstance(left, float) else (right, left)
            leaves = field.lower()
            for leaf in leaves:
                leaf.coeff *= scale
            return leaves
        leaves = left.lower()
        rhs = right.lower()
        if op == '-':
            for leaf in rhs:
                leaf.coeff = -leaf.coeff
        return leaves + rhs


class DiscreteFF(FlowField):
    """Flow field sampled on a regular grid, steady (nx, ny, 2) or unsteady (nt, nx, ny, 2), evaluated by N-linear
    interpolation with clipping (flowfield.py:178-515)."""

    def __init__(self, values: ndarray, bounds: ndarray, coords: Coords, grad_values: Optional[ndarray] = None,
                 force_no_diff=False):
        super().__init__(nt_int=values.shape[0] if values.ndim == 4 else None)
        self.is_dumpable = 2
        if bounds.shape[0] != values.ndim - 1:
            raise Exception(f'Incompatible shape for values and bounds: values has {values.ndim - 1} + 1 dimensions '
                            f'so bounds must be of shape ({values.ndim - 1}, 2) ({bounds.shape} given)')
        self.values = values
        self.bounds = bounds
        self.coords = coords
        # (hi - lo) / (n - 1) per axis (flowfield.py:198-199)
        self.spacings = (bounds[:, 1] - bounds[:, 0]) / (np.array(values.shape[:-1]) - np.ones(values.ndim - 1))
        self._grad_values = None
        self._grad_is_default = False
        self._grad_fingerprint = None
        self.grad_values = grad_values
        # The reference differentiates the samples in its constructor (flowfield.py:209-210).  Here the device does that
        # at upload (PackGridPhase), so the host copy of the default Jacobian - 2 x the size of `values` - is only
        # computed when somebody reads `grad_values` on the host (d_value, save, the oracle): same values, on demand.
        self._grad_lazy = not force_no_diff and grad_values is None
        if values.ndim == 3:
            self.value = self._value_steady
            self.d_value = self._d_value_steady
        else:
            self.value = self._value_unsteady
            self.d_value = self._d_value_unsteady
            self.t_start = bounds[0, 0]
            self.t_end = bounds[0, 1]

    @property
    def grad_values(self):
        if self._grad_values is None and getattr(self, '_grad_lazy', False):
            self._grad_lazy = False
            self.compute_derivatives()
        return self._grad_values

    @grad_values.setter
    def grad_values(self, value):
        # a caller-supplied Jacobian is uploaded as it is; only the central differences of compute_derivatives are
        # recomputed on the device instead of being shipped (dabry_b200/_lowering.py)
        self._grad_values = value
        self._grad_is_default = False
        self._grad_fingerprint = None
        if value is not None:
            self._grad_lazy = False

    @staticmethod
    def _fingerprint(a):
        flat = a.reshape(-1)
        return hash(flat[::max(1, flat.shape[0] // 4096)].tobytes())

    def grad_is_device_default(self):
        """True when `grad_values` is still exactly what `compute_derivatives` produced (same array, not edited in
        place as far as a strided fingerprint of 4 096 entries can tell): the device then differentiates by itself."""
        if self._grad_values is None:
            return bool(getattr(self, '_grad_lazy', False))  # never materialised on the host: the device's own
        return bool(self._grad_is_default and self._grad_fingerprint == self._fingerprint(self._grad_values))

    def has_host_jacobian(self):
        """True when a caller-supplied (or edited) Jacobian must be uploaded; does not trigger the lazy default."""
        return self._grad_values is not None and not self.grad_is_device_default()

    @property
    def times(self):
        if self.t_end is None:
            raise ValueError('Flowfield is steady')
        return np.linspace(self.t_start, self.t_end, self.values.shape[0])

    def pin_memory(self):
        """Page-lock `values` (and host-supplied `grad_values`) in place so that every upload of this field to a GPU is
        one DMA at the PCIe rate instead of a driver-staged pageable copy (no reference counterpart; the arrays stay
        ordinary numpy arrays).  Returns self."""
        from dabry_b200 import _native as nat
        if getattr(self, '_pins', None):
            return self
        self.values = np.ascontiguousarray(self.values, dtype=np.float64)
        if not self.values.flags['WRITEABLE']:
            self.values = self.values.copy()
        pins = [nat.PinnedArray(self.values)]
        if self.has_host_jacobian():
            self.grad_values = np.array(self._grad_values, dtype=np.float64, order='C', copy=None)
            pins.append(nat.PinnedArray(self._grad_values))
        self._pins = pins
        return self

    @classmethod
    def from_npz(cls, filepath, pin=False):
        """Flow file of docs/flow_format.md: arrays `values`, `bounds`, `coords` (flowfield.py:222-225).
        pin=True (no reference counterpart): the samples are read straight into a contiguous float64 array that is
        page-locked in place, i.e. ready for the streamed upload of `dabry_set_flow_grid`; the Jacobian is left to the
        device and only materialised on the host if `grad_values` is read."""
        with np.load(filepath) as data:
            values = np.ascontiguousarray(data['values'], dtype=np.float64)
            ff = cls(values, data['bounds'], Coords.from_string(data['coords']))
        return ff.pin_memory() if pin else ff

    @classmethod
    def from_h5(cls, filepath, **kwargs):
        """Legacy h5 wind file (flowfield.py:227-255).  Needs h5py."""
        try:
            import h5py
        except ImportError as exc:
            raise ImportError('DiscreteFF.from_h5 needs the optional dependency h5py') from exc
        with h5py.File(filepath, 'r') as f:
            coords = Coords.from_string(f.attrs['coords'])
            values = np.array(f['data']).squeeze()
            ts = np.array(f['ts'], dtype=float)
            if np.any(ts > 1e11):  # millisecond time stamps
                ts = ts / 1000.
            from dabry_b200.misc import Units
            factor = Utils.DEG_TO_RAD if Units.from_string(f.attrs['units_grid']) == Units.DEGREES else 1.
            rows = [] if ts.shape[0] == 1 else [np.array((ts[0], ts[-1]))]
            lo, hi = factor * np.array(f['grid'][0, 0]), factor * np.array(f['grid'][-1, -1])
            rows += [np.array((lo[0], hi[0])), np.array((lo[1], hi[1]))]
        return cls(values, np.stack(rows, axis=0), coords, **kwargs)

    @classmethod
    def from_ff(cls, ff: FlowField, grid_bounds: Union[tuple, ndarray], nx=100, ny=100, nt=50,
                coords: Optional[Coords] = None, device: Optional[int] = None, **kwargs):
        """Sample another flow field (values and Jacobian) on a regular grid (flowfield.py:257-309).

        device=None runs the reference's node-by-node Python loop.  device=<ordinal> evaluates all nodes in one
        `dabry_eval_flow` launch on that GPU (the field must be lowerable, no silent fall-back):
        node positions are the same expressions, values agree with the loop to 1e-12 relative (device libm)."""
        if isinstance(grid_bounds, tuple):
            if len(grid_bounds) != 2:
                raise ValueError('"grid_bounds" provided as a tuple must have two elements')
            grid_bounds = np.array(grid_bounds).transpose()
        if isinstance(grid_bounds, ndarray) and grid_bounds.shape != (2, 2):
            raise ValueError('"grid_bounds" provided as an array must have shape (2, 2)')
        steady = ff.t_end is None
        rows = ([] if steady else [np.array((ff.t_start, ff.t_end))]) + [grid_bounds[0], grid_bounds[1]]
        bounds = np.stack(rows, axis=0)
        shape = (() if steady else (nt,)) + (nx, ny)
        spacings = (bounds[:, 1] - bounds[:, 0]) / (np.array(shape) - np.ones(bounds.shape[0]))
        no_diff = bool(kwargs.get('force_no_diff'))
        if coords is None:
            coords = Coords.GCS if getattr(ff, 'coords', None) == Coords.GCS else Coords.CARTESIAN
        if device is not None:
            from dabry_b200 import _lowering as low
            ks = np.arange(1 if steady else nt)
            ts = np.full(1, ff.t_start if ff.t_start is not None else 0.) if steady else bounds[0, 0] + ks * spacings[0]
            xs = bounds[-2, 0] + np.arange(nx) * spacings[-2]
            ys = bounds[-1, 0] + np.arange(ny) * spacings[-1]
            tt, xx, yy = np.meshgrid(ts, xs, ys, indexing='ij')
            handle = low.flow_evaluator(ff, device)
            try:
                w, jac = handle.eval_flow(tt.reshape(-1), np.stack((xx.reshape(-1), yy.reshape(-1)), -1))
            finally:
                handle.close()
            values = w.reshape(shape + (2,))
            grad_values = None if no_diff else jac.reshape(shape + (2, 2))
            return cls(values, bounds, coords, grad_values=grad_values, **kwargs)
        values = np.zeros(shape + (2,))
        grad_values = None if no_diff else np.zeros(shape + (2, 2))
        for k in range(1 if steady else nt):
            t = ff.t_start if steady else bounds[0, 0] + k * spacings[0]
            for i in range(nx):
                for j in range(ny):
                    state = bounds[-2:, 0] + np.diag((i, j)) @ spacings[-2:]
                    node = (i, j) if steady else (k, i, j)
                    values[node] = ff.value(t, state)
                    if not no_diff:
                        grad_values[node] = ff.d_value(t, state)
        return cls(values, bounds, coords, grad_values=grad_values, **kwargs)

    def _value_steady(self, _, x):
        return Utils.interpolate(self.values, self.bounds[:, 0], self.spacings, np.asarray(x))

    def _value_unsteady(self, t, x):
        return Utils.interpolate(self.values, self.bounds[:, 0], self.spacings, np.array((t,) + tuple(x)))

    def _d_value_steady(self, _, x):
        return Utils.interpolate(self.grad_values, self.bounds[:, 0], self.spacings, np.asarray(x), ndim_values_data=2)

    def _d_value_unsteady(self, t, x):
        return Utils.interpolate(self.grad_values, self.bounds[:, 0], self.spacings, np.array((t,) + tuple(x)),
                                 ndim_values_data=2)

    def compute_derivatives(self):
        """Second-order central differences on the native grid; border lines copy the adjacent interior line and
        corners the diagonal interior node (flowfield.py:474-504)."""
        v = self.values
        dx, dy = self.spacings[-2], self.spacings[-1]
        g = np.zeros(v.shape + (2,))
        inner = g[..., 1:-1, 1:-1, :, :]
        inner[..., 0, 0] = (v[..., 2:, 1:-1, 0] - v[..., :-2, 1:-1, 0]) / (2 * dx)
        inner[..., 0, 1] = (v[..., 1:-1, 2:, 0] - v[..., 1:-1, :-2, 0]) / (2 * dy)
        inner[..., 1, 0] = (v[..., 2:, 1:-1, 1] - v[..., :-2, 1:-1, 1]) / (2 * dx)
        inner[..., 1, 1] = (v[..., 1:-1, 2:, 1] - v[..., 1:-1, :-2, 1]) / (2 * dy)
        g[..., 0, 1:-1, :, :] = g[..., 1, 1:-1, :, :]
        g[..., -1, 1:-1, :, :] = g[..., -2, 1:-1, :, :]
        g[..., 1:-1, 0, :, :] = g[..., 1:-1, 1, :, :]
        g[..., 1:-1, -1, :, :] = g[..., 1:-1, -2, :, :]
        for a, b in ((0, 1), (-1, -2)):
            for c, d in ((0, 1), (-1, -2)):
                g[..., a, c, :, :] = g[..., b, d, :, :]
        self._grad_values = g
        self._grad_is_default = True
        self._grad_fingerprint = self._fingerprint(g)

    def dualize(self):
        ff = DiscreteFF.from_ff(-1. * self, self.bounds, self.values.shape[1], self.values.shape[2])
        if self.t_end is not None:
            ff.t_start, ff.t_end = self.t_end, self.t_start
        else:
            ff.t_start = self.t_start
        return ff

    def lower(self):
        raise LoweringError('DiscreteFF lowers to a grid, not to leaves')


class WrapperFF(FlowField):
    """Non-dimensionalising view of a flow field (flowfield.py:518-562): w'(t, x) = w(t0 + t S_t, bl + x S_l) / S_v."""

    def __init__(self, ff: FlowField, scale_length: float, bl: ndarray, scale_time: float, time_origin: float):
        self.ff = ff
        self.scale_length = scale_length
        self.bl = bl.copy()
        self.scale_time = scale_time
        self.time_origin = time_origin
        # GCS winds are in metres per second and become radians per second first (flowfield.py:529-531)
        self.scale_speed = scale_length / scale_time * (Utils.EARTH_RADIUS if ff.coords == Coords.GCS else 1.)
        self._scaler_speed = 1 / self.scale_speed
        self._scaler_dspeed = self.scale_length / self.scale_speed
        super().__init__(coords=ff.coords, t_start=(ff.t_start - self.time_origin) / self.scale_time,
                         t_end=None if ff.t_end is None else (ff.t_end - self.time_origin) / self.scale_time)

    def value(self, t, x):
        return self._scaler_speed * self.ff.value(self.time_origin + t * self.scale_time,
                                                  self.bl + x * self.scale_length)

    def d_value(self, t, x):
        return self._scaler_dspeed * self.ff.d_value(self.time_origin + t * self.scale_time,
                                                     self.bl + x * self.scale_length)

    def __getattr__(self, item):
        if item in ('ff', '_lch', '_rch', '_op'):
            raise AttributeError(item)
        if item == 'values':
            return self._scaler_speed * self.ff.values
        if item == 'bounds':
            inner = self.ff.bounds
            space = np.vstack(((0., (inner[-2, 1] - self.bl[0]) / self.scale_length),
                               (0., (inner[-1, 1] - self.bl[1]) / self.scale_length)))
            if self.ff.t_end is None:
                return space
            window = ((self.ff.t_start - self.time_origin) / self.scale_time,
                      (self.ff.t_end - self.time_origin) / self.scale_time)
            return np.vstack((window, space))
        return self.ff.__getattribute__(item)

    def lower(self):
        return self.ff.lower()


# ---- analytic fields ------------------------------------------------------------------------------------------

class TwoSectorsFF(FlowField):
    """Constant cross flow on each side of x = x_switch (flowfield.py:565-588)."""

    def __init__(self, v_w1: float, v_w2: float, x_switch: float):
        super().__init__()
        self.v_w1 = v_w1
        self.v_w2 = v_w2
        self.x_switch = x_switch

    def value(self, t, x):
        return np.array([0, self.v_w1 * np.heaviside(self.x_switch - x[0], 0.)
                         + self.v_w2 * np.heaviside(x[0] - self.x_switch, 0.)])

    def d_value(self, t, x):
        return np.array([[0, 0], [0, 0]])

    def _leaf(self):
        return Leaf(LEAF_TWO_SECTORS, (self.v_w1, self.v_w2, self.x_switch))


class TSEqualFF(TwoSectorsFF):
    def __init__(self, v_w1, v_w2, x_f):
        super().__init__(v_w1, v_w2, x_f / 2)


class UniformFF(FlowField):
    def __init__(self, ff_val: ndarray):
        super().__init__()
        self.ff_val = ff_val.copy()

    def _value(self, t, x):
        return self.ff_val

    def _d_value(self, t, x):
        return np.zeros((2, 2))

    def _leaf(self):
        return Leaf(LEAF_UNIFORM, self.ff_val)


class ZeroFF(UniformFF):
    def __init__(self):
        super().__init__(np.zeros(2))


def _point_vortex_jacobian(dx, dy, factor):
    """factor * [[2 dx dy, dy^2 - dx^2], [dy^2 - dx^2, -2 dx dy]]"""
    return factor * np.array([[2 * dx * dy, dy ** 2 - dx ** 2],
                              [dy ** 2 - dx ** 2, -2 * dx * dy]])


class VortexFF(FlowField):
    """Potential-flow point vortex (flowfield.py:623-649)."""

    def __init__(self, center: ndarray, circulation: float):
        super().__init__()
        self.center = center.copy()
        self.circulation = circulation

    def value(self, t, x):
        d = x - self.center
        r = np.linalg.norm(d)
        e_theta = np.array([-d[1] / r, d[0] / r])
        return self.circulation / (2 * np.pi * r) * e_theta

    def d_value(self, t, x):
        r = np.linalg.norm(x - self.center)
        return _point_vortex_jacobian(x[0] - self.center[0], x[1] - self.center[1],
                                      self.circulation / (2 * np.pi * r ** 4))

    def _leaf(self):
        return Leaf(LEAF_VORTEX, (self.center[0], self.center[1], self.circulation))


class RankineVortexFF(FlowField):
    """Rankine vortex: solid-body core of radius `radius`, potential flow outside; parameters may be tabulated in
    time over [t_start, t_end] (flowfield.py:652-727)."""

    def __init__(self, center: Union[float, ndarray], circulation: Union[float, ndarray],
                 radius: Union[float, ndarray], t_start=0., t_end=None):
        nt_int = center.shape[0] if len(center.shape) > 1 else None
        if nt_int is not None and t_end is None:
            print('Missing t_end', file=sys.stderr)
            exit(1)
        super().__init__(t_start=t_start, t_end=t_end, nt_int=nt_int)
        self.center = np.array(center)
        self.circulation = np.array(circulation)
        self.radius = np.array(radius)
        if nt_int is not None and not (self.center.shape[0] == self.circulation.shape[0] == self.radius.shape[0]):
            print('Incoherent sizes', file=sys.stderr)
            exit(1)
        self.zero_ceil = 1e-3

    def max_speed(self, t=0.):
        _, gamma, radius = self.params(t)
        return gamma / (radius * 2 * np.pi)

    def params(self, t):
        if self.t_end is None:
            return self.center, self.circulation, self.radius
        i, alpha = self._index(t)
        return ((1 - alpha) * self.center[i] + alpha * self.center[i + 1],
                (1 - alpha) * self.circulation[i] + alpha * self.circulation[i + 1],
                (1 - alpha) * self.radius[i] + alpha * self.radius[i + 1])

    def value(self, t, x):
        omega, gamma, radius = self.params(t)
        d = x - omega
        r = np.linalg.norm(d)
        if r < self.zero_ceil * radius:
            return np.zeros(2)
        e_theta = np.array([-d[1] / r, d[0] / r])
        f = gamma / (2 * np.pi)
        if r <= radius:
            return f * r / (radius ** 2) * e_theta
        return f * 1 / r * e_theta

    def d_value(self, t, x):
        omega, gamma, radius = self.params(t)
        d = x - omega
        r = np.linalg.norm(d)
        if r < self.zero_ceil * radius:
            return np.zeros((2, 2))
        f = gamma / (2 * np.pi)
        if r <= radius:
            a, b = d[0], d[1]
            e_theta = np.array([-b / r, a / r])
            de_dx = np.array([a * b / r ** 3, b ** 2 / r ** 3])
            de_dy = np.array([-a ** 2 / r ** 3, -a * b / r ** 3])
            return f / radius ** 2 * np.stack((a / r * e_theta + r * de_dx, b / r * e_theta + r * de_dy), axis=1)
        return _point_vortex_jacobian(x[0] - omega[0], x[1] - omega[1], f / r ** 4)

    def _leaf(self):
        if self.t_end is None:
            return Leaf(LEAF_RANKINE, (self.center[0], self.center[1], float(self.circulation), float(self.radius)))
        table = np.column_stack((self.center, self.circulation, self.radius))
        return Leaf(LEAF_RANKINE, table=table, t_start=self.t_start, t_end=self.t_end)


class StateLinearFF(FlowField):
    """w = gradient @ (x - origin) + value_origin (flowfield.py:730-752)."""

    def __init__(self, gradient: ndarray, origin: ndarray, value_origin: ndarray):
        super().__init__()
        self.gradient = gradient.copy()
        self.origin = origin.copy()
        self.value_origin = value_origin.copy()

    def value(self, _, x):
        return self.gradient.dot(x - self.origin) + self.value_origin

    def d_value(self, _, x):
        return self.gradient

    def _leaf(self):
        return Leaf(LEAF_STATE_LINEAR, tuple(self.gradient.reshape(-1)) + tuple(self.origin) + tuple(self.value_origin))


class LinearFFT(FlowField):
    """Linear in space with a gradient that varies linearly in time (flowfield.py:755-780).  The reference's `value`
    has no return statement, which makes the class unusable in a solver; this one returns the documented value."""

    def __init__(self, gradient_diff_time: ndarray, gradient_init: ndarray, origin: ndarray, value_origin: ndarray):
        super().__init__()
        self.gradient_diff_time = gradient_diff_time.copy()
        self.gradient_init = gradient_init.copy()
        self.origin = origin.copy()
        self.value_origin = value_origin.copy()

    def value(self, t, x):
        return (self.gradient_diff_time * t + self.gradient_init) @ (x - self.origin) + self.value_origin

    def d_value(self, t, x):
        return self.gradient_diff_time * t + self.gradient_init

    def _leaf(self):
        table = np.concatenate((np.asarray(self.gradient_diff_time, dtype=float).reshape(-1),
                                np.asarray(self.gradient_init, dtype=float).reshape(-1),
                                np.asarray(self.origin, dtype=float), np.asarray(self.value_origin, dtype=float)))
        return Leaf(LEAF_LINEAR_T, table=table.reshape(3, 4))


class PointSymFF(FlowField):
    """Point-symmetric linear field with divergence gamma and curl omega, Techy 2011 (flowfield.py:783-804)."""

    def __init__(self, center, gamma: float, omega: float):
        super().__init__()
        self.center = center.copy() if isinstance(center, ndarray) else np.array(center)
        self.mat = np.array([[gamma, -omega], [omega, gamma]])

    def value(self, t, x):
        return self.mat @ (x - self.center)

    def d_value(self, t, x):
        return self.mat

    def _leaf(self):
        return Leaf(LEAF_STATE_LINEAR, tuple(self.mat.reshape(-1)) + tuple(self.center) + (0., 0.))


class GyreFF(FlowField):
    """Cellular gyre w = A (-sin(kx X) cos(ky Y), cos(kx X) sin(ky Y)), (X, Y) = x - center (flowfield.py:807-833)."""

    def __init__(self, x_center: float, y_center: float, x_wl: float, y_wl: float, ampl: float):
        super().__init__()
        self.center = np.array((x_center, y_center))
        self.kx = 2 * np.pi / x_wl
        self.ky = 2 * np.pi / y_wl
        self.ampl = ampl

    def _phase(self, x):
        return np.diag((self.kx, self.ky)) @ (x - self.center)

    def value(self, t, x):
        a, b = self._phase(x)
        return self.ampl * np.array((-np.sin(a) * np.cos(b), np.cos(a) * np.sin(b)))

    def d_value(self, t, x):
        a, b = self._phase(x)
        return self.ampl * np.array([[-self.kx * np.cos(a) * np.cos(b), self.ky * np.sin(a) * np.sin(b)],
                                     [-self.kx * np.sin(a) * np.sin(b), self.ky * np.cos(a) * np.cos(b)]])

    def _leaf(self):
        return Leaf(LEAF_GYRE, (self.center[0], self.center[1], self.kx, self.ky, self.ampl))


class DoubleGyreDampedFF(FlowField):
    """Gyre damped by 1 / (1 + |D^-1 (x - c)|^2) (flowfield.py:836-857; the reference defines no Jacobian)."""

    def __init__(self, x_center, y_center, x_wl, y_wl, ampl, x_center_damp, y_center_damp, lambda_x_damp,
                 lambda_y_damp):
        super().__init__()
        self.double_gyre = GyreFF(x_center, y_center, x_wl, y_wl, ampl)
        self.center_damp = np.array((x_center_damp, y_center_damp))
        self.lambda_x_damp = lambda_x_damp
        self.lambda_y_damp = lambda_y_damp

    def value(self, t, x):
        xx = (x - self.center_damp) / np.array((self.lambda_x_damp, self.lambda_y_damp))
        return self.double_gyre.value(t, x) * (1 / (1 + xx[0] ** 2 + xx[1] ** 2))


def _radial_gauss_value(x, center, radius, sdev, v_max, zero_ceil):
    d = x - center
    r = np.linalg.norm(d)
    if r < zero_ceil * radius:
        return np.zeros(2)
    return v_max * exp(-(log(r / radius)) ** 2 / (2 * sdev ** 2)) * (d / r)


def _radial_gauss_jacobian(x, center, radius, sdev, v_max, zero_ceil):
    d = x - center
    r = np.linalg.norm(d)
    if r < zero_ceil * radius:
        return np.zeros((2, 2))
    e_r = d / r
    a, b = d
    amp = v_max * exp(-(log(r / radius)) ** 2 / (2 * sdev ** 2))
    dv = -log(r / radius) * amp / (r ** 2 * sdev ** 2) * np.array([d[0], d[1]])
    nabla_e_r = np.array([[b ** 2 / r ** 3, -a * b / r ** 3], [-a * b / r ** 3, a ** 2 / r ** 3]])
    return np.einsum('i,j->ij', e_r, dv) + amp * nabla_e_r


class RadialGaussFF(FlowField):
    """Radial outflow with a log-normal profile peaking at `radius` (flowfield.py:860-903)."""

    def __init__(self, center: ndarray, radius: float, sdev: float, v_max: float):
        super().__init__()
        self.center = center.copy()
        self.radius = radius
        self.sdev = sdev
        self.v_max = v_max
        self.zero_ceil = 1e-3
        self.dualizable = False

    def ampl(self, r):
        if r < self.zero_ceil * self.radius:
            return 0.
        return self.v_max * exp(-(log(r / self.radius)) ** 2 / (2 * self.sdev ** 2))

    def value(self, t, x):
        return _radial_gauss_value(x, self.center, self.radius, self.sdev, self.v_max, self.zero_ceil)

    def d_value(self, t, x):
        return _radial_gauss_jacobian(x, self.center, self.radius, self.sdev, self.v_max, self.zero_ceil)

    def _leaf(self):
        return Leaf(LEAF_RADIAL_GAUSS, (self.center[0], self.center[1], self.radius, self.sdev, self.v_max))


class RadialGaussFFT(FlowField):
    """Time-tabulated radial Gauss field (flowfield.py:906-973)."""

    def __init__(self, center, radius, sdev, v_max, t_end, t_start=0.):
        super().__init__(t_start=t_start, t_end=t_end, nt_int=radius.shape[0])
        self.center = np.array(center, dtype=float)
        self.radius = np.array(radius, dtype=float)
        self.sdev = np.array(sdev, dtype=float)
        self.v_max = np.array(v_max, dtype=float)
        self.zero_ceil = 1e-3
        self.is_analytical = True
        self.dualizable = False

    def _params(self, t):
        i, alpha = self._index(t)
        mix = lambda arr: (1 - alpha) * arr[i] + alpha * arr[i + 1]
        return mix(self.center), mix(self.radius), mix(self.sdev), mix(self.v_max)

    def value(self, t, x):
        center, radius, sdev, v_max = self._params(t)
        return _radial_gauss_value(x, center, radius, sdev, v_max, self.zero_ceil)

    def d_value(self, t, x):
        center, radius, sdev, v_max = self._params(t)
        return _radial_gauss_jacobian(x, center, radius, sdev, v_max, self.zero_ceil)

    def _leaf(self):
        nt = self.radius.shape[0]
        table = np.zeros((2 * nt, 4))
        table[:nt, 0:2] = self.center
        table[:nt, 2] = self.radius
        table[:nt, 3] = self.sdev
        table[nt:, 0] = self.v_max
        leaf = Leaf(LEAF_RADIAL_GAUSS_T, table=table, t_start=self.t_start, t_end=self.t_end)
        leaf.nt = nt
        return leaf


def _fd_jacobian_central(value, t, x, eps):
    ex, ey = eps * np.array((1, 0)), eps * np.array((0, 1))
    return np.column_stack(((value(t, x + ex) - value(t, x - ex)) / (2 * eps),
                            (value(t, x + ey) - value(t, x - ey)) / (2 * eps)))


class BandGaussFF(FlowField):
    """Gaussian jet along `vect` through `origin`; forward-difference Jacobian, step 1e-6 (flowfield.py:976-999)."""

    def __init__(self, origin, vect, ampl, sdev):
        super().__init__()
        self.origin = np.array(origin, dtype=float)
        self.vect = np.asarray(vect, dtype=float) / np.linalg.norm(vect)
        self.ampl = ampl
        self.sdev = sdev

    def value(self, t, x):
        d = x - self.origin
        dist = np.abs(self.vect[0] * d[1] - self.vect[1] * d[0])
        return self.vect * (self.ampl * np.exp(-0.5 * (dist / self.sdev) ** 2))

    def d_value(self, t, x):
        dx = 1e-6
        return np.column_stack((1 / dx * (self.value(t, x + np.array((dx, 0.))) - self.value(t, x)),
                                1 / dx * (self.value(t, x + np.array((0., dx))) - self.value(t, x))))

    def _leaf(self):
        return Leaf(LEAF_BAND_GAUSS, (self.origin[0], self.origin[1], self.vect[0], self.vect[1], self.ampl, self.sdev))


class BandFF(FlowField):
    """Uniform jet of width `width` with linear edges; central-difference Jacobian, step 1e-6
    (flowfield.py:1002-1032)."""

    def __init__(self, origin, vect, w_value, width):
        super().__init__()
        self.origin = np.array(origin, dtype=float)
        self.vect = np.asarray(vect, dtype=float) / np.linalg.norm(vect)
        self.w_value = w_value
        self.width = width
        self.transition_width = width / 10

    def value(self, t, x):
        d = x - self.origin
        dist = np.abs(self.vect[0] * d[1] - self.vect[1] * d[0])
        outer = (self.width + self.transition_width / 2) / 2
        inner = (self.width - self.transition_width / 2) / 2
        if dist >= outer:
            return np.zeros(2)
        if dist >= inner:
            return (1 - (dist - inner) / self.transition_width) * self.w_value
        return self.w_value

    def d_value(self, t, x):
        return _fd_jacobian_central(self.value, t, x, 1e-6)

    def _leaf(self):
        w = np.asarray(self.w_value, dtype=float)
        if w.shape != (2,):
            raise LoweringError('BandFF on the device needs a 2-vector w_value')
        return Leaf(LEAF_BAND, (self.origin[0], self.origin[1], self.vect[0], self.vect[1], w[0], w[1], self.width))


class TrapFF(FlowField):
    """Inward radial flow switched on by a sigmoid at distance `radius` from a moving centre
    (flowfield.py:1035-1086)."""

    def __init__(self, ff_val: ndarray, center: ndarray, radius: ndarray, rel_wid=0.05, t_start=0.,
                 t_end: Optional[float] = None):
        if ff_val.shape[0] != radius.shape[0]:
            raise Exception('Incoherent shapes')
        super().__init__(t_start=t_start, t_end=t_end, nt_int=ff_val.shape[0])
        self.ff_val = ff_val.copy()
        self.center = center.copy()
        self.radius = radius.copy()
        self.rel_wid = rel_wid

    @staticmethod
    def sigmoid(r, wid):
        return 1 / (1 + np.exp(-4 / wid * r))

    @staticmethod
    def d_sigmoid(r, wid):
        s = TrapFF.sigmoid(r, wid)
        return 4 / wid * s * (1 - s)

    def _params(self, t):
        i, alpha = self._index(t)
        # the upper frame is dropped (not weighted) when alpha < 1e-3 (flowfield.py:1064-1066)
        mix = lambda arr: (1 - alpha) * arr[i] + (0. if alpha < 1e-3 else alpha * arr[i + 1])
        return mix(self.ff_val), mix(self.center), mix(self.radius)

    def value(self, t, x):
        ff_val, center, radius = self._params(t)
        d = x - center
        if d @ d < 1e-8 * radius ** 2:
            return np.zeros(2)
        r = np.linalg.norm(d)
        return -ff_val * TrapFF.sigmoid(r - radius, radius * self.rel_wid) * (d / r)

    def d_value(self, t, x):
        ff_val, center, radius = self._params(t)
        d = x - center
        if d @ d < 1e-8 * radius ** 2:
            return np.zeros((2, 2))
        r = np.linalg.norm(d)
        e_r = d / r
        frame = np.array(((e_r[0], e_r[1]), (-e_r[1] / r, e_r[0] / r)))
        wid = radius * self.rel_wid
        return -ff_val * (frame.transpose() @ np.diag((TrapFF.d_sigmoid(r - radius, wid),
                                                       TrapFF.sigmoid(r - radius, wid))) @ frame)

    def _leaf(self):
        table = np.column_stack((self.ff_val, self.center, self.radius))
        t_end = self.t_end if self.t_end is not None else self.t_start
        leaf = Leaf(LEAF_TRAP, (self.rel_wid,), table=table, t_start=self.t_start, t_end=t_end)
        if self.t_end is None:
            leaf.table = leaf.table[:1]
        return leaf


class ChertovskihFF(FlowField):
    """w = (x (y + 3) / 4, -x^2) (flowfield.py:1089-1099)."""

    def __init__(self):
        super().__init__()

    def value(self, t, x):
        return np.array((x[0] * (x[1] + 3) / 4, -x[0] * x[0]))

    def d_value(self, t, x):
        return np.array((((x[1] + 3) / 4, x[0] / 4), (-2 * x[0], 0)))

    def _leaf(self):
        return Leaf(LEAF_CHERTOVSKIH)


class GyreMSEASFF(FlowField):
    """Time-dependent double gyre of the LCS tutorial; central-difference Jacobian, step 1e-6
    (flowfield.py:1102-1134)."""

    def __init__(self):
        super().__init__(t_end=2)
        self.A = 1
        self.eps = 0.1
        self.T = 0.25

    def _a(self, t):
        return self.eps * np.sin(2 * np.pi * t / self.T)

    def _b(self, t):
        return 1 - 2 * self._a(t)

    def _f(self, t, x):
        return self._a(t) * x ** 2 + self._b(t) * x

    def _dfdx(self, t, x):
        return 2 * self._a(t) * x + self._b(t)

    def value(self, t, x):
        f = self._f(t, x[0])
        return np.array((-np.pi * self.A * np.sin(np.pi * f) * np.cos(np.pi * x[1]),
                         np.pi * self.A * np.cos(np.pi * f) * np.sin(np.pi * x[1]) * self._dfdx(t, x[0])))

    def d_value(self, t, x):
        return _fd_jacobian_central(self.value, t, x, 1e-6)

    def _leaf(self):
        return Leaf(LEAF_GYRE_MSEAS, (self.A, self.eps, self.T))


def discretize_ff(ff: FlowField, nx=None, ny=None, nt=None, bl=None, tr=None):
    """A `DiscreteFF` view of any field, values only (flowfield.py:1137-1154)."""
    if hasattr(ff, 'bounds'):
        return ff
    if bl is None or tr is None:
        raise Exception(f'Missing bounding box (bl, tr) to sample unbounded {ff}')
    return DiscreteFF.from_ff(ff, np.array((bl, tr)).transpose(), nx=50 if nx is None else nx,
                              ny=50 if ny is None else ny, nt=25 if nt is None else nt, force_no_diff=True)


def save_ff(ff: FlowField, filepath: str, fmt='npz', nx=None, ny=None, nt=None, bl=None, tr=None):
    """Write the `ff.npz` flow file (`values`, `bounds`, `coords`; docs/flow_format.md, flowfield.py:1157-1170)."""
    if fmt != 'npz':
        raise Exception(f'Unsupported output format "{fmt}" (npz only)')
    dff = discretize_ff(ff, nx, ny, nt, bl, tr)
    np.savez(filepath, values=dff.values, bounds=dff.bounds, coords=np.array(dff.coords.value))
