"""Running-cost penalties of the drop-in API (reference: dabry/penalty.py).

The costate equation subtracts `penalty.d_value(t, x)` (problem.py:174-175); the default gradient is a central
finite difference with step 1e-8 (penalty.py:39-44), which the device reproduces operation by operation
(csrc/model.cuh `penalty_grad`).
"""
import numpy as np
from numpy import ndarray

PEN_NONE, PEN_CIRCLE, PEN_GRID = 0, 1, 2


class Penalty:
    def __init__(self, length_scale=None):
        self._dx = 1e-8 * length_scale if length_scale is not None else 1e-8

    def value(self, t, x):
        raise NotImplementedError

    def d_value(self, t, x):
        dx = self._dx
        ex, ey = dx * np.array((1, 0)), dx * np.array((0, 1))
        a1 = 1 / (2 * dx) * (self.value(t, x + ex) - self.value(t, x - ex))
        a2 = 1 / (2 * dx) * (self.value(t, x + ey) - self.value(t, x - ey))
        return np.hstack((a1, a2))

    def lower(self):
        raise TypeError(f'{type(self).__name__} has no device implementation')


class WrapperPen(Penalty):
    """Penalty seen in non-dimensional variables (penalty.py:47-69)."""

    def __init__(self, penalty: Penalty, scale_length: float, bl: ndarray, scale_time: float, time_origin: float):
        super().__init__()
        self.penalty = penalty
        self.scale_length = scale_length
        self.bl = bl.copy()
        self.scale_time = scale_time
        self.time_origin = time_origin
        self.scale_speed = scale_length / scale_time
        self._scaler_speed = 1 / self.scale_speed

    def _inner_args(self, t, x):
        return self.time_origin + t * self.scale_time, self.bl + x * self.scale_length

    def value(self, t, x):
        return self._scaler_speed * self.penalty.value(*self._inner_args(t, x))

    def d_value(self, t, x):
        return self._scaler_speed * self.penalty.d_value(*self._inner_args(t, x))

    def lower(self):
        inner = self.penalty.lower()
        if inner['kind'] != PEN_NONE and (inner['scale_length'] != 1. or inner['scaler'] != 1.):
            raise TypeError('nested WrapperPen has no device implementation')
        inner.update(time_origin=float(self.time_origin), scale_time=float(self.scale_time),
                     scale_length=float(self.scale_length), bl=(float(self.bl[0]), float(self.bl[1])),
                     scaler=float(self._scaler_speed))
        return inner


def _identity_lowering(kind, dx, params=()):
    return dict(kind=kind, time_origin=0., scale_time=1., scale_length=1., bl=(0., 0.), scaler=1., dx=float(dx),
                params=tuple(float(p) for p in params))


class NullPenalty(Penalty):
    def __init__(self):
        super().__init__()

    def value(self, t, x):
        return 0.

    def d_value(self, t, x):
        return np.array((0., 0.))

    def lower(self):
        return _identity_lowering(PEN_NONE, self._dx)


class CirclePenalty(Penalty):
    """amplitude / 2 * max(r^2 - |x - c|^2, 0) (penalty.py:84-93)."""

    def __init__(self, center: ndarray, radius: float, amplitude: float):
        super().__init__()
        self.center = center.copy()
        self.radius = radius
        self.amplitude = amplitude

    def value(self, t, x):
        return np.max((self.amplitude * 0.5 * (self.radius ** 2 - np.sum(np.square(x - self.center))), 0.))

    def lower(self):
        return _identity_lowering(PEN_CIRCLE, self._dx,
                                  (self.center[0], self.center[1], self.radius ** 2, self.amplitude))


class DiscretePenalty(Penalty):
    """Penalty sampled on a (t, x, y) grid, linear interpolation, zero outside (penalty.py:96-134).  On the device it is
    a second gathered grid (`dabry_set_penalty_grid`): scipy's RegularGridInterpolator restated with numpy's roundings,
    differentiated by the same central differences (dx = 1e-8)."""

    def __init__(self, *args):
        super().__init__()
        self.data = self.ts = self.grid = self.itp = None
        if len(args) != 0:
            if len(args) != 3:
                raise Exception('Requires 3 arguments : data, timestamps, grid')
            self.data, self.ts, self.grid = (np.array(a) for a in args)
            self._setup()

    def value(self, t, x):
        return self.itp(np.array((t, x[0], x[1])))

    def _setup(self):
        from scipy.interpolate import RegularGridInterpolator
        self.itp = RegularGridInterpolator((self.ts, self.grid[:, 0, 0].copy(), self.grid[0, :, 1].copy()), self.data,
                                           method='linear', bounds_error=False, fill_value=0.)

    def lower(self):
        if self.data is None:
            raise TypeError('DiscretePenalty without data has no device implementation')
        d = _identity_lowering(PEN_GRID, self._dx)
        d['grid'] = (np.asarray(self.data, dtype=float), np.asarray(self.ts, dtype=float),
                     np.asarray(self.grid[:, 0, 0], dtype=float), np.asarray(self.grid[0, :, 1], dtype=float))
        return d
