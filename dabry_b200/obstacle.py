"""Obstacles of the drop-in API (reference: dabry/obstacle.py): scalar fields negative inside, with a gradient.

`lower()` gives the device description (include/dabry_b200.h `dabry_obstacle`); obstacles without one raise.
"""
from abc import ABC, abstractmethod
from typing import Union

import numpy as np
from numpy import ndarray

from dabry_b200.misc import terminal, Utils

OBS_CIRCLE, OBS_FRAME, OBS_GREAT_CIRCLE = 0, 1, 2


class Obstacle(ABC):
    def __init__(self):
        pass

    @terminal
    def event(self, time: float, state_aug: ndarray):
        """scipy event: the obstacle value at the position part of the augmented state (obstacle.py:37-39)."""
        return self.value(state_aug[:2])

    @abstractmethod
    def value(self, x: ndarray):
        """Negative inside the obstacle, zero on its border."""

    def d_value(self, x: ndarray) -> ndarray:
        """Central finite differences with step 1e-8 (obstacle.py:50-63)."""
        eps = 1e-8
        n = x.shape[0]
        steps = np.diag(n * (eps,))
        return np.array([(self.value(x + steps[i]) - self.value(x - steps[i])) / (2 * eps) for i in range(n)])

    def lower(self):
        raise TypeError(f'{type(self).__name__} has no device implementation')


class WrapperObs(Obstacle):
    """Obstacle seen in non-dimensional coordinates (obstacle.py:66-81); no chain-rule factor on the gradient."""

    def __init__(self, obs: Obstacle, scale_length: float, bl: ndarray):
        super().__init__()
        self.obs = obs
        self.scale_length = scale_length
        self.bl = bl.copy()

    def value(self, x):
        return self.obs.value(self.bl + x * self.scale_length)

    def d_value(self, x):
        return self.obs.d_value(self.bl + x * self.scale_length)

    def lower(self):
        kind, scale, bl, params = self.obs.lower()
        if scale != 1. or np.any(np.asarray(bl) != 0.):
            raise TypeError('nested WrapperObs has no device implementation')
        return kind, float(self.scale_length), (float(self.bl[0]), float(self.bl[1])), params


class CircleObs(Obstacle):
    """Disc: value = (|x - c|^2 - r^2) / 2, gradient x - c (obstacle.py:84-99)."""

    def __init__(self, center: Union[ndarray, tuple], radius: float):
        self.center = center.copy() if isinstance(center, ndarray) else np.array(center)
        self.radius = radius
        self._sqradius = radius ** 2
        super().__init__()

    def value(self, x: ndarray):
        return 0.5 * (np.sum(np.square(x - self.center)) - self._sqradius)

    def d_value(self, x: ndarray) -> ndarray:
        return x - self.center

    def lower(self):
        return OBS_CIRCLE, 1., (0., 0.), (float(self.center[0]), float(self.center[1]), float(self._sqradius))


class FrameObs(Obstacle):
    """The outside of the box [bl, tr]: value = 1 - max_i |x_i - c_i| / h_i (obstacle.py:102-128)."""

    def __init__(self, bl: ndarray, tr: ndarray):
        self.bl = bl.copy()
        self.tr = tr.copy()
        self.center = 0.5 * (bl + tr)
        self.scaler = np.diag(1 / (0.5 * (self.tr - self.bl)))
        super().__init__()

    def value(self, x):
        return 1. - np.max(np.abs(np.dot(x[..., :2] - self.center, self.scaler)), -1)

    def d_value(self, x):
        # outward unit normal of the nearest side, chosen by the two diagonals of the box
        xx = np.dot(x[..., :2] - self.center, self.scaler)
        c1 = xx[..., 0] + xx[..., 1]
        c2 = xx[..., 0] - xx[..., 1]
        shape = xx.shape
        right = np.broadcast_to(np.array((1., 0.)), shape)
        top = np.broadcast_to(np.array((0., 1.)), shape)
        left = np.broadcast_to(np.array((-1., 0.)), shape)
        bottom = np.broadcast_to(np.array((0., -1.)), shape)
        pick = lambda cond, a, b: np.where(np.expand_dims(cond, -1) if xx.ndim > 1 else cond, a, b)
        return pick((c1 > 0) & (c2 > 0), right,
                    pick((c1 > 0) & (c2 < 0), top,
                         pick((c1 < 0) & (c2 < 0), left, bottom)))

    def lower(self):
        return OBS_FRAME, 1., (0., 0.), (float(self.center[0]), float(self.center[1]),
                                          float(self.scaler[0, 0]), float(self.scaler[1, 1]))


class GreatCircleObs(Obstacle):
    """Half-space beyond the great circle through p1 and p2, optionally limited to a lon/lat zone [z1, z2]
    (obstacle.py:131-172).  value = X(x) . dir with X the unit vector of (lon, lat) and dir = -(X1 x X2) / |X1 x X2|;
    outside the zone the value is 1 and the gradient (1, 1), as written."""

    def __init__(self, p1, p2, z1=None, z2=None, autobox=False):
        def unit(p):
            return np.array((np.cos(p[0]) * np.cos(p[1]), np.sin(p[0]) * np.cos(p[1]), np.sin(p[1])))

        if autobox:
            # as in the reference, the latitude extent is p1[1] - p2[0] (lat of p1 minus LON of p2, obstacle.py:151)
            delta_lon = Utils.angular_diff(p1[0], p2[0])
            delta_lat = p1[1] - p2[0]
            self.z1 = np.array((min(p1[0] - delta_lon / 2., p2[0] - delta_lon / 2.),
                                min(p1[1] - delta_lat / 2., p2[1] - delta_lat / 2.)))
            self.z2 = np.array((max(p1[0] + delta_lon / 2., p2[0] + delta_lon / 2.),
                                max(p1[1] + delta_lat / 2., p2[1] + delta_lat / 2.)))
        else:
            self.z1, self.z2 = z1, z2
        self.dir_vect = -np.cross(unit(p1), unit(p2))
        self.dir_vect /= np.linalg.norm(self.dir_vect)
        super().__init__()

    def _outside_zone(self, x):
        return self.z1 is not None and not Utils.in_lonlat_box(self.z1, self.z2, x)

    def value(self, x):
        if self._outside_zone(x):
            return 1.
        return np.array((np.cos(x[0]) * np.cos(x[1]), np.sin(x[0]) * np.cos(x[1]), np.sin(x[1]))) @ self.dir_vect

    def d_value(self, x):
        if self._outside_zone(x):
            return np.array((1., 1.))
        d_lon = np.array((-np.sin(x[0]) * np.cos(x[1]), np.cos(x[0]) * np.cos(x[1]), 0))
        d_lat = np.array((-np.cos(x[0]) * np.sin(x[1]), -np.sin(x[0]) * np.sin(x[1]), np.cos(x[1])))
        return np.array((self.dir_vect @ d_lon, self.dir_vect @ d_lat))

    def lower(self):
        zone = (0., 0., 0., 0., 0.) if self.z1 is None else (float(self.z1[0]), float(self.z1[1]), float(self.z2[0]),
                                                             float(self.z2[1]), 1.)
        return OBS_GREAT_CIRCLE, 1., (0., 0.), tuple(float(v) for v in self.dir_vect) + zone
