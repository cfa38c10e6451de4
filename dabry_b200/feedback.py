"""Feedback control laws (reference: dabry/feedback.py).  They are used only by
`NavigationProblem.apply_feedback` to bound the time horizon (two scalar ODE solves per problem, host side;
SURVEY 8f rank 1)."""
from abc import ABC, abstractmethod
import numpy as np
from numpy import cos, sin
from numpy import arctan2 as atan2  # numpy's, as in the reference (feedback.py:4): differs from math.atan2 in the last bit

from dabry_b200.misc import Coords, Utils, directional_timeopt_control


class Feedback(ABC):
    @abstractmethod
    def __call__(self, t, x):
        """control vector at (t, x)"""


class ConstantFB(Feedback):
    def __init__(self, value):
        self.value = value.copy()

    def __call__(self, t, x):
        return self.value


def _local_direction_to(target, x, coords):
    """Direction from x to target: planar difference, or the chord to the target expressed in the local
    (east, north) frame of the sphere (feedback.py:69-85)."""
    if coords != Coords.GCS:
        return np.array(target - x, dtype=float)
    lon, lat = x[0], x[1]
    here = Utils.EARTH_RADIUS * np.array((cos(lon) * cos(lat), sin(lon) * cos(lat), sin(lat)))
    east = np.array((-sin(lon), cos(lon), 0.))
    north = np.array((-sin(lat) * cos(lon), -sin(lat) * sin(lon), cos(lat)))
    lon_t, lat_t = target[0], target[1]
    there = Utils.EARTH_RADIUS * np.array((cos(lon_t) * cos(lat_t), sin(lon_t) * cos(lat_t), sin(lat_t)))
    return np.array(((there - here) @ east, (there - here) @ north))


class GSTargetFB(Feedback):
    """Points the GROUND velocity at a fixed target (feedback.py:58-89)."""

    def __init__(self, ff, srf: float, target):
        self.ff = ff
        self.srf = srf
        self.target = target.copy()

    def __call__(self, t, x):
        e_target = _local_direction_to(self.target, x, self.ff.coords)
        if np.isclose(np.linalg.norm(e_target), 0):
            return np.zeros(2)
        return directional_timeopt_control(self.ff.value(t, x), e_target, self.srf)


class HTargetFB(Feedback):
    """Points the HEADING at a fixed target; returns a unit vector, not scaled by the vehicle speed
    (feedback.py:92-130)."""

    def __init__(self, target, coords: Coords):
        self.target = target.copy()
        self.coords = coords

    def __call__(self, t, x):
        e_target = _local_direction_to(self.target, x, self.coords)
        if np.isclose(np.linalg.norm(e_target), 0):
            return np.zeros(2)
        e_target = e_target / np.linalg.norm(e_target)
        heading = atan2(e_target[1], e_target[0])
        return np.array((np.cos(heading), np.sin(heading)))


class MapFB(Feedback):
    """Piecewise-constant heading map on a grid (feedback.py:133-151)."""

    def __init__(self, grid, values):
        self.grid = np.array(grid)
        self.values = np.array(values)

    def __call__(self, _, x):
        nx, ny, _ = self.grid.shape
        fx = (x[0] - self.grid[0, 0, 0]) / (self.grid[-1, 0, 0] - self.grid[0, 0, 0])
        fy = (x[1] - self.grid[0, 0, 1]) / (self.grid[0, -1, 1] - self.grid[0, 0, 1])
        return self.values[int((nx - 1) * fx), int((ny - 1) * fy)]
