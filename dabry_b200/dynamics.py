"""Vehicle dynamics x' = f(t, x, u) of the drop-in API (reference: dabry/dynamics.py).

Host-side numpy evaluation for problem set-up and the public API; the solver evaluates the same expressions on the
device (csrc/model.cuh `rhs_aug`, `dyn_value`).
"""
from abc import ABC, abstractmethod

import numpy as np
from numpy import cos, sin

from dabry_b200.flowfield import FlowField


class Dynamics(ABC):
    def __init__(self, ff: FlowField):
        self.ff = ff

    @abstractmethod
    def value(self, t, x, u):
        """dx/dt"""

    @abstractmethod
    def d_value__d_state(self, t, x, u):
        """Partial derivative of f with respect to the state."""

    def __call__(self, t, x, u):
        return self.value(t, x, u)


class ZermeloR2Dyn(Dynamics):
    """Planar Zermelo: x' = u + w(t, x) (dynamics.py:70-82)."""

    def value(self, t, x, u):
        return u + self.ff.value(t, x)

    def d_value__d_state(self, t, x, u):
        return self.ff.d_value(t, x)


class ZermeloS2Dyn(Dynamics):
    """Zermelo on the plate-carree chart of the sphere, x = (lon, lat): x' = diag(1/cos(lat), 1) (u + w)
    (dynamics.py:85-99)."""

    def value(self, t, x, u):
        return np.diag([1 / cos(x[1]), 1.]) @ (u + self.ff.value(t, x))

    def d_value__d_state(self, t, x, u):
        # As the reference evaluates it (dynamics.py:97-99): the second column is
        # diag(sin/cos^2, 0) @ u + w  - operator precedence binds the product before the sum.
        metric = np.diag((1 / cos(x[1]), 1)) @ self.ff.d_value(t, x)
        column = np.diag((sin(x[1]) / (cos(x[1]) ** 2), 0.)) @ u + self.ff.value(t, x)
        return metric + np.column_stack((np.zeros(2), column))
