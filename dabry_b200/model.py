"""`Model`: a dynamics bound to a flow field (reference: dabry/model.py:31-68)."""
from typing import Optional

import numpy as np

from dabry_b200.dynamics import Dynamics, ZermeloR2Dyn, ZermeloS2Dyn
from dabry_b200.flowfield import FlowField, UniformFF
from dabry_b200.misc import Coords


class Model:
    def __init__(self, dyn: Dynamics):
        self.dyn = dyn
        self.coords = Model.which_coords(dyn.ff)

    @classmethod
    def which_coords(cls, ff: FlowField):
        return Coords.GCS if getattr(ff, 'coords', None) == Coords.GCS else Coords.CARTESIAN

    @property
    def ff(self):
        return self.dyn.ff

    @classmethod
    def zermelo_R2(cls, ff: Optional[FlowField]):
        return cls(ZermeloR2Dyn(UniformFF(np.zeros(2)) if ff is None else ff))

    @classmethod
    def zermelo_S2(cls, ff: FlowField):
        return cls(ZermeloS2Dyn(ff))

    @classmethod
    def zermelo(cls, ff: FlowField):
        """Planar or spherical Zermelo dynamics according to `ff.coords` (model.py:62-68)."""
        return cls.zermelo_S2(ff) if Model.which_coords(ff) == Coords.GCS else cls.zermelo_R2(ff)
