"""`Trajectory` container and its on-disk format (reference: dabry/trajectory.py).

File format (trajectory.py:165-201): `traj_<name>.npz` with arrays `times, states, costates, controls, cost`
(missing optional arrays stored empty) plus `traj_<name>_meta.json` = {"coords": ..., "events": {name: [t, ...]}}.
"""
import json
import os
import warnings
from typing import Dict, Optional

import numpy as np
from numpy import ndarray

from dabry_b200.misc import Coords


def traj_name_to_filename(name, meta=False):
    return 'traj_%s%s' % (name, '_meta.json' if meta else '.npz')


def traj_filepath_to_name(filepath):
    filename = os.path.basename(filepath)
    if not filename.startswith('traj_'):
        raise ValueError('Filename "%s" is not properly formatted' % filename)
    return filename[5:].split('.')[0]


def _nan_rows(n):
    return np.full((n, 2), np.nan)


class Trajectory:
    def __init__(self, times: ndarray, states: ndarray, coords: Coords, controls: Optional[ndarray] = None,
                 costates: Optional[ndarray] = None, cost: Optional[ndarray] = None,
                 events: Optional[Dict[str, ndarray]] = None):
        self.times = times.copy()
        self.states = states.copy()
        self.controls = None if controls is None else controls.copy()
        self.costates = None if costates is None else costates.copy()
        self.cost = None if cost is None else cost.copy()
        self.events = events if events is not None else {}
        self.coords = coords

    @classmethod
    def empty(cls):
        none2 = np.array(((), ()))
        return cls(np.array(()), none2, Coords.CARTESIAN, none2, none2, np.array(()))

    @classmethod
    def cartesian(cls, times, states, controls=None, costates=None, cost=None, events=None):
        return cls(times, states, Coords.CARTESIAN, controls=controls, costates=costates, cost=cost, events=events)

    @classmethod
    def gcs(cls, times, states, controls=None, costates=None, cost=None, events=None):
        return cls(times, states, Coords.GCS, controls=controls, costates=costates, cost=cost, events=events)

    def __len__(self):
        return self.times.shape[0]

    def __add__(self, other):
        """Concatenation in time; an optional array missing on one side is filled with NaN
        (trajectory.py:104-140)."""
        if not isinstance(other, Trajectory):
            raise ValueError('Trajectory addition with type %s' % type(other))
        if self.coords != other.coords:
            raise ValueError('Incompatible coord types %s and %s' % (self.coords, other.coords))
        if len(self) == 0 or len(other) == 0:
            src = other if len(self) == 0 else self
            times, states, controls, costates, cost = src.times, src.states, src.controls, src.costates, src.cost
        else:
            na, nb = len(self), len(other)
            times = np.concatenate((self.times, other.times))
            states = np.concatenate((self.states, other.states))
            controls = np.concatenate((self.controls if self.controls is not None else _nan_rows(na),
                                       other.controls if other.controls is not None else _nan_rows(nb)))
            costates = np.concatenate((self.costates if self.costates is not None else _nan_rows(na),
                                       other.costates if other.costates is not None else _nan_rows(nb)))
            cost = np.concatenate((self.cost if self.cost is not None else np.full(na, np.nan),
                                   other.cost if other.cost is not None else np.full(nb, np.nan)))
        events = dict(self.events)
        events.update(other.events)
        return Trajectory(times, states, self.coords, controls=controls, costates=costates, cost=cost, events=events)

    def copy(self):
        return Trajectory(self.times, self.states, self.coords, controls=self.controls, costates=self.costates,
                          cost=self.cost, events=dict(self.events))

    def save(self, name, dir_name):
        none2 = np.array(((), ()))
        np.savez(os.path.join(dir_name, traj_name_to_filename(name)), times=self.times, states=self.states,
                 costates=self.costates if self.costates is not None else none2,
                 controls=self.controls if self.controls is not None else none2,
                 cost=self.cost if self.cost is not None else np.array(()))
        meta = {'coords': self.coords.value, 'events': {k: np.asarray(v).tolist() for k, v in self.events.items()}}
        with open(os.path.join(dir_name, traj_name_to_filename(name, meta=True)), 'w') as f:
            json.dump(meta, f)

    @classmethod
    def from_npz(cls, filepath):
        if not filepath.endswith('.npz'):
            raise ValueError('Not a NPZ file %s' % filepath)
        data = np.load(filepath)
        try:
            with open(filepath[:-4] + '_meta.json') as f:
                meta = json.load(f)
            coords = Coords.from_string(meta['coords'])
            events = {k: np.array(v) for k, v in meta['events'].items()}
        except FileNotFoundError:
            warnings.warn('Metadata not found for trajectory', category=UserWarning)
            coords, events = Coords.CARTESIAN, {}
        opt = lambda key: data[key] if data[key].size > 0 else None
        return cls(data['times'], data['states'], coords, opt('controls'), opt('costates'), opt('cost'), events)
