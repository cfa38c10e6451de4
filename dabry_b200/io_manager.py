"""Case-directory output of the drop-in API (reference: dabry/io_manager.py:137-201).

The writers the extremal-field path uses: `save_traj(s)`, `save_ff`, the batched `save_trajs_bundle` (one npz for a
whole front; SURVEY 8f rank 3 - a million per-site files is unusable), and exporters for the two legacy h5 layouts the
reference documents and its viewer still reads (`docs/traj_data_h5_format.txt`, `docs/rff_data_h5_format.txt`,
`display/display.py:823-855`; v1.1 itself has no writer for them): `save_trajs_h5`, `save_rff`.  They need the
optional dependency h5py; `save_rff` falls back to an npz with the same dataset names when it is absent.
GRIB / CDS ingestion is out of scope (needs pygrib and network access).
"""
import os
import shutil
from typing import List, Optional

import numpy as np

from dabry_b200.flowfield import FlowField, save_ff
from dabry_b200.trajectory import Trajectory


class IOManager:
    def __init__(self, name: str, case_dir: Optional[str] = None, cache_ff=False, cache_rff=False):
        self.case_name = name
        root = os.environ.get('DABRY_OUTPUT_DIR', os.path.join(os.getcwd(), 'output'))
        self.case_dir = case_dir if case_dir is not None else os.path.join(root, str(name).replace(' ', '_'))
        self.cache_ff = cache_ff
        self.cache_rff = cache_rff
        self.ff_filename = 'ff.npz'

    @property
    def output_dir(self):
        return os.path.dirname(self.case_dir)

    def set_case_dir(self, dirpath: str):
        self.case_dir = os.path.abspath(dirpath)

    @property
    def pb_data_fpath(self):
        return os.path.join(self.case_dir, self.case_name + '.json')

    @property
    def trajs_dir(self):
        return os.path.join(self.case_dir, 'trajs')

    @property
    def ff_fpath(self):
        return os.path.join(self.case_dir, self.ff_filename)

    def setup_dir(self):
        os.makedirs(self.case_dir, exist_ok=True)

    def setup_trajs(self):
        self.setup_dir()
        os.makedirs(self.trajs_dir, exist_ok=True)

    def clean_output_dir(self):
        if not os.path.exists(self.case_dir):
            return
        for filename in os.listdir(self.case_dir):
            if (filename.endswith('ff.npz') and self.cache_ff) or (filename.endswith('rff.h5') and self.cache_rff):
                continue
            path = os.path.join(self.case_dir, filename)
            shutil.rmtree(path) if os.path.isdir(path) else os.remove(path)

    def save_traj(self, traj: Trajectory, name: str, target_dir: Optional[str] = None):
        if target_dir is None:
            self.setup_trajs()
            target_dir = self.trajs_dir
        traj.save(name, target_dir)

    def save_trajs(self, trajs: List[Trajectory], group_name: Optional[str] = None):
        """`trajs/<group>/traj_<zero padded index>.npz` + meta json per trajectory (io_manager.py:142-154)."""
        if len(trajs) == 0:
            return
        self.setup_trajs()
        target_dir = self.trajs_dir if group_name is None else os.path.join(self.trajs_dir, group_name)
        os.makedirs(target_dir, exist_ok=True)
        width = 1 + (int(np.log10(len(trajs) - 1)) if len(trajs) > 1 else 0)
        for i_traj, traj in enumerate(trajs):
            self.save_traj(traj, str(i_traj).rjust(width, '0'), target_dir)

    def save_trajs_bundle(self, name: str, **arrays):
        """One compressed npz holding a whole front (padded arrays + per-site records)."""
        self.setup_trajs()
        np.savez_compressed(os.path.join(self.trajs_dir, name + '.npz'), **arrays)

    def save_ff(self, ff: FlowField, nx=None, ny=None, nt=None, bl=None, tr=None):
        if os.path.exists(self.ff_fpath) and self.cache_ff:
            return
        self.setup_dir()
        save_ff(ff, self.ff_fpath, nx=nx, ny=ny, nt=nt, bl=bl, tr=tr)

    # ---- legacy h5 layouts -------------------------------------------------------------------------------------
    @staticmethod
    def _h5py():
        try:
            import h5py
        except ImportError as exc:
            raise ImportError('the h5 exporters need the optional dependency h5py') from exc
        return h5py

    def save_trajs_h5(self, trajs: List[Trajectory], filename: str = 'trajectories.h5', names=None,
                      traj_type: str = 'pmp'):
        """`docs/traj_data_h5_format.txt`: one group per trajectory with attributes `coords, type, last_index, label,
        info, interrupted` and datasets `data (nt, 2)`, `ts (nt,)`, `controls (nt,)` = heading angle of the control
        vector, `adjoints (nt, 2)`."""
        h5py = self._h5py()
        self.setup_dir()
        path = os.path.join(self.case_dir, filename)
        with h5py.File(path, 'w') as f:
            for i, traj in enumerate(trajs):
                g = f.create_group(str(i))
                g.attrs['coords'] = traj.coords.value
                g.attrs['type'] = traj_type
                g.attrs['last_index'] = len(traj) - 1
                g.attrs['label'] = i
                g.attrs['info'] = '' if names is None else str(names[i])
                g.attrs['interrupted'] = False
                g.create_dataset('data', data=np.asarray(traj.states, dtype=float))
                g.create_dataset('ts', data=np.asarray(traj.times, dtype=float))
                if traj.controls is not None:
                    u = np.asarray(traj.controls, dtype=float)
                    g.create_dataset('controls', data=np.arctan2(u[:, 1], u[:, 0]))
                if traj.costates is not None:
                    g.create_dataset('adjoints', data=np.asarray(traj.costates, dtype=float))
        return path

    def save_rff(self, data: np.ndarray, ts: np.ndarray, grid: np.ndarray, coords, fmt: Optional[str] = None):
        """`docs/rff_data_h5_format.txt`: attribute `coords`, datasets `data (nt, nx, ny)`, `ts (nt,)`, `grid (nx, ny,
        2)` - a reachability-front function whose zero level set at ts[k] is the front at that time.  fmt: 'h5', 'npz',
        or None = h5 when h5py can be imported, else npz with the same names."""
        data, ts, grid = np.asarray(data, dtype=float), np.asarray(ts, dtype=float), np.asarray(grid, dtype=float)
        if data.shape != (ts.shape[0],) + grid.shape[:2] or grid.shape[2:] != (2,):
            raise ValueError('rff: data (nt, nx, ny), ts (nt,), grid (nx, ny, 2) expected')
        coords_name = getattr(coords, 'value', coords)
        if fmt is None:
            try:
                import h5py  # noqa: F401
                fmt = 'h5'
            except ImportError:
                fmt = 'npz'
        self.setup_dir()
        if fmt == 'npz':
            path = os.path.join(self.case_dir, 'rff.npz')
            np.savez(path, data=data, ts=ts, grid=grid, coords=coords_name)
            return path
        if fmt != 'h5':
            raise ValueError('fmt must be "h5", "npz" or None')
        h5py = self._h5py()
        path = os.path.join(self.case_dir, 'rff.h5')
        with h5py.File(path, 'w') as f:
            f.attrs['coords'] = coords_name
            f.create_dataset('data', data=data)
            f.create_dataset('ts', data=ts)
            f.create_dataset('grid', data=grid)
        return path
