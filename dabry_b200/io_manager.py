"""Case-directory output of the drop-in API (reference: dabry/io_manager.py:137-201).

Only the writers the extremal-field path uses are provided: `save_traj(s)`, `save_ff`, plus the batched
`save_trajs_bundle` (one npz for a whole front; SURVEY 8f rank 3 - a million per-site files is unusable).
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
