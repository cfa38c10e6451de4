"""BASELINE config 3 asks for a front of 0.9 - 1.1 M extremals from 2^19 roots at max_depth = 4: realised site count as a
function of max_dist (run on a GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dabry_b200.solver_ef import SolverEFResampling  # noqa: E402

pb, total, kw, descr = bench.build_problem('planar')
kw.pop('max_dist')
for max_dist in [float(a) for a in sys.argv[1:]] or (0.002, 0.0014, 0.001, 0.0007, 0.0005):
    s = SolverEFResampling(pb, total, n_costate_sectors=1 << 19, max_dist=max_dist, max_sites=3 << 20, device_roots=True,
                           **kw)
    s.solve()
    st = s.native_stats()
    print('max_dist %-8g sites %8d  rk_steps %.4g  ms_total %.2f  cost %r' % (
        max_dist, st['n_sites'], st['rk_steps'], st['ms_total'], s.solution_cost), flush=True)
    s.close()
