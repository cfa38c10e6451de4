"""How often does one end-to-end solve hit a slow dabry_create / dabry_destroy?  N solves of the bench workload through the
public API with the wall time of every native call; prints the steps in which a call other than the solve took > 2 ms.
    DABRY_B200_DEBUG_TIMING=1 python profiles/e2e_outliers.py [n_steps]      (on a GPU box)"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dabry_b200 import _native as nat  # noqa: E402
from dabry_b200.solver_ef import SolverEFResampling  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
pb, total, kw, descr = bench.build_problem('era5')
pb.model.ff.ff.pin_memory()
roots = 1_250_000
walls = []
for k in range(n):
    nat.CALL_TIMES = {}
    t0 = time.perf_counter()
    s = SolverEFResampling(pb, total, n_costate_sectors=roots, max_sites=int(roots * 1.02) + 4096, **kw)
    s.solve()
    cost = s.solution_cost
    s.close()
    wall = 1e3 * (time.perf_counter() - t0)
    walls.append(wall)
    slow = {c: round(1e3 * v, 1) for c, v in nat.CALL_TIMES.items() if v > 2e-3 and 'solve' not in c}
    if slow or k < 3:
        print('step %2d: %.1f ms wall, slow calls %s' % (k, wall, slow), flush=True)
walls = sorted(walls[2:])
print('%d steps after 2 warm-ups: median %.1f ms, max %.1f ms, mean %.1f ms' % (len(walls), walls[len(walls) // 2], walls[-1], sum(walls) / len(walls)))
