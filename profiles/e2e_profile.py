"""Where the host time of one end-to-end solve goes (public API, host arrays in, results out).  Run on a GPU box:
    python profiles/e2e_profile.py [roots]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dabry_b200.solver_ef import SolverEFResampling  # noqa: E402

roots = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
pb, total, kw, descr = bench.build_problem('era5')


def step():
    s = SolverEFResampling(pb, total, n_costate_sectors=roots, max_sites=int(roots * 1.02) + 4096, **kw)
    s.solve()
    cost = s.solution_cost
    traj = s.solution_site.traj if s.solution_site is not None else None
    st = s.native_stats()
    s.close()
    return cost, st


step()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
cost, st = step()
pr.disable()
print('e2e step %.3f s, cost %r' % (time.perf_counter() - t0, cost))
print({k: v for k, v in st.items() if k.startswith('ms_')})
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
