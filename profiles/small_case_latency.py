"""Time-to-solution of the reference's own small cases (30 roots, default depth) on the GPU: these are latency-bound
(every launch has a few hundred threads).  Run on a GPU box:  python profiles/small_case_latency.py [case ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dabry_b200.problem import NavigationProblem  # noqa: E402
from dabry_b200.solver_ef import SolverEFResampling, SolverEFTrimming  # noqa: E402

for name in (sys.argv[1:] or ('movor', 'three_vortices', 'gyre', 'obstacle')):
    for cls in (SolverEFResampling, SolverEFTrimming):
        pb = NavigationProblem.from_name(name).rescale()
        best = None
        for _ in range(3):
            s = cls(pb, NavigationProblem.time_bound(name))
            t0 = time.perf_counter()
            s.solve()
            dt = time.perf_counter() - t0
            st = s.native_stats()
            n = len(s.sites)
            s.close()
            best = dt if best is None else min(best, dt)
        print('%-16s %-20s %5d sites  %8d rk steps  solve %7.1f ms (device %6.1f ms, %d launches)' % (
            name, cls.__name__, n, st['rk_steps'], 1e3 * best, st['ms_total'], st['kernel_launches']), flush=True)
