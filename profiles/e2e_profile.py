"""Where the time of one end-to-end solve goes (public API, pinned host arrays in, results out): wall time per native
call and per host function.  Run on a GPU box:
    python profiles/e2e_profile.py [roots] [--no-pin] [--torch]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if '--torch' in sys.argv:
    import torch
    torch.cuda.set_device(0)
    torch.zeros(1, device='cuda')
import bench  # noqa: E402
from dabry_b200 import _native as nat  # noqa: E402
from dabry_b200.solver_ef import SolverEFResampling  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
roots = int(args[0]) if args else 1_250_000
pb, total, kw, descr = bench.build_problem('era5')
if '--no-pin' not in sys.argv:
    pb.model.ff.ff.pin_memory()

spent = collections.OrderedDict()


def timed(label, fn):
    def wrapper(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            spent[label] = spent.get(label, 0.) + time.perf_counter() - t0
    return wrapper


_call = nat.Handle.call
nat.Handle.call = lambda self, name, *a: timed(name, _call)(self, name, *a)
nat.Handle.__init__ = timed('dabry_create', nat.Handle.__init__)
nat.Handle.close = timed('dabry_destroy', nat.Handle.close)


def step():
    s = SolverEFResampling(pb, total, n_costate_sectors=roots, max_sites=int(roots * 1.02) + 4096, **kw)
    s.solve()
    cost = s.solution_cost
    traj = s.solution_site.traj if s.solution_site is not None else None
    st = s.native_stats()
    s.close()
    return cost, st


step()
for k in range(3):
    spent.clear()
    t0 = time.perf_counter()
    cost, st = step()
    wall = time.perf_counter() - t0
    print('e2e step %d: %.1f ms wall, cost %r' % (k, 1e3 * wall, cost))
    print('   native calls (ms): ' + ', '.join('%s %.1f' % (n, 1e3 * v) for n, v in spent.items()))
    print('   host python outside native calls: %.1f ms' % (1e3 * (wall - sum(spent.values()))))
    print('   device phases (ms): ' + ', '.join('%s %.1f' % (n[3:], v) for n, v in st.items() if n.startswith('ms_')))
