"""Why the hot kernel runs with ~23 of 32 lanes: per-extremal RK attempt counts of the bench workload on the host.

Runs a 1 % block-cyclic sample (blocks of 128 consecutive roots = one CUDA block) of BASELINE config 4 at the bench's
theta_0 density through tests/hostsim (the device headers compiled by g++; DABRY_HOSTSIM_TRACE writes one
`site attempts` line per free arc) and prints what a lockstep warp pays: rounds = max over its 32 lanes.

    python profiles/lane_imbalance.py [n_roots] [world]
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402


def main():
    n_roots = int(sys.argv[1]) if len(sys.argv) > 1 else 1_280_000
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    trace = tempfile.mktemp(suffix='.trace')
    os.environ['DABRY_HOSTSIM_TRACE'] = trace
    import parity_util
    from dabry_b200 import _native
    _native.use_library(parity_util.build_hostsim())
    import bench
    from dabry_b200.solver_ef import SolverEFResampling
    pb, total, kw, desc = bench.build_problem('era5')
    solver = SolverEFResampling(pb, total_duration=total, n_costate_sectors=n_roots, shard=(0, world, 128),
                                max_depth=1)
    solver.solve()
    stats = solver.native_stats()
    a = np.loadtxt(trace, dtype=np.int64)
    os.unlink(trace)
    order = np.argsort(a[:, 0], kind='stable')
    att = a[order, 1]
    att = att[:len(att) // 32 * 32].reshape(-1, 32)
    rounds = att.max(axis=1)
    print('%s: %d extremals sampled of %d, rk_steps %d' % (desc, att.size, n_roots, stats['rk_steps']))
    print('attempts per extremal: mean %.1f  median %.0f  p99 %.0f  max %d' % (
        att.mean(), np.median(att), np.percentile(att, 99), att.max()))
    print('lockstep rounds per warp (max over lanes): mean %.1f  -> lane utilisation %.1f / 32' % (
        rounds.mean(), 32 * att.sum() / (32. * rounds.sum())))
    spread = rounds - np.median(att, axis=1)
    print('warp max - warp median: mean %.1f  p50 %.0f  p90 %.0f  p99 %.0f' % (
        spread.mean(), np.median(spread), np.percentile(spread, 90), np.percentile(spread, 99)))
    # how many lanes are still running in the extra rounds
    tail = [(w > np.median(w) + 2).sum() for w in att]
    print('lanes more than 2 attempts above their warp median: mean %.2f per warp' % np.mean(tail))
    np.save(os.path.join(ROOT, 'gpurun_out', 'lane_attempts.npy'), att)


if __name__ == '__main__':
    main()
