"""Multi-GPU driver: the ring of root extremals is sharded by contiguous theta_0 ranges, one process per GPU, the
flow grid replicated.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) carries only

  * the ring boundary: after the roots are integrated, each rank sends the trajectory of its FIRST root to the
    previous rank, where it becomes the ghost neighbour of that rank's LAST root (solver_ef.py:468-469), so that
    compute_new_sites sees the same neighbour pairs as a single-GPU run;
  * the final arrival-time argmin: all_reduce(MIN) of the arrival cost, then of the owning rank.

Children are created between their parents and stay on the left parent's rank.
"""
import numpy as np

from dabry_b200 import _native as nat


def shard_range(n_roots: int, rank: int, world: int):
    """Contiguous theta_0 range [offset, offset + count) of `rank`; the first `n_roots % world` ranks get one more."""
    base, extra = divmod(n_roots, world)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


class ShardedResampling:
    """Runs `SolverEFResampling` over all ranks of the default process group.

    >>> runner = ShardedResampling(lambda **kw: SolverEFResampling(pb, T, max_depth=4, n_costate_sectors=N, **kw))
    >>> runner.solve(); runner.global_cost, runner.owner_rank
    """

    def __init__(self, make_solver, n_roots: int, device=None, group=None, block=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.root_range = shard_range(n_roots, self.rank, self.world)
        # block=None: one contiguous theta_0 range per rank.  block=B: block-cyclic, B roots per block - extremals
        # heading out of the domain are captured early, so contiguous ranges are unevenly loaded
        if block and self.world > 1:
            kwargs = dict(shard=(self.rank, self.world, block))
        else:
            kwargs = dict(root_range=self.root_range if self.world > 1 else None)
        if device is not None:
            kwargs['device'] = device
        self.solver = make_solver(**kwargs)
        self.global_cost = float('nan')
        self.owner_rank = -1
        self.total_rk_steps = 0

    def _tensor(self, array):
        import torch
        t = torch.from_numpy(array)
        if self.dist.get_backend(self.group) == 'nccl':
            t = t.cuda(self.solver.device)
        return t

    def _exchange_boundary(self, handle):
        """first root of every block of rank r -> ghost of the preceding block, which lives on rank r - 1 (ring)."""
        import ctypes as C
        size, count = C.c_int64(), C.c_int64()
        handle.call('dabry_boundary_doubles', C.byref(size))
        handle.call('dabry_boundary_count', C.byref(count))
        send = np.empty((count.value, size.value))
        handle.call('dabry_export_boundary', nat.dptr(send))
        t_send, t_recv = self._tensor(send), self._tensor(np.empty_like(send))
        dist = self.dist
        ops = [dist.P2POp(dist.isend, t_send, (self.rank - 1) % self.world, self.group),
               dist.P2POp(dist.irecv, t_recv, (self.rank + 1) % self.world, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        recv = t_recv.cpu().numpy()
        # block k of rank r is followed by block k of rank r + 1, except on the last rank: by block k + 1 of rank 0
        if self.rank == self.world - 1:
            recv = np.roll(recv, -1, axis=0)
        recv = np.ascontiguousarray(recv)
        handle.call('dabry_import_boundary', nat.dptr(recv))

    def solve(self, fetch=True):
        """fetch=False keeps every per-site array on the device (only the 8-byte optimum crosses to the host)."""
        s = self.solver
        handle = s._ensure_handle()
        handle.call('dabry_setup', s._VARIANT, nat.dptr(s._root_costates()))
        for depth in range(s.max_depth):
            handle.call('dabry_step_integrate')
            if self.world > 1 and depth == 0:
                self._exchange_boundary(handle)
            handle.call('dabry_step_resample')
        s._invalidate()
        s.depth = s.max_depth
        return self._finish(handle, fetch)

    def _finish(self, handle, fetch):
        s = self.solver
        if fetch:
            s.check_solutions()
        else:
            handle.call('dabry_finish')
        stats = handle.stats()
        # arrival-time argmin over ranks, and the step total for the throughput metric
        cost = handle.solution()[1]
        local = np.array([np.inf if np.isnan(cost) else cost])
        t = self._tensor(local.copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        best = float(t.cpu().numpy()[0])
        owner = self._tensor(np.array([float(self.rank) if local[0] == best and np.isfinite(best) else 1e18]))
        self.dist.all_reduce(owner, op=self.dist.ReduceOp.MIN, group=self.group)
        steps = self._tensor(np.array([float(stats['rk_steps'])]))
        self.dist.all_reduce(steps, op=self.dist.ReduceOp.SUM, group=self.group)
        self.global_cost = best if np.isfinite(best) else float('nan')
        own = float(owner.cpu().numpy()[0])
        self.owner_rank = int(own) if own < 1e17 else -1
        self.total_rk_steps = int(steps.cpu().numpy()[0])
        return self


class ShardedTrimming(ShardedResampling):
    """`SolverEFTrimming` over all ranks: per sub-frame the ring boundary is re-sent (the ghost root grows with its
    owner) and the 100 x 100 trimming cost map is all_reduce(MIN)-ed (80 kB) before the dominated extremals are
    closed, so that every rank trims against the global front (solver_ef.py:819-841)."""

    def solve(self, fetch=True):
        s = self.solver
        handle = s._ensure_handle()
        handle.call('dabry_setup', s._VARIANT, nat.dptr(s._root_costates()))
        for _ in range(s.n_subframes):
            handle.call('dabry_step_trim_integrate')
            if self.world > 1:
                self._exchange_boundary(handle)
            handle.call('dabry_step_trim_refine')
            if self.world > 1:
                local = handle.cost_map(100, 100, full=False)
                t = self._tensor(local)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
                merged = np.ascontiguousarray(t.cpu().numpy())
                handle.call('dabry_set_cost_map', nat.dptr(merged))
            handle.call('dabry_step_trim_close')
        s._invalidate()
        s.i_subframe = s.n_subframes
        return self._finish(handle, fetch)
