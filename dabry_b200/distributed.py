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

    def _on_device(self):
        return self.dist.get_backend(self.group) == 'nccl'

    def _buffer(self, name, shape):
        """Persistent exchange buffer: a CUDA tensor on the solver's device under NCCL (the library reads and writes it
        in place through the *_device entry points), a CPU tensor under gloo (tests/hostsim: 'device' memory is host
        memory there, so the same entry points apply)."""
        import torch
        bufs = self.__dict__.setdefault('_bufs', {})
        t = bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            device = torch.device('cuda', self.solver.device) if self._on_device() else torch.device('cpu')
            t = bufs[name] = torch.empty(shape, dtype=torch.float64, device=device)
        return t

    def _sync_collective(self):
        """NCCL work is ordered on torch's stream, the library's on its own: drain torch's before handing a buffer over."""
        if self._on_device():
            import torch
            torch.cuda.current_stream(self.solver.device).synchronize()

    def _exchange_boundary(self, handle):
        """first root of every block of rank r -> ghost of the preceding block, which lives on rank r - 1 (ring).
        Device to device: the packets are written into the NCCL send buffer by the export kernel and read from the
        receive buffer by the import kernel (dabry_export/import_boundary_device); nothing is staged on the host."""
        import ctypes as C
        import torch
        size, count = C.c_int64(), C.c_int64()
        handle.call('dabry_boundary_doubles', C.byref(size))
        handle.call('dabry_boundary_count', C.byref(count))
        t_send = self._buffer('send', (count.value, size.value))
        t_recv = self._buffer('recv', (count.value, size.value))
        handle.call('dabry_export_boundary_device', C.c_void_p(t_send.data_ptr()))
        dist = self.dist
        ops = [dist.P2POp(dist.isend, t_send, (self.rank - 1) % self.world, self.group),
               dist.P2POp(dist.irecv, t_recv, (self.rank + 1) % self.world, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        # block k of rank r is followed by block k of rank r + 1, except on the last rank: by block k + 1 of rank 0
        if self.rank == self.world - 1:
            t_recv = torch.roll(t_recv, -1, dims=0).contiguous()
        self._sync_collective()
        handle.call('dabry_import_boundary_device', C.c_void_p(t_recv.data_ptr()))

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
                import ctypes as C
                t = self._buffer('cost_map', (100, 100))
                handle.call('dabry_get_cost_map_device', 0, C.c_void_p(t.data_ptr()))
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
                self._sync_collective()
                handle.call('dabry_set_cost_map_device', C.c_void_p(t.data_ptr()))
            handle.call('dabry_step_trim_close')
        s._invalidate()
        s.i_subframe = s.n_subframes
        return self._finish(handle, fetch)
