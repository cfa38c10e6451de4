"""Multi-GPU driver: the ring of root extremals is sharded by theta_0 (contiguous ranges or block-cyclic), one process
per GPU, the flow grid replicated.  The exchanges themselves live INSIDE the library
(`dabry_solve_sharded`, csrc/engine.cuh): NCCL calls enqueued on the solver's stream -

  * the ring boundary: after the roots are integrated (Trimming: every sub-frame) each rank sends the FIRST root of
    every local block to the previous rank, where it becomes the ghost neighbour of that rank's LAST root of the block
    before (solver_ef.py:468-469) - ncclSend / ncclRecv in one group;
  * Trimming: ncclAllReduce(min) of the 100 x 100 cost map, in place, every sub-frame (solver_ef.py:825-828);
  * the arrival-time argmin (solver_ef.py:667-685): one ncclAllGather of (cost, step count) per rank.

This module only bootstraps the library's communicator: rank 0 draws the unique id, `torch.distributed` (any backend)
broadcasts its 128 bytes, every rank calls `dabry_comm_create`.  In the CPU test tier the same code runs over
tests/hostsim, whose communicator is a shared-memory test double, with gloo carrying the id.

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


_COMMS = {}


def library_comm(device=0, group=None):
    """The library's communicator over the ranks of `group` (default process group), created once per (group, device)
    and kept for the life of the process: a solver is created per solve(), the communicator is not."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return None
    key = (id(group), device, rank, world, nat._lib_path)
    comm = _COMMS.get(key)
    if comm is None:
        box = [nat.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        comm = _COMMS[key] = nat.Comm(device, rank, world, box[0])
    return comm


def close_comms():
    for comm in _COMMS.values():
        comm.close()
    _COMMS.clear()


class ShardedResampling:
    """Runs `SolverEFResampling` over all ranks of the default process group.

    >>> runner = ShardedResampling(lambda **kw: SolverEFResampling(pb, T, max_depth=4, n_costate_sectors=N, **kw), N)
    >>> runner.solve(); runner.global_cost, runner.owner_rank
    """

    def __init__(self, make_solver, n_roots: int, device=None, group=None, block=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        # no process group: one rank (dabry_solve_sharded is dabry_solve then)
        live = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if live else 0
        self.world = dist.get_world_size(group) if live else 1
        self.root_range = shard_range(n_roots, self.rank, self.world)
        # block=None: one contiguous theta_0 range per rank.  block=B: block-cyclic, B roots per block - extremals
        # heading out of the domain are captured early, so contiguous ranges are unevenly loaded
        if block and self.world > 1:
            kwargs = dict(shard=(self.rank, self.world, block))
        else:
            kwargs = dict(root_range=self.root_range if self.world > 1 else None)
        if device is not None:
            kwargs['device'] = device
        self.comm = library_comm(device or 0, group)
        kwargs['comm'] = self.comm
        self.solver = make_solver(**kwargs)
        if self.comm is not None and getattr(self.solver, '_comm', None) is not self.comm:
            self.solver._comm = self.comm  # a make_solver that swallowed the keyword
        self.global_cost = float('nan')
        self.owner_rank = -1
        self.total_rk_steps = 0

    def solve(self, fetch=True):
        """fetch=False keeps every per-site array on the device (only the optimum crosses to the host)."""
        s = self.solver
        handle = s._ensure_handle()
        handle.call('dabry_solve_sharded', s._VARIANT, nat.dptr(s._root_costates()))
        s._invalidate()
        s._after_solve()
        if fetch:
            s.check_solutions(_finished=True)
        cost, owner, steps = handle.global_solution()
        self.global_cost = cost
        self.owner_rank = owner
        self.total_rk_steps = int(steps)
        return self


class ShardedTrimming(ShardedResampling):
    """`SolverEFTrimming` over all ranks: per sub-frame the ring boundary is re-sent (the ghost root grows with its
    owner) and the 100 x 100 trimming cost map is all-reduced (min) before the dominated extremals are closed, so that
    every rank trims against the global front (solver_ef.py:819-841).  The loop itself is the library's."""
