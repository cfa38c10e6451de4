"""The multi-GPU path on CPU: two `gloo` ranks, each driving its own handle of tests/hostsim (the device headers
compiled by g++), shard the ring of root extremals by theta_0 and exchange the ring boundary + the arrival-time
argmin through torch.distributed exactly like the NCCL run does (dabry_b200/distributed.py).  The union of the two
shards must reproduce the single-handle result site by site."""
import os
import sys

import numpy as np
import pytest

import parity_util as pu

WORKER = r'''
import os, sys, pickle
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests')); sys.path.insert(0, os.path.join({root!r}, 'oracle'))
import numpy as np
import torch.distributed as dist
import parity_util as pu
from dabry_b200 import _native
backend = {backend!r}
if backend == 'nccl':
    import torch
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    _native.use_library(None)          # the product library, one GPU per rank
    extra = dict(device=local)
else:
    _native.use_library(pu.HOSTSIM_LIB)
    extra = dict()
from dabry_b200.problem import NavigationProblem
from dabry_b200.solver_ef import SolverEFResampling, SolverEFTrimming
from dabry_b200.distributed import ShardedResampling, ShardedTrimming
dist.init_process_group(backend)
pb = NavigationProblem.from_name({case!r}).rescale()
Runner, Solver = (ShardedTrimming, SolverEFTrimming) if {trimming} else (ShardedResampling, SolverEFResampling)
runner = Runner(lambda **kw: Solver(pb, max_depth={depth}, n_costate_sectors=30, **kw), 30, block={block}, **extra)
runner.solve()
s = runner.solver
rec = {{'rank': runner.rank, 'names': list(s.sites.keys()), 'cost': runner.global_cost, 'owner': runner.owner_rank,
       'steps': runner.total_rk_steps, 'local_cost': s.solution_cost,
       'last': {{n: np.array(site.traj.states[-1]) for n, site in s.sites.items()}},
       'len': {{n: len(site.traj) for n, site in s.sites.items()}},
       'closure': {{n: (-1 if site.closure_reason is None else site.closure_reason.value) for n, site in s.sites.items()}}}}
pickle.dump(rec, open(os.path.join({out!r}, 'rank%d.pkl' % runner.rank), 'wb'))
dist.destroy_process_group()
'''


CASES = [('linear', 5, False, None), ('moving_vortex', 4, False, 5), ('gyre', 6, True, None), ('obstacle', 6, True, 3),
         ('three_vortices', 5, False, 1)]


@pytest.mark.parametrize('case,depth,trimming,block', CASES)
def test_two_rank_sharding_matches_single(case, depth, trimming, block, tmp_path, hostsim):
    _two_ranks_match_single(case, depth, trimming, block, tmp_path, 'gloo')


@pytest.mark.gpu
@pytest.mark.parametrize('case,depth,trimming,block', CASES)
def test_two_gpu_nccl_sharding_matches_single(case, depth, trimming, block, tmp_path, cuda_lib):
    """The same check on two B200s over NCCL: the ring boundary and the cost map travel device to device
    (dabry_*_device entry points).  Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    _two_ranks_match_single(case, depth, trimming, block, tmp_path, 'nccl')


def _two_ranks_match_single(case, depth, trimming, block, tmp_path, backend):
    import pickle
    import subprocess
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=pu.ROOT, case=case, depth=depth, out=str(tmp_path), trimming=trimming,
                                    block=block, backend=backend))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    subprocess.run([sys.executable, '-W', 'ignore', '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                    '--master-addr', '127.0.0.1', '--master-port', '29611', str(script)], check=True, env=env,
                   timeout=600)
    recs = [pickle.load(open(tmp_path / ('rank%d.pkl' % r), 'rb')) for r in range(2)]

    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling, SolverEFTrimming
    pb = NavigationProblem.from_name(case).rescale()
    single = (SolverEFTrimming if trimming else SolverEFResampling)(pb, max_depth=depth, n_costate_sectors=30)
    single.solve()
    # same extremals: the ghost copies are not reported, every site lives on exactly one rank
    names = recs[0]['names'] + recs[1]['names']
    assert sorted(names) == sorted(single.sites.keys())
    assert len(set(names)) == len(names)
    for rec in recs:
        for n in rec['names']:
            site = single.sites[n]
            assert rec['len'][n] == len(site.traj), n
            assert rec['closure'][n] == (-1 if site.closure_reason is None else site.closure_reason.value), n
            assert np.allclose(rec['last'][n], site.traj.states[-1], rtol=1e-12, atol=1e-12), n
    # arrival-time argmin over ranks == single-handle optimum; both ranks agree on it
    assert recs[0]['cost'] == recs[1]['cost'] == pytest.approx(single.solution_cost, rel=1e-12)
    assert recs[0]['owner'] == recs[1]['owner']
    owner = recs[0]['owner']
    assert recs[owner]['local_cost'] == recs[0]['cost']
    assert recs[0]['steps'] == recs[1]['steps'] == single.native_stats()['rk_steps']


@pytest.mark.gpu
def test_two_devices_in_one_process(cuda_lib):
    """One process, one solver per GPU, calls interleaved: every entry point makes its handle's device current, and the
    device block cache keeps the blocks of the two devices apart.  Needs two GPUs; skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from dabry_b200.problem import NavigationProblem
    from dabry_b200.solver_ef import SolverEFResampling
    pb = NavigationProblem.from_name('moving_vortex').rescale()
    bound = NavigationProblem.time_bound('moving_vortex')
    for _ in range(2):  # second pass re-uses parked blocks
        a = SolverEFResampling(pb, bound, max_depth=4, device=0, max_sites=1 << 17)
        b = SolverEFResampling(pb, bound, max_depth=4, device=1, max_sites=1 << 17)
        a.setup()
        b.setup()
        for _ in range(4):
            a.step()
            b.step()
        a.check_solutions()
        b.check_solutions()
        assert list(a.sites) == list(b.sites) and a.solution_cost == b.solution_cost
        for name, site in a.sites.items():
            assert np.array_equal(site.traj.states, b.sites[name].traj.states, equal_nan=True), name
        b.close()
        a.close()
