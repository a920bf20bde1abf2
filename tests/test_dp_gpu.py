"""Data-parallel decomposition on the GPU: (a) two ranks emulated on ONE device through the C ABI's fql_step_grads /
fql_step_apply split reproduce the single-rank oracle step; (b) with >= 2 GPUs, two NCCL processes do the same."""
import copy
import os
import socket

import numpy as np
import pytest
import torch

from oracle import fql_oracle as O
from tests.helpers import check_update_delta, cuda_agent_from_state, f32, info_close, make_case, rel_err

pytestmark = pytest.mark.gpu


def _mk(cfg, B, F, A, world, precision):
    from fql_b200 import FQLAgent
    c = dict(cfg)
    c['batch_size'] = B // world
    return FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), c, precision=precision, world_size=world)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 8e-2)])
def test_two_emulated_ranks_match_single_rank_oracle(precision, tol):
    from fql_b200 import dist as fdist
    B, F, A, H = 128, 29, 8, 512
    cfg, state, batch, noise = make_case(dict(q_agg='min', alpha=10.0), B, F, A, seed=5, hidden=H)
    new_state, ref_info, ref_grads = O.update(copy.deepcopy(state), cfg, batch, noise)
    agents, bufs = [], []
    for r in range(2):
        a = _mk(cfg, B, F, A, 2, precision)
        a.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
        b = a.stage(f32(fdist.shard_rows(batch, r, 2)), f32(fdist.shard_rows(noise, r, 2)))
        a.grads_phase(b)
        agents.append(a)
        bufs.append(b)
    torch.cuda.synchronize()
    g = agents[0]._grads + agents[1]._grads
    raw = bufs[0]['raw'].clone()
    raw[:, :9] = bufs[0]['raw'][:, :9] + bufs[1]['raw'][:, :9]
    raw[:, 9:11] = torch.maximum(bufs[0]['raw'][:, 9:11], bufs[1]['raw'][:, 9:11])
    for a, b in zip(agents, bufs):
        a._grads.copy_(g)
        b['raw'].copy_(raw)
        a.apply_phase(b)
    info = agents[0]._info_out(bufs[0]['info'])
    for k in O.INFO_KEYS:
        if k.startswith('grad/') and precision == 'bf16':
            continue
        info_close(k, info[k], ref_info, 3 * tol if precision == 'fp32' else 5e-2)
    for which, ref, t in (('grads', ref_grads, tol), ('params', new_state['params'], tol if precision == 'fp32' else 3e-3)):
        got = agents[0].export_tree(which)
        for (path, r), (_, gg) in zip(O.tree_leaves(ref), O.tree_leaves(got)):
            assert rel_err(gg, r) <= t, (which, path, rel_err(gg, r))
    p0, p1 = agents[0].export_tree('params'), agents[1].export_tree('params')
    for (_, x), (_, y) in zip(O.tree_leaves(p0), O.tree_leaves(p1)):
        assert np.array_equal(x, y)                      # replicas stay bit-identical without a parameter broadcast


def _dp_worker(rank, world, port, ret, backend, precision, over):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ['FQL_DP_BACKEND'] = 'nccl' if backend == 'nccl' else 'peer'
    os.environ['FQL_DP_MULTICAST'] = '0' if backend == 'peer-loads' else '1'
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    from fql_b200 import FQLAgent, dist as fdist
    B, F, A, H = 256, 29, 8, 512
    cfg, state, batch, noise = make_case(over, B, F, A, seed=6, hidden=H)
    c = dict(cfg)
    c['batch_size'] = B // world
    agent = FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), c, process_group=dist.group.WORLD, precision=precision)
    assert agent._dp_peer == (backend != 'nccl')
    agent.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    tol_info, tol_grad = (3e-5, 1e-5) if precision == 'fp32' else (5e-2, 4e-2)
    st = copy.deepcopy(state)
    for i in range(3):                                   # eager, graph capture, graph replay
        ba, nz = (batch, noise) if i == 0 else (O.make_batch(40 + i, B, F, A, np.float64), O.make_noise(50 + i, B, A, np.float64))
        prev, prev_mu, prev_nu = copy.deepcopy(st['params']), copy.deepcopy(st['mu']), copy.deepcopy(st['nu'])
        st, ref_info, ref_grads = O.update(st, cfg, ba, nz)
        _, info = agent.update(f32(fdist.shard_rows(ba, rank, world)), noise=f32(fdist.shard_rows(nz, rank, world)))
        for k in O.INFO_KEYS[:10]:
            info_close(k, info[k], ref_info, tol_info)
        if i == 0:
            for (path, r), (_, g) in zip(O.tree_leaves(ref_grads), O.tree_leaves(agent.export_tree('grads'))):
                assert rel_err(g, r) <= tol_grad, ('grads', path, rel_err(g, r))   # the arena holds the REDUCED gradient
            check_update_delta(prev, st['params'], agent.export_tree('params'), 2e-3 if precision == 'fp32' else 0.9, what=f'rank {rank}',
                               opt=dict(state=dict(st, params=prev, mu=prev_mu, nu=prev_nu, count=st['count'] - 1), cfg=cfg, grads=agent.export_tree('grads')))
    if precision == 'fp32':
        worst = max(rel_err(g, r) for (_, r), (_, g) in zip(O.tree_leaves(st['params']), O.tree_leaves(agent.export_tree('params'))))
        assert worst <= 3e-5, worst
    # forward-only losses describe the GLOBAL batch on every rank (peer-memory transport; with the NCCL A/B backend the forward-only
    # entry point reports this rank's rows)
    if backend != 'nccl':
        loss, vinfo = agent.total_loss(f32(fdist.shard_rows(batch, rank, world)), noise=f32(fdist.shard_rows(noise, rank, world)))
        _, ref_v, _ = O.total_loss(st['params'], cfg, batch, noise, with_grads=False)
        info_close('critic/critic_loss', vinfo['critic/critic_loss'], ref_v, 10 * tol_info)
    # replicas stay bit-identical without a parameter broadcast
    mine = agent._params.view(torch.int32).sum(dtype=torch.int64).reshape(1)
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert all(int(x) == int(both[0]) for x in both), [int(x) for x in both]
    # production noise (noise=None): ranks share the key but draw disjoint rows of the global tensors
    agent.update(f32(fdist.shard_rows(batch, rank, world)))
    z = agent._bufs[B // world]['dev']['z'].clone()
    zs = [torch.zeros_like(z) for _ in range(world)]
    dist.all_gather(zs, z)
    assert not torch.equal(zs[0], zs[1])
    if rank == 0:
        ret.put('ok')
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('backend,precision,over', [
    ('peer', 'fp32', dict(q_agg='min', alpha=10.0)),
    ('peer', 'bf16', dict(q_agg='min', alpha=10.0)),
    ('peer', 'fp32', dict(normalize_q_loss=True, alpha=1000.0)),     # global mean|q| exchange inside the loss kernel
    ('peer-loads', 'bf16', dict(alpha=300.0)),                       # no NVLS multicast: peer loads / stores
    ('nccl', 'bf16', dict(q_agg='min', alpha=10.0)),
], ids=['peer-fp32', 'peer-bf16', 'peer-fp32-normq', 'peerloads-bf16', 'nccl-bf16'])
def test_two_ranks_data_parallel_step(backend, precision, over):
    """Two processes, one per GPU: the data-parallel step (library kernels over NVLink peer memory / NVLS multicast by default,
    NCCL as the A/B alternative) reproduces the single-device oracle step on the concatenated batch."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ret = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, ret, backend, precision, over)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(420)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 'ok'


def _dp_pixel_worker(rank, world, port, ret, precision):
    """Pixel configuration, two ranks: rows shard, the encoder + MLP gradients of the whole trainable arena are exchanged once behind
    the backward (the encoder backwards finish on forked streams), every rank applies the same update."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    from fql_b200 import FQLAgent, dist as fdist
    from oracle import fql_pixel_oracle as PO
    from oracle.encoder_oracle import bf16_round
    B, hw, ch, A, hidden = 8, 16, 6, 3, 512
    cfg = dict(O.DEFAULT_CONFIG)
    cfg.update(alpha=10.0, actor_hidden_dims=(hidden,) * 4, value_hidden_dims=(hidden,) * 4, encoder='impala_small')
    params = PO.init_params(3, ch, A, cfg, dtype=np.float64, hw=hw, jitter=0.05, target_equals_critic=False)
    state = O.init_state(params, warm=True, seed=3)
    # frames on which no max-pool window is decided by fp32 rounding (seed 4 has a 1.3e-8 near-tie at 8 rows: the pooling gradient
    # of one channel then lands on another pixel than in the fp64 oracle -- measured, profiles/dbg_pix_fp32.py)
    batch, _ = PO.well_posed_pixel_batch(params, B, A, hw, ch, dtype=np.float64, first_seed=4)
    noise = O.make_noise(5, B, A, np.float64)
    q = None if precision == 'fp32' else bf16_round
    new_state, ref_info, ref_grads = PO.update(copy.deepcopy(state), cfg, batch, noise, enc_q=q)
    c = dict(cfg)
    c['batch_size'] = B // world
    agent = FQLAgent.create(0, np.zeros((1, hw, hw, ch), np.uint8), np.zeros((1, A), np.float32), c, process_group=dist.group.WORLD, precision=precision)
    agent.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    sh = lambda d: {k: (v if v.dtype == np.uint8 else v.astype(np.float32)) for k, v in fdist.shard_rows(d, rank, world).items()}
    _, info = agent.update(sh(batch), noise=sh(noise))
    tol_info, tol_grad = (3e-5, 1e-4) if precision == 'fp32' else (5e-2, 8e-2)
    for k in O.INFO_KEYS[:10]:
        info_close(k, info[k], ref_info, tol_info)
    for (path, r), (_, g) in zip(O.tree_leaves(ref_grads), O.tree_leaves(agent.export_tree('grads'))):
        if np.abs(r).max() > 0:
            assert rel_err(g, r) <= tol_grad, ('grads', path, rel_err(g, r))      # the arena holds the REDUCED gradient
    mine = agent._params.view(torch.int32).sum(dtype=torch.int64).reshape(1)
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert all(int(x) == int(both[0]) for x in both), [int(x) for x in both]
    if rank == 0:
        ret.put('ok')
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_two_ranks_data_parallel_pixel_step(precision):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ret = ctx.Queue()
    procs = [ctx.Process(target=_dp_pixel_worker, args=(r, 2, port, ret, precision)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(420)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 'ok'


def _nccl_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    from fql_b200 import FQLAgent, dist as fdist
    B, F, A, H = 256, 29, 8, 512
    cfg, state, batch, noise = make_case(dict(q_agg='min', alpha=10.0), B, F, A, seed=6, hidden=H)
    c = dict(cfg)
    c['batch_size'] = B // world
    agent = FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), c, process_group=dist.group.WORLD)
    agent.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    _, info = agent.update(f32(fdist.shard_rows(batch, rank, world)), noise=f32(fdist.shard_rows(noise, rank, world)))
    new_state, ref_info, _ = O.update(copy.deepcopy(state), cfg, batch, noise)
    for k in O.INFO_KEYS:
        info_close(k, info[k], ref_info, 3e-5)
    got = agent.export_tree('params')
    worst = max(rel_err(g, r) for (_, r), (_, g) in zip(O.tree_leaves(new_state['params']), O.tree_leaves(got)))
    assert worst <= 1e-5, worst
    if rank == 0:
        ret.put(worst)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_nccl_ranks_match_single_rank_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ret = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=5) <= 1e-5
