"""Host-side logic of the data-parallel path on CPU with gloo, world_size 2: sharding + the all-reduce semantics of
fql_b200.dist reproduce the single-rank step.  The per-rank compute stand-in here is the oracle (the product has no CPU path);
the CUDA side of the same decomposition is tests/test_dp_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fql_oracle as O
from tests.helpers import info_from_raw, raw_from_losses


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(tree):
    return np.concatenate([np.ravel(v) for _, v in O.tree_leaves(tree)])


def _worker(rank, world, port, q_agg, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fql_b200 import dist as fdist
    cfg = dict(O.DEFAULT_CONFIG)
    cfg.update(actor_hidden_dims=(32,) * 4, value_hidden_dims=(32,) * 4, q_agg=q_agg, alpha=10.0)
    B, F, A = 16, 6, 3
    params = O.init_params(0, F, A, cfg, dtype=np.float64, jitter=0.1, target_equals_critic=False)
    batch, noise = O.make_batch(1, B, F, A, np.float64), O.make_noise(2, B, A, np.float64)
    lb, ln = fdist.shard_rows(batch, rank, world), fdist.shard_rows(noise, rank, world)
    assert lb['actions'].shape[0] == B // world
    _, info, grads = O.total_loss(params, cfg, lb, ln)
    # a rank contributes its local-mean gradient scaled by local/global rows (the CUDA step divides by global_batch directly)
    g = torch.from_numpy(_flat(grads) * (B // world) / B).reshape(1, -1)
    qs = O.critic_forward(params['modules_critic'], cfg, lb['observations'], np.clip(O.actor_forward(
        params['modules_actor_onestep_flow'], cfg, lb['observations'], ln['z']), -1, 1)).mean(0)
    raw = torch.from_numpy(raw_from_losses({k: float(v) for k, v in info.items()}, None, qs, B // world, A)).reshape(1, -1)
    fdist.allreduce_step(g, raw, group=None)
    if rank == 0:
        _, info_full, grads_full = O.total_loss(params, cfg, batch, noise)
        np.testing.assert_allclose(g.numpy()[0], _flat(grads_full), rtol=1e-9, atol=1e-12)
        got = info_from_raw(raw.numpy()[0], B, A, cfg['alpha'], cfg['normalize_q_loss'])
        for k, v in got.items():
            np.testing.assert_allclose(v, float(info_full[k]), rtol=1e-9, atol=1e-12, err_msg=k)
        out.put('ok')
    dist.destroy_process_group()


@pytest.mark.parametrize('q_agg', ['mean', 'min'])
def test_two_rank_gloo_matches_single_rank(q_agg):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_agg, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 'ok'


def test_shard_helpers():
    from fql_b200 import dist as fdist
    b = {'x': np.arange(12).reshape(6, 2)}
    assert np.array_equal(np.concatenate([fdist.shard_rows(b, r, 3)['x'] for r in range(3)]), b['x'])
    with pytest.raises(ValueError):
        fdist.shard_rows(b, 0, 4)
    assert [list(fdist.shard_seeds(64, r, 8)) for r in (0, 7)] == [list(range(8)), list(range(56, 64))]
    with pytest.raises(ValueError):
        fdist.shard_seeds(10, 0, 4)
