"""Checkpoint wire format (utils/flax_utils.py:162-202): params_{epoch}.pkl = pickle of dict(agent=<Flax state-dict nesting>)."""
import copy
import pickle

import numpy as np
import pytest

from oracle import fql_oracle as O
from tests.helpers import cuda_agent_from_state, f32, make_case

pytestmark = pytest.mark.gpu


def test_save_restore_round_trip(tmp_path):
    from fql_b200.checkpoint import restore_agent, save_agent
    B, F, A = 32, 13, 5
    cfg, state, batch, noise = make_case(dict(q_agg='min'), B, F, A, seed=7)
    a = cuda_agent_from_state(cfg, state, B, F, A)
    a.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    a.update(f32(batch), noise=f32(noise))                      # a non-trivial optimizer state
    save_agent(a, str(tmp_path), 1000)
    with open(tmp_path / 'params_1000.pkl', 'rb') as f:           # readable without this package: plain dicts of numpy arrays
        raw = pickle.load(f)
    net = raw['agent']['network']
    assert set(raw['agent']) == {'rng', 'network'} and set(net) == {'step', 'params', 'opt_state'}
    assert set(net['opt_state']) == {'0', '1'} and set(net['opt_state']['0']) == {'count', 'mu', 'nu'}
    assert set(net['params']) == {'modules_critic', 'modules_target_critic', 'modules_actor_bc_flow', 'modules_actor_onestep_flow'}
    k = net['params']['modules_critic']['value_net']['Dense_0']['kernel']
    assert isinstance(k, np.ndarray) and k.shape == (2, F + A, cfg['value_hidden_dims'][0])   # ensemblize axis first, [in, out]

    b = cuda_agent_from_state(cfg, state, B, F, A)               # fresh agent, different content
    restore_agent(b, str(tmp_path), 1000)
    for which in ('params', 'mu', 'nu'):
        for (path, x), (_, y) in zip(O.tree_leaves(a.export_tree(which)), O.tree_leaves(b.export_tree(which))):
            assert np.array_equal(x, y), (which, path)
    assert b.network.step == a.network.step
    ba, nz = O.make_batch(11, B, F, A, np.float64), O.make_noise(12, B, A, np.float64)
    _, ia = a.update(f32(ba), noise=f32(nz))
    _, ib = b.update(f32(ba), noise=f32(nz))
    for key in O.INFO_KEYS:
        assert float(ia[key]) == float(ib[key]), key             # bit-identical continuation
