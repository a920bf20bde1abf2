"""Checkpoint wire format on the host side only (no GPU): fql_b200.checkpoint writes the reference's `params_{epoch}.pkl`
(utils/flax_utils.py:162-202) as a pickle of plain dicts / numpy arrays and restores through `load_state_dict`."""
import pickle

import numpy as np

from fql_b200.checkpoint import restore_agent, save_agent


class _StubAgent:
    """Only the two methods the checkpoint module relies on."""

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        leaf = lambda *s: rng.standard_normal(s).astype(np.float32)
        tree = lambda: {'modules_critic': {'value_net': {'Dense_0': {'kernel': leaf(2, 7, 4), 'bias': leaf(2, 4)}}}}
        self.sd = {'rng': np.array([seed, 1], np.uint32),
                   'network': {'step': 5, 'params': tree(), 'opt_state': {'0': {'count': 4, 'mu': tree(), 'nu': tree()}, '1': {}}}}

    def state_dict(self):
        return self.sd

    def load_state_dict(self, sd):
        self.sd = sd
        return self


def test_checkpoint_file_is_plain_numpy_and_round_trips(tmp_path):
    a, b = _StubAgent(1), _StubAgent(2)
    save_agent(a, str(tmp_path), 7)
    with open(tmp_path / 'params_7.pkl', 'rb') as f:
        raw = pickle.load(f)
    assert set(raw) == {'agent'} and set(raw['agent']) == {'rng', 'network'}
    k = raw['agent']['network']['params']['modules_critic']['value_net']['Dense_0']['kernel']
    assert type(k) is np.ndarray and k.dtype == np.float32 and k.shape == (2, 7, 4)
    assert raw['agent']['network']['opt_state']['1'] == {}
    restore_agent(b, str(tmp_path), 7)
    np.testing.assert_array_equal(b.sd['network']['opt_state']['0']['nu']['modules_critic']['value_net']['Dense_0']['bias'],
                                  a.sd['network']['opt_state']['0']['nu']['modules_critic']['value_net']['Dense_0']['bias'])
    assert b.sd['network']['step'] == 5 and int(b.sd['network']['opt_state']['0']['count']) == 4
    np.testing.assert_array_equal(b.sd['rng'], a.sd['rng'])
