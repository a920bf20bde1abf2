"""update(host batch) stages the batch (and draws the noise) on a copy stream into the input set the running step is not reading
(fql_b200/agent.py `_update_overlapped`): same results, bit for bit, as the single-stream path, whatever is interleaved with it."""
import numpy as np
import pytest

from oracle import fql_oracle as O
from tests.helpers import cuda_agent_from_state, f32, make_case

pytestmark = pytest.mark.gpu


def _run(overlap, precision, B, F, A, hidden, steps=7, explicit_noise=False):
    cfg, state, _, _ = make_case(dict(q_agg='min', alpha=10.0), B, F, A, seed=3, hidden=hidden)
    a = cuda_agent_from_state(cfg, state, B, F, A, precision=precision)
    a._overlap_h2d = overlap
    a.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    infos = []
    for i in range(steps):
        batch = f32(O.make_batch(100 + i, B, F, A, np.float64))
        noise = f32(O.make_noise(200 + i, B, A, np.float64)) if explicit_noise else None
        _, info = a.update(batch, noise=noise)
        infos.append(info)
        if i == 2:      # forward-only call on the base input set between two overlapped updates
            a.total_loss(f32(O.make_batch(999, B, F, A, np.float64)), noise=f32(O.make_noise(998, B, A, np.float64)))
        if i == 4:      # the stage / step API (device-resident loop of bench.py) between two overlapped updates
            bufs = a.stage(f32(O.make_batch(777, B, F, A, np.float64)), f32(O.make_noise(776, B, A, np.float64)))
            a.step(bufs, fill_noise=False)
    vals = [float(np.ravel(x['critic/critic_loss'])[0]) for x in infos]
    return a.export_tree('params'), a.export_tree('mu'), vals


@pytest.mark.parametrize('precision,B,hidden,explicit', [('fp32', 48, 64, True), ('bf16', 256, 512, False), ('bf16', 256, 512, True)])
def test_overlapped_host_updates_are_bit_identical(precision, B, hidden, explicit):
    F, A = 29, 8
    p1, m1, v1 = _run(True, precision, B, F, A, hidden, explicit_noise=explicit)
    p0, m0, v0 = _run(False, precision, B, F, A, hidden, explicit_noise=explicit)
    assert v0 == v1, (v0, v1)
    for (path, x), (_, y) in zip(O.tree_leaves(p0), O.tree_leaves(p1)):
        assert np.array_equal(x, y), ('params', path)
    for (path, x), (_, y) in zip(O.tree_leaves(m0), O.tree_leaves(m1)):
        assert np.array_equal(x, y), ('mu', path)
    assert len(set(v1)) == len(v1)       # every step saw its own batch


def test_overlapped_path_uses_two_input_sets():
    B, F, A = 32, 13, 5
    cfg, state, batch, noise = make_case(dict(), B, F, A, seed=5, hidden=64)
    a = cuda_agent_from_state(cfg, state, B, F, A)
    a.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    for _ in range(3):
        a.update(f32(batch), noise=f32(noise))
    base = a._bufs[B]
    assert len(base['sets']) == 2 and base['sets'][0]['dev_block'].data_ptr() != base['sets'][1]['dev_block'].data_ptr()
    assert base['sets'][0]['ws'].data_ptr() == base['sets'][1]['ws'].data_ptr()      # workspace, state and metrics are shared
