"""Parity of the CUDA path (through the C ABI) against the oracle.  FP32 mode tolerance: 1e-5 tensor-norm-relative
(north_star: "within 1e-5 relative in fp32 mode"), measured against the fp64 oracle; the fp32 oracle's own distance
to fp64 (~1e-6, tests/test_oracle.py::test_fp32_vs_fp64_gap) calibrates it."""
import copy

import numpy as np
import pytest

from oracle import fql_oracle as O
from tests.helpers import cuda_agent_from_state, f32, info_close, make_case, rel_err, stack_trees

pytestmark = pytest.mark.gpu
TOL = 1e-5

CASES = [
    # (name, cfg overrides, B, F, A, hidden)
    ('tiny-default', dict(), 24, 7, 3, 32),
    ('tiny-min-odd', dict(q_agg='min', alpha=10.0), 37, 11, 5, 48),
    ('tiny-normq', dict(normalize_q_loss=True, alpha=1000.0), 16, 9, 2, 32),
    ('tiny-actor-ln', dict(actor_layer_norm=True), 20, 6, 4, 64),
    ('tiny-no-critic-ln', dict(layer_norm=False, flow_steps=3, discount=0.995), 33, 5, 1, 32),
    ('cube-single', dict(alpha=300.0), 256, 28, 5, 512),                       # BASELINE config 1
    ('antmaze-large', dict(q_agg='min', alpha=10.0), 256, 29, 8, 512),         # BASELINE config 2
    ('humanoidmaze-medium', dict(discount=0.995, alpha=30.0), 256, 69, 21, 512),  # BASELINE config 3 @256
    ('puzzle-4x4', dict(normalize_q_loss=True, alpha=1000.0), 256, 83, 5, 512),   # BASELINE config 4, one seed
]


def check_update(agent, cfg, state, batch, noise, seeds=None):
    ref_states, ref_infos, ref_grads = [], [], []
    seeds = seeds or [(state, batch, noise)]
    for st, ba, nz in seeds:
        ns, info, grads = O.update(copy.deepcopy(st), cfg, ba, nz)
        ref_states.append(ns)
        ref_infos.append(info)
        ref_grads.append(grads)
    S = len(seeds)
    cat = (lambda xs: xs[0]) if S == 1 else stack_trees
    agent.load_tree(f32(cat([s[0]['params'] for s in seeds])), f32(cat([s[0]['mu'] for s in seeds])),
                    f32(cat([s[0]['nu'] for s in seeds])), seeds[0][0]['count'])
    b = {k: np.stack([s[1][k] for s in seeds]) if S > 1 else seeds[0][1][k] for k in seeds[0][1]}
    n = {k: np.stack([s[2][k] for s in seeds]) if S > 1 else seeds[0][2][k] for k in seeds[0][2]}
    _, info = agent.update(f32(b), noise=f32(n))
    got = {w: agent.export_tree(w) for w in ('params', 'mu', 'nu', 'grads')}
    worst = {}
    for si in range(S):
        pick = (lambda x: x) if S == 1 else (lambda x: x[si])
        for k in O.INFO_KEYS:
            info_close(k, info[k] if S == 1 else info[k][si], ref_infos[si], 3 * TOL)
        for which, ref in (('grads', ref_grads[si]), ('params', ref_states[si]['params']), ('mu', ref_states[si]['mu']),
                           ('nu', ref_states[si]['nu'])):
            for (path, r), (_, g) in zip(O.tree_leaves(ref), O.tree_leaves(got[which])):
                e = rel_err(pick(g), r)
                worst[which] = max(worst.get(which, 0), e)
                assert e <= TOL, (which, path, e)
        # the parameter UPDATE itself (new - old), not just the new value
        for (path, new), (_, old), (_, g) in zip(O.tree_leaves(ref_states[si]['params']), O.tree_leaves(seeds[si][0]['params']),
                                                 O.tree_leaves(got['params'])):
            d_ref = new - old
            d_got = pick(g).astype(np.float64) - old.astype(np.float32).astype(np.float64)
            if np.abs(d_ref).max() > 0:
                assert rel_err(d_got, d_ref) <= 2e-3, ('delta', path, rel_err(d_got, d_ref))  # fp32 rounding of p+dp
    assert agent.network.step == seeds[0][0]['step'] + 1
    return worst


@pytest.mark.parametrize('name,over,B,F,A,hidden', CASES, ids=[c[0] for c in CASES])
def test_update_step_parity(name, over, B, F, A, hidden):
    cfg, state, batch, noise = make_case(over, B, F, A, seed=sum(map(ord, name)) % 1000, hidden=hidden)
    agent = cuda_agent_from_state(cfg, state, B, F, A)
    worst = check_update(agent, cfg, state, batch, noise)
    print(name, {k: f'{v:.2e}' for k, v in worst.items()})


def test_two_consecutive_steps_and_graph_replay():
    """Step 1 runs eagerly, step 2 captures the CUDA graph, step 3 replays it: all three must match the oracle."""
    cfg, state, batch, noise = make_case(dict(q_agg='min', alpha=10.0), 64, 29, 8, seed=3, hidden=128)
    agent = cuda_agent_from_state(cfg, state, 64, 29, 8)
    agent.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    st = copy.deepcopy(state)
    for i in range(4):
        ba = O.make_batch(100 + i, 64, 29, 8, np.float64)
        nz = O.make_noise(200 + i, 64, 8, np.float64)
        st, info_ref, _ = O.update(st, cfg, ba, nz)
        _, info = agent.update(f32(ba), noise=f32(nz))
        for k in O.INFO_KEYS:
            info_close(k, info[k], info_ref, 5e-5)
        got = agent.export_tree('params')
        for (path, r), (_, g) in zip(O.tree_leaves(st['params']), O.tree_leaves(got)):
            assert rel_err(g, r) <= 1e-5 * (i + 1), (i, path)
    assert agent.network.step == state['step'] + 4


def test_multi_seed_matches_per_seed():
    """BASELINE config 4 shape: seeds are one more leading axis; every seed must equal its own single-seed oracle run."""
    over = dict(normalize_q_loss=True, alpha=1000.0)
    seeds = []
    for s in range(3):
        cfg, state, batch, noise = make_case(over, 32, 13, 5, seed=50 + s, hidden=64)
        seeds.append((state, batch, noise))
    agent = cuda_agent_from_state(cfg, seeds[0][0], 32, 13, 5, num_seeds=3)
    check_update(agent, cfg, None, None, None, seeds=seeds)


def test_total_loss_forward_only():
    cfg, state, batch, noise = make_case(dict(), 48, 17, 6, seed=8, hidden=64)
    agent = cuda_agent_from_state(cfg, state, 48, 17, 6)
    agent.load_tree(f32(state['params']))
    before = agent.export_tree('params')
    loss, info = agent.total_loss(f32(batch), noise=f32(noise))
    ref_loss, ref_info, _ = O.total_loss(state['params'], cfg, batch, noise, with_grads=False)
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    for k in ref_info:
        info_close(k, info[k], ref_info, 3 * TOL)
    after = agent.export_tree('params')
    for (_, a), (_, b) in zip(O.tree_leaves(before), O.tree_leaves(after)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('rows', [1, 10, 300])
def test_sample_actions_and_flow_actions(rows):
    cfg, state, _, _ = make_case(dict(), 8, 29, 8, seed=21, hidden=512)
    agent = cuda_agent_from_state(cfg, state, 8, 29, 8)
    agent.load_tree(f32(state['params']))
    rng = np.random.default_rng(rows)
    obs = rng.standard_normal((rows, 29))
    nz = rng.standard_normal((rows, 8))
    a = agent.sample_actions(obs.astype(np.float32), noise=nz.astype(np.float32))
    ref = O.sample_actions_given_noise(state['params'], cfg, obs, nz)
    assert a.shape == (rows, 8) and rel_err(a, ref) <= TOL
    fa = agent.compute_flow_actions(obs.astype(np.float32), nz.astype(np.float32))
    ref = O.compute_flow_actions(state['params'], cfg, obs, nz)
    assert rel_err(fa, ref) <= TOL
    # unbatched observation, like the online loop (main.py:225)
    a1 = agent.sample_actions(obs[0].astype(np.float32), noise=nz[0].astype(np.float32))
    assert a1.shape == (8,) and rel_err(a1, O.sample_actions_given_noise(state['params'], cfg, obs[:1], nz[:1])[0]) <= TOL
    # seed-driven draw is deterministic and clipped
    s1 = agent.sample_actions(obs.astype(np.float32), seed=np.array([1, 2], np.uint32))
    s2 = agent.sample_actions(obs.astype(np.float32), seed=np.array([1, 2], np.uint32))
    assert np.array_equal(s1, s2) and np.abs(s1).max() <= 1.0


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_q_values_and_best_of_n_action_selection(precision):
    """IFQLAgent.sample_actions (agents/ifql.py:122-149) on the FQL networks: Euler-integrate N noises through the bc-flow field, pick the
    action with the largest min-over-heads Q.  Q values against the oracle's critic (1e-5); the chosen action is the oracle's argmax
    (or, in bf16 where the Euler integration carries 1e-3, an action whose oracle Q is within 1e-2 of the best)."""
    cfg, state, _, _ = make_case(dict(), 8, 29, 8, seed=23, hidden=512)
    agent = cuda_agent_from_state(cfg, state, 8, 29, 8, precision=precision)
    agent.load_tree(f32(state['params']))
    rng = np.random.default_rng(3)
    ob = rng.standard_normal(29)
    nz = rng.standard_normal((32, 8))
    obs_n = np.repeat(ob[None], 32, 0)
    acts = O.compute_flow_actions(state['params'], cfg, obs_n, nz)
    q_ref = O.critic_forward(state['params']['modules_critic'], cfg, obs_n, acts)
    q = agent.q_values(obs_n.astype(np.float32), acts.astype(np.float32))
    assert q.shape == (2, 32) and rel_err(q, q_ref) <= TOL
    a = agent.sample_actions_best_of_n(ob.astype(np.float32), noise=nz.astype(np.float32))
    qmin = np.asarray(q_ref).min(axis=0)
    best = int(np.argmax(qmin))
    if precision == 'fp32':
        assert rel_err(a, acts[best]) <= TOL
    else:
        k = int(np.argmin(np.abs(acts - a[None]).max(axis=1)))           # which sample the device picked
        assert np.abs(acts[k] - a).max() <= 5e-2 and qmin[k] >= qmin[best] - 1e-2 * np.abs(qmin).max()


def test_large_batch_properties():
    """BASELINE config 3 at a large batch: checked against the oracle through size-independent properties:
    (a) info of a batch made of 8 copies of a 256-row block == info of the block (means are replication-invariant),
    (b) its gradients equal the block's gradients."""
    cfg, state, batch, noise = make_case(dict(discount=0.995, alpha=30.0), 256, 69, 21, seed=77, hidden=512, warm=False)
    _, info_ref, grads_ref = O.update(copy.deepcopy(state), cfg, batch, noise)
    rep = 8
    big_b = {k: np.concatenate([v] * rep, 0) for k, v in batch.items()}
    big_n = {k: np.concatenate([v] * rep, 0) for k, v in noise.items()}
    agent = cuda_agent_from_state(cfg, state, 256 * rep, 69, 21)
    agent.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    _, info = agent.update(f32(big_b), noise=f32(big_n))
    for k in O.INFO_KEYS:
        info_close(k, info[k], info_ref, 3 * TOL)
    got = agent.export_tree('grads')
    for (path, r), (_, g) in zip(O.tree_leaves(grads_ref), O.tree_leaves(got)):
        assert rel_err(g, r) <= TOL, path


def test_device_noise_statistics():
    """The production noise source (Philox, fql_fill_noise): N(0,1) / U[0,1) moments and step-to-step independence."""
    import ctypes as C
    import torch
    from fql_b200 import _lib
    d = _lib.make_dims(4096, 8, 16)
    bufs = [torch.empty(4096 * 16, device='cuda') for _ in range(5)]
    bufs[2] = torch.empty(4096, device='cuda')
    P = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().fql_fill_noise(C.byref(d), C.c_uint64(7), C.c_uint64(0), *[P(b) for b in bufs], None), 'noise')
    first = [b.clone() for b in bufs]
    _lib.check(_lib.lib().fql_fill_noise(C.byref(d), C.c_uint64(7), C.c_uint64(1), *[P(b) for b in bufs], None), 'noise')
    torch.cuda.synchronize()
    for i in (0, 1, 3, 4):
        x = first[i].cpu().numpy()
        assert abs(x.mean()) < 0.02 and abs(x.std() - 1) < 0.02 and abs((x ** 3).mean()) < 0.05 and abs((x ** 4).mean() - 3) < 0.15
        assert not np.array_equal(x, bufs[i].cpu().numpy())
    t = first[2].cpu().numpy()
    assert t.min() >= 0 and t.max() < 1 and abs(t.mean() - 0.5) < 0.02
    assert abs(np.corrcoef(first[0].cpu().numpy(), first[1].cpu().numpy())[0, 1]) < 0.02
