import copy

import numpy as np

from oracle import fql_oracle as O


def rel_err(a, ref):
    """tensor-norm-relative error max|a-ref| / max|ref| (element-wise rtol is meaningless on near-zero grads)."""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    den = np.abs(ref).max()
    if den == 0:
        return np.abs(a).max()
    return np.abs(a - ref).max() / den


def make_case(cfg_over, B, F, A, seed=0, hidden=None, jitter=0.05, warm=True, onestep_bias_push=True):
    cfg = dict(O.DEFAULT_CONFIG)
    cfg.update(cfg_over)
    if hidden is not None:
        cfg.update(actor_hidden_dims=(hidden,) * 4, value_hidden_dims=(hidden,) * 4)
    params = O.init_params(seed, F, A, cfg, dtype=np.float64, jitter=jitter, target_equals_critic=False)
    if onestep_bias_push:
        # push part of the one-step policy's outputs outside [-1, 1] so the clip gates (fql.py:26,69,152) are exercised
        b = params['modules_actor_onestep_flow']['mlp']['Dense_4']['bias']
        b[0] += 1.1
        if A > 1:
            b[1] -= 1.1
    state = O.init_state(params, warm=warm, seed=seed)
    batch = O.make_batch(seed + 1, B, F, A, np.float64)
    noise = O.make_noise(seed + 2, B, A, np.float64)
    return cfg, state, batch, noise


def f32(t):
    return O.cast_tree(t, np.float32)


def cuda_agent_from_state(cfg, state, B, F, A, num_seeds=1, precision='fp32'):
    from fql_b200 import FQLAgent
    c = dict(cfg)
    c['batch_size'] = B
    agent = FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), c, num_seeds=num_seeds, precision=precision)
    return agent


def stack_trees(trees):
    if isinstance(trees[0], dict):
        return {k: stack_trees([t[k] for t in trees]) for k in trees[0]}
    return np.stack(trees, 0)


def info_close(key, got, ref_info, tol):
    """Scalars that are means of signed quantities (q, q_loss, q_mean) cancel: their tolerance is relative to the scale
    of the quantity being averaged (max |q|), not to the possibly near-zero mean itself."""
    r = float(ref_info[key])
    scale = max(abs(r), 1e-3)
    if key in ('actor/q', 'actor/q_loss', 'critic/q_mean', 'actor/actor_loss'):
        qs = max(abs(float(ref_info['critic/q_max'])), abs(float(ref_info['critic/q_min'])), abs(float(ref_info['actor/q'])))
        if key == 'actor/q_loss' and abs(float(ref_info['actor/q'])) > 0:
            qs *= abs(r) / abs(float(ref_info['actor/q']))       # lam factor when normalize_q_loss
        scale = max(scale, qs)
    assert abs(got - r) <= tol * scale, (key, got, r, scale)
