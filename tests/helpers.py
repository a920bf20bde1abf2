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


def check_update_delta(old_params, ref_new_params, got_new_params, tol, pick=None, what='', opt=None):
    """The parameter UPDATE itself (new - old), per leaf.  Comparing the new VALUES has no power for one Adam step (the step moves
    a weight by ~lr = 3e-4 against |p| ~ 0.1: a value-level bound a bf16 path can meet is also met by an optimizer that never
    ran, whose update error is 1.0).  Two checks:
      (a) opt = dict(state=..., cfg=..., grads=<the gradients the device step produced>): the device's update equals the oracle's
          Adam / Polyak arithmetic applied to THOSE gradients, to 2e-3 (fp32 rounding of p + dp) -- the optimizer itself is exact
          whatever the precision of the gradients;
      (b) against the reference update: tensor-norm-relative error <= tol (and, without (a), the sign of every significant
          element agrees).  Adam divides by sqrt(v): with the warm moments of these cases d(update)/d(gradient) reaches ~16 lr per
          unit gradient, so a gradient error of 1e-2 of the leaf's max legitimately moves some elements by a large fraction of a
          whole step (measured up to 0.67 of the leaf's largest step).  In bf16 mode the PRECISION of the update is therefore
          established by (a) together with the per-leaf gradient bound; (b) only has to tell a real update (error < 1) from a missing
          (1.0) or mirrored (2.0) one."""
    pick = pick or (lambda x: x)
    worst = 0.0
    exp_new = None
    if opt is not None:
        st, cfg = opt['state'], opt['cfg']
        old32 = O.cast_tree(O.cast_tree(st['params'], np.float32), np.float64)
        g64 = O.tree_map(lambda g: np.asarray(pick(g), np.float64), opt['grads'])
        exp_new, _, _, _ = O.adam_update(old32, g64, O.cast_tree(O.cast_tree(st['mu'], np.float32), np.float64),
                                         O.cast_tree(O.cast_tree(st['nu'], np.float32), np.float64), st['count'], cfg['lr'])
        tau = cfg['tau']
        exp_new['modules_target_critic'] = O.tree_map(lambda p, tp: p * tau + tp * (1 - tau), old32['modules_critic'], old32['modules_target_critic'])
        for (path, e_new), (_, old), (_, g) in zip(O.tree_leaves(exp_new), O.tree_leaves(old32), O.tree_leaves(got_new_params)):
            d_exp = e_new - old
            d_got = np.asarray(pick(g), np.float64) - old
            if np.abs(d_exp).max() > 0:
                assert rel_err(d_got, d_exp) <= 2e-3, (what, 'optimizer arithmetic on the device gradients', path, rel_err(d_got, d_exp))
    for (path, new), (_, old), (_, g) in zip(O.tree_leaves(ref_new_params), O.tree_leaves(old_params), O.tree_leaves(got_new_params)):
        new, old = np.asarray(new, np.float64), np.asarray(old, np.float64)
        d_ref = new - old
        if np.abs(d_ref).max() == 0:
            assert np.array_equal(np.asarray(pick(g), np.float64), old.astype(np.float32).astype(np.float64)), (what, 'frozen leaf moved', path)
            continue
        d_got = np.asarray(pick(g), np.float64) - old.astype(np.float32).astype(np.float64)
        e = rel_err(d_got, d_ref)
        worst = max(worst, e)
        # with (a) the end-to-end figure is reported, not bounded tightly: it is ill-conditioned (see the docstring); 1.5 still
        # separates it from a mirrored update (2.0), and a missing update (1.0) has already failed (a)
        assert e <= (tol if opt is None else max(tol, 1.5)), (what, 'delta', path, e)
        if opt is None:
            big = np.abs(d_ref) > 0.5 * np.abs(d_ref).max()
            assert np.array_equal(np.sign(d_got[big]), np.sign(d_ref[big])), (what, 'delta sign', path)
    return worst


# ---- reduction semantics of the data-parallel metric accumulators, in NumPy (used by the CPU tests; include/fql_b200.h raw[]) ----
def raw_from_losses(info: dict, q: np.ndarray, q_pi: np.ndarray, local_rows: int, action_dim: int) -> np.ndarray:
    """Raw accumulators a rank would produce, reconstructed from per-rank MEAN metrics (used by the CPU tests to exercise the
    reduction semantics against the oracle)."""
    raw = np.zeros(16, np.float64)
    B, A = local_rows, action_dim
    raw[0] = info['critic/critic_loss'] * 2 * B
    raw[1] = info['critic/q_mean'] * 2 * B
    raw[2] = info['actor/bc_flow_loss'] * B * A
    raw[3] = info['actor/distill_loss'] * B * A
    raw[4] = info['actor/q'] * B
    raw[5] = np.abs(q_pi).sum()
    raw[6] = info['actor/mse'] * B * A
    raw[9] = info['critic/q_max']
    raw[10] = -info['critic/q_min']
    return raw


def info_from_raw(raw: np.ndarray, global_rows: int, action_dim: int, alpha: float, normalize_q_loss: bool) -> dict:
    """finalize_info_kernel in NumPy (losses.cu): the 10 loss metrics from the reduced accumulators."""
    gb, A = float(global_rows), float(action_dim)
    bc, distill, q = raw[2] / (gb * A), raw[3] / (gb * A), raw[4] / gb
    q_loss = -q
    if normalize_q_loss:
        q_loss = q_loss / (raw[5] / gb)
    return {'critic/critic_loss': raw[0] / (2 * gb), 'critic/q_mean': raw[1] / (2 * gb), 'critic/q_max': raw[9], 'critic/q_min': -raw[10],
            'actor/actor_loss': bc + alpha * distill + q_loss, 'actor/bc_flow_loss': bc, 'actor/distill_loss': distill, 'actor/q_loss': q_loss,
            'actor/q': q, 'actor/mse': raw[6] / (gb * A)}
