"""CPU oracle for FQLAgent.update with a visual encoder (BASELINE config 5) -- TEST INFRASTRUCTURE ONLY, parity unpinned.

Wraps oracle/fql_oracle.py with the `impala_small` encoders of oracle/encoder_oracle.py, following agents/fql.py:
  :196-202  three encoders (critic, actor_bc_flow, actor_onestep_flow); the target critic is a deepcopy incl. its own encoder (:226)
  :230-232  the bc-flow encoder is also a ModuleDict entry ('actor_bc_flow_encoder'); its parameters are kept THERE
            (modules_actor_bc_flow holds only `mlp`) -- SURVEY 8a, unverified against a real checkpoint
  :25       sample_actions(next_obs)        -> onestep encoder on next_obs (stored params)
  :28       target_critic(next_obs, a')     -> target critic encoder on next_obs
  :36       critic(obs, a; grad)            -> critic encoder on obs, gradient flows
  :58       actor_bc_flow(obs, x_t, t; grad)-> bc-flow encoder on obs, gradient flows
  :64,162   compute_flow_actions(obs)       -> bc-flow encoder (stored params) once, then is_encoded=True
  :65       actor_onestep_flow(obs; grad)   -> onestep encoder on obs, gradient flows
  :70       critic(obs, clip a_pi)          -> stored critic (incl. encoder): no encoder gradient from the Q loss
  :82       sample_actions(obs)             -> onestep encoder on obs (stored)
Stored and grad parameters are the same values inside one step, so 5 unique encoder forwards and 3 backwards.
"""
from __future__ import annotations

import numpy as np

from oracle import encoder_oracle as E
from oracle import fql_oracle as O

FEAT = 512


def init_params(seed, in_ch, action_dim, cfg, dtype=np.float32, hw=64, jitter=0.0, target_equals_critic=True):
    params = O.init_params(seed, FEAT, action_dim, cfg, dtype=dtype, jitter=jitter, target_equals_critic=target_equals_critic)
    rng = np.random.default_rng(seed + 4242)
    params['modules_critic']['encoder'] = E.init_encoder(rng, in_ch, dtype, hw, jitter)
    params['modules_actor_onestep_flow']['encoder'] = E.init_encoder(rng, in_ch, dtype, hw, jitter)
    params['modules_actor_bc_flow_encoder'] = E.init_encoder(rng, in_ch, dtype, hw, jitter)
    if target_equals_critic:
        params['modules_target_critic']['encoder'] = O.tree_map(lambda x: x.copy(), params['modules_critic']['encoder'])
    else:
        params['modules_target_critic']['encoder'] = E.init_encoder(rng, in_ch, dtype, hw, jitter)
    return params


def total_loss(params, cfg, batch, noise, with_grads=True, enc_q=None):
    """enc_q (optional): storage rounding of the encoders' activations / weight operands (oracle/encoder_oracle.py bf16_round)."""
    obs, nobs = batch['observations'], batch['next_observations']
    dt = batch['actions'].dtype
    enc = lambda p, x, save=False: E.encoder_forward(p, x, dtype=dt, save=save, q=enc_q)
    fC, sC = enc(params['modules_critic']['encoder'], obs, True)
    fF, sF = enc(params['modules_actor_bc_flow_encoder'], obs, True)
    fO, sO = enc(params['modules_actor_onestep_flow']['encoder'], obs, True)
    feats = {'C': fC, 'F': fF, 'O': fO, 'O_next': enc(params['modules_actor_onestep_flow']['encoder'], nobs),
             'T_next': enc(params['modules_target_critic']['encoder'], nobs)}
    loss, info, grads, dfeat = O.total_loss(params, cfg, batch, noise, with_grads=with_grads, feats=feats)
    if not with_grads:
        return loss, info, None
    grads['modules_critic']['encoder'] = E.encoder_backward(params['modules_critic']['encoder'], sC, dfeat['C'], q=enc_q)
    grads['modules_actor_bc_flow_encoder'] = E.encoder_backward(params['modules_actor_bc_flow_encoder'], sF, dfeat['F'], q=enc_q)
    grads['modules_actor_onestep_flow']['encoder'] = E.encoder_backward(params['modules_actor_onestep_flow']['encoder'], sO, dfeat['O'], q=enc_q)
    return loss, info, grads


def update(state, cfg, batch, noise, enc_q=None):
    """agents/fql.py:122-133 for the pixel configuration (same optimizer / Polyak as the state oracle)."""
    params = state['params']
    loss, info, grads = total_loss(params, cfg, batch, noise, enc_q=enc_q)
    gmax, gmin, gnorm = O.grad_stats(grads)
    info['grad/max'], info['grad/min'], info['grad/norm'] = gmax, gmin, gnorm
    new_p, new_m, new_v, new_count = O.adam_update(params, grads, state['mu'], state['nu'], state['count'], cfg['lr'])
    dtp = batch['actions'].dtype.type
    tau = dtp(cfg['tau'])
    new_p['modules_target_critic'] = O.tree_map(lambda p, tp: p * tau + tp * (dtp(1) - tau), params['modules_critic'],
                                                params['modules_target_critic'])
    return dict(params=new_p, mu=new_m, nu=new_v, count=new_count, step=state['step'] + 1), info, grads


def make_pixel_batch(seed, B, action_dim, hw=64, ch=9, dtype=np.float32):
    rng = np.random.default_rng(seed)
    b = O.make_batch(seed, B, 4, action_dim, dtype)
    b['observations'] = rng.integers(0, 256, (B, hw, hw, ch), dtype=np.uint8)
    b['next_observations'] = rng.integers(0, 256, (B, hw, hw, ch), dtype=np.uint8)
    return b


def min_pool_gap(params, batch):
    """Smallest gap between the two largest entries of any max-pool window of the three gradient-carrying encoder passes on
    batch['observations'] (encoders.py:41 `nn.max_pool`).  The pooling gradient goes to the argmax: with a gap below the rounding
    noise of the arithmetic under test (~1e-7 of the activation scale in fp32) the winner -- and with it the gradient of the
    convolution in front of the pool -- is decided by rounding, not by the algorithm.  Parity tests use this to pick inputs on
    which the comparison with the fp64 oracle is well-posed."""
    gap = np.inf
    for enc in (params['modules_actor_onestep_flow']['encoder'], params['modules_critic']['encoder'], params['modules_actor_bc_flow_encoder']):
        x = batch['observations'].astype(np.float64) / 255.0
        for i in range(len(E.STACKS)):
            blk = enc[f'stack_blocks_{i}']
            c0 = E.conv_fwd(x, blk['Conv_0']['kernel'].astype(np.float64), blk['Conv_0']['bias'].astype(np.float64))
            B, H, W, C = c0.shape
            Ho, Wo = (H + 1) // 2, (W + 1) // 2
            ph, pw = max((Ho - 1) * 2 + 3 - H, 0), max((Wo - 1) * 2 + 3 - W, 0)
            xp = np.pad(c0, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)), constant_values=-np.inf)
            wins = np.sort(np.stack([xp[:, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2, :] for ky in range(3) for kx in range(3)], axis=0), axis=0)
            gap = min(gap, float((wins[-1] - wins[-2]).min()))
            pl = wins[-1]
            c1 = E.conv_fwd(np.maximum(pl, 0), blk['Conv_1']['kernel'].astype(np.float64), blk['Conv_1']['bias'].astype(np.float64))
            x = E.conv_fwd(np.maximum(c1, 0), blk['Conv_2']['kernel'].astype(np.float64), blk['Conv_2']['bias'].astype(np.float64)) + pl
    return gap


def well_posed_pixel_batch(params, B, action_dim, hw, ch, dtype=np.float64, first_seed=4, min_gap=2e-6):
    """make_pixel_batch with the first seed >= first_seed whose pooling windows are all decided by more than `min_gap`."""
    for seed in range(first_seed, first_seed + 64):
        batch = make_pixel_batch(seed, B, action_dim, hw=hw, ch=ch, dtype=dtype)
        if min_pool_gap(params, batch) > min_gap:
            return batch, seed
    raise RuntimeError('no well-posed pixel batch found')
