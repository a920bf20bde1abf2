"""Pins the oracle itself: finite differences (fp64) and an independent torch-autograd transcription.
The reference has no tests/goldens (parity unpinned), so these are what stand behind the oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import fql_oracle as O


def small_cfg(**kw):
    cfg = dict(O.DEFAULT_CONFIG)
    cfg.update(actor_hidden_dims=(24, 24, 24, 24), value_hidden_dims=(24, 24, 24, 24))
    cfg.update(kw)
    return cfg


def setup(cfg, B=6, F=5, A=3, dtype=np.float64, seed=0):
    params = O.init_params(seed, F, A, cfg, dtype=dtype, jitter=0.1, target_equals_critic=False)
    batch = O.make_batch(seed + 1, B, F, A, dtype)
    noise = O.make_noise(seed + 2, B, A, dtype)
    return params, batch, noise


@pytest.mark.parametrize('kw', [
    dict(), dict(q_agg='min', alpha=10.0), dict(normalize_q_loss=True), dict(actor_layer_norm=True),
    dict(layer_norm=False), dict(flow_steps=3, discount=0.995),
])
def test_finite_differences(kw):
    cfg = small_cfg(**kw)
    params, batch, noise = setup(cfg)
    # push some actor outputs outside [-1,1] so the clip gate is exercised
    params['modules_actor_onestep_flow']['mlp']['Dense_4']['bias'] += np.array([1.2, -0.1, 0.0])
    loss, info, grads = O.total_loss(params, cfg, batch, noise)
    rng = np.random.default_rng(5)
    leaves = O.tree_leaves(params)
    gl = dict(O.tree_leaves(grads))
    eps = 1e-6
    for path, arr in leaves:
        g = gl[path]
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in arr.shape)
            old = arr[idx]
            arr[idx] = old + eps
            lp = O.total_loss(params, cfg, batch, noise, with_grads=False)[0]
            arr[idx] = old - eps
            lm = O.total_loss(params, cfg, batch, noise, with_grads=False)[0]
            arr[idx] = old
            fd = (lp - lm) / (2 * eps)
            if path[0] == 'modules_target_critic':
                # stored-params-only module: the true derivative is non-zero but the algorithm stops it
                assert g[idx] == 0.0
                continue
            # stop-gradient structure: stored-param uses of critic/bc_flow/onestep contribute to FD but not to the
            # algorithm's gradient, so compare against a FD that perturbs only the grad_params copy (below)
    # proper check: FD on a loss where stored and grad params are separated
    _fd_separated(cfg, params, batch, noise, grads)


def _fd_separated(cfg, params, batch, noise, grads):
    """FD of L(grad_params; stored_params) w.r.t. grad_params only, emulating `params=grad_params` routing."""
    stored = copy.deepcopy(params)

    def loss_of(gp):
        return _routed_loss(gp, stored, cfg, batch, noise)

    rng = np.random.default_rng(11)
    gl = dict(O.tree_leaves(grads))
    eps = 1e-6
    gp = copy.deepcopy(params)
    for path, arr in O.tree_leaves(gp):
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in arr.shape)
            old = arr[idx]
            arr[idx] = old + eps
            lp = loss_of(gp)
            arr[idx] = old - eps
            lm = loss_of(gp)
            arr[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - gl[path][idx]) <= 1e-6 * max(1.0, abs(fd)) + 1e-8, (path, idx, fd, gl[path][idx])


def _routed_loss(gp, sp, cfg, batch, noise):
    """total_loss with explicit routing: which calls see grad_params (gp) and which see stored params (sp);
    agents/fql.py:25,28,36,58,64,65,70,82."""
    obs, act, nobs = batch['observations'], batch['actions'], batch['next_observations']
    na = np.clip(O.actor_forward(sp['modules_actor_onestep_flow'], cfg, nobs, noise['z_next']), -1, 1)
    nq = O.critic_forward(sp['modules_target_critic'], cfg, nobs, na)
    nq = nq.min(0) if cfg['q_agg'] == 'min' else nq.mean(0)
    tq = batch['rewards'] + cfg['discount'] * batch['masks'] * nq
    q = O.critic_forward(gp['modules_critic'], cfg, obs, act)
    cl = ((q - tq) ** 2).mean()
    x0, t = noise['x0'], noise['t']
    xt = (1 - t) * x0 + t * act
    pred = O.actor_forward(gp['modules_actor_bc_flow'], cfg, obs, xt, t)
    bc = ((pred - (act - x0)) ** 2).mean()
    tgt = O.compute_flow_actions(sp, cfg, obs, noise['z'])
    api = O.actor_forward(gp['modules_actor_onestep_flow'], cfg, obs, noise['z'])
    dl = ((api - tgt) ** 2).mean()
    qs = O.critic_forward(sp['modules_critic'], cfg, obs, np.clip(api, -1, 1))
    qm = qs.mean(0)
    ql = -qm.mean()
    if cfg['normalize_q_loss']:
        lam = 1 / np.abs(O.critic_forward(sp['modules_critic'], cfg, obs, np.clip(
            O.actor_forward(sp['modules_actor_onestep_flow'], cfg, obs, noise['z']), -1, 1)).mean(0)).mean()
        ql = lam * ql
    return cl + bc + cfg['alpha'] * dl + ql


# ---------------- independent torch-autograd transcription ----------------
def _t_mlp(p, x, ln):
    n = O.n_dense(p)
    for i in range(n):
        W, b = p[f'Dense_{i}']['kernel'], p[f'Dense_{i}']['bias']
        if W.ndim == 3:
            x = torch.matmul(x, W) + b[:, None, :]
        else:
            x = x @ W + b
        if i + 1 < n:
            x = torch.nn.functional.gelu(x, approximate='tanh')
            if ln:
                sc, bi = p[f'LayerNorm_{i}']['scale'], p[f'LayerNorm_{i}']['bias']
                mu = x.mean(-1, keepdim=True)
                var = (x * x).mean(-1, keepdim=True) - mu * mu
                xh = (x - mu) / torch.sqrt(var.clamp_min(0) + 1e-6)
                x = xh * (sc[:, None, :] if W.ndim == 3 else sc) + (bi[:, None, :] if W.ndim == 3 else bi)
    return x


def _to_torch(t, grad):
    if isinstance(t, dict):
        return {k: _to_torch(v, grad) for k, v in t.items()}
    return torch.tensor(t, dtype=torch.float64, requires_grad=grad)


@pytest.mark.parametrize('kw', [dict(), dict(q_agg='min', alpha=10.0), dict(normalize_q_loss=True, alpha=1000.0),
                                dict(actor_layer_norm=True)])
def test_against_torch_autograd(kw):
    cfg = small_cfg(**kw)
    params, batch, noise = setup(cfg, B=9, F=7, A=4, seed=3)
    loss, info, grads = O.total_loss(params, cfg, batch, noise)
    gp = _to_torch(params, True)
    sp = _to_torch(params, False)
    b = {k: torch.tensor(v) for k, v in batch.items()}
    nz = {k: torch.tensor(v) for k, v in noise.items()}
    cat = lambda *a: torch.cat(a, -1)
    aln, cln = cfg['actor_layer_norm'], cfg['layer_norm']
    obs, act, nobs = b['observations'], b['actions'], b['next_observations']
    na = _t_mlp(sp['modules_actor_onestep_flow']['mlp'], cat(nobs, nz['z_next']), aln).clamp(-1, 1)
    nq = _t_mlp(sp['modules_target_critic']['value_net'], cat(nobs, na), cln)[..., 0]
    nq = nq.min(0).values if cfg['q_agg'] == 'min' else nq.mean(0)
    tq = b['rewards'] + cfg['discount'] * b['masks'] * nq
    q = _t_mlp(gp['modules_critic']['value_net'], cat(obs, act), cln)[..., 0]
    cl = ((q - tq) ** 2).mean()
    x0, t = nz['x0'], nz['t']
    pred = _t_mlp(gp['modules_actor_bc_flow']['mlp'], cat(obs, (1 - t) * x0 + t * act, t), aln)
    bc = ((pred - (act - x0)) ** 2).mean()
    a = nz['z']
    for i in range(cfg['flow_steps']):
        tt = torch.full((obs.shape[0], 1), i / cfg['flow_steps'], dtype=torch.float64)
        a = a + _t_mlp(sp['modules_actor_bc_flow']['mlp'], cat(obs, a, tt), aln) / cfg['flow_steps']
    tgt = a.clamp(-1, 1)
    api = _t_mlp(gp['modules_actor_onestep_flow']['mlp'], cat(obs, nz['z']), aln)
    dl = ((api - tgt) ** 2).mean()
    qm = _t_mlp(sp['modules_critic']['value_net'], cat(obs, api.clamp(-1, 1)), cln)[..., 0].mean(0)
    ql = -qm.mean()
    if cfg['normalize_q_loss']:
        ql = ql * (1 / qm.abs().mean()).detach()
    total = cl + bc + cfg['alpha'] * dl + ql
    total.backward()
    assert abs(total.item() - loss) < 1e-10 * max(1, abs(loss))
    np.testing.assert_allclose(info['critic/critic_loss'], cl.item(), rtol=1e-12)
    np.testing.assert_allclose(info['actor/distill_loss'], dl.item(), rtol=1e-12)
    np.testing.assert_allclose(info['actor/q_loss'], ql.item(), rtol=1e-12)
    gl = dict(O.tree_leaves(grads))
    for path, tg in O.tree_leaves(gp):
        ref = tg.grad.numpy() if tg.grad is not None else np.zeros(tg.shape)
        np.testing.assert_allclose(gl[path], ref, rtol=1e-9, atol=1e-12, err_msg=str(path))


def test_adam_polyak_against_torch():
    cfg = small_cfg()
    params, batch, noise = setup(cfg, seed=4)
    state = O.init_state(params, warm=True)
    new_state, info, grads = O.update(copy.deepcopy(state), cfg, batch, noise)
    # torch.optim.Adam is the same arithmetic as optax.adam (eps outside the sqrt, bias-corrected)
    for net in ('modules_critic', 'modules_actor_bc_flow'):
        for path, p in O.tree_leaves(state['params'][net]):
            g = dict(O.tree_leaves(grads[net]))[path]
            tp = torch.tensor(p.copy(), requires_grad=True)
            opt = torch.optim.Adam([tp], lr=cfg['lr'], betas=(0.9, 0.999), eps=1e-8)
            tp.grad = torch.tensor(g)
            opt.state[tp] = dict(step=torch.tensor(float(state['count'])),
                                 exp_avg=torch.tensor(dict(O.tree_leaves(state['mu'][net]))[path].copy()),
                                 exp_avg_sq=torch.tensor(dict(O.tree_leaves(state['nu'][net]))[path].copy()))
            opt.step()
            # the oracle evaluates the bias correction in float32 like optax under jax's default dtype (1.3e-5 away from
            # the double value at small counts), torch evaluates it in double: agreement of the new parameters to ~1e-8
            np.testing.assert_allclose(dict(O.tree_leaves(new_state['params'][net]))[path], tp.detach().numpy(),
                                       rtol=1e-6, atol=2e-8)
    # Polyak uses pre-step critic (F6); target Adam step is a no-op (F7)
    for path, tp_old in O.tree_leaves(state['params']['modules_target_critic']):
        p_old = dict(O.tree_leaves(state['params']['modules_critic']))[path]
        np.testing.assert_allclose(dict(O.tree_leaves(new_state['params']['modules_target_critic']))[path],
                                   cfg['tau'] * p_old + (1 - cfg['tau']) * tp_old, rtol=1e-15)
    assert new_state['count'] == state['count'] + 1 and new_state['step'] == state['step'] + 1
    # grad stats: L1 norm of per-leaf L2 norms, zeros of the target critic included
    leaves = [g for _, g in O.tree_leaves(grads)]
    assert np.isclose(info['grad/norm'], sum(np.linalg.norm(l.ravel()) for l in leaves))
    assert info['grad/max'] == max(l.max() for l in leaves) and info['grad/min'] == min(l.min() for l in leaves)


def test_fp32_vs_fp64_gap():
    """Calibrates the 1e-5 tolerance: the fp32 restatement sits ~1e-6 from fp64 on tensor-norm-relative error."""
    cfg = dict(O.DEFAULT_CONFIG)
    cfg.update(actor_hidden_dims=(128,) * 4, value_hidden_dims=(128,) * 4)
    params, batch, noise = setup(cfg, B=64, F=29, A=8, seed=9)
    l64, i64, g64 = O.total_loss(params, cfg, batch, noise)
    c = lambda t: O.cast_tree(t, np.float32)
    l32, i32, g32 = O.total_loss(c(params), cfg, c(batch), c(noise))
    assert abs(l32 - l64) / abs(l64) < 1e-5
    for (path, a), (_, b) in zip(O.tree_leaves(g64), O.tree_leaves(g32)):
        if path[0] == 'modules_target_critic':
            continue
        assert np.abs(a - b).max() / np.abs(a).max() < 2e-5, path


def test_torch_cpu_restatement_matches_numpy_oracle():
    """oracle/fql_torch_cpu.py (the timed CPU baseline) is the same algorithm as oracle/fql_oracle.py: one whole update in
    fp64 from identical state -> identical info, grads, params, mu, nu."""
    from oracle.fql_torch_cpu import TorchCpuAgent
    for kw in (dict(q_agg='min', alpha=10.0), dict(normalize_q_loss=True, alpha=1000.0)):
        cfg = small_cfg(**kw)
        params, batch, noise = setup(cfg, B=12, F=6, A=3, seed=6)
        state = O.init_state(params, warm=True)
        ta = TorchCpuAgent(state['params'], cfg, state['mu'], state['nu'], state['count'], dtype=torch.float64)
        new_state, info, grads = O.update(copy.deepcopy(state), cfg, batch, noise)
        tinfo, tgrads = ta.update(batch, noise)
        for k in O.INFO_KEYS:
            np.testing.assert_allclose(tinfo[k], float(info[k]), rtol=1e-9, atol=1e-12, err_msg=k)
        for which, ref in (('p', new_state['params']), ('m', new_state['mu']), ('v', new_state['nu'])):
            for (path, r), (_, g) in zip(O.tree_leaves(ref), O.tree_leaves(ta.tree(which))):
                np.testing.assert_allclose(g, r, rtol=1e-9, atol=1e-13, err_msg=f'{which} {path}')
        for (path, r), (_, g) in zip(O.tree_leaves(grads), O.tree_leaves(tgrads)):
            np.testing.assert_allclose(g, r, rtol=1e-8, atol=1e-12, err_msg=str(path))


def test_encoder_oracle_against_torch_autograd():
    """oracle/encoder_oracle.py (impala_small restatement, manual backward) vs torch conv2d / max_pool2d autograd in fp64."""
    import torch.nn.functional as Fn
    from oracle import encoder_oracle as E
    rng = np.random.default_rng(3)
    p = E.init_encoder(rng, 6, dtype=np.float64, hw=16, jitter=0.1)
    obs = rng.integers(0, 256, (3, 16, 16, 6), dtype=np.uint8)
    out, saved = E.encoder_forward(p, obs, save=True)
    dout = rng.standard_normal(out.shape)
    g = E.encoder_backward(p, saved, dout)
    tp = _to_torch(p, True)
    x = torch.tensor(obs.astype(np.float64) / 255.0).permute(0, 3, 1, 2)
    conv = lambda x, q: Fn.conv2d(x, q['kernel'].permute(3, 2, 0, 1), q['bias'], padding=1)
    for i in range(3):
        b = tp[f'stack_blocks_{i}']
        x = conv(x, b['Conv_0'])
        x = Fn.max_pool2d(Fn.pad(x, (0, 1, 0, 1), value=float('-inf')), 3, 2)
        y = conv(torch.relu(x), b['Conv_1'])
        y = conv(torch.relu(y), b['Conv_2'])
        x = y + x
    flat = torch.relu(x).permute(0, 2, 3, 1).reshape(x.shape[0], -1)          # NHWC flatten (encoders.py:96)
    tout = Fn.gelu(flat @ tp['MLP_0']['Dense_0']['kernel'] + tp['MLP_0']['Dense_0']['bias'], approximate='tanh')
    np.testing.assert_allclose(out, tout.detach().numpy(), rtol=1e-10, atol=1e-12)
    (tout * torch.tensor(dout)).sum().backward()
    for (path, a), (_, t) in zip(O.tree_leaves(g), O.tree_leaves(tp)):
        np.testing.assert_allclose(a, t.grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=str(path))
    # 64x64x9 (frame_stack 3) geometry of BASELINE config 5: 64 -> 32 -> 16 -> 8, flatten 8*8*32 = 2048, 1,105,920 parameters
    p64 = E.init_encoder(rng, 9, hw=64)
    assert p64['MLP_0']['Dense_0']['kernel'].shape == (2048, 512)
    assert sum(v.size for _, v in O.tree_leaves(p64)) == 1105920


def test_pixel_oracle_finite_differences_and_state_oracle_unchanged():
    """FD check of the pixel-config total gradient (encoders included) with separated grad/stored params, fp64, 16x16 images."""
    from oracle import fql_pixel_oracle as PO
    cfg = small_cfg(alpha=10.0, encoder='impala_small')
    B, A = 3, 2
    params = PO.init_params(0, 6, A, cfg, dtype=np.float64, hw=16, jitter=0.1, target_equals_critic=False)
    batch = PO.make_pixel_batch(1, B, A, hw=16, ch=6, dtype=np.float64)
    noise = O.make_noise(2, B, A, np.float64)
    loss, info, grads = PO.total_loss(params, cfg, batch, noise)
    stored = copy.deepcopy(params)

    def routed(gp):  # only the three gradient-carrying call sites see gp (fql.py:36,58,65); everything else sees stored
        mix = copy.deepcopy(stored)
        fC = PO.E.encoder_forward(gp['modules_critic']['encoder'], batch['observations'])
        fF = PO.E.encoder_forward(gp['modules_actor_bc_flow_encoder'], batch['observations'])
        fO = PO.E.encoder_forward(gp['modules_actor_onestep_flow']['encoder'], batch['observations'])
        sfC = PO.E.encoder_forward(stored['modules_critic']['encoder'], batch['observations'])
        sfF = PO.E.encoder_forward(stored['modules_actor_bc_flow_encoder'], batch['observations'])
        sfO = PO.E.encoder_forward(stored['modules_actor_onestep_flow']['encoder'], batch['observations'])
        fOn = PO.E.encoder_forward(stored['modules_actor_onestep_flow']['encoder'], batch['next_observations'])
        fTn = PO.E.encoder_forward(stored['modules_target_critic']['encoder'], batch['next_observations'])
        b = dict(batch)
        na = np.clip(O.actor_forward(stored['modules_actor_onestep_flow'], cfg, fOn, noise['z_next']), -1, 1)
        tq = b['rewards'] + cfg['discount'] * b['masks'] * O.critic_forward(stored['modules_target_critic'], cfg, fTn, na).mean(0)
        q = O.critic_forward(gp['modules_critic'], cfg, fC, b['actions'])
        cl = ((q - tq) ** 2).mean()
        x0, t = noise['x0'], noise['t']
        pred = O.actor_forward(gp['modules_actor_bc_flow'], cfg, fF, (1 - t) * x0 + t * b['actions'], t)
        bc = ((pred - (b['actions'] - x0)) ** 2).mean()
        tgt = O.compute_flow_actions(stored, cfg, sfF, noise['z'])
        api = O.actor_forward(gp['modules_actor_onestep_flow'], cfg, fO, noise['z'])
        dl = ((api - tgt) ** 2).mean()
        ql = -O.critic_forward(stored['modules_critic'], cfg, sfC, np.clip(api, -1, 1)).mean(0).mean()
        return cl + bc + cfg['alpha'] * dl + ql

    assert abs(routed(copy.deepcopy(params)) - loss) < 1e-10
    rng = np.random.default_rng(4)
    gl = dict(O.tree_leaves(grads))
    gp = copy.deepcopy(params)
    eps = 1e-6
    for path, arr in O.tree_leaves(gp):
        if path[0] == 'modules_target_critic':
            assert not gl[path].any()
            continue
        for _ in range(2):
            idx = tuple(rng.integers(0, s) for s in arr.shape)
            old = arr[idx]
            arr[idx] = old + eps
            lp = routed(gp)
            arr[idx] = old - eps
            lm = routed(gp)
            arr[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - gl[path][idx]) <= 2e-6 * max(1.0, abs(fd)) + 1e-8, (path, idx, fd, gl[path][idx])


def test_torch_cpu_restatement_matches_pixel_oracle():
    """The timed CPU baseline with encoders == the pixel oracle (one whole update, fp64, 16x16 images)."""
    from oracle import fql_pixel_oracle as PO
    from oracle.fql_torch_cpu import TorchCpuAgent
    cfg = small_cfg(alpha=10.0, encoder='impala_small', q_agg='min')
    B, A = 4, 3
    params = PO.init_params(0, 6, A, cfg, dtype=np.float64, hw=16, jitter=0.1, target_equals_critic=False)
    state = O.init_state(params, warm=True)
    batch = PO.make_pixel_batch(1, B, A, hw=16, ch=6, dtype=np.float64)
    noise = O.make_noise(2, B, A, np.float64)
    new_state, info, grads = PO.update(copy.deepcopy(state), cfg, batch, noise)
    ta = TorchCpuAgent(state['params'], cfg, state['mu'], state['nu'], state['count'], dtype=torch.float64)
    tinfo, tgrads = ta.update(batch, noise)
    for k in O.INFO_KEYS:
        np.testing.assert_allclose(tinfo[k], float(info[k]), rtol=1e-8, atol=1e-11, err_msg=k)
    for (path, r), (_, g) in zip(O.tree_leaves(grads), O.tree_leaves(tgrads)):
        np.testing.assert_allclose(g, r, rtol=1e-7, atol=1e-11, err_msg=str(path))
    for (path, r), (_, g) in zip(O.tree_leaves(new_state['params']), O.tree_leaves(ta.tree('p'))):
        np.testing.assert_allclose(g, r, rtol=1e-8, atol=1e-12, err_msg=str(path))


def test_bf16_storage_rounding_hook_of_the_encoder_oracle():
    """oracle/encoder_oracle.py `q=bf16_round` (the reference for FQL_PRECISION_BF16_ENC / BF16_TC on pixel configs): the rounding is
    torch's round-to-nearest-even bfloat16 cast, the rounded forward stays within bf16 distance of the exact one, and the rounded
    backward is a gradient of about the same function (cosine with the exact gradient)."""
    import torch
    from oracle import encoder_oracle as E
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4096) * np.exp(rng.uniform(-20, 20, 4096))
    assert np.array_equal(E.bf16_round(x), torch.tensor(x, dtype=torch.float32).to(torch.bfloat16).double().numpy())
    p = E.init_encoder(rng, 6, np.float64, hw=16, jitter=0.05)
    obs = rng.integers(0, 256, (4, 16, 16, 6), dtype=np.uint8)
    f0, s0 = E.encoder_forward(p, obs, dtype=np.dtype(np.float64), save=True)
    fq, sq = E.encoder_forward(p, obs, dtype=np.dtype(np.float64), save=True, q=E.bf16_round)
    assert np.abs(fq - f0).max() <= 2e-2 * np.abs(f0).max()
    dout = rng.standard_normal(f0.shape)
    g0, gq = E.encoder_backward(p, s0, dout), E.encoder_backward(p, sq, dout, q=E.bf16_round)
    a, b = g0['MLP_0']['Dense_0']['kernel'].ravel(), gq['MLP_0']['Dense_0']['kernel'].ravel()
    assert np.vdot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)) >= 0.999
    a, b = g0['stack_blocks_0']['Conv_0']['kernel'].ravel(), gq['stack_blocks_0']['Conv_0']['kernel'].ravel()
    assert np.vdot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)) >= 0.9


def test_pixel_batch_conditioning_helper():
    """The max-pool gradient goes to the argmax (encoders.py:41): a window whose two largest entries differ by less than the rounding
    noise of the arithmetic under test is decided by rounding.  `min_pool_gap` measures the closest call of a batch; seed 4 at 8 frames
    of 16x16x6 is the near-tie (1.3e-8) that the 2-GPU fp32 pixel test ran into, and `well_posed_pixel_batch` steps past it."""
    from oracle import fql_pixel_oracle as PO
    hw, ch, A, hidden = 16, 6, 3, 32
    cfg = dict(O.DEFAULT_CONFIG)
    cfg.update(alpha=10.0, actor_hidden_dims=(hidden,) * 4, value_hidden_dims=(hidden,) * 4, encoder='impala_small')
    params = PO.init_params(3, ch, A, cfg, dtype=np.float64, hw=hw, jitter=0.05, target_equals_critic=False)
    tie = PO.make_pixel_batch(4, 8, A, hw=hw, ch=ch, dtype=np.float64)
    assert PO.min_pool_gap(params, tie) < 1e-7
    ok = PO.make_pixel_batch(4, 6, A, hw=hw, ch=ch, dtype=np.float64)
    assert PO.min_pool_gap(params, ok) > 2e-6
    batch, seed = PO.well_posed_pixel_batch(params, 8, A, hw, ch, first_seed=4)
    assert seed > 4 and PO.min_pool_gap(params, batch) > 2e-6
