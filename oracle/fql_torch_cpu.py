"""torch-CPU fp32 restatement of FQLAgent.update -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/fql_oracle.py header).

Same algorithm and citations as oracle/fql_oracle.py, written with torch autograd so that it (a) is an independent check of
the NumPy oracle's manual backward (tests/test_oracle.py) and (b) uses every host core for both the matmuls (MKL) and the
elementwise work, which makes it the fairest stand-in available here for "the reference's JAX CPU path" (jax/flax/optax
are not installed).  Used by bench.py's `cpu_baseline` / `--impl reference` legs.  kind = "port".
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as Fn


def to_torch(tree, dtype=torch.float32):
    if isinstance(tree, dict):
        return {k: to_torch(v, dtype) for k, v in tree.items()}
    return torch.tensor(np.asarray(tree), dtype=dtype)


def leaves(tree, prefix=()):
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out += leaves(tree[k], prefix + (k,))
        return out
    return [(prefix, tree)]


def mlp(p, x, ln):
    """utils/networks.py:34-61 (Dense -> gelu(tanh) -> LayerNorm eps 1e-6, fast variance); ensemble via broadcasting."""
    n = sum(1 for k in p if k.startswith('Dense_'))
    for i in range(n):
        W, b = p[f'Dense_{i}']['kernel'], p[f'Dense_{i}']['bias']
        x = torch.matmul(x, W) + (b[:, None, :] if W.dim() == 3 else b)
        if i + 1 < n:
            x = Fn.gelu(x, approximate='tanh')
            if ln:
                sc, bi = p[f'LayerNorm_{i}']['scale'], p[f'LayerNorm_{i}']['bias']
                mu = x.mean(-1, keepdim=True)
                var = ((x * x).mean(-1, keepdim=True) - mu * mu).clamp_min(0)
                x = (x - mu) * torch.rsqrt(var + 1e-6)
                x = x * (sc[:, None, :] if W.dim() == 3 else sc) + (bi[:, None, :] if W.dim() == 3 else bi)
    return x


def encoder(p, obs_u8):
    """ImpalaEncoder('impala_small') (utils/encoders.py:60-100) with torch ops; obs uint8 [B,H,W,C] -> [B,512]."""
    x = obs_u8.to(p['MLP_0']['Dense_0']['kernel'].dtype).permute(0, 3, 1, 2) / 255.0
    conv = lambda x, q: Fn.conv2d(x, q['kernel'].permute(3, 2, 0, 1), q['bias'], padding=1)
    for i in range(3):
        blk = p[f'stack_blocks_{i}']
        x = conv(x, blk['Conv_0'])
        ph, pw = x.shape[2] % 2, x.shape[3] % 2   # SAME 3x3/2: pad (0,1) for even sizes, (1,1) for odd
        x = Fn.max_pool2d(Fn.pad(x, (pw, 1, ph, 1), value=float('-inf')), 3, 2)
        y = conv(torch.relu(x), blk['Conv_1'])
        y = conv(torch.relu(y), blk['Conv_2'])
        x = y + x
    flat = torch.relu(x).permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    return Fn.gelu(flat @ p['MLP_0']['Dense_0']['kernel'] + p['MLP_0']['Dense_0']['bias'], approximate='tanh')


def total_loss(gp, sp, cfg, b, nz):
    """agents/fql.py:22-111.  gp: params that receive gradients (`params=grad_params`), sp: stored params (stop-gradient)."""
    cat = lambda *a: torch.cat(a, -1)
    aln, cln = cfg['actor_layer_norm'], cfg['layer_norm']
    obs, act, nobs = b['observations'], b['actions'], b['next_observations']
    info = {}
    pix = cfg.get('encoder') is not None
    if pix:  # agents/fql.py with encoders: per call site features (see oracle/fql_pixel_oracle.py for the routing)
        with torch.no_grad():
            nobs_O = encoder(sp['modules_actor_onestep_flow']['encoder'], nobs)
            nobs_T = encoder(sp['modules_target_critic']['encoder'], nobs)
            s_fC = encoder(sp['modules_critic']['encoder'], obs)
            s_fF = encoder(sp['modules_actor_bc_flow_encoder'], obs)
            s_fO = encoder(sp['modules_actor_onestep_flow']['encoder'], obs)
        g_fC = encoder(gp['modules_critic']['encoder'], obs)
        g_fF = encoder(gp['modules_actor_bc_flow_encoder'], obs)
        g_fO = encoder(gp['modules_actor_onestep_flow']['encoder'], obs)
    else:
        nobs_O = nobs_T = nobs
        s_fC = s_fF = s_fO = g_fC = g_fF = g_fO = obs
    with torch.no_grad():
        na = mlp(sp['modules_actor_onestep_flow']['mlp'], cat(nobs_O, nz['z_next']), aln).clamp(-1, 1)
        nq = mlp(sp['modules_target_critic']['value_net'], cat(nobs_T, na), cln)[..., 0]
        nq = nq.min(0).values if cfg['q_agg'] == 'min' else nq.mean(0)
        tq = b['rewards'] + cfg['discount'] * b['masks'] * nq
    q = mlp(gp['modules_critic']['value_net'], cat(g_fC, act), cln)[..., 0]
    cl = ((q - tq) ** 2).mean()
    info.update({'critic/critic_loss': cl, 'critic/q_mean': q.mean(), 'critic/q_max': q.max(), 'critic/q_min': q.min()})
    x0, t = nz['x0'], nz['t']
    pred = mlp(gp['modules_actor_bc_flow']['mlp'], cat(g_fF, (1 - t) * x0 + t * act, t), aln)
    bc = ((pred - (act - x0)) ** 2).mean()
    with torch.no_grad():
        a = nz['z']
        n = cfg['flow_steps']
        for i in range(n):
            tt = torch.full((act.shape[0], 1), i / n, dtype=act.dtype)
            a = a + mlp(sp['modules_actor_bc_flow']['mlp'], cat(s_fF, a, tt), aln) / n
        tgt = a.clamp(-1, 1)
    api = mlp(gp['modules_actor_onestep_flow']['mlp'], cat(g_fO, nz['z']), aln)
    dl = ((api - tgt) ** 2).mean()
    qm = mlp(sp['modules_critic']['value_net'], cat(s_fC, api.clamp(-1, 1)), cln)[..., 0].mean(0)
    ql = -qm.mean()
    if cfg['normalize_q_loss']:
        ql = ql * (1 / qm.abs().mean()).detach()
    al = bc + cfg['alpha'] * dl + ql
    with torch.no_grad():
        ma = mlp(sp['modules_actor_onestep_flow']['mlp'], cat(s_fO, nz['z_metric']), aln).clamp(-1, 1)
        mse = ((ma - act) ** 2).mean()
    info.update({'actor/actor_loss': al, 'actor/bc_flow_loss': bc, 'actor/distill_loss': dl, 'actor/q_loss': ql,
                 'actor/q': qm.mean(), 'actor/mse': mse})
    return cl + al, info


class TorchCpuAgent:
    """State held as torch CPU tensors; update() = agents/fql.py:122-133 + utils/flax_utils.py:120-159."""

    def __init__(self, params, cfg, mu=None, nu=None, count=0, dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.p = [t for _, t in leaves(to_torch(params, dtype))]
        self.paths = [pth for pth, _ in leaves(params)]
        self.m = [t for _, t in leaves(to_torch(mu, dtype))] if mu is not None else [torch.zeros_like(t) for t in self.p]
        self.v = [t for _, t in leaves(to_torch(nu, dtype))] if nu is not None else [torch.zeros_like(t) for t in self.p]
        self.count = count

    def _tree(self, flat):
        out = {}
        for pth, t in zip(self.paths, flat):
            d = out
            for k in pth[:-1]:
                d = d.setdefault(k, {})
            d[pth[-1]] = t
        return out

    def tree(self, which='p'):
        return self._tree([t.detach().numpy() for t in getattr(self, which)])

    def update(self, batch, noise):
        cfg = self.cfg
        b = {k: (torch.as_tensor(v) if np.asarray(v).dtype == np.uint8 else torch.as_tensor(v, dtype=self.dtype)) for k, v in batch.items()}
        nz = {k: torch.as_tensor(v, dtype=self.dtype) for k, v in noise.items()}
        gp_flat = [t.detach().requires_grad_(True) for t in self.p]
        loss, info = total_loss(self._tree(gp_flat), self._tree(self.p), cfg, b, nz)
        grads = torch.autograd.grad(loss, gp_flat, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, self.p)]
        info['grad/max'] = max(g.max() for g in grads)
        info['grad/min'] = min(g.min() for g in grads)
        info['grad/norm'] = sum(torch.linalg.vector_norm(g) for g in grads)
        t = self.count + 1
        bc1 = float(np.float32(1) - np.power(np.float32(0.9), np.float32(t)))     # optax: float32 bias correction
        bc2 = float(np.float32(1) - np.power(np.float32(0.999), np.float32(t)))
        tau = cfg['tau']
        new_p = []
        with torch.no_grad():
            for i, (pth, p, g) in enumerate(zip(self.paths, self.p, grads)):
                if pth[0] == 'modules_target_critic':
                    src = self.p[self.paths.index(('modules_critic',) + pth[1:])]  # incl. the encoder (deepcopy of the critic def)
                    new_p.append(src * tau + p * (1 - tau))                         # pre-step critic (fql.py:113-120)
                    continue
                self.m[i] = 0.9 * self.m[i] + (1 - 0.9) * g
                self.v[i] = 0.999 * self.v[i] + (1 - 0.999) * g * g
                new_p.append(p - cfg['lr'] * ((self.m[i] / bc1) / (torch.sqrt(self.v[i] / bc2) + 1e-8)))
        self.p = new_p
        self.count = t
        return {k: float(v.detach()) for k, v in info.items()}, self._tree([g.numpy() for g in grads])
