"""CPU oracle for the visual encoder -- TEST INFRASTRUCTURE ONLY (see oracle/fql_oracle.py header; parity unpinned).

NumPy restatement of `ImpalaEncoder('impala_small')` = ImpalaEncoder(num_blocks=1, stack_sizes=(16,32,32), mlp_hidden_dims=(512,)):
  utils/encoders.py:83-100  ImpalaEncoder.__call__: x = u8/255 -> 3 x ResnetStack -> relu -> flatten (NHWC) -> MLP((512,), activate_final=True)
  utils/encoders.py:17-57   ResnetStack: conv3x3 SAME -> max_pool 3x3 stride 2 SAME -> [relu -> conv -> relu -> conv] + skip
  utils/networks.py:34-61   MLP with activate_final: Dense -> gelu(tanh) (no LayerNorm: layer_norm=False)
Third-party semantics encoded (flax, unverifiable here): nn.Conv 'SAME' stride 1 = zero pad 1, kernel HWIO [3,3,cin,cout] + bias;
nn.max_pool 3x3/2 'SAME' pads with -inf, for even inputs the pad is (0,1): window i covers input rows [2i, 2i+2]; the gradient goes
to the first maximum of a window in row-major order (select_and_scatter).
Parameter names follow Flax: stack_blocks_{i}/Conv_{j}/{kernel,bias}, MLP_0/Dense_0/{kernel,bias}.
"""
from __future__ import annotations

import math

import numpy as np

from oracle.fql_oracle import gelu_tanh, gelu_tanh_grad

STACKS = (16, 32, 32)


def init_encoder(rng, in_ch, dtype=np.float32, hw=64, jitter=0.0):
    """xavier_uniform convs (encoders.py:19), default_init Dense (networks.py:9-11), zero biases (+ jitter for tests)."""
    p = {}
    c = in_ch
    for i, f in enumerate(STACKS):
        blk = {}
        for j, (ci, co) in enumerate([(c, f), (f, f), (f, f)]):
            fan_in, fan_out = 9 * ci, 9 * co
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            blk[f'Conv_{j}'] = {'kernel': rng.uniform(-lim, lim, (3, 3, ci, co)).astype(dtype),
                                'bias': (jitter * rng.standard_normal(co)).astype(dtype)}
        p[f'stack_blocks_{i}'] = blk
        c = f
        hw = (hw + 1) // 2
    flat = hw * hw * c
    lim = math.sqrt(6.0 / (flat + 512))
    p['MLP_0'] = {'Dense_0': {'kernel': rng.uniform(-lim, lim, (flat, 512)).astype(dtype), 'bias': (jitter * rng.standard_normal(512)).astype(dtype)}}
    return p


def _im2col(x):
    """x [B,H,W,C] -> [B,H,W,9*C] patches of the zero-padded 3x3 neighbourhood, (ky,kx,c) order."""
    B, H, W, C = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    cols = [xp[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)]
    return np.concatenate(cols, axis=-1)


def conv_fwd(x, W, b):
    cols = _im2col(x)
    return cols @ W.reshape(-1, W.shape[-1]) + b


def conv_bwd(x, W, dy):
    """-> dx, dW, db"""
    B, H, Wd, C = x.shape
    cols = _im2col(x)
    dW = (cols.reshape(-1, cols.shape[-1]).T @ dy.reshape(-1, dy.shape[-1])).reshape(W.shape)
    db = dy.reshape(-1, dy.shape[-1]).sum(0)
    dcols = dy @ W.reshape(-1, W.shape[-1]).T                       # [B,H,W,9C]
    dxp = np.zeros((B, H + 2, Wd + 2, C), x.dtype)
    k = 0
    for ky in range(3):
        for kx in range(3):
            dxp[:, ky:ky + H, kx:kx + Wd, :] += dcols[..., k * C:(k + 1) * C]
            k += 1
    return dxp[:, 1:-1, 1:-1, :], dW, db


def pool_fwd(x):
    """max_pool 3x3 stride 2 SAME (-inf padding, pad (0,1) for even sizes). Returns y and the argmax (0..8, row-major, first max)."""
    B, H, W, C = x.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    ph, pw = max((Ho - 1) * 2 + 3 - H, 0), max((Wo - 1) * 2 + 3 - W, 0)
    xp = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)), constant_values=-np.inf)
    wins = np.stack([xp[:, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2, :] for ky in range(3) for kx in range(3)], axis=0)  # [9,B,Ho,Wo,C]
    arg = wins.argmax(axis=0)          # first maximum in (ky,kx) row-major order
    return wins.max(axis=0), arg, (ph // 2, pw // 2)


def pool_bwd(dy, arg, in_shape, pad_lo):
    B, H, W, C = in_shape
    Ho, Wo = dy.shape[1:3]
    dxp = np.zeros((B, H + 3, W + 3, C), dy.dtype)
    for k in range(9):
        ky, kx = divmod(k, 3)
        dxp[:, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2, :] += np.where(arg == k, dy, 0)
    return dxp[:, pad_lo[0]:pad_lo[0] + H, pad_lo[1]:pad_lo[1] + W, :]


def bf16_round(x):
    """round-to-nearest-even to bfloat16 (returned in x's dtype): the storage rounding of the tensor-core encoder path"""
    x = np.asarray(x)
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).view(np.float32)
    return r.astype(x.dtype)


def encoder_forward(p, obs_u8, dtype=None, save=False, q=None):
    """obs_u8 [B,H,W,C] uint8 -> features [B,512].  With save=True also returns the cache for encoder_backward.
    q (optional, e.g. bf16_round): the same mathematics with every STORED activation and every matrix-product weight operand rounded
    by q -- the reference's arithmetic as FQL_PRECISION_BF16_ENC evaluates it (fql_b200/csrc/encoder_tc.cu: bf16 NHWC activations,
    bf16 weight operands, fp32 accumulation, fp32 biases; the 0..255 frames are exact in bf16 and 1/255 scales the accumulator)."""
    dt = dtype or p['MLP_0']['Dense_0']['kernel'].dtype
    q = q or (lambda a: a)
    x = obs_u8.astype(dt) / dt.type(255.0)
    cache = []
    for i in range(len(STACKS)):
        blk = p[f'stack_blocks_{i}']
        x_in = x
        c0 = q(conv_fwd(x_in, q(blk['Conv_0']['kernel']), blk['Conv_0']['bias']))
        pl, arg, pad_lo = pool_fwd(c0)
        r1 = np.maximum(pl, 0)
        c1 = conv_fwd(r1, q(blk['Conv_1']['kernel']), blk['Conv_1']['bias'])
        r2 = q(np.maximum(c1, 0))
        c2 = conv_fwd(r2, q(blk['Conv_2']['kernel']), blk['Conv_2']['bias'])
        x = q(c2 + pl)
        cache.append((x_in, c0.shape, arg, pad_lo, pl, r1, c1, r2))
    xr = np.maximum(x, 0)
    flat = xr.reshape(xr.shape[0], -1)
    d = p['MLP_0']['Dense_0']
    z = flat @ q(d['kernel']) + d['bias']
    out = gelu_tanh(z)
    if save:
        return out, (cache, x, flat, z)
    return out


def encoder_backward(p, saved, dout, q=None):
    """-> grads with the layout of p (no input gradient: the input is pixels).  q: see encoder_forward (gradient tensors that the
    tensor-core path stores -- dz as a GEMM operand, dx, dc1, dpool, dc0 -- are rounded by q; parameter gradients are fp32 sums)."""
    cache, x_last, flat, z = saved
    q = q or (lambda a: a)
    g = {}
    d = p['MLP_0']['Dense_0']
    dz = dout * gelu_tanh_grad(z)
    dzq = q(dz)
    g['MLP_0'] = {'Dense_0': {'kernel': flat.T @ dzq, 'bias': dz.sum(0)}}
    dx = q((dzq @ q(d['kernel']).T).reshape(x_last.shape) * (x_last > 0))
    for i in reversed(range(len(STACKS))):
        blk = p[f'stack_blocks_{i}']
        x_in, c0_shape, arg, pad_lo, pl, r1, c1, r2 = cache[i]
        gb = {}
        dr2, dW2, db2 = conv_bwd(r2, q(blk['Conv_2']['kernel']), dx)
        gb['Conv_2'] = {'kernel': dW2, 'bias': db2}
        dc1 = q(dr2 * (c1 > 0))
        dr1, dW1, db1 = conv_bwd(r1, q(blk['Conv_1']['kernel']), dc1)
        gb['Conv_1'] = {'kernel': dW1, 'bias': db1}
        dpl = q(dx + dr1 * (pl > 0))                               # skip connection + residual branch
        dc0 = q(pool_bwd(dpl, arg, c0_shape, pad_lo))
        dx, dW0, db0 = conv_bwd(x_in, q(blk['Conv_0']['kernel']), dc0)
        dx = q(dx)
        gb['Conv_0'] = {'kernel': dW0, 'bias': db0}
        g[f'stack_blocks_{i}'] = gb
    return g
