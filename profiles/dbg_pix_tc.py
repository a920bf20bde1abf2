"""Diagnostics: per-leaf gradient error of the tensor-core encoder step (FQL_PRECISION_BF16_ENC) against the fp64 oracle and against
the fp32 CUDA-core step on the same inputs."""
import copy, os, sys
import numpy as np
sys.path.insert(0, '.')
from oracle import fql_oracle as O
from oracle import fql_pixel_oracle as PO
from tests.helpers import f32, rel_err
from fql_b200 import FQLAgent
B, hw, ch, A, hidden = int(os.environ.get('B', 6)), int(os.environ.get('HW', 16)), int(os.environ.get('CH', 6)), 3, int(os.environ.get('HIDDEN', 64))
cfg = dict(O.DEFAULT_CONFIG); cfg.update(alpha=10.0)
cfg.update(actor_hidden_dims=(hidden,) * 4, value_hidden_dims=(hidden,) * 4, encoder='impala_small')
params = PO.init_params(3, ch, A, cfg, dtype=np.float64, hw=hw, jitter=0.05, target_equals_critic=False)
state = O.init_state(params, warm=True, seed=3)
batch = PO.make_pixel_batch(4, B, A, hw=hw, ch=ch, dtype=np.float64)
noise = O.make_noise(5, B, A, np.float64)
st64, info64, grads64 = PO.update(copy.deepcopy(state), cfg, batch, noise)
from oracle.encoder_oracle import bf16_round
stq, infoq, gradsq = PO.update(copy.deepcopy(state), cfg, batch, noise, enc_q=bf16_round)
b32 = {k: (v if v.dtype == np.uint8 else v.astype(np.float32)) for k, v in batch.items()}
res = {}
for prec in ('fp32', 'bf16'):
    c = dict(cfg); c['batch_size'] = B
    ag = FQLAgent.create(0, np.zeros((1, hw, hw, ch), np.uint8), np.zeros((1, A), np.float32), c, precision=prec)
    ag.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
    _, info = ag.update(b32, noise=f32(noise))
    res[prec] = (ag.export_tree('grads'), dict(info))
print('  bf16-vs-64  fp32-vs-64  bf16-vs-quantized-oracle')
for (path, r), (_, g32), (_, g16), (_, rq) in zip(O.tree_leaves(grads64), O.tree_leaves(res['fp32'][0]), O.tree_leaves(res['bf16'][0]), O.tree_leaves(gradsq)):
    if np.abs(r).max() == 0:
        continue
    print(f"{rel_err(g16, r):9.2e} {rel_err(g32, r):9.2e} {rel_err(g16, rq):9.2e} |ref|max {np.abs(r).max():9.2e}  {'/'.join(path)} {r.shape}")
for k in O.INFO_KEYS:
    print(k, res['bf16'][1][k], res['fp32'][1][k], info64[k], infoq[k])
