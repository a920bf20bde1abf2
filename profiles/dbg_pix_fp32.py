"""Diagnostics: per-leaf gradient error of the fp32 pixel step against the fp64 oracle (leaves above 1e-4 are printed)."""
import copy, os, sys
import numpy as np
sys.path.insert(0, '.')
from oracle import fql_oracle as O
from oracle import fql_pixel_oracle as PO
from tests.helpers import f32, rel_err
from fql_b200 import FQLAgent
B, hw, ch, A, hidden = int(os.environ.get('B', 8)), int(os.environ.get('HW', 16)), int(os.environ.get('CH', 6)), 3, int(os.environ.get('HIDDEN', 512))
cfg = dict(O.DEFAULT_CONFIG); cfg.update(alpha=10.0)
cfg.update(actor_hidden_dims=(hidden,) * 4, value_hidden_dims=(hidden,) * 4, encoder='impala_small')
params = PO.init_params(3, ch, A, cfg, dtype=np.float64, hw=hw, jitter=0.05, target_equals_critic=False)
state = O.init_state(params, warm=True, seed=3)
batch = PO.make_pixel_batch(4, B, A, hw=hw, ch=ch, dtype=np.float64)
noise = O.make_noise(5, B, A, np.float64)
st64, info64, grads64 = PO.update(copy.deepcopy(state), cfg, batch, noise)
b32 = {k: (v if v.dtype == np.uint8 else v.astype(np.float32)) for k, v in batch.items()}
c = dict(cfg); c['batch_size'] = B
ag = FQLAgent.create(0, np.zeros((1, hw, hw, ch), np.uint8), np.zeros((1, A), np.float32), c, precision='fp32')
ag.load_tree(f32(state['params']), f32(state['mu']), f32(state['nu']), state['count'])
_, info = ag.update(b32, noise=f32(noise))
g = ag.export_tree('grads')
bad = 0
for (path, r), (_, g32) in zip(O.tree_leaves(grads64), O.tree_leaves(g)):
    if np.abs(r).max() == 0:
        continue
    e = rel_err(g32, r)
    if e > 1e-4:
        bad += 1
        d = np.abs(np.asarray(g32, np.float64) - r) / np.abs(r).max()
        idx = np.argwhere(d > 1e-4)
        print(f"BAD {e:9.2e} {'/'.join(path)} {r.shape} n_bad={len(idx)} of {r.size}; first {idx[:12].tolist()}")
print('variant', os.environ.get('VARIANT'), 'B', B, 'bad leaves', bad)
