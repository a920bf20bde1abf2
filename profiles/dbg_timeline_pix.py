"""Diagnostics: %globaltimer stamps at the schedule points of one captured update step of the pixel configuration (config 5)."""
import ctypes as C
import os
import sys
os.environ['FQL_B200_STAMPS'] = '1'
import numpy as np
import torch
sys.path.insert(0, '.')
from fql_b200 import FQLAgent, get_config, _lib

NAMES = {0: 'step start', 1: 'encoders + prep done', 2: 'Euler done (S1)', 3: 'one-step fwd done', 4: 'bc-flow backward chain (+ encoder backward) done (S2)',
         5: 'critic fwd done', 6: 'critic input-grad chain done', 7: 'join Euler + dL/da done', 8: 'bc+critic grads (+ encoder backwards) complete (S2)',
         10: 'one-step grads (+ encoder backward) complete', 11: 'optimizer pass done', 12: 'step end'}
B, A = int(os.environ.get('B', 256)), 5
cfg = get_config()
cfg.update(alpha=300.0, encoder='impala_small', batch_size=B)
rng = np.random.default_rng(0)
agent = FQLAgent.create(0, np.zeros((1, 64, 64, 9), np.uint8), np.zeros((1, A), np.float32), cfg, precision=os.environ.get('PREC', 'bf16'))
batch = dict(observations=torch.as_tensor(rng.integers(0, 256, (B, 64, 64, 9), dtype=np.uint8), device='cuda'),
             next_observations=torch.as_tensor(rng.integers(0, 256, (B, 64, 64, 9), dtype=np.uint8), device='cuda'),
             actions=torch.as_tensor(rng.uniform(-1, 1, (B, A)).astype(np.float32), device='cuda'),
             rewards=torch.as_tensor(rng.standard_normal(B).astype(np.float32), device='cuda'), masks=torch.ones(B, device='cuda'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
lib = C.CDLL(_lib.LIB_PATH)
lib.fql_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
acc = []
for it in range(20):
    flush.fill_(it & 1)
    agent.update(batch)
    torch.cuda.synchronize()
    out = np.zeros(576, np.uint64)
    assert lib.fql_debug_stamps(agent._ctx, out.ctypes.data, 576) == 0
    if it >= 6:
        acc.append((out[:19].astype(np.int64) - int(out[0])) / 1e3)
m = np.median(np.stack(acc), axis=0)
for i in np.argsort(m):
    if i in NAMES and abs(m[i]) < 1e6:
        print(f'{m[i]:8.1f} us  [{i:2d}] {NAMES[i]}')
