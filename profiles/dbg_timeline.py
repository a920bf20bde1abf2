"""Diagnostics: %globaltimer stamps at the schedule points of one captured update step (FQL_B200_STAMPS=1 adds one 1-thread
kernel per point, ~1-2 us each on its stream; the step is therefore a little slower than the bench value)."""
import ctypes as C
import os
import sys
os.environ['FQL_B200_STAMPS'] = '1'
import numpy as np
import torch
sys.path.insert(0, '.')
from fql_b200 import FQLAgent, get_config, _lib

NAMES = {0: 'step start', 1: 'prep done', 2: 'Euler done (S1)', 3: 'one-step fwd done', 4: 'bc-flow dgrad chain done (S2)',
         5: 'critic fwd done', 6: 'critic input-grad chain done', 7: 'join Euler + dL/da done', 8: 'bc+critic grads complete (S2)',
         9: 'early optimizer pass done (S2)', 10: 'one-step grads complete', 11: 'optimizer pass done', 12: 'step end'}
B, F, A = int(os.environ.get('B', 256)), 29, 8
cfg = get_config()
cfg['q_agg'] = 'min'
cfg['alpha'] = 10.0
cfg['batch_size'] = B
rng = np.random.default_rng(0)
agent = FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), cfg, precision='bf16')
batch = {k: torch.as_tensor(v, device='cuda') for k, v in dict(
    observations=rng.standard_normal((B, F)).astype(np.float32), next_observations=rng.standard_normal((B, F)).astype(np.float32),
    actions=rng.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rng.standard_normal(B).astype(np.float32),
    masks=np.ones(B, np.float32)).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
lib = C.CDLL(_lib.LIB_PATH)
lib.fql_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
acc = []
for it in range(30):
    flush.fill_(it & 1)
    agent.update(batch)
    torch.cuda.synchronize()
    out = np.zeros(16, np.uint64)
    assert lib.fql_debug_stamps(agent._ctx, out.ctypes.data, 16) == 0
    if it >= 10:
        acc.append((out[:13].astype(np.int64) - int(out[0])) / 1e3)
m = np.median(np.stack(acc), axis=0)
for i in np.argsort(m):
    print(f'{m[i]:8.1f} us  [{i:2d}] {NAMES[i]}')
