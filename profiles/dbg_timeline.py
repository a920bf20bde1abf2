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
         9: 'early optimizer pass done (S2)', 13: 'Euler kernel CTA 0 started', 14: 'Euler kernel CTA 0 finished', 15: 'one-step cluster fwd CTA 0 started', 16: 'one-step cluster fwd CTA 0 finished', 17: 'one-step cluster dgrad CTA 0 started', 18: 'one-step cluster dgrad CTA 0 finished', 10: 'one-step grads complete', 11: 'optimizer pass done', 12: 'step end'}
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
eul = []
for it in range(30):
    flush.fill_(it & 1)
    agent.update(batch)
    torch.cuda.synchronize()
    out = np.zeros(576, np.uint64)
    assert lib.fql_debug_stamps(agent._ctx, out.ctypes.data, 576) == 0
    if it >= 10:
        acc.append((out[:19].astype(np.int64) - int(out[0])) / 1e3)
        eul.append(out[64:].astype(np.int64).reshape(32, 16))
m = np.median(np.stack(acc), axis=0)
for i in np.argsort(m):
    if i in NAMES and abs(m[i]) < 1e6:
        print(f'{m[i]:8.1f} us  [{i:2d}] {NAMES[i]}')

e = np.stack(eul).astype(np.float64)            # [iters][cta][16]
e = e[:, e[0, :, 0] > 0]
ph = {'epilogue (acc_full -> stored)': e[..., 1] - e[..., 0], 'publish (stored -> multicast issued)': e[..., 3] - e[..., 1],
      'landing (issued -> first sub-block seen by MMA warp 0)': e[..., 5] - e[..., 3], 'first -> last sub-block seen (MMA issue)': e[..., 6] - e[..., 5],
      'last seen -> acc_full': e[..., 7] - e[..., 6], 'layer total': e[..., 7] - e[..., 0]}
print('Euler cluster kernel, one hidden layer inside the step (median over CTAs and iterations, ns):')
for k, v in ph.items():
    print(f'  {np.median(v):7.0f}  {k}')

if int(os.environ.get('FQL_B200_EULER_DBG_IT', '7')) % 5 == 0:
    tr = {'acc_full(last) -> accumulators read and summed': e[..., 13] - e[..., 0], 'acc_full(last) -> operand tile updated': e[..., 11] - e[..., 0], '-> fence + a_ready arrive': e[..., 12] - e[..., 11],
          'acc_full(last) -> MMA warp sees a_ready': e[..., 8] - e[..., 0], '-> sees first-layer weights': e[..., 9] - e[..., 8],
          '-> MMAs issued': e[..., 10] - e[..., 9], '-> epilogue sees acc_full(first)': e[..., 7] - e[..., 10]}
    print('last layer -> first layer of the next Euler step (ns):')
    for k, v in tr.items():
        print(f'  {np.median(v):7.0f}  {k}')
