"""Diagnostics (needs a build with `make -C fql_b200/csrc EXTRA=-DFQL_C2_DBG`): per-stage %globaltimer stamps of the MMA issuer and the
TMA producer of CTA 0 inside mlp_chain2_kernel<EULER>, one 512 x 512 layer (iteration FQL_C2_DBG_N) of a humanoidmaze-shaped batch."""
import ctypes as C
import os
import sys
os.environ.setdefault('FQL_B200_GRAPH', '0')
import numpy as np
import torch
sys.path.insert(0, '.')
from fql_b200 import FQLAgent, get_config, _lib
B, F, A = int(os.environ.get('B', 16384)), 69, 21
cfg = get_config(); cfg.update(discount=0.995, batch_size=B)
rng = np.random.default_rng(0)
agent = FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), cfg, precision='bf16')
batch = dict(observations=torch.as_tensor(rng.standard_normal((B, F)).astype(np.float32), device='cuda'),
             next_observations=torch.as_tensor(rng.standard_normal((B, F)).astype(np.float32), device='cuda'),
             actions=torch.as_tensor(rng.uniform(-1, 1, (B, A)).astype(np.float32), device='cuda'),
             rewards=torch.as_tensor(rng.standard_normal(B).astype(np.float32), device='cuda'), masks=torch.ones(B, device='cuda'))
lib = C.CDLL(_lib.LIB_PATH)
for it in range(4):
    agent.update(batch)
    torch.cuda.synchronize()
out = np.zeros(256, np.uint64)
assert lib.fql_debug_chain2_stamps(out.ctypes.data_as(C.c_void_p), 256) == 0
t = out.astype(np.int64)
t0 = t[0]
for h in range(2):
    print(f'half {h}:  stage: wait start, full seen (+wait), committed (+issue)   | producer saw the slot empty')
    for ks in range(16):
        a, b, c = t[h * 48 + ks * 3: h * 48 + ks * 3 + 3] - t0
        pe = t[128 + h * 16 + ks] - t0
        print(f'  ks {ks:2d}: {a:6d} ns  {b:6d} (+{b - a:4d})  {c:6d} (+{c - b:4d})   | {pe:6d}')
