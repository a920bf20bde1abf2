"""Diagnostics: the %globaltimer stamps of profiles/dbg_timeline.py for the data-parallel step (run under torchrun, one rank per
GPU): where the bucket reductions sit in the step and how long each waits for its peers."""
import ctypes as C
import os
import sys
os.environ['FQL_B200_STAMPS'] = '1'
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
from fql_b200 import FQLAgent, get_config, _lib

NAMES = {0: 'step start', 1: 'prep done', 2: 'Euler done (S1)', 3: 'one-step fwd done', 4: 'bc-flow dgrad chain done (S2)',
         5: 'critic fwd done', 6: 'critic input-grad chain done', 7: 'join Euler + dL/da done', 8: 'bc+critic grads complete (S2)',
         10: 'one-step grads complete', 11: 'optimizer pass done', 12: 'step end',
         20: 'DP bc-flow bucket: gradients final', 21: 'DP bc-flow bucket reduced', 22: 'DP critic bucket: gradients final',
         23: 'DP critic bucket reduced', 24: 'DP one-step bucket: gradients final', 25: 'DP one-step bucket reduced'}
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
B, F, A = int(os.environ.get('B', 256)), 29, 8
cfg = get_config()
cfg.update(q_agg='min', alpha=10.0, batch_size=B)
rng = np.random.default_rng(rank)
agent = FQLAgent.create(0, np.zeros((1, F), np.float32), np.zeros((1, A), np.float32), cfg, precision='bf16', process_group=dist.group.WORLD)
batch = {k: torch.as_tensor(v, device='cuda') for k, v in dict(
    observations=rng.standard_normal((B, F)).astype(np.float32), next_observations=rng.standard_normal((B, F)).astype(np.float32),
    actions=rng.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rng.standard_normal(B).astype(np.float32),
    masks=np.ones(B, np.float32)).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
lib = C.CDLL(_lib.LIB_PATH)
lib.fql_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
acc = []
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    bufs = agent.stage(batch)
    for it in range(40):
        flush.fill_(it & 1)
        if os.environ.get('SYNC_EACH', '1') == '1':
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        agent.step(bufs)
        torch.cuda.synchronize()
        out = np.zeros(576, np.uint64)
        assert lib.fql_debug_stamps(agent._ctx, out.ctypes.data, 576) == 0
        if it >= 10:
            acc.append((out[:26].astype(np.int64) - int(out[0])) / 1e3)
m = np.median(np.stack(acc), axis=0)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f'---- rank {rank} ({agent.dp_transport})')
        for i in np.argsort(m):
            if i in NAMES and abs(m[i]) < 1e6:
                print(f'{m[i]:8.1f} us  [{i:2d}] {NAMES[i]}')
        sys.stdout.flush()
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)      # diagnostics only: skip the communicator / symmetric-memory teardown
