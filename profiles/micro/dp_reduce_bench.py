"""Microbenchmark of the data-parallel bucket exchange (dp_comm.cu) alone: time per call of fql_dp_allreduce for a few range
sizes, both ranks in lockstep (torchrun, one rank per GPU).  Grid / unroll / threads are taken from FQL_DP_CTAS / FQL_DP_UNROLL /
FQL_DP_THREADS by the library."""
import ctypes as C
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
from fql_b200 import FQLAgent, get_config, _lib

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
cfg = get_config()
cfg.update(q_agg='min', alpha=10.0, batch_size=256)
agent = FQLAgent.create(0, np.zeros((1, 29), np.float32), np.zeros((1, 8), np.float32), cfg, precision='bf16', process_group=dist.group.WORLD)
lib = agent._lib
st = torch.cuda.Stream()
res = {}
with torch.cuda.stream(st):
    for n in (0, 4096, 409600, 1638400, 3276800):
        agent._grads.fill_(1.0)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        for _ in range(5):
            _lib.check(lib.fql_dp_allreduce(agent._ctx, 3, 0, n, C.c_void_p(st.cuda_stream)), 'dp')
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 50
        e0.record()
        for _ in range(K):
            _lib.check(lib.fql_dp_allreduce(agent._ctx, 3, 0, n, C.c_void_p(st.cuda_stream)), 'dp')
        e1.record()
        torch.cuda.synchronize()
        res[n] = 1e3 * e0.elapsed_time(e1) / K
        if n:
            v = float(agent._grads[0, 0].item())
            assert np.isfinite(v)
if rank == 0:
    tag = f"ctas={os.environ.get('FQL_DP_CTAS', 'def')} unroll={os.environ.get('FQL_DP_UNROLL', 'def')} threads={os.environ.get('FQL_DP_THREADS', 'def')} mc={os.environ.get('FQL_DP_MULTICAST', '1')}"
    print(tag, ' '.join(f'{n * 4 / 1e6:.2f}MB:{t:.1f}us' for n, t in res.items()), flush=True)
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
