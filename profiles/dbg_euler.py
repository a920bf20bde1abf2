"""Diagnostics: phase stamps of one layer iteration inside euler_cluster_kernel + total kernel time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
NCTA = 32 if os.environ.get('FQL_B200_EULER_NC', '16') != '8' else 16
dbg = torch.zeros(NCTA * 16, dtype=torch.int64, device='cuda')
os.environ['FQL_B200_EULER_DBG'] = hex(dbg.data_ptr())
from fql_b200 import FQLAgent, get_config
cfg = get_config(); cfg.update(dict(q_agg='min', alpha=10.0)); cfg['batch_size'] = 256
ag = FQLAgent.create(0, np.zeros((1, 29), np.float32), np.zeros((1, 8), np.float32), cfg, precision='bf16')
obs = np.random.randn(256, 29).astype(np.float32); nz = np.random.randn(256, 8).astype(np.float32)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(3): ag.compute_flow_actions(obs, nz)
    torch.cuda.synchronize()
    d = dbg.cpu().numpy().reshape(NCTA, 16)
    base = d[:, 0].min()
    print('ns relative to the earliest acc_full(l-1) over all CTAs')
    print('cta  acc_full(l-1) | half0 stored, half1 stored | mcast h0, h1 issued | mma: first sub-block seen, last seen | acc_full(l)')
    for cta in range(0, NCTA, 3):
        r = d[cta]
        print(cta, int(r[0] - base), '|', int(r[1] - base), int(r[2] - base), '|', int(r[3] - base), int(r[4] - base), '|', int(r[5] - base), int(r[6] - base), '|', int(r[7] - base))
    print('median layer time (acc_full to acc_full):', float(np.median(d[:, 7] - d[:, 0])), 'ns')
    import time
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    obs_t = torch.tensor(obs, device='cuda'); nz_t = torch.tensor(nz, device='cuda')
    e0.record()
    for _ in range(20): ag._fwd_call(ag._lib.fql_compute_flow_actions, obs_t, nz_t)
    e1.record(); torch.cuda.synchronize()
    print('compute_flow_actions (concat+pad+cluster kernel) us per call:', e0.elapsed_time(e1) * 1000 / 20)
