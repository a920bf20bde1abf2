"""Diagnostics: per-phase globaltimer stamps of tc_gemm_kernel (one hidden-layer forward GEMM, M=256 N=512 K=512)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from fql_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
lib.fql_debug_tc_gemm.argtypes = [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p] * 2
M, N, K = 256, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 512
X = torch.randn(M, K, device='cuda').bfloat16()
W = (torch.randn(K, N, device='cuda') / K ** 0.5).bfloat16()
b = torch.zeros(N, device='cuda')
H = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
nct = (N // 64) * ((M + 127) // 128)
dbg = torch.zeros(nct * 8, dtype=torch.int64, device='cuda')
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for it in range(5):
        rc = lib.fql_debug_tc_gemm(X.data_ptr(), W.data_ptr(), b.data_ptr(), H.data_ptr(), M, N, K, dbg.data_ptr(), st.cuda_stream)
        assert rc == 0
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(X.float() @ W.float(), approximate='tanh')
    print('max err', (H.float() - ref).abs().max().item())
    d = dbg.cpu().numpy().reshape(nct, 8)
    t0 = d[:, 0].min()
    print('per-CTA ns since first CTA start: start, setup_done, first_full, last_full, acc_full, epi_done, end')
    for r in d[:4]:
        print([int(x - t0) for x in r[:7]])
    print('mean', [float((d[:, i] - d[:, 0]).mean()) for i in range(7)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(100):
        lib.fql_debug_tc_gemm(X.data_ptr(), W.data_ptr(), b.data_ptr(), H.data_ptr(), M, N, K, None, st.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print('back-to-back launches: us per GEMM', e0.elapsed_time(e1) * 10)
