"""Diagnostics: per-phase globaltimer stamps of tc_gemm_kernel for the four operand-majorness combinations (M=256 N=512 K=512)."""
import ctypes as C
import os
import sys
os.environ['FQL_B200_PDL'] = '0'
import torch
sys.path.insert(0, '.')
from fql_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
lib.fql_debug_tc_gemm.argtypes = [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p] * 2 + [C.c_int] * 2
M, N, K = 256, 512, 512
X = torch.randn(M, K, device='cuda').bfloat16()
W = (torch.randn(K, N, device='cuda') / K ** 0.5).bfloat16()
b = torch.zeros(N, device='cuda')
H = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
nct = (N // 64) * ((M + 127) // 128)
dbg = torch.zeros(nct * 8, dtype=torch.int64, device='cuda')
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            for it in range(4):
                assert lib.fql_debug_tc_gemm(X.data_ptr(), W.data_ptr(), b.data_ptr(), H.data_ptr(), M, N, K, dbg.data_ptr(), st.cuda_stream, a_mn, b_mn) == 0
            torch.cuda.synchronize()
            d = dbg.cpu().numpy().reshape(nct, 8)
            m = [float((d[:, i] - d[:, 0]).mean()) for i in range(7)]
            print(f'a_mn={a_mn} b_mn={b_mn}: setup {m[1]:.0f}  first block seen {m[2]:.0f}  last block seen {m[3]:.0f}  acc_full {m[4]:.0f}  epilogue done {m[5]:.0f}  end {m[6]:.0f} ns'
                  f'   => MMA phase per k-block {(m[3] - m[2]) / 7:.0f} ns')
