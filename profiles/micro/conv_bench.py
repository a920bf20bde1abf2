"""Microbenchmark of the tensor-core convolution kernels (encoder_tc.cu) at the config-5 shapes: us per launch (CUDA events, L2
flushed between launches) and the HBM traffic each needs at least (read x once, write y once / read x and dy once)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from fql_b200 import _lib
lib = _lib.lib()
dev = torch.device('cuda')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wsb = int(lib.fql_conv3x3_workspace_bytes())
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
for (H, cin, cout) in [(64, 16, 16), (32, 16, 16), (32, 16, 32), (16, 32, 32), (8, 32, 32)]:
    x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, H, H, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(3, 3, cin, cout, device=dev) * 0.1
    b = torch.zeros(cout, device=dev)
    out = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device=dev)
    gw = torch.zeros(3, 3, cin, cout, device=dev); gb = torch.zeros(cout, device=dev)
    res = {}
    for name, fn in (('fwd', lambda: lib.fql_conv3x3_bf16(p(x), p(w), p(b), B, H, H, cin, cout, 0, None, None, 1, p(out), p(ws), wsb, st)),
                     ('wgrad', lambda: lib.fql_conv3x3_wgrad_bf16(p(x), p(dy), B, H, H, cin, cout, p(gw), p(gb), p(ws), wsb, st))):
        for _ in range(3):
            assert fn() == 0
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[name] = float(np.median(ts))
    npix = B * H * H
    mb_f = npix * (cin + cout) * 2 / 1e6
    fl = 2 * npix * 9 * cin * cout / 1e9
    print(f'B={B} {H}x{H} {cin}->{cout}: fwd {res["fwd"]:7.1f} us (incl. ~5 us weight prep; {mb_f:.1f} MB min = {mb_f / 6.5e3 * 1e3:.1f} us at HBM peak, {fl:.2f} GFLOP)'
          f'  wgrad {res["wgrad"]:7.1f} us (incl. reduce)')
