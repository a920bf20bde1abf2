"""The implicit-GEMM 3x3 convolutions of the tensor-core ImpalaEncoder (fql_b200/csrc/encoder_tc.cu) through the C ABI, checked
EXACTLY: inputs and weights are bf16-representable, the kernels accumulate in fp32, so forward / input-gradient outputs equal the
fp64 convolution up to the final bf16 rounding (2^-8 relative) and the weight / bias gradients (fp32 outputs) to 1e-5.
Reference semantics: flax nn.Conv(features, (3, 3), padding='SAME') on NHWC with an HWIO kernel (utils/encoders.py:19-54) and
jax.grad of it; the NumPy convolution below is oracle/encoder_oracle.py's."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as E

pytestmark = pytest.mark.gpu


def bf16_exact(rng, shape, scale=1.0):
    """random values that survive a round trip through bf16"""
    x = torch.as_tensor(rng.standard_normal(shape).astype(np.float32) * scale)
    return x.to(torch.bfloat16)


def conv_ref(x, w, b):
    """SAME 3x3 convolution, NHWC x HWIO, fp64"""
    B, H, W, Ci = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    out = np.zeros((B, H, W, w.shape[3]))
    for ky in range(3):
        for kx in range(3):
            out += np.einsum('bhwc,co->bhwo', xp[:, ky:ky + H, kx:kx + W, :], w[ky, kx])
    return out + b


@pytest.mark.parametrize('B,H,W,cin,cout', [(3, 8, 8, 16, 16), (2, 16, 16, 16, 32), (5, 20, 12, 32, 32), (1, 5, 3, 32, 32), (64, 32, 32, 16, 16),
                                            (300, 8, 8, 32, 32), (128, 64, 64, 16, 16), (192, 32, 32, 32, 32)])   # the last two: ~28 / ~10 tiles per persistent CTA
def test_conv3x3_forward_dgrad_wgrad_exact(B, H, W, cin, cout):
    from fql_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(B * 1000 + H)
    dev = torch.device('cuda')
    x = bf16_exact(rng, (B, H, W, cin)).to(dev)
    w = bf16_exact(rng, (3, 3, cin, cout), 0.2).float().to(dev)
    b = torch.as_tensor(rng.standard_normal(cout).astype(np.float32)).to(dev)
    dy = bf16_exact(rng, (B, H, W, cout)).to(dev)
    skip = bf16_exact(rng, (B, H, W, cout)).to(dev)
    wsb = int(lib.fql_conv3x3_workspace_bytes())
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    x64, w64, dy64 = x.float().cpu().numpy().astype(np.float64), w.cpu().numpy().astype(np.float64), dy.float().cpu().numpy().astype(np.float64)

    def close_bf16(got, ref, what):
        got = got.float().cpu().numpy().astype(np.float64)
        tol = 2.0 ** -8 * np.abs(ref) + 1e-6 * np.abs(ref).max()
        bad = np.abs(got - ref) > tol
        assert not bad.any(), (what, int(bad.sum()), float(np.abs(got - ref).max()), float(np.abs(ref).max()))

    # forward: relu(conv + bias), and conv + bias + skip
    out = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.fql_conv3x3_bf16(p(x), p(w), p(b), B, H, W, cin, cout, 0, None, None, 1, p(out), p(ws), wsb, st), 'conv fwd')
    ref = conv_ref(x64, w64, b.cpu().numpy().astype(np.float64))
    close_bf16(out, np.maximum(ref, 0), 'forward relu')
    _lib.check(lib.fql_conv3x3_bf16(p(x), p(w), p(b), B, H, W, cin, cout, 0, None, p(skip), 0, p(out), p(ws), wsb, st), 'conv fwd skip')
    close_bf16(out, ref + skip.float().cpu().numpy(), 'forward + skip')
    # input gradient with a relu mask and an upstream addend: dX = mask * conv^T(dY) + add   (wgrad has no (32 -> 16) instantiation)
    if not (cin == 32 and cout == 16):
        maskt = bf16_exact(rng, (B, H, W, cin)).to(dev)
        addt = bf16_exact(rng, (B, H, W, cin)).to(dev)
        dx = torch.empty(B, H, W, cin, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.fql_conv3x3_bf16(p(dy), p(w), None, B, H, W, cin, cout, 1, p(maskt), p(addt), 0, p(dx), p(ws), wsb, st), 'conv dgrad')
        wt = np.flip(w64, (0, 1)).transpose(0, 1, 3, 2)               # conv^T = conv with flipped taps and swapped channel roles
        ref_dx = conv_ref(dy64, wt, 0.0) * (maskt.float().cpu().numpy() > 0) + addt.float().cpu().numpy()
        close_bf16(dx, ref_dx, 'input gradient')
        # weight / bias gradient: fp32 outputs of exact products
        gw = torch.zeros(3, 3, cin, cout, device=dev)
        gb = torch.zeros(cout, device=dev)
        _lib.check(lib.fql_conv3x3_wgrad_bf16(p(x), p(dy), B, H, W, cin, cout, p(gw), p(gb), p(ws), wsb, st), 'conv wgrad')
        xp = np.pad(x64, ((0, 0), (1, 1), (1, 1), (0, 0)))
        ref_gw = np.stack([np.stack([np.einsum('bhwc,bhwo->co', xp[:, ky:ky + H, kx:kx + W, :], dy64) for kx in range(3)]) for ky in range(3)])
        ref_gb = dy64.sum((0, 1, 2))
        assert np.abs(gw.cpu().numpy() - ref_gw).max() <= 1e-5 * np.abs(ref_gw).max(), np.abs(gw.cpu().numpy() - ref_gw).max() / np.abs(ref_gw).max()
        assert np.abs(gb.cpu().numpy() - ref_gb).max() <= 1e-5 * max(np.abs(ref_gb).max(), np.abs(dy64).sum((0, 1, 2)).max() * 0 + 1.0)
    torch.cuda.synchronize()
