"""Times the token-path kernels through the C ABI (CUDA events): attention core and fused encoder-layer tail.

    python -m disentangledcolorization_b200.tools.token_probe
"""
import ctypes as C

import torch

from .. import _lib


def _time(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    hd = _lib.Handle.get(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for B, S in ((64, 256), (32, 1024)):
        M = B * S
        qkv = torch.randn(M, 192, device="cuda")
        out = torch.empty(M, 64, device="cuda")
        us = _time(lambda: _lib.check(hd.lib.disco_attention(hd.h, p(qkv), B, S, p(out), st), "attention"))
        pairs = B * 8 * S * S
        print(f"attention    B={B} S={S}: {us:7.1f} us  ({pairs * 16 / us / 1e6:.1f} TFMA/s of ~36 peak)")
        x, att, y = (torch.randn(M, 64, device="cuda") for _ in range(3))
        wo, w1, w2 = torch.randn(64, 64, device="cuda"), torch.randn(256, 64, device="cuda"), torch.randn(64, 256, device="cuda")
        v64 = [torch.randn(64, device="cuda") for _ in range(6)]
        b1 = torch.randn(256, device="cuda")
        us = _time(lambda: _lib.check(hd.lib.disco_encoder_tail(hd.h, p(att), p(x), p(y), M, p(wo), p(v64[0]), p(v64[1]), p(v64[2]),
                                                                 p(w1), p(b1), p(w2), p(v64[3]), p(v64[4]), p(v64[5]), st), "tail"))
        print(f"encoder_tail M={M}:        {us:7.1f} us  ({M * (64 * 64 + 2 * 64 * 256) / us / 1e6:.1f} TFMA/s of ~36 peak)")


if __name__ == "__main__":
    main()
