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


def stacks():
    """Whole 6-layer stack: fused one-launch kernel (bf16 engine) vs the per-layer fp32 SIMT path (fp32 engine)."""
    from .. import synth
    from ..engine import Engine
    sd = synth.make_state_dict(seed=0)
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for B, H in ((64, 256), (32, 512)):
        for prec in ("bf16", "fp32"):
            eng = Engine(sd, dev, precision=prec, n_clusters=8)
            # token buffers only (the conv workspace of the fp32 engine at batch 64 is not needed here)
            S = (H // 16) ** 2
            from ..engine import position_table
            eng._pos[(H // 16, H // 16)] = position_table(H // 16, H // 16, dev)
            M = B * S
            f32 = dict(dtype=torch.float32, device=dev)
            tok = dict(qkv=torch.empty(M, 192, **f32), att=torch.empty(M, 64, **f32), xa=torch.empty(M, 64, **f32),
                       xb=torch.empty(M, 64, **f32))
            ws = dict(tok=tok, S=S, h=H // 16, w=H // 16)
            x = torch.randn(M, 64, **f32)
            out = torch.empty(M, 64, **f32)
            us = _time(lambda: eng._encoder_stack("wildpath", x, out, ws, B, st), iters=10)
            flops = 6 * (2 * M * 64 * (192 + 64 + 512) + 4 * M * S * 64)
            print(f"encoder stack ({'fused mma.sync split-bf16' if eng.fused_tokens else 'per-layer fp32 SIMT'}) B={B} S={S}: "
                  f"{us:8.1f} us  = {flops / us / 1e6:.1f} TFLOP/s useful")


if __name__ == "__main__":
    import sys
    if "--stacks" in sys.argv:
        stacks()
        sys.exit(0)
    main()
