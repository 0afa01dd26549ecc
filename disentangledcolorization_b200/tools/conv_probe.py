"""Times one fused-conv configuration through the C ABI (CUDA events) -- kernel tuning / ncu target.

    python -m disentangledcolorization_b200.tools.conv_probe --cin 64 --cout 64 --hw 256 --batch 64
"""
import argparse
import ctypes as C

import torch

from .. import _lib


def build_desc(handle, cin, cout, B, H, W, stride=1, up2=0, head=0, act=1, post=True, res=False, kind=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    keep = {}
    if kind == _lib.DECONV4:
        w = torch.randn(cin, cout, 4, 4, generator=g) / (cin * 4) ** 0.5
        blk = w.permute(2, 3, 0, 1).reshape(-1)
        Ho, Wo = 2 * H, 2 * W
    else:
        w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        blk = w.permute(2, 3, 1, 0).reshape(-1)
        Ho, Wo = (H << up2) // stride, (W << up2) // stride
    x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16).cuda()
    w32_host = blk.contiguous()
    d = _lib.ConvDesc()
    d.kind, d.stride, d.dtype, d.batch, d.Ho, d.Wo, d.Cout, d.n_src = kind, stride, _lib.BF16, B, Ho, Wo, cout, 1
    d.src[0].ptr, d.src[0].H, d.src[0].W, d.src[0].C, d.src[0].up2 = x.data_ptr(), H, W, cin, up2
    bias = torch.randn(cout, generator=g).cuda()
    d.bias = bias.data_ptr()
    keep.update(x=x, bias=bias)
    if post:
        ps, pb = (torch.rand(cout, generator=g) + 0.5).cuda(), torch.randn(cout, generator=g).cuda()
        d.post_scale, d.post_shift = ps.data_ptr(), pb.data_ptr()
        keep.update(ps=ps, pb=pb)
    if res:
        r = torch.randn(B, Ho, Wo, cout, generator=g).to(torch.bfloat16).cuda()
        d.residual = r.data_ptr()
        keep.update(r=r)
    d.act, d.slope, d.head = act, 0.2, head
    bh = bias.cpu().contiguous()
    d.bias_host = bh.data_ptr()
    keep.update(bh=bh)
    if post:
        psh, pbh = keep["ps"].cpu().contiguous(), keep["pb"].cpu().contiguous()
        d.post_scale_host, d.post_shift_host = psh.data_ptr(), pbh.data_ptr()
        keep.update(psh=psh, pbh=pbh)
    out = (torch.empty(B, cout, Ho, Wo, device="cuda") if head else
           torch.empty(B, Ho, Wo, cout, device="cuda", dtype=torch.bfloat16))
    d.out = out.data_ptr()
    keep.update(out=out)
    if handle.lib.disco_conv_tc_supported(handle.h, C.byref(d)):
        n = int(handle.lib.disco_conv_tc_weight_elems(C.byref(d)))
        w16 = torch.empty(n, dtype=torch.int16)
        _lib.check(handle.lib.disco_conv_tc_pack_weights(C.byref(d), C.c_void_p(w32_host.data_ptr()),
                                                         C.c_void_p(w16.data_ptr())), "pack")
        wd = w16.cuda()
        tc = True
    else:
        wd = w32_host.cuda()
        tc = False
    d.weights = wd.data_ptr()
    keep.update(w=wd)
    flops = 2.0 * B * Ho * Wo * cin * cout * (16 / 4 if kind == _lib.DECONV4 else 9)
    return d, keep, flops, tc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--up2", type=int, default=0)
    ap.add_argument("--head", type=int, default=0)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    h = _lib.Handle.get(0)
    d, keep, flops, tc = build_desc(h, a.cin, a.cout, a.batch, a.hw, a.hw, a.stride, a.up2, a.head, res=bool(a.res), kind=a.kind)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        _lib.check(h.lib.disco_conv(h.h, C.byref(d), st), "conv")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        _lib.check(h.lib.disco_conv(h.h, C.byref(d), st), "conv")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    import os
    if os.environ.get("DISCO_TC_DEBUG"):
        import numpy as np
        buf = np.zeros(4 * 48 * 8, dtype=np.int64)
        h.lib.disco_debug_timeline(C.c_void_p(buf.ctypes.data))
        t = buf.reshape(4, 48, 8)
        t0 = t[t > 0].min() if (t > 0).any() else 0
        rel = np.where(t > 0, t - t0, -1)
        names = ["producer(step issue)", "mma(0 pre-tempty,1 post,2.. full[i],6 commit)", "epi w2 (0 pre,1 got tfull,2 done)", "epi w6"]
        for r in range(4):
            print(names[r])
            for ti in range(0, 24):
                print("  tile", ti, rel[r, ti].tolist())
    print(f"cin={a.cin} cout={a.cout} hw={a.hw} B={a.batch} stride={a.stride} up2={a.up2} head={a.head} kind={a.kind} "
          f"tc={tc}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s (reference-formulation FLOPs)")


if __name__ == "__main__":
    main()
