import ctypes as C, torch
from disentangledcolorization_b200 import _lib
hd = _lib.Handle.get(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for B, S in ((64, 256), (32, 1024)):
    qkv = torch.randn(B * S, 192, device="cuda")
    out = torch.empty(B * S, 64, device="cuda")
    for _ in range(3):
        _lib.check(hd.lib.disco_attention(hd.h, C.c_void_p(qkv.data_ptr()), B, S, C.c_void_p(out.data_ptr()), st), "att")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.check(hd.lib.disco_attention(hd.h, C.c_void_p(qkv.data_ptr()), B, S, C.c_void_p(out.data_ptr()), st), "att")
    e1.record()
    torch.cuda.synchronize()
    print(f"attention B={B} S={S}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
