"""Drop-in `network` module: SpixelNet / ColorProbNet / HourGlass2 with the reference's
constructor signatures, `forward(x)` contracts (NCHW fp32 in/out) and state_dict keys
(reference models/network.py:125-144,147-236,260-313) -- parameters only; the arithmetic is the
fused-conv launch plan of netspec.py executed by libdisco_b200.so.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib, netspec


class ParamTree(nn.Module):
    """Parameter/buffer container whose nesting reproduces dotted state_dict keys."""

    def __init__(self, entries, prefix=""):
        super().__init__()
        for key, shape, is_param in entries:
            if not key.startswith(prefix):
                continue
            parts = key[len(prefix):].split(".")
            mod = self
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, ParamTree([], ""))
                mod = getattr(mod, p)
            dtype = torch.int64 if parts[-1] == "num_batches_tracked" else torch.float32
            t = torch.zeros(shape, dtype=dtype)
            if is_param:
                mod.register_parameter(parts[-1], nn.Parameter(t))
            else:
                mod.register_buffer(parts[-1], t)


def _init_from_synth(module, prefix):
    """Deterministic finite initialisation (the reference's own init overflows in eval mode)."""
    from . import synth
    sd = synth.make_state_dict(seed=0)
    own = module.state_dict()
    module.load_state_dict({k: sd[prefix + k] for k in own}, strict=True)


class _ConvNet(nn.Module):
    """Base: a standalone conv network executed through disco_conv."""
    _PREFIX = ""
    _NET = ""
    precision = "bf16"

    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_key = None
        self.register_load_state_dict_post_hook(lambda m, keys: m._invalidate())

    def _ops(self):
        raise NotImplementedError

    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def _get_engine(self, device):
        from .engine import _PackedConv
        key = (str(device), self.precision)
        if self._engine is None or self._engine_key != key:
            sd = {self._PREFIX + k: v.detach().cpu() for k, v in self.state_dict().items()}
            packed = [_PackedConv(netspec.fold(sd, op), device) for op in self._ops()]
            self._engine, self._engine_key = packed, key
        return self._engine

    @staticmethod
    def _route(handle, d, pc, dev):
        if d.dtype != _lib.BF16 or not handle.lib.disco_conv_tc_supported(handle.h, C.byref(d)):
            return
        if pc.w16 is None:
            n = int(handle.lib.disco_conv_tc_weight_elems(C.byref(d)))
            w16 = torch.empty(n, dtype=torch.int16)
            _lib.check(handle.lib.disco_conv_tc_pack_weights(C.byref(d), C.c_void_p(pc.w32_host.data_ptr()),
                                                             C.c_void_p(w16.data_ptr())), "disco_conv_tc_pack_weights")
            pc.w16 = w16.to(dev)
        d.weights = pc.w16.data_ptr()
        for i in range(d.n_src):
            if d.src[i].is_f32:
                d.gray_weights = pc.w32.data_ptr() + 4 * int(d.src[i].w_off)

    def _run(self, inputs, out_name):
        """inputs: dict buffer-name -> tensor (gray: (N,1,H,W) fp32; others NHWC in the working dtype)."""
        from .engine import _DT, _ACT, _HEAD
        gray = inputs["gray"]
        dev = gray.device
        if dev.type != "cuda":
            raise _lib.DiscoError("this module runs on a CUDA (B200) device only -- there is no CPU path")
        handle = _lib.Handle.get(dev.index if dev.index is not None else torch.cuda.current_device())
        code, tdt = _DT[self.precision]
        B, _, H, W = gray.shape
        bufs = dict(inputs)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for pc in self._get_engine(dev):
            op = pc.op
            Ho, Wo = H // op.scale, W // op.scale
            if op.head:
                out = torch.empty(B, op.cout, Ho, Wo, dtype=torch.float32, device=dev)
            else:
                out = torch.empty(B, Ho, Wo, op.cout, dtype=tdt, device=dev)
            bufs[op.out] = out
            d = _lib.ConvDesc()
            d.kind = _lib.DECONV4 if op.kind == "deconv4" else _lib.CONV3
            d.stride, d.dtype, d.batch, d.Ho, d.Wo, d.Cout, d.n_src = op.stride, code, B, Ho, Wo, op.cout, len(op.srcs)
            for i, s in enumerate(op.srcs):
                t = bufs[s.buf]
                if s.buf == "gray":
                    d.src[i].H, d.src[i].W, d.src[i].C, d.src[i].is_f32 = t.shape[2], t.shape[3], 1, 1
                else:
                    d.src[i].H, d.src[i].W, d.src[i].C, d.src[i].is_f32 = t.shape[1], t.shape[2], t.shape[3], 0
                d.src[i].ptr, d.src[i].up2, d.src[i].w_off = t.data_ptr(), int(s.up2), pc.w_off[i]
            d.weights, d.bias = pc.w32.data_ptr(), pc.bias.data_ptr()
            d.post_scale = pc.post_scale.data_ptr() if pc.post_scale is not None else None
            d.post_shift = pc.post_shift.data_ptr() if pc.post_shift is not None else None
            d.residual = bufs[op.res].data_ptr() if op.res else None
            d.act, d.slope, d.head, d.out = _ACT[op.act], op.slope, _HEAD[op.head], out.data_ptr()
            self._route(handle, d, pc, dev)
            _lib.check(handle.lib.disco_conv(handle.h, C.byref(d), stream), "disco_conv")
        return bufs[out_name]

    def _run_on(self, inputs, out_name):
        """`_run` with the input's device made current (kernels, allocations and the stream lookup all follow it)."""
        with torch.cuda.device(inputs["gray"].device):
            return self._run(inputs, out_name)


class SpixelNet(_ConvNet):
    """reference models/network.py:260-313: (N,1,H,W) -> (N,9,H,W) soft 9-neighbour assignment."""
    _PREFIX, _NET = "segnet.net.", "segnet"

    def __init__(self, inChannel=3, outChannel=9, batchNorm=True):
        super().__init__()
        if inChannel != 1 or outChannel != 9 or not batchNorm:
            raise _lib.DiscoError("SpixelNet is built for inChannel=1, outChannel=9, batchNorm=True "
                                  "(the only configuration AnchorColorProb uses, models/model.py:40)")
        tree = ParamTree(netspec.schema(nets=("segnet",)), self._PREFIX)
        for name, child in list(tree.named_children()):
            self.add_module(name, child)
        _init_from_synth(self, self._PREFIX)

    def _ops(self):
        return netspec.segnet_ops()

    @torch.no_grad()
    def forward(self, x):
        return self._run_on({"gray": x.float().contiguous()}, "affinity")


class ColorProbNet(_ConvNet):
    """reference models/network.py:147-236: (N,1,H,W) -> (N,64,H,W) non-negative features."""
    _PREFIX, _NET = "repnet.", "repnet"

    def __init__(self, inChannel=1, outChannel=2, with_SA=False):
        super().__init__()
        if inChannel != 1 or outChannel != 64 or with_SA:
            raise _lib.DiscoError("ColorProbNet is built for inChannel=1, outChannel=64, with_SA=False "
                                  "(models/model.py:41; `Self_Attn` is undefined in the reference)")
        tree = ParamTree(netspec.schema(nets=("repnet",)), self._PREFIX)
        for name, child in list(tree.named_children()):
            self.add_module(name, child)
        _init_from_synth(self, self._PREFIX)

    def _ops(self):
        return netspec.repnet_ops()

    @torch.no_grad()
    def forward(self, input_grays):
        out = self._run_on({"gray": input_grays.float().contiguous()}, "pred_feats")
        return out.permute(0, 3, 1, 2).float().contiguous()


class HourGlass2(_ConvNet):
    """reference models/network.py:125-144 with inChannel=65, outChannel=2: (N,65,H,W) -> (N,2,H,W), the pre-tanh map
    exactly as the reference's `HourGlass2.forward` returns it (the tanh belongs to `AnchorColorProb.forward`,
    models/model.py:197, where it is fused into the same conv's epilogue as the `tanh2` head)."""
    _PREFIX, _NET = "enhanceNet.", "enhanceNet"

    def __init__(self, inChannel=3, outChannel=1, resNum=3, normLayer=None):
        super().__init__()
        if inChannel != 65 or outChannel != 2 or resNum != 3 or normLayer is not nn.BatchNorm2d:
            raise _lib.DiscoError("HourGlass2 is built for inChannel=65, outChannel=2, resNum=3, "
                                  "normLayer=nn.BatchNorm2d (models/model.py:44)")
        tree = ParamTree(netspec.schema(nets=("enhanceNet",)), self._PREFIX)
        for name, child in list(tree.named_children()):
            self.add_module(name, child)
        _init_from_synth(self, self._PREFIX)

    def _ops(self):
        ops = netspec.enhancenet_ops()
        assert ops[-1].head == "tanh2"
        ops[-1].head = "raw2"            # standalone: no tanh (DISCO_HEAD_RAW2)
        return ops

    @torch.no_grad()
    def forward(self, x):
        from .engine import _DT
        gray = x[:, :1].float().contiguous()
        feats = x[:, 1:].permute(0, 2, 3, 1).contiguous().to(_DT[self.precision][1])
        return self._run_on({"gray": gray, "full_feats": feats}, "pred_colors")
