"""Drop-in `model` module: SpixelSeg and AnchorColorProb with the reference's constructor
signature, forward contract, helper methods and state_dict schema (reference models/model.py:12-199).

    model = AnchorColorProb(inChannel=1, outChannel=313, sp_size=16, d_model=64, use_dense_pos=True,
                            n_clusters=8, enhanced=True).cuda()
    load_checkpoint(path, model)            # strict load_state_dict, same 461 keys
    model.eval()
    pal_logit, ref_logit, pred_colors, affinity_map, spix_colors, hint_mask = \
        model(input_grays, input_colors, True, 0)

All arithmetic runs in libdisco_b200.so (hand-written sm_100a kernels); there is no PyTorch or CPU
fallback -- a missing library or a non-CUDA tensor raises.
"""
import torch
import torch.nn as nn

from . import _lib, netspec
from .network import ParamTree, SpixelNet, ColorProbNet, HourGlass2


class SpixelSeg(nn.Module):
    """reference models/model.py:12-29."""

    def __init__(self, inChannel=1, outChannel=9, batchNorm=True):
        super().__init__()
        self.net = SpixelNet(inChannel=inChannel, outChannel=outChannel, batchNorm=batchNorm)

    def get_trainable_params(self, lr=1.0):
        return [{"params": p} for _, p in self.named_parameters()]

    def forward(self, input_grays):
        return self.net(input_grays)


class AnchorColorProb(nn.Module):
    """reference models/model.py:32-199 (inference / test_mode path)."""

    #: 'bf16' = tensor-core path (bf16 storage, fp32 accumulate); 'fp32' = exact CUDA-core path
    precision = "bf16"
    #: replay the forward from a CUDA graph (one graph per input shape; sampled_T == 0 only).  Off by default:
    #: the CLI feeds images of arbitrary sizes one at a time; batch serving with fixed shapes should turn it on.
    use_cuda_graph = False
    #: graph mode only: return the graph-owned output tensors (overwritten by the next forward) instead of copies
    graph_static_outputs = False
    #: [extension, SURVEY 8f N2] allow sampled_T > 0 (--diverse) on batches: the reference's expand (models/model.py:155-159)
    #: only works for N == 1; with this flag a batch of N returns 3N variants, variant-major (index v * N + n)
    batched_diverse = False
    #: defer the host-RNG fix-up (one tiny D2H read + sync) to the next forward / engine().sync_rng()
    lazy_rng = False

    def __init__(self, inChannel=1, outChannel=313, sp_size=16, d_model=64, use_dense_pos=True, spix_pos=False,
                 learning_pos=False, n_clusters=8, random_hint=False, hint2regress=False, enhanced=False,
                 use_mask=False, rank=0, precision=None):
        super().__init__()
        unsupported = []
        if inChannel != 1 or outChannel != 313 or d_model != 64:
            unsupported.append("inChannel/outChannel/d_model other than 1/313/64")
        if not use_dense_pos:
            unsupported.append("use_dense_pos=False (the inference CLI forces True, inference.py:165)")
        if spix_pos or hint2regress or use_mask:
            unsupported.append("spix_pos / hint2regress / use_mask")
        if unsupported:
            raise _lib.DiscoError("AnchorColorProb configuration not built: " + "; ".join(unsupported))
        self.sp_size = sp_size
        self.spix_pos = spix_pos
        self.use_token_mask = use_mask
        self.hint2regress = hint2regress
        self.enhanced = enhanced
        self.n_vocab = 313
        self.hint_num = n_clusters
        self.random_hint = random_hint
        if precision is not None:
            self.precision = precision
        self.segnet = SpixelSeg(inChannel=1, outChannel=9, batchNorm=True)
        self.repnet = ColorProbNet(inChannel=inChannel, outChannel=64)
        if enhanced:
            self.enhanceNet = HourGlass2(inChannel=64 + 1, outChannel=2, resNum=3, normLayer=nn.BatchNorm2d)
        tok = ParamTree(netspec.schema(nets=("tokens",)), "")
        for name, child in list(tok.named_children()):
            self.add_module(name, child)
        self._init_tokens()
        self._engine = None
        self._engine_key = None
        self.register_load_state_dict_post_hook(lambda m, keys: m._invalidate())

    def _init_tokens(self):
        from . import synth
        sd = synth.make_state_dict(seed=0)
        own = self.state_dict()
        self.load_state_dict({k: sd[k] for k in own}, strict=True)

    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    # --- reference helper methods -------------------------------------------------------------
    def load_and_froze_weight(self, checkpt_path):
        """reference models/model.py:78-87."""
        data_dict = torch.load(checkpt_path, map_location=torch.device("cpu"))
        self.segnet.load_state_dict(data_dict["state_dict"])
        for _, param in self.segnet.named_parameters():
            param.requires_grad = False
        self.segnet.eval()
        self._invalidate()

    def set_train(self):
        """reference models/model.py:89-95 (mode flags only; the training step is not built)."""
        self.repnet.train()
        self.wildpath.train()
        self.hintpath.train()
        if self.enhanced:
            self.enhanceNet.train()

    def get_entry_mask(self, mask_tensor):
        return None if mask_tensor is None else mask_tensor.flatten(1)

    # --- engine -------------------------------------------------------------------------------
    def engine(self, device=None):
        from .engine import Engine
        if device is None:
            device = next(self.parameters()).device
        key = (str(device), self.precision, self.hint_num, self.random_hint)
        if self._engine is None or self._engine_key != key:
            self._engine = Engine(self.state_dict(), device, precision=self.precision, n_clusters=self.hint_num,
                                  sp_size=self.sp_size, enhanced=self.enhanced, random_hint=self.random_hint)
            self._engine_key = key
        self._engine.use_graph = self.use_cuda_graph
        self._engine.static_outputs = self.graph_static_outputs
        self._engine.lazy_rng = self.lazy_rng
        self._engine.batched_diverse = self.batched_diverse
        return self._engine

    def forward(self, input_grays, input_colors, test_mode=False, sampled_T=0, hint_mask=None, init_idx=None):
        """Same contract as the reference forward (models/model.py:103,199).  `hint_mask` (extension):
        inject anchor sites instead of running k-means; `init_idx` (extension): k-means init rows for this
        shard when the batch is split over ranks (dist.sharded_init_draws)."""
        if not test_mode:
            raise _lib.DiscoError("the training branch (test_mode=False) is not built; SURVEY.md section 8f N3")
        if self.training:
            raise _lib.DiscoError("call model.eval() first: only the eval-mode forward is built")
        return self.engine(input_grays.device).forward(input_grays, input_colors, sampled_T=sampled_T,
                                                       hint_mask=hint_mask, init_idx=init_idx)
