"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference on CPU.

Only usable in the authoring container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to pin oracle/disco_oracle.py against the real reference code and to
generate the committed fixtures under tests/golden/.  Nothing in the product package, the
`-m gpu` tests, smoke() or bench.py imports this file.

Three harness-side shims (reference files are untouched):
  1. stub modules for matplotlib / skimage / tensorboardX (utils/cielab.py:2, utils/util.py:6,
     main/colorizer/inference.py:20 import them but the forward path never calls them);
  2. CPU only: torch.Tensor.cuda -> identity and ColorLabel(device='cpu')
     (models/model.py:68,122; models/basic.py:267,284,330,332 hard-code .cuda());
  3. chdir to <ref>/main/colorizer so '../../utils/gamut_*.npy' resolves (utils/cielab.py:6-7),
     and put <ref>, <ref>/main, <ref>/models, <ref>/utils on sys.path (main/_init_paths.py:10-13).
"""
import os
import sys
import types

REF = os.environ.get("DISCO_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models"))


_loaded = None


def load():
    """Returns the reference `model` module (models/model.py) importable on CPU."""
    global _loaded
    if _loaded is not None:
        return _loaded
    import torch

    for name in ["matplotlib", "matplotlib.pyplot", "skimage", "skimage.segmentation", "skimage.color",
                 "tensorboardX"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["skimage.segmentation"].mark_boundaries = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].segmentation = sys.modules["skimage.segmentation"]
    sys.modules["skimage"].color = sys.modules["skimage.color"]

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "main", "colorizer"))
    for p in [REF, os.path.join(REF, "main"), os.path.join(REF, "models"), os.path.join(REF, "utils")]:
        if p not in sys.path:
            sys.path.append(p)
    try:
        import basic as ref_basic  # noqa
        import model as ref_model  # noqa

        if not torch.cuda.is_available():
            _orig = ref_basic.ColorLabel.__init__

            def _cpu_init(self, lambda_=0.5, device="cpu"):
                _orig(self, lambda_, torch.device("cpu"))

            ref_basic.ColorLabel.__init__ = _cpu_init
    finally:
        os.chdir(cwd)
    _loaded = ref_model
    return ref_model


def build_model(n_clusters=8, **kw):
    """AnchorColorProb constructed exactly as main/colorizer/inference.py:71-74,165 does."""
    ref_model = load()
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "main", "colorizer"))
    try:
        m = ref_model.AnchorColorProb(inChannel=1, outChannel=313, sp_size=16, d_model=64, use_dense_pos=True,
                                      spix_pos=False, learning_pos=False, n_clusters=n_clusters,
                                      random_hint=kw.get("random_hint", False), hint2regress=False, enhanced=True)
    finally:
        os.chdir(cwd)
    return m
