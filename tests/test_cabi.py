"""CPU: the C-ABI library builds, loads, and exports every symbol include/disco_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "disco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(disco_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from disentangledcolorization_b200 import build, _lib
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in disco_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(declared)
    lib.disco_version.restype = ctypes.c_int
    assert lib.disco_version() >= 100


def test_struct_layout_matches_header():
    """ctypes mirrors of the descriptor structs have the C layout (sizes from the compiler's rules)."""
    from disentangledcolorization_b200 import _lib
    assert ctypes.sizeof(_lib.ConvSrc) == 40
    assert ctypes.sizeof(_lib.ConvDesc) == 32 + 80 + 5 * 8 + 16 + 8 + 8 + 4 * 8
    assert _lib.ConvDesc.src.offset == 32 and _lib.ConvDesc.out.offset == 168
    assert ctypes.sizeof(_lib.LinearDesc) == 136
    # ... and equal what the compiler laid out (the library describes its own ABI; no GPU needed)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    assert lib.disco_abi_size(0) == ctypes.sizeof(_lib.ConvSrc)
    assert lib.disco_abi_size(1) == ctypes.sizeof(_lib.ConvDesc)
    assert lib.disco_abi_size(2) == ctypes.sizeof(_lib.LinearDesc)
    assert lib.disco_abi_size(3) == _lib.ConvDesc.out.offset
    assert lib.disco_abi_size(4) == _lib.ConvDesc.bias_host.offset


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    from disentangledcolorization_b200 import model, _lib
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = model.AnchorColorProb(n_clusters=8, enhanced=True).eval()
    with pytest.raises(_lib.DiscoError):
        m(torch.zeros(1, 1, 32, 32), torch.zeros(1, 2, 32, 32), True, 0)


def test_loss_and_helper_modules_fail_loudly_without_cuda():
    """The training-side terms and the super-pixel helpers have no CPU path either."""
    import pytest
    import torch
    from disentangledcolorization_b200 import basic, loss, _lib
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(_lib.DiscoError):
        basic.lab2rgb(torch.zeros(1, 3, 16, 16))
    with pytest.raises(_lib.DiscoError):
        basic.split_spixels(torch.full((1, 9, 16, 16), 1.0 / 9))
    with pytest.raises(_lib.DiscoError):
        loss.AnchorColorProbLoss()._laplace_gradient(torch.zeros(1, 2, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(_lib.DiscoError):
        loss.SPixelLoss(psize=16)({"pred_prob": torch.full((1, 9, 16, 16), 1.0 / 9), "target_feat": torch.zeros(1, 5, 16, 16)}, 0)
