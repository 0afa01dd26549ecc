"""ctypes binding of libdisco_b200.so (the C ABI declared in include/disco_b200.h).

There is no fallback: if the library is missing or the device is not a B200 the import of the
compute path raises.  Set DISCO_B200_BUILD=1 to (re)build with nvcc on first use.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdisco_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
HEAD_NONE, HEAD_SOFTMAX9, HEAD_TANH2, HEAD_RAW2 = 0, 1, 2, 3
CONV3, DECONV4 = 0, 1

EXPORTS = ["disco_version", "disco_abi_size", "disco_last_error", "disco_create", "disco_destroy", "disco_launch_count",
           "disco_reset_launch_count", "disco_add_launch_count", "disco_conv", "disco_poolfeat", "disco_upfeat", "disco_linear",
           "disco_attention", "disco_kmeans_anchor", "disco_token_labels", "disco_set_tensor_core",
           "disco_conv_tc_supported", "disco_conv_tc_weight_elems", "disco_conv_tc_pack_weights",
           "disco_debug_timeline", "disco_token_sample3", "disco_encoder_tail", "disco_conv_tc_cache_clear", "disco_host_choice_rows",
           "disco_encoder_stack", "disco_encoder_stack_pack", "disco_encoder_stack_scratch_elems", "disco_segnet_head", "disco_lab2rgb_u8", "disco_ce_rebalance",
           "disco_spixel_recon_loss", "disco_encode_ab2ind", "disco_host_png_bound", "disco_host_png_encode", "disco_host_png_write",
           "disco_lab2rgb_norm", "disco_rgb_norm", "disco_maxpool2", "disco_l1_mean", "disco_laplace_l1", "disco_spixel_ids"]


class ConvSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("up2", C.c_int32),
                ("is_f32", C.c_int32), ("w_off", C.c_int64)]


class ConvDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("stride", C.c_int32), ("dtype", C.c_int32), ("batch", C.c_int32),
                ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32), ("n_src", C.c_int32),
                ("src", ConvSrc * 2), ("weights", C.c_void_p), ("bias", C.c_void_p), ("post_scale", C.c_void_p),
                ("post_shift", C.c_void_p), ("residual", C.c_void_p), ("act", C.c_int32), ("slope", C.c_float),
                ("head", C.c_int32), ("out", C.c_void_p), ("gray_weights", C.c_void_p),
                ("bias_host", C.c_void_p), ("post_scale_host", C.c_void_p), ("post_shift_host", C.c_void_p),
                ("gray_weights_host", C.c_void_p)]


class LinearDesc(C.Structure):
    _fields_ = [("X", C.c_void_p), ("W", C.c_void_p), ("b", C.c_void_p), ("M", C.c_int32), ("N", C.c_int32),
                ("K", C.c_int32), ("pos", C.c_void_p), ("pos_cols", C.c_int32), ("S", C.c_int32),
                ("col_scale", C.c_float), ("scale_cols", C.c_int32), ("relu", C.c_int32), ("residual", C.c_void_p),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("hint_mask", C.c_void_p), ("labels", C.c_void_p),
                ("emb", C.c_void_p), ("transpose_S", C.c_int32), ("Y", C.c_void_p)]


class EncoderLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("in_w", "in_b", "out_w", "out_b", "l1_w", "l1_b", "l2_w", "l2_b",
                                          "n1_w", "n1_b", "n2_w", "n2_b")]


_lib = None


class DiscoError(RuntimeError):
    pass


def load():
    """Loads the shared library (building it first when DISCO_B200_BUILD=1).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("DISCO_B200_BUILD") == "1":
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise DiscoError(f"{LIB_PATH} not found: build it with `python -m disentangledcolorization_b200.build` "
                         "(there is no CPU / PyTorch fallback for the compute path)")
    lib = C.CDLL(LIB_PATH)
    lib.disco_last_error.restype = C.c_char_p
    lib.disco_launch_count.restype = C.c_int64
    lib.disco_launch_count.argtypes = [C.c_void_p]
    lib.disco_reset_launch_count.argtypes = [C.c_void_p]
    lib.disco_add_launch_count.argtypes = [C.c_void_p, C.c_int64]
    lib.disco_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.disco_destroy.argtypes = [C.c_void_p]
    lib.disco_conv.argtypes = [C.c_void_p, C.POINTER(ConvDesc), C.c_void_p]
    lib.disco_set_tensor_core.argtypes = [C.c_void_p, C.c_int]
    lib.disco_conv_tc_supported.argtypes = [C.c_void_p, C.POINTER(ConvDesc)]
    lib.disco_conv_tc_weight_elems.argtypes = [C.POINTER(ConvDesc)]
    lib.disco_conv_tc_weight_elems.restype = C.c_int64
    lib.disco_conv_tc_pack_weights.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p]
    lib.disco_conv_tc_cache_clear.argtypes = [C.c_void_p]
    lib.disco_encoder_stack_scratch_elems.argtypes = [C.c_int, C.c_int]
    lib.disco_encoder_stack_scratch_elems.restype = C.c_int64
    lib.disco_encoder_stack_pack.argtypes = [C.POINTER(EncoderLayerWeights), C.c_int, C.c_void_p, C.c_void_p]
    lib.disco_encoder_stack.argtypes = [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p] * 3
    lib.disco_segnet_head.argtypes = [C.c_void_p] * 8 + [C.c_float] + [C.c_int] * 3 + [C.c_void_p] * 3
    lib.disco_lab2rgb_u8.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p] * 2
    lib.disco_encode_ab2ind.argtypes = [C.c_void_p] * 3 + [C.c_int] * 2 + [C.c_void_p] * 2
    lib.disco_ce_rebalance.argtypes = [C.c_void_p] * 4 + [C.c_int] * 2 + [C.c_void_p] * 4
    lib.disco_spixel_recon_loss.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.disco_host_choice_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int32)] + [C.c_int] * 5 + [C.c_void_p]
    lib.disco_lab2rgb_norm.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p] * 3
    lib.disco_rgb_norm.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p] + [C.c_int] * 2 + [C.c_void_p] * 3
    lib.disco_maxpool2.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2
    lib.disco_l1_mean.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p]
    lib.disco_laplace_l1.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 3
    lib.disco_spixel_ids.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2
    lib.disco_host_png_bound.argtypes = [C.c_int, C.c_int]
    lib.disco_host_png_bound.restype = C.c_longlong
    lib.disco_host_png_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]
    lib.disco_host_png_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong]
    lib.disco_poolfeat.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 6
    lib.disco_upfeat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2
    lib.disco_linear.argtypes = [C.c_void_p, C.POINTER(LinearDesc), C.c_void_p]
    lib.disco_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.disco_kmeans_anchor.argtypes = ([C.c_void_p] * 4 + [C.c_int, C.c_void_p] + [C.c_int] * 4 + [C.c_float]
                                        + [C.c_void_p] * 5)
    lib.disco_token_labels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    lib.disco_token_sample3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    lib.disco_encoder_tail.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 11
    _lib = lib
    return lib


def check(rc, what="disco call"):
    if rc != 0:
        raise DiscoError(f"{what} failed ({rc}): {load().disco_last_error().decode()}")


class Handle:
    """One library handle per (process, device)."""
    _cache = {}

    def __init__(self, device_index):
        lib = load()
        h = C.c_void_p()
        check(lib.disco_create(C.byref(h), int(device_index)), "disco_create")
        self.lib, self.h, self.device_index = lib, h, int(device_index)

    @classmethod
    def get(cls, device_index):
        if device_index not in cls._cache:
            cls._cache[device_index] = cls(device_index)
        return cls._cache[device_index]

    def launches(self):
        return int(self.lib.disco_launch_count(self.h))

    def reset_launches(self):
        self.lib.disco_reset_launch_count(self.h)


def choice_rows(n_tokens, n_clusters, rows, keep=None):
    """`rows` x np.random.choice(n_tokens, n_clusters, replace=False) on numpy's global generator, in one native call
    (disco_host_choice_rows): same numbers, same final generator state, ~15x less host time.  Returns int32 (n_keep, K)
    for rows keep=(lo, hi) (default: all).  Falls back to the python loop for a non-MT19937 global generator."""
    import numpy as np
    lo, hi = (0, rows) if keep is None else keep
    st = np.random.get_state()
    if st[0] != "MT19937":
        allidx = [np.random.choice(n_tokens, n_clusters, replace=False) for _ in range(rows)]
        return np.stack(allidx[lo:hi]).astype(np.int32) if hi > lo else np.zeros((0, n_clusters), np.int32)
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(st[2]))
    out = np.empty((hi - lo, n_clusters), np.int32)
    check(load().disco_host_choice_rows(C.c_void_p(key.ctypes.data), C.byref(pos), int(n_tokens), int(n_clusters), int(rows),
                                        int(lo), int(hi), C.c_void_p(out.ctypes.data)), "disco_host_choice_rows")
    np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))
    return out


def _rgb_u8(rgb):
    import numpy as np
    if rgb.dtype != np.uint8 or rgb.ndim != 3 or rgb.shape[2] != 3 or rgb.strides[2] != 1 or rgb.strides[1] != 3:
        raise ValueError("expected an (H, W, 3) uint8 array with contiguous rows")
    return rgb


def png_encode(rgb):
    """(H, W, 3) uint8 RGB -> the bytes of a PNG file (disco_host_png_encode; host only, the GIL is released)."""
    import numpy as np
    rgb = _rgb_u8(rgb)
    lib = load()
    H, W = int(rgb.shape[0]), int(rgb.shape[1])
    cap = int(lib.disco_host_png_bound(H, W))
    buf = np.empty(cap, np.uint8)
    n = C.c_longlong(0)
    check(lib.disco_host_png_encode(C.c_void_p(rgb.ctypes.data), H, W, int(rgb.strides[0]), C.c_void_p(buf.ctypes.data), cap, C.byref(n)),
          "disco_host_png_encode")
    return buf[:n.value].tobytes()


def png_write(path, rgb):
    """Encode and write one PNG file (disco_host_png_write), replacing Image.fromarray(rgb).save(path) of utils/util.py:106."""
    rgb = _rgb_u8(rgb)
    check(load().disco_host_png_write(os.fsencode(path), C.c_void_p(rgb.ctypes.data), int(rgb.shape[0]), int(rgb.shape[1]),
                                      int(rgb.strides[0])), "disco_host_png_write")
