"""ctypes binding of include/xdtts_b200.h (the same symbols a Rust/cgo/JNI shim would bind)."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libxdtts_b200.so")

OK, ERR_BAD_ARG, ERR_SHAPE, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
RUN_FROM_MAG, RUN_USE_PHASE, RUN_NO_GRAPH, RUN_PER_LAUNCH = 1, 2, 4, 16

_CODE_NAMES = {
    ERR_BAD_ARG: "BAD_ARG", ERR_SHAPE: "SHAPE", ERR_CUDA: "CUDA", ERR_OOM: "OOM", ERR_UNSUPPORTED: "UNSUPPORTED",
}


class XdttsError(RuntimeError):
    """Raised for every non-zero return of the C ABI (the shim's anyhow::bail!)."""

    def __init__(self, code, message):
        super().__init__("xdtts_b200 error %s (%d): %s" % (_CODE_NAMES.get(code, "?"), code, message))
        self.code = code
        self.message = message


class PostnetOpts(ctypes.Structure):
    _fields_ = [("precision", ctypes.c_int)]


class DecoderWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.POINTER(ctypes.c_float)) for n in (
        "prenet1", "prenet2", "att_w_ih", "att_w_hh", "att_b_ih", "att_b_hh", "query", "v", "loc_conv", "loc_dense",
        "dec_w_ih", "dec_w_hh", "dec_b_ih", "dec_b_hh", "proj_w", "proj_b", "gate_w", "gate_b")]


class DecoderOpts(ctypes.Structure):
    _fields_ = [("gate_threshold", ctypes.c_float), ("max_steps", ctypes.c_int), ("prenet_dropout", ctypes.c_int),
                ("seed", ctypes.c_ulonglong)]


class GlOpts(ctypes.Structure):
    _fields_ = [
        ("delog", ctypes.c_int),
        ("pad_mode", ctypes.c_int),
        ("normalise", ctypes.c_int),
        ("run_frames", ctypes.c_int),
        ("seed", ctypes.c_ulonglong),
        ("persistent", ctypes.c_int),
        ("lift", ctypes.c_int),
        ("nnls_iters", ctypes.c_int),
        ("fixed_seed", ctypes.c_int),
        ("exponent", ctypes.c_int),
    ]


_fp = ctypes.POINTER(ctypes.c_float)
_fpp = ctypes.POINTER(_fp)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p
_sp = ctypes.POINTER(ctypes.c_short)
_spp = ctypes.POINTER(_sp)

# every symbol include/xdtts_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "xdtts_mel_filter_bank": (ctypes.c_int, [ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, _fp]),
    "xdtts_pinv": (ctypes.c_int, [_fp, ctypes.c_int, ctypes.c_int, _fp]),
    "xdtts_gl_create": (ctypes.c_int, [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       ctypes.c_float, ctypes.POINTER(GlOpts), ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_gl_destroy": (None, [_vp]),
    "xdtts_gl_out_len": (ctypes.c_int, [_vp, ctypes.c_int]),
    "xdtts_gl_get_pinv": (ctypes.c_int, [_vp, _fp]),
    "xdtts_gl_infer": (ctypes.c_int, [_vp, _fp, ctypes.c_int, _fp, _fp, ctypes.c_int]),
    "xdtts_gl_infer_batch": (ctypes.c_int, [_vp, _fpp, _ip, ctypes.c_int, _fpp, _fpp]),
    "xdtts_gl_from_mag_batch": (ctypes.c_int, [_vp, _fpp, _ip, ctypes.c_int, _fpp, _fpp]),
    "xdtts_gl_infer_batch_pcm16": (ctypes.c_int, [_vp, _fpp, _ip, ctypes.c_int, _fpp, _spp]),
    "xdtts_gl_plan_download_pcm16": (ctypes.c_int, [_vp, _spp]),
    "xdtts_gl_plan_lift_ms": (ctypes.c_int, [_vp, _fp]),
    "xdtts_gl_plan_time_lift": (ctypes.c_int, [_vp, ctypes.c_int, _fp]),
    "xdtts_pool_create": (ctypes.c_int, [_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                         ctypes.c_float, ctypes.POINTER(GlOpts), _ip, ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_pool_destroy": (None, [_vp]),
    "xdtts_pool_n_devices": (ctypes.c_int, [_vp]),
    "xdtts_pool_out_len": (ctypes.c_int, [_vp, ctypes.c_int]),
    "xdtts_shard_assign": (ctypes.c_int, [_ip, ctypes.c_int, ctypes.c_int, _ip]),
    "xdtts_pool_assignment": (ctypes.c_int, [_vp, _ip, ctypes.c_int, _ip]),
    "xdtts_pool_infer_batch": (ctypes.c_int, [_vp, _fpp, _ip, ctypes.c_int, _fpp, _fpp]),
    "xdtts_pool_from_mag_batch": (ctypes.c_int, [_vp, _fpp, _ip, ctypes.c_int, _fpp, _fpp]),
    "xdtts_gl_plan_create": (ctypes.c_int, [_vp, _ip, ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_gl_plan_destroy": (None, [_vp]),
    "xdtts_gl_plan_upload": (ctypes.c_int, [_vp, ctypes.c_int, _fpp]),
    "xdtts_gl_plan_run": (ctypes.c_int, [_vp, ctypes.c_int, _fp, _fp, _ip]),
    "xdtts_gl_plan_download": (ctypes.c_int, [_vp, _fpp]),
    "xdtts_gl_plan_peek": (ctypes.c_int, [_vp, ctypes.c_int, _fp, ctypes.c_longlong]),
    "xdtts_gl_plan_is_persistent": (ctypes.c_int, [_vp]),
    "xdtts_gl_plan_info": (ctypes.c_int, [_vp, _ip]),
    "xdtts_postnet_create": (ctypes.c_int, [ctypes.c_int, _ip, ctypes.c_int, _fpp, _fpp, _fpp, _fpp, _fpp, _fpp, ctypes.c_float,
                                            ctypes.POINTER(PostnetOpts), ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_postnet_destroy": (None, [_vp]),
    "xdtts_postnet_infer": (ctypes.c_int, [_vp, _fp, ctypes.c_int, _fp]),
    "xdtts_postnet_infer_batch": (ctypes.c_int, [_vp, _fpp, _ip, ctypes.c_int, _fpp]),
    "xdtts_postnet_plan_create": (ctypes.c_int, [_vp, _ip, ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_postnet_plan_destroy": (None, [_vp]),
    "xdtts_postnet_plan_upload": (ctypes.c_int, [_vp, _fpp]),
    "xdtts_postnet_plan_run": (ctypes.c_int, [_vp, _vp, _fp]),
    "xdtts_postnet_plan_download": (ctypes.c_int, [_vp, _fpp]),
    "xdtts_postnet_create_from_onnx": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(PostnetOpts), ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_onnx_postnet_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "xdtts_onnx_postnet_close": (None, [_vp]),
    "xdtts_onnx_postnet_n_layers": (ctypes.c_int, [_vp]),
    "xdtts_onnx_postnet_layer_info": (ctypes.c_int, [_vp, ctypes.c_int, _ip, _ip, _ip, _ip, _ip, _fp]),
    "xdtts_onnx_postnet_layer_copy": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _fp]),
    "xdtts_tail_infer_batch": (ctypes.c_int, [_vp, _vp, _fpp, _ip, ctypes.c_int, _fpp, _fpp, _fpp]),
    "xdtts_decoder_create_from_onnx": (ctypes.c_int, [ctypes.c_char_p, _vp, ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_onnx_decoder_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "xdtts_onnx_decoder_close": (None, [_vp]),
    "xdtts_onnx_decoder_dims": (ctypes.c_int, [_vp, _ip]),
    "xdtts_onnx_decoder_tensor": (ctypes.c_longlong, [_vp, ctypes.c_int, _fp, ctypes.c_longlong]),
    "xdtts_decoder_create": (ctypes.c_int, [ctypes.POINTER(DecoderWeights), ctypes.POINTER(DecoderOpts), ctypes.c_int,
                                            ctypes.POINTER(_vp)]),
    "xdtts_decoder_destroy": (None, [_vp]),
    "xdtts_decoder_max_steps": (ctypes.c_int, [_vp]),
    "xdtts_decoder_infer_batch": (ctypes.c_int, [_vp, _fpp, _fpp, ctypes.c_int, _ip, ctypes.c_int, _fpp, _ip, _fpp, _fpp]),
    "xdtts_decoder_last_timing": (ctypes.c_int, [_vp, _fp, _ip]),
    "xdtts_decoder_info": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]),
    "xdtts_pipe_create": (ctypes.c_int, [_vp, _vp, _ip, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    "xdtts_pipe_push": (ctypes.c_int, [_vp, _fpp, _fpp, _fpp, _fpp]),
    "xdtts_pipe_pop": (ctypes.c_int, [_vp]),
    "xdtts_pipe_flush": (ctypes.c_int, [_vp]),
    "xdtts_pipe_pending": (ctypes.c_int, [_vp]),
    "xdtts_pipe_destroy": (None, [_vp]),
    "xdtts_npy_write_f32": (ctypes.c_int, [ctypes.c_char_p, _fp, ctypes.c_int, ctypes.c_int]),
    "xdtts_npy_read_f32": (ctypes.c_int, [ctypes.c_char_p, _fp, ctypes.c_longlong, _ip, _ip]),
    "xdtts_host_alloc": (_vp, [ctypes.c_ulonglong]),
    "xdtts_host_free": (None, [_vp]),
    "xdtts_last_error": (ctypes.c_char_p, []),
    "xdtts_kernel_launches": (ctypes.c_ulonglong, []),
    "xdtts_version": (ctypes.c_char_p, []),
}

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("XDTTS_B200_LIB", LIB_PATH)   # tuning builds (tools/build_variants.py) override the path
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first; there is no CPU fallback" % path
            )
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise XdttsError(rc, load_library().xdtts_last_error().decode("utf-8", "replace"))
    return rc


def version():
    return load_library().xdtts_version().decode()


def fptr(a):
    return a.ctypes.data_as(_fp)


def sptr_array(arrays):
    arr = (_sp * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data_as(_sp)
    return arr


def fptr_array(arrays):
    arr = (_fp * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = fptr(a)
    return arr
