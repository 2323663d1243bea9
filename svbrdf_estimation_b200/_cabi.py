"""ctypes binding of ``libsvbrdf_b200.so`` (C ABI declared in ``include/svbrdf_b200.h``).

There is deliberately no fallback: if the shared library cannot be loaded the import of any
product entry point raises, and every non-zero status from the library becomes an exception.
"""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# SVBRDF_B200_LIB points at an alternative build of the same ABI (variant builds of scripts/variant_bench.py)
LIB_PATH = os.environ.get("SVBRDF_B200_LIB") or os.path.join(_PKG, "libsvbrdf_b200.so")

E_INVALID, E_TOO_LARGE, E_STATE = -1, -2, -3
ABI_VERSION = 2

_c_float_p = ctypes.c_void_p   # raw addresses (tensor.data_ptr()) are passed as integers
_c_stream = ctypes.c_void_p

# name -> (restype, argtypes); kept in the order of include/svbrdf_b200.h
PROTOTYPES = {
    "svbrdf_b200_abi_version": (ctypes.c_int, []),
    "svbrdf_b200_last_error": (ctypes.c_char_p, []),
    "svbrdf_b200_build_id": (ctypes.c_char_p, []),
    "svbrdf_b200_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 4),
    "svbrdf_b200_coordinate_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "svbrdf_b200_sample_scenes": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_void_p]),
    "svbrdf_b200_reference_draws": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_void_p, ctypes.c_void_p]),
    "svbrdf_b200_reference_scenes_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "svbrdf_b200_reference_scenes_finish": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7),
    "svbrdf_b200_render_forward": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                  ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _c_stream]),
    "svbrdf_b200_render_backward": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                   ctypes.c_int, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p,
                                                   _c_stream]),
    "svbrdf_b200_loss_forward": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                _c_float_p, ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_void_p,
                                                ctypes.c_size_t, _c_stream]),
    "svbrdf_b200_loss_forward_backward": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_int, _c_float_p, ctypes.c_int, _c_float_p,
                                                         _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_size_t,
                                                         _c_stream]),
    "svbrdf_b200_loss_forward_backward_accurate": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                                  ctypes.c_int, _c_float_p, ctypes.c_int, _c_float_p,
                                                                  _c_float_p, _c_float_p, ctypes.c_void_p, ctypes.c_size_t,
                                                                  _c_stream]),
    "svbrdf_b200_loss_forward_accurate": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                         _c_float_p, ctypes.c_int, _c_float_p, _c_float_p, ctypes.c_void_p,
                                                         ctypes.c_size_t, _c_stream]),
    "svbrdf_b200_scale_grad": (ctypes.c_int, [_c_float_p, ctypes.c_size_t, _c_float_p, _c_stream]),
    "svbrdf_b200_mixed_loss_forward_backward": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                               ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_float,
                                                               _c_float_p, _c_float_p, _c_float_p, ctypes.c_void_p,
                                                               ctypes.c_size_t, _c_stream]),
    "svbrdf_b200_mixed_loss_encoded_forward_backward": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                                       ctypes.c_int, _c_float_p, ctypes.c_int,
                                                                       ctypes.c_float, _c_float_p, _c_float_p, _c_float_p,
                                                                       ctypes.c_void_p, ctypes.c_size_t, _c_stream]),
    "svbrdf_b200_loss_layouts": (ctypes.c_int, [_c_float_p, ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_float, _c_float_p, _c_float_p,
                                                _c_float_p, ctypes.c_void_p, ctypes.c_size_t, _c_stream]),
    "svbrdf_b200_ctx_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int]),
    "svbrdf_b200_ctx_destroy": (None, [ctypes.c_void_p]),
    "svbrdf_b200_ctx_pinned": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "svbrdf_b200_rendering_loss_host": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_int,
                                                       _c_float_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                                       _c_float_p]),
    "svbrdf_b200_loss_host": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_int,
                                             _c_float_p, ctypes.c_int, ctypes.c_float, _c_float_p, _c_float_p]),
}
LAYOUT_MAPS12, LAYOUT_MAPS10, LAYOUT_ENCODED9 = 12, 10, 9

_lib = None
_lock = threading.Lock()


class SvbrdfB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__("svbrdf_b200 status %d: %s" % (status, message))
        self.status = status


def lib():
    """The loaded library.  The in-tree binary carries a hash of the sources it was built from
    (``svbrdf_b200_build_id``); if it is missing or was built from other sources it is rebuilt when nvcc is
    available, otherwise loading fails - a stale kernel is never served silently."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.environ.get("SVBRDF_B200_LIB"):
            from . import _build
            if _build.is_stale():
                try:
                    _build.find_nvcc()
                except RuntimeError:
                    raise RuntimeError("%s is %s and nvcc is not available to rebuild it" % (
                        LIB_PATH, "missing" if not os.path.exists(LIB_PATH) else "older than the CUDA sources"))
                _build.build()          # raises if the build fails: no silent fallback
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)   # AttributeError if the .so does not export the header's symbol
            fn.restype, fn.argtypes = restype, argtypes
        if handle.svbrdf_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("libsvbrdf_b200.so has ABI version %d, expected %d" % (handle.svbrdf_b200_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        msg = lib().svbrdf_b200_last_error()
        raise SvbrdfB200Error(status, msg.decode("utf-8", "replace") if msg else "")
