"""ctypes binding of libpwc_b200.so (include/pwc_b200.h).  No torch types cross this layer:
only raw device pointers, sizes, dtype codes and a cudaStream_t.

There is deliberately no CPU or PyTorch fallback: if the library is missing or the device is not
a B200 the call fails loudly."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PWC_B200_LIB") or os.path.join(_HERE, "_lib", "libpwc_b200.so")

PWC_F32, PWC_F16, PWC_BF16 = 0, 1, 2
ABI_VERSION = 1


class PwcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpwc_b200 error {code}: {msg}")
        self.code = code


class LevelT(ctypes.Structure):  # pwc_level_t
    _fields_ = [("c1", ctypes.c_void_p), ("c2", ctypes.c_void_p), ("flow", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int), ("C", ctypes.c_int),
                ("flow_scale", ctypes.c_float), ("out_pixel_stride", ctypes.c_int64)]


_vp, _i, _i64, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
# name -> (restype, argtypes); every symbol include/pwc_b200.h declares
SIGNATURES = {
    "pwc_abi_version": (_i, []),
    "pwc_last_error": (ctypes.c_char_p, []),
    "pwc_launch_count": (ctypes.c_uint64, []),
    "pwc_set_option": (_i, [ctypes.c_char_p, _i]),
    "pwc_get_option": (_i, [ctypes.c_char_p, ctypes.POINTER(_i)]),
    "pwc_debug_read_trace": (_i, [_vp, _i]),
    "pwc_cost_volume_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i64, _vp]),
    "pwc_dense_image_warp_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pwc_warp_cost_volume_fwd": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _i, _i, _i64, _vp]),
    "pwc_pyramid_fwd": (_i, [ctypes.POINTER(LevelT), _i, _i, _i, _i, _i, _vp]),
    "pwc_host_ctx_create": (_i, [ctypes.POINTER(_vp)]),
    "pwc_host_ctx_destroy": (_i, [_vp]),
    "pwc_pyramid_fwd_host": (_i, [_vp, ctypes.POINTER(LevelT), _i, _i, _i, _i, _i]),
    "pwc_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "pwc_cost_volume_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i64, _i64, _vp]),
    "pwc_dense_image_warp_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "pwc_warp_cost_volume_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i,
                                      _i64, _i64, _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None


def load(path: str | None = None):
    """Loads the shared library and declares every prototype.  Raises if it is missing."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise RuntimeError(
                f"{p} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (or python tfoptflow_b200/csrc/build.py). There is no CPU fallback.")
        lib = ctypes.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        v = lib.pwc_abi_version()
        if v != ABI_VERSION:
            raise RuntimeError(f"libpwc_b200 ABI version {v}, binding expects {ABI_VERSION}")
        if path is None:
            _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        raise PwcError(rc, load().pwc_last_error().decode("utf-8", "replace"))


def set_option(key: str, value: int):
    check(load().pwc_set_option(key.encode(), int(value)))


def get_option(key: str) -> int:
    v = ctypes.c_int(0)
    check(load().pwc_get_option(key.encode(), ctypes.byref(v)))
    return v.value


def launch_count() -> int:
    return int(load().pwc_launch_count())
