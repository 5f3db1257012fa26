"""ctypes binding of libphdfx.so (C ABI: include/phdfx.h).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG_ROOT = Path(__file__).resolve().parent.parent  # implementation-phd-lab-vision_b200/
LIB_PATH = _PKG_ROOT / "lib" / "libphdfx.so"

# every symbol include/phdfx.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "phdfx_version",
    "phdfx_last_error",
    "phdfx_create",
    "phdfx_destroy",
    "phdfx_load_weights",
    "phdfx_preprocess_u8",
    "phdfx_preprocess_u8_jitter",
    "phdfx_extract_u8_jitter",
    "phdfx_nchw_f32_to_nhwc_bf16",
    "phdfx_forward",
    "phdfx_forward_timed",
    "phdfx_extract_u8",
    "phdfx_run_layer",
    "phdfx_run_layer2",
    "phdfx_chain_span",
    "phdfx_run_chain",
    "phdfx_set_schedule",
    "phdfx_get_schedule",
    "phdfx_layer_count",
    "phdfx_layer_info",
    "phdfx_last_launch_count",
    "phdfx_linked_launches",
]

PHDFX_CONV, PHDFX_STEM, PHDFX_MAXPOOL, PHDFX_STEM_POOL = 0, 1, 2, 3
IMG, IN_WPAD, IN_LPAD, IN_CPAD, FEAT_DIM = 224, 232, 4, 4, 2048
SCHED_REUSE = 1  # PHDFX_SCHED_REUSE
JITTER_FLOATS = 12  # one row of phdfx_preprocess_u8_jitter's parameter array


class LayerDesc(C.Structure):
    """Mirror of phdfx_layer_desc (include/phdfx.h)."""

    _fields_ = [
        ("kind", C.c_int32),
        ("cin", C.c_int32),
        ("cout", C.c_int32),
        ("r", C.c_int32),
        ("s", C.c_int32),
        ("stride", C.c_int32),
        ("pad", C.c_int32),
        ("hin", C.c_int32),
        ("win", C.c_int32),
        ("relu", C.c_int32),
        ("in_buf", C.c_int32),
        ("out_buf", C.c_int32),
        ("res_buf", C.c_int32),
        ("gap", C.c_int32),
        ("in2_buf", C.c_int32),
        ("cin2", C.c_int32),
        ("stride2", C.c_int32),
        ("hin2", C.c_int32),
        ("w_off", C.c_int64),
        ("b_off", C.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load libphdfx.so and declare its signatures.  Raises RuntimeError when the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("PHDFX_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"libphdfx.so not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for this backend)"
        )
    lib = C.CDLL(str(path))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.phdfx_version.restype = i32
    lib.phdfx_version.argtypes = []
    lib.phdfx_last_error.restype = C.c_char_p
    lib.phdfx_last_error.argtypes = [vp]
    lib.phdfx_create.restype = i32
    lib.phdfx_create.argtypes = [C.POINTER(vp), i32, i32]
    lib.phdfx_destroy.restype = i32
    lib.phdfx_destroy.argtypes = [vp]
    lib.phdfx_load_weights.restype = i32
    lib.phdfx_load_weights.argtypes = [vp, vp, i64, vp, i64, C.POINTER(LayerDesc), i32]
    lib.phdfx_preprocess_u8.restype = i32
    lib.phdfx_preprocess_u8.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp]
    lib.phdfx_preprocess_u8_jitter.restype = i32
    lib.phdfx_preprocess_u8_jitter.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp, vp]
    lib.phdfx_extract_u8_jitter.restype = i32
    lib.phdfx_extract_u8_jitter.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp, vp]
    lib.phdfx_nchw_f32_to_nhwc_bf16.restype = i32
    lib.phdfx_nchw_f32_to_nhwc_bf16.argtypes = [vp, vp, i32, vp, vp]
    lib.phdfx_forward.restype = i32
    lib.phdfx_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.phdfx_forward_timed.restype = i32
    lib.phdfx_forward_timed.argtypes = [vp, vp, i32, vp, vp, C.POINTER(C.c_float), i32]
    lib.phdfx_extract_u8.restype = i32
    lib.phdfx_extract_u8.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp]
    lib.phdfx_run_layer.restype = i32
    lib.phdfx_run_layer.argtypes = [vp, i32, vp, vp, vp, i32, vp]
    lib.phdfx_run_layer2.restype = i32
    lib.phdfx_run_layer2.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp]
    lib.phdfx_chain_span.restype = i32
    lib.phdfx_chain_span.argtypes = [vp, i32]
    lib.phdfx_run_chain.restype = i32
    lib.phdfx_run_chain.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp]
    lib.phdfx_set_schedule.restype = i32
    lib.phdfx_set_schedule.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i32, i32]
    lib.phdfx_get_schedule.restype = i32
    lib.phdfx_get_schedule.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i32, C.POINTER(C.c_int32)]
    lib.phdfx_layer_count.restype = i32
    lib.phdfx_layer_count.argtypes = [vp]
    lib.phdfx_layer_info.restype = i32
    lib.phdfx_layer_info.argtypes = [vp, i32, C.POINTER(LayerDesc)]
    lib.phdfx_linked_launches.restype = i32
    lib.phdfx_linked_launches.argtypes = [vp, i32]
    lib.phdfx_last_launch_count.restype = i32
    lib.phdfx_last_launch_count.argtypes = [vp]
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    """Turn a negative phdfx_status into a RuntimeError carrying the library's message."""
    if rc == 0:
        return
    msg = load().phdfx_last_error(handle)
    raise RuntimeError(f"libphdfx error {rc}: {msg.decode() if msg else '?'}")
