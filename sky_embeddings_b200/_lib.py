"""ctypes binding of libskysearch.so (the C ABI declared in include/sky_search.h).

The product path has NO fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libskysearch.so")

# constants mirrored from include/sky_search.h
F32, BF16 = 0, 1
COSINE, MSE, MAE = 0, 1, 2
MEAN, MIN, MAX = 0, 1, 2
TOK_ALL, TOK_CLS, TOK_PATCHES, TOK_MAXPOOL = 0, 1, 2, 3
PATH_AUTO, PATH_SIMT, PATH_TENSOR, PATH_GENERIC, PATH_BATCH = 0, 1, 2, 3, 4

METRICS = {"cosine": COSINE, "MSE": MSE, "MAE": MAE}
COMBINES = {"mean": MEAN, "min": MIN, "max": MAX}
PATHS = {"auto": PATH_AUTO, "simt": PATH_SIMT, "tensor": PATH_TENSOR, "generic": PATH_GENERIC, "batch": PATH_BATCH}

# every symbol include/sky_search.h declares: (restype, argtypes)
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
SIGNATURES = {
    "sky_last_error": (C.c_char_p, []),
    "sky_abi_version": (_i, []),
    "sky_bank_create": (_i, [C.POINTER(_vp), _i, _i64, _i, _i, _i]),
    "sky_bank_destroy": (_i, [_vp]),
    "sky_bank_fit_norm": (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "sky_bank_set_norm": (_i, [_vp, _vp, _vp, _vp]),
    "sky_bank_get_norm": (_i, [_vp, _vp, _vp, _vp]),
    "sky_bank_upload": (_i, [_vp, _vp, _i, _i64, _i64, _i, _i, _i, _vp]),
    "sky_bank_finalize": (_i, [_vp, _vp]),
    "sky_bank_resize": (_i, [_vp, _i64]),
    "sky_bank_download": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "sky_bank_info": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "sky_query_from_targets": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    "sky_search": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _vp, _vp, _i, _vp]),
    "sky_search_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _vp, _vp, _i, _vp]),
    "sky_score": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _vp, _vp]),
    "sky_pixel_bank_create": (_i, [C.POINTER(_vp), _i, _i64, _i, _i, _i]),
    "sky_pixel_bank_upload": (_i, [_vp, _vp, _i64, _i64, _vp]),
    "sky_search_pixels": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp, _vp, _vp]),
    "sky_score_pixels": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _vp, _vp]),
    "sky_pixel_snr": (_i, [_vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sky_pixel_bank_snr": (_i, [_vp, _i, _i, _i, _i64, _i64, _i, _i, _vp, _vp, _vp]),
    "sky_tile_cutouts": (_i, [_vp, _i, _i, _i, _vp, _i64, _i, C.c_float, C.c_float, _vp, _i, _vp]),
    "sky_center_clip": (_i, [_vp, _i64, _i, _i, _i, _i, C.c_float, C.c_float, _vp, _i, _vp]),
    "sky_merge_candidates": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sky_merge_candidates_strided": (_i, [_vp, _vp, _i, _i, _i, _i64, _i64, _i, _i, _vp, _vp, _i, _vp]),
    "sky_exchange_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "sky_exchange_handle_bytes": (_i, []),
    "sky_exchange_handle": (_i, [_vp, _vp]),
    "sky_exchange_open": (_i, [_vp, _vp]),
    "sky_exchange_open_local": (_i, [_vp, _vp]),
    "sky_exchange_local_ptr": (_vp, [_vp]),
    "sky_exchange_destroy": (_i, [_vp]),
    "sky_exchange_merge": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sky_search_sharded": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _vp, _vp, _i, _vp]),
    "sky_profile_enable": (_i, [_vp, _i]),
    "sky_profile_read": (_i, [_vp, C.POINTER(_i64), C.POINTER(C.c_double), _i]),
    "sky_launch_count": (_i64, [_i]),
}


class SkyError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libskysearch error {code}: {text}")
        self.code = code


_lib = None


def load():
    """Load libskysearch.so (built in-tree by sky_embeddings_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m sky_embeddings_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the search path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SkyError(rc, load().sky_last_error().decode("utf-8", "replace"))


def launch_count(reset=False):
    return int(load().sky_launch_count(1 if reset else 0))
