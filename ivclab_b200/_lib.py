"""ctypes binding of libivcb200.so (the C ABI declared in include/ivclab_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an
exception is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``python build_ext.py``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libivcb200.so")

# element type codes (include/ivclab_b200.h)
U8, I32, F32, F64, I64, I16 = 0, 1, 2, 3, 4, 5
ABI_VERSION = 10            # include/ivclab_b200.h IVC_ABI_VERSION (2: entry points added in round 1, zero-run write gained a length; 3: ivc_rgb8_to_luma8;
                           # 4: ivc_pframe_forward_ch, ivc_zerorun_symbol_histogram; 5: ivc_pframe_search_forward; 6: the _zr variants; 7: ivc_pframe_step; 8: ivc_dct8x8_norm; 9: ivc_intra_forward_rgb8_multi; 10: ivc_me_workspace_bytes_planes)
ME_AUTO, ME_EXACT, ME_INT = 0, 1, 2
SSE_RGB8_AS_YCBCR = 103
DIST_RGB, DIST_YCBCR = 1, 2
OK, ERR_ARG, ERR_DTYPE, ERR_SHAPE, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4, -5

_i, _i64, _p = C.c_int, C.c_int64, C.c_void_p
_s5 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "ivc_abi_version": (_i, []),
    "ivc_build_info": (C.c_char_p, []),
    "ivc_error_string": (C.c_char_p, [_i]),
    "ivc_last_cuda_error": (_i, []),
    "ivc_last_cuda_error_string": (C.c_char_p, []),
    "ivc_dct8x8": (_i, [_i, _p, _i, _p, _i, _i64, _i64, _i64, _s5, _p, _i]),
    "ivc_dct8x8_norm": (_i, [_i, _p, _i, _i, _p, _i, _i64, _i64, _i64, _s5, _p, _i]),
    "ivc_quantize": (_i, [_i, _p, _p, _i, _i64, _i64, _i64, _s5, _p, _i, _i, _p]),
    "ivc_dequantize": (_i, [_i, _p, _p, _i, _i64, _i64, _i64, _s5, _p, _i, _i, _p]),
    "ivc_zigzag": (_i, [_i, _p, _i, _p, _i, _i64, _p]),
    "ivc_intra_forward": (_i, [_i, _p, _p, _i, _i64, _i64, _i64, _i64, _i64, _p, _i, _p]),
    "ivc_intra_inverse": (_i, [_i, _p, _p, _i64, _i64, _i64, _i64, _p, _i, _p, _i]),
    "ivc_intra_inverse_rgb": (_i, [_i, _p, _p, _i64, _i64, _i64, _p, _i, _p]),
    "ivc_intra_inverse_sse_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "ivc_intra_inverse_sse": (_i, [_i, _p, _p, _i64, _i64, _i64, _p, _i, _p, _p, _i64, _i, _p, _i64, _p]),
    "ivc_me_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "ivc_me_workspace_bytes_planes": (_i64, [_i64, _i64, _i64]),
    "ivc_me_full_search": (_i, [_i, _p, _p, _p, _i, _i64, _i64, _i64, _i64, _i64, _i, _i, _p, _p, _i64]),
    "ivc_me_full_search_intdtype": (_i, [_i, _p, _p, _p, _i, _i64, _i64, _i64, _i64, _i64, _i, _p]),
    "ivc_mc_reconstruct": (_i, [_i, _p, _p, _i, _i64, _i64, _i64, _i64, _p, _i, _p]),
    "ivc_pframe_forward": (_i, [_i, _p, _p, _p, _p, _i, _i64, _i64, _i64, _i, _p, _i, _p, _p]),
    "ivc_pframe_forward_ch": (_i, [_i, _p, _p, _p, _p, _i, _i64, _i64, _i64, _i, _p, _i, _p, _i, _p]),
    "ivc_pframe_search_forward": (_i, [_i, _p, _p, _p, _i, _i64, _i64, _i64, _i, _i, _p, _i, _i, _p, _p, _p, _i64]),
    "ivc_pframe_search_forward_zr": (_i, [_i, _p, _p, _p, _i, _i64, _i64, _i64, _i, _i, _p, _i, _i, _p, _p, _p, _i64, _p, _p]),
    "ivc_pframe_step": (_i, [_i, _p, _p, _p, _i, _i64, _i64, _i64, _i, _p, _i, _i, _p, _p, _p]),
    "ivc_pframe_inverse": (_i, [_i, _p, _p, _i64, _p, _p, _p, _i, _i64, _i64, _i64, _i, _p, _i, _p]),
    "ivc_sse_workspace_bytes": (_i64, [_i64, _i64]),
    "ivc_sum_squared_error": (_i, [_i, _p, _p, _i, _p, _i, _i64, _i64, _i, _p, _i64, _p]),
    "ivc_zerorun_count": (_i, [_i, _p, _p, _i64, _p]),
    "ivc_zerorun_write": (_i, [_i, _p, _p, _i64, C.c_int32, _p, _p]),
    "ivc_zerorun_count_masks": (_i, [_i, _p, _p, _i64, _p, _p]),
    "ivc_zerorun_write_masks": (_i, [_i, _p, _p, _i64, C.c_int32, _p, _p, _p, _i64]),
    "ivc_zerorun_write_masks_i16": (_i, [_i, _p, _p, _i64, C.c_int32, _p, _p, _p, _i64]),
    "ivc_zerorun_symbol_histogram": (_i, [_i, _p, _p, _i64, _i64, C.c_int32, _i64, _i64, _p, _p]),
    "ivc_zerorun_offsets_workspace_bytes": (_i64, [_i64]),
    "ivc_zerorun_offsets": (_i, [_i, _p, _p, _i64, _p, _p, _i64, _p, _p]),
    "ivc_post_words_to_host": (_i, [_i, _p, _p, _p, _i]),
    "ivc_zerorun_decode_mark": (_i, [_i, _p, _p, _i64, C.c_int32, _p]),
    "ivc_zerorun_decode_ends": (_i, [_i, _p, _p, _p, _i64, _i64, _p]),
    "ivc_zerorun_decode_write": (_i, [_i, _p, _p, _p, _i64, _p, _p]),
    "ivc_symbol_minmax": (_i, [_i, _p, _p, _i, _i64, _p]),
    "ivc_symbol_histogram": (_i, [_i, _p, _p, _i, _i64, _i64, _i64, _i64, _p]),
    "ivc_rgb2ycbcr": (_i, [_i, _p, _p, _i, _i64, _p]),
    "ivc_ycbcr2rgb": (_i, [_i, _p, _p, _i64, _p]),
    "ivc_rgb8_to_luma8": (_i, [_i, _p, _p, _i64, _p, _p]),
    "ivc_intra_forward_rgb8": (_i, [_i, _p, _p, _i64, _i64, _i64, _i64, _p, _i, _p]),
    "ivc_intra_forward_rgb8_zr": (_i, [_i, _p, _p, _i64, _i64, _i64, _i64, _p, _i, _p, _p, _p]),
    "ivc_intra_forward_rgb8_multi": (_i, [_i, _p, _p, _i64, _i64, _i64, _i64, _p, _i, _i, _p, _p, _p]),
}


class IvcError(RuntimeError):
    """A C-ABI call returned a negative status."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. ivclab_b200 has no CPU "
            "fallback. Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ivc_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.ivc_abi_version()} != {ABI_VERSION} (stale build? run build_ext.py --force)")
    return lib


lib = _load()


def check(status: int, what: str) -> None:
    """Map a status code to the exception the reference would have produced."""
    if status == OK:
        return
    msg = f"{what}: {lib.ivc_error_string(status).decode()}"
    if status == ERR_CUDA:
        msg += f" [{lib.ivc_last_cuda_error_string().decode()}]"
        raise IvcError(msg)
    if status in (ERR_SHAPE, ERR_DTYPE):
        raise ValueError(msg)
    raise IvcError(msg)


def strides5(s):
    return (C.c_int64 * 5)(*[int(v) for v in s])
