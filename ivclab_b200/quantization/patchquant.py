"""``ivclab.quantization.PatchQuant`` on the B200 (reference: ivclab/quantization/patchquant.py:3-78)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._runtime import code, dev_index, stream_ptr, to_device, to_host

__all__ = ["PatchQuant"]

_TORCH_OF = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
_NP_OF = {torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32,
          torch.int64: np.int64, torch.float16: np.float16, torch.float32: np.float32,
          torch.float64: np.float64, torch.bool: np.bool_}


class PatchQuant:
    """JPEG-table quantiser with the reference's exact arithmetic.

    ``quantize``: ``int32(np.round(x / table))`` -- IEEE division, round half to
    even (patchquant.py:59-60).  ``dequantize``: ``int32(q * table)`` -- product
    in numpy's promoted type, truncation toward zero (patchquant.py:77-78).
    numpy broadcasting of the channel axis against the ``[3, 8, 8]`` table is
    reproduced (a 1-channel input yields 3 output channels).  The table itself
    is built on the host with the very numpy expression of the reference
    (patchquant.py:40-41) so its dtype follows numpy's promotion rules
    (python-float scale -> float32 table, np.float64 scale -> float64 table).
    """

    def __init__(self, quantization_scale=1.0, luminance=None, chrominance=None):
        self.quantization_scale = quantization_scale
        self.luminance = luminance
        self.chrominance = chrominance
        if self.luminance is None:
            self.luminance = np.asarray([
                [16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55],
                [14, 13, 16, 24, 40, 57, 69, 56], [14, 17, 22, 29, 51, 87, 80, 62],
                [18, 55, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99],
            ]).astype(np.float32)
        if self.chrominance is None:
            c = np.full((8, 8), 99)
            c[0, :4] = (17, 18, 24, 47)
            c[1, :4] = (18, 21, 26, 66)
            c[2, :3] = (24, 13, 56)
            c[3, :2] = (47, 66)
            self.chrominance = c.astype(np.float32)
        self._dev_tables = {}

    def get_quantization_table(self):
        return np.stack([self.luminance, self.chrominance, self.chrominance], axis=0) * self.quantization_scale

    # -- internals ---------------------------------------------------------
    def _table_on(self, device):
        tab = np.ascontiguousarray(self.get_quantization_table())
        if tab.shape != (3, 8, 8):
            raise ValueError(f"quantisation table must be [3,8,8], got {tab.shape}")
        if tab.dtype not in _TORCH_OF:
            tab = tab.astype(np.float64)
        key = (str(device), tab.dtype.str, tab.tobytes())
        hit = self._dev_tables.get("t")
        if hit is None or hit[0] != key:
            hit = (key, torch.from_numpy(tab).to(device))
            self._dev_tables["t"] = hit
        return tab, hit[1]

    def _run(self, x, dequant: bool):
        t, was_np = to_device(x)
        tab, dtab = self._table_on(t.device)
        if t.ndim < 2 or tuple(t.shape[-2:]) != (8, 8):
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(t.shape)} (1,1,3,8,8)")
        shape = (1,) * max(0, 5 - t.ndim) + tuple(t.shape)
        C = shape[-3]
        if C not in (1, 3):
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(t.shape)} (1,1,3,8,8)")
        # numpy's promotion decides the arithmetic type of x / table and q * table
        res = np.result_type(_NP_OF[t.dtype], tab.dtype)
        compute = _lib.F32 if res == np.float32 else _lib.F64
        if t.dtype in (torch.int8, torch.int16, torch.bool):
            t = t.to(torch.int32)
        elif t.dtype == torch.float16:
            t = t.to(torch.float32)
        v = t.reshape(shape)
        if v.ndim > 5:
            v = v.reshape((-1,) + shape[-4:])
        n0, n1 = v.shape[0], v.shape[1]
        out = torch.empty((n0, n1, 3, 8, 8), dtype=torch.int32, device=v.device)
        fn = _lib.lib.ivc_dequantize if dequant else _lib.lib.ivc_quantize
        st = fn(dev_index(v), stream_ptr(v.device), v.data_ptr(), code(v.dtype), n0, n1, C,
                _lib.strides5(v.stride()), dtab.data_ptr(), code(dtab.dtype), compute, out.data_ptr())
        _lib.check(st, "ivc_dequantize" if dequant else "ivc_quantize")
        return to_host(out.reshape(shape[:-3] + (3, 8, 8)), was_np)

    def quantize(self, patched_img):
        """[H_patch, W_patch, C, 8, 8] -> int32 [H_patch, W_patch, 3, 8, 8] (patchquant.py:44-60)."""
        return self._run(patched_img, False)

    def dequantize(self, quantized_img):
        """[H_patch, W_patch, C, 8, 8] -> int32 [H_patch, W_patch, 3, 8, 8] (patchquant.py:62-78)."""
        return self._run(quantized_img, True)
