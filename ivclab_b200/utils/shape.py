"""``ivclab.utils.ZigZag`` / ``Patcher`` on the B200 (reference: ivclab/utils/shape.py:4-65)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._runtime import dev_index, stream_ptr, to_device, to_host

__all__ = ["ZigZag", "Patcher"]


class ZigZag:
    """8x8 zig-zag scan / un-scan.  ``zigzag_order[k]`` is the scan position of
    raster index ``k`` (shape.py:10-19); ``flatten`` scatters, ``unflatten``
    gathers (shape.py:26, :32); dtype is preserved; no debug printing."""

    def __init__(self):
        self.zigzag_order = np.asarray([
            0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42,
            3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
            10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
            21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63])

    @staticmethod
    def _run(t: torch.Tensor, inverse: bool) -> torch.Tensor:
        t = t.contiguous()
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        out = torch.empty_like(t)
        st = _lib.lib.ivc_zigzag(dev_index(t), stream_ptr(t.device), int(inverse), t.data_ptr(),
                                 t.element_size(), t.numel() // 64, out.data_ptr())
        _lib.check(st, "ivc_zigzag")
        return out

    def flatten(self, patched_img):
        """[h, w, c, p0, p1] (p0*p1 == 64) -> [h, w, c, 64] in scan order (shape.py:21-28)."""
        t, was_np = to_device(patched_img)
        if t.ndim != 5:
            raise ValueError(f"ZigZag.flatten expects 'h w c p0 p1' (5 axes), got shape {tuple(t.shape)}")
        if t.shape[3] * t.shape[4] != 64:
            raise ValueError(f"shape mismatch: cannot scan blocks of {t.shape[3]}x{t.shape[4]} into 64 positions")
        out = self._run(t.reshape(t.shape[:3] + (64,)), False)
        return to_host(out, was_np)

    def unflatten(self, unshuffled):
        """[h, w, c, 64] -> [h, w, c, 8, 8] (shape.py:30-36)."""
        t, was_np = to_device(unshuffled)
        if t.ndim != 4:
            raise IndexError(f"ZigZag.unflatten expects 4 axes [h, w, c, 64], got shape {tuple(t.shape)}")
        if t.shape[3] < 64:
            raise IndexError(f"index 63 is out of bounds for axis 3 with size {t.shape[3]}")
        if t.shape[3] > 64:
            t = t[..., :64]                      # fancy indexing with zigzag_order only touches 0..63
        out = self._run(t, True)
        return to_host(out.reshape(t.shape[:3] + (8, 8)), was_np)


class Patcher:
    """Block view ``'(h p0) (w p1) c -> h w c p0 p1'`` and its inverse (shape.py:38-65).
    Pure layout: numpy arrays give numpy views, tensors give tensor views; no kernel runs."""

    def __init__(self, window_size=(8, 8)):
        self.window_size = window_size

    def patch(self, img):
        p0, p1 = self.window_size
        H, W, C = img.shape
        if H % p0 or W % p1:
            raise ValueError(f"image sides ({H}, {W}) are not multiples of the window {self.window_size}")
        v = img.reshape(H // p0, p0, W // p1, p1, C)
        return v.transpose(0, 2, 4, 1, 3) if isinstance(img, np.ndarray) else v.permute(0, 2, 4, 1, 3)

    def unpatch(self, patched_img):
        h, w, c, p0, p1 = patched_img.shape
        if isinstance(patched_img, np.ndarray):
            return np.ascontiguousarray(patched_img.transpose(0, 3, 1, 4, 2)).reshape(h * p0, w * p1, c)
        return patched_img.permute(0, 3, 1, 4, 2).reshape(h * p0, w * p1, c)
