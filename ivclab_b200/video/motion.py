"""``ivclab.video.MotionCompensator`` on the B200 (reference: ivclab/video/motion.py:3-97)."""
from __future__ import annotations

import torch

from .. import _lib
from .._runtime import code, dev_index, stream_ptr, to_device, to_host

__all__ = ["MotionCompensator"]

_MODES = {"auto": _lib.ME_AUTO, "exact": _lib.ME_EXACT, "int": _lib.ME_INT}


class MotionCompensator:
    """Full-search SSD block matching on 8x8 blocks + block-copy compensation.

    ``compute_motion_vector(ref_image[H,W], image[H,W]) -> int64 [H/8, W/8, 1]``
    with ``index = (dy+sr)*(2sr+1) + (dx+sr)``, candidates visited in (dy, dx)
    raster order, out-of-frame windows skipped, first strict minimum kept
    (motion.py:28-57).  SSDs are accumulated in numpy's summation order with
    individually rounded operations, so vectors are bit-exact for float32 /
    float64 frames; integer-valued [0,255] frames take an exact packed-integer
    kernel (``me_mode='auto'``, decided on the device).  uint8 / int16 / int32
    frames (both of one dtype) get numpy's own wrap-around arithmetic
    (``(block - ref_block) ** 2`` evaluated in that dtype, summed in 64 bits:
    SURVEY.md A10), so the vectors equal the reference's there too; other
    integer dtypes and mixed integer pairs are upcast to float64.  ``reconstruct_with_motion_vector(ref[H,W,C], mv) -> [H,W,C]``
    copies blocks, leaving out-of-frame sources zero (motion.py:76-95).
    """

    def __init__(self, search_range=4, me_mode="auto"):
        self.search_range = search_range
        if me_mode not in _MODES:
            raise ValueError(f"me_mode must be one of {sorted(_MODES)}")
        self.me_mode = me_mode

    def compute_motion_vector(self, ref_image, image):
        ref, was_np = to_device(ref_image)
        cur, _ = to_device(image, ref.device)
        if ref.ndim != 2:
            raise ValueError(f"not enough values to unpack: ref_image must be [H, W], got shape {tuple(ref.shape)}")
        H, W = ref.shape
        if H % 8 or W % 8:
            raise IndexError(f"frame sides ({H}, {W}) must be multiples of the 8x8 block size")
        if cur.ndim != 2 or cur.shape[0] < H or cur.shape[1] < W:
            raise ValueError(f"operands could not be broadcast together: image {tuple(cur.shape)} vs ref_image {(H, W)}")
        cur = cur[:H, :W]
        wrap = {torch.uint8: _lib.U8, torch.int16: _lib.I16, torch.int32: _lib.I32}
        if ref.dtype == cur.dtype and ref.dtype in wrap:          # numpy's integer-dtype arithmetic, replayed literally
            ref, cur = ref.contiguous(), cur.contiguous()
            mv = torch.empty((H // 8, W // 8, 1), dtype=torch.int64, device=ref.device)
            st = _lib.lib.ivc_me_full_search_intdtype(dev_index(ref), stream_ptr(ref.device), ref.data_ptr(), cur.data_ptr(),
                                                      wrap[ref.dtype], 1, H, W, H * W, H * W, int(self.search_range),
                                                      mv.data_ptr())
            _lib.check(st, "ivc_me_full_search_intdtype")
            return to_host(mv, was_np)
        both_f32 = ref.dtype == torch.float32 and cur.dtype == torch.float32
        dt = torch.float32 if both_f32 else torch.float64
        ref = ref.to(dt).contiguous()
        cur = cur.to(dt).contiguous()
        mv = torch.empty((H // 8, W // 8, 1), dtype=torch.int64, device=ref.device)
        mode = _MODES[self.me_mode]
        ws, ws_bytes = None, 0
        if mode != _lib.ME_EXACT:
            planes = dt == torch.float64 and int(self.search_range) >= 8      # wide search: convert the frames to bytes once
            ws_bytes = (_lib.lib.ivc_me_workspace_bytes_planes if planes else _lib.lib.ivc_me_workspace_bytes)(1, H, W)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ref.device)
        st = _lib.lib.ivc_me_full_search(dev_index(ref), stream_ptr(ref.device), ref.data_ptr(), cur.data_ptr(),
                                         code(dt), 1, H, W, H * W, H * W, int(self.search_range), mode,
                                         mv.data_ptr(), ws.data_ptr() if ws is not None else None, ws_bytes)
        _lib.check(st, "ivc_me_full_search")
        return to_host(mv, was_np)

    def reconstruct_with_motion_vector(self, ref_image, motion_vector):
        ref, was_np = to_device(ref_image)
        if ref.ndim != 3:
            raise ValueError(f"ref_image must be [H, W, C], got shape {tuple(ref.shape)}")
        H, W, C = ref.shape
        if H % 8 or W % 8:
            raise IndexError(f"frame sides ({H}, {W}) must be multiples of the 8x8 block size")
        mv, _ = to_device(motion_vector, ref.device)
        if mv.ndim != 3 or mv.shape[0] < H // 8 or mv.shape[1] < W // 8 or mv.shape[2] < 1:
            raise IndexError(f"motion_vector must be [H/8, W/8, 1], got shape {tuple(mv.shape)}")
        mv = mv[:H // 8, :W // 8, 0].to(torch.int64).contiguous()
        ref = ref.contiguous()
        view = ref.view(torch.uint8) if ref.dtype == torch.bool else ref
        out = torch.empty_like(view)
        st = _lib.lib.ivc_mc_reconstruct(dev_index(ref), stream_ptr(ref.device), view.data_ptr(), view.element_size(),
                                         1, H, W, C, mv.data_ptr(), int(self.search_range), out.data_ptr())
        _lib.check(st, "ivc_mc_reconstruct")
        if ref.dtype == torch.bool:
            out = out.view(torch.bool)
        return to_host(out, was_np)
