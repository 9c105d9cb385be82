"""``ivclab.entropy.ZeroRunCoder`` on the B200 (reference: ivclab/entropy/zerorun.py:4-87; SURVEY.md
section 8f row N2).  Both directions run on the device: ``encode`` is two kernels around one prefix sum,
``decode`` three kernels around one prefix sum; each needs a single device->host read (the stream length /
the block count and error flag)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._runtime import aligned16, dev_index, stream_ptr, to_device, to_host

__all__ = ["ZeroRunCoder"]


class _PendingEncode:
    """Device-side state of an encode whose stream length is still on its way to the host."""
    __slots__ = ("blocks", "offsets", "masks", "total_host", "event", "nblk", "stream")


class ZeroRunCoder:
    def __init__(self, end_of_block=4000, block_size=64):
        self.EOB = end_of_block
        self.block_size = block_size

    # ---- encode --------------------------------------------------------------------------------
    def _check(self, flat_patch_img):
        if self.block_size != 64:
            raise NotImplementedError("only 64-coefficient blocks are implemented on the device")
        t, was_np = to_device(flat_patch_img)
        if t.ndim not in (4, 5) or t.shape[-1] != 64:        # 5-D = a batch of frames, streams concatenated in order
            raise ValueError(f"expected [h, w, c, 64] (or [n, h, w, c, 64]) scan blocks, got shape {tuple(t.shape)}")
        return aligned16(t.to(torch.int32)), was_np

    def encode_begin(self, flat_patch_img, total_host=None, record=True, counts=None, masks=None) -> _PendingEncode:
        """First half of :meth:`encode` without a host synchronisation: counts the symbols of every block,
        scans them and starts an asynchronous copy of the stream length into pinned host memory.  Lets a
        pipeline enqueue further work before :meth:`encode_finish` waits for that one number.
        ``total_host``: a caller-owned pinned int64[1] to receive the length (nothing is allocated on the host
        then -- required inside CUDA-graph capture); ``record=False`` leaves the synchronisation to the caller
        (who must have waited for this call's work before calling :meth:`encode_finish`).  ``counts`` / ``masks``: the
        per-block symbol counts (int32) and non-zero masks (int64) when the kernel that produced the blocks has
        already emitted them (``forward_rgb(zr=True)``, ``estimate_forward(zr=True)``): the count pass is skipped."""
        t, _ = self._check(flat_patch_img)
        p = _PendingEncode()
        p.blocks, p.nblk, p.stream = t, t.numel() // 64, torch.cuda.current_stream(t.device)
        dev, sp = dev_index(t), stream_ptr(t.device)
        if counts is not None and masks is not None:
            if counts.numel() != p.nblk or masks.numel() != p.nblk or counts.dtype != torch.int32 or masks.dtype != torch.int64:
                raise ValueError("counts (int32) and masks (int64) must hold one entry per scan block")
            p.masks = masks
        else:
            counts = torch.empty(p.nblk, dtype=torch.int32, device=t.device)
            p.masks = torch.empty(p.nblk, dtype=torch.int64, device=t.device)  # 64-bit non-zero masks, count pass -> write pass
            _lib.check(_lib.lib.ivc_zerorun_count_masks(dev, sp, t.data_ptr(), p.nblk, counts.data_ptr(), p.masks.data_ptr()),
                       "ivc_zerorun_count_masks")
        p.offsets = torch.empty(p.nblk, dtype=torch.int64, device=t.device)
        p.total_host = total_host if total_host is not None else torch.zeros(1, dtype=torch.int64).pin_memory()
        if p.nblk:              # one scan kernel; it posts the stream length into the mapped pinned word itself
            wsb = _lib.lib.ivc_zerorun_offsets_workspace_bytes(p.nblk)
            ws = torch.empty(wsb, dtype=torch.uint8, device=t.device)
            _lib.check(_lib.lib.ivc_zerorun_offsets(dev, sp, counts.data_ptr(), p.nblk, p.offsets.data_ptr(), ws.data_ptr(), wsb,
                                                    p.total_host.data_ptr(), None), "ivc_zerorun_offsets")
        else:
            p.total_host.zero_()
        p.event = None
        if record:
            p.event = torch.cuda.Event()
            p.event.record(p.stream)
        return p

    def encode_finish(self, p: _PendingEncode, dtype=torch.int32) -> torch.Tensor:
        """Second half: waits for the stream length (one event), then writes the symbols on the current stream.
        ``dtype=torch.int16`` writes a lossless 16-bit transfer format; the CALLER guarantees that every coefficient
        and the EOB marker fit (values are not checked) -- the reference's dtype, and the default, is int32."""
        if p.event is not None:
            p.event.synchronize()
        total = int(p.total_host[0])
        t = p.blocks
        if dtype not in (torch.int32, torch.int16):
            raise ValueError("symbol dtype must be torch.int32 or torch.int16")
        out = torch.empty(total, dtype=dtype, device=t.device)
        cur = torch.cuda.current_stream(t.device)
        if cur != p.stream and p.event is not None:
            cur.wait_event(p.event)
        fn = _lib.lib.ivc_zerorun_write_masks if dtype == torch.int32 else _lib.lib.ivc_zerorun_write_masks_i16
        _lib.check(fn(dev_index(t), stream_ptr(t.device), t.data_ptr(), p.nblk, int(self.EOB), p.offsets.data_ptr(),
                      p.masks.data_ptr(), out.data_ptr(), total), "ivc_zerorun_write_masks")
        return out

    def encode(self, flat_patch_img):
        """[h, w, c, 64] scan blocks -> int32 symbol stream, blocks in (h w c) order (zerorun.py:10-43)."""
        was_np = not isinstance(flat_patch_img, torch.Tensor)
        return to_host(self.encode_finish(self.encode_begin(flat_patch_img)), was_np)

    # ---- decode --------------------------------------------------------------------------------
    def decode(self, encoded, original_shape):
        """Symbol stream -> ``[h, w, c, 64]`` int32 blocks with the reference's semantics (zerorun.py:44-87):
        parsing stops after ``h*w*c`` blocks (later symbols are ignored), a stream that ends early or a
        block that expands past 64 coefficients raises ``ValueError``.  Divergence: a run length of 0 --
        which the encoder never emits and the reference silently accepts -- raises ``ValueError`` here."""
        if self.block_size != 64:
            raise NotImplementedError("only 64-coefficient blocks are implemented on the device")
        if not isinstance(encoded, (torch.Tensor, np.ndarray)):
            encoded = np.asarray(encoded, dtype=np.int64)
        sym, was_np = to_device(encoded)
        sym = sym.reshape(-1)
        if sym.dtype != torch.int32:
            sym = sym.to(torch.int32)
        sym = sym.contiguous()
        h, w, c = (int(v) for v in original_shape)
        want, n = h * w * c, sym.numel()
        dev, sp = dev_index(sym), stream_ptr(sym.device)
        out = torch.empty((h, w, c, 64), dtype=torch.int32, device=sym.device)
        if want == 0:
            return to_host(out, was_np)
        found, last_is_eob = 0, True
        if n:
            is_eob = torch.empty(n, dtype=torch.int32, device=sym.device)
            _lib.check(_lib.lib.ivc_zerorun_decode_mark(dev, sp, sym.data_ptr(), n, int(self.EOB), is_eob.data_ptr()),
                       "ivc_zerorun_decode_mark")
            rank = torch.cumsum(is_eob, 0, dtype=torch.int64)
            ends = torch.empty(want, dtype=torch.int64, device=sym.device)
            err = torch.zeros(1, dtype=torch.int32, device=sym.device)
            _lib.check(_lib.lib.ivc_zerorun_decode_ends(dev, sp, is_eob.data_ptr(), rank.data_ptr(), n, want, ends.data_ptr()),
                       "ivc_zerorun_decode_ends")
            tail = torch.stack([rank[-1], is_eob[-1].to(torch.int64)])
            found, last_is_eob = (int(v) for v in tail.tolist())
            nb = min(found, want)
            _lib.check(_lib.lib.ivc_zerorun_decode_write(dev, sp, sym.data_ptr(), ends.data_ptr(), nb, out.data_ptr(),
                                                         err.data_ptr()), "ivc_zerorun_decode_write")
            flags = int(err.item()) if nb else 0
            if flags & 1:
                raise ValueError("Block size exceeded: a block expands to more than 64 coefficients")
            if flags & 2:
                raise ValueError("zero-length run in the symbol stream (never produced by ZeroRunCoder.encode)")
        if found < want:
            if n and not last_is_eob:
                raise ValueError("Unexpected end of encoded symbols")
            raise ValueError(f"Expected {want} blocks, got {found}")
        return to_host(out, was_np)
