"""``ivclab.entropy.ZeroRunCoder.encode`` on the B200 (reference: ivclab/entropy/zerorun.py:4-43;
SURVEY.md section 8f row N2).  ``decode`` is host-side parsing of a variable-length stream and is
not on the device path; it is provided for round trips with the reference's exact semantics."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._runtime import aligned16, dev_index, stream_ptr, to_device, to_host

__all__ = ["ZeroRunCoder"]


class ZeroRunCoder:
    def __init__(self, end_of_block=4000, block_size=64):
        self.EOB = end_of_block
        self.block_size = block_size

    def encode(self, flat_patch_img):
        """[h, w, c, 64] scan blocks -> int32 symbol stream, blocks in (h w c) order (zerorun.py:15-41).
        Two kernels around one ``cumsum``; the stream length needs one device->host read."""
        if self.block_size != 64:
            raise NotImplementedError("only 64-coefficient blocks are implemented on the device")
        t, was_np = to_device(flat_patch_img)
        if t.ndim not in (4, 5) or t.shape[-1] != 64:        # 5-D = a batch of frames, streams concatenated in order
            raise ValueError(f"expected [h, w, c, 64] (or [n, h, w, c, 64]) scan blocks, got shape {tuple(t.shape)}")
        t = aligned16(t.to(torch.int32))
        nblk = t.numel() // 64
        dev, sp = dev_index(t), stream_ptr(t.device)
        counts = torch.empty(nblk, dtype=torch.int32, device=t.device)
        _lib.check(_lib.lib.ivc_zerorun_count(dev, sp, t.data_ptr(), nblk, counts.data_ptr()), "ivc_zerorun_count")
        ends = torch.cumsum(counts, 0, dtype=torch.int64)
        offsets = (ends - counts).contiguous()
        total = int(ends[-1].item()) if nblk else 0
        out = torch.empty(total, dtype=torch.int32, device=t.device)
        _lib.check(_lib.lib.ivc_zerorun_write(dev, sp, t.data_ptr(), nblk, int(self.EOB), offsets.data_ptr(), out.data_ptr()),
                   "ivc_zerorun_write")
        return to_host(out, was_np)

    def decode(self, encoded, original_shape):
        """Host-side inverse with the reference's stop-after-h*w*c-blocks rule and error conditions
        (zerorun.py:44-87)."""
        enc = encoded.cpu().numpy() if isinstance(encoded, torch.Tensor) else np.asarray(encoded)
        h, w, c = original_shape
        want = h * w * c
        blocks = np.zeros((want, self.block_size), dtype=np.int32)
        i = n = 0
        while i < len(enc) and n < want:
            pos = 0
            while True:
                if i >= len(enc):
                    raise ValueError("Unexpected end of encoded symbols")
                s = int(enc[i])
                i += 1
                if s == self.EOB:
                    break
                if s == 0:
                    pos += int(enc[i])
                    i += 1
                else:
                    if pos >= self.block_size:
                        raise ValueError(f"Block size exceeded: {pos + 1}")
                    blocks[n, pos] = s
                    pos += 1
                if pos > self.block_size:
                    raise ValueError(f"Block size exceeded: {pos}")
            n += 1
        if n != want:
            raise ValueError(f"Expected {want} blocks, got {n}")
        return blocks.reshape(h, w, c, self.block_size)
