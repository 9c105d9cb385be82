"""``ivclab.image.IntraCodec`` at the symbol level on the B200 (reference: ivclab/image/intracodec.py:11-168) --
the caller on either side of the hot path: ``image2symbols`` = rgb2ycbcr -> patch -> DCT ->
quantise -> zig-zag -> zero-run encode, ``symbols2image`` = its inverse, ``train_huffman_from_image`` up to the
point where the third-party Huffman coder takes over (alphabet bounds and smoothed pmf).  Every stage runs on the
device; a numpy image costs one upload and the symbol stream one download.

The Huffman stage itself (``constriction``, Rust, not installed here, no tests in the reference) is outside this
package: ``intra_encode`` / ``intra_decode`` / ``encode_decode`` raise ``NotImplementedError``."""
from __future__ import annotations

import numpy as np
import torch

from .._runtime import to_device, to_host
from ..codec import IntraBlockCoder
from ..entropy import ZeroRunCoder, stats_marg, symbol_minmax
from ..signal import DiscreteCosineTransform, rgb2ycbcr
from ..utils import Patcher, ZigZag

__all__ = ["IntraCodec"]


class IntraCodec:
    def __init__(self, quantization_scale=1.0, bounds=(-1000, 4000), end_of_block=4000, block_shape=(8, 8)):
        if tuple(block_shape) != (8, 8):
            raise NotImplementedError("the device path implements the reference's 8x8 blocks only")
        self.quantization_scale = quantization_scale
        self.bounds = None                                  # like the reference: set by train_huffman_from_image (:28)
        self.end_of_block = end_of_block
        self.block_shape = tuple(block_shape)
        self._coder = IntraBlockCoder(quantization_scale)
        self.dct = DiscreteCosineTransform()
        self.quant = self._coder.quant
        self.zigzag = ZigZag()
        self.zerorun = ZeroRunCoder(end_of_block=end_of_block, block_size=64)
        self.huffman = None
        self.pmf = None
        self.patcher = Patcher()

    # ---- intracodec.py:34-79 -----------------------------------------------------------------------
    def image2symbols(self, img, is_source_rgb=True):
        """[H, W, C] (or [H, W]) image -> zero-run symbol stream (int32), identical to the reference's."""
        t, was_np = to_device(img)
        fused = (is_source_rgb and t.dtype == torch.uint8 and t.ndim == 3 and t.shape[2] == 3
                 and t.shape[0] % 8 == 0 and t.shape[1] % 16 == 0)
        if fused:
            zz = self._coder.forward_rgb(t)                 # colour transform inside the forward kernel
        else:
            y = rgb2ycbcr(t) if is_source_rgb else t
            if y.ndim == 2:
                y = y[:, :, None]
            H, W, C = y.shape
            ph, pw = (8 - H % 8) % 8, (8 - W % 8) % 8
            keep_f32 = y.dtype == torch.float32             # only reachable with is_source_rgb=False (rgb2ycbcr returns float64)
            if ph or pw:                                    # np.pad(..., mode='edge') of intracodec.py:57-62
                y = torch.nn.functional.pad(y.permute(2, 0, 1)[None].to(torch.float32 if keep_f32 else torch.float64),
                                            (0, pw, 0, ph), mode="replicate")[0]
                y = y.permute(1, 2, 0)
            if keep_f32:
                # the reference keeps float32 end to end here (scipy's dct stays float32 and the division by a float32
                # table happens in float32, dct.py:24-26, patchquant.py:59): the per-method classes have those paths
                zz = self.zigzag.flatten(self.quant.quantize(self.dct.transform(self.patcher.patch(y.contiguous()))))
            else:
                zz = self._coder.forward(y.to(torch.float64).contiguous())
        return to_host(self.zerorun.encode(zz), was_np)

    # ---- intracodec.py:82-141 ----------------------------------------------------------------------
    def symbols2image(self, symbols, original_shape):
        """Symbol stream -> image: ``[H, W, 3]`` float64 RGB for a 3-element shape; for a 2-element (luma) shape
        the reference decodes the first Hp*Wp blocks against all three tables and returns ``[H, W, 3]`` as well."""
        if len(original_shape) == 2:
            (H, W), C, is_rgb = original_shape, 1, False
        else:
            (H, W, C), is_rgb = original_shape, True
        was_np = not isinstance(symbols, torch.Tensor)
        decoded = self.zerorun.decode(symbols if isinstance(symbols, (torch.Tensor, np.ndarray)) else np.asarray(symbols),
                                      [H // 8, W // 8, C])
        d, _ = to_device(decoded)
        # [8Hp, 8Wp, 3]; colour transform inside the decoder's store.  A 3-element shape with C == 1 takes the
        # reference's `if C == 1` branch (intracodec.py:130-136): the three-table decode is returned unconverted
        out = self._coder.inverse(d, to_rgb=(is_rgb and C == 3))
        if out.shape[0] != H or out.shape[1] != W:
            out = out[:H, :W, :]
        return to_host(out, was_np)

    # ---- intracodec.py:144-168, up to the Huffman coder -----------------------------------------------
    def train_huffman_from_image(self, training_img, is_source_rgb=True):
        """Alphabet bounds (min - 20, max + 21) and the smoothed pmf of the image's symbols -- what the reference
        hands to ``HuffmanCoder.train``.  Stored as ``self.bounds`` / ``self.pmf``; returns None like the reference."""
        t, _ = to_device(training_img)
        sym = self.image2symbols(t, is_source_rgb)
        lo, hi = symbol_minmax(sym)
        self.bounds = (lo - 20, hi + 20 + 1)
        pmf = stats_marg(sym, np.arange(self.bounds[0], self.bounds[1]))
        pmf = pmf + 1e-9                                    # smooth_pmf (entropy.py:31-35)
        self.pmf = pmf / pmf.sum()
        return None

    def _needs_huffman(self, *a, **k):
        raise NotImplementedError("Huffman coding is done by the third-party `constriction` package in the reference "
                                  "(ivclab/entropy/huffman.py); it is outside this package -- feed `image2symbols` "
                                  "output and `self.pmf` / `self.bounds` to it")

    intra_encode = intra_decode = encode_decode = _needs_huffman
