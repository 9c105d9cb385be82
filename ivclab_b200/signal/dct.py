"""``ivclab.signal.DiscreteCosineTransform`` on the B200 (reference: ivclab/signal/dct.py:4-46)."""
from __future__ import annotations

import torch

from .. import _lib
from .._runtime import code, dev_index, stream_ptr, to_device, to_host

__all__ = ["DiscreteCosineTransform"]


class DiscreteCosineTransform:
    """Forward / inverse 8x8 DCT-II over the last two axes (orthonormal by default; `norm` as scipy spells it).

    Same constructor, methods and array conventions as the reference class
    (dct.py:9, :12-28, :30-46): input ``[..., 8, 8]`` (typically the strided
    ``[H_patch, W_patch, C, 8, 8]`` view made by ``Patcher.patch``), output a
    fresh C-contiguous array of the same shape; float32 stays float32, every
    other dtype is computed and returned in float64 -- bit-identical to
    ``scipy.fft.dct/idct(norm='ortho')``.  numpy in -> numpy out, CUDA tensor
    in -> CUDA tensor out (same device, current stream, no synchronisation).
    """

    def __init__(self, norm='ortho'):
        self.norm = norm

    # scipy's spellings of `norm` (the reference forwards it, dct.py:24,26,42,44; ivclab itself only uses 'ortho',
    # intracodec.py:25, tests/ch3.py:15) -> IVC_NORM_*
    _NORMS = {'ortho': 0, None: 1, 'backward': 1, 'forward': 2}

    def _run(self, x, inverse: bool):
        try:
            norm = self._NORMS[self.norm]
        except (KeyError, TypeError):
            # what scipy raises, at the call (not in the constructor), for anything else
            raise ValueError(f'Invalid norm value {self.norm!r}; should be "backward", "ortho" or "forward".') from None
        t, was_np = to_device(x)
        if t.ndim < 2 or t.shape[-1] != 8 or t.shape[-2] != 8:
            raise ValueError(f"expected [..., 8, 8] patches, got shape {tuple(t.shape)}")
        if t.dtype not in (torch.uint8, torch.int32, torch.float32, torch.float64):
            t = t.to(torch.float64)
        shape = tuple(t.shape)
        if t.ndim == 5:
            v = t
        else:
            v = t.reshape(-1, 1, 1, 8, 8)
        out_dtype = torch.float32 if v.dtype == torch.float32 else torch.float64
        out = torch.empty(v.shape, dtype=out_dtype, device=v.device)
        n0, n1, c = v.shape[:3]
        st = _lib.lib.ivc_dct8x8_norm(dev_index(v), stream_ptr(v.device), int(inverse), norm, v.data_ptr(), code(v.dtype),
                                      n0, n1, c, _lib.strides5(v.stride()), out.data_ptr(), code(out_dtype))
        _lib.check(st, "ivc_dct8x8_norm")
        return to_host(out.reshape(shape), was_np)

    def transform(self, patched_img):
        """[H_patch, W_patch, C, 8, 8] -> same shape; DCT-II along axis -1 then -2 (dct.py:24-26)."""
        return self._run(patched_img, False)

    def inverse_transform(self, transformed):
        """Inverse of :meth:`transform` (dct.py:42-44)."""
        return self._run(transformed, True)
