"""``ivclab.entropy.stats_marg`` on the B200 (reference: ivclab/entropy/entropy.py:6-29; SURVEY.md section 8f
row N3) in the form the coding loop uses it -- the marginal pmf of a symbol stream over unit-width integer
bins (``IntraCodec.train_huffman_from_image``, intracodec.py:160-166) -- plus the min/max reduction that
picks those bins.  Counting happens on the device; only the histogram crosses to the host."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._runtime import code, dev_index, stream_ptr, to_device

__all__ = ["stats_marg", "symbol_minmax", "symbol_histogram", "zerorun_symbol_histogram"]

_INT = (torch.uint8, torch.int32, torch.int64)


def _as_int(t: torch.Tensor, upper=None) -> torch.Tensor:
    """Integer view of the data for unit-bin counting: integer dtypes pass, floats are floored (a value lies
    in bin floor(x - lo)); float values above the last edge are moved out of range first."""
    if t.dtype in _INT:
        return t.contiguous()
    if t.dtype in (torch.int8, torch.int16, torch.bool):
        return t.to(torch.int32).contiguous()
    f = t.to(torch.float64)
    if upper is not None:
        f = torch.where(f > upper, torch.full_like(f, float(upper) + 2.0), f)
    return torch.floor(f).to(torch.int64).contiguous()


def symbol_minmax(symbols):
    """(min, max) of an integer array as Python ints -- ``img_symbols.min()`` / ``.max()`` of
    intracodec.py:162-163 -- reduced on the device."""
    t, _ = to_device(symbols)
    t = _as_int(t.reshape(-1))
    if t.numel() == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    out = torch.empty(2, dtype=torch.int64, device=t.device)
    _lib.check(_lib.lib.ivc_symbol_minmax(dev_index(t), stream_ptr(t.device), t.data_ptr(), code(t.dtype), t.numel(),
                                          out.data_ptr()), "ivc_symbol_minmax")
    lo, hi = out.tolist()
    return int(lo), int(hi)


def symbol_histogram(symbols, lo: int, n_bins: int, hot: int = 4000) -> torch.Tensor:
    """Counts over the unit bins ``[lo+k, lo+k+1)``, the last one closed (== ``np.histogram(x,
    bins=np.arange(lo, lo+n_bins+1))[0]``) as an int64 CUDA tensor; no host synchronisation."""
    t, _ = to_device(symbols)
    t = _as_int(t.reshape(-1), upper=lo + n_bins)
    counts = torch.empty(max(int(n_bins), 0), dtype=torch.int64, device=t.device)
    _lib.check(_lib.lib.ivc_symbol_histogram(dev_index(t), stream_ptr(t.device), t.data_ptr(), code(t.dtype), t.numel(),
                                             int(lo), int(n_bins), int(hot), counts.data_ptr()), "ivc_symbol_histogram")
    return counts


def zerorun_symbol_histogram(flat_patch_img, lo: int = -4096, n_bins: int = 8192, end_of_block: int = 4000):
    """Per-frame histogram of the zero-run symbol stream, computed from the scan blocks WITHOUT writing the stream:
    for ``[N, h, w, c, 64]`` (or ``[h, w, c, 64]``) int32 blocks returns ``(counts, outside)`` with ``counts[i]``
    == ``np.histogram(ZeroRunCoder(end_of_block).encode(blocks[i]), bins=np.arange(lo, lo + n_bins + 1))[0]`` as
    int64-compatible uint32 counts (an int32 CUDA tensor ``[N, n_bins]``) and ``outside[i]`` the number of symbols
    that fall outside ``[lo, lo + n_bins)`` (zero when the range is wide enough: then ``counts[i].sum()`` is the
    stream length, ``counts / counts.sum()`` the ``stats_marg`` pmf and the first / last non-zero bin the min / max
    that ``train_huffman_from_image`` derives its bounds from, intracodec.py:160-166).  No host synchronisation."""
    t, _ = to_device(flat_patch_img)
    if t.ndim not in (4, 5) or t.shape[-1] != 64:
        raise ValueError(f"expected [h, w, c, 64] or [n, h, w, c, 64] scan blocks, got shape {tuple(t.shape)}")
    single = t.ndim == 4
    t = t.to(torch.int32).contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    n = 1 if single else t.shape[0]
    bpu = t.numel() // 64 // max(n, 1)
    counts = torch.empty((n, int(n_bins)), dtype=torch.int32, device=t.device)
    outside = torch.empty(n, dtype=torch.int32, device=t.device)
    _lib.check(_lib.lib.ivc_zerorun_symbol_histogram(dev_index(t), stream_ptr(t.device), t.data_ptr(), n, bpu, int(end_of_block),
                                                     int(lo), int(n_bins), counts.data_ptr(), outside.data_ptr()),
               "ivc_zerorun_symbol_histogram")
    return (counts[0], outside[0]) if single else (counts, outside)


def stats_marg(image, pixel_range):
    """Marginal pmf of ``image`` over ``pixel_range`` (entropy.py:6-29): ``np.histogram(image.flatten(),
    bins=pixel_range)[0] / image.size``.  ``pixel_range`` must be consecutive integers (``np.arange(lo, hi)``,
    the only form the coding loop uses); returns a float64 numpy array of ``len(pixel_range) - 1`` entries."""
    pr = np.asarray(pixel_range.cpu() if isinstance(pixel_range, torch.Tensor) else pixel_range)
    if pr.ndim != 1 or pr.size < 2:
        raise ValueError("`bins` must be a 1-D array of at least two edges")
    if not (np.all(np.diff(pr) == 1) and np.all(pr == np.round(pr))):
        raise NotImplementedError("stats_marg on the device supports consecutive integer bin edges (np.arange(lo, hi)) only")
    lo, n_bins = int(pr[0]), int(pr.size - 1)
    t, _ = to_device(image)
    counts = symbol_histogram(t, lo, n_bins).cpu().numpy()
    return counts / int(t.numel())
