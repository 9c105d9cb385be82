"""Intra rate-distortion sweep of host-resident frames (BASELINE.json configuration 3; the loops of
exercises/ch4/ex1.py:385-437 and the ch3 RD exercises): every frame is uploaded ONCE and coded at all
quantisation scales on the device; per (scale, frame) only what the sweep consumes comes back --

* the squared error between the uint8 RGB original and ``ycbcr2rgb(reconstruction)``, i.e. the PSNR that
  ``calc_psnr(img, codec.symbols2image(symbols, img.shape))`` reports (utils/metrics.py:3-40,
  intracodec.py:139-141), out of the decoder kernel itself (nothing is reconstructed into memory);
* the histogram of the zero-run symbols of ``image2symbols(img)``, i.e. the counts behind
  ``stats_marg(symbols, np.arange(min - 20, max + 21))`` with which ``train_huffman_from_image`` trains the
  entropy coder (intracodec.py:160-166), counted from the scan indices without writing the symbol stream.

Neither scan indices nor symbols nor reconstructions cross PCIe: 3 bytes per pixel go up once, a few tens of
kilobytes per rate-distortion point come down.  Uploads, the ``len(qscales)`` coding passes of a chunk and the
downloads of the previous chunk's results overlap on three streams."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib  # noqa: F401
from .codec import IntraBlockCoder, forward_rgb_multi

__all__ = ["RateDistortionSweep"]

QSCALES_EX1 = (0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5)          # exercises/ch4/ex1.py:385


class RateDistortionSweep:
    def __init__(self, qscales=QSCALES_EX1, chunk_frames=8, device=None, hist_lo=-4096, hist_bins=8192,
                 end_of_block=4000, slots=3):
        self.qscales = tuple(qscales)
        self.coders = [IntraBlockCoder(q) for q in self.qscales]
        self.chunk = int(chunk_frames)
        self.lo, self.nb, self.eob = int(hist_lo), int(hist_bins), int(end_of_block)
        self.nslots = max(2, int(slots))
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._s_in, self._s_cmp, self._s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self._slots = None
        self._host = None

    # ---- one chunk on the device: all scales over the same resident frames ------------------------------
    def code(self, d_rgb, out=None, at=0):
        """d_rgb [n,H,W,3] uint8 on the device -> (sse [Q,n] float64, hist [Q,n,bins] int32, outside [Q,n] int32),
        device tensors, enqueued on the current stream without synchronisation.  ``out=(sse, hist, outside)`` with
        F >= at + n columns: the results are written at columns ``at .. at+n`` of those tensors instead."""
        n = d_rgb.shape[0]
        Q = len(self.coders)
        if out is None:
            sse = torch.empty((Q, n), dtype=torch.float64, device=d_rgb.device)
            hist = torch.empty((Q, n, self.nb), dtype=torch.int32, device=d_rgb.device)
            outside = torch.empty((Q, n), dtype=torch.int32, device=d_rgb.device)
            at = 0
        else:
            sse, hist, outside = out
        sp = torch.cuda.current_stream(d_rgb.device).cuda_stream
        # rgb2ycbcr + DCT ONCE, quantise + zig-zag per scale (widths that are not a multiple of 16: one forward per scale)
        zz_all = forward_rgb_multi(self.coders, d_rgb) if d_rgb.dtype == torch.uint8 and d_rgb.shape[2] % 16 == 0 else None
        for qi, coder in enumerate(self.coders):
            zz = zz_all[qi] if zz_all is not None else coder.forward_rgb(d_rgb)
            _lib.check(_lib.lib.ivc_zerorun_symbol_histogram(
                d_rgb.device.index, sp, zz.data_ptr(), n, zz.numel() // 64 // n, self.eob, self.lo, self.nb,
                hist[qi, at:at + n].data_ptr(), outside[qi, at:at + n].data_ptr()), "ivc_zerorun_symbol_histogram")
            sse[qi, at:at + n] = coder.inverse_with_distortion(zz, d_rgb, space="rgb")   # decode + ycbcr2rgb + error, nothing stored
        return sse, hist, outside

    # ---- the host-fed pipeline ------------------------------------------------------------------------------
    def _buffers(self, F, H, W, need_host=True):
        Q = len(self.coders)
        key = (F, H, W, Q, self.nb)
        if need_host and (self._host is None or self._host[0] != key):
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._host = (key, {"sse": pin((Q, F), torch.float64), "hist": pin((Q, F, self.nb), torch.int32),
                                "outside": pin((Q, F), torch.int32)})
        skey = (self.chunk, H, W)
        if self._slots is None or self._slots[0] != skey:
            self._slots = (skey, [torch.empty((self.chunk, H, W, 3), dtype=torch.uint8, device=self.device)
                                  for _ in range(self.nslots)])
        return (self._host[1] if need_host else None), self._slots[1]

    def run(self, rgb, to_host=True, out=None, out_at=0):
        """rgb [F,H,W,3] uint8, a pinned host tensor (numpy arrays are accepted and pinned once; H a multiple of 8,
        W of 16).  Returns host arrays -- views of this object's pinned buffers, valid until the next ``run`` --
        ``sse`` [Q,F] float64 (``psnr = 10 log10(255^2 / (sse / (H W 3)))``), ``hist`` [Q,F,bins] int32 with
        ``hist[q,f,k]`` = number of symbols equal to ``hist_lo + k``, ``outside`` [Q,F] (symbols outside the
        histogram's range; 0 for 8-bit frames with the default range), plus ``h2d_bytes`` / ``d2h_bytes``.
        ``to_host=False`` leaves the three results on the device (CUDA tensors, complete when the compute stream
        ``self._s_cmp`` is): the form a multi-GPU run hands to ``shard.gather_rows``.  ``out`` (a dict of pinned host
        tensors ``sse`` [Q,Ftot], ``hist`` [Q,Ftot,bins], ``outside`` [Q,Ftot]) with ``out_at``: the results of frame f
        are downloaded to column ``out_at + f`` of those tensors instead of this object's own buffers -- with
        ``shard.SharedPinned`` buffers every rank of a sharded run writes its frame range straight into the array
        rank 0 reads (the host gather without a collective)."""
        if isinstance(rgb, np.ndarray):
            rgb = torch.from_numpy(rgb)
        if rgb.dtype != torch.uint8 or rgb.ndim != 4 or rgb.shape[-1] != 3 or rgb.shape[1] % 8 or rgb.shape[2] % 16:
            raise ValueError(f"expected uint8 RGB frames [F,H,W,3] with H % 8 == 0 and W % 16 == 0, got {rgb.dtype} {tuple(rgb.shape)}")
        if not rgb.is_pinned():
            rgb = rgb.pin_memory()
        F, H, W, _ = rgb.shape
        hb, slots = self._buffers(F, H, W, need_host=out is None and to_host)
        if out is not None:
            hb, to_host = out, True
        Q, C, S = len(self.coders), self.chunk, self.nslots
        bounds = [(lo, min(lo + C, F)) for lo in range(0, F, C)]
        n = len(bounds)
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_cmp = [torch.cuda.Event() for _ in range(n)]
        results = {}
        dev_out = None
        if not to_host:
            dev_out = (torch.empty((Q, F), dtype=torch.float64, device=self.device),
                       torch.empty((Q, F, self.nb), dtype=torch.int32, device=self.device),
                       torch.empty((Q, F), dtype=torch.int32, device=self.device))

        def upload(k):
            lo, hi = bounds[k]
            with torch.cuda.stream(self._s_in):
                if k >= S:
                    self._s_in.wait_event(ev_cmp[k - S])                  # the slot's previous chunk has been coded
                slots[k % S][:hi - lo].copy_(rgb[lo:hi], non_blocking=True)
                ev_in[k].record(self._s_in)

        def compute(k):
            lo, hi = bounds[k]
            with torch.cuda.stream(self._s_cmp):
                self._s_cmp.wait_event(ev_in[k])
                if to_host:
                    results[k] = self.code(slots[k % S][:hi - lo])
                else:
                    self.code(slots[k % S][:hi - lo], out=dev_out, at=lo)
                ev_cmp[k].record(self._s_cmp)

        def download(k):
            lo, hi = bounds[k]
            sse, hist, outside = results.pop(k)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev_cmp[k])
                for t in (sse, hist, outside):
                    t.record_stream(self._s_out)
                a, b = out_at + lo, out_at + hi
                for qi in range(Q):                                        # [qi, a:b] is contiguous on both sides
                    hb["hist"][qi, a:b].copy_(hist[qi], non_blocking=True)
                    hb["sse"][qi, a:b].copy_(sse[qi], non_blocking=True)
                    hb["outside"][qi, a:b].copy_(outside[qi], non_blocking=True)

        for k in range(min(S - 1, n)):
            upload(k)
        for k in range(n):
            if k + S - 1 < n:
                upload(k + S - 1)
            compute(k)
            if to_host and k >= 1:
                download(k - 1)
        if not to_host:
            for t in dev_out:
                t.record_stream(self._s_cmp)
            return {"sse": dev_out[0], "hist": dev_out[1], "outside": dev_out[2], "qscales": self.qscales,
                    "hist_lo": self.lo, "h2d_bytes": rgb.numel(), "d2h_bytes": 0}
        if n:
            download(n - 1)
        self._s_out.synchronize()
        view = (lambda t: t[:, out_at:out_at + F].numpy()) if out is not None else (lambda t: t.numpy())
        return {"sse": view(hb["sse"]), "hist": view(hb["hist"]), "outside": view(hb["outside"]),
                "qscales": self.qscales, "hist_lo": self.lo, "h2d_bytes": rgb.numel(),
                "d2h_bytes": Q * F * (8 + 4 * self.nb + 4)}

    # ---- what the sweep's consumers compute from the results (host, numpy) -------------------------------------
    @staticmethod
    def psnr(sse, samples_per_frame, maxval=255.0):
        """calc_psnr of utils/metrics.py:21-40 from summed squared errors."""
        mse = np.asarray(sse, dtype=np.float64) / float(samples_per_frame)
        with np.errstate(divide="ignore"):
            return 10.0 * np.log10(maxval * maxval / mse)

    def entropy_bits(self, hist):
        """Zeroth-order entropy of every symbol histogram times its symbol count: the bit budget an ideal entropy
        coder trained on that frame would need (calc_entropy of entropy.py:37-50 on the stats_marg pmf)."""
        h = np.asarray(hist, dtype=np.float64)
        n = h.sum(axis=-1, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            p = np.where(h > 0, h / n, 1.0)
            return -(h * np.log2(p)).sum(axis=-1)
