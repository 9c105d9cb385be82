"""Host <-> device streaming of the coding loop: while chunk k is being coded, chunk k+1 is uploaded
and the results of chunk k-1 are downloaded, on three CUDA streams with pinned staging buffers.

This is the form in which the path is fed from HOST memory (what ``bench.py`` reports as ``e2e``):
uint8 RGB frames and uint8 luma planes go up (4-5 bytes per pixel), zero-run symbol streams, motion
vectors and squared errors come back; the colour transform, both transform loops, the motion search,
the zero-run coder and the error reduction all run on the device in between.  Frames are independent
(intra) or frame pairs (inter), so chunks never depend on each other."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib  # noqa: F401
from .codec import IntraBlockCoder, PFrameBlockCoder
from .entropy import ZeroRunCoder
from .signal.color import rgb2ycbcr
from .utils.metrics import frame_sse

__all__ = ["StreamedCoder"]


class StreamedCoder:
    def __init__(self, quantization_scale=1.0, search_range=4, me_mode="auto", chunk_frames=2, device=None):
        self.intra = IntraBlockCoder(quantization_scale)
        self.pframe = PFrameBlockCoder(quantization_scale, search_range, me_mode)
        self.zr = ZeroRunCoder()
        self.chunk = int(chunk_frames)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._s_in, self._s_cmp, self._s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self._host = None

    def _host_buffers(self, F, H, W):
        key = (F, H, W)
        if self._host is None or self._host[0] != key:
            nb = (H // 8) * (W // 8)
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._host = (key, {
                "sym_intra": pin(F * nb * 3 * 65, torch.int32), "sym_inter": pin(F * nb * 3 * 65, torch.int32),
                "mv": pin((F, H // 8, W // 8, 1), torch.int64), "sse": pin((2, F), torch.float64)})
        return self._host[1]

    def run(self, rgb, cur, ref):
        """rgb [F,H,W,3] uint8, cur/ref [F,H,W] uint8 luma planes -- pinned host tensors (numpy arrays are
        accepted and pinned once).  Returns host-side results: ``sym_intra`` / ``sym_inter`` (int32 streams
        and per-chunk lengths), ``mv`` [F,Hp,Wp,1] int64, ``sse`` [2,F] (intra on YCbCr, inter on luma)."""
        pinned = lambda x: (torch.from_numpy(x) if isinstance(x, np.ndarray) else x)
        rgb, cur, ref = (t if t.is_pinned() else t.pin_memory() for t in map(pinned, (rgb, cur, ref)))
        F, H, W, _ = rgb.shape
        hb = self._host_buffers(F, H, W)
        dev = self.device
        C = self.chunk
        nchunks = (F + C - 1) // C
        slots = [{"rgb": torch.empty((C, H, W, 3), dtype=torch.uint8, device=dev),
                  "cur": torch.empty((C, H, W), dtype=torch.uint8, device=dev),
                  "ref": torch.empty((C, H, W), dtype=torch.uint8, device=dev)} for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(nchunks)]
        ev_cmp = [torch.cuda.Event() for _ in range(nchunks)]
        ev_out = [torch.cuda.Event() for _ in range(nchunks)]
        keep = {}                                     # device results of in-flight chunks

        def upload(k):
            lo, hi = k * C, min(F, (k + 1) * C)
            with torch.cuda.stream(self._s_in):
                if k >= 2:
                    self._s_in.wait_event(ev_cmp[k - 2])          # the slot's previous chunk has been consumed
                s = slots[k & 1]
                for name, src in (("rgb", rgb), ("cur", cur), ("ref", ref)):
                    s[name][:hi - lo].copy_(src[lo:hi], non_blocking=True)
                ev_in[k].record(self._s_in)

        lens_i, lens_p = [], []
        off_i = off_p = 0
        upload(0)
        for k in range(nchunks):
            lo, hi = k * C, min(F, (k + 1) * C)
            n = hi - lo
            if k + 1 < nchunks:
                upload(k + 1)                                      # overlaps with the work below
            with torch.cuda.stream(self._s_cmp):
                self._s_cmp.wait_event(ev_in[k])
                s = slots[k & 1]
                d_rgb, d_cur, d_ref = s["rgb"][:n], s["cur"][:n].double(), s["ref"][:n].double()
                zz = self.intra.forward_rgb(d_rgb)
                rec = self.intra.inverse(zz)
                sse_i = frame_sse(rgb2ycbcr(d_rgb), rec)
                mv = self.pframe.estimate(d_ref, d_cur)
                zzp = self.pframe.forward(d_cur, d_ref, mv)
                recp = self.pframe.inverse(zzp, ref=d_ref, mv=mv)
                sse_p = frame_sse(d_cur, recp)
                sym_i = self.zr.encode(zz)                         # (reads the stream lengths: syncs this stream only)
                sym_p = self.zr.encode(zzp)
                ev_cmp[k].record(self._s_cmp)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev_cmp[k])
                for t in (sym_i, sym_p, mv, sse_i, sse_p):
                    t.record_stream(self._s_out)
                hb["sym_intra"][off_i:off_i + sym_i.numel()].copy_(sym_i, non_blocking=True)
                hb["sym_inter"][off_p:off_p + sym_p.numel()].copy_(sym_p, non_blocking=True)
                hb["mv"][lo:hi].copy_(mv, non_blocking=True)
                hb["sse"][0, lo:hi].copy_(sse_i, non_blocking=True)
                hb["sse"][1, lo:hi].copy_(sse_p, non_blocking=True)
                ev_out[k].record(self._s_out)
            keep[k] = (sym_i, sym_p, mv, sse_i, sse_p)
            keep.pop(k - 2, None)
            lens_i.append(sym_i.numel())
            lens_p.append(sym_p.numel())
            off_i += sym_i.numel()
            off_p += sym_p.numel()
        ev_out[-1].synchronize()
        self._s_cmp.synchronize()
        return {"sym_intra": hb["sym_intra"][:off_i], "sym_inter": hb["sym_inter"][:off_p], "len_intra": lens_i,
                "len_inter": lens_p, "mv": hb["mv"], "sse": hb["sse"],
                "h2d_bytes": rgb.numel() + cur.numel() + ref.numel(),
                "d2h_bytes": (off_i + off_p) * 4 + hb["mv"].numel() * 8 + hb["sse"].numel() * 8}
