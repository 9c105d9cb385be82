"""Closed-loop luma video coding on the B200: the per-frame loop of the ch4 exercise codecs
(exercises/ch4/E4-1.py:212-306; ivclab/video/videocodec.py:41-75) at the symbol level.

Per P-frame three kernels run back to back on one stream, with no host synchronisation and every
intermediate resident in HBM:  K3 motion search (decoder reconstruction vs current frame) ->
K1p (MC + residual + DCT + quantise + zig-zag) -> K2p (dequantise + IDCT + prediction add), the last
of which produces the reference frame of the next iteration.  A whole sequence can be captured in
one CUDA graph (``use_graph=True``) so the sequential dependency costs launches, not host time.
The entropy coder (zero-run + Huffman) stays on the host and consumes the returned scan indices and
motion vectors; it is outside the hot path."""
from __future__ import annotations

import torch

from .. import _lib
from .._runtime import code, to_device, to_host
from ..quantization import PatchQuant

__all__ = ["ClosedLoopLumaCoder"]


class ClosedLoopLumaCoder:
    """``decode='faithful'`` reproduces the reference's luma decode bit for bit (``symbols2image``
    with a 2-D shape consumes the first Hp*Wp scan blocks of the flat (h w c) list, SURVEY.md
    section 0 item 10); ``decode='luma'`` decodes channel 0 of every block."""

    def __init__(self, quantization_scale=1.0, search_range=4, decode="faithful", me_mode="exact", use_graph=False):
        if decode not in ("faithful", "luma"):
            raise ValueError("decode must be 'faithful' or 'luma'")
        self.quant = PatchQuant(quantization_scale)
        self.search_range = int(search_range)
        self.decode = decode
        self.me_mode = {"auto": _lib.ME_AUTO, "exact": _lib.ME_EXACT, "int": _lib.ME_INT}[me_mode]
        self.use_graph = use_graph
        self._graphs = {}
        # decode == "luma" needs nothing across tiles: one fused kernel per P-frame (the reference frames of a closed loop
        # are reconstructions, so the search is the order-exact one whatever me_mode says -- the vectors are the same)
        import os
        self.fused_step = decode == "luma" and os.environ.get("IVC_CLOSED_LOOP_FUSED", "1") != "0"

    # S sequences in lockstep, time-major [T,S,H,W] float64 on the device: every launch codes frame t of all S
    def _enqueue(self, frames, zz, mv, recon, ws, zero, dtab, tcode):
        L = _lib.lib
        T, S, H, W = frames.shape
        dev = frames.device.index
        sp = torch.cuda.current_stream(frames.device).cuda_stream
        czz = 1 if self.decode == "faithful" else 3
        fsz, zsz, msz = S * H * W * 8, S * (H // 8) * (W // 8) * 192 * 4, S * (H // 8) * (W // 8) * 8
        fp, zp, mp, rp = frames.data_ptr(), zz.data_ptr(), mv.data_ptr(), recon.data_ptr()
        chk = _lib.check
        # I-frames: intra forward with the 3-table broadcast, decode against a zero prediction
        def inverse(z, pred, ref, m, out):
            # 'faithful' reads each frame's first Hp*Wp scan blocks as [Hp,Wp,1,64]: that view is contiguous per
            # frame only, so lockstep batches decode frame by frame; 'luma' decodes the whole batch at once
            groups = [(0, S)] if (czz == 3 or S == 1) else [(i, 1) for i in range(S)]
            for i, n in groups:
                chk(L.ivc_pframe_inverse(dev, sp, z + i * (zsz // S), czz, pred + i * (fsz // S) if pred else None,
                                         ref + i * (fsz // S) if ref else None, m + i * (msz // S) if m else None,
                                         _lib.F64, n, H, W, self.search_range, dtab, tcode, out + i * (fsz // S)),
                    "ivc_pframe_inverse")

        chk(L.ivc_intra_forward(dev, sp, fp, _lib.F64, S, H, W, 1, H * W, dtab, tcode, zp), "ivc_intra_forward")
        inverse(zp, zero.data_ptr(), None, None, rp)
        for t in range(1, T):
            cur, ref, out = fp + t * fsz, rp + (t - 1) * fsz, rp + t * fsz
            z, m = zp + t * zsz, mp + (t - 1) * msz
            if self.fused_step:                   # decode == "luma": search, encoder half and decoder half in ONE kernel
                chk(L.ivc_pframe_step(dev, sp, cur, ref, _lib.F64, S, H, W, self.search_range, dtab, tcode, 3, m, z, out),
                    "ivc_pframe_step")
                continue
            chk(L.ivc_me_full_search(dev, sp, ref, cur, _lib.F64, S, H, W, H * W, H * W, self.search_range,
                                     self.me_mode, m, ws.data_ptr(), ws.numel()), "ivc_me_full_search")
            chk(L.ivc_pframe_forward(dev, sp, cur, ref, m, _lib.F64, S, H, W, self.search_range, dtab, tcode, None, z),
                "ivc_pframe_forward")
            inverse(z, None, ref, m, out)

    def code_sequence(self, frames):
        """frames [T,H,W] (numpy or CUDA tensor, float64) -> dict(zz [T,Hp,Wp,3,64] int32,
        mv [T-1,Hp,Wp,1] int64, recon [T,H,W] float64 = the decoder's reconstructions)."""
        shape = tuple(frames.shape)
        if len(shape) != 3:
            raise ValueError(f"expected [T,H,W] with H,W multiples of 8, got {shape}")
        out = self.code_sequences(frames[None])
        return {k: v[0] for k, v in out.items()}

    def code_sequences(self, frames):
        """S independent sequences coded in LOCKSTEP: frames [S,T,H,W] -> dict(zz [S,T,Hp,Wp,3,64], mv
        [S,T-1,Hp,Wp,1], recon [S,T,H,W]).  Each sequence is its own closed loop (frame t needs its own
        reconstruction t-1), but frame t of all S sequences goes through one launch per kernel, which is what
        fills the GPU when a single frame does not (sequences / GOPs are the unit of parallelism: SURVEY 8e)."""
        f, was_np = to_device(frames)
        f = f.to(torch.float64)
        if f.ndim != 4 or f.shape[2] % 8 or f.shape[3] % 8:
            raise ValueError(f"expected [S,T,H,W] with H,W multiples of 8, got {tuple(f.shape)}")
        f = f.permute(1, 0, 2, 3).contiguous()                                  # time-major
        T, S, H, W = f.shape
        dev = f.device
        _, dtab_t = self.quant._table_on(dev)
        dtab, tcode = dtab_t.data_ptr(), code(dtab_t.dtype)
        key = (T, S, H, W, str(dev), self.decode, self.me_mode, dtab)
        st = self._graphs.get(key) if self.use_graph else None
        if st is None:
            st = {"frames": torch.empty_like(f) if self.use_graph else f,
                  "zz": torch.empty((T, S, H // 8, W // 8, 3, 64), dtype=torch.int32, device=dev),
                  "mv": torch.empty((max(T - 1, 0), S, H // 8, W // 8, 1), dtype=torch.int64, device=dev),
                  "recon": torch.empty((T, S, H, W), dtype=torch.float64, device=dev),
                  "ws": torch.empty(256, dtype=torch.uint8, device=dev),
                  # the I-frames' all-zero prediction: owned by this state, so a captured graph keeps ITS buffer alive
                  "zero": torch.zeros((S, H, W), dtype=torch.float64, device=dev)}
            if self.use_graph:
                st["frames"].copy_(f)
                # warm up once outside capture (sets kernel attributes), then capture the whole sequence
                self._enqueue(st["frames"], st["zz"], st["mv"], st["recon"], st["ws"], st["zero"], dtab, tcode)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(st["frames"], st["zz"], st["mv"], st["recon"], st["ws"], st["zero"], dtab, tcode)
                st["graph"] = g
                self._graphs[key] = st
        if self.use_graph:
            st["frames"].copy_(f)
            st["graph"].replay()
        else:
            self._enqueue(st["frames"], st["zz"], st["mv"], st["recon"], st["ws"], st["zero"], dtab, tcode)
        # back to sequence-major (a copy, so a graph's static buffers are never handed out)
        out = {k: st[k].transpose(0, 1).contiguous() for k in ("zz", "mv", "recon")}
        return {k: to_host(v, was_np) for k, v in out.items()}
