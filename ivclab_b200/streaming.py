"""Host <-> device streaming of the coding loop: while chunk k is being coded, chunk k+1 is uploaded
and the results of chunk k-1 are downloaded, on three CUDA streams with pinned staging buffers.

This is the form in which the path is fed from HOST memory (what ``bench.py`` reports as ``e2e``):
uint8 RGB frames and uint8 luma planes go up -- 5 bytes per pixel with explicit references, 4 in sequence
mode, 3 when the luma plane is the Y channel of the RGB frame and is derived on the device, as the
reference's video codecs do (``rgb2ycbcr(frame)[..., 0]``, videocodec.py:38) -- and zero-run symbol
streams, motion vectors and squared errors come back; the colour transform, both transform loops, the motion search,
the zero-run coder and the error reduction all run on the device in between.  Frames are independent
(intra) or frame pairs (inter), so chunks never depend on each other.

Two things keep the device busy rather than waiting for Python:
* the ~25 launches that code one chunk are captured ONCE per input slot as a CUDA graph (the slots and the
  graph's outputs are static buffers) and replayed with a single call per chunk;
* the only host-dependent step -- sizing the symbol streams -- is software-pipelined: the host waits for the
  two stream lengths of chunk k-1 (copied to pinned memory inside the graph) only after chunk k's graph has
  been enqueued."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _lib  # noqa: F401
from .codec import IntraBlockCoder, PFrameBlockCoder
from .signal.color import luma8_from_rgb8
from .entropy import ZeroRunCoder
from .utils.metrics import frame_sse

__all__ = ["StreamedCoder"]


class _Slot:
    """Static device buffers of one pipeline slot, the graph that codes them and that graph's outputs."""
    __slots__ = ("rgb", "cur", "ref", "graphs", "totals", "derive")


class StreamedCoder:
    def __init__(self, quantization_scale=1.0, search_range=4, me_mode="auto", chunk_frames=2, device=None, use_graph=True,
                 slots=3, compute_streams=2, symbol_dtype="auto", ramp=(), inter_channels="auto", mv_dtype="auto"):
        self.intra = IntraBlockCoder(quantization_scale)
        self.pframe = PFrameBlockCoder(quantization_scale, search_range, me_mode)
        self.zr = ZeroRunCoder()
        self.chunk = int(chunk_frames)
        # The P-frame stream of the reference carries the luma residual quantised against [lum, chrom, chrom]
        # (patchquant.py:40,59): channel 2 repeats channel 1 bit for bit.  "auto" codes and sends channels 0 and 1
        # only whenever the two chrominance tables really are the same bits (`expand_inter` rebuilds the
        # reference's three-channel stream on the host); 3 forces the reference's layout.
        t3 = np.asarray(self.intra.quant.get_quantization_table())
        same_chroma = t3.shape == (3, 8, 8) and t3[1].tobytes() == t3[2].tobytes()
        if inter_channels == "auto":
            inter_channels = 2 if same_chroma else 3
        if inter_channels not in (2, 3) or (inter_channels == 2 and not same_chroma):
            raise ValueError("inter_channels must be 'auto', 3, or 2 with identical chrominance tables")
        self.inter_channels = int(inter_channels)
        # motion vectors are indices below (2 sr + 1)^2: int16 on the wire when that fits (the reference's dtype is
        # int64, `dtype=int` at motion.py:26; `.astype(np.int64)` restores it)
        span2 = (2 * int(search_range) + 1) ** 2
        if mv_dtype == "auto":
            mv_dtype = torch.int16 if span2 <= 32767 else torch.int64
        if mv_dtype not in (torch.int16, torch.int32, torch.int64) or (mv_dtype == torch.int16 and span2 > 32767):
            raise ValueError("mv_dtype must be 'auto', torch.int64, torch.int32, or torch.int16 with (2 sr + 1)^2 <= 32767")
        self.mv_dtype = mv_dtype
        import os
        self.fused_count = os.environ.get("IVC_STREAM_FUSED_COUNT", "1") != "0"    # A/B switch: zero-run counts from the forward kernels
        # Symbol streams are the bulk of the download.  The inputs are 8-bit images, so every DCT coefficient is
        # bounded by 8 * 255 and every quantised value by 2040 / min(table): when that (and the EOB marker) fits 16
        # bits, "auto" sends int16 symbols -- a lossless transfer format, half the bytes; torch.int32 forces the
        # reference's dtype.
        tab = np.asarray(self.intra.quant.get_quantization_table(), dtype=np.float64)
        fits16 = bool(np.all(tab > 0)) and 2040.0 / float(tab.min()) < 32000 and abs(int(self.zr.EOB)) < 32768
        if symbol_dtype == "auto":
            symbol_dtype = torch.int16 if fits16 else torch.int32
        if symbol_dtype not in (torch.int16, torch.int32) or (symbol_dtype == torch.int16 and not fits16):
            raise ValueError("symbol_dtype must be 'auto', torch.int32, or torch.int16 with a table that keeps symbols in 16 bits")
        self.symbol_dtype = symbol_dtype
        self.use_graph = bool(use_graph)
        self.ramp = tuple(ramp) if ramp else ()   # sizes of shorter chunks at both ends of a run (see _schedule)
        self.nslots = max(2, int(slots))          # input buffers in rotation: uploads run ahead of the coder by nslots-1 chunks
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._s_in, self._s_cmp, self._s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        # chunks alternate between two compute streams: the tail of one chunk's kernels overlaps the head of the next
        self._s_cmp2 = torch.cuda.Stream(self.device) if compute_streams > 1 else self._s_cmp
        self._host = None
        self._slots = None         # ((C, H, W), [_Slot, _Slot])
        self.trace = None          # set to [] to collect (label, chunk, start event, end event, host t0, host t1) per stage

    # ---- buffers -------------------------------------------------------------------------------
    def _host_buffers(self, F, H, W):
        key = (F, H, W, self.symbol_dtype, self.mv_dtype)
        if self._host is None or self._host[0] != key:
            nb = (H // 8) * (W // 8)
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._host = (key, {
                # symbol streams: room for 24 symbols per block to start with (a block emits 1..97); grown on demand
                "sym_intra": pin(F * nb * 3 * 24, self.symbol_dtype), "sym_inter": pin(F * nb * 3 * 24, self.symbol_dtype),
                "mv": pin((F, H // 8, W // 8, 1), self.mv_dtype), "sse": pin((2, F), torch.float64)})
        return self._host[1]

    def _device_slots(self, C, H, W, seq, derive=False):
        key = (C, H, W, seq, derive)
        if self._slots is None or self._slots[0] != key:
            slots = []
            for _ in range(self.nslots):
                s = _Slot()
                s.derive = derive
                # derived luma: C + 1 RGB frames as well, frame 0 = the frame before the chunk (the source of its reference plane)
                s.rgb = torch.empty((C + 1 if derive else C, H, W, 3), dtype=torch.uint8, device=self.device)
                # sequence mode: one luma buffer of C+1 frames, frame 0 = the frame before the chunk (its reference)
                s.cur = torch.empty((C + 1 if seq else C, H, W), dtype=torch.uint8, device=self.device)
                s.ref = None if seq else torch.empty((C, H, W), dtype=torch.uint8, device=self.device)
                s.graphs = {}                       # frames in the chunk -> (CUDA graph, its static outputs)
                s.totals = torch.zeros(2, dtype=torch.int64).pin_memory()      # the two stream lengths of the chunk in flight
                slots.append(s)
            self._slots = (key, slots)
        return self._slots[1]

    def _mark(self, label, k, stream):
        """Tracing aid: returns a function that closes the interval opened here (no-op unless ``self.trace`` is a list)."""
        if self.trace is None:
            return lambda: None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        t0 = time.perf_counter()

        def close():
            b.record(stream)
            self.trace.append((label, k, a, b, t0, time.perf_counter()))
        return close

    # ---- one chunk on the device -------------------------------------------------------------------
    def _code(self, s: _Slot, n: int):
        """Everything of a chunk that does not need a stream length on the host (current stream = compute stream)."""
        d_rgb = s.rgb[1:n + 1] if s.derive else s.rgb[:n]
        if s.ref is None:                              # sequence mode: frame t is predicted from frame t-1
            luma8 = s.cur[:n + 1]
            if s.derive:                               # the planes are the rounded Y channel of the RGB frames (videocodec.py:38)
                luma = torch.empty(luma8.shape, dtype=torch.float64, device=luma8.device)
                luma8_from_rgb8(s.rgb[:n + 1], out=luma8, out_f64=luma)
            else:
                luma = luma8.double()
            d_ref, d_cur, r8, c8 = luma[:n], luma[1:], luma8[:n], luma8[1:]
        else:
            r8, c8 = s.ref[:n], s.cur[:n]
            d_cur, d_ref = c8.double(), r8.double()
        # the forward kernels hand the zero-run coder its per-block symbol counts and non-zero masks (taken from their
        # staging buffers): no pass re-reads the indices from HBM to count
        if self.fused_count:
            zz, cnt_i, msk_i = self.intra.forward_rgb(d_rgb, zr=True)
        else:
            zz, cnt_i, msk_i = self.intra.forward_rgb(d_rgb), None, None
        pend_i = self.zr.encode_begin(zz, total_host=s.totals[0:1], record=False, counts=cnt_i, masks=msk_i)
        sse_i = self.intra.inverse_with_distortion(zz, d_rgb, space="ycbcr")   # decode + error in one kernel, nothing stored
        if int(self.pframe.search_range) == 4 and self.pframe.motion_comp.me_mode != "exact":
            # one kernel: the search reads the uint8 planes as they arrived and codes its tiles' blocks from the same bytes
            if self.fused_count:
                mv, zzp, cnt_p, msk_p = self.pframe.estimate_forward(r8, c8, channels=self.inter_channels, zr=True)
            else:
                (mv, zzp), cnt_p, msk_p = self.pframe.estimate_forward(r8, c8, channels=self.inter_channels), None, None
        else:
            mv = self.pframe.estimate(r8, c8)
            zzp = self.pframe.forward(d_cur, d_ref, mv, channels=self.inter_channels)
            cnt_p = msk_p = None
        pend_p = self.zr.encode_begin(zzp, total_host=s.totals[1:2], record=False, counts=cnt_p, masks=msk_p)
        recp = self.pframe.inverse(zzp, ref=d_ref, mv=mv)
        sse_p = frame_sse(d_cur, recp)
        if self.mv_dtype != torch.int64:
            mv = mv.to(self.mv_dtype)
        return pend_i, pend_p, mv, sse_i, sse_p

    def _launch(self, s: _Slot, n: int):
        if not self.use_graph:
            return self._code(s, n)
        if n not in s.graphs:                           # one graph per slot and chunk size, captured on first use
            self._code(s, n)                            # warm-up outside capture (function attributes, workspaces)
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.current_stream(self.device)):
                out = self._code(s, n)
            s.graphs[n] = (g, out)
        g, out = s.graphs[n]
        g.replay()
        return out

    def _schedule(self, F):
        """Chunk boundaries: full chunks in the middle, optionally shorter chunks (``self.ramp``, e.g. ``(2, 2)``) at
        both ends -- the first upload and the last download are the part of the pipeline nothing overlaps with."""
        C = self.chunk
        ramp = [min(int(n), C) for n in self.ramp if int(n) > 0]
        if not ramp or F < 2 * sum(ramp) + C:
            sizes = [C] * (F // C) + ([F % C] if F % C else [])
        else:
            body = F - 2 * sum(ramp)
            sizes = ramp + [C] * (body // C) + ([body % C] if body % C else []) + ramp[::-1]
        bounds, lo = [], 0
        for n in sizes:
            bounds.append((lo, lo + n))
            lo += n
        return bounds

    # ---- the pipeline ----------------------------------------------------------------------------
    def run(self, rgb, cur=None, ref=None, first_ref=None):
        """rgb [F,H,W,3] uint8, cur [F,H,W] uint8 luma planes -- pinned host tensors (numpy arrays are accepted and
        pinned once).  The P-frame references are either given frame by frame (``ref`` [F,H,W]) or implied by the
        sequence (``ref=None``): frame t is predicted from frame t-1 and frame 0 from ``first_ref`` [H,W] -- every
        luma frame then crosses PCIe once instead of twice.  ``cur=None`` (sequence mode only): the luma plane of a
        frame is ``luma8_from_rgb8`` of its RGB frame, derived on the device -- only the RGB frames cross PCIe;
        ``first_ref`` is then the RGB frame [H,W,3] before frame 0.  Returns host-side results: ``sym_intra`` /
        ``sym_inter`` (symbol streams in ``self.symbol_dtype`` -- int16 by default, a lossless transfer format --
        with per-chunk lengths ``len_intra`` / ``len_inter``; the inter stream holds ``self.inter_channels`` scan
        blocks per image block, see :meth:`expand_inter`), ``mv`` [F,Hp,Wp,1] in ``self.mv_dtype``, ``sse`` [2,F]
        (intra on YCbCr, inter on luma).  The arrays are VIEWS of this coder's persistent pinned buffers: they are
        valid until the next ``run`` with the same geometry -- copy what must outlive it."""
        pinned = lambda x: (torch.from_numpy(x) if isinstance(x, np.ndarray) else x)
        seq = ref is None
        derive = cur is None
        if seq and first_ref is None:
            raise ValueError("sequence mode (ref=None) needs first_ref, the reference of frame 0")
        if derive and not seq:
            raise ValueError("cur=None (luma derived from the RGB frames) needs sequence mode (ref=None)")
        rgb, ref = (t if t.is_pinned() else t.pin_memory() for t in map(pinned, (rgb, first_ref if seq else ref)))
        if not derive:
            cur = pinned(cur)
            cur = cur if cur.is_pinned() else cur.pin_memory()
        F, H, W, _ = rgb.shape
        if derive and tuple(ref.shape) != (H, W, 3):
            raise ValueError(f"cur=None: first_ref must be the RGB frame [H,W,3] before frame 0, got {tuple(ref.shape)}")
        hb = self._host_buffers(F, H, W)
        if F == 0:
            return {"sym_intra": hb["sym_intra"][:0], "sym_inter": hb["sym_inter"][:0], "len_intra": [], "len_inter": [],
                    "mv": hb["mv"], "sse": hb["sse"], "h2d_bytes": 0, "d2h_bytes": 0}
        C = self.chunk
        bounds = self._schedule(F)
        nchunks = len(bounds)
        slots = self._device_slots(C, H, W, seq, derive)
        S = self.nslots
        ev_in = [torch.cuda.Event() for _ in range(nchunks)]
        ev_cmp = [torch.cuda.Event() for _ in range(nchunks)]
        ev_fin = [torch.cuda.Event() for _ in range(nchunks)]
        ev_out = [torch.cuda.Event() for _ in range(nchunks)]
        lens_i, lens_p = [], []
        off = [0, 0]
        pending = {}                                  # chunk -> device results whose symbol streams are not written yet

        def upload(k):
            lo, hi = bounds[k]
            with torch.cuda.stream(self._s_in):
                if k >= S:
                    self._s_in.wait_event(ev_cmp[k - S])          # the slot's previous chunk has been consumed
                s = slots[k % S]
                done = self._mark("h2d", k, self._s_in)
                n_prev = bounds[k - 1][1] - bounds[k - 1][0] if k else 0
                prev = slots[(k - 1) % S]                         # its frames were uploaded by this very stream
                if derive:                                        # RGB only; frame 0 of the slot = the frame before the chunk
                    s.rgb[1:1 + hi - lo].copy_(rgb[lo:hi], non_blocking=True)
                    if k == 0:
                        s.rgb[0].copy_(ref, non_blocking=True)
                    else:
                        s.rgb[0].copy_(prev.rgb[n_prev])          # device-to-device
                elif seq:
                    s.rgb[:hi - lo].copy_(rgb[lo:hi], non_blocking=True)
                    s.cur[1:1 + hi - lo].copy_(cur[lo:hi], non_blocking=True)
                    if k == 0:
                        s.cur[0].copy_(ref.reshape(H, W), non_blocking=True)
                    else:                                         # the last plane of the previous chunk: device-to-device
                        s.cur[0].copy_(prev.cur[n_prev])
                else:
                    s.rgb[:hi - lo].copy_(rgb[lo:hi], non_blocking=True)
                    s.cur[:hi - lo].copy_(cur[lo:hi], non_blocking=True)
                    s.ref[:hi - lo].copy_(ref[lo:hi], non_blocking=True)
                done()
                ev_in[k].record(self._s_in)

        def compute(k):
            n = bounds[k][1] - bounds[k][0]
            sc = self._s_cmp if k % 2 == 0 else self._s_cmp2
            with torch.cuda.stream(sc):
                sc.wait_event(ev_in[k])
                if k >= S:
                    sc.wait_event(ev_fin[k - S])                  # the slot's previous results have been picked up
                done = self._mark("code", k, sc)
                pending[k] = self._launch(slots[k % S], n)
                done()
                ev_cmp[k].record(sc)                               # the input slot may be overwritten from here on

        def finish(k):
            """Write chunk k's symbol streams (their lengths have arrived by now) and send the results home."""
            lo, hi = bounds[k]
            pend_i, pend_p, mv, sse_i, sse_p = pending.pop(k)
            ev_cmp[k].synchronize()                                # waits for TWO numbers, with chunk k+1 already queued
            sc = self._s_cmp if k % 2 == 0 else self._s_cmp2
            with torch.cuda.stream(sc):
                tr = self._mark("symbols", k, sc)
                sym_i = self.zr.encode_finish(pend_i, self.symbol_dtype)
                sym_p = self.zr.encode_finish(pend_p, self.symbol_dtype)
                mv, sse_i, sse_p = mv.clone(), sse_i.clone(), sse_p.clone()     # frees the slot's (static) result buffers
                tr()
                ev_fin[k].record(sc)
            for name, o, sym in (("sym_intra", off[0], sym_i), ("sym_inter", off[1], sym_p)):
                if o + sym.numel() > hb[name].numel():             # rare: denser streams than provisioned
                    self._s_out.synchronize()                      # earlier downloads into the old buffer are complete
                    grown = torch.empty(max(2 * hb[name].numel(), o + sym.numel()), dtype=self.symbol_dtype).pin_memory()
                    grown[:o].copy_(hb[name][:o])
                    hb[name] = grown
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev_fin[k])
                for t in (sym_i, sym_p, mv, sse_i, sse_p):
                    t.record_stream(self._s_out)
                tr = self._mark("d2h", k, self._s_out)
                hb["sym_intra"][off[0]:off[0] + sym_i.numel()].copy_(sym_i, non_blocking=True)
                hb["sym_inter"][off[1]:off[1] + sym_p.numel()].copy_(sym_p, non_blocking=True)
                hb["mv"][lo:hi].copy_(mv, non_blocking=True)
                hb["sse"][0, lo:hi].copy_(sse_i, non_blocking=True)
                hb["sse"][1, lo:hi].copy_(sse_p, non_blocking=True)
                tr()
                ev_out[k].record(self._s_out)
            lens_i.append(sym_i.numel())
            lens_p.append(sym_p.numel())
            off[0] += sym_i.numel()
            off[1] += sym_p.numel()

        # software pipeline on the host: chunk k is enqueued BEFORE the host waits for the stream lengths of chunk
        # k-1, so the device always has work queued and the host never waits on fresh work
        for k in range(min(S - 1, nchunks)):
            upload(k)
        for k in range(nchunks):
            if k + S - 1 < nchunks:
                upload(k + S - 1)
            compute(k)
            if k >= 1:
                finish(k - 1)
        finish(nchunks - 1)
        ev_out[-1].synchronize()
        self._s_cmp.synchronize()
        self._s_cmp2.synchronize()
        return {"sym_intra": hb["sym_intra"][:off[0]], "sym_inter": hb["sym_inter"][:off[1]], "len_intra": lens_i,
                "len_inter": lens_p, "mv": hb["mv"], "sse": hb["sse"],
                "h2d_bytes": rgb.numel() + (0 if derive else cur.numel()) + ref.numel(),   # ref = one frame in sequence mode
                "d2h_bytes": (off[0] + off[1]) * hb["sym_intra"].element_size() + hb["mv"].numel() * hb["mv"].element_size()
                             + hb["sse"].numel() * 8}

    # ---- the transfer format of the inter stream -----------------------------------------------------
    @staticmethod
    def expand_inter(symbols, end_of_block=4000):
        """Host side: rebuild the reference's three-channel P-frame symbol stream (what ``ZeroRunCoder.encode`` of the
        ``[Hp, Wp, 3, 64]`` indices yields, E4-1.py:271-279) from the two-channel stream this coder ships: blocks
        come in (h w c) order, channel 2 repeats channel 1, so every second block is emitted twice.  numpy, vectorised;
        an EOB is a block end unless it is the run length after a zero marker (zerorun.py:25-35)."""
        s = np.asarray(symbols.numpy() if isinstance(symbols, torch.Tensor) else symbols)
        if s.size == 0:
            return s.astype(np.int32)
        prev0 = np.concatenate(([False], s[:-1] == 0))
        runlen = np.zeros(s.size, dtype=bool)                 # a slot after a zero marker is a run length ...
        idx = np.flatnonzero(prev0)
        # ... unless that "marker" is itself a run length: impossible, run lengths are >= 1
        runlen[idx] = True
        ends = np.flatnonzero((s == end_of_block) & ~runlen)   # last symbol of every block
        starts = np.concatenate(([0], ends[:-1] + 1))
        if ends.size % 2:
            raise ValueError("the two-channel inter stream must hold an even number of blocks")
        # output order per image block: block 2k, block 2k+1, block 2k+1
        sel = np.stack([np.arange(0, ends.size, 2), np.arange(1, ends.size, 2), np.arange(1, ends.size, 2)], axis=1).reshape(-1)
        lens = (ends - starts + 1)[sel]
        out_off = np.concatenate(([0], np.cumsum(lens)))
        src = np.repeat(starts[sel] - out_off[:-1], lens) + np.arange(out_off[-1])
        return s[src].astype(np.int32)
