"""Fused per-block coding loop: what ``IntraCodec.image2symbols`` / ``symbols2image``
(ivclab/image/intracodec.py:66-75, :115-124) and the P-frame branch of the video codecs
(ivclab/video/videocodec.py:52-75; exercises/ch4/E4-1.py:257-306) chain together, as ONE kernel
per direction.  Results are identical to calling the four drop-in classes one after another
(and therefore to the reference); the difference is HBM traffic: each sample is read once and
written once.  All entry points accept a single image/frame or a batch (leading axis)."""
from __future__ import annotations

import torch

from . import _lib
from ._runtime import aligned16, code, dev_index, stream_ptr, to_device, to_host
from .quantization import PatchQuant
from .video import MotionCompensator

__all__ = ["IntraBlockCoder", "PFrameBlockCoder", "forward_rgb_multi"]


class IntraBlockCoder:
    """patch -> DCT -> quantize -> zig-zag (``forward``) and its inverse (``inverse``)."""

    def __init__(self, quantization_scale=1.0, luminance=None, chrominance=None):
        self.quant = PatchQuant(quantization_scale, luminance, chrominance)

    @property
    def quantization_scale(self):
        return self.quant.quantization_scale

    def forward(self, img):
        """HWC float64 image ``[H, W, C]`` or batch ``[N, H, W, C]`` (C in {1,3}; a 2-D ``[H, W]``
        plane is treated as C=1) -> int32 scan indices ``[(N,) H/8, W/8, 3, 64]``."""
        t, was_np = to_device(img)
        if t.ndim == 2:
            t = t[..., None]
        batched = t.ndim == 4
        if t.ndim not in (3, 4):
            raise ValueError(f"expected [H,W], [H,W,C] or [N,H,W,C], got shape {tuple(t.shape)}")
        t = aligned16(t.to(torch.float64))
        v = t if batched else t[None]
        N, H, W, C = v.shape
        _, dtab = self.quant._table_on(v.device)
        out = torch.empty((N, H // 8, W // 8, 3, 64), dtype=torch.int32, device=v.device)
        st = _lib.lib.ivc_intra_forward(dev_index(v), stream_ptr(v.device), v.data_ptr(), _lib.F64, N, H, W, C,
                                        H * W * C, dtab.data_ptr(), code(dtab.dtype), out.data_ptr())
        _lib.check(st, "ivc_intra_forward")
        return to_host(out if batched else out[0], was_np)

    def forward_rgb(self, rgb, zr=False):
        """uint8 RGB ``[H, W, 3]`` / ``[N, H, W, 3]`` -> the scan indices of ``forward(rgb2ycbcr(rgb))``
        (intracodec.py:45 + :66-75) in one kernel that reads 3 bytes per pixel.  Frames whose width is
        not a multiple of 16 take the two-kernel route (colour kernel, then ``forward``).
        ``zr=True`` (CUDA tensors, W % 16 == 0): returns ``(zz, counts, masks)`` -- per scan block the number of
        symbols ``ZeroRunCoder.encode`` emits for it and its 64-bit non-zero mask, taken while the block is still in
        shared memory; ``ZeroRunCoder.encode_begin(zz, counts=counts, masks=masks)`` then skips its count pass."""
        t, was_np = to_device(rgb)
        batched = t.ndim == 4
        if t.ndim not in (3, 4) or t.shape[-1] != 3:
            raise ValueError(f"expected [H,W,3] or [N,H,W,3], got shape {tuple(t.shape)}")
        v = t if batched else t[None]
        N, H, W, _ = v.shape
        if v.dtype != torch.uint8 or W % 16:
            if zr:
                raise ValueError("zr=True needs uint8 frames whose width is a multiple of 16")
            from .signal.color import rgb2ycbcr
            out = self.forward(rgb2ycbcr(v))
            return to_host(out if batched else out[0], was_np)
        v = aligned16(v)
        _, dtab = self.quant._table_on(v.device)
        out = torch.empty((N, H // 8, W // 8, 3, 64), dtype=torch.int32, device=v.device)
        if zr:
            if was_np:
                raise ValueError("zr=True is the device-resident pipeline's form: pass CUDA tensors")
            nsb = out.numel() // 64
            counts = torch.empty(nsb, dtype=torch.int32, device=v.device)
            masks = torch.empty(nsb, dtype=torch.int64, device=v.device)
            st = _lib.lib.ivc_intra_forward_rgb8_zr(dev_index(v), stream_ptr(v.device), v.data_ptr(), N, H, W, H * W * 3,
                                                    dtab.data_ptr(), code(dtab.dtype), out.data_ptr(), counts.data_ptr(),
                                                    masks.data_ptr())
            _lib.check(st, "ivc_intra_forward_rgb8_zr")
            return (out if batched else out[0]), counts, masks
        st = _lib.lib.ivc_intra_forward_rgb8(dev_index(v), stream_ptr(v.device), v.data_ptr(), N, H, W, H * W * 3,
                                             dtab.data_ptr(), code(dtab.dtype), out.data_ptr())
        _lib.check(st, "ivc_intra_forward_rgb8")
        return to_host(out if batched else out[0], was_np)

    def inverse(self, zz, to_rgb=False):
        """int32 scan indices ``[(N,) Hp, Wp, C, 64]`` (C in {1,3}) -> float64 ``[(N,) 8Hp, 8Wp, 3]``.
        ``to_rgb=True`` (C = 3) applies ``ycbcr2rgb`` + clip in the decoder's store, the last step of
        ``symbols2image`` (intracodec.py:139-141): the same bits as ``ycbcr2rgb(inverse(zz))`` in one pass."""
        t, was_np = to_device(zz)
        batched = t.ndim == 5
        if t.ndim not in (4, 5) or t.shape[-1] != 64:
            raise ValueError(f"expected [Hp,Wp,C,64] or [N,Hp,Wp,C,64], got shape {tuple(t.shape)}")
        t = aligned16(t.to(torch.int32))
        v = t if batched else t[None]
        N, Hp, Wp, C, _ = v.shape
        _, dtab = self.quant._table_on(v.device)
        out = torch.empty((N, Hp * 8, Wp * 8, 3), dtype=torch.float64, device=v.device)
        if to_rgb and C == 3:
            st = _lib.lib.ivc_intra_inverse_rgb(dev_index(v), stream_ptr(v.device), v.data_ptr(), N, Hp, Wp,
                                                dtab.data_ptr(), code(dtab.dtype), out.data_ptr())
            _lib.check(st, "ivc_intra_inverse_rgb")
            return to_host(out if batched else out[0], was_np)
        st = _lib.lib.ivc_intra_inverse(dev_index(v), stream_ptr(v.device), v.data_ptr(), N, Hp, Wp, C,
                                        dtab.data_ptr(), code(dtab.dtype), out.data_ptr(), _lib.F64)
        _lib.check(st, "ivc_intra_inverse")
        if to_rgb:                                      # luma-only scan blocks: broadcast decode, then the colour kernel
            from .signal.color import ycbcr2rgb
            out = ycbcr2rgb(out)
        return to_host(out if batched else out[0], was_np)


    def inverse_with_distortion(self, zz, original_rgb8, space="rgb", return_reconstruction=False):
        """One rate-distortion point without intermediate images: decode ``zz`` ``[(N,) Hp, Wp, 3, 64]`` and, in the
        same kernel, sum the squared error of every frame against its uint8 RGB original ``[(N,) H, W, 3]``.
        ``space='rgb'``: error between the original and ``ycbcr2rgb(reconstruction)`` -- what
        ``calc_psnr(img, codec.symbols2image(...))`` measures (metrics.py:3-40, intracodec.py:139-141);
        ``space='ycbcr'``: error between ``rgb2ycbcr(original)`` and the reconstruction.
        Returns the float64 CUDA tensor ``sse [N]`` (``mse = sse / (H*W*3)``), plus the float64 YCbCr reconstruction
        when asked for; without it the decoder stores nothing (15 bytes of traffic per pixel)."""
        t, _ = to_device(zz)
        o, _ = to_device(original_rgb8, t.device)
        batched = t.ndim == 5
        if t.ndim not in (4, 5) or t.shape[-1] != 64 or t.shape[-2] != 3:
            raise ValueError(f"expected [Hp,Wp,3,64] or [N,Hp,Wp,3,64], got shape {tuple(t.shape)}")
        t = aligned16(t.to(torch.int32))
        v = t if batched else t[None]
        ov = o if o.ndim == 4 else o[None]
        N, Hp, Wp, _, _ = v.shape
        if o.dtype != torch.uint8 or tuple(ov.shape) != (N, Hp * 8, Wp * 8, 3):
            raise ValueError(f"original must be uint8 RGB of shape {(N, Hp * 8, Wp * 8, 3)}, got {o.dtype} {tuple(o.shape)}")
        if Wp % 2:
            raise ValueError("the fused distortion path needs an image width that is a multiple of 16")
        if space not in ("rgb", "ycbcr"):
            raise ValueError("space must be 'rgb' or 'ycbcr'")
        ov = aligned16(ov)
        _, dtab = self.quant._table_on(v.device)
        out = torch.empty((N, Hp * 8, Wp * 8, 3), dtype=torch.float64, device=v.device) if return_reconstruction else None
        sse = torch.empty(N, dtype=torch.float64, device=v.device)
        wsb = _lib.lib.ivc_intra_inverse_sse_workspace_bytes(N, Hp, Wp)
        ws = torch.empty(wsb, dtype=torch.uint8, device=v.device)
        st = _lib.lib.ivc_intra_inverse_sse(dev_index(v), stream_ptr(v.device), v.data_ptr(), N, Hp, Wp, dtab.data_ptr(),
                                            code(dtab.dtype), out.data_ptr() if out is not None else None, ov.data_ptr(),
                                            Hp * 8 * Wp * 8 * 3, _lib.DIST_RGB if space == "rgb" else _lib.DIST_YCBCR,
                                            ws.data_ptr(), wsb, sse.data_ptr())
        _lib.check(st, "ivc_intra_inverse_sse")
        if not batched:
            sse, out = sse[0], (out[0] if out is not None else None)
        return (sse, out) if return_reconstruction else sse


MAX_FORWARD_TABLES = 16                  # include/ivclab_b200.h IVC_MAX_FORWARD_TABLES


def forward_rgb_multi(coders, rgb, zr=False):
    """``[c.forward_rgb(rgb) for c in coders]`` as ONE kernel: the colour transform and the DCT of a frame do not depend on
    the quantisation scale, so a rate-distortion sweep (exercises/ch4/ex1.py:385-405) transforms each frame once and
    quantises it ``len(coders)`` times.  rgb: CUDA uint8 ``[N, H, W, 3]`` with ``W % 16 == 0``; returns int32
    ``[Q, N, Hp, Wp, 3, 64]`` (``zr=True``: also counts ``[Q, N*Hp*Wp*3]`` int32 and masks int64, as ``forward_rgb``).
    Coders whose tables differ in dtype (float32 for a Python-float scale, float64 for ``np.float64``) and lists longer
    than ``MAX_FORWARD_TABLES`` are handled in groups."""
    coders = list(coders)
    if not isinstance(rgb, torch.Tensor) or not rgb.is_cuda or rgb.dtype != torch.uint8 or rgb.ndim != 4 or rgb.shape[-1] != 3:
        raise ValueError("forward_rgb_multi takes a CUDA uint8 tensor [N, H, W, 3]")
    N, H, W, _ = rgb.shape
    if H % 8 or W % 16:
        raise ValueError("forward_rgb_multi needs H % 8 == 0 and W % 16 == 0")
    v = aligned16(rgb)
    Q = len(coders)
    out = torch.empty((Q, N, H // 8, W // 8, 3, 64), dtype=torch.int32, device=v.device)
    nsb = N * (H // 8) * (W // 8) * 3
    counts = torch.empty((Q, nsb), dtype=torch.int32, device=v.device) if zr else None
    masks = torch.empty((Q, nsb), dtype=torch.int64, device=v.device) if zr else None
    tabs = [c.quant._table_on(v.device)[1] for c in coders]
    i = 0
    while i < Q:
        j = i + 1
        while j < Q and j - i < MAX_FORWARD_TABLES and tabs[j].dtype == tabs[i].dtype:
            j += 1
        dtab = torch.stack(tabs[i:j]).contiguous()
        st = _lib.lib.ivc_intra_forward_rgb8_multi(
            dev_index(v), stream_ptr(v.device), v.data_ptr(), N, H, W, H * W * 3, dtab.data_ptr(), code(dtab.dtype), j - i,
            out[i].data_ptr(), counts[i].data_ptr() if zr else None, masks[i].data_ptr() if zr else None)
        _lib.check(st, "ivc_intra_forward_rgb8_multi")
        i = j
    return (out, counts, masks) if zr else out


class PFrameBlockCoder:
    """Motion estimation, then MC + residual fused into the forward transform and
    prediction + residual fused into the inverse transform (luma planes, float64)."""

    def __init__(self, quantization_scale=1.0, search_range=4, me_mode="auto", luminance=None, chrominance=None):
        self.quant = PatchQuant(quantization_scale, luminance, chrominance)
        self.motion_comp = MotionCompensator(search_range, me_mode)

    @property
    def search_range(self):
        return self.motion_comp.search_range

    def estimate(self, ref, cur):
        """Batched ``compute_motion_vector``: ``[(N,) H, W]`` x2 -> int64 ``[(N,) H/8, W/8, 1]``."""
        r, was_np = to_device(ref)
        c, _ = to_device(cur, r.device)
        batched = r.ndim == 3
        # uint8 planes keep their dtype: the packed-integer kernel stages them without any conversion (float
        # semantics -- the values, not numpy's uint8 wrap-around, which is MotionCompensator's business)
        u8 = r.dtype == torch.uint8 and c.dtype == torch.uint8
        dt = torch.uint8 if u8 else torch.float32 if (r.dtype == torch.float32 and c.dtype == torch.float32) else torch.float64
        r = r.to(dt).contiguous()
        c = c.to(dt).contiguous()
        rv, cv = (r, c) if batched else (r[None], c[None])
        N, H, W = rv.shape
        if H % 8 or W % 8:
            raise IndexError(f"frame sides ({H}, {W}) must be multiples of the 8x8 block size")
        mv = torch.empty((N, H // 8, W // 8, 1), dtype=torch.int64, device=r.device)
        mode = {"auto": _lib.ME_AUTO, "exact": _lib.ME_EXACT, "int": _lib.ME_INT}[self.motion_comp.me_mode]
        ws, ws_bytes = None, 0
        if mode != _lib.ME_EXACT:
            # wide searches of float64 frames: room for uint8 planes, so that the frames are converted once
            planes = dt == torch.float64 and int(self.search_range) >= 8
            ws_bytes = (_lib.lib.ivc_me_workspace_bytes_planes if planes else _lib.lib.ivc_me_workspace_bytes)(N, H, W)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=r.device)
        st = _lib.lib.ivc_me_full_search(dev_index(rv), stream_ptr(rv.device), rv.data_ptr(), cv.data_ptr(), code(dt),
                                         N, H, W, H * W, H * W, int(self.search_range), mode, mv.data_ptr(),
                                         ws.data_ptr() if ws is not None else None, ws_bytes)
        _lib.check(st, "ivc_me_full_search")
        return to_host(mv if batched else mv[0], was_np)

    def estimate_forward(self, ref, cur, channels=3, zr=False):
        """Search and encoder half in one call: ``mv = estimate(ref, cur)`` and ``zz = forward(cur, ref, mv)``
        (videocodec.py:52 + :68-71) -> ``(mv [(N,) Hp, Wp, 1] int64, zz [(N,) Hp, Wp, channels, 64] int32)``.
        For +-4 searches of integer-valued float64 frames a single kernel does both -- its tiles code their blocks
        from the bytes the search staged, so the frames are read once; every other case (other search ranges,
        non-integer frames -- detected on the device in ``me_mode='auto'`` --, ``me_mode='exact'``) runs the two
        stand-alone kernels.  Same results either way.  uint8 planes (the frames' values) are taken as they are when the
        fused kernel applies.  ``zr=True`` (CUDA tensors): returns ``(mv, zz, counts, masks)`` with the zero-run coder's
        per-block symbol counts and non-zero masks (see :meth:`IntraBlockCoder.forward_rgb`)."""
        r, was_np = to_device(ref)
        c, _ = to_device(cur, r.device)
        batched = r.ndim == 3
        # uint8 planes (the frames' values) stay bytes where the fused kernel can take them: nothing is converted
        u8 = r.dtype == torch.uint8 and c.dtype == torch.uint8 and int(self.search_range) == 4 and self.motion_comp.me_mode != "exact"
        if u8:
            r, c = r.contiguous(), c.contiguous()
        else:
            r = aligned16(r.to(torch.float64))
            c = aligned16(c.to(torch.float64))
        rv, cv = (r, c) if batched else (r[None], c[None])
        N, H, W = rv.shape
        if H % 8 or W % 8:
            raise IndexError(f"frame sides ({H}, {W}) must be multiples of the 8x8 block size")
        if channels not in (2, 3):
            raise ValueError("channels must be 3 (the reference's layout) or 2 (channels 0 and 1)")
        _, dtab = self.quant._table_on(r.device)
        mv = torch.empty((N, H // 8, W // 8, 1), dtype=torch.int64, device=r.device)
        zz = torch.empty((N, H // 8, W // 8, channels, 64), dtype=torch.int32, device=r.device)
        mode = {"auto": _lib.ME_AUTO, "exact": _lib.ME_EXACT, "int": _lib.ME_INT}[self.motion_comp.me_mode]
        ws_bytes = _lib.lib.ivc_me_workspace_bytes(N, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=r.device)
        args = (dev_index(rv), stream_ptr(rv.device), cv.data_ptr(), rv.data_ptr(), _lib.U8 if u8 else _lib.F64, N, H, W,
                int(self.search_range), mode, dtab.data_ptr(), code(dtab.dtype), channels, mv.data_ptr(), zz.data_ptr(),
                ws.data_ptr(), ws_bytes)
        if zr:
            if was_np:
                raise ValueError("zr=True is the device-resident pipeline's form: pass CUDA tensors")
            nsb = zz.numel() // 64
            counts = torch.empty(nsb, dtype=torch.int32, device=r.device)
            masks = torch.empty(nsb, dtype=torch.int64, device=r.device)
            _lib.check(_lib.lib.ivc_pframe_search_forward_zr(*args, counts.data_ptr(), masks.data_ptr()), "ivc_pframe_search_forward_zr")
            return (mv, zz, counts, masks) if batched else (mv[0], zz[0], counts, masks)
        st = _lib.lib.ivc_pframe_search_forward(*args)
        _lib.check(st, "ivc_pframe_search_forward")
        if not batched:
            mv, zz = mv[0], zz[0]
        return to_host(mv, was_np), to_host(zz, was_np)

    def forward(self, cur, ref, mv, return_prediction=False, channels=3):
        """residual = cur - MC(ref, mv); -> scan indices ``[(N,) Hp, Wp, 3, 64]`` (and the prediction).
        ``channels=2`` stores channels 0 and 1 only (``[(N,) Hp, Wp, 2, 64]``): numpy broadcasting quantises the one
        luma residual against ``[lum, chrom, chrom]`` (patchquant.py:40,59), so channel 2 repeats channel 1 bit for
        bit whenever the two chrominance tables are the same array -- a pipeline that ships symbols to the host
        need not compute, store or send it (``inverse`` reads channel 0 of either layout)."""
        c, was_np = to_device(cur)
        r, _ = to_device(ref, c.device)
        m, _ = to_device(mv, c.device)
        batched = c.ndim == 3
        c = aligned16(c.to(torch.float64))
        r = r.to(torch.float64).contiguous()
        m = m.to(torch.int64).contiguous()
        cv = c if batched else c[None]
        N, H, W = cv.shape
        _, dtab = self.quant._table_on(c.device)
        if channels not in (2, 3):
            raise ValueError("channels must be 3 (the reference's layout) or 2 (channels 0 and 1)")
        zz = torch.empty((N, H // 8, W // 8, channels, 64), dtype=torch.int32, device=c.device)
        pred = torch.empty_like(cv) if return_prediction else None
        st = _lib.lib.ivc_pframe_forward_ch(dev_index(c), stream_ptr(c.device), cv.data_ptr(), r.data_ptr(), m.data_ptr(),
                                            _lib.F64, N, H, W, int(self.search_range), dtab.data_ptr(), code(dtab.dtype),
                                            pred.data_ptr() if pred is not None else None, channels, zz.data_ptr())
        _lib.check(st, "ivc_pframe_forward_ch")
        zz = zz if batched else zz[0]
        if return_prediction:
            return to_host(zz, was_np), to_host(pred if batched else pred[0], was_np)
        return to_host(zz, was_np)

    def inverse(self, zz, pred=None, ref=None, mv=None):
        """recon = prediction + idct(dequantize(unflatten(zz[..., 0, :]))) with the luminance table
        (what ``symbols2image(...)[..., 0]`` returns, E4-1.py:284-306).  Give either ``pred`` or ``(ref, mv)``."""
        z, was_np = to_device(zz)
        batched = z.ndim == 5
        z = aligned16(z.to(torch.int32))
        zv = z if batched else z[None]
        N, Hp, Wp, Czz, _ = zv.shape
        H, W = Hp * 8, Wp * 8
        _, dtab = self.quant._table_on(z.device)
        p = r = m = None
        if pred is not None:
            p = aligned16(to_device(pred, z.device)[0].to(torch.float64))
        else:
            if ref is None or mv is None:
                raise ValueError("give either pred, or ref and mv")
            r = to_device(ref, z.device)[0].to(torch.float64).contiguous()
            m = to_device(mv, z.device)[0].to(torch.int64).contiguous()
        out = torch.empty((N, H, W), dtype=torch.float64, device=z.device)
        st = _lib.lib.ivc_pframe_inverse(dev_index(z), stream_ptr(z.device), zv.data_ptr(), Czz,
                                         p.data_ptr() if p is not None else None,
                                         r.data_ptr() if r is not None else None,
                                         m.data_ptr() if m is not None else None,
                                         _lib.F64, N, H, W, int(self.search_range), dtab.data_ptr(), code(dtab.dtype),
                                         out.data_ptr())
        _lib.check(st, "ivc_pframe_inverse")
        return to_host(out if batched else out[0], was_np)
