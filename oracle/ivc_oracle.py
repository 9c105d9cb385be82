"""numpy restatement of the ivclab per-block coding loop (TEST INFRASTRUCTURE).

Every function cites the reference file:line (under /root/reference) whose
behaviour it restates.  The restatement is deliberately *independent* of the
reference's code structure: the 8-point DCT is written out op-for-op the way
scipy's ducc0 back-end evaluates it (the reference calls ``scipy.fft.dct`` at
ivclab/signal/dct.py:24,26,42,44), so the oracle is bit-identical to scipy and
is at the same time the executable specification of the arithmetic the CUDA
kernels perform.  ``oracle/gen_golden.py`` pins all of it against the real
reference modules.

Nothing here is imported by the product package ``ivclab_b200``.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "DUCC_TW", "DUCC_WA", "SQRT2", "ZIGZAG_ORDER", "ZIGZAG_SCAN",
    "LUMINANCE", "CHROMINANCE",
    "ducc_unity_root", "derive_ducc_constants",
    "dct2_8", "dct3_8", "dct8x8_forward", "dct8x8_inverse",
    "quant_table", "quantize", "dequantize",
    "zigzag_flatten", "zigzag_unflatten", "patch", "unpatch",
    "np_sum64", "me_full_search_loops", "me_full_search", "mc_reconstruct",
    "intra_forward", "intra_inverse", "pframe_forward", "pframe_inverse",
    "rgb2ycbcr", "ycbcr2rgb", "calc_mse", "calc_psnr",
    "zerorun_encode", "zerorun_decode",
    "smooth_noise_rgb", "smooth_noise_luma", "moving_sequence",
]

# --------------------------------------------------------------------------
# 1. ducc0's constants for the length-8 DCT-II / DCT-III
# --------------------------------------------------------------------------
# scipy.fft.dct(x, type=2, norm='ortho') on length 8 runs ducc0's T_dcst23:
# a length-8 real FFT (factors 2,4) wrapped by pre/post twiddling with
# twiddle[i] = Re(UnityRoots(32)[i+1]).  ducc0 evaluates those roots in
# *double* (cos/sin of a rounded angle pi/128*8i), so several of them are one
# ulp away from the correctly rounded cosine.  The values are frozen here as
# hex doubles (the CUDA kernels carry the same literals); ``derive_ducc_
# constants`` re-derives them with the same recipe and tests assert equality.
DUCC_TW = np.array([float.fromhex(h) for h in (
    "0x1.f6297cff75cb0p-1",   # ~cos(1*pi/16)
    "0x1.d906bcf328d46p-1",   # ~cos(2*pi/16)
    "0x1.a9b66290ea1a3p-1",   # ~cos(3*pi/16)
    "0x1.6a09e667f3bccp-1",   # ~cos(4*pi/16)   (1 ulp below correctly rounded)
    "0x1.1c73b39ae68c8p-1",   # ~cos(5*pi/16)   (1 ulp below)
    "0x1.87de2a6aea963p-2",   # ~cos(6*pi/16)   (1 ulp below)
    "0x1.8f8b83c69a60ap-3",   # ~cos(7*pi/16)   (3 ulp below)
)], dtype=np.float64)
# radix-2 pass twiddle of the length-8 real FFT: UnityRoots(8)[1] = (re, im)
DUCC_WA = np.array([float.fromhex("0x1.6a09e667f3bccp-1"),
                    float.fromhex("0x1.6a09e667f3bcdp-1")], dtype=np.float64)
SQRT2 = float(np.float64(np.longdouble("1.414213562373095048801688724209698")))


def ducc_unity_root(n: int):
    """Emulate ducc0's ``UnityRoots<double>(n)`` -> callable idx -> (re, im).

    Two-table scheme: v1[i] = calc(i), v2[i] = calc(i*(mask+1)), root(idx) =
    v1[idx&mask]*v2[idx>>shift] (complex product in double).  ``calc`` takes
    cos/sin of ``x * ang`` with ``ang = double(0.25L*pi/n)`` and x = 8*idx
    folded into the first octant.
    """
    pi_ld = np.longdouble("3.141592653589793238462643383279502884197")
    ang = float(np.longdouble(0.25) * pi_ld / np.longdouble(n))

    def calc(x):
        x <<= 3
        if x < 4 * n:
            if x < 2 * n:
                if x < n:
                    return (math.cos(x * ang), math.sin(x * ang))
                return (math.sin((2 * n - x) * ang), math.cos((2 * n - x) * ang))
            x -= 2 * n
            if x < n:
                return (-math.sin(x * ang), math.cos(x * ang))
            return (-math.cos((2 * n - x) * ang), math.sin((2 * n - x) * ang))
        x = 8 * n - x
        if x < 2 * n:
            if x < n:
                return (math.cos(x * ang), -math.sin(x * ang))
            return (math.sin((2 * n - x) * ang), -math.cos((2 * n - x) * ang))
        x -= 2 * n
        if x < n:
            return (-math.sin(x * ang), -math.cos(x * ang))
        return (-math.cos((2 * n - x) * ang), -math.sin((2 * n - x) * ang))

    nval = (n + 2) // 2
    shift = 1
    while (1 << shift) * (1 << shift) < nval:
        shift += 1
    mask = (1 << shift) - 1
    v1 = [(1.0, 0.0)] + [calc(i) for i in range(1, mask + 1)]
    nv2 = (nval + mask) // (mask + 1)
    v2 = [(1.0, 0.0)] + [calc(i * (mask + 1)) for i in range(1, nv2)]

    def root(idx):
        neg = False
        if 2 * idx > n:
            idx = n - idx
            neg = True
        a, b = v1[idx & mask], v2[idx >> shift]
        re = a[0] * b[0] - a[1] * b[1]
        im = a[0] * b[1] + a[1] * b[0]
        return (re, -im if neg else im)

    return root


def derive_ducc_constants():
    """Re-derive (DUCC_TW, DUCC_WA) with ducc0's recipe (depends on libm)."""
    r32, r8 = ducc_unity_root(32), ducc_unity_root(8)
    tw = np.array([r32(i + 1)[0] for i in range(7)], dtype=np.float64)
    wa = np.array(r8(1), dtype=np.float64)
    return tw, wa


def _consts(T):
    tw = [T(v) for v in DUCC_TW]
    return tw, T(DUCC_WA[0]), T(DUCC_WA[1]), T(np.longdouble("1.414213562373095048801688724209698"))


# --------------------------------------------------------------------------
# 2. 8-point DCT-II / DCT-III, op-for-op as ducc0 evaluates them
# --------------------------------------------------------------------------
_NORMS = ("ortho", "backward", "forward")


def _norm_name(norm):
    """scipy's spelling: None means 'backward' (the forward transform is unscaled)."""
    norm = "backward" if norm is None else norm
    if norm not in _NORMS:
        raise ValueError(f'Invalid norm value {norm!r}; should be "backward", "ortho" or "forward".')
    return norm


def dct2_8(x: np.ndarray, norm="ortho") -> np.ndarray:
    """DCT-II along the last axis (length 8), bit-identical to
    ``scipy.fft.dct(x, axis=-1, norm=norm)`` (reference dct.py:24,26 forward ``self.norm``; 'ortho' everywhere in ivclab).

    float32 in -> float32 arithmetic; anything else -> float64.
    Every ``*2``/``*0.25``/``*0.5`` is an exact power-of-two scaling; all other
    operations are individually rounded (no FMA), in exactly this order.
    ducc0's factor for length 8: ortho 1/sqrt(16), backward 1, forward 1/16 -- all powers of two.
    """
    norm = _norm_name(norm)
    T = np.float32 if x.dtype == np.float32 else np.float64
    tw, wa0, wa1, sq2 = _consts(T)
    c = [x[..., i].astype(T) for i in range(8)]
    two, half, fct = T(2), T(0.5), T({"ortho": 0.25, "backward": 1.0, "forward": 0.0625}[norm])
    # T_dcst23::exec, type 2 pre-processing
    c[0] = c[0] * two
    c[7] = c[7] * two
    for k in (1, 3, 5):
        a, b = c[k + 1], c[k]
        c[k + 1] = a - b
        c[k] = b + a
    # real FFT backward, radix-2 pass (ido=4, l1=1)
    ch = [None] * 8
    ch[0] = c[0] + c[7]
    ch[4] = c[0] - c[7]
    ch[3] = two * c[3]
    ch[7] = -two * c[4]
    ch[1] = c[1] + c[5]
    tr2 = c[1] - c[5]
    ti2 = c[2] + c[6]
    ch[2] = c[2] - c[6]
    ch[6] = wa0 * ti2 + wa1 * tr2
    ch[5] = wa0 * tr2 - wa1 * ti2
    # radix-4 pass (ido=1, l1=2)
    o = [None] * 8
    for k in range(2):
        tr2 = ch[4 * k] + ch[4 * k + 3]
        tr1 = ch[4 * k] - ch[4 * k + 3]
        tr3 = two * ch[4 * k + 1]
        tr4 = two * ch[4 * k + 2]
        o[k] = tr2 + tr3
        o[k + 4] = tr2 - tr3
        o[k + 6] = tr1 + tr4
        o[k + 2] = tr1 - tr4
    o = [v * fct for v in o]                       # ortho: fct = 1/sqrt(2*8) = 0.25
    r = [None] * 8
    r[0] = o[0] * (sq2 * half) if norm == "ortho" else o[0]
    for k, kc in ((1, 7), (2, 6), (3, 5)):
        t1 = tw[k - 1] * o[kc] + tw[kc - 1] * o[k]
        t2 = tw[k - 1] * o[k] - tw[kc - 1] * o[kc]
        r[k] = half * (t1 + t2)
        r[kc] = half * (t1 - t2)
    r[4] = o[4] * tw[3]
    return np.stack(r, axis=-1)


def dct3_8(x: np.ndarray, norm="ortho") -> np.ndarray:
    """DCT-III (= inverse of DCT-II) along the last axis, length 8,
    bit-identical to ``scipy.fft.idct(x, axis=-1, norm=norm)`` (reference
    dct.py:42,44).  The inverse carries the factor the forward direction left out: backward 1/16, forward 1."""
    norm = _norm_name(norm)
    T = np.float32 if x.dtype == np.float32 else np.float64
    tw, wa0, wa1, sq2 = _consts(T)
    c = [x[..., i].astype(T) for i in range(8)]
    two, fct = T(2), T({"ortho": 0.25, "backward": 0.0625, "forward": 1.0}[norm])
    # T_dcst23::exec, type 3 pre-processing
    if norm == "ortho":
        c[0] = c[0] * sq2
    for k, kc in ((1, 7), (2, 6), (3, 5)):
        t1 = c[k] + c[kc]
        t2 = c[k] - c[kc]
        c[k] = tw[k - 1] * t2 + tw[kc - 1] * t1
        c[kc] = tw[k - 1] * t1 - tw[kc - 1] * t2
    c[4] = c[4] * (two * tw[3])
    # real FFT forward, radix-4 pass (ido=1, l1=2)
    ch = [None] * 8
    for k in range(2):
        tr1 = c[k + 6] + c[k + 2]
        ch[2 + 4 * k] = c[k + 6] - c[k + 2]
        tr2 = c[k] + c[k + 4]
        ch[1 + 4 * k] = c[k] - c[k + 4]
        ch[0 + 4 * k] = tr2 + tr1
        ch[3 + 4 * k] = tr2 - tr1
    # radix-2 pass (ido=4, l1=1)
    o = [None] * 8
    o[0] = ch[0] + ch[4]
    o[7] = ch[0] - ch[4]
    o[4] = -ch[7]
    o[3] = ch[3]
    tr2 = wa0 * ch[5] + wa1 * ch[6]
    ti2 = wa0 * ch[6] - wa1 * ch[5]
    o[1] = ch[1] + tr2
    o[5] = ch[1] - tr2
    o[2] = ti2 + ch[2]
    o[6] = ti2 - ch[2]
    o = [v * fct for v in o]
    for k in (1, 3, 5):
        a, b = o[k], o[k + 1]
        o[k] = a - b
        o[k + 1] = b + a
    return np.stack(o, axis=-1)


def _apply_2d(fn, x):
    t = fn(x)                                      # axis -1 first (dct.py:24 / :42)
    return np.ascontiguousarray(np.swapaxes(fn(np.swapaxes(t, -1, -2)), -1, -2))


def dct8x8_forward(patches: np.ndarray, norm="ortho") -> np.ndarray:
    """``DiscreteCosineTransform.transform`` (dct.py:12-28): DCT-II over axis -1
    then axis -2 of ``[..., 8, 8]``; f32->f32, everything else ->f64."""
    return _apply_2d(lambda v: dct2_8(v, norm), np.asarray(patches))


def dct8x8_inverse(patches: np.ndarray, norm="ortho") -> np.ndarray:
    """``DiscreteCosineTransform.inverse_transform`` (dct.py:30-46)."""
    return _apply_2d(lambda v: dct3_8(v, norm), np.asarray(patches))


# --------------------------------------------------------------------------
# 3. PatchQuant
# --------------------------------------------------------------------------
LUMINANCE = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 55, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99,
], dtype=np.float32).reshape(8, 8)                 # patchquant.py:16-25 (NB row 4: 18,55,37,...)
CHROMINANCE = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 13, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
] + [99] * 32, dtype=np.float32).reshape(8, 8)     # patchquant.py:28-37 (NB row 2: 24,13,56)


def quant_table(scale=1.0, luminance=None, chrominance=None) -> np.ndarray:
    """``PatchQuant.get_quantization_table`` (patchquant.py:39-42): stack
    [lum, chrom, chrom] then multiply by ``scale`` under numpy promotion rules
    (python float -> stays float32; np.float64 scalar -> float64)."""
    lum = LUMINANCE if luminance is None else luminance
    chrom = CHROMINANCE if chrominance is None else chrominance
    return np.stack([lum, chrom, chrom], axis=0) * scale


def quantize(x: np.ndarray, table: np.ndarray) -> np.ndarray:
    """``PatchQuant.quantize`` (patchquant.py:56-60): true division, round half
    to even, cast to int32; numpy broadcasting of ``x[..., C, 8, 8]`` against
    ``table[None, None]`` (C=1 -> 3 channels)."""
    return np.round(np.asarray(x) / table[None, None, :, :, :]).astype(np.int32)


def dequantize(q: np.ndarray, table: np.ndarray) -> np.ndarray:
    """``PatchQuant.dequantize`` (patchquant.py:74-78): product in the promoted
    float type, truncation toward zero on the int32 cast."""
    return (np.asarray(q) * table[None, None, :, :, :]).astype(np.int32)


# --------------------------------------------------------------------------
# 4. ZigZag / Patcher
# --------------------------------------------------------------------------
def _build_zigzag():
    # Independent construction (anti-diagonal walk) of the JPEG scan; equals
    # the literal table at shape.py:10-19 (asserted by gen_golden.py / tests).
    order = np.zeros(64, dtype=np.int64)
    pos = 0
    for s in range(15):
        rng_ = range(max(0, s - 7), min(7, s) + 1)
        cells = [(i, s - i) for i in rng_]         # (row, col) with row+col = s
        if s % 2 == 0:
            cells.reverse()                        # even diagonals run bottom-left -> top-right
        for (i, j) in cells:
            order[i * 8 + j] = pos
            pos += 1
    return order


ZIGZAG_ORDER = _build_zigzag()          # ZIGZAG_ORDER[raster k] = scan position (shape.py:10-19)
ZIGZAG_SCAN = np.argsort(ZIGZAG_ORDER)  # ZIGZAG_SCAN[scan pos] = raster index


def zigzag_flatten(p: np.ndarray) -> np.ndarray:
    """``ZigZag.flatten`` (shape.py:21-28): [h,w,c,8,8] -> [h,w,c,64] with
    ``out[..., order[k]] = in[..., k]`` (scatter); dtype preserved."""
    p = np.asarray(p)
    if p.ndim != 5:
        raise ValueError("zigzag_flatten expects a 5-D [h,w,c,8,8] array")
    flat = p.reshape(p.shape[:3] + (64,))
    out = np.zeros_like(flat)
    out[..., ZIGZAG_ORDER] = flat
    return out


def zigzag_unflatten(z: np.ndarray) -> np.ndarray:
    """``ZigZag.unflatten`` (shape.py:30-36): gather ``out[..., k] = in[..., order[k]]``."""
    z = np.asarray(z)
    if z.ndim != 4:
        raise ValueError("zigzag_unflatten expects a 4-D [h,w,c,64] array")
    return z[..., ZIGZAG_ORDER].reshape(z.shape[:3] + (8, 8))


def patch(img: np.ndarray) -> np.ndarray:
    """``Patcher.patch`` (shape.py:45-54): '(h p0)(w p1) c -> h w c p0 p1' view."""
    H, W, C = img.shape
    return img.reshape(H // 8, 8, W // 8, 8, C).transpose(0, 2, 4, 1, 3)


def unpatch(p: np.ndarray) -> np.ndarray:
    """``Patcher.unpatch`` (shape.py:56-65)."""
    hp, wp, C = p.shape[:3]
    return np.ascontiguousarray(p.transpose(0, 3, 1, 4, 2)).reshape(hp * 8, wp * 8, C)


# --------------------------------------------------------------------------
# 5. MotionCompensator
# --------------------------------------------------------------------------
def np_sum64(sq: np.ndarray) -> np.ndarray:
    """``np.sum`` of contiguous 64-element blocks ``sq[..., 8, 8]`` in numpy's
    exact order (motion.py:46): 8 strided accumulators r[j] = sum_i a[8i+j]
    taken sequentially in i, then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))."""
    r = [sq[..., 0, j] for j in range(8)]
    for i in range(1, 8):
        r = [r[j] + sq[..., i, j] for j in range(8)]
    return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))


def me_full_search_loops(ref: np.ndarray, cur: np.ndarray, sr: int) -> np.ndarray:
    """Literal restatement of ``compute_motion_vector`` (motion.py:8-58) with
    the same four nested loops -- used for small cases and as the CPU
    "as shipped" timing arm."""
    H, W = ref.shape
    mv = np.zeros((H // 8, W // 8, 1), dtype=int)
    span = 2 * sr + 1
    for by in range(H // 8):
        for bx in range(W // 8):
            y, x = 8 * by, 8 * bx
            blk = cur[y:y + 8, x:x + 8]
            best, bdy, bdx = float("inf"), 0, 0
            for dy in range(-sr, sr + 1):
                for dx in range(-sr, sr + 1):
                    yy, xx = y + dy, x + dx
                    if yy < 0 or yy + 8 > H or xx < 0 or xx + 8 > W:
                        continue
                    ssd = np.sum((blk - ref[yy:yy + 8, xx:xx + 8]) ** 2)
                    if ssd < best:
                        best, bdy, bdx = ssd, dy, dx
            mv[by, bx, 0] = (bdy + sr) * span + (bdx + sr)
    return mv


def me_full_search(ref: np.ndarray, cur: np.ndarray, sr: int) -> np.ndarray:
    """Vectorised ``compute_motion_vector`` (motion.py:8-58): candidates are
    visited in (dy asc, dx asc) order, a candidate is skipped when its window
    leaves the frame (:41-43), strict ``<`` keeps the first minimum (:48), and
    each SSD is summed in numpy's own order (``np_sum64``).  Arithmetic dtype
    follows numpy promotion of ``cur - ref`` (float32 stays float32)."""
    ref = np.asarray(ref)
    cur = np.asarray(cur)
    H, W = ref.shape
    hp, wp = H // 8, W // 8
    span = 2 * sr + 1
    dt = np.result_type(ref.dtype, cur.dtype)
    curb = cur[:hp * 8, :wp * 8].reshape(hp, 8, wp, 8).transpose(0, 2, 1, 3)
    best = np.full((hp, wp), np.inf, dtype=np.float64)
    idx = np.full((hp, wp), sr * span + sr, dtype=np.int64)   # default (0,0) (:31-33)
    by = np.arange(hp)[:, None] * 8
    bx = np.arange(wp)[None, :] * 8
    for dy in range(-sr, sr + 1):
        oky = (by + dy >= 0) & (by + dy + 8 <= H)
        if not oky.any():
            continue
        y0, y1 = np.nonzero(oky[:, 0])[0][[0, -1]]
        for dx in range(-sr, sr + 1):
            okx = (bx + dx >= 0) & (bx + dx + 8 <= W)
            if not okx.any():
                continue
            x0, x1 = np.nonzero(okx[0, :])[0][[0, -1]]
            nby, nbx = y1 - y0 + 1, x1 - x0 + 1
            r = ref[8 * y0 + dy: 8 * (y1 + 1) + dy, 8 * x0 + dx: 8 * (x1 + 1) + dx]
            rb = r.reshape(nby, 8, nbx, 8).transpose(0, 2, 1, 3)
            d = (curb[y0:y1 + 1, x0:x1 + 1] - rb).astype(dt, copy=False)
            ssd = np_sum64(d * d)
            sub_best = best[y0:y1 + 1, x0:x1 + 1]
            sub_idx = idx[y0:y1 + 1, x0:x1 + 1]
            take = ssd < sub_best
            sub_best[take] = ssd[take]
            sub_idx[take] = (dy + sr) * span + (dx + sr)
    return idx[:, :, None].astype(int)


def mc_reconstruct(ref: np.ndarray, mv: np.ndarray, sr: int) -> np.ndarray:
    """``reconstruct_with_motion_vector`` (motion.py:60-97): per 8x8 block copy
    ``ref[y+dy:y+dy+8, x+dx:x+dx+8, :]``; a block whose source window leaves
    the frame stays zero (:90-92); dtype of ``ref`` preserved (:74)."""
    ref = np.asarray(ref)
    H, W, C = ref.shape
    span = 2 * sr + 1
    out = np.zeros_like(ref)
    idx = np.asarray(mv)[:, :, 0]
    dy = idx // span - sr
    dx = idx % span - sr
    for by in range(H // 8):
        for bx in range(W // 8):
            yy, xx = 8 * by + int(dy[by, bx]), 8 * bx + int(dx[by, bx])
            if yy < 0 or yy + 8 > H or xx < 0 or xx + 8 > W:
                continue
            out[8 * by:8 * by + 8, 8 * bx:8 * bx + 8, :] = ref[yy:yy + 8, xx:xx + 8, :]
    return out


# --------------------------------------------------------------------------
# 6. Fused restatements (what the fused kernels must equal)
# --------------------------------------------------------------------------
def intra_forward(img_hwc: np.ndarray, table: np.ndarray) -> np.ndarray:
    """patch -> transform -> quantize -> flatten, as chained by
    ``IntraCodec.image2symbols`` (intracodec.py:66-75).  [H,W,C] -> [Hp,Wp,3,64] int32."""
    return zigzag_flatten(quantize(dct8x8_forward(patch(img_hwc)), table))


def intra_inverse(zz: np.ndarray, table: np.ndarray) -> np.ndarray:
    """unflatten -> dequantize -> inverse_transform -> un-patch, as chained by
    ``IntraCodec.symbols2image`` (intracodec.py:115-124).  [Hp,Wp,C,64] -> [H,W,3] f64."""
    return unpatch(dct8x8_inverse(dequantize(zigzag_unflatten(zz), table)))


def pframe_forward(cur: np.ndarray, ref: np.ndarray, mv: np.ndarray, sr: int, table: np.ndarray):
    """P-frame encoder half (videocodec.py:68-71; exercises/ch4/E4-1.py:268-275):
    prediction = MC(ref, mv); residual = cur - prediction; symbols-side indices =
    flatten(quantize(dct(patch(residual[...,None])))).  Returns (prediction, zz)."""
    pred = mc_reconstruct(ref[..., None], mv, sr)[..., 0]
    residual = cur - pred
    return pred, intra_forward(residual[..., None], table)


def pframe_inverse(zz_luma: np.ndarray, pred: np.ndarray, table: np.ndarray) -> np.ndarray:
    """P-frame decoder half (videocodec.py:74; E4-1.py:284-306) for indices that
    are already laid out ``[Hp,Wp,1,64]``: recon = pred + idct(dequant(.))[...,0]
    (channel 0 of the 3-channel broadcast = luminance table)."""
    rec = intra_inverse(zz_luma, table)[..., 0]
    return pred + rec


# --------------------------------------------------------------------------
# 7. Neighbours of the path ("next" rows N1-N3 of SURVEY.md section 8f)
# --------------------------------------------------------------------------
_RGB2YCBCR = np.array([[0.299, 0.587, 0.114],
                       [-0.168736, -0.331264, 0.5],
                       [0.5, -0.418688, -0.081312]])


def rgb2ycbcr(image: np.ndarray) -> np.ndarray:
    """``rgb2ycbcr`` (ivclab/signal/color.py:15-37)."""
    return image @ _RGB2YCBCR.T + np.array([0, 128, 128])


def ycbcr2rgb(image: np.ndarray) -> np.ndarray:
    """``ycbcr2rgb`` (color.py:39-63)."""
    Y = image[:, :, 0]
    Cb = image[:, :, 1] - 128.0
    Cr = image[:, :, 2] - 128.0
    R = Y + 1.402 * Cr
    G = Y - 0.344136 * Cb - 0.714136 * Cr
    B = Y + 1.772 * Cb
    return np.clip(np.stack([R, G, B], axis=-1), 0, 255)


def calc_mse(orig: np.ndarray, rec: np.ndarray) -> float:
    """``calc_mse`` (ivclab/utils/metrics.py:3-23)."""
    if orig.ndim == 2 and rec.ndim == 3:
        orig = np.stack([orig] * 3, axis=-1)
    elif orig.ndim == 3 and rec.ndim == 2:
        rec = np.stack([rec] * 3, axis=-1)
    return float(np.mean((orig.astype(np.float64) - rec.astype(np.float64)) ** 2))


def calc_psnr(orig: np.ndarray, rec: np.ndarray, maxval=255) -> float:
    """``calc_psnr`` (metrics.py:25-40)."""
    return float(20 * np.log10(maxval / np.sqrt(calc_mse(orig, rec))))


def zerorun_encode(zz: np.ndarray, eob: int = 4000) -> np.ndarray:
    """``ZeroRunCoder.encode`` (ivclab/entropy/zerorun.py:10-43): per 64-block,
    non-zero -> itself, a run of zeros before the last non-zero -> (0, run),
    trailing zeros -> EOB; all-zero block -> EOB."""
    flat = np.asarray(zz).reshape(-1, 64)
    out = []
    for blk in flat:
        nz = np.nonzero(blk)[0]
        if nz.size == 0:
            out.append(eob)
            continue
        last = nz[-1]
        i = 0
        while i <= last:
            v = int(blk[i])
            if v == 0:
                j = i
                while j <= last and blk[j] == 0:
                    j += 1
                out.extend((0, j - i))
                i = j
            else:
                out.append(v)
                i += 1
        out.append(eob)
    return np.array(out, dtype=np.int32)


def zerorun_encode_fast(zz: np.ndarray, eob: int = 4000) -> np.ndarray:
    """The same stream as :func:`zerorun_encode` (zerorun.py:10-43), vectorised over blocks so that full 1080p
    frames are affordable: position p of a block emits its value if non-zero, the pair (0, run length) if it starts
    a run of zeros that a non-zero value follows, nothing otherwise; every block ends with EOB.  Checked against
    the loop form in tests/test_oracle_cpu.py."""
    b = np.asarray(zz).reshape(-1, 64)
    nb = b.shape[0]
    nz = b != 0
    idx = np.arange(64)
    last = np.where(nz.any(axis=1), 63 - np.argmax(nz[:, ::-1], axis=1), -1)
    below = idx[None, :] <= last[:, None]
    prev_nz = np.concatenate([np.ones((nb, 1), dtype=bool), nz[:, :-1]], axis=1)
    start = ~nz & prev_nz & below                                   # a zero run that is followed by a value
    nxt = np.where(nz, idx[None, :], 64)
    nxt = np.minimum.accumulate(nxt[:, ::-1], axis=1)[:, ::-1]     # next non-zero position at or after p
    cnt = nz.astype(np.int64) + 2 * start
    per_block = cnt.sum(axis=1) + 1
    offs = np.concatenate([[0], np.cumsum(per_block)])
    pos = offs[:-1, None] + (np.cumsum(cnt, axis=1) - cnt)
    out = np.empty(int(offs[-1]), dtype=np.int32)
    out[pos[nz]] = b[nz]
    out[pos[start]] = 0
    out[pos[start] + 1] = (nxt - idx[None, :])[start]
    out[offs[1:] - 1] = eob
    return out


def symbol_histogram(symbols: np.ndarray, lo: int, n_bins: int) -> np.ndarray:
    """np.histogram(symbols, bins=np.arange(lo, lo + n_bins + 1))[0] for integer symbols inside the range
    (entropy.py:24: unit bins), as int64 counts."""
    s = np.asarray(symbols).astype(np.int64) - lo
    s = s[(s >= 0) & (s < n_bins)]
    return np.bincount(s, minlength=n_bins).astype(np.int64)


def zerorun_decode(symbols, shape, eob: int = 4000) -> np.ndarray:
    """``ZeroRunCoder.decode`` (zerorun.py:44-87) incl. the "stop after h*w*c
    blocks" rule (:60-62)."""
    h, w, c = shape
    want = h * w * c
    blocks = np.zeros((want, 64), dtype=np.int32)
    i = 0
    n = 0
    symbols = np.asarray(symbols)
    while i < len(symbols) and n < want:
        pos = 0
        while True:
            if i >= len(symbols):
                raise ValueError("Unexpected end of encoded symbols")
            s = int(symbols[i])
            i += 1
            if s == eob:
                break
            if s == 0:
                pos += int(symbols[i])
                i += 1
            else:
                if pos >= 64:
                    raise ValueError("Block size exceeded")
                blocks[n, pos] = s
                pos += 1
            if pos > 64:
                raise ValueError("Block size exceeded")
        n += 1
    if n != want:
        raise ValueError(f"Expected {want} blocks, got {n}")
    return blocks.reshape(h, w, c, 64)


def stats_marg(image: np.ndarray, pixel_range: np.ndarray) -> np.ndarray:
    """``stats_marg`` (entropy.py:6-29): histogram over the bin EDGES ``pixel_range`` (last bin closed),
    normalised by the number of samples (not by the number of counted samples)."""
    flat = np.asarray(image).astype(np.float64).flatten()
    counts, _ = np.histogram(flat, bins=pixel_range)
    return counts / flat.size


def symbol_bounds(symbols: np.ndarray, margin: int = 20):
    """Bin range of ``IntraCodec.train_huffman_from_image`` (intracodec.py:161-164)."""
    s = np.asarray(symbols, dtype=np.int32)
    return int(s.min()) - margin, int(s.max()) + margin + 1


# --------------------------------------------------------------------------
# 8. Seeded synthetic inputs
# --------------------------------------------------------------------------
def _box5(a: np.ndarray) -> np.ndarray:
    """5x5 box mean with edge replication, separable, plain numpy (no scipy so
    the generator is identical on every box)."""
    p = np.pad(a, ((2, 2), (2, 2)) + ((0, 0),) * (a.ndim - 2), mode="edge")
    s = sum(p[i:i + a.shape[0]] for i in range(5))
    s = sum(s[:, j:j + a.shape[1]] for j in range(5))
    return s / 25.0


def smooth_noise_rgb(seed: int, H: int, W: int) -> np.ndarray:
    """uint8 RGB 'smooth-noise' image: clip(box5(U{0..255}) + N(0,4^2), 0, 255)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(H, W, 3)).astype(np.float64)
    img = _box5(base) * 1.0 + rng.normal(0.0, 4.0, size=(H, W, 3))
    # stretch contrast so that AC coefficients survive coarse quantisation
    img = (img - 127.5) * 3.0 + 127.5
    return np.clip(img, 0, 255).astype(np.uint8)


def smooth_noise_luma(seed: int, H: int, W: int) -> np.ndarray:
    """integer-valued float64 luma plane of the same texture."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(H, W)).astype(np.float64)
    img = (_box5(base) - 127.5) * 3.0 + 127.5 + rng.normal(0.0, 4.0, size=(H, W))
    return np.clip(np.round(img), 0, 255)


def moving_sequence(seed: int, n_frames: int, H: int, W: int, max_shift: int = 3,
                    obj: int = 32) -> np.ndarray:
    """[n,H,W] float64 integer-valued luma frames: crops of one canvas under a
    per-frame global translation plus an object moving (+2,+1)/frame plus
    N(0,1) sensor noise."""
    rng = np.random.default_rng(seed)
    m = max_shift * 2 + 8
    canvas = smooth_noise_luma(seed + 1, H + 2 * m, W + 2 * m)
    patch_tex = smooth_noise_luma(seed + 2, obj, obj)
    frames = np.empty((n_frames, H, W), dtype=np.float64)
    for t in range(n_frames):
        dx, dy = rng.integers(-max_shift, max_shift + 1, size=2)
        f = canvas[m + dy:m + dy + H, m + dx:m + dx + W].copy()
        oy = (H // 4 + t) % max(1, H - obj)
        ox = (W // 4 + 2 * t) % max(1, W - obj)
        f[oy:oy + obj, ox:ox + obj] = patch_tex[:min(obj, H - oy), :min(obj, W - ox)]
        f = f + rng.normal(0.0, 1.0, size=f.shape)
        frames[t] = np.clip(np.round(f), 0, 255)
    return frames
