"""CPU tests (no GPU): the oracle against the golden vectors recorded from the real reference,
the C restatement and the scipy-call port against the oracle, the host-side logic of the package,
and that the C-ABI library loads and exports every symbol include/ivclab_b200.h declares."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLD, ME_CASES, QSCALES, ROOT, case_sr
from oracle import c_oracle, ivc_oracle as O, ref_port as R


# ---------------------------------------------------------------- oracle vs golden (reference outputs)
def test_pinning_report_is_complete():
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))
    assert len(pin["checks"]) >= 90 and all(c["ok"] for c in pin["checks"].values())


def test_constants():
    assert np.array_equal(O.ZIGZAG_SCAN[:10], [0, 1, 8, 16, 9, 2, 3, 10, 17, 24])      # JPEG scan
    assert sorted(O.ZIGZAG_ORDER.tolist()) == list(range(64))
    tw, wa = O.derive_ducc_constants()          # ducc0's recipe re-run on this box's libm
    assert np.array_equal(tw, O.DUCC_TW) and np.array_equal(wa, O.DUCC_WA)
    assert O.LUMINANCE[4, 1] == 55 and O.CHROMINANCE[2, 1] == 13                       # the reference's own table quirks


def test_oracle_dct_matches_scipy_here():
    from scipy.fft import dct, idct
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        x = rng.uniform(-300, 300, size=(4000, 8)).astype(dt)
        assert np.array_equal(O.dct2_8(x), dct(x, axis=-1, norm="ortho"))
        assert np.array_equal(O.dct3_8(x), idct(x, axis=-1, norm="ortho"))
    xi = rng.integers(-500, 500, size=(4000, 8)).astype(np.int32)
    assert np.array_equal(O.dct3_8(xi), idct(xi, axis=-1, norm="ortho"))


def test_oracle_dct_other_norms_match_scipy_here():
    """The reference forwards `norm` to scipy (dct.py:24,26,42,44): None / 'backward' / 'forward' as scipy computes them,
    one axis and the 2-D composition the class applies."""
    from scipy.fft import dct, idct
    rng = np.random.default_rng(1)
    for dt in (np.float64, np.float32):
        for scale in (1.0, 300.0, 1e-4):
            x = (rng.standard_normal((600, 8, 8)) * scale).astype(dt)
            for norm in (None, "backward", "forward", "ortho"):
                assert np.array_equal(O.dct2_8(x, norm), dct(x, axis=-1, norm=norm))
                assert np.array_equal(O.dct3_8(x, norm), idct(x, axis=-1, norm=norm))
                assert np.array_equal(O.dct8x8_forward(x, norm), dct(dct(x, axis=-1, norm=norm), axis=-2, norm=norm))
                assert np.array_equal(O.dct8x8_inverse(x, norm), idct(idct(x, axis=-1, norm=norm), axis=-2, norm=norm))
    with pytest.raises(ValueError):
        O.dct2_8(np.zeros((1, 8)), "bogus")


@pytest.mark.parametrize("qi", range(4))
def test_oracle_intra_golden(g1, g2, qi):
    tab = O.quant_table(QSCALES[qi])
    assert np.array_equal(tab, g1[f"table{qi}"]) and tab.dtype == g1[f"table{qi}"].dtype
    assert np.array_equal(O.dct8x8_forward(O.patch(g1["img"])), g1["coef"])
    assert np.array_equal(O.intra_forward(g1["img"], tab), g1[f"zz{qi}"])
    assert np.array_equal(O.dequantize(O.zigzag_unflatten(g1[f"zz{qi}"]), tab), g1[f"dq{qi}"])
    assert np.array_equal(O.intra_inverse(g1[f"zz{qi}"], tab), g1[f"rec{qi}"])
    assert np.array_equal(O.intra_forward(g2["luma"][..., None], tab), g2[f"zz{qi}"])
    assert np.array_equal(O.intra_inverse(g2[f"zz{qi}"][:, :, :1], tab), g2[f"rec{qi}"])


def test_oracle_dtypes_golden(g3):
    assert np.array_equal(O.dct8x8_forward(g3["x32"]), g3["c32"])
    assert np.array_equal(O.dct8x8_inverse(g3["x32"]), g3["i32"])
    assert np.array_equal(O.dct8x8_forward(g3["xu8"]), g3["cu8"])
    assert np.array_equal(O.dct8x8_forward(g3["one"]), g3["c_one"])
    assert np.array_equal(O.quantize(g3["xu8"], g3["tab1"]), g3["q_u8"])
    assert np.array_equal(O.quantize(g3["c32"], g3["tab007"]), g3["q_f32"])


@pytest.mark.parametrize("name", ME_CASES)
def test_oracle_motion_golden(g4, name):
    sr = case_sr(name)
    r, c = g4[f"{name}__ref"], g4[f"{name}__cur"]
    assert np.array_equal(O.me_full_search(r, c, sr), g4[f"{name}__mv"])
    if r.size <= 48 * 64 and sr <= 4:
        assert np.array_equal(O.me_full_search_loops(r, c, sr), g4[f"{name}__mv"])
    assert np.array_equal(O.mc_reconstruct(r[..., None], g4[f"{name}__mv"], sr), g4[f"{name}__pred"])


def test_oracle_pframe_and_neighbours_golden(g1, g4, g5, g6):
    assert np.array_equal(O.mc_reconstruct(g4["mc_ref3"], g4["mc_mv_rand"], 4), g4["mc_pred3"])
    pred, zz = O.pframe_forward(g5["cur"], g5["ref"], g5["mv"], 4, g5["table"])
    assert np.array_equal(pred, g5["pred"]) and np.array_equal(zz, g5["zz"])
    assert np.array_equal(O.pframe_inverse(g5["zz"][:, :, :1], g5["pred"], g5["table"]), g5["recon"])
    assert np.array_equal(O.zerorun_encode(g1["zz1"]), g6["sym"])
    assert np.array_equal(O.zerorun_decode(g6["sym"], (6, 8, 1)), g6["dec_trunc"])
    assert np.array_equal(O.ycbcr2rgb(g1["rec1"]), g6["rec_rgb"])
    assert O.calc_psnr(g1["rgb"], g6["rec_rgb"]) == float(g6["psnr"])


def test_oracle_qcif_golden(g7):
    seq = O.moving_sequence(2, 6, 144, 176)
    for t in (1, 3):
        assert np.array_equal(O.me_full_search(seq[t - 1], seq[t], 4)[..., 0], g7["mvs"][t - 1][..., 0])


def test_oracle_entropy_golden(g9):
    """Zero-run decoder (incl. the stop-after-h*w*c rule) and stats_marg against outputs recorded from the
    real reference (oracle/gen_golden_entropy.py)."""
    assert np.array_equal(O.zerorun_decode(g9["sym"], (6, 8, 3)), g9["dec_full"])
    assert np.array_equal(O.zerorun_decode(g9["sym"], (6, 8, 1)), g9["dec_trunc"])
    assert np.array_equal(O.zerorun_encode(g9["blocks"]), g9["sym2"])
    assert np.array_equal(O.zerorun_decode(g9["sym2"], (5, 7, 3)), g9["blocks"])
    lo, hi = int(g9["lo"]), int(g9["hi"])
    assert O.symbol_bounds(g9["sym"]) == (lo, hi)
    assert np.array_equal(O.stats_marg(g9["sym"], np.arange(lo, hi)), g9["pmf"])
    assert np.array_equal(O.stats_marg(g9["sym"], np.arange(-3, 9)), g9["pmf_cut"])
    assert np.array_equal(O.stats_marg(g9["img8"], np.arange(256)), g9["pmf8"])
    with pytest.raises(ValueError, match="Unexpected end"):
        O.zerorun_decode(g9["sym2"][:-1], (5, 7, 3))
    with pytest.raises(ValueError, match="Expected 140 blocks, got 105"):
        O.zerorun_decode(g9["sym2"], (5, 7, 4))


def test_streamed_coder_chunk_schedule():
    """Host logic of the pipeline's chunking (no device needed): chunks tile the run exactly, never exceed the slot
    size, and the ramped schedule puts the short chunks at both ends."""
    from ivclab_b200.streaming import StreamedCoder
    sc = object.__new__(StreamedCoder)
    for C in (1, 2, 4, 8):
        for ramp in ((), (2, 2), (1, 1, 2)):
            sc.chunk, sc.ramp = C, ramp
            for F in (1, 2, 5, 8, 31, 32, 33, 100):
                b = sc._schedule(F)
                assert b[0][0] == 0 and b[-1][1] == F and all(x[1] == y[0] for x, y in zip(b, b[1:]))
                assert all(0 < hi - lo <= C for lo, hi in b)
    sc.chunk, sc.ramp = 4, (1, 1, 2)
    assert [hi - lo for lo, hi in sc._schedule(32)] == [1, 1, 2, 4, 4, 4, 4, 4, 4, 2, 1, 1]


def test_flat_frame_tie_break():
    """all-tie search: interior -> first candidate (index 0), borders -> first in-bounds."""
    z = np.zeros((40, 48))
    mv = O.me_full_search(z, z, 4)[..., 0]
    assert mv[2, 2] == 0 and mv[0, 2] == 36 and mv[2, 0] == 4 and mv[0, 0] == 40


# ---------------------------------------------------------------- C restatement and scipy port vs oracle
needs_c = pytest.mark.skipif(not c_oracle.available(), reason="oracle/_build/libivc_oracle.so not built")


@needs_c
def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(5)
    for C_, shape in ((3, (40, 56, 3)), (1, (24, 64, 1))):
        img = rng.uniform(-20, 280, size=shape)
        for q in (0.07, 1.0, np.float64(0.4)):
            tab = O.quant_table(q)
            zz = c_oracle.intra_forward(img, tab, threads=2)
            assert np.array_equal(zz, O.intra_forward(img, tab))
            zin = zz if C_ == 3 else zz[:, :, :1]
            assert np.array_equal(c_oracle.intra_inverse(zin, tab, threads=2), O.intra_inverse(zin, tab))
    ref = rng.uniform(0, 255, size=(48, 72))
    cur = np.roll(ref, (1, -2), (0, 1)) + rng.normal(0, 1, ref.shape)
    for dt in (np.float64, np.float32):
        for sr in (2, 4, 9):
            assert np.array_equal(c_oracle.me_full_search(ref.astype(dt), cur.astype(dt), sr, threads=3),
                                  O.me_full_search(ref.astype(dt), cur.astype(dt), sr))
    mv = rng.integers(0, 81, size=(6, 9, 1))
    assert np.array_equal(c_oracle.mc_reconstruct(ref[..., None], mv, 4), O.mc_reconstruct(ref[..., None], mv, 4))


@needs_c
def test_c_oracle_motion_golden(g4):
    for name in ME_CASES:
        r, c = g4[f"{name}__ref"], g4[f"{name}__cur"]
        if r.dtype != c.dtype:
            continue
        assert np.array_equal(c_oracle.me_full_search(r, c, case_sr(name)), g4[f"{name}__mv"])


def test_scipy_port_matches_oracle():
    rng = np.random.default_rng(6)
    img = rng.uniform(0, 255, size=(32, 48, 3))
    tab = O.quant_table(0.2)
    zz, rec = R.intra_loop(img, tab)
    assert np.array_equal(zz, O.intra_forward(img, tab)) and np.array_equal(rec, O.intra_inverse(zz, tab))
    seq = O.moving_sequence(3, 2, 24, 32)
    ref = seq[0] + rng.normal(0, 0.3, seq[0].shape)
    mv, zzp, recon = R.pframe_loop(seq[1], ref, 3, tab)
    assert np.array_equal(mv, O.me_full_search(ref, seq[1], 3))
    pred, z = O.pframe_forward(seq[1], ref, mv, 3, tab)
    assert np.array_equal(z, zzp) and np.array_equal(recon, O.pframe_inverse(z[:, :, :1], pred, tab))


# ---------------------------------------------------------------- C ABI surface and host logic
def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ivclab_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ivc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import ivclab_b200
    from ivclab_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 16
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/ivclab_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib.ivc_abi_version() == _lib.ABI_VERSION == 10
    assert b"sm_100a" in _lib.lib.ivc_build_info()
    assert _lib.lib.ivc_me_workspace_bytes(2, 16, 16) >= 4
    assert ivclab_b200.__version__


def test_argument_validation_without_gpu():
    """status codes that are decided before any CUDA call"""
    from ivclab_b200 import _lib
    L = _lib.lib
    s5 = _lib.strides5((64, 64, 64, 8, 1))
    one = ctypes.c_void_p(16)
    assert L.ivc_dct8x8(0, None, 0, one, _lib.I64, 1, 1, 1, s5, one, _lib.F64) == _lib.ERR_DTYPE
    assert L.ivc_dct8x8(0, None, 0, one, _lib.F32, 1, 1, 1, s5, one, _lib.F64) == _lib.ERR_DTYPE
    assert L.ivc_dct8x8(0, None, 0, None, _lib.F64, 0, 1, 1, s5, None, _lib.F64) == _lib.OK      # empty input
    assert L.ivc_quantize(0, None, one, _lib.F64, 1, 1, 2, s5, one, _lib.F32, _lib.F64, one) == _lib.ERR_SHAPE
    assert L.ivc_intra_forward(0, None, one, _lib.F64, 1, 12, 16, 3, 0, one, _lib.F32, one) == _lib.ERR_SHAPE
    assert L.ivc_intra_forward(0, None, one, _lib.F32, 1, 16, 16, 3, 0, one, _lib.F32, one) == _lib.ERR_DTYPE
    assert L.ivc_intra_forward(0, None, ctypes.c_void_p(8), _lib.F64, 1, 16, 16, 3, 768, one, _lib.F32, one) == _lib.ERR_ARG
    assert L.ivc_me_full_search(0, None, one, one, _lib.F64, 1, 20, 24, 480, 480, 4, 0, one, None, 0) == _lib.ERR_SHAPE
    assert L.ivc_me_full_search(0, None, one, one, _lib.I32, 1, 16, 16, 256, 256, 4, 0, one, None, 0) == _lib.ERR_DTYPE
    assert L.ivc_me_full_search(0, None, one, one, _lib.F64, 1, 16, 16, 256, 256, 4, 7, one, None, 0) == _lib.ERR_ARG
    assert L.ivc_zigzag(0, None, 0, one, 3, 1, one) == _lib.ERR_DTYPE
    assert L.ivc_pframe_inverse(0, None, one, 3, None, None, None, _lib.F64, 1, 16, 16, 4, one, _lib.F32, one) == _lib.ERR_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_SHAPE, "x")
    with pytest.raises(_lib.IvcError):
        _lib.check(_lib.ERR_ARG, "x")


def test_host_mirror_of_reference_interface():
    import ivclab_b200 as ivc
    pq = ivc.PatchQuant()
    assert np.array_equal(pq.luminance, O.LUMINANCE) and np.array_equal(pq.chrominance, O.CHROMINANCE)
    assert pq.quantization_scale == 1.0
    for q in (0.07, 1.0, np.float64(0.4), 3):
        t = ivc.PatchQuant(quantization_scale=q).get_quantization_table()
        assert np.array_equal(t, O.quant_table(q)) and t.dtype == O.quant_table(q).dtype and t.shape == (3, 8, 8)
    assert np.array_equal(ivc.ZigZag().zigzag_order, O.ZIGZAG_ORDER)
    assert ivc.DiscreteCosineTransform().norm == "ortho" and ivc.MotionCompensator().search_range == 4
    img = np.arange(16 * 24 * 3, dtype=np.float64).reshape(16, 24, 3)
    P = ivc.Patcher()
    assert P.window_size == (8, 8)
    v = P.patch(img)
    assert v.shape == (2, 3, 3, 8, 8) and np.shares_memory(v, img) and np.array_equal(v, O.patch(img))
    assert v.strides == (8 * 24 * 3 * 8, 8 * 3 * 8, 8, 24 * 3 * 8, 3 * 8)
    assert np.array_equal(P.unpatch(v), img)
    with pytest.raises(ValueError):
        ivc.MotionCompensator(me_mode="fast")


def test_product_fails_loudly_without_gpu():
    import torch
    import ivclab_b200 as ivc
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    for call in (lambda: ivc.DiscreteCosineTransform().transform(np.zeros((1, 1, 1, 8, 8))),
                 lambda: ivc.PatchQuant().quantize(np.zeros((1, 1, 3, 8, 8))),
                 lambda: ivc.ZigZag().flatten(np.zeros((1, 1, 3, 8, 8))),
                 lambda: ivc.MotionCompensator().compute_motion_vector(np.zeros((16, 16)), np.zeros((16, 16))),
                 lambda: ivc.IntraBlockCoder().forward(np.zeros((16, 16, 3)))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ivclab_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert not re.search(r"^\s*(from|import)\s+scipy\b", src, flags=re.M), f"{f} imports scipy (CPU math)"


def test_install_swaps_classes_into_a_fake_ivclab():
    import sys
    import types
    import ivclab_b200 as ivc
    mods = {}
    for name in ("ivclab", "ivclab.signal", "ivclab.quantization", "ivclab.utils", "ivclab.video"):
        mods[name] = types.ModuleType(name)
    try:
        sys.modules.update(mods)
        done = ivc.install()
        assert set(done) == {"ivclab.signal", "ivclab.quantization", "ivclab.utils", "ivclab.video"}
        assert sys.modules["ivclab.signal"].DiscreteCosineTransform is ivc.DiscreteCosineTransform
        assert sys.modules["ivclab.utils"].ZigZag is ivc.ZigZag and sys.modules["ivclab.utils"].Patcher is ivc.Patcher
        assert sys.modules["ivclab.video"].MotionCompensator is ivc.MotionCompensator
    finally:
        for name in mods:
            sys.modules.pop(name, None)

    class FakeCodec:                      # attribute surface of IntraCodec (intracodec.py:25-30)
        def __init__(self):
            self.dct, self.quant, self.zigzag, self.patcher = object(), types.SimpleNamespace(quantization_scale=0.4), 1, 2

    c = ivc.inject(FakeCodec())
    assert isinstance(c.dct, ivc.DiscreteCosineTransform) and c.quant.quantization_scale == 0.4
    assert isinstance(c.zigzag, ivc.ZigZag) and isinstance(c.patcher, ivc.Patcher)


def test_vectorised_zerorun_oracle_equals_loop_form(g1, g2, g9):
    """oracle.zerorun_encode_fast (used for full-size 1080p checks) == the loop restatement of zerorun.py:10-43."""
    rng = np.random.default_rng(17)
    cases = [g1["zz0"], g1["zz1"], g1["zz2"], g2["zz1"]]
    zz = np.where(rng.random((4, 5, 3, 64)) < 0.2, rng.integers(-3000, 3000, (4, 5, 3, 64)), 0).astype(np.int32)
    zz[0, 0, 0] = 0
    zz[0, 0, 1] = 9
    zz[0, 0, 2] = 0
    zz[0, 0, 2, 63] = -1
    zz[0, 1, 0] = 0
    zz[0, 1, 0, 0] = 4
    cases.append(zz)
    for c in cases:
        a, b = O.zerorun_encode(c), O.zerorun_encode_fast(c)
        assert a.dtype == b.dtype and np.array_equal(a, b)
        assert np.array_equal(O.symbol_histogram(a, -4096, 8192), np.histogram(a, bins=np.arange(-4096, 4097))[0])
    assert O.zerorun_encode_fast(np.zeros((0, 64), dtype=np.int32)).size == 0


def test_expand_inter_rebuilds_the_three_channel_stream():
    """StreamedCoder.expand_inter: host-side expansion of the two-channel P-frame transfer format."""
    from ivclab_b200.streaming import StreamedCoder
    rng = np.random.default_rng(3)
    zz2 = np.where(rng.random((5, 7, 2, 64)) < 0.1, rng.integers(-40, 40, (5, 7, 2, 64)), 0).astype(np.int32)
    zz2[0, 0, 0] = 0
    zz2[1, 1, 1] = 7
    zz2[2, 2, 0, 5] = 4000 - 3937                                     # a run length that collides with nothing
    zz3 = np.concatenate([zz2, zz2[:, :, 1:2]], axis=2)
    got = StreamedCoder.expand_inter(O.zerorun_encode(zz2).astype(np.int16))
    assert got.dtype == np.int32 and np.array_equal(got, O.zerorun_encode(zz3))
    assert StreamedCoder.expand_inter(np.zeros(0, dtype=np.int16)).size == 0
    with pytest.raises(ValueError):
        StreamedCoder.expand_inter(O.zerorun_encode(zz2[:1, :1, :1]))
