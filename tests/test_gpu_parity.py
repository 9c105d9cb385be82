"""GPU parity tests (run with ``-m gpu`` on the B200): the CUDA path, called through the drop-in
classes and hence through the C ABI, against (a) the golden vectors produced by the real
reference (tests/golden, oracle/gen_golden.py) and (b) the oracle on seeded inputs.
Bar: bit-exact for indices, motion vectors and -- because the DCT replays scipy's operation
order -- also for every float64/float32 transform output (tolerance 0)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, ME_CASES, QSCALES, case_sr

pytestmark = pytest.mark.gpu

import ivclab_b200 as ivc  # noqa: E402
from oracle import ivc_oracle as O  # noqa: E402  (the checker)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- per-op, golden vectors
def test_dct_forward_golden_strided_view(g1):
    img = g1["img"]
    patches = ivc.Patcher().patch(img)                       # strided view, like intracodec.py:66
    assert not patches.flags.c_contiguous
    out = ivc.DiscreteCosineTransform().transform(patches)
    assert out.dtype == np.float64 and out.flags.c_contiguous
    assert np.array_equal(out, g1["coef"])                   # bit-exact vs scipy


@pytest.mark.parametrize("qi", range(4))
def test_quantize_zigzag_dequantize_idct_golden(g1, qi):
    q = QSCALES[qi]
    pq = ivc.PatchQuant(quantization_scale=q)
    tab = pq.get_quantization_table()
    assert np.array_equal(tab, g1[f"table{qi}"]) and tab.dtype == g1[f"table{qi}"].dtype
    qz = pq.quantize(g1["coef"])
    assert qz.dtype == np.int32 and qz.shape == (6, 8, 3, 8, 8)
    zz = ivc.ZigZag().flatten(qz)
    assert np.array_equal(zz, g1[f"zz{qi}"])
    un = ivc.ZigZag().unflatten(zz)
    assert np.array_equal(un, qz)
    dq = pq.dequantize(un)
    assert dq.dtype == np.int32 and np.array_equal(dq, g1[f"dq{qi}"])
    rec = ivc.DiscreteCosineTransform().inverse_transform(dq)
    assert rec.dtype == np.float64
    assert np.array_equal(ivc.Patcher().unpatch(rec), g1[f"rec{qi}"])


@pytest.mark.parametrize("qi", range(4))
def test_luma_broadcast_and_ties_golden(g2, qi):
    """C=1 input broadcasts to 3 channels; the frame holds exact rounding ties (DC/16 = k+0.5)."""
    luma = g2["luma"]
    pq = ivc.PatchQuant(quantization_scale=QSCALES[qi])
    coef = ivc.DiscreteCosineTransform().transform(ivc.Patcher().patch(luma[..., None]))
    assert np.array_equal(coef, g2["coef"])
    qz = pq.quantize(coef)
    assert qz.shape == (5, 7, 3, 8, 8)
    assert np.array_equal(ivc.ZigZag().flatten(qz), g2[f"zz{qi}"])
    dq = pq.dequantize(qz[:, :, :1])
    assert dq.shape == (5, 7, 3, 8, 8)
    rec = ivc.Patcher().unpatch(ivc.DiscreteCosineTransform().inverse_transform(dq))
    assert np.array_equal(rec, g2[f"rec{qi}"])


def test_other_dtypes_golden(g3):
    D = ivc.DiscreteCosineTransform()
    c32 = D.transform(g3["x32"])
    assert c32.dtype == np.float32 and np.array_equal(c32, g3["c32"])
    i32 = D.inverse_transform(g3["x32"])
    assert i32.dtype == np.float32 and np.array_equal(i32, g3["i32"])
    cu8 = D.transform(g3["xu8"])
    assert cu8.dtype == np.float64 and np.array_equal(cu8, g3["cu8"])
    c_one = D.transform(g3["one"])                                  # plain (8,8) input
    assert c_one.shape == (8, 8) and np.array_equal(c_one, g3["c_one"])
    assert np.array_equal(ivc.PatchQuant().quantize(g3["xu8"]), g3["q_u8"])       # float32 division path
    assert np.array_equal(ivc.PatchQuant(0.07).quantize(g3["c32"]), g3["q_f32"])
    q388 = ivc.PatchQuant().quantize(g3["c32"][0, 0])
    assert q388.shape == (1, 1, 3, 8, 8)
    assert np.array_equal(q388, O.quantize(g3["c32"][0, 0], g3["tab1"]))


def test_zigzag_dtypes_and_roundtrip():
    rng = np.random.default_rng(3)
    Z = ivc.ZigZag()
    for dt in (np.int32, np.float64, np.float32, np.int16, np.uint8, np.int64):
        x = rng.integers(0, 100, size=(3, 5, 2, 8, 8)).astype(dt)
        f = Z.flatten(x)
        assert f.dtype == dt and np.array_equal(f, O.zigzag_flatten(x))
        assert np.array_equal(Z.unflatten(f), x)
    with pytest.raises(ValueError):
        Z.flatten(np.zeros((2, 2, 8, 8)))


# ---------------------------------------------------------------- fused kernels
@pytest.mark.parametrize("qi", range(4))
def test_fused_intra_forward_inverse_golden(g1, qi):
    coder = ivc.IntraBlockCoder(quantization_scale=QSCALES[qi])
    zz = coder.forward(g1["img"])
    assert zz.dtype == np.int32 and np.array_equal(zz, g1[f"zz{qi}"])
    rec = coder.inverse(g1[f"zz{qi}"])
    assert rec.dtype == np.float64 and np.array_equal(rec, g1[f"rec{qi}"])


@pytest.mark.parametrize("qi", range(4))
def test_fused_luma_golden(g2, qi):
    coder = ivc.IntraBlockCoder(quantization_scale=QSCALES[qi])
    zz = coder.forward(g2["luma"])                                   # [H,W] -> C=1 -> 3 tables
    assert np.array_equal(zz, g2[f"zz{qi}"])
    rec = coder.inverse(g2[f"zz{qi}"][:, :, :1])                     # 1 scan channel -> 3 image channels
    assert np.array_equal(rec, g2[f"rec{qi}"])


@pytest.mark.parametrize("shape", [(8, 8, 3), (8, 24, 3), (16, 40, 3), (24, 104, 1), (8, 8, 1), (32, 200, 1),
                                   (40, 136, 3)])
def test_fused_ragged_tile_widths_vs_oracle(shape):
    """widths that are not a multiple of the 4-block (C=3) / 12-block (C=1) tile"""
    rng = np.random.default_rng(sum(shape))
    img = rng.uniform(-50, 300, size=shape)
    coder = ivc.IntraBlockCoder(quantization_scale=0.3)
    tab = coder.quant.get_quantization_table()
    zz = coder.forward(img)
    assert np.array_equal(zz, O.intra_forward(img, tab))
    zin = zz if shape[2] == 3 else zz[:, :, :1]
    assert np.array_equal(coder.inverse(zin), O.intra_inverse(zin, tab))


def test_fused_batch_equals_per_frame():
    rng = np.random.default_rng(8)
    batch = rng.uniform(0, 255, size=(5, 24, 40, 3))
    coder = ivc.IntraBlockCoder(quantization_scale=1.0)
    zz = coder.forward(batch)
    for i in range(5):
        assert np.array_equal(zz[i], O.intra_forward(batch[i], coder.quant.get_quantization_table()))
    rec = coder.inverse(zz)
    for i in range(5):
        assert np.array_equal(rec[i], O.intra_inverse(zz[i], coder.quant.get_quantization_table()))


def test_quantizer_extremes_match_x86_semantics():
    """huge / non-finite coefficients: the int32 cast follows numpy-on-x86 (0x80000000)."""
    x = np.zeros((1, 1, 3, 8, 8))
    x[0, 0, 0, 0, :4] = [1e300, -1e300, np.inf, np.nan]
    x[0, 0, 1, 0, :4] = [2147483647.0 * 17, -2147483648.0 * 17, 40000.0 * 17, -40000.5 * 17]
    tab = O.quant_table(1.0)
    with np.errstate(all="ignore"):
        want = O.quantize(x, tab)
    assert np.array_equal(ivc.PatchQuant().quantize(x), want)


# ---------------------------------------------------------------- motion
@pytest.mark.parametrize("name", ME_CASES)
@pytest.mark.parametrize("mode", ["auto", "exact"])
def test_motion_vectors_golden(g4, name, mode):
    sr = case_sr(name)
    mc = ivc.MotionCompensator(search_range=sr, me_mode=mode)
    mv = mc.compute_motion_vector(g4[f"{name}__ref"], g4[f"{name}__cur"])
    assert mv.dtype == np.int64 and mv.shape == g4[f"{name}__mv"].shape
    assert np.array_equal(mv, g4[f"{name}__mv"])
    pred = mc.reconstruct_with_motion_vector(g4[f"{name}__ref"][..., None], g4[f"{name}__mv"])
    assert pred.dtype == g4[f"{name}__pred"].dtype and np.array_equal(pred, g4[f"{name}__pred"])


def test_motion_int_kernel_forced(g4):
    for name in ("int_sr4", "int_sr2", "int_sr16", "flat_sr4", "flat255_sr3"):
        mc = ivc.MotionCompensator(search_range=case_sr(name), me_mode="int")
        assert np.array_equal(mc.compute_motion_vector(g4[f"{name}__ref"], g4[f"{name}__cur"]), g4[f"{name}__mv"])


def test_mc_out_of_frame_vectors_golden(g4):
    mc = ivc.MotionCompensator(search_range=4)
    assert np.array_equal(mc.reconstruct_with_motion_vector(g4["mc_ref3"], g4["mc_mv_rand"]), g4["mc_pred3"])


def test_qcif_motion_golden(g7):
    """cfg2: QCIF sequence, vectors produced by the reference's own loops (open loop + float reference)."""
    seq = O.moving_sequence(2, 6, 144, 176)
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))
    assert sha(seq) == pin["hashes"]["cfg2_seq_sha"]
    mc = ivc.MotionCompensator(search_range=4)
    for t in range(1, 6):
        assert np.array_equal(mc.compute_motion_vector(seq[t - 1], seq[t])[..., 0], g7["mvs"][t - 1][..., 0])
    ref_q = seq[0] + np.random.default_rng(9).normal(0, 0.5, size=seq[0].shape)
    assert np.array_equal(mc.compute_motion_vector(ref_q, seq[1])[..., 0], g7["mv_float"][..., 0])


@pytest.mark.parametrize("sr", [1, 4, 7, 16])
def test_motion_float_frames_vs_oracle(sr):
    rng = np.random.default_rng(100 + sr)
    ref = rng.uniform(0, 255, size=(72, 200))
    cur = np.roll(ref, (2, -3), axis=(0, 1)) + rng.normal(0, 2.0, size=ref.shape)
    for dt in (np.float64, np.float32):
        r, c = ref.astype(dt), cur.astype(dt)
        mv = ivc.MotionCompensator(search_range=sr).compute_motion_vector(r, c)
        assert np.array_equal(mv, O.me_full_search(r, c, sr))


def test_motion_near_ties_float():
    """periodic texture: many candidates have almost equal SSD, so the summation order matters."""
    rng = np.random.default_rng(77)
    base = np.tile(rng.uniform(0, 255, size=(8, 8)), (6, 12))
    ref = base + rng.normal(0, 1e-7, size=base.shape)
    cur = base + rng.normal(0, 1e-7, size=base.shape)
    mv = ivc.MotionCompensator(search_range=8).compute_motion_vector(ref, cur)
    assert np.array_equal(mv, O.me_full_search(ref, cur, 8))


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32])
def test_motion_integer_dtype_frames_wrap_like_numpy(dtype):
    """on integer-dtype frames numpy evaluates (block - ref_block)**2 in that dtype (uint8 wraps mod 256,
    255**2 == -511 in int16) and sums in 64 bits; the vectors must equal the reference loop's on such inputs."""
    rng = np.random.default_rng(77)
    for sr in (2, 4):
        hi = 256 if dtype != np.int32 else 70000
        ref = rng.integers(0, hi, size=(24, 40)).astype(dtype)
        cur = np.roll(ref, (1, -2), axis=(0, 1))
        cur[8:16, 8:24] = rng.integers(0, hi, size=(8, 16)).astype(dtype)
        want = O.me_full_search_loops(ref, cur, sr)
        got = ivc.MotionCompensator(search_range=sr).compute_motion_vector(ref, cur)
        assert got.dtype == want.dtype and np.array_equal(got, want)
        if dtype == np.uint8:                                   # and it differs from the float result: the wrap matters
            assert not np.array_equal(want, O.me_full_search(ref.astype(np.float64), cur.astype(np.float64), sr))


@pytest.mark.parametrize("sr", [3, 4, 16])
def test_motion_uint8_planes_equal_float_search(sr):
    """PFrameBlockCoder.estimate on uint8 planes (float semantics, no conversion pass) == the same search on their
    float64 values == the oracle; aligned (sr % 4 == 0) and byte-wise staging paths, batches, odd tile counts."""
    rng = np.random.default_rng(200 + sr)
    seq = np.clip(O.moving_sequence(90 + sr, 4, 72, 136) + rng.integers(-3, 4, size=(4, 72, 136)), 0, 255).astype(np.uint8)
    ref8, cur8 = torch.from_numpy(seq[:-1]).cuda(), torch.from_numpy(seq[1:]).cuda()
    for mode in ("auto", "int", "exact"):
        pc = ivc.PFrameBlockCoder(1.0, sr, me_mode=mode)
        got = pc.estimate(ref8, cur8)
        assert torch.equal(got, pc.estimate(ref8.double(), cur8.double()))
    want = np.stack([O.me_full_search(seq[i].astype(np.float64), seq[i + 1].astype(np.float64), sr) for i in range(3)])
    assert np.array_equal(got.cpu().numpy(), want)
    # planes at an odd byte address: the byte-wise staging path
    buf_r = torch.empty(72 * 136 + 1, dtype=torch.uint8, device="cuda")
    buf_c = torch.empty(72 * 136 + 3, dtype=torch.uint8, device="cuda")
    r_odd, c_odd = buf_r[1:].view(72, 136), buf_c[3:].view(72, 136)
    r_odd.copy_(ref8[1]); c_odd.copy_(cur8[1])
    assert r_odd.data_ptr() % 4 and np.array_equal(pc.estimate(r_odd, c_odd).cpu().numpy(), want[1])


def test_motion_more_frames_than_one_grid_dimension():
    """The search grid is (tiles, frames); batches above 65535 frames go out in chunks -- 70 000 tiny frame pairs through
    both kernels and the uint8 entry give what the oracle gives on sampled frames (and agree everywhere)."""
    n = 70000
    g = torch.Generator(device="cuda").manual_seed(5)
    ref = torch.randint(0, 256, (n, 8, 16), generator=g, device="cuda", dtype=torch.uint8)
    cur = torch.roll(ref, 1, dims=2)
    a = ivc.PFrameBlockCoder(1.0, 2, me_mode="int").estimate(ref.double(), cur.double())
    b = ivc.PFrameBlockCoder(1.0, 2, me_mode="exact").estimate(ref.double(), cur.double())
    c = ivc.PFrameBlockCoder(1.0, 2).estimate(ref, cur)
    assert torch.equal(a, b) and torch.equal(a, c)
    for i in (0, 65534, 65535, 65536, n - 1):
        want = O.me_full_search(ref[i].cpu().numpy().astype(np.float64), cur[i].cpu().numpy().astype(np.float64), 2)
        assert np.array_equal(a[i].cpu().numpy(), want), i


def test_motion_ragged_frame_raises():
    with pytest.raises(IndexError):
        ivc.MotionCompensator().compute_motion_vector(np.zeros((20, 24)), np.zeros((20, 24)))


# ---------------------------------------------------------------- P-frame step
def test_pframe_fused_golden(g5):
    coder = ivc.PFrameBlockCoder(quantization_scale=0.4, search_range=4)
    mv = coder.estimate(g5["ref"], g5["cur"])
    assert np.array_equal(mv, g5["mv"])
    zz, pred = coder.forward(g5["cur"], g5["ref"], g5["mv"], return_prediction=True)
    assert np.array_equal(pred, g5["pred"]) and np.array_equal(zz, g5["zz"])
    assert np.array_equal(coder.forward(g5["cur"], g5["ref"], g5["mv"]), g5["zz"])
    rec_a = coder.inverse(g5["zz"], pred=g5["pred"])
    rec_b = coder.inverse(g5["zz"], ref=g5["ref"], mv=g5["mv"])
    rec_c = coder.inverse(g5["zz"][:, :, :1], pred=g5["pred"])
    for rec in (rec_a, rec_b, rec_c):
        assert rec.dtype == np.float64 and np.array_equal(rec, g5["recon"])


def test_pframe_out_of_frame_vectors_vs_oracle():
    rng = np.random.default_rng(21)
    H, W, sr = 32, 208, 4
    ref = rng.uniform(0, 255, size=(H, W))
    cur = rng.uniform(0, 255, size=(H, W))
    mv = rng.integers(0, 81, size=(H // 8, W // 8, 1))
    coder = ivc.PFrameBlockCoder(quantization_scale=0.2, search_range=sr)
    tab = coder.quant.get_quantization_table()
    pred_o, zz_o = O.pframe_forward(cur, ref, mv, sr, tab)
    zz, pred = coder.forward(cur, ref, mv, return_prediction=True)
    assert np.array_equal(pred, pred_o) and np.array_equal(zz, zz_o)
    assert np.array_equal(coder.inverse(zz, ref=ref, mv=mv), O.pframe_inverse(zz_o[:, :, :1], pred_o, tab))


def test_pframe_every_vector_ragged_batch_vs_oracle():
    """Every vector of the +-4 window (even and odd dx: the tensor-map gather fetches boxes at even columns only),
    three frames, a width that leaves a ragged last tile (11 blocks = 8 + 3)."""
    rng = np.random.default_rng(33)
    F, H, W, sr = 3, 40, 88, 4
    ref = rng.uniform(0, 255, size=(F, H, W))
    cur = rng.uniform(0, 255, size=(F, H, W))
    mv = (np.arange(F * (H // 8) * (W // 8)) * 7 % 81).reshape(F, H // 8, W // 8, 1)
    coder = ivc.PFrameBlockCoder(quantization_scale=1.0, search_range=sr)
    tab = coder.quant.get_quantization_table()
    d = lambda a: torch.from_numpy(a).cuda()
    zz, pred = coder.forward(d(cur), d(ref), d(mv), return_prediction=True)
    rec = coder.inverse(zz, ref=d(ref), mv=d(mv))
    for f in range(F):
        pred_o, zz_o = O.pframe_forward(cur[f], ref[f], mv[f], sr, tab)
        assert np.array_equal(pred[f].cpu().numpy(), pred_o) and np.array_equal(zz[f].cpu().numpy(), zz_o)
        assert np.array_equal(rec[f].cpu().numpy(), O.pframe_inverse(zz_o[:, :, :1], pred_o, tab))


def test_pframe_reference_plane_only_8_byte_aligned():
    """A reference plane whose base is not 16-byte aligned cannot be described by a tensor map: the cp.async
    gather takes over, same results."""
    rng = np.random.default_rng(34)
    H, W, sr = 24, 72, 4
    buf = torch.from_numpy(rng.uniform(0, 255, size=H * W + 1)).cuda()
    ref_t = buf[1:].view(H, W)
    assert ref_t.data_ptr() % 16 == 8
    cur = rng.uniform(0, 255, size=(H, W))
    mv = rng.integers(0, 81, size=(H // 8, W // 8, 1))
    coder = ivc.PFrameBlockCoder(quantization_scale=0.4, search_range=sr)
    tab = coder.quant.get_quantization_table()
    pred_o, zz_o = O.pframe_forward(cur, ref_t.cpu().numpy(), mv, sr, tab)
    zz = coder.forward(torch.from_numpy(cur).cuda(), ref_t, torch.from_numpy(mv).cuda())
    assert np.array_equal(zz.cpu().numpy(), zz_o)
    rec = coder.inverse(zz, ref=ref_t, mv=torch.from_numpy(mv).cuda())
    assert np.array_equal(rec.cpu().numpy(), O.pframe_inverse(zz_o[:, :, :1], pred_o, tab))


def test_pframe_forward_three_distinct_tables_through_the_abi():
    """PatchQuant always hands over [lum, chrom, chrom]; the C ABI takes any [3, 64] table, and the kernel must not
    assume channels 1 and 2 are equal (it only computes them once when they are)."""
    from ivclab_b200 import _lib
    from ivclab_b200._runtime import dev_index, stream_ptr
    rng = np.random.default_rng(35)
    H, W, sr = 24, 136, 4
    ref, cur = rng.uniform(0, 255, size=(H, W)), rng.uniform(0, 255, size=(H, W))
    mv = rng.integers(0, 81, size=(H // 8, W // 8, 1))
    tab = rng.uniform(0.5, 40.0, size=(3, 8, 8))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dc, dr, dm, dt = d(cur), d(ref), d(mv), d(tab)
    zz = torch.empty((H // 8, W // 8, 3, 64), dtype=torch.int32, device="cuda")
    st = _lib.lib.ivc_pframe_forward(dev_index(dc), stream_ptr(dc.device), dc.data_ptr(), dr.data_ptr(), dm.data_ptr(),
                                     _lib.F64, 1, H, W, sr, dt.data_ptr(), _lib.F64, None, zz.data_ptr())
    _lib.check(st, "ivc_pframe_forward")
    _, zz_o = O.pframe_forward(cur, ref, mv, sr, tab)
    assert np.array_equal(zz.cpu().numpy(), zz_o)


# ---------------------------------------------------------------- full-size configs
def test_cfg1_full_size_hashes():
    """cfg1 (512x768 RGB->YCbCr, qScale 0.07/1/4.5): hashes recorded from the real reference."""
    pin = json.load(open(os.path.join(GOLD, "PINNING.json")))["hashes"]
    img = O.rgb2ycbcr(O.smooth_noise_rgb(0, 512, 768))
    assert sha(img) == pin["cfg1_img_sha"]
    for q in (0.07, 1.0, 4.5):
        coder = ivc.IntraBlockCoder(quantization_scale=q)
        zz = coder.forward(img)
        assert sha(zz) == pin[f"cfg1_zz_sha[{q}]"]
        assert sha(coder.inverse(zz)) == pin[f"cfg1_rec_sha[{q}]"]
        # and the unfused drop-in classes give the same bytes
        pq = ivc.PatchQuant(q)
        zz2 = ivc.ZigZag().flatten(pq.quantize(ivc.DiscreteCosineTransform().transform(ivc.Patcher().patch(img))))
        assert sha(zz2) == pin[f"cfg1_zz_sha[{q}]"]


def test_1080p_batch_vs_oracle_and_roundtrip_property():
    """cfg3 shape: 1080p frames.  Parity vs the oracle on 2 frames; on the device-resident batch the
    size-independent property decode(encode(x)) ~ x within half a quantisation step per coefficient."""
    frames = np.stack([O.rgb2ycbcr(O.smooth_noise_rgb(3000 + i, 1080, 1920)) for i in range(2)])
    coder = ivc.IntraBlockCoder(quantization_scale=0.07)
    tab = coder.quant.get_quantization_table()
    d = torch.from_numpy(frames).cuda()
    zz = coder.forward(d)
    assert zz.is_cuda and zz.dtype == torch.int32
    rec = coder.inverse(zz)
    for i in range(2):
        zo = O.intra_forward(frames[i], tab)
        assert np.array_equal(zz[i].cpu().numpy(), zo)
        assert np.array_equal(rec[i].cpu().numpy(), O.intra_inverse(zo, tab))
    # orthonormal transform: reconstruction error energy <= sum((t/2 + 1)^2) per block
    err = (rec - d).reshape(2, 135, 8, 240, 8, 3)
    e_blk = (err ** 2).sum(dim=(2, 4))
    bound = float(((tab.astype(np.float64) / 2 + 1.0) ** 2).sum(axis=(1, 2)).max())
    assert float(e_blk.max()) <= bound


def test_4k_motion_crops_and_linearity():
    """cfg4 shape: +-16 search.  Bit-exact vs the oracle on 256x256 crops; on the full 4K frame the
    estimator must be invariant to adding a constant to both frames (SSD is translation invariant)
    and the int and exact kernels must agree."""
    seq = O.moving_sequence(4000, 2, 2160, 3840, max_shift=12, obj=128)
    mc = ivc.MotionCompensator(search_range=16)
    crop_r, crop_c = seq[0][512:768, 1024:1280], seq[1][512:768, 1024:1280]
    assert np.array_equal(mc.compute_motion_vector(crop_r, crop_c), O.me_full_search(crop_r, crop_c, 16))
    r, c = torch.from_numpy(seq[0]).cuda(), torch.from_numpy(seq[1]).cuda()
    mv_auto = mc.compute_motion_vector(r, c)
    mv_exact = ivc.MotionCompensator(16, me_mode="exact").compute_motion_vector(r, c)
    assert torch.equal(mv_auto, mv_exact)
    r2, c2 = (r * 0.5 + 3.0), (c * 0.5 + 3.0)           # exact in binary: SSD scales by 1/4, argmin unchanged
    assert torch.equal(ivc.MotionCompensator(16).compute_motion_vector(r2, c2), mv_exact)


def test_cuda_tensor_in_out_no_host_copy(g1):
    d = torch.from_numpy(g1["img"]).cuda()
    p = ivc.Patcher().patch(d)
    out = ivc.DiscreteCosineTransform().transform(p)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert np.array_equal(out.cpu().numpy(), g1["coef"])
    q = ivc.PatchQuant(1.0).quantize(out)
    assert q.is_cuda and np.array_equal(ivc.ZigZag().flatten(q).cpu().numpy(), g1["zz1"])


def test_unmodified_caller_chain_like_intracodec(g1, g6):
    """The call chain of IntraCodec.image2symbols / symbols2image (intracodec.py:66-75, :115-124)
    executed with the drop-in objects, checked against the reference's zero-run symbol stream."""
    dct, quant, zigzag, patcher = ivc.DiscreteCosineTransform(), ivc.PatchQuant(1.0), ivc.ZigZag(), ivc.Patcher()
    patches = patcher.patch(g1["img"])
    zz = zigzag.flatten(quant.quantize(dct.transform(patches)))
    assert np.array_equal(O.zerorun_encode(zz), g6["sym"])
    dec = O.zerorun_decode(g6["sym"], zz.shape[:3])
    ycbcr = patcher.unpatch(dct.inverse_transform(quant.dequantize(zigzag.unflatten(dec))))
    assert np.array_equal(O.ycbcr2rgb(ycbcr), g6["rec_rgb"])
    assert O.calc_psnr(g1["rgb"], O.ycbcr2rgb(ycbcr)) == float(g6["psnr"])


# ---------------------------------------------------------------- closed loop (cfg2 / cfg5 shape)
def test_closed_loop_golden_from_reference_codec_classes():
    """5-frame closed loop recorded from the real IntraCodec + MotionCompensator (oracle/gen_golden_video.py):
    no teacher forcing -- every frame depends on the previous reconstruction and still matches bit for bit."""
    from conftest import load_golden
    g8 = load_golden("g8_closed_loop.npz")
    for use_graph in (False, True):
        coder = ivc.ClosedLoopLumaCoder(float(g8["qscale"]), int(g8["sr"]), decode="faithful", use_graph=use_graph)
        out = coder.code_sequence(g8["frames"])
        assert np.array_equal(out["mv"], g8["mv"])
        assert np.array_equal(out["zz"], g8["zz"])
        assert np.array_equal(out["recon"], g8["recon"])
        out2 = coder.code_sequence(g8["frames"])            # second call replays the captured graph
        assert np.array_equal(out2["recon"], g8["recon"])
    luma = ivc.ClosedLoopLumaCoder(float(g8["qscale"]), int(g8["sr"]), decode="luma").code_sequence(g8["frames"])
    assert np.array_equal(luma["recon"], g8["recon_luma"]) and np.array_equal(luma["mv"], g8["mv_luma"])


def test_closed_loop_qcif_21_frames_vs_oracle():
    """cfg2: QCIF 176x144, 21 frames, +-4 search, closed loop, against the oracle's loop."""
    from oracle import closed_loop as CL
    frames = O.moving_sequence(2, 21, 144, 176)
    want = CL.code_sequence(frames, 1.0, 4, "luma")
    got = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="auto").code_sequence(frames)
    assert np.array_equal(got["mv"], want["mv"])
    assert np.array_equal(got["zz"], want["zz"])
    assert np.array_equal(got["recon"], want["recon"])
    psnr = 10 * np.log10(255.0 ** 2 / np.mean((got["recon"] - frames) ** 2))
    assert psnr > 20.0


@pytest.mark.parametrize("decode", ["faithful", "luma"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_closed_loop_lockstep_sequences(decode, use_graph):
    """Several independent sequences coded in lockstep (one launch per kernel and time step) give exactly what
    coding them one after the other gives -- which is what the oracle's loop gives."""
    from oracle import closed_loop as CL
    seqs = np.stack([O.moving_sequence(31 + s, 4, 40, 72) for s in range(3)])
    coder = ivc.ClosedLoopLumaCoder(0.4, 3, decode=decode, me_mode="auto", use_graph=use_graph)
    got = coder.code_sequences(seqs)
    assert got["zz"].shape == (3, 4, 5, 9, 3, 64) and got["mv"].shape == (3, 3, 5, 9, 1) and got["recon"].shape == seqs.shape
    for s in range(3):
        want = CL.code_sequence(seqs[s], 0.4, 3, decode)
        for k in ("zz", "mv", "recon"):
            assert np.array_equal(got[k][s], want[k]), (s, k)
        one = coder.code_sequence(seqs[s])
        for k in ("zz", "mv", "recon"):
            assert np.array_equal(one[k], want[k]), (s, k)


def test_rd_sweep_properties_at_1080p():
    """cfg3 shape (1080p colour frames, the ten qScales of the sweep), properties that need no oracle: re-encoding a
    reconstruction reproduces its indices when every table entry exceeds 2 (dequantisation truncates by < 1, i.e. by
    less than half a quantisation step); distortion grows and the symbol count shrinks monotonically with qScale; the
    fused decoder's distortion equals the three-kernel form's."""
    g = torch.Generator(device="cuda").manual_seed(33)
    rgb = (torch.nn.functional.avg_pool2d(torch.rand((6, 3, 1080, 1920), generator=g, device="cuda") * 255, 5, 1, 2)
           .permute(0, 2, 3, 1).contiguous()).to(torch.uint8)
    qs = [0.07, 0.2, 0.4, 0.8, 1.0, 1.5, 2, 3, 4, 4.5]
    sse_all, nsym = [], []
    for q in qs:
        coder = ivc.IntraBlockCoder(q)
        zz = coder.forward_rgb(rgb)
        sse, rec = coder.inverse_with_distortion(zz, rgb, space="rgb", return_reconstruction=True)
        if q >= 0.4:
            assert torch.equal(coder.forward(rec), zz), q                     # idempotent
        if q in (0.07, 1.0, 4.5):
            want = ivc.frame_sse(rgb, ivc.ycbcr2rgb(rec))
            assert torch.allclose(sse, want, rtol=1e-12, atol=0)
        sse_all.append(sse.sum().item())
        nsym.append(ivc.ZeroRunCoder().encode(zz).numel())
    assert all(a < b for a, b in zip(sse_all, sse_all[1:])), sse_all          # distortion up
    assert all(a > b for a, b in zip(nsym, nsym[1:])), nsym                   # rate down


def test_closed_loop_full_size_decoder_replay():
    """cfg5 at its full size (300 x 1080p luma, +-4): a DECODER that only sees the scan indices and vectors rebuilds
    every reconstruction bit for bit (I-frame: intra inverse; P-frames: prediction from its own previous output),
    the vectors are what an independent search on (previous reconstruction, frame) gives, and coding is deterministic."""
    from bench_configs import luma_seq
    T = 300
    seq = luma_seq(torch, torch.device("cuda"), T, 1080, 1920, 5000)
    enc = ivc.ClosedLoopLumaCoder(1.0, 4, decode="luma", me_mode="exact")
    out = enc.code_sequence(seq)
    assert out["zz"].shape == (T, 135, 240, 3, 64) and out["mv"].shape == (T - 1, 135, 240, 1)
    pc = ivc.PFrameBlockCoder(1.0, 4, me_mode="exact")
    prev = pc.inverse(out["zz"][:1], pred=torch.zeros((1, 1080, 1920), dtype=torch.float64, device="cuda"))[0]
    assert torch.equal(prev, out["recon"][0])
    for t in range(1, T):
        if t in (1, 2, 150, T - 1):                              # the encoder's search, re-run from the outside
            assert torch.equal(pc.estimate(prev, seq[t]), out["mv"][t - 1])
        prev = pc.inverse(out["zz"][t], ref=prev, mv=out["mv"][t - 1])
        assert torch.equal(prev, out["recon"][t]), t
    psnr = 10 * torch.log10(255.0 ** 2 / ((out["recon"] - seq) ** 2).mean(dim=(1, 2)))
    assert float(psnr.min()) > 24.0 and float(psnr.max()) < 40.0       # busy synthetic texture at qScale 1
    again = enc.code_sequence(seq)
    assert all(torch.equal(out[k], again[k]) for k in ("zz", "mv", "recon"))


# ---------------------------------------------------------------- "next" rows N2 / N3
def test_zerorun_encode_matches_reference_stream(g1, g6):
    """N2: the GPU zero-run encoder reproduces the reference's symbol list (zerorun.py:10-43)."""
    zr = ivc.ZeroRunCoder()
    sym = zr.encode(g1["zz1"])
    assert sym.dtype == np.int32 and np.array_equal(sym, g6["sym"])
    assert np.array_equal(zr.decode(sym, g1["zz1"].shape[:3]), g1["zz1"])
    assert np.array_equal(zr.decode(sym, (6, 8, 1)), g6["dec_trunc"])            # the 2-D-shape truncation rule
    rng = np.random.default_rng(4)
    for density in (0.0, 0.02, 0.3, 1.0):
        zz = (rng.integers(-9, 10, size=(7, 9, 3, 64)) * (rng.random((7, 9, 3, 64)) < density)).astype(np.int32)
        zz[0, 0, 0, :] = 0
        zz[0, 0, 1, 63] = 5                                                      # run of 63 zeros then a value
        zz[0, 0, 2, :] = np.arange(1, 65)                                        # no zeros at all
        assert np.array_equal(zr.encode(zz), O.zerorun_encode(zz))
    big = ivc.IntraBlockCoder(0.4).forward(O.rgb2ycbcr(O.smooth_noise_rgb(5, 256, 384)))
    assert np.array_equal(zr.encode(big), O.zerorun_encode(big))


def test_zerorun_decode_on_device(g9):
    """N2, decoding direction (zerorun.py:44-87): blocks, truncation rule and error conditions recorded from the
    real reference; round trips at 1080p size; cuda tensors in -> cuda tensor out."""
    zr = ivc.ZeroRunCoder()
    assert np.array_equal(zr.decode(g9["sym"], (6, 8, 3)), g9["dec_full"])
    assert np.array_equal(zr.decode(g9["sym"], [6, 8, 1]), g9["dec_trunc"])        # later symbols are ignored
    assert np.array_equal(zr.decode(g9["sym2"], (5, 7, 3)), g9["blocks"])
    assert np.array_equal(zr.decode(list(g9["sym2"]), (5, 7, 3)), g9["blocks"])    # the reference is handed lists too
    with pytest.raises(ValueError, match="Unexpected end"):
        zr.decode(g9["sym2"][:-1], (5, 7, 3))
    with pytest.raises(ValueError, match="Expected 140 blocks, got 105"):
        zr.decode(g9["sym2"], (5, 7, 4))
    with pytest.raises(ValueError, match="Block size exceeded"):
        zr.decode(np.concatenate([np.arange(1, 66), [4000]]).astype(np.int32), (1, 1, 1))
    with pytest.raises(ValueError, match="Block size exceeded"):
        zr.decode(np.array([7, 0, 4000, 4000], dtype=np.int32), (1, 1, 1))         # EOB value in a run-length slot
    with pytest.raises(ValueError, match="Expected 2 blocks, got 0"):
        zr.decode(np.zeros(0, dtype=np.int32), (1, 2, 1))
    rng = np.random.default_rng(11)
    for density in (0.0, 0.03, 0.5, 1.0):
        zz = (rng.integers(-300, 301, size=(9, 11, 3, 64)) * (rng.random((9, 11, 3, 64)) < density)).astype(np.int32)
        zz[0, 0, 1, 63] = 5
        zz[0, 0, 2, :] = np.arange(1, 65)
        sym = O.zerorun_encode(zz)
        assert np.array_equal(zr.decode(sym, zz.shape[:3]), zz)
        assert np.array_equal(zr.decode(sym, zz.shape[:3]), O.zerorun_decode(sym, zz.shape[:3]))
    big = ivc.IntraBlockCoder(0.2).forward(torch.from_numpy(O.rgb2ycbcr(O.smooth_noise_rgb(6, 1080, 1920))).cuda())
    sym = zr.encode(big)
    back = zr.decode(sym, big.shape[:3])
    assert back.is_cuda and back.dtype == torch.int32 and torch.equal(back, big)
    luma_only = zr.decode(sym, (135, 240, 1))                                      # first Hp*Wp blocks
    assert torch.equal(luma_only.reshape(-1, 64), big.reshape(-1, 64)[:135 * 240])


def test_intracodec_symbol_level_calls(g9):
    """IntraCodec.image2symbols / symbols2image / the statistics half of train_huffman_from_image against outputs of
    the REAL reference codec (oracle/gen_golden_entropy.py): colour image, luma plane (2-D), ragged size."""
    codec = ivc.IntraCodec(quantization_scale=0.4)
    sym = codec.image2symbols(g9["img"], is_source_rgb=True)                       # uint8, 48x64: fused colour front end
    assert sym.dtype == np.int32 and np.array_equal(sym, g9["sym"])
    assert np.array_equal(codec.image2symbols(g9["img"].astype(np.float64), True), g9["sym"])   # two-kernel route
    assert np.array_equal(codec.symbols2image(sym, g9["img"].shape), g9["rec_rgb"])
    assert np.array_equal(codec.symbols2image(list(sym), g9["img"].shape), g9["rec_rgb"])
    sl = codec.image2symbols(g9["luma"], is_source_rgb=False)
    assert np.array_equal(sl, g9["sym_luma"])
    assert np.array_equal(codec.symbols2image(sl, g9["luma"].shape), g9["rec_luma"])           # (H, W, 3): SURVEY A13
    assert np.array_equal(codec.image2symbols(g9["ragged"], True), g9["sym_ragged"])           # edge padding to 48x64
    assert codec.bounds is None
    codec.train_huffman_from_image(g9["img"], is_source_rgb=True)
    assert codec.bounds == (int(g9["lo"]), int(g9["hi"]))
    want = g9["pmf"] + 1e-9
    assert np.array_equal(codec.pmf, want / want.sum())
    dev = codec.image2symbols(torch.from_numpy(g9["img"]).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), g9["sym"])
    with pytest.raises(NotImplementedError):
        codec.intra_encode(g9["img"])


def test_symbol_statistics_match_reference(g9):
    """N3: stats_marg over the bins IntraCodec.train_huffman_from_image picks (entropy.py:6-29,
    intracodec.py:160-166), bit-identical pmf; min/max on the device."""
    lo, hi = int(g9["lo"]), int(g9["hi"])
    assert ivc.symbol_minmax(g9["sym"]) == (lo + 20, hi - 21)
    pmf = ivc.stats_marg(g9["sym"], np.arange(lo, hi))
    assert pmf.dtype == np.float64 and np.array_equal(pmf, g9["pmf"])
    assert np.array_equal(ivc.stats_marg(g9["sym"], np.arange(-3, 9)), g9["pmf_cut"])      # clipped range, closed last bin
    assert np.array_equal(ivc.stats_marg(g9["img8"], np.arange(256)), g9["pmf8"])
    assert np.array_equal(ivc.stats_marg(g9["img8"].astype(np.float64), np.arange(256)), g9["pmf8"])
    with pytest.raises(NotImplementedError):
        ivc.stats_marg(g9["sym"], np.arange(0, 64, 2))
    rng = np.random.default_rng(12)
    x = rng.integers(-70000, 70001, size=300001).astype(np.int32)                         # wide range: global-atomic path
    assert np.array_equal(ivc.stats_marg(x, np.arange(-70000, 70002)), O.stats_marg(x, np.arange(-70000, 70002)))
    assert ivc.symbol_minmax(x) == (int(x.min()), int(x.max()))
    zz = ivc.IntraBlockCoder(1.0).forward(torch.from_numpy(O.rgb2ycbcr(O.smooth_noise_rgb(7, 1080, 1920))).cuda())
    sym = ivc.ZeroRunCoder().encode(zz)
    s = sym.cpu().numpy()
    b = O.symbol_bounds(s)
    assert ivc.symbol_minmax(sym) == (b[0] + 20, b[1] - 21)
    assert np.array_equal(ivc.stats_marg(sym, np.arange(*b)), O.stats_marg(s, np.arange(*b)))


def test_fused_colour_sse_equals_two_step_form():
    """frame_sse_rgb8_vs_ycbcr(rgb, rec) == frame_sse(rgb2ycbcr(rgb), rec) bit for bit (same visiting order)."""
    rng = np.random.default_rng(21)
    rgb = torch.from_numpy(rng.integers(0, 256, size=(3, 72, 104, 3), dtype=np.uint8)).cuda()
    rec = ivc.rgb2ycbcr(rgb) + torch.from_numpy(rng.normal(0, 3, size=(3, 72, 104, 3))).cuda()
    assert torch.equal(ivc.frame_sse_rgb8_vs_ycbcr(rgb, rec), ivc.frame_sse(ivc.rgb2ycbcr(rgb), rec))
    want = ((O.rgb2ycbcr(rgb.cpu().numpy()) - rec.cpu().numpy()) ** 2).reshape(3, -1).sum(1)
    assert np.allclose(ivc.frame_sse_rgb8_vs_ycbcr(rgb, rec).cpu().numpy(), want, rtol=1e-12, atol=0)


def test_decoder_with_rgb_store_equals_two_passes():
    """ivc_intra_inverse_rgb == ycbcr2rgb(ivc_intra_inverse(...)) bit for bit == the oracle's symbols2image tail."""
    rgb = np.stack([O.smooth_noise_rgb(70 + i, 56, 88) for i in range(2)])
    for q in (0.07, 1.0):
        coder = ivc.IntraBlockCoder(q)
        zz = coder.forward(np.stack([O.rgb2ycbcr(f) for f in rgb]))
        got = coder.inverse(zz, to_rgb=True)
        assert np.array_equal(got, ivc.ycbcr2rgb(coder.inverse(zz)))
        tab = coder.quant.get_quantization_table()
        assert np.array_equal(got[1], O.ycbcr2rgb(O.intra_inverse(zz[1], tab)))
    luma = ivc.IntraBlockCoder(1.0).forward(O.smooth_noise_luma(3, 40, 48)[..., None])[:, :, :1]
    assert np.array_equal(ivc.IntraBlockCoder(1.0).inverse(luma, to_rgb=True),
                          ivc.ycbcr2rgb(ivc.IntraBlockCoder(1.0).inverse(luma)))


@pytest.mark.parametrize("space", ["rgb", "ycbcr"])
def test_fused_decode_and_distortion(space):
    """ivc_intra_inverse_sse: the reconstruction equals the plain decoder's bit for bit; the squared error equals
    the reference's PSNR pipeline (ycbcr2rgb + calc_mse, or rgb2ycbcr of the original) within reduction order."""
    rgb = np.stack([O.smooth_noise_rgb(40 + i, 72, 112) for i in range(3)])
    for q in (0.07, 1.0, 4.5):
        coder = ivc.IntraBlockCoder(q)
        zz = coder.forward_rgb(torch.from_numpy(rgb).cuda())
        sse, rec = coder.inverse_with_distortion(zz, torch.from_numpy(rgb).cuda(), space=space, return_reconstruction=True)
        assert torch.equal(rec, coder.inverse(zz))
        only = coder.inverse_with_distortion(zz, torch.from_numpy(rgb).cuda(), space=space)
        assert torch.equal(only, sse)                                              # same sums without the store
        rec_np = rec.cpu().numpy()
        if space == "rgb":
            want = [((rgb[i].astype(np.float64) - O.ycbcr2rgb(rec_np[i])) ** 2).sum() for i in range(3)]
            psnr_ref = [O.calc_psnr(rgb[i], O.ycbcr2rgb(rec_np[i])) for i in range(3)]
            psnr = [20 * np.log10(255.0 / np.sqrt(float(s) / rgb[i].size)) for i, s in enumerate(sse.cpu().numpy())]
            assert np.allclose(psnr, psnr_ref, rtol=0, atol=1e-9)                 # north star: PSNR within 0.01 dB
        else:
            want = [((O.rgb2ycbcr(rgb[i]) - rec_np[i]) ** 2).sum() for i in range(3)]
        assert np.allclose(sse.cpu().numpy(), want, rtol=1e-12, atol=0)
    one = coder.inverse_with_distortion(zz[0], torch.from_numpy(rgb[0]).cuda(), space=space)
    assert one.ndim == 0 and float(one) == float(sse[0])
    with pytest.raises(ValueError):
        coder.inverse_with_distortion(zz, torch.from_numpy(rgb.astype(np.float64)).cuda())


def test_metrics_match_reference(g1, g6):
    """N3: calc_mse / calc_psnr (metrics.py:3-40); reduction order differs from numpy's pairwise mean,
    so the comparison is relative 1e-12 (PSNR: far below the 0.01 dB bar)."""
    rgb, rec = g1["rgb"], g6["rec_rgb"]
    assert abs(ivc.calc_mse(rgb, rec) / float(g6["mse"]) - 1.0) < 1e-12
    assert abs(ivc.calc_psnr(rgb, rec) - float(g6["psnr"])) < 1e-9
    gray = rgb[..., 0]
    assert abs(ivc.calc_mse(gray, rec) / O.calc_mse(gray, rec) - 1.0) < 1e-12     # gray vs RGB stacking
    assert abs(ivc.calc_mse(rec, gray) / O.calc_mse(rec, gray) - 1.0) < 1e-12
    a = np.random.default_rng(2).uniform(0, 255, size=(5, 40, 56, 3))
    b = a + np.random.default_rng(3).normal(0, 2, size=a.shape)
    sse = ivc.frame_sse(a, b).cpu().numpy()
    want = ((a - b) ** 2).reshape(5, -1).sum(axis=1)
    assert np.allclose(sse, want, rtol=1e-12, atol=0)
    assert np.array_equal(sse, ivc.frame_sse(a, b).cpu().numpy())                 # deterministic


def test_colour_transforms_and_fused_rgb_front_end(g1, g6):
    """N1: rgb2ycbcr replays numpy's BLAS FMA chain, ycbcr2rgb the elementwise ops; both bit-identical
    to the reference's outputs recorded in the golden files, and the fused uint8 front end of K1
    produces the same indices as the two-step route."""
    rgb = g1["rgb"]
    ycc = ivc.rgb2ycbcr(rgb)
    assert ycc.dtype == np.float64 and np.array_equal(ycc, g1["img"])            # golden: reference rgb2ycbcr here
    assert np.array_equal(ivc.rgb2ycbcr(rgb.astype(np.float32)), g1["img"])      # videocodec.py:38 feeds float32
    assert np.array_equal(ivc.ycbcr2rgb(g1["rec1"]), g6["rec_rgb"])
    big = O.smooth_noise_rgb(8, 128, 192)
    assert np.array_equal(ivc.rgb2ycbcr(big), O.rgb2ycbcr(big))                  # vs numpy on this box
    wild = np.random.default_rng(1).uniform(-500, 700, size=(64, 64, 3))
    assert np.array_equal(ivc.ycbcr2rgb(wild), O.ycbcr2rgb(wild))
    for q in (0.07, 1.0):
        coder = ivc.IntraBlockCoder(q)
        assert np.array_equal(coder.forward_rgb(rgb), g1[f"zz{0 if q == 0.07 else 1}"])
        batch = np.stack([O.smooth_noise_rgb(30 + i, 64, 160) for i in range(3)])
        assert np.array_equal(coder.forward_rgb(batch), coder.forward(np.stack([O.rgb2ycbcr(b) for b in batch])))
        odd = O.smooth_noise_rgb(9, 24, 40)                                      # W % 16 != 0: two-kernel route
        assert np.array_equal(coder.forward_rgb(odd), O.intra_forward(O.rgb2ycbcr(odd), coder.quant.get_quantization_table()))


def test_streamed_coder_matches_direct_calls():
    """host-fed streaming (3 streams, chunked) returns exactly what the direct API calls return"""
    F, H, W = 5, 64, 96
    rgb = np.stack([O.smooth_noise_rgb(50 + i, H, W) for i in range(F)])
    seq = O.moving_sequence(60, F + 1, H, W).astype(np.uint8)
    cur, ref = seq[1:], seq[:-1]
    sc = ivc.StreamedCoder(0.4, 4, chunk_frames=2)
    out = sc.run(rgb, cur, ref)
    assert sc.symbol_dtype == torch.int16 and out["sym_intra"].dtype == torch.int16       # lossless 16-bit transfer format
    out32 = ivc.StreamedCoder(0.4, 4, chunk_frames=2, symbol_dtype=torch.int32).run(rgb, cur, ref)
    assert out32["sym_intra"].dtype == torch.int32 and torch.equal(out32["sym_intra"], out["sym_intra"].to(torch.int32))
    assert torch.equal(out32["sym_inter"], out["sym_inter"].to(torch.int32)) and out32["d2h_bytes"] > out["d2h_bytes"]
    with pytest.raises(ValueError):
        ivc.StreamedCoder(0.001, 4, symbol_dtype=torch.int16)                              # 2040 / 0.01 does not fit
    assert ivc.StreamedCoder(0.001, 4).symbol_dtype == torch.int32
    intra, pc, zr = ivc.IntraBlockCoder(0.4), ivc.PFrameBlockCoder(0.4, 4), ivc.ZeroRunCoder()
    tab = intra.quant.get_quantization_table()
    sym_i, sym_p = [], []
    for i in range(F):
        zz = O.intra_forward(O.rgb2ycbcr(rgb[i]), tab)
        sym_i.append(O.zerorun_encode(zz))
        mv = O.me_full_search(ref[i].astype(np.float64), cur[i].astype(np.float64), 4)
        assert np.array_equal(out["mv"][i].numpy(), mv)
        pred, zzp = O.pframe_forward(cur[i].astype(np.float64), ref[i].astype(np.float64), mv, 4, tab)
        sym_p.append(O.zerorun_encode(zzp))
        rec = O.intra_inverse(zz, tab)
        assert abs(out["sse"][0, i].item() / ((O.rgb2ycbcr(rgb[i]) - rec) ** 2).sum() - 1) < 1e-12
    assert np.array_equal(out["sym_intra"].numpy(), np.concatenate(sym_i))
    assert sc.inter_channels == 2                 # transfer format: channels 0 and 1; expand_inter rebuilds the reference's stream
    assert np.array_equal(ivc.StreamedCoder.expand_inter(out["sym_inter"]), np.concatenate(sym_p))
    assert sum(out["len_intra"]) == out["sym_intra"].numel()
    # sequence mode: the references are implied (frame t-1), every luma frame is uploaded once -- same results
    keep = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    for chunk, slots in ((2, 3), (3, 2), (8, 2)):
        out2 = ivc.StreamedCoder(0.4, 4, chunk_frames=chunk, slots=slots).run(rgb, cur, first_ref=seq[0])
        for k in ("sym_intra", "sym_inter", "mv", "sse"):
            assert torch.equal(out2[k], keep[k]), (chunk, k)
        assert out2["h2d_bytes"] == rgb.size + cur.size + seq[0].size


def test_luma8_from_rgb8_vs_numpy_statement():
    """clip(np.round(rgb2ycbcr(rgb)[..., 0]), 0, 255) on every (r, g, b) on a 64-step lattice plus random pixels and
    the extremes; odd pixel counts and unaligned bases take the scalar path."""
    rng = np.random.default_rng(71)
    lat = np.stack(np.meshgrid(*[np.r_[0:256:5, 255]] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.uint8)
    rgb = np.concatenate([lat, rng.integers(0, 256, size=(100003, 3), dtype=np.uint8)])
    want = np.clip(np.round(O.rgb2ycbcr(rgb)[..., 0]), 0, 255).astype(np.uint8)
    assert np.array_equal(ivc.luma8_from_rgb8(rgb), want)
    d = torch.from_numpy(rgb).cuda()
    flat = torch.empty(rgb.size + 1, dtype=torch.uint8, device="cuda")
    flat[1:] = d.reshape(-1)
    assert np.array_equal(ivc.luma8_from_rgb8(flat[1:].view(-1, 3)).cpu().numpy(), want)       # base not 4-byte aligned
    f64 = torch.empty(len(rgb), dtype=torch.float64, device="cuda")               # the same plane as float64, in one pass
    u8 = ivc.luma8_from_rgb8(d, out_f64=f64)
    assert np.array_equal(u8.cpu().numpy(), want) and np.array_equal(f64.cpu().numpy(), want.astype(np.float64))
    img = rgb[: 24 * 40].reshape(24, 40, 3)
    assert ivc.luma8_from_rgb8(img).shape == (24, 40)
    with pytest.raises(ValueError):
        ivc.luma8_from_rgb8(img.astype(np.float64))


def test_streamed_coder_luma_derived_on_device():
    """cur=None: the luma planes are derived from the RGB frames on the device -- same results as handing over the
    planes, one byte per pixel less across PCIe."""
    F, H, W = 7, 64, 96
    rgb = np.stack([O.smooth_noise_rgb(80 + i, H, W) for i in range(F + 1)])
    luma = np.clip(np.round(np.stack([O.rgb2ycbcr(f)[..., 0] for f in rgb])), 0, 255).astype(np.uint8)
    want = ivc.StreamedCoder(0.4, 4, chunk_frames=2).run(rgb[1:], luma[1:], first_ref=luma[0])
    keep = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in want.items()}
    for chunk, slots, graph in ((2, 3, True), (3, 2, True), (4, 2, False), (16, 2, True)):
        got = ivc.StreamedCoder(0.4, 4, chunk_frames=chunk, slots=slots, use_graph=graph).run(rgb[1:], first_ref=rgb[0])
        for k in ("sym_intra", "sym_inter", "mv", "sse"):
            assert torch.equal(got[k], keep[k]), (chunk, k)
        assert got["h2d_bytes"] == rgb.size
    with pytest.raises(ValueError):
        ivc.StreamedCoder(0.4, 4).run(rgb[1:], first_ref=luma[0])                # the first reference must be an RGB frame
    with pytest.raises(ValueError):
        ivc.StreamedCoder(0.4, 4).run(rgb[1:], None, ref=luma[:-1])              # derived planes need sequence mode
