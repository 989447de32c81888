"""Prediction of the encode pass (inter_predictor, VPX/vp9_reconinter.h:23) and the full-loop distortion
(EbEncDecProcess.c:723): the oracle restatement against the reference's own functions (C and AVX2 flavours), the CUDA
path against the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

api = H.api
PAD = 80  # reference pictures carry 64 + 16 samples of padding on every side


def _ref_plane(rng, w, h, kind):
    if kind == "noise":
        a = rng.integers(0, 256, size=(h + 2 * PAD, w + 2 * PAD), dtype=np.uint8)
    elif kind == "extreme":  # saturating both ways: the intermediate and final clips matter
        a = (rng.integers(0, 2, size=(h + 2 * PAD, w + 2 * PAD)) * 255).astype(np.uint8)
    else:
        yy, xx = np.mgrid[0:h + 2 * PAD, 0:w + 2 * PAD]
        a = ((np.sin(xx / 5.0) + np.cos(yy / 7.0)) * 60 + 128 + rng.integers(-6, 7, size=xx.shape)).clip(0, 255).astype(np.uint8)
    return np.ascontiguousarray(a)


def _cases(rng, n, w, h):
    """n blocks of every size the encode pass predicts (luma 4..64, chroma 4..32; the reference's SIMD kernels do not take widths below 4, its encoder never asks for them), every phase pair"""
    sizes = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (64, 32), (32, 64), (16, 8), (8, 16), (32, 16), (4, 8), (8, 4), (64, 16), (16, 64), (4, 16), (16, 32)]
    out = []
    for i in range(n):
        bw, bh = sizes[i % len(sizes)]
        blk = dict(w=bw, h=bh, compound=int(rng.integers(0, 3) == 0), bx=int(rng.integers(0, (w - bw) // 2 + 1)) * 2, by=int(rng.integers(0, (h - bh) // 2 + 1)) * 2)
        for r in range(2):
            # all 256 phase pairs appear in the first 256 blocks; motion reaches into the padding
            ph = (i + 97 * r) % 256
            blk["sx%d" % r], blk["sy%d" % r] = ph & 15, ph >> 4
            blk["mx%d" % r], blk["my%d" % r] = int(rng.integers(-70, 71)), int(rng.integers(-70, 71))
        out.append(blk)
    return out


def _clampi(v, lo, hi):
    return max(lo, min(hi, v))


def _descriptors(cases, refs, ref_addr, dst_addr, stride, dst_stride, w, h):
    arr = (api.PredBlock * len(cases))()
    for d, c in zip(arr, cases):
        for r in range(2):
            d.ref[r] = ref_addr[r] + PAD * stride + PAD
            d.ref_stride[r] = stride
            # the reference clamps motion vectors so that the 8-tap window stays inside the padded plane (eb_vp9_clamp_mv_to_umv_border_sb)
            d.x[r] = _clampi(c["bx"] + c["mx%d" % r], -PAD + 4, w + PAD - c["w"] - 5)
            d.y[r] = _clampi(c["by"] + c["my%d" % r], -PAD + 4, h + PAD - c["h"] - 5)
            d.subpel_x[r], d.subpel_y[r] = c["sx%d" % r], c["sy%d" % r]
        d.dst = dst_addr + c["by"] * dst_stride + c["bx"]
        d.dst_stride = dst_stride
        d.w, d.h, d.compound = c["w"], c["h"], c["compound"]
    return arr


def _non_overlapping(cases):
    """keep blocks whose destinations do not overlap (one batch writes each destination sample once)"""
    taken, keep = [], []
    for c in cases:
        box = (c["bx"], c["by"], c["bx"] + c["w"], c["by"] + c["h"])
        if all(box[2] <= t[0] or t[2] <= box[0] or box[3] <= t[1] or t[3] <= box[1] for t in taken):
            taken.append(box); keep.append(c)
    return keep


def _oracle_predict(cases, refs, w, h):
    dst = np.zeros((h, w), dtype=np.uint8)
    arr = _descriptors(cases, refs, [r.ctypes.data for r in refs], dst.ctypes.data, refs[0].shape[1], w, w, h)
    O = H.oracle()
    O.svtb200_oracle_inter_predict.argtypes = [C.c_void_p, C.c_int]
    O.svtb200_oracle_inter_predict(C.addressof(arr), len(arr))
    return dst, arr


def _reference_predict(cases, refs, w, h, simd):
    """the reference's functions block by block (simd 0: the plain-C ones, 1: the RTCD selection of this machine, AVX2 here).
    The SIMD copy / average kernels use aligned stores, so each block is predicted into an aligned buffer like the encoder's."""
    R = H.refdrv()
    R.refdrv_inter_predict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    dst = np.zeros((h, w), dtype=np.uint8)
    arr = _descriptors(cases, refs, [r.ctypes.data for r in refs], dst.ctypes.data, refs[0].shape[1], w, w, h)
    blk = H.aligned_zeros((64, 64), np.uint8)
    for d, c in zip(arr, cases):
        for r in range(1 + d.compound):
            assert R.refdrv_inter_predict(d.ref[r] + d.y[r] * d.ref_stride[r] + d.x[r], d.ref_stride[r], blk.ctypes.data, 64, d.subpel_x[r], d.subpel_y[r], d.w, d.h, r, simd) == 0
        dst[c["by"]:c["by"] + c["h"], c["bx"]:c["bx"] + c["w"]] = blk[:c["h"], :c["w"]]
    return dst


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("simd", [0, 1])
@pytest.mark.parametrize("kind", ["noise", "extreme", "smooth"])
def test_oracle_prediction_matches_reference_functions(kind, simd):
    rng = np.random.default_rng(11 + simd)
    w, h = 640, 384
    refs = [_ref_plane(rng, w, h, kind), _ref_plane(rng, w, h, kind)]
    cases = _non_overlapping(_cases(rng, 1200, w, h))
    assert len(cases) > 150 and len({(c["sx0"], c["sy0"]) for c in cases}) > 100
    exp = _reference_predict(cases, refs, w, h, simd)
    got, _ = _oracle_predict(cases, refs, w, h)
    assert (exp == got).all()
    assert exp.any()


def test_oracle_prediction_phase_table():
    """every one of the 16 x 16 phase pairs on one block position, against the reference when it is built, else self-consistency"""
    rng = np.random.default_rng(5)
    w, h = 64, 64
    refs = [_ref_plane(rng, w, h, "noise")] * 2
    outs = []
    for ph in range(256):
        case = dict(w=16, h=16, compound=0, bx=8, by=8, sx0=ph & 15, sy0=ph >> 4, mx0=3, my0=-2, sx1=0, sy1=0, mx1=0, my1=0)
        got, _ = _oracle_predict([case], refs, w, h)
        if H.have_refdrv():
            assert (got == _reference_predict([case], refs, w, h, 1)).all(), ph
        outs.append(got[8:24, 8:24].copy())
    assert len({o.tobytes() for o in outs}) == 256  # all phases differ on noise
    assert (outs[0] == refs[0][PAD + 8 - 2:PAD + 24 - 2, PAD + 8 + 3:PAD + 24 + 3]).all()  # phase (0,0) is a copy


def _distortion_case(rng, tx_size, n):
    nn = (4 << tx_size) ** 2
    coeff = rng.integers(-3000, 3001, size=(n, nn)).astype(np.int16)
    rc = (coeff + rng.integers(-40, 41, size=(n, nn))).astype(np.int16)
    rc[rng.integers(0, 2, size=n) == 0] = 0
    eob = rng.integers(0, 3, size=n).astype(np.uint16)
    rc[eob == 0] = 0
    return coeff, rc, eob


def _oracle_distortion(tx_size, coeff, rc, eob):
    O = H.oracle()
    O.svtb200_oracle_full_distortion.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.zeros((coeff.shape[0], 2), dtype=np.uint64)
    O.svtb200_oracle_full_distortion(tx_size, coeff.shape[0], coeff.ctypes.data, rc.ctypes.data, eob.ctypes.data, out.ctypes.data)
    return out


@pytest.mark.parametrize("tx_size", [0, 1, 2, 3])
def test_kernel_source_full_distortion_on_the_cpu(tx_size):
    """full_dist_kernel (csrc/pred_device.cuh) under the CPU emulation against the oracle: 32-bit wrapping sums, the eob == 0 rule"""
    rng = np.random.default_rng(40 + tx_size)
    nn, n = (4 << tx_size) ** 2, 37
    coeff = rng.integers(-32768, 32768, size=(n, nn), dtype=np.int16)
    rc = (coeff.astype(np.int32) + rng.integers(-3000, 3001, size=coeff.shape)).clip(-32768, 32767).astype(np.int16)
    eob = rng.integers(0, 3, size=n).astype(np.uint16)
    want = _oracle_distortion(tx_size, coeff, rc, eob)
    got = np.zeros_like(want)
    E = H.emu()
    E.full_dist_emu_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    E.full_dist_emu_run(tx_size, n, coeff.ctypes.data, rc.ctypes.data, eob.ctypes.data, got.ctypes.data)
    assert (got == want).all() and (eob == 0).any()


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("simd", [0, 1])
@pytest.mark.parametrize("tx_size", [0, 1, 2, 3])
def test_oracle_distortion_matches_reference_functions(tx_size, simd):
    rng = np.random.default_rng(tx_size)
    coeff, rc, eob = _distortion_case(rng, tx_size, 64)
    R = H.refdrv()
    R.refdrv_full_distortion.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    exp = np.zeros((64, 2), dtype=np.uint64)
    for k in range(64):
        a, b = H.aligned_copy(coeff[k]), H.aligned_copy(rc[k])
        assert R.refdrv_full_distortion(tx_size, a.ctypes.data, b.ctypes.data, int(eob[k]), exp[k].ctypes.data, simd) == 0
    assert (exp == _oracle_distortion(tx_size, coeff, rc, eob)).all()


@pytest.mark.parametrize("kind", ["noise", "extreme"])
def test_emulated_prediction_kernel_matches_oracle(kind):
    """csrc/pred_device.cuh (the CUDA kernel's source: packed dot products, 16-bit-halves vertical pass, reciprocal indexing)
    executed on the CPU; fewer CTAs than blocks, so the grid-stride loop is exercised too"""
    rng = np.random.default_rng(91)
    w, h = 320, 192
    refs = [_ref_plane(rng, w, h, kind), _ref_plane(rng, w, h, kind)]
    cases = _non_overlapping(_cases(rng, 700, w, h))
    assert len(cases) > 40 and any(c["compound"] for c in cases)
    exp, _ = _oracle_predict(cases, refs, w, h)
    got = np.zeros((h, w), dtype=np.uint8)
    arr = _descriptors(cases, refs, [r.ctypes.data for r in refs], got.ctypes.data, refs[0].shape[1], w, w, h)
    E = H.emu()
    E.pred_emu_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
    assert E.pred_emu_run(C.addressof(arr), len(arr), 7) == 0
    assert (got == exp).all()


# ---- intra prediction ------------------------------------------------------------------------------------------------
def _intra_cases(rng, n, kind):
    """n predictor calls: every mode x size x DC variant, edge samples random / saturating / smooth.
    Returns (edges, cases): edges is one byte pool, a case is (mode, tx_size, have_left, have_top, above_offset, left_offset)."""
    pool = np.zeros(n * 128, dtype=np.uint8)
    cases = []
    for i in range(n):
        mode, ts = i % 10, (i // 10) % 4
        bs = 4 << ts
        if kind == "noise":
            e = rng.integers(0, 256, size=112)
        elif kind == "extreme":
            e = rng.integers(0, 2, size=112) * 255
        else:
            e = (128 + 60 * np.sin(np.arange(112) / 9.0 + i) + rng.integers(-4, 5, size=112)).clip(0, 255)
        pool[i * 128:i * 128 + 112] = e
        # layout inside the 128-byte record: left at 0..31, above[-1] at 47, above[0..63] at 48..111
        cases.append((mode, ts, (i // 40) & 1, (i // 80) & 1, i * 128 + 48, i * 128))
        assert 2 * bs <= 64
    return pool, cases


def _intra_blocks(cases, pool_addr, dst_addr):
    arr = (api.IntraBlock * len(cases))()
    for k, (d, (mode, ts, hl, ht, ao, lo)) in enumerate(zip(arr, cases)):
        d.above, d.left, d.dst, d.dst_stride = pool_addr + ao, pool_addr + lo, dst_addr + k * 1024, 32
        d.mode, d.tx_size, d.have_left, d.have_top = mode, ts, hl, ht
    return arr


def _oracle_intra(cases, pool):
    out = np.zeros((len(cases), 32, 32), dtype=np.uint8)
    arr = _intra_blocks(cases, pool.ctypes.data, out.ctypes.data)
    O = H.oracle()
    O.svtb200_oracle_intra_predict.argtypes = [C.c_void_p, C.c_int]
    O.svtb200_oracle_intra_predict(C.addressof(arr), len(arr))
    return out


@pytest.mark.parametrize("kind", ["noise", "extreme", "smooth"])
def test_emulated_intra_closed_forms_match_oracle(kind):
    """the per-sample closed forms of csrc/intra_device.cuh (what the CUDA kernel evaluates), run on the CPU"""
    rng = np.random.default_rng(53)
    pool, cases = _intra_cases(rng, 960, kind)
    exp = _oracle_intra(cases, pool)
    got = np.zeros_like(exp)
    arr = _intra_blocks(cases, pool.ctypes.data, got.ctypes.data)
    E = H.emu()
    E.intra_emu_run.argtypes = [C.c_void_p, C.c_int]
    assert E.intra_emu_run(C.addressof(arr), len(arr)) == 0
    assert (got == exp).all()


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("simd", [0, 1])
@pytest.mark.parametrize("kind", ["noise", "extreme", "smooth"])
def test_oracle_intra_matches_reference_functions(kind, simd):
    rng = np.random.default_rng(31 + simd)
    pool, cases = _intra_cases(rng, 480, kind)
    got = _oracle_intra(cases, pool)
    R = H.refdrv()
    R.refdrv_intra_predict.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    blk = H.aligned_zeros((32, 32), np.uint8)
    edges = H.aligned_zeros(256, np.uint8)  # the SIMD predictors load the above row with aligned 16-byte loads: above[0] at a 16-byte boundary
    for k, (mode, ts, hl, ht, ao, lo) in enumerate(cases):
        bs = 4 << ts
        edges[:] = 0
        edges[64 - 1:64 + 64] = pool[ao - 1:ao + 64]
        edges[160:160 + 32] = pool[lo:lo + 32]
        blk[:] = 0
        assert R.refdrv_intra_predict(mode, ts, hl, ht, blk.ctypes.data, 32, edges.ctypes.data + 64, edges.ctypes.data + 160, simd) == 0
        assert (blk[:bs, :bs] == got[k, :bs, :bs]).all(), (mode, ts, hl, ht)
    assert len({(c[0], c[1]) for c in cases}) == 40


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["noise", "extreme", "smooth"])
def test_cuda_prediction_matches_oracle(kind):
    rng = np.random.default_rng(23)
    w, h = 1920, 1088
    refs = [_ref_plane(rng, w, h, kind), _ref_plane(rng, w, h, kind)]
    cases = _non_overlapping(_cases(rng, 6000, w, h))
    assert len(cases) > 1500
    exp, _ = _oracle_predict(cases, refs, w, h)
    ctx = api.Context(0)
    L, hd = ctx.L, ctx.h
    d_ref = [L.svtb200_dev_alloc(hd, r.nbytes) for r in refs]
    d_dst = L.svtb200_dev_alloc(hd, w * h)
    for d, r in zip(d_ref, refs):
        ctx.check(L.svtb200_memcpy_h2d(hd, d, r.ctypes.data, r.nbytes), "h2d")
    zero = np.zeros((h, w), dtype=np.uint8)
    ctx.check(L.svtb200_memcpy_h2d(hd, d_dst, zero.ctypes.data, zero.nbytes), "h2d")
    arr = _descriptors(cases, refs, d_ref, d_dst, refs[0].shape[1], w, w, h)
    before = ctx.launches()
    ctx.inter_predict(arr)
    ctx.inter_predict(arr)  # a second batch re-using the descriptor buffer while the first may still run
    ctx.sync()
    assert ctx.launches() == before + 2
    got = np.zeros((h, w), dtype=np.uint8)
    ctx.check(L.svtb200_memcpy_d2h(hd, got.ctypes.data, d_dst, got.nbytes), "d2h")
    for d in d_ref + [d_dst]:
        L.svtb200_dev_free(hd, d)
    ctx.close()
    assert (got == exp).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tx_size", [0, 1, 2, 3])
def test_cuda_distortion_matches_oracle(tx_size):
    rng = np.random.default_rng(40 + tx_size)
    coeff, rc, eob = _distortion_case(rng, tx_size, 1000)
    coeff[0, :] = 32767; rc[0, :] = -32768; eob[0] = 1  # the 16-bit difference and the 32-bit sums wrap like the C functions
    ctx = api.Context(0)
    got = ctx.full_distortion(tx_size, coeff, rc, eob)
    ctx.close()
    assert (got == _oracle_distortion(tx_size, coeff, rc, eob)).all()


@pytest.mark.gpu
def test_prediction_argument_checks():
    ctx = api.Context(0)
    arr = (api.PredBlock * 1)()
    assert ctx.L.svtb200_inter_predict_batch(ctx.h, C.addressof(arr), 1) == -2  # null planes
    assert ctx.L.svtb200_inter_predict_batch(ctx.h, None, 0) == 0
    arr[0].ref[0] = 4096; arr[0].dst = 4096; arr[0].w = 2; arr[0].h = 8
    assert ctx.L.svtb200_inter_predict_batch(ctx.h, C.addressof(arr), 1) == -2  # narrower than the reference's kernels take
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["noise", "extreme", "smooth"])
def test_cuda_intra_matches_oracle(kind):
    rng = np.random.default_rng(77)
    pool, cases = _intra_cases(rng, 1600, kind)
    exp = _oracle_intra(cases, pool)
    ctx = api.Context(0)
    L, hd = ctx.L, ctx.h
    d_pool = L.svtb200_dev_alloc(hd, pool.nbytes)
    d_out = L.svtb200_dev_alloc(hd, exp.nbytes)
    ctx.check(L.svtb200_memcpy_h2d(hd, d_pool, pool.ctypes.data, pool.nbytes), "h2d")
    got = np.zeros_like(exp)
    ctx.check(L.svtb200_memcpy_h2d(hd, d_out, got.ctypes.data, got.nbytes), "h2d")
    ctx.intra_predict(_intra_blocks(cases, d_pool, d_out))
    ctx.sync()
    ctx.check(L.svtb200_memcpy_d2h(hd, got.ctypes.data, d_out, got.nbytes), "d2h")
    L.svtb200_dev_free(hd, d_pool); L.svtb200_dev_free(hd, d_out)
    arr = (api.IntraBlock * 1)()
    arr[0].above = arr[0].left = arr[0].dst = 4096; arr[0].mode = 10
    assert L.svtb200_intra_predict_batch(hd, C.addressof(arr), 1) == -2
    ctx.close()
    assert (got == exp).all()
