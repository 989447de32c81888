"""Rest of the picture-analysis front end: the noise filters / noise detection statistics
(eb_vp9_noise_extract_*, detect_input_picture_noise: EbPictureAnalysisProcess.c:1450-1760, :3512) and the two edge maps
(edge_detection :4545, edge_detection_mean_luma_chroma16x16 :4434).  The oracle restatement (oracle/pa_noise_port.c) is pinned against the
reference's own functions -- plain C and the AVX2 entries of its dispatch tables -- through the path driver; the CUDA kernels
(csrc/pa_noise.cu) against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

api = H.api
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOISE_MIN_LEVEL_0, NOISE_MIN_LEVEL_1, FLAT_MAX_VAR = 70000, 120000, 50   # EbPictureAnalysisProcess.c:32-35
KINDS = ["luma_weak", "luma_strong", "chroma_weak", "chroma_strong"]


def _O():
    O = H.oracle()
    O.svtb200_oracle_noise_filter.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    O.svtb200_oracle_noise_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    O.svtb200_oracle_noise_detect_decimated.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    O.svtb200_oracle_edge_detection.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    O.svtb200_oracle_homogeneous_regions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    O.svtb200_oracle_edge_cu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
    return O


def _R():
    R = H.refdrv()
    R.refdrv_noise_filter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    R.refdrv_edge_detection.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    R.refdrv_edge_cu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
    return R


def _plane(rng, w, h, kind, pad=96):
    """a plane with room around it (the reference's SIMD filters load whole vectors): returns (backing array, y0, x0, stride)"""
    stride = (w + 2 * pad + 63) & ~63
    a = rng.integers(0, 256, size=(h + 2 * pad, stride), dtype=np.uint8)
    if kind == "smooth":   # gradients + mild noise: the filters' rounding cases without saturation everywhere
        yy, xx = np.mgrid[0:h + 2 * pad, 0:stride]
        a = np.clip((yy * 3 + xx * 2) // 7 % 256 + rng.integers(-6, 7, size=a.shape), 0, 255).astype(np.uint8)
    elif kind == "extremes":
        a = rng.choice(np.array([0, 255, 254, 1], dtype=np.uint8), size=a.shape)
    return a, pad, pad, stride


def oracle_noise_filter(kind, a, y0, x0, stride, w, h):
    den = np.full_like(a, 0x5a)
    noise = np.full_like(a, 0x5a) if kind == 0 else None
    _O().svtb200_oracle_noise_filter(kind, a.ctypes.data + y0 * stride + x0, stride, w, h, den.ctypes.data + y0 * stride + x0, stride,
                                     noise.ctypes.data + y0 * stride + x0 if kind == 0 else None)
    return den[y0:y0 + h, x0:x0 + w], (noise[y0:y0 + h, x0:x0 + w] if kind == 0 else None)


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("flavour", [0, 1], ids=["c", "avx2"])
@pytest.mark.parametrize("kind", range(4), ids=KINDS)
@pytest.mark.parametrize("w,h,content", [(256, 128, "noise"), (200, 136, "smooth"), (328, 72, "extremes"), (64, 64, "noise")])
def test_oracle_noise_filters_match_reference(kind, flavour, w, h, content):
    """w x h = luma size; chroma kinds filter the (w/2) x (h/2) plane.  Ragged sizes reach the partial last super-block column / row."""
    rng = np.random.default_rng(w * 7 + h + kind)
    pw, ph = (w // 2, h // 2) if kind >= 2 else (w, h)
    a, y0, x0, stride = _plane(rng, pw, ph, content)
    want_den, want_noise = oracle_noise_filter(kind, a, y0, x0, stride, pw, ph)
    den = np.full_like(a, 0x5a)
    noise = np.zeros((ph + 64, stride), np.uint8)
    _R().refdrv_noise_filter(kind, flavour, a.ctypes.data + y0 * stride + x0, stride, pw, ph, den.ctypes.data + y0 * stride + x0, stride, noise.ctypes.data)
    assert (den[y0:y0 + ph, x0:x0 + pw] == want_den).all(), "denoised plane"
    if kind == 0:
        assert (noise[:ph, :pw] == want_noise).all(), "noise plane"


def _stats(rng, w, h, bright=False):
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    y_mean = rng.integers(0, 256, size=(n_sb, 85), dtype=np.uint8)
    var = rng.integers(0, 600, size=(n_sb, 85)).astype(np.uint16)
    var[:, 5:21] = rng.integers(0, 60, size=(n_sb, 16))
    if bright:   # bright super-blocks beside dark ones: the isolated-high-intensity neighbourhoods
        y_mean[:, 0] = rng.choice(np.array([10, 100, 130, 185, 250], dtype=np.uint8), size=n_sb)
    return y_mean, var


def oracle_edge_detection(y_mean, var, w, h, avg):
    n = y_mean.shape[0]
    out = [np.full(n, 7, np.uint8) for _ in range(3)]
    _O().svtb200_oracle_edge_detection(y_mean.ctypes.data, var.ctypes.data, w, h, avg, *[o.ctypes.data for o in out])
    return out


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 2160), (640, 480), (832, 704)])
def test_oracle_edge_detection_matches_reference(w, h):
    rng = np.random.default_rng(w)
    for bright in (False, True):
        y_mean, var = _stats(rng, w, h, bright)
        avg = int(var[:, 0].astype(np.int64).sum() // var.shape[0])
        want = oracle_edge_detection(y_mean, var, w, h, avg)
        got = [np.full(y_mean.shape[0], 9, np.uint8) for _ in range(3)]
        _R().refdrv_edge_detection(y_mean.ctypes.data, var.ctypes.data, w, h, avg, *[g.ctypes.data for g in got])
        for a, b, name in zip(got, want, ("edge_block_num", "isolated_high_intensity_sb", "sharp_edge_sb_flag")):
            assert (a == b).all(), name
        if bright and w >= 832:
            assert want[1].any() and not want[1].all()
    assert any(o.any() for o in want)


def _logo_flags(w, h, rng):
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    return (rng.integers(0, 3, size=n_sb) > 0).astype(np.uint8)


def oracle_edge_cu(y_mean, cb, cr, logo, w, h, pic):
    sbs_w, sbs_h = (w + 63) // 64, (h + 63) // 64
    complete = np.array([[x * 64 + 64 <= w and y * 64 + 64 <= h for x in range(sbs_w)] for y in range(sbs_h)], dtype=np.uint8).reshape(-1)
    flags = (logo & complete).astype(np.uint8)
    out = np.full((y_mean.shape[0], 16), 5, np.uint8)
    _O().svtb200_oracle_edge_cu(y_mean.ctypes.data, cb.ctypes.data, cr.ctypes.data, flags.ctypes.data, y_mean.shape[0], pic, out.ctypes.data)
    return out, flags


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("w,h,pic", [(1920, 1080, 8), (640, 480, 0), (456, 280, 4), (1920, 1080, 6)])
def test_oracle_edge_cu_matches_reference(w, h, pic):
    rng = np.random.default_rng(h + pic)
    y_mean, _ = _stats(rng, w, h)
    y_mean[:, 5:21] = (y_mean[:, :1].astype(int) + rng.integers(-12, 13, size=(y_mean.shape[0], 16))).clip(0, 255)   # mostly gentle gradients, some edges
    cb = rng.integers(100, 140, size=(y_mean.shape[0], 21), dtype=np.uint8)
    cr = rng.integers(0, 256, size=(y_mean.shape[0], 21), dtype=np.uint8) if pic == 0 else cb.copy()
    logo = _logo_flags(w, h, rng)
    want, _ = oracle_edge_cu(y_mean, cb, cr, logo, w, h, pic)
    got = np.full_like(want, 3)
    _R().refdrv_edge_cu(y_mean.ctypes.data, cb.ctypes.data, cr.ctypes.data, logo.ctypes.data, w, h, pic, got.ctypes.data)
    assert (got == want).all()
    assert want.any() == (pic % 4 == 0)


# ---------------------------------------------------------------------------------------------- noise detection statistics
def oracle_noise_detect(a, y0, x0, stride, w, h, noise_th=NOISE_MIN_LEVEL_0, den_th=FLAT_MAX_VAR):
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    den = np.full_like(a, 0x5a)
    nv, dv, flat, tot = np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint8), np.zeros(2, np.uint64)
    _O().svtb200_oracle_noise_detect(a.ctypes.data + y0 * stride + x0, stride, w, h, den.ctypes.data + y0 * stride + x0, stride, noise_th, den_th,
                                     nv.ctypes.data, dv.ctypes.data, flat.ctypes.data, tot.ctypes.data)
    return den[y0:y0 + h, x0:x0 + w], nv, dv, flat, tot


def _noisy_picture(rng, w, h):
    """flat areas with noise of different strength per super-block (flat-noise flags of both kinds), one textured corner"""
    a, y0, x0, stride = _plane(rng, w, h, "smooth")
    base = np.full((h, w), 120, np.int64)
    for sy in range(0, h, 64):
        for sx in range(0, w, 64):
            amp = int(rng.integers(0, 40))
            base[sy:sy + 64, sx:sx + 64] += rng.integers(-amp, amp + 1, size=base[sy:sy + 64, sx:sx + 64].shape)
    base[:64, :64] = rng.integers(0, 256, size=(64, 64))
    a[y0:y0 + h, x0:x0 + w] = np.clip(base, 0, 255)
    return a, y0, x0, stride


def test_oracle_noise_detect_is_consistent_with_the_pinned_restatements():
    """the denoised picture is the weak filter's (pinned above against the reference); the 64x64 variance of the denoised samples is the
    64x64 point of the block statistics' hierarchy (compute_variance64x64 'is exactly the same' code, EbPictureAnalysisProcess.c:401-407),
    whose restatement the seam check pins inside real encodes"""
    rng = np.random.default_rng(11)
    w, h = 320, 200
    a, y0, x0, stride = _noisy_picture(rng, w, h)
    den, nv, dv, flat, tot = oracle_noise_detect(a, y0, x0, stride, w, h)
    want_den, want_noise = oracle_noise_filter(0, a, y0, x0, stride, w, h)
    assert (den == want_den).all()
    sbs_w, sbs_h = (w + 63) // 64, (h + 63) // 64
    O = H.oracle()
    for which, plane in (("den", want_den), ("noise", want_noise)):
        full = np.zeros((sbs_h * 64, sbs_w * 64), np.uint8)
        full[:h, :w] = plane
        st = api.plane_of(full, 0, 0, sbs_w * 64, sbs_h * 64)
        ym, var = np.zeros((sbs_w * sbs_h, 85), np.uint8), np.zeros((sbs_w * sbs_h, 85), np.uint16)
        O.svtb200_oracle_block_stats_picture(C.byref(st), sbs_w * 64, sbs_h * 64, ym.ctypes.data, var.ctypes.data)
        for sb in range(sbs_w * sbs_h):
            complete = (sb % sbs_w) * 64 + 64 <= w and (sb // sbs_w) * 64 + 64 <= h
            got = int(dv[sb]) if which == "den" else int(nv[sb]) >> 16
            assert got == (int(var[sb, 0]) if complete else 0), (which, sb)
    assert tot[1] == (w // 64) * (h // 64) and tot[0] == sum(int(v) >> 16 for v in nv)
    assert flat.any() and not flat.all()


# ---------------------------------------------------------------------------------------------- kernel sources on the CPU
def _E():
    E = H.emu()
    E.pa_emu_noise_filter.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    E.pa_emu_noise_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    E.pa_emu_edge_detection.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    E.pa_emu_homogeneous.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    E.pa_emu_edge_cu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return E


@pytest.mark.parametrize("kind", range(4), ids=KINDS)
def test_kernel_source_noise_filter_on_the_cpu(kind):
    rng = np.random.default_rng(kind)
    w, h = 203, 70
    a, y0, x0, stride = _plane(rng, w, h, "noise", pad=8)
    want_den, want_noise = oracle_noise_filter(kind, a, y0, x0, stride, w, h)
    den, noise = np.full_like(a, 0x33), np.full_like(a, 0x33)
    _E().pa_emu_noise_filter(kind, a.ctypes.data + y0 * stride + x0, stride, w, h, den.ctypes.data + y0 * stride + x0, stride,
                             noise.ctypes.data + y0 * stride + x0 if kind == 0 else None)
    assert (den[y0:y0 + h, x0:x0 + w] == want_den).all()
    den[y0:y0 + h, x0:x0 + w] = 0x33
    assert (den == 0x33).all()   # nothing written outside the plane
    if kind == 0:
        assert (noise[y0:y0 + h, x0:x0 + w] == want_noise).all()


def test_kernel_source_noise_detect_on_the_cpu():
    rng = np.random.default_rng(5)
    w, h = 200, 136
    a, y0, x0, stride = _noisy_picture(rng, w, h)
    want = oracle_noise_detect(a, y0, x0, stride, w, h)
    n_sb = want[1].size
    den = np.full_like(a, 0x33)
    nv, dv, flat, tot = np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint8), np.zeros(2, np.uint64)
    _E().pa_emu_noise_detect(a.ctypes.data + y0 * stride + x0, stride, w, h, 64, (w + 63) // 64, n_sb, den.ctypes.data + y0 * stride + x0, stride, NOISE_MIN_LEVEL_0,
                             FLAT_MAX_VAR, nv.ctypes.data, dv.ctypes.data, flat.ctypes.data, tot.ctypes.data)
    for g, wv, name in zip((den[y0:y0 + h, x0:x0 + w], nv, dv, flat, tot), want, ("denoised", "noise_var", "den_var", "flat", "totals")):
        assert (g == wv).all(), name


# ---- the decimated detection methods (quarter_sample_detect_noise: block 32 on the 1/4 plane, sub_sample_detect_noise: block 16 on the 1/16 plane).
# The oracle restatement is pinned inside real encodes of the unmodified reference (oracle/ref_seam_check.c: sb_flat_noise_array and pic_noise_class
# of every comparable picture, both methods, flags set: test_oracle_noise_detection_pinned_inside_real_encodes below).
def _decimated_case(rng, luma_w, luma_h, block):
    step = 64 // block                       # 2: quarter plane, 4: sixteenth plane
    w, h = luma_w // step, luma_h // step
    a, y0, x0, stride = _noisy_picture(rng, w, h)
    pw, n_sb = (luma_w + 63) // 64, ((luma_w + 63) // 64) * ((luma_h + 63) // 64)
    return a, y0, x0, stride, w, h, pw, n_sb


def oracle_noise_detect_decimated(a, y0, x0, stride, w, h, block, pw, n_sb, noise_th, den_th=FLAT_MAX_VAR):
    den = np.full_like(a, 0x5a)
    nv, dv, flat, tot = np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint8), np.zeros(2, np.uint64)
    _O().svtb200_oracle_noise_detect_decimated(a.ctypes.data + y0 * stride + x0, stride, w, h, block, pw, n_sb, den.ctypes.data + y0 * stride + x0, stride, noise_th, den_th,
                                               nv.ctypes.data, dv.ctypes.data, flat.ctypes.data, tot.ctypes.data)
    return nv, dv, flat, tot


@pytest.mark.parametrize("luma_w,luma_h,block", [(704, 456, 32), (1088, 840, 16)])
def test_kernel_source_decimated_noise_detect_on_the_cpu(luma_w, luma_h, block):
    rng = np.random.default_rng(block)
    a, y0, x0, stride, w, h, pw, n_sb = _decimated_case(rng, luma_w, luma_h, block)
    want = oracle_noise_detect_decimated(a, y0, x0, stride, w, h, block, pw, n_sb, NOISE_MIN_LEVEL_0)
    den = np.full_like(a, 0x33)
    got = [np.full(n_sb, 7, np.uint64), np.full(n_sb, 7, np.uint64), np.full(n_sb, 7, np.uint8), np.zeros(2, np.uint64)]
    _E().pa_emu_noise_detect(a.ctypes.data + y0 * stride + x0, stride, w, h, block, pw, n_sb, den.ctypes.data + y0 * stride + x0, stride, NOISE_MIN_LEVEL_0, FLAT_MAX_VAR,
                             *[g.ctypes.data for g in got])
    for g, wv, name in zip(got, want, ("noise_var", "den_var", "flat", "totals")):
        assert (g == wv).all(), name
    visited = (w // 64) * (h // 64) * (64 // block) ** 2
    assert want[3][1] == visited and visited < (w // block) * (h // block)      # areas right of / below the last whole 64x64 area are not walked
    assert want[2].any() and (want[0][want[1] > 0] > 0).any()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "SvtVp9EncApp_seam")), reason="oracle/_ref seam-check flavour not built")
@pytest.mark.parametrize("args,method", [("960 544 8 0 6 8 20 1", 2), ("1920 1080 8 1 5 6 16 1", 1), ("3840 2160 8 1 3 5 12 1", 0)], ids=["full", "quarter", "half"])
def test_oracle_noise_detection_pinned_inside_real_encodes(args, method):
    """the unmodified reference encodes a smooth, mildly noisy clip; the seam check recomputes the noise detection of every picture the denoiser left
    alone with the oracle (the method the reference picked: full sample at 540p, quarter at 1080p, half = 1/16 plane at 4K) and compares the
    flat-noise flags of all super-blocks and the picture's noise class; the edge maps are compared on every picture as well"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_seam_check.py")] + args.split(), stdout=subprocess.PIPE, timeout=900)
    out = dict(l.split() for l in r.stdout.decode().strip().splitlines() if len(l.split()) == 2)
    assert int(out["noise_method"]) == method, out
    assert int(out["noise_pics"]) >= 2 and int(out["noise_flat_set"]) > 0, out       # pictures were comparable and flags were set
    for k in ("noise_bad", "edge_bad", "edge_cu_bad", "homog_bad", "sb_bad", "stats_bad"):
        assert int(out[k]) == 0, out


@pytest.mark.parametrize("w,h", [(832, 704), (640, 480)])
def test_kernel_source_edge_maps_on_the_cpu(w, h):
    rng = np.random.default_rng(w)
    y_mean, var = _stats(rng, w, h, bright=True)
    avg = int(var[:, 0].astype(np.int64).sum() // var.shape[0])
    want = oracle_edge_detection(y_mean, var, w, h, avg)
    got = [np.full(y_mean.shape[0], 9, np.uint8) for _ in range(3)]
    got_avg = np.zeros(1, np.uint16)
    _E().pa_emu_edge_detection(y_mean.ctypes.data, var.ctypes.data, w, h, got[0].ctypes.data, got[1].ctypes.data, got[2].ctypes.data, got_avg.ctypes.data)
    assert got_avg[0] == avg
    for a, b in zip(got, want):
        assert (a == b).all()
    cb = rng.integers(100, 140, size=(y_mean.shape[0], 21), dtype=np.uint8)
    cr = rng.integers(0, 256, size=(y_mean.shape[0], 21), dtype=np.uint8)
    logo = _logo_flags(w, h, rng)
    want_cu, _ = oracle_edge_cu(y_mean, cb, cr, logo, w, h, 4)
    got_cu, mg = np.full_like(want_cu, 3), np.zeros(1, np.uint32)
    _E().pa_emu_edge_cu(y_mean.ctypes.data, cb.ctypes.data, cr.ctypes.data, logo.ctypes.data, w, h, got_cu.ctypes.data, mg.ctypes.data)
    assert (got_cu == want_cu).all() and want_cu.any()


def _variances(rng, w, h):
    """block statistics with flat and busy super-blocks (the homogeneous-area flag both ways), some 64x64 variances below SB_LOW_VAR_TH"""
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    var = rng.integers(0, 40, size=(n_sb, 85)).astype(np.uint16)
    busy = rng.integers(0, 2, size=n_sb).astype(bool)
    var[busy, 21:] = rng.integers(0, 1500, size=(int(busy.sum()), 64))
    var[:, 0] = rng.integers(0, 12, size=n_sb)
    return var


def oracle_homogeneous(var, w, h):
    n_sb = var.shape[0]
    vv, hom, cnt = np.zeros((n_sb, 4), np.uint64), np.full(n_sb, 7, np.uint8), np.zeros(3, np.uint64)
    _O().svtb200_oracle_homogeneous_regions(var.ctypes.data, w, h, 5, 64 * 64, vv.ctypes.data, hom.ctypes.data, cnt.ctypes.data)
    return vv, hom, cnt


@pytest.mark.parametrize("w,h", [(832, 704), (456, 280)])
def test_kernel_source_homogeneous_regions_on_the_cpu(w, h):
    """(the oracle itself is pinned inside real encodes: homog_bad 0 in test_oracle_noise_detection_pinned_inside_real_encodes)"""
    rng = np.random.default_rng(h)
    var = _variances(rng, w, h)
    want = oracle_homogeneous(var, w, h)
    got = [np.zeros_like(want[0]), np.full_like(want[1], 9), np.zeros(3, np.uint64)]
    _E().pa_emu_homogeneous(var.ctypes.data, w, h, 5, 64 * 64, *[g.ctypes.data for g in got])
    for g, wv in zip(got, want):
        assert (g == wv).all()
    assert want[1].any() and not want[1].all() and want[2][1] > 0


# ---------------------------------------------------------------------------------------------- CUDA
class _Dev:
    def __init__(self):
        self.ctx = api.Context(0)
        self.bufs = []

    def up(self, a):
        a = np.ascontiguousarray(a)
        d = self.ctx.L.svtb200_dev_alloc(self.ctx.h, max(a.nbytes, 16))
        self.ctx.check(self.ctx.L.svtb200_memcpy_h2d(self.ctx.h, d, a.ctypes.data, a.nbytes), "h2d")
        self.bufs.append(d)
        return d

    def down(self, d, like):
        out = np.zeros_like(like)
        self.ctx.check(self.ctx.L.svtb200_memcpy_d2h(self.ctx.h, out.ctypes.data, d, out.nbytes), "d2h")
        return out

    def close(self):
        for d in self.bufs:
            self.ctx.L.svtb200_dev_free(self.ctx.h, d)
        self.ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(200, 136), (1920, 1080), (3840, 2160)])
def test_cuda_noise_filters_and_detection(w, h):
    rng = np.random.default_rng(h)
    D = _Dev()
    L, hd = D.ctx.L, D.ctx.h
    try:
        for kind in range(4):
            pw, ph = (w // 2, h // 2) if kind >= 2 else (w, h)
            a, y0, x0, stride = _plane(rng, pw, ph, "noise" if kind & 1 else "smooth", pad=16)
            want_den, want_noise = oracle_noise_filter(kind, a, y0, x0, stride, pw, ph)
            d_in, d_den, d_nz = D.up(a), D.up(np.full_like(a, 0x33)), D.up(np.full_like(a, 0x33))
            off = y0 * stride + x0
            D.ctx.check(L.svtb200_noise_filter_dev(hd, kind, d_in + off, stride, pw, ph, d_den + off, stride, d_nz + off if kind == 0 else None), "noise_filter")
            D.ctx.sync()
            den = D.down(d_den, a)
            assert (den[y0:y0 + ph, x0:x0 + pw] == want_den).all(), KINDS[kind]
            den[y0:y0 + ph, x0:x0 + pw] = 0x33
            assert (den == 0x33).all()
            if kind == 0:
                assert (D.down(d_nz, a)[y0:y0 + ph, x0:x0 + pw] == want_noise).all()
        a, y0, x0, stride = _noisy_picture(rng, w, h)
        want = oracle_noise_detect(a, y0, x0, stride, w, h, NOISE_MIN_LEVEL_1, FLAT_MAX_VAR)
        n_sb = want[1].size
        d_in, d_den = D.up(a), D.up(np.full_like(a, 0x33))
        outs = [np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint64), np.zeros(n_sb, np.uint8), np.zeros(2, np.uint64)]
        d_out = [D.up(o) for o in outs]
        off = y0 * stride + x0
        D.ctx.check(L.svtb200_noise_detect_dev(hd, d_in + off, stride, w, h, d_den + off, stride, NOISE_MIN_LEVEL_1, FLAT_MAX_VAR, *d_out), "noise_detect")
        D.ctx.sync()
        assert (D.down(d_den, a)[y0:y0 + h, x0:x0 + w] == want[0]).all()
        for d, o, wv, name in zip(d_out, outs, want[1:], ("noise_var", "den_var", "flat", "totals")):
            assert (D.down(d, o) == wv).all(), name
        assert want[3].any() or w < 1000   # (a handful of super-blocks need not hold a flat noisy one)
    finally:
        D.close()


@pytest.mark.gpu
@pytest.mark.parametrize("luma_w,luma_h,block", [(1920, 1080, 32), (3840, 2160, 16), (704, 456, 32)], ids=["1080p-quarter", "4k-sixteenth", "704x456-quarter"])
def test_cuda_decimated_noise_detection(luma_w, luma_h, block):
    rng = np.random.default_rng(luma_w + block)
    a, y0, x0, stride, w, h, pw, n_sb = _decimated_case(rng, luma_w, luma_h, block)
    want = oracle_noise_detect_decimated(a, y0, x0, stride, w, h, block, pw, n_sb, NOISE_MIN_LEVEL_1)
    D = _Dev()
    L, hd = D.ctx.L, D.ctx.h
    L.svtb200_noise_detect_decimated_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                                     C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    try:
        d_in, d_den = D.up(a), D.up(np.full_like(a, 0x33))
        outs = [np.full(n_sb, 7, np.uint64), np.full(n_sb, 7, np.uint64), np.full(n_sb, 7, np.uint8), np.zeros(2, np.uint64)]
        d_out = [D.up(o) for o in outs]
        off = y0 * stride + x0
        D.ctx.check(L.svtb200_noise_detect_decimated_dev(hd, d_in + off, stride, w, h, block, pw, n_sb, d_den + off, stride, NOISE_MIN_LEVEL_1, FLAT_MAX_VAR, *d_out),
                    "noise_detect_decimated")
        D.ctx.sync()
        for d, o, wv, name in zip(d_out, outs, want, ("noise_var", "den_var", "flat", "totals")):
            assert (D.down(d, o) == wv).all(), name
        assert want[2].any()
    finally:
        D.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,pic", [(3840, 2160, 8), (1920, 1080, 0), (640, 480, 4), (832, 704, 2)])
def test_cuda_edge_maps(w, h, pic):
    rng = np.random.default_rng(w + pic)
    y_mean, var = _stats(rng, w, h, bright=True)
    n_sb = y_mean.shape[0]
    avg = int(var[:, 0].astype(np.int64).sum() // n_sb)
    want = oracle_edge_detection(y_mean, var, w, h, avg)
    cb = rng.integers(100, 140, size=(n_sb, 21), dtype=np.uint8)
    cr = rng.integers(0, 256, size=(n_sb, 21), dtype=np.uint8)
    logo = _logo_flags(w, h, rng)
    want_cu, _ = oracle_edge_cu(y_mean, cb, cr, logo, w, h, pic)
    D = _Dev()
    L, hd = D.ctx.L, D.ctx.h
    try:
        d_ym, d_var = D.up(y_mean), D.up(var)
        outs = [np.full(n_sb, 9, np.uint8) for _ in range(3)] + [np.zeros(1, np.uint16)]
        d_out = [D.up(o) for o in outs]
        D.ctx.check(L.svtb200_edge_detection_dev(hd, d_ym, d_var, w, h, *d_out), "edge_detection")
        d_cb, d_cr, d_logo, d_cu, d_mg = D.up(cb), D.up(cr), D.up(logo), D.up(np.full_like(want_cu, 3)), D.up(np.zeros(4, np.uint32))
        D.ctx.check(L.svtb200_edge_cu_dev(hd, d_ym, d_cb, d_cr, d_logo, w, h, pic, d_cu, d_mg), "edge_cu")
        D.ctx.sync()
        for d, o, wv, name in zip(d_out, outs, want + [np.array([avg], np.uint16)], ("edge_block_num", "isolated", "sharp_edge", "pic_avg_variance")):
            assert (D.down(d, o) == wv).all(), name
        assert (D.down(d_cu, want_cu) == want_cu).all()
        assert want_cu.any() == (pic % 4 == 0)
        # determine_homogeneous_region_in_picture on a second set of statistics
        L.svtb200_homogeneous_regions_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        var2 = _variances(rng, w, h)
        want_h = oracle_homogeneous(var2, w, h)
        outs_h = [np.zeros_like(want_h[0]), np.full_like(want_h[1], 9), np.zeros(3, np.uint64)]
        d_h = [D.up(o) for o in outs_h]
        D.ctx.check(L.svtb200_homogeneous_regions_dev(hd, D.up(var2), w, h, 5, 64 * 64, *d_h), "homogeneous_regions")
        D.ctx.sync()
        for d, o, wv in zip(d_h, outs_h, want_h):
            assert (D.down(d, o) == wv).all()
    finally:
        D.close()
