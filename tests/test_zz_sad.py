"""compute_zz_sad: the oracle restatement against the reference's own function, the CUDA path against both."""
import ctypes as C
import importlib

import numpy as np
import pytest

import helpers as H

api = H.api
frames = importlib.import_module("svt-vp9_b200.frames")
TH_SHIFT = [4, 2, 0, 0]  # non_moving_th_shift[input_resolution], EbMotionEstimationProcess.c:353

CASES = [(640, 384, 0, 0), (640, 384, 2, 1), (456, 248, 1, 2), (1920, 1088, 3, 1), (200, 136, 2, 8)]


def _pair(w, h, noise, step):
    """two pictures: a still background with a moving patch, so that all four score classes appear"""
    rng = np.random.default_rng(w * 7 + h + noise)
    base = H.synth.make_clip(w, h, 1, seed=w + h, k=3, g=4, a=0, b=0, m=32, noise=0)[0][0].astype(np.int16)
    prev = base.copy()
    cur = base.copy()
    cur[:, w // 3:] += rng.integers(-noise, noise + 1, size=(h, w - w // 3), dtype=np.int16) if noise else 0
    cur[h // 2:, : w // 2] = np.roll(base, step, axis=1)[h // 2:, : w // 2]
    to8 = lambda a: np.clip(a, 0, 255).astype(np.uint8)
    pl = [frames.pa_planes(to8(prev)), frames.pa_planes(to8(cur))]
    return pl, [frames.plane_structs(p) for p in pl]


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("w,h,res,noise", CASES)
def test_oracle_zz_sad_matches_reference_function(w, h, res, noise):
    pl, st = _pair(w, h, noise, 3)
    ref = H.refdrv_zz_sad(st[1], st[0], w, h, res)
    got, sads = H.oracle_zz_sad(st[1], st[0], w, h, TH_SHIFT[res])
    assert (ref == got).all()
    if w % 64 or h % 64:
        assert (sads == 0xffffffff).any() and (got[sads == 0xffffffff] == 30).all()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,res,noise", CASES + [(3840, 2160, 3, 2)])
def test_cuda_zz_sad_matches_oracle_and_reference(w, h, res, noise):
    pl, st = _pair(w, h, noise, 5)
    exp, exp_sad = H.oracle_zz_sad(st[1], st[0], w, h, TH_SHIFT[res])
    ctx = api.Context(0)
    ctx.picture_upload(0, st[0][0], st[0][1], st[0][2])
    ctx.picture_upload_luma(1, st[1][0])          # the current picture's 1/16 plane is produced on the device
    got, sad = ctx.zz_sad_picture(1, 0, w, h, TH_SHIFT[res])
    ctx.close()
    assert (got == exp).all() and (sad == exp_sad).all()
    assert len(set(exp.tolist())) >= 2
    if H.have_refdrv():
        assert (got == H.refdrv_zz_sad(st[1], st[0], w, h, res)).all()


def _oracle_stats(st, w, h):
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    ym, var = np.zeros((n_sb, 85), dtype=np.uint8), np.zeros((n_sb, 85), dtype=np.uint16)
    L = H.oracle()
    L.svtb200_oracle_block_stats_picture.argtypes = [C.POINTER(api.Plane), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.svtb200_oracle_block_stats_picture(C.byref(st[0]), w, h, ym.ctypes.data, var.ctypes.data)
    return ym, var


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
def test_oracle_block_stats_8x8_level_matches_reference_intrinsics():
    """the 8x8 entries (21..84) against the reference's own eb_vp9_compute_sub_mean8x8_sse2_intrin /
    eb_vp9_compute_subd_mean_of_squared_values8x8_sse2_intrin; the whole function is pinned inside real encodes by the seam check
    (oracle/ref_seam_check.c: stats_bad 0)"""
    import os
    R = C.CDLL(os.path.join(H.ROOT, "oracle", "_ref", "libSvtVp9Enc_ref.so"))
    for f in (R.eb_vp9_compute_sub_mean8x8_sse2_intrin, R.eb_vp9_compute_subd_mean_of_squared_values8x8_sse2_intrin):
        f.argtypes, f.restype = [C.c_void_p, C.c_uint16], C.c_uint64
    w, h = 200, 136
    pl, st = _pair(w, h, 6, 3)
    ym, var = _oracle_stats(st[1], w, h)
    full = pl[1]["full"]
    sbs_w = (w + 63) // 64
    for sb in range(ym.shape[0]):
        ox, oy = (sb % sbs_w) * 64, (sb // sbs_w) * 64
        for b in range(64):
            p = full[frames.FULL_PAD + oy + (b >> 3) * 8:, frames.FULL_PAD + ox + (b & 7) * 8:]
            addr = full.ctypes.data + (frames.FULL_PAD + oy + (b >> 3) * 8) * full.shape[1] + frames.FULL_PAD + ox + (b & 7) * 8
            m = R.eb_vp9_compute_sub_mean8x8_sse2_intrin(addr, full.shape[1])
            q = R.eb_vp9_compute_subd_mean_of_squared_values8x8_sse2_intrin(addr, full.shape[1])
            assert ym[sb, 21 + b] == (m >> 8) & 0xff and var[sb, 21 + b] == ((q - m * m) >> 16) & 0xffff


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(640, 384), (200, 136), (1920, 1088), (3840, 2160)])
def test_cuda_block_stats_match_oracle(w, h):
    pl, st = _pair(w, h, 5, 2)
    ctx = api.Context(0)
    ctx.picture_upload(0, st[1][0], st[1][1], st[1][2])
    ym, var = ctx.block_stats_picture(0, w, h)
    ctx.close()
    eym, evar = _oracle_stats(st[1], w, h)
    assert (ym == eym).all() and (var == evar).all()
    assert var.max() > 0


def _oracle_hist(st, rw, rh):
    hist = np.zeros((rw, rh, 256), dtype=np.uint32)
    avg = np.zeros((rw, rh), dtype=np.uint8)
    s16 = C.c_uint64()
    L = H.oracle()
    L.svtb200_oracle_luma_histogram.argtypes = [C.POINTER(api.Plane), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.svtb200_oracle_luma_histogram(C.byref(st[2]), rw, rh, hist.ctypes.data, avg.ctypes.data, C.byref(s16))
    return hist, avg, s16.value


def test_oracle_histogram_properties():
    """the restatement is pinned inside real encodes by the seam check (hist_bad 0 over 31 pictures); here its invariants"""
    w, h = 456, 264
    pl, st = _pair(w, h, 4, 2)
    hist, avg, s16 = _oracle_hist(st[1], 4, 4)
    six = pl[1]["sixteenth"][frames.SIXTEENTH_PAD:frames.SIXTEENTH_PAD + (h + 3) // 4, frames.SIXTEENTH_PAD:frames.SIXTEENTH_PAD + (w + 3) // 4]
    assert int((hist >> 4).sum()) - 16 * 256 == six.size and (hist & 15).sum() == 0
    assert s16 == int(six.astype(np.int64).sum()) << 4
    assert abs(int(avg[0, 0]) - int(round(six[: six.shape[0] // 4, : six.shape[1] // 4].mean()))) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,rw,rh", [(640, 384, 4, 4), (456, 264, 4, 4), (3840, 2160, 4, 4), (200, 136, 2, 3)])
def test_cuda_histogram_matches_oracle(w, h, rw, rh):
    pl, st = _pair(w, h, 6, 3)
    ctx = api.Context(0)
    ctx.picture_upload_luma(0, st[1][0])
    hist, avg, s16 = ctx.luma_histogram_picture(0, rw, rh)
    ctx.close()
    eh, ea, es = _oracle_hist(st[1], rw, rh)
    assert (hist == eh).all() and (avg == ea).all() and s16 == es


# ---------------------------------------------------------------------------------------------- kernel sources on the CPU
@pytest.mark.parametrize("w,h,res,noise", [(640, 384, 2, 1), (456, 248, 1, 2), (200, 136, 2, 8)])
def test_kernel_sources_zz_sad_block_stats_histogram_on_the_cpu(w, h, res, noise):
    """zz_sad_kernel, block_stats_kernel and the two histogram kernels (csrc/zz_kernel.cuh) under the CPU emulation, against the oracle"""
    pl, st = _pair(w, h, noise, 5)
    E = H.emu()
    E.pa_emu_zz_sad.argtypes = [C.POINTER(api.Plane), C.POINTER(api.Plane), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    E.pa_emu_block_stats.argtypes = [C.POINTER(api.Plane), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    E.pa_emu_luma_histogram.argtypes = [C.POINTER(api.Plane), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    want, want_sad = H.oracle_zz_sad(st[1], st[0], w, h, TH_SHIFT[res])
    got, sad = np.full_like(want, 77), np.zeros_like(want_sad)
    E.pa_emu_zz_sad(C.byref(st[1][2]), C.byref(st[0][0]), w, h, TH_SHIFT[res], got.ctypes.data, sad.ctypes.data)
    assert (got == want).all() and (sad == want_sad).all()
    eym, evar = _oracle_stats(st[1], w, h)
    # (the statistics kernel reads whole super-blocks: the padded plane holds the samples right of / below a ragged picture)
    ym, var = np.zeros_like(eym), np.zeros_like(evar)
    E.pa_emu_block_stats(C.byref(st[1][0]), w, h, ym.ctypes.data, var.ctypes.data)
    assert (ym == eym).all() and (var == evar).all() and var.max() > 0
    for rw, rh in ((4, 4), (2, 3)):
        eh, ea, es = _oracle_hist(st[1], rw, rh)
        hist, avg, s16 = np.zeros_like(eh), np.zeros_like(ea), C.c_uint64()
        E.pa_emu_luma_histogram(C.byref(st[1][2]), rw, rh, hist.ctypes.data, avg.ctypes.data, C.byref(s16))
        assert (hist == eh).all() and (avg == ea).all() and s16.value == es
