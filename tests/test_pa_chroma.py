"""Chroma statistics of the picture-analysis stage (compute_chroma_block_mean, EbPictureAnalysisProcess.c:1828; chroma histograms :4313):
the oracle restatement (oracle/pa_port.c) is pinned inside real encodes of the unmodified reference by the seam check
(oracle/ref_seam_check.c, chroma_bad / chroma_hist_bad), the CUDA kernels (csrc/pa_chroma.cu) against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

api = H.api
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle(cb, cr, w, h, rw, rh):
    O = H.oracle()
    O.svtb200_oracle_chroma_block_mean.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    O.svtb200_oracle_chroma_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    mb, mr = np.zeros((n_sb, 21), np.uint8), np.zeros((n_sb, 21), np.uint8)
    O.svtb200_oracle_chroma_block_mean(cb.ctypes.data, cr.ctypes.data, cb.shape[1], w, h, mb.ctypes.data, mr.ctypes.data)
    hist, avg, s16 = np.zeros((rw, rh, 2, 256), np.uint32), np.zeros((rw, rh, 2), np.uint8), np.zeros(2, np.uint64)
    O.svtb200_oracle_chroma_histogram(cb.ctypes.data, cr.ctypes.data, cb.shape[1], w, h, rw, rh, hist.ctypes.data, avg.ctypes.data, s16.ctypes.data)
    return mb, mr, hist, avg, s16


def test_oracle_invariants():
    """flat planes: every mean is the sample value (the 64x64 formula included); histogram counts add up to the sampled grid"""
    w, h = 256, 192
    cb, cr = np.full((h // 2, w // 2 + 8), 77, np.uint8), np.full((h // 2, w // 2 + 8), 200, np.uint8)
    mb, mr, hist, avg, s16 = _oracle(cb, cr, w, h, 4, 4)
    assert (mb == 77).all() and (mr == 200).all() and (avg[..., 0] == 77).all() and (avg[..., 1] == 200).all()
    assert ((hist >> 4).sum(axis=-1) == 256 + (w // 4 // 2 // 4) * (h // 4 // 2 // 4)).all()
    # the 64x64 entry follows the reference's formula: quadrant 3 twice, quadrant 2 never
    rng = np.random.default_rng(1)
    cb = rng.integers(0, 256, size=(h // 2, w // 2 + 8), dtype=np.uint8)
    mb, _, _, _, _ = _oracle(cb, cr, w, h, 4, 4)
    s = cb[0:32:2, 0:32].astype(np.int64).reshape(4, 4, 4, 8).sum(axis=(1, 3)) << 3     # [block row][block col] sums of rows 0,2,4,6
    m32 = [(s[r:r + 2, c:c + 2].sum()) >> 2 for r in (0, 2) for c in (0, 2)]
    assert mb[0][0] == ((m32[0] + m32[1] + m32[3] + m32[3]) >> 2) >> 8


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "SvtVp9EncApp_seam")), reason="oracle/_ref seam-check flavour not built")
def test_oracle_pinned_inside_a_real_encode():
    """the unmodified reference encodes a clip; the seam check compares its cb_mean / cr_mean / chroma histograms of every picture
    with the oracle restatement (and, as before, motion estimation, block statistics and luma histograms)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_seam_check.py"), "320", "192", "8", "1", "10"], stdout=subprocess.PIPE, timeout=600)
    out = dict(l.split() for l in r.stdout.decode().strip().splitlines() if len(l.split()) == 2)
    assert int(out["hist_pics"]) >= 8
    for k in ("chroma_bad", "chroma_hist_bad", "hist_bad", "stats_bad", "sb_bad"):
        assert int(out[k]) == 0, out


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(640, 384), (456, 280), (3840, 2160)])
def test_cuda_chroma_statistics(size):
    w, h = size
    rng = np.random.default_rng(w)
    stride = ((w // 2 + 7) & ~7) + 16
    cb = rng.integers(0, 256, size=((h + 63) // 64 * 32 + 8, stride), dtype=np.uint8)
    cr = (rng.integers(0, 64, size=cb.shape) + 100).astype(np.uint8)
    rw = rh = 4
    mb, mr, hist, avg, s16 = _oracle(cb, cr, w, h, rw, rh)
    ctx = api.Context(0)
    L, hd = ctx.L, ctx.h
    dev = []

    def up(a):
        d = L.svtb200_dev_alloc(hd, a.nbytes)
        ctx.check(L.svtb200_memcpy_h2d(hd, d, a.ctypes.data, a.nbytes), "h2d")
        dev.append(d)
        return d
    try:
        d_cb, d_cr = up(cb), up(cr)
        outs = [np.zeros_like(x) for x in (mb, mr, hist, avg, s16)]
        d_out = [up(o) for o in outs]
        ctx.check(L.svtb200_chroma_block_mean_dev(hd, d_cb, d_cr, stride, w, h, d_out[0], d_out[1]), "chroma_block_mean")
        ctx.check(L.svtb200_chroma_histogram_dev(hd, d_cb, d_cr, stride, w, h, rw, rh, d_out[2], d_out[3], d_out[4]), "chroma_histogram")
        ctx.sync()
        for o, d, want in zip(outs, d_out, (mb, mr, hist, avg, s16)):
            ctx.check(L.svtb200_memcpy_d2h(hd, o.ctypes.data, d, o.nbytes), "d2h")
            assert (o == want).all()
        assert mb.any() and (mb[-1] == 0).all() == (h % 64 != 0 or w % 64 != 0)
    finally:
        for d in dev:
            L.svtb200_dev_free(hd, d)
        ctx.close()


@pytest.mark.parametrize("size", [(640, 384), (456, 280)])
def test_kernel_sources_chroma_statistics_on_the_cpu(size):
    """chroma_mean_kernel and the two chroma histogram kernels (csrc/pa_chroma_kernel.cuh) under the CPU emulation, against the oracle"""
    w, h = size
    rng = np.random.default_rng(w + 1)
    stride = ((w // 2 + 7) & ~7) + 16
    cb = H.aligned_copy(rng.integers(0, 256, size=((h + 63) // 64 * 32 + 8, stride), dtype=np.uint8))
    cr = H.aligned_copy((rng.integers(0, 64, size=cb.shape) + 100).astype(np.uint8))
    rw = rh = 4
    mb, mr, hist, avg, s16 = _oracle(cb, cr, w, h, rw, rh)
    E = H.emu()
    E.pa_emu_chroma_mean.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    E.pa_emu_chroma_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    gb, gr = np.full_like(mb, 9), np.full_like(mr, 9)
    E.pa_emu_chroma_mean(cb.ctypes.data, cr.ctypes.data, stride, w, h, gb.ctypes.data, gr.ctypes.data)
    assert (gb == mb).all() and (gr == mr).all() and mb.any()
    gh, ga, gs = np.zeros_like(hist), np.zeros_like(avg), np.zeros_like(s16)
    E.pa_emu_chroma_histogram(cb.ctypes.data, cr.ctypes.data, stride, w, h, rw, rh, gh.ctypes.data, ga.ctypes.data, gs.ctypes.data)
    assert (gh == hist).all() and (ga == avg).all() and (gs == s16).all()
