"""compute_zz_sad: the oracle restatement against the reference's own function, the CUDA path against both."""
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
