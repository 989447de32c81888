"""Deblocking: oracle vs reference dumps, emulated kernel (no GPU) vs reference dumps, CUDA kernel vs both."""
import glob
import os

import numpy as np
import pytest

import helpers as H

api = H.api
FIX = sorted(glob.glob(os.path.join(H.GOLDEN, "lf_*.bin.gz")))


def test_lf_fixtures_present():
    assert len(FIX) >= 3


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-7] for p in FIX])
def test_oracle_matches_reference_dump(path):
    p = H.LfPicture(path)
    y, u, v = H.oracle_deblock(p.params, p.y, p.u, p.v, p.lfm, p.lfm_stride, p.thr)
    assert (y == p.ey).all() and (u == p.eu).all() and (v == p.ev).all()
    assert (p.y != p.ey).any(), "fixture must exercise the filter"


@pytest.mark.parametrize("path", FIX[:2], ids=[os.path.basename(p)[:-7] for p in FIX[:2]])
def test_emulated_kernel_matches_reference_dump(path):
    p = H.LfPicture(path)
    y, u, v = H.emu_deblock(p.params, p.y, p.u, p.v, p.lfm, p.lfm_stride, p.thr)
    assert (y == p.ey).all() and (u == p.eu).all() and (v == p.ev).all()


@pytest.mark.parametrize("w,h", [(200, 136), (264, 120), (136, 48)])
def test_emulated_kernel_matches_oracle_ragged(w, h):
    rng = np.random.default_rng(3)
    P, y, u, v, lfm, stride = H.synthetic_lf_case(rng, w, h, 20)  # e.g. 25 x 17 blocks: ragged right / bottom, odd rows
    thr = H.lf_thresholds()
    a = H.oracle_deblock(P, y, u, v, lfm, stride, thr)
    b = H.emu_deblock(P, y, u, v, lfm, stride, thr)
    for x, z in zip(a, b):
        assert (x == z).all()
    assert (a[0] != y).any()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-7] for p in FIX])
def test_cuda_matches_reference_dump(path):
    ctx = api.Context(0)
    p = H.LfPicture(path)
    y, u, v = ctx.deblock(p.params, p.y, p.u, p.v, p.lfm, p.lfm_stride, p.thr)
    ctx.close()
    assert (y == p.ey).all() and (u == p.eu).all() and (v == p.ev).all()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,level", [(200, 136, 20), (1920, 1080, 12), (3840, 2160, 32), (64, 64, 63), (72, 8, 5)])
def test_cuda_matches_oracle(w, h, level):
    rng = np.random.default_rng(w * 7 + h)
    P, y, u, v, lfm, stride = H.synthetic_lf_case(rng, w, h, level)
    thr = H.lf_thresholds()
    exp = H.oracle_deblock(P, y, u, v, lfm, stride, thr)
    ctx = api.Context(0)
    got = ctx.deblock(P, y, u, v, lfm, stride, thr)
    # idempotence of the no-op cases: level 0 leaves the picture untouched
    P0 = api.LfParams.from_buffer_copy(bytes(P))
    P0.filter_level = 0
    same = ctx.deblock(P0, y, u, v, lfm, stride, thr)
    ctx.close()
    for a, b in zip(exp, got):
        assert (a == b).all(), "%d samples differ" % (a != b).sum()
    assert (same[0] == y).all()
