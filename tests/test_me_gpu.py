"""GPU parity tests for motion estimation (through the C ABI of libsvtvp9_b200.so).

  * golden: every picture dumped from the unmodified reference encoder -> results must be identical, PU by PU;
  * oracle: larger synthetic pictures with the 4K (single-quadrant HME) and 1080p parameter sets and ragged
    picture sizes, compared with the CPU oracle that is itself pinned against the reference;
  * full size: one 3840x2160 picture pair: structural invariants + a sample of super-blocks against the oracle.
"""
import glob
import zlib
import importlib
import os

import numpy as np
import pytest

import helpers as H

api = H.api
frames = importlib.import_module("svt-vp9_b200.frames")
presets = importlib.import_module("svt-vp9_b200.presets")

pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(H.GOLDEN, "me_*.bin.gz")))


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


LAYOUTS = ["host_layout", "device_layout"]


def _upload(ctx, pic_id, planes3, layout="host_layout"):
    """host_layout: the three planes as the reference holds them (svtb200_picture_upload; pitch 2 * 68 + width: windows staged by the
    threads); device_layout: svtb200_picture_upload_luma (aligned pitch, padding and decimations made on the device: windows staged by
    TMA)."""
    if layout == "host_layout":
        ctx.picture_upload(pic_id, planes3[0], planes3[1] if planes3[1].buf else None, planes3[2] if planes3[2].buf else None)
    else:
        ctx.picture_upload_luma(pic_id, planes3[0], planes3[1].origin_x if planes3[1].buf else 32, planes3[2].origin_x, bool(planes3[1].buf))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-7] for p in FIXTURES])
def test_cuda_matches_reference_dump(ctx, path, layout):
    pic = H.MePicture(path)
    cur, r0 = pic.cur_planes(), pic.ref_planes(0)
    _upload(ctx, 1, cur, layout)
    _upload(ctx, 2, r0, layout)
    ids = [2]
    if pic.params.num_lists == 2:
        _upload(ctx, 3, pic.ref_planes(1), layout)
        ids.append(3)
    out = ctx.me_picture(pic.params, 1, ids)
    ok = H.me_equal(out, pic.expected, pic.params.num_lists)
    assert ok.all(), "CUDA ME differs from the reference on %d of %d PUs" % ((~ok).sum(), ok.size)


def _synthetic_pair(w, h, seed, shift=(5, 3)):
    clip = H.synth.make_clip(w, h, 3, seed=seed, k=3, g=4, a=shift[0], b=shift[1], m=32, noise=2)
    return [frames.pa_planes(f[0]) for f in clip]


CASES = [
    ("4k_params_tl4", 640, 384, lambda w, h: presets.me_params_4k_m8(w, h, 4, 2, (0, 2))),
    ("4k_params_tl0_samepoc", 512, 320, lambda w, h: presets.me_params_4k_m8(w, h, 0, 2, (0, 0))),
    ("4k_params_P", 448, 304, lambda w, h: presets.me_params_4k_m8(w, h, 3, 1, (16, 0))),
    ("1080p_params_ragged", 456, 248, lambda w, h: presets.me_params_1080p_m8(w, h, 2, 2, (0, 8))),
    ("1080p_params_tl1", 384, 256, lambda w, h: presets.me_params_1080p_m8(w, h, 1, 2, (0, 16))),
]


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("name,w,h,mk", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_oracle(ctx, name, w, h, mk, layout):
    P = mk(w, h)
    pl = _synthetic_pair(w, h, seed=zlib.crc32(name.encode()) & 0xffff)
    structs = [frames.plane_structs(p) for p in pl]
    for i, s in enumerate(structs):
        _upload(ctx, 10 + i, s, layout)
    ids = [10, 12] if P.num_lists == 2 else [10]
    out, rcme = ctx.me_picture(P, 11, ids, want_rcme=True)
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    exp = np.zeros(n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    H.oracle().svtb200_oracle_me_picture(P, structs[1], structs[0], structs[2] if P.num_lists == 2 else None,
                                         exp.ctypes.data, None, 0, n_sb)
    ok = H.me_equal(out, exp, P.num_lists)
    assert ok.all(), "%s: CUDA ME differs from the oracle on %d of %d PUs" % (name, (~ok).sum(), ok.size)
    # rcme_distortion = sum of the 16 best 16x16 distortions
    d16 = out["cand"]["distortion"][:, 0].reshape(n_sb, 85)[:, 5:21].sum(axis=1)
    assert (rcme == d16).all()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_full_size_4k_picture(ctx, layout):
    w, h = 3840, 2160
    P = presets.me_params_4k_m8(w, h, 2, 2, (0, 8))
    pl = _synthetic_pair(w, h, seed=11, shift=(7, 5))
    structs = [frames.plane_structs(p) for p in pl]
    for i, s in enumerate(structs):
        _upload(ctx, 20 + i, s, layout)
    out = ctx.me_picture(P, 21, [20, 22])
    n_sb = 60 * 34
    r = out.reshape(n_sb, 85)
    # invariants: candidate counts, sorted distortions, MVs inside the padded reference
    assert (r["total"][:, :21] == 3).all() and (r["total"][:, 21:] == 2).all()
    d = r["cand"]["distortion"]
    assert (d[:, :21, 0] <= d[:, :21, 1]).all() and (d[:, :21, 1] <= d[:, :21, 2]).all()
    assert (np.abs(r["x_mv_l0"].astype(int)) <= 4 * (w + 64)).all()
    # a sample of super-blocks (corners, edges, bottom row without HME, interior) against the oracle
    sample = [0, 1, 59, 60, 1019, 1020 + 30, 33 * 60, 33 * 60 + 59, 17 * 60 + 31, 2039]
    exp = np.zeros(85, dtype=api.ME_RESULT_DTYPE)
    for sb in sample:
        H.oracle().svtb200_oracle_me_sb(P, structs[1], structs[0], structs[2], sb % 60, sb // 60, exp.ctypes.data, None, None)
        ok = H.me_equal(r[sb], exp, 2)
        assert ok.all(), "SB %d differs on %d PUs" % (sb, (~ok).sum())
