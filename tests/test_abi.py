"""The C-ABI shared library: it loads without a GPU, exports every entry point include/svtb200.h declares, and its
structs have the sizes the header (and the reference's own structs, via the seam check) fix."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

import helpers as H

api = H.api
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "svtb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svtb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(os.path.join(ROOT, "svt-vp9_b200", "libsvtvp9_b200.so"))
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_sizes_match_header():
    assert api.ME_RESULT_DTYPE.itemsize == 40          # MeCuResults, EbMotionEstimationLcuResults.h:36-56
    assert api.LF_MASK_DTYPE.itemsize == 160           # LOOP_FILTER_MASK, vp9_loopfilter.h
    assert api.TX_JOB_DTYPE.itemsize == 16
    assert C.sizeof(api.Plane) == 32


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        api.Context(0)


@pytest.mark.gpu
@pytest.mark.parametrize("staged", [False, True], ids=["direct", "staged"])
@pytest.mark.parametrize("w,h,hme1", [(640, 384, True), (200, 136, True), (456, 248, False)])
def test_device_decimation_and_padding_match_host_front_end(w, h, hme1, staged):
    frames = importlib.import_module("svt-vp9_b200.frames")
    rng = np.random.default_rng(w)
    luma = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    pl = frames.pa_planes(luma)
    st = frames.plane_structs(pl)
    poisoned = pl["full"].copy()
    poisoned[:frames.FULL_PAD] = 0x55            # the padding must be regenerated on the device, not copied
    poisoned[:, :frames.FULL_PAD] = 0xAA
    st_p = frames.plane_structs({**pl, "full": poisoned})
    ctx = api.Context(0)
    if staged:  # more uploads than the ring has slots: every slot is reused while its last transfer may still be in flight
        junk_full = np.ascontiguousarray(poisoned[::-1])
        junk = frames.plane_structs({**pl, "full": junk_full})
        for k in range(12):
            ctx.picture_upload_luma(100 + k, junk[0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, hme1, staged=True)
    ctx.picture_upload_luma(7, st_p[0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, hme1, staged=staged)
    for which, key in ((0, "full"), (1, "quarter"), (2, "sixteenth")):
        if which == 1 and not hme1:
            continue
        got, d = ctx.picture_plane(7, which)
        want = pl[key]
        assert (d.origin_y, d.width, d.height) == (st[which].origin_y, st[which].width, st[which].height)
        if which:
            assert d.origin_x == st[which].origin_x
            assert (got[:, :want.shape[1]] == want).all(), key
            continue
        # the full-resolution plane has the device's own geometry: aligned pitch / origin, the host's padded row inside it ...
        assert d.stride % 16 == 0 and d.origin_x % 16 == 0 and d.origin_x >= frames.FULL_PAD
        x0, hp = d.origin_x - frames.FULL_PAD, want.shape[1]
        assert (got[:, x0:x0 + hp] == want).all(), key
        # ... and beyond it the bytes a read past the end (before the start) of the contiguous host row finds: the next (previous) row
        right = d.stride - (x0 + hp)
        assert right >= 24
        assert (got[:-1, x0 + hp:] == want[1:, :right]).all()
        assert (got[1:, :x0] == want[:-1, hp - x0:]).all()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,px,py", [(640, 384, 68, 68), (200, 136, 32, 32), (96, 40, 16, 4)])
def test_device_padding_matches_reference_generate_padding(w, h, px, py):
    """eb_vp9_generate_padding (EbMcp.c:17) from the reference library vs svtb200_pad_plane_dev on the same buffer"""
    if not H.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(w + px)
    stride = w + 2 * px
    buf = rng.integers(0, 256, size=(h + 2 * py, stride), dtype=np.uint8)
    want = buf.copy()
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libSvtVp9Enc_ref.so"))
    R.eb_vp9_generate_padding.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    R.eb_vp9_generate_padding(want.ctypes.data, stride, w, h, px, py)
    ctx = api.Context(0)
    L = ctx.L
    L.svtb200_pad_plane_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    d = L.svtb200_dev_alloc(ctx.h, buf.nbytes)
    L.svtb200_memcpy_h2d(ctx.h, d, buf.ctypes.data, buf.nbytes)
    ctx.check(L.svtb200_pad_plane_dev(ctx.h, d, stride, w, h, px, py), "pad_plane_dev")
    got = np.zeros_like(buf)
    L.svtb200_memcpy_d2h(ctx.h, got.ctypes.data, d, buf.nbytes)
    ctx.close()
    assert (got == want).all()
    assert (np.pad(buf[py:py + h, px:px + w], ((py, py), (px, px)), mode="edge") == got).all()


@pytest.mark.parametrize("w,h", [(640, 384), (200, 136), (456, 248)])
def test_kernel_sources_front_end_on_the_cpu(w, h):
    """pa_pad_full_kernel and pa_decimate_kernel (csrc/pa_front_kernel.cuh) under the CPU emulation: what svtb200_picture_upload_luma runs on the device
    after its 2-D copy of the picture area -- the device layout of the full plane (the host's padded row inside it, beyond it the bytes a read past the
    end of the contiguous host row finds) and the 1/4, 1/16 planes with their padding -- against the reference layout (frames.pa_planes)"""
    import ctypes as C
    frames = importlib.import_module("svt-vp9_b200.frames")
    rng = np.random.default_rng(h)
    luma = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    pl = frames.pa_planes(luma)
    poisoned = pl["full"].copy()
    poisoned[:frames.FULL_PAD] = 0x55            # the padding must be regenerated, not copied
    poisoned[:, :frames.FULL_PAD] = 0xAA
    dev_ox, dpitch = 80, (w + 80 + 96 + 15) & ~15           # SVTB200_DEV_PAD_L / _R of csrc/common.cuh
    rows = h + 2 * frames.FULL_PAD
    dev = np.full((rows, dpitch), 0x33, np.uint8)
    qw, qh, sw, sh = (w + 1) // 2, (h + 1) // 2, (w + 3) // 4, (h + 3) // 4
    qstride, sstride = (qw + 2 * frames.QUARTER_PAD + 15) & ~15, (sw + 2 * frames.SIXTEENTH_PAD + 15) & ~15
    quarter = np.full((qh + 2 * frames.QUARTER_PAD, qstride), 0x33, np.uint8)
    sixteenth = np.full((sh + 2 * frames.SIXTEENTH_PAD, sstride), 0x33, np.uint8)
    E = H.emu()
    E.pa_emu_front.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    E.pa_emu_front(poisoned.ctypes.data, poisoned.shape[1], frames.FULL_PAD, w, h, dev.ctypes.data, dpitch, dev_ox, quarter.ctypes.data, qstride, frames.QUARTER_PAD,
                   sixteenth.ctypes.data, sstride, frames.SIXTEENTH_PAD)
    want = pl["full"]
    x0, hp = dev_ox - frames.FULL_PAD, want.shape[1]
    assert (dev[:, x0:x0 + hp] == want).all()
    right = dpitch - (x0 + hp)
    assert right >= 24 and (dev[:-1, x0 + hp:] == want[1:, :right]).all() and (dev[1:, :x0] == want[:-1, hp - x0:]).all()
    assert (quarter[:, :pl["quarter"].shape[1]] == pl["quarter"]).all() and (sixteenth[:, :pl["sixteenth"].shape[1]] == pl["sixteenth"]).all()
