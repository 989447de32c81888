"""Intra reference-sample preparation of the encode pass on the device (rest of SURVEY row N2): generate_intra_reference_samples
(Source/Lib/Codec/EbEncDecProcess.c:1128-1315).  The oracle restatement (oracle/pred_port.c: svtb200_oracle_intra_edges) is pinned against the
reference's own function through the path driver (an EncDecContext filled like the encode pass's, neighbour arrays holding the samples around
the block, everything else poisoned); the CUDA kernel (csrc/intra_device.cuh: intra_edges_kernel) against the oracle, its source also on the CPU;
and the records feed svtb200_intra_predict_batch without a host round trip."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

api = H.api

BLOCK_8X8, BLOCK_16X16, BLOCK_32X32, BLOCK_64X64 = 3, 6, 9, 12
N8 = {BLOCK_8X8: 1, BLOCK_16X16: 2, BLOCK_32X32: 4, BLOCK_64X64: 8}


class EdgeJob(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("tx_size", C.c_uint8), ("have_left", C.c_uint8), ("have_top", C.c_uint8), ("have_right", C.c_uint8),
                ("extend_right", C.c_uint8), ("extend_bottom", C.c_uint8), ("pad", C.c_uint8 * 2)]


class RefBlk(C.Structure):
    _fields_ = [("mi_row", C.c_int32), ("mi_col", C.c_int32), ("bsize", C.c_uint8), ("tx_size", C.c_uint8), ("bmi_index", C.c_uint8), ("is_4x4", C.c_uint8)]


def job_of(blk, plane, luma_w, luma_h):
    """what a reference-side binding computes from the block (the locals of generate_intra_reference_samples, :1131-1172)"""
    ss = 1 if plane else 0
    mi_cols, mi_rows = (luma_w + 7) >> 3, (luma_h + 7) >> 3
    aoff, loff = blk.bmi_index & 1, blk.bmi_index >> 1
    bw4 = max((2 * N8[blk.bsize]) >> ss, 1)
    ox, oy = blk.mi_col * 8 + 4 * aoff, blk.mi_row * 8 + 4 * loff          # block_origin_x / _y
    round_uv = lambda v: (v >> 3) << 3
    j = EdgeJob()
    j.x0, j.y0 = ((blk.mi_col * 8) >> ss) + 4 * aoff, ((blk.mi_row * 8) >> ss) + 4 * loff
    j.tx_size = blk.tx_size
    j.have_top = int((plane == 0 and oy > 0) or (round_uv(oy) >> 1) > 0)
    j.have_left = int((plane == 0 and ox > 0) or (round_uv(ox) >> 1) > 0)
    j.have_right = int(aoff + (1 << blk.tx_size) < bw4)
    j.extend_right = int(blk.mi_col + N8[blk.bsize] > mi_cols)
    j.extend_bottom = int(blk.mi_row + N8[blk.bsize] > mi_rows)
    return j


def make_blocks(rng, n, plane, luma_w, luma_h):
    mi_cols, mi_rows = (luma_w + 7) >> 3, (luma_h + 7) >> 3
    out = []
    while len(out) < n:
        b = RefBlk()
        kind = int(rng.integers(0, 6))
        b.bsize = [BLOCK_8X8, BLOCK_8X8, BLOCK_16X16, BLOCK_32X32, BLOCK_64X64, BLOCK_16X16][kind]
        full_tx = {BLOCK_8X8: 1, BLOCK_16X16: 2, BLOCK_32X32: 3, BLOCK_64X64: 3}[b.bsize]
        if plane == 0:
            b.is_4x4 = int(kind == 0)
            b.bmi_index = int(rng.integers(0, 4)) if b.is_4x4 else 0
            b.tx_size = 0 if b.is_4x4 else (1 if kind == 5 else full_tx)     # kind 5: a transform smaller than the block
        else:
            b.is_4x4, b.bmi_index = 0, 0
            b.tx_size = {BLOCK_8X8: 0, BLOCK_16X16: 1, BLOCK_32X32: 2, BLOCK_64X64: 3}[b.bsize]
        n8 = N8[b.bsize]
        edge = rng.integers(0, 4)
        b.mi_col = 0 if edge == 0 else int(rng.integers(0, mi_cols)) // n8 * n8
        b.mi_row = 0 if edge == 1 else int(rng.integers(0, mi_rows)) // n8 * n8
        if edge == 2:
            b.mi_col = (mi_cols - 1) // n8 * n8     # the last block column: crosses the frame edge when the width is ragged
        if edge == 3:
            b.mi_row = (mi_rows - 1) // n8 * n8
        if b.is_4x4 and ((b.mi_col * 8 + 4 * (b.bmi_index & 1)) >= luma_w or (b.mi_row * 8 + 4 * (b.bmi_index >> 1)) >= luma_h):
            continue
        out.append(b)
    return out


def _plane(rng, w, h, pad=80):
    stride = w + 2 * pad
    a = rng.integers(0, 256, size=(h + 2 * pad, stride), dtype=np.uint8)
    return a, pad, stride


def _oracle(a, pad, stride, fw, fh, jobs):
    O = H.oracle()
    O.svtb200_oracle_intra_edges.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    out = np.full((len(jobs), 128), 0xAB, np.uint8)
    arr = (EdgeJob * len(jobs))(*jobs)
    O.svtb200_oracle_intra_edges(a.ctypes.data + pad * stride + pad, stride, fw, fh, C.addressof(arr), len(jobs), out.ctypes.data)
    return out


CASES = [(200, 136), (64, 64), (328, 72), (1920, 1080)]


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("plane", [0, 1, 2])
@pytest.mark.parametrize("w,h", CASES)
def test_oracle_intra_edges_match_reference(w, h, plane):
    rng = np.random.default_rng(w + 3 * plane)
    fw, fh = (w >> 1, h >> 1) if plane else (w, h)
    a, pad, stride = _plane(rng, fw, fh)
    blks = make_blocks(rng, 600, plane, w, h)
    jobs = [job_of(b, plane, w, h) for b in blks]
    want = _oracle(a, pad, stride, fw, fh, jobs)
    R = H.refdrv()
    R.refdrv_pred_new.restype = C.c_void_p
    R.refdrv_intra_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    R.refdrv_pred_free.argtypes = [C.c_void_p]
    hnd = R.refdrv_pred_new()
    got = np.full_like(want, 0xCD)
    arr = (RefBlk * len(blks))(*blks)
    assert R.refdrv_intra_edges(hnd, a.ctypes.data + pad * stride + pad, stride, plane, w, h, C.addressof(arr), len(blks), got.ctypes.data) == 0
    R.refdrv_pred_free(hnd)
    bad = np.flatnonzero((got != want).any(axis=1))
    assert bad.size == 0, "block %d: %s" % (bad[0], [(f, getattr(jobs[bad[0]], f)) for f, _ in EdgeJob._fields_[:-1]])
    flags = np.array([[j.have_top, j.have_left, j.have_right and j.tx_size == 0, j.extend_right, j.extend_bottom] for j in jobs])
    assert flags.any(axis=0)[[0, 1]].all() and (flags == 0).any(axis=0).all()     # the cases really vary
    if plane == 0:
        assert flags[:, 2].any()
    if w % 64:
        assert flags[:, 3].any()


@pytest.mark.parametrize("plane", [0, 1])
def test_kernel_source_intra_edges_on_the_cpu(plane):
    rng = np.random.default_rng(plane)
    w, h = 200, 136
    fw, fh = (w >> 1, h >> 1) if plane else (w, h)
    a, pad, stride = _plane(rng, fw, fh)
    jobs = [job_of(b, plane, w, h) for b in make_blocks(rng, 120, plane, w, h)]
    want = _oracle(a, pad, stride, fw, fh, jobs)
    E = H.emu()
    E.intra_edges_emu_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    got = np.full_like(want, 0xCD)
    arr = (EdgeJob * len(jobs))(*jobs)
    E.intra_edges_emu_run(a.ctypes.data + pad * stride + pad, stride, fw, fh, C.addressof(arr), len(jobs), got.ctypes.data)
    assert (got == want).all()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(200, 136), (1920, 1080), (3840, 2160)])
def test_cuda_intra_edges_and_prediction_from_them(w, h):
    """edge records made on the device equal the oracle's, and intra prediction fed with those records (device pointers, no host copy of
    the edges) equals the oracle's predictors on the oracle's edges"""
    rng = np.random.default_rng(h)
    ctx = api.Context(0)
    L, hd = ctx.L, ctx.h
    L.svtb200_intra_edges_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int, C.c_void_p]
    dev = []

    def up(buf, nbytes):
        d = L.svtb200_dev_alloc(hd, nbytes)
        ctx.check(L.svtb200_memcpy_h2d(hd, d, buf, nbytes), "h2d")
        dev.append(d)
        return d
    try:
        for plane in (0, 1):
            fw, fh = (w >> 1, h >> 1) if plane else (w, h)
            a, pad, stride = _plane(rng, fw, fh)
            n = 20000 if w > 1000 else 1500
            jobs = [job_of(b, plane, w, h) for b in make_blocks(rng, n, plane, w, h)]
            want = _oracle(a, pad, stride, fw, fh, jobs)
            arr = (EdgeJob * n)(*jobs)
            d_plane, d_jobs = up(a.ctypes.data, a.nbytes), up(C.addressof(arr), C.sizeof(arr))
            d_edges = up(np.full_like(want, 0xCD).ctypes.data, want.nbytes)
            ctx.check(L.svtb200_intra_edges_dev(hd, d_plane + pad * stride + pad, stride, fw, fh, d_jobs, n, d_edges), "intra_edges")
            ctx.sync()
            got = np.zeros_like(want)
            ctx.check(L.svtb200_memcpy_d2h(hd, got.ctypes.data, d_edges, got.nbytes), "d2h")
            assert (got == want).all()
            # prediction from the device-resident records: 64 blocks of every mode into a packed destination
            m = 640
            dst = np.zeros((m, 32 * 32), np.uint8)
            d_dst = up(dst.ctypes.data, dst.nbytes)
            blocks = (api.IntraBlock * m)()
            host_blocks = (api.IntraBlock * m)()
            exp = np.zeros_like(dst)
            for i in range(m):
                j = jobs[i]
                bs = 4 << j.tx_size
                for B, edges_base, dbase in ((blocks[i], d_edges, d_dst), (host_blocks[i], want.ctypes.data, exp.ctypes.data)):
                    B.above, B.left, B.dst = edges_base + 128 * i + 16, edges_base + 128 * i + 80, dbase + i * 1024
                    B.dst_stride, B.mode, B.tx_size, B.have_left, B.have_top = bs, i % 10, j.tx_size, j.have_left, j.have_top
            ctx.intra_predict(blocks)
            ctx.sync()
            ctx.check(L.svtb200_memcpy_d2h(hd, dst.ctypes.data, d_dst, dst.nbytes), "d2h")
            O = H.oracle()
            O.svtb200_oracle_intra_predict.argtypes = [C.c_void_p, C.c_int]
            O.svtb200_oracle_intra_predict(C.addressof(host_blocks), m)
            assert (dst == exp).all()
    finally:
        for d in dev:
            L.svtb200_dev_free(hd, d)
        ctx.close()
