"""Direct comparisons with the UNMODIFIED reference's own hot-path functions (oracle/_ref/libSvtVp9Enc_refdrv.so, built by
oracle/build_ref.sh from /root/reference; the .so travels to the GPU box).  Skipped where that library is absent.

Note: for picture widths that are not a multiple of 64 the reference's AVX2 HME kernels and its plain-C kernels disagree
with each other (the last SB column is 8..56 samples wide); this implementation follows the plain-C semantics, which is
what the AVX2 path also computes for every BASELINE configuration (widths 640, 1920, 3840)."""
import importlib
import zlib

import numpy as np
import pytest

import helpers as H

api = H.api
frames = importlib.import_module("svt-vp9_b200.frames")
presets = importlib.import_module("svt-vp9_b200.presets")
pytestmark = pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")

ME_CASES = [
    ("4k_params_tl4", 640, 384, lambda w, h: presets.me_params_4k_m8(w, h, 4, 2, (0, 2))),
    ("4k_params_tl0_samepoc", 512, 320, lambda w, h: presets.me_params_4k_m8(w, h, 0, 2, (0, 0))),
    ("4k_params_P", 448, 304, lambda w, h: presets.me_params_4k_m8(w, h, 3, 1, (16, 0))),
    ("1080p_params_tl2_ragged_height", 448, 248, lambda w, h: presets.me_params_1080p_m8(w, h, 2, 2, (0, 8))),
    ("1080p_params_tl1", 384, 256, lambda w, h: presets.me_params_1080p_m8(w, h, 1, 2, (0, 16))),
    ("480p_m9_params_tl0", 640, 480, lambda w, h: presets.me_params_480p_m9(w, h, 0, 2, (0, 0))),
    ("480p_m9_params_tl3", 640, 480, lambda w, h: presets.me_params_480p_m9(w, h, 3, 2, (0, 4))),
    ("4k_m5t0_params_tl0_stripes", 704, 448, lambda w, h: presets.me_params_4k_m5_tune0(w, h, 0, 2, (0, 0))),
    ("4k_m5t0_params_tl2_8x8", 512, 320, lambda w, h: presets.me_params_4k_m5_tune0(w, h, 2, 2, (0, 4))),
    ("4k_m5t0_params_tl3_nosubpel", 448, 256, lambda w, h: presets.me_params_4k_m5_tune0(w, h, 3, 1, (8, 0))),
]


def _me_inputs(name, w, h, mk):
    P = mk(w, h)
    clip = H.synth.make_clip(w, h, 3, seed=zlib.crc32(name.encode()) & 0xffff, k=3, g=4, a=5, b=3, m=32, noise=2)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    return P, pl, st, ((w + 63) // 64) * ((h + 63) // 64)


@pytest.mark.parametrize("name,w,h,mk", ME_CASES, ids=[c[0] for c in ME_CASES])
def test_oracle_me_matches_reference_functions(name, w, h, mk):
    P, pl, st, n_sb = _me_inputs(name, w, h, mk)
    ref = H.refdrv_me_picture(P, st[1], st[0], st[2] if P.num_lists == 2 else None, n_sb)
    exp = np.zeros(n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    H.oracle().svtb200_oracle_me_picture(P, st[1], st[0], st[2] if P.num_lists == 2 else None, exp.ctypes.data, None, 0, n_sb)
    assert H.me_equal(ref, exp, P.num_lists).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,mk", ME_CASES, ids=[c[0] for c in ME_CASES])
def test_cuda_me_matches_reference_functions(name, w, h, mk):
    P, pl, st, n_sb = _me_inputs(name, w, h, mk)
    ref = H.refdrv_me_picture(P, st[1], st[0], st[2] if P.num_lists == 2 else None, n_sb)
    ctx = api.Context(0)
    for i, s in enumerate(st):
        ctx.picture_upload(i, s[0], s[1], s[2])
    out = ctx.me_picture(P, 1, [0, 2] if P.num_lists == 2 else [0])
    ctx.close()
    ok = H.me_equal(out, ref, P.num_lists)
    assert ok.all(), "%d PUs differ from the reference" % (~ok).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("tl,nl,poc", [(0, 2, (0, 0)), (1, 2, (0, 16)), (4, 2, (0, 2)), (3, 1, (4, 0))], ids=["tl0", "tl1", "tl4", "P"])
def test_cuda_me_full_4k_picture_matches_reference_functions(tl, nl, poc):
    """BASELINE config C at full size: every one of the 2040 super-blocks against the reference's motion_estimate_sb
    (run on all host cores through the path driver)."""
    import os
    w, h = 3840, 2160
    P = presets.me_params_4k_m8(w, h, tl, nl, poc)
    clip = H.synth.make_clip(w, h, 3, seed=31 + tl, k=3, g=4, a=7, b=5, m=48, noise=2)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    threads = max(1, min(64, len(os.sched_getaffinity(0))))
    ref = H.refdrv_me_picture(P, st[1], st[0], st[2] if nl == 2 else None, n_sb, threads=threads)
    ctx = api.Context(0)
    for i, s_ in enumerate(st):   # the bench's path: device geometry, windows staged by TMA
        ctx.picture_upload_luma(i, s_[0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, bool(P.enable_hme_level1))
    out = ctx.me_picture(P, 1, [0, 2] if nl == 2 else [0])
    ctx.close()
    ok = H.me_equal(out, ref, nl)
    assert ok.all(), "%d of %d PUs differ from the reference" % ((~ok).sum(), ok.size)


@pytest.mark.gpu
@pytest.mark.parametrize("tl,nl,poc", [(0, 2, (0, 0)), (1, 2, (0, 8)), (2, 2, (0, 4)), (3, 2, (0, 2)), (3, 1, (8, 0))], ids=["tl0", "tl1", "tl2", "tl3", "P"])
def test_cuda_me_full_4k_m5_tune0_picture_matches_reference_functions(tl, nl, poc):
    """BASELINE config D at full size (3840x2160, enc-mode 5, tune 0: 2x2 quadrants of 64x32 scaled up to 200 % -- the HME row-stripe
    path --, HME levels 1 / 2, 16x16 full-pel area, 8x8 refinement): every super-block against the reference's motion_estimate_sb."""
    import os
    w, h = 3840, 2160
    P = presets.me_params_4k_m5_tune0(w, h, tl, nl, poc)
    clip = H.synth.make_clip(w, h, 3, seed=51 + tl, k=3, g=4, a=9, b=5, m=64, noise=2)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    threads = max(1, min(64, len(os.sched_getaffinity(0))))
    ref = H.refdrv_me_picture(P, st[1], st[0], st[2] if nl == 2 else None, n_sb, threads=threads)
    ctx = api.Context(0)
    for i, s_ in enumerate(st):   # the bench's path: device geometry, windows staged by TMA
        ctx.picture_upload_luma(i, s_[0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, bool(P.enable_hme_level1))
    out = ctx.me_picture(P, 1, [0, 2] if nl == 2 else [0])
    ctx.close()
    ok = H.me_equal(out, ref, nl)
    assert ok.all(), "%d of %d PUs differ from the reference" % ((~ok).sum(), ok.size)


@pytest.mark.gpu
@pytest.mark.parametrize("tl,nl,poc", [(0, 2, (0, 0)), (1, 2, (0, 16)), (2, 2, (0, 8)), (3, 2, (0, 4)), (4, 2, (0, 2)), (3, 1, (16, 0))], ids=["tl0", "tl1", "tl2", "tl3", "tl4", "P"])
def test_cuda_me_full_480p_m9_picture_matches_reference_functions(tl, nl, poc):
    """BASELINE config A at full size (640x480, enc-mode 9, tune 1)."""
    w, h = 640, 480
    P = presets.me_params_480p_m9(w, h, tl, nl, poc)
    clip = H.synth.make_clip(w, h, 3, seed=71 + tl, k=3, g=3, a=3, b=2, m=32, noise=0)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    ref = H.refdrv_me_picture(P, st[1], st[0], st[2] if nl == 2 else None, n_sb, threads=4)
    ctx = api.Context(0)
    for i, s_ in enumerate(st):   # the bench's path: device geometry, windows staged by TMA
        ctx.picture_upload_luma(i, s_[0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, bool(P.enable_hme_level1))
    out = ctx.me_picture(P, 1, [0, 2] if nl == 2 else [0])
    ctx.close()
    ok = H.me_equal(out, ref, nl)
    assert ok.all(), "%d of %d PUs differ from the reference" % ((~ok).sum(), ok.size)


def _txq_case(ts, seed):
    rng = np.random.default_rng(seed)
    N, Hh, W = 4 << ts, 256, 384
    src = rng.integers(0, 256, size=(Hh, W), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-30, 31, size=src.shape), 0, 255).astype(np.uint8)
    rows = H.quant_rows([20, 45, 90, 180])
    by, bx = np.meshgrid(np.arange(0, Hh, N), np.arange(0, W, N), indexing="ij")
    nb = by.size
    jobs = np.zeros(nb, dtype=api.TX_JOB_DTYPE)
    jobs["src_offset"] = by.ravel() * W + bx.ravel()
    jobs["pred_offset"] = jobs["src_offset"]
    jobs["recon_offset"] = jobs["src_offset"]
    jobs["qrow"] = rng.integers(0, len(rows), size=nb)
    jobs["tx_type"] = 0 if ts == 3 else rng.integers(0, 4, size=nb)
    jobs["flags"] = 1
    return jobs, src, pred, np.zeros_like(src), rows


@pytest.mark.gpu
@pytest.mark.parametrize("ts", [0, 1, 2, 3])
def test_cuda_txq_matches_reference_functions(ts):
    jobs, src, pred, recon, rows = _txq_case(ts, 40 + ts)
    c0, q0, d0, e0, r0 = H.refdrv_txq(ts, jobs, src, pred, recon, rows)
    ctx = api.Context(0)
    ctx.register_scans(H.vp9_tables())
    c1, q1, d1, e1, r1 = ctx.txq(ts, jobs, src, pred, recon, rows)
    ctx.close()
    assert (c0 == c1).all() and (q0 == q1).all() and (d0 == d1).all() and (e0 == e1).all() and (r0 == r1).all()


@pytest.mark.parametrize("w,h,level", [(200, 136, 20), (640, 384, 40)])
def test_oracle_deblock_matches_reference_functions(w, h, level):
    rng = np.random.default_rng(w + h)
    P, y, u, v, lfm, stride = H.synthetic_lf_case(rng, w, h, level)
    thr = H.lf_thresholds()
    a = H.oracle_deblock(P, y, u, v, lfm, stride, thr)
    b = H.refdrv_deblock(P, y, u, v, lfm, stride, thr)
    for x, z in zip(a, b):
        assert (x == z).all()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,level", [(200, 136, 20), (1920, 1080, 25)])
def test_cuda_deblock_matches_reference_functions(w, h, level):
    rng = np.random.default_rng(w + h)
    P, y, u, v, lfm, stride = H.synthetic_lf_case(rng, w, h, level)
    thr = H.lf_thresholds()
    b = H.refdrv_deblock(P, y, u, v, lfm, stride, thr)
    ctx = api.Context(0)
    a = ctx.deblock(P, y, u, v, lfm, stride, thr)
    ctx.close()
    for x, z in zip(a, b):
        assert (x == z).all()


def test_thread_pool_of_the_cpu_arm_equals_direct_calls():
    """bench.py's CPU arm runs the pictures of a mini-GOP through refdrv_pool_run (native threads, SB / TU granular tasks,
    deblocking gated on its picture's transform tasks): same results as the single calls."""
    import ctypes as C
    L = H.refdrv()
    w, h = 384, 256
    clip = H.synth.make_clip(w, h, 3, seed=77, k=3, g=4, a=5, b=3, m=32, noise=2)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    rows = H.quant_rows([45, 180])
    thr = H.lf_thresholds()
    rng = np.random.default_rng(9)
    P_lf, _, _, _, lfm, lfm_stride = H.synthetic_lf_case(rng, w, h, 28)
    keep, pics, want = [], (H.RefdrvPicture * 2)(), []
    for b, (cur, r0, r1, tl) in enumerate([(1, 0, 2, 4), (2, 0, 0, 0)]):
        P = presets.me_params_1080p_m8(w, h, tl, 2, (r0, r1))
        d = pics[b]
        d.P, d.cur, d.ref0, d.ref1 = P, st[cur], st[r0], st[r1]
        me = np.zeros(n_sb * 85, dtype=api.ME_RESULT_DTYPE)
        d.me_out = me.ctypes.data
        planes = [H.aligned_copy(clip[cur][0]), H.aligned_copy(clip[cur][1])]
        preds = [H.aligned_copy(clip[r0][0]), H.aligned_copy(clip[r0][1])]
        recs = [H.aligned_zeros((h + 16, w + 16), np.uint8), H.aligned_zeros((h // 2 + 16, w // 2 + 16), np.uint8), H.aligned_zeros((h // 2 + 16, w // 2 + 16), np.uint8)]
        outs = []
        d.n_pass = 2
        for k, (ts, N, pw, ph) in enumerate([(3, 32, w, h), (2, 16, w // 2, h // 2)]):
            by, bx = np.meshgrid(np.arange(0, ph, N), np.arange(0, pw, N), indexing="ij")
            jobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
            jobs["src_offset"] = jobs["pred_offset"] = by.ravel() * pw + bx.ravel()
            jobs["recon_offset"] = by.ravel() * recs[k].shape[1] + bx.ravel()
            jobs["qrow"], jobs["flags"] = np.arange(by.size) & 1, 1
            q, dq = H.aligned_zeros((by.size, N, N), np.int16), H.aligned_zeros((by.size, N, N), np.int16)
            eob = np.zeros(by.size, np.uint16)
            tp = d.passes[k]
            tp.tx_size, tp.n, tp.jobs = ts, by.size, jobs.ctypes.data
            tp.pl.src, tp.pl.src_stride, tp.pl.src_bytes = planes[k].ctypes.data, pw, planes[k].nbytes
            tp.pl.pred, tp.pl.pred_stride, tp.pl.pred_bytes = preds[k].ctypes.data, pw, preds[k].nbytes
            tp.pl.recon, tp.pl.recon_stride, tp.pl.recon_bytes = recs[k].ctypes.data, recs[k].shape[1], recs[k].nbytes
            tp.coeff, tp.qcoeff, tp.dqcoeff, tp.eob = None, q.ctypes.data, dq.ctypes.data, eob.ctypes.data
            outs.append((jobs, q, dq, eob))
        d.qrows = rows.ctypes.data
        d.deblock = 1 if b == 1 else 0
        d.lf = P_lf
        d.y, d.u, d.v, d.ys, d.cs = recs[0].ctypes.data, recs[1].ctypes.data, recs[2].ctypes.data, recs[0].shape[1], recs[1].shape[1]
        d.lfm, d.lfm_stride, d.thr = lfm.ctypes.data, lfm_stride, thr.ctypes.data
        keep.append((me, planes, preds, recs, outs, P))
        want.append(H.refdrv_me_picture(P, st[cur], st[r0], st[r1], n_sb))
    pool = L.refdrv_pool_new(4, 1)
    assert pool
    cpu_s = (C.c_double * 4)()
    assert L.refdrv_pool_run(pool, pics, 2, 2, 8, cpu_s) == 0
    L.refdrv_pool_free(pool)
    assert cpu_s[0] > 0 and cpu_s[1] > 0 and cpu_s[2] > 0
    for b, (me, planes, preds, recs, outs, P) in enumerate(keep):
        assert H.me_equal(me, want[b], 2).all()
        for k, (jobs, q, dq, eob) in enumerate(outs):
            j2 = jobs.copy()
            pw = planes[k].shape[1]
            j2["recon_offset"] = j2["src_offset"]
            _, q0, d0, e0, r0 = H.refdrv_txq(3 - k, j2, planes[k], preds[k], np.zeros_like(planes[k]), rows)
            assert (q == q0).all() and (dq == d0).all() and (eob == e0).all()
            if b == 0:
                assert (recs[k][:planes[k].shape[0], :pw] == r0).all()
        if b == 1:   # deblocked in place after its transform tasks: equal to the direct call on the unfiltered reconstruction
            unf = []
            for k in range(2):
                j2 = outs[k][0].copy()
                j2["recon_offset"] = j2["src_offset"]
                unf.append(H.refdrv_txq(3 - k, j2, planes[k], preds[k], np.zeros_like(planes[k]), rows)[4])
            assert not (recs[0][:h, :w] == unf[0]).all()
            fy, fu, fv = H.refdrv_deblock(P_lf, unf[0], unf[1], np.zeros_like(unf[1]), lfm, lfm_stride, thr)
            assert (recs[0][:h, :w] == fy).all() and (recs[1][:h // 2, :w // 2] == fu).all() and (recs[2][:h // 2, :w // 2] == fv).all()


def test_thread_pool_with_inter_prediction_and_rate_equals_direct_calls():
    """the CPU arm's round-2 chain inside refdrv_pool_run -- motion estimation -> open-loop partition -> the reference's inter_prediction()
    into prediction planes -> transform passes reading them -> coeff_rate_estimate of every block -- against the same steps called one
    by one (oracle partition, refdrv_inter_prediction, refdrv_txq, refdrv_coeff_rate)"""
    import ctypes as C
    import test_rate as TR
    L = H.refdrv()
    w, h = 384, 256
    clip = H.synth.make_clip(w, h, 3, seed=78, k=3, g=4, a=5, b=3, m=32, noise=2)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    cpad = [[H.aligned_copy(np.pad(f[1], 40, mode="edge")), H.aligned_copy(np.pad(f[2], 40, mode="edge"))] for f in clip]
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    rows = H.quant_rows([120])
    tc = np.ascontiguousarray(H.vp9_tables()["token_costs_default"])
    cur, r0, r1 = 1, 0, 2
    P = presets.me_params_1080p_m8(w, h, 4, 2, (r0, r1))
    n_blk = (w // 32) * (h // 32)

    def pred_picture(dst):
        pp = api.PredPicture()
        pp.width, pp.height, pp.use_subpel_flag = w, h, 1
        for l, r in enumerate((r0, r1)):
            full = pl[r]["full"]
            pp.ref[l][0], pp.ref_stride[l][0] = full.ctypes.data + frames.FULL_PAD * full.shape[1] + frames.FULL_PAD, full.shape[1]
            for c in range(2):
                a = cpad[r][c]
                pp.ref[l][1 + c], pp.ref_stride[l][1 + c] = a.ctypes.data + 40 * a.shape[1] + 40, a.shape[1]
        for k in range(3):
            pp.dst[k], pp.dst_stride[k] = dst[k].ctypes.data, dst[k].shape[1]
        return pp
    planes = [H.aligned_copy(clip[cur][0]), H.aligned_copy(clip[cur][1]), H.aligned_copy(clip[cur][2])]
    preds = [H.aligned_zeros((h, w), np.uint8), H.aligned_zeros((h // 2, w // 2), np.uint8), H.aligned_zeros((h // 2, w // 2), np.uint8)]
    recs = [H.aligned_zeros(p.shape, np.uint8) for p in preds]   # the reference's SIMD reconstruction stores aligned
    pics = (H.RefdrvPicture * 1)()
    d = pics[0]
    d.P, d.cur, d.ref0, d.ref1 = P, st[cur], st[r0], st[r1]
    me = np.zeros(n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    d.me_out = me.ctypes.data
    d.n_pass, outs = 3, []
    for k, (ts, N, plane) in enumerate([(3, 32, 0), (2, 16, 1), (2, 16, 2)]):
        pw, ph = planes[plane].shape[1], planes[plane].shape[0]
        by, bx = np.meshgrid(np.arange(0, ph, N), np.arange(0, pw, N), indexing="ij")
        jobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
        jobs["src_offset"] = jobs["pred_offset"] = jobs["recon_offset"] = by.ravel() * pw + bx.ravel()
        jobs["flags"] = 1
        q, dq = H.aligned_zeros((by.size, N, N), np.int16), H.aligned_zeros((by.size, N, N), np.int16)
        eob, rate = np.zeros(by.size, np.uint16), np.zeros(by.size, np.int32)
        rb = np.zeros(by.size, dtype=api.RATE_BLOCK_DTYPE)
        rb["plane_type"], rb["is_inter"] = (1 if plane else 0), 1
        tp = d.passes[k]
        tp.tx_size, tp.n, tp.jobs = ts, by.size, jobs.ctypes.data
        tp.pl.src, tp.pl.src_stride, tp.pl.src_bytes = planes[plane].ctypes.data, pw, planes[plane].nbytes
        tp.pl.pred, tp.pl.pred_stride, tp.pl.pred_bytes = preds[plane].ctypes.data, pw, preds[plane].nbytes
        tp.pl.recon, tp.pl.recon_stride, tp.pl.recon_bytes = recs[plane].ctypes.data, pw, recs[plane].nbytes
        tp.coeff, tp.qcoeff, tp.dqcoeff, tp.eob = None, q.ctypes.data, dq.ctypes.data, eob.ctypes.data
        tp.rblk, tp.rate = rb.ctypes.data, rate.ctypes.data
        outs.append((ts, plane, jobs, q, eob, rate, rb, dq))   # (dq only to keep the array alive)
    d.qrows = rows.ctypes.data
    d.n_blocks, d.pred, d.token_costs_all = n_blk, pred_picture(preds), tc.ctypes.data
    pool = L.refdrv_pool_new(4, 1)
    cpu_s = (C.c_double * 4)()
    assert L.refdrv_pool_run(pool, pics, 1, 2, 8, cpu_s) == 0
    L.refdrv_pool_free(pool)
    assert cpu_s[0] > 0 and cpu_s[1] > 0 and cpu_s[3] > 0
    # the same, step by step
    want_me = H.refdrv_me_picture(P, st[cur], st[r0], st[r1], n_sb)
    assert H.me_equal(me, want_me, 2).all()
    O = H.oracle()
    O.svtb200_oracle_me_best_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    blocks = np.zeros(n_blk, dtype=api.INTER_BLOCK_DTYPE)
    assert O.svtb200_oracle_me_best_blocks(want_me.ctypes.data, w, h, blocks.ctypes.data) == n_blk
    want_pred = [H.aligned_zeros(p.shape, np.uint8) for p in preds]
    L.refdrv_pred_new.restype = C.c_void_p
    L.refdrv_inter_prediction.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(api.PredPicture)]
    L.refdrv_pred_free.argtypes = [C.c_void_p]
    hd = L.refdrv_pred_new()
    pp = pred_picture(want_pred)
    assert L.refdrv_inter_prediction(hd, blocks.ctypes.data, 0, n_blk, C.byref(pp)) == 0
    L.refdrv_pred_free(hd)
    for a, b in zip(preds, want_pred):
        assert (a == b).all() and a.any()
    assert (blocks["list"][:, 1] != 0xff).any()   # some blocks are compound: both lists were read
    for ts, plane, jobs, q, eob, rate, rb, _dq in outs:
        _, q0, d0, e0, r0_ = H.refdrv_txq(ts, jobs, planes[plane], want_pred[plane], np.zeros_like(planes[plane]), rows)
        assert (q == q0).all() and (eob == e0).all() and (recs[plane] == r0_).all()
        assert (rate == TR._ref_rate(ts, np.ascontiguousarray(q0.reshape(len(jobs), -1)), e0, rb, tc)).all() and rate.any()
