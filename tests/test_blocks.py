"""The block walk in front of the inter predictors (inter_prediction -> build_inter_predictors, Source/Lib/VPX/vp9_reconinter.c:102-250,
with eb_vp9_clamp_mv_to_umv_border_sb :68) and the bench's open-loop partition: the oracle restatement (oracle/blocks_port.c +
pred_port.c) against the reference's own inter_prediction() on a filled-in encode-pass context, the CUDA kernels
(csrc/blocks.cu) against the oracle -- descriptor by descriptor and through the whole device chain ME -> blocks -> descriptors ->
prediction."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import test_pred as TP

api = H.api
PAD = 80


def _planes(rng, w, h, kind="noise"):
    """Y, U, V reference planes with the reference's 80 / 40 sample padding"""
    return [TP._ref_plane(rng, w, h, kind), TP._ref_plane(rng, w // 2 + 40, h // 2 + 40, kind)[:h // 2 + 80, :w // 2 + 80].copy(),
            TP._ref_plane(rng, w // 2 + 40, h // 2 + 40, kind)[:h // 2 + 80, :w // 2 + 80].copy()]


def _pad_of(plane):
    return PAD if plane == 0 else PAD // 2


def _picture(w, h, refs, dst, addr_of, subpel=1):
    pic = api.PredPicture()
    pic.width, pic.height, pic.use_subpel_flag = w, h, subpel
    for l in range(2):
        for p in range(3):
            a = refs[l][p]
            pic.ref[l][p] = addr_of(a) + _pad_of(p) * a.shape[1] + _pad_of(p)
            pic.ref_stride[l][p] = a.shape[1]
    for p in range(3):
        pic.dst[p] = addr_of(dst[p])
        pic.dst_stride[p] = dst[p].shape[1]
    return pic


def _random_blocks(rng, w, h, n):
    """non-overlapping blocks, every size 8..64 (squares and 2:1 rectangles), vectors far enough to hit
    every clamp (left / right / top / bottom), one or two references"""
    sizes = [(64, 64), (32, 32), (16, 16), (8, 8), (64, 32), (32, 64), (32, 16), (16, 32), (16, 8), (8, 16)]
    taken, out = np.zeros(((h + 63) // 8 + 8, (w + 63) // 8 + 8), bool), []
    mi_w, mi_h = (w + 7) // 8, (h + 7) // 8
    for _ in range(n * 4):
        bw, bh = sizes[int(rng.integers(0, len(sizes)))]
        # positions are multiples of the block size as in a partition tree (the reference's SIMD predictors store aligned);
        # the block may hang over the right / bottom edge
        x, y = int(rng.integers(0, (mi_w * 8 + bw - 1) // bw)) * bw, int(rng.integers(0, (mi_h * 8 + bh - 1) // bh)) * bh
        if x + bw > ((w + 63) & ~63) or y + bh > ((h + 63) & ~63):
            continue
        if taken[y // 8:(y + bh) // 8, x // 8:(x + bw) // 8].any():
            continue
        taken[y // 8:(y + bh) // 8, x // 8:(x + bw) // 8] = True
        b = np.zeros((), dtype=api.INTER_BLOCK_DTYPE)
        b["x"], b["y"], b["bw"], b["bh"] = x, y, bw, bh
        comp = int(rng.integers(0, 3)) == 0
        b["list"] = [int(rng.integers(0, 2)), int(rng.integers(0, 2)) if comp else 0xff]
        scale = [40, 400, 3000, 20000][int(rng.integers(0, 4))]
        b["mv"] = rng.integers(-scale, scale + 1, size=(2, 2))
        if not comp:
            b["mv"][1] = 0
        out.append(b)
        if len(out) == n:
            break
    return np.array(out, dtype=api.INTER_BLOCK_DTYPE)


def _oracle_expand(blocks, pic):
    O = H.oracle()
    O.svtb200_oracle_inter_expand.argtypes = [C.c_void_p, C.c_int, C.POINTER(api.PredPicture), C.c_void_p]
    O.svtb200_oracle_inter_expand.restype = None
    out = (api.PredBlock * (3 * len(blocks)))()
    O.svtb200_oracle_inter_expand(blocks.ctypes.data, len(blocks), C.byref(pic), C.addressof(out))
    return out


def _oracle_chain(blocks, w, h, refs, subpel):
    dst = [np.zeros(((h + 63) & ~63, (w + 63) & ~63), np.uint8), np.zeros((((h + 63) & ~63) // 2, ((w + 63) & ~63) // 2), np.uint8),
           np.zeros((((h + 63) & ~63) // 2, ((w + 63) & ~63) // 2), np.uint8)]
    pic = _picture(w, h, refs, dst, lambda a: a.ctypes.data, subpel)
    desc = _oracle_expand(blocks, pic)
    O = H.oracle()
    O.svtb200_oracle_inter_predict.argtypes = [C.c_void_p, C.c_int]
    O.svtb200_oracle_inter_predict(C.addressof(desc), len(desc))
    return dst, desc


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("subpel", [1, 0])
@pytest.mark.parametrize("size", [(200, 136), (640, 384)])
def test_oracle_block_walk_matches_reference_inter_prediction(size, subpel):
    w, h = size
    rng = np.random.default_rng(w + subpel)
    refs = [_planes(rng, (w + 63) & ~63, (h + 63) & ~63), _planes(rng, (w + 63) & ~63, (h + 63) & ~63)]
    blocks = _random_blocks(rng, w, h, 300)
    assert len(blocks) > 40 and (blocks["list"][:, 1] != 0xff).any() and (blocks["list"][:, 1] == 0xff).any()
    got, desc = _oracle_chain(blocks, w, h, refs, subpel)
    R = H.refdrv()
    R.refdrv_pred_new.restype = C.c_void_p
    R.refdrv_inter_prediction.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(api.PredPicture)]
    R.refdrv_pred_free.argtypes = [C.c_void_p]
    want = [H.aligned_zeros(g.shape, np.uint8) for g in got]
    pic = _picture(w, h, refs, want, lambda a: a.ctypes.data, subpel)
    hd = R.refdrv_pred_new()
    assert R.refdrv_inter_prediction(hd, blocks.ctypes.data, 0, len(blocks), C.byref(pic)) == 0
    R.refdrv_pred_free(hd)
    for p in range(3):
        assert (got[p] == want[p]).all(), (p, int((got[p] != want[p]).sum()))
        assert got[p].any()
    # the clamps were exercised on every side
    xs = np.array([d.x[0] for d in desc[0::3]]); ys = np.array([d.y[0] for d in desc[0::3]])
    assert xs.min() < 0 and ys.min() < 0 and xs.max() > w - 8 and ys.max() > h - 8


def _synthetic_me(rng, w, h, num_lists=2):
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    me = np.zeros(n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    for k in ("x_mv_l0", "y_mv_l0", "x_mv_l1", "y_mv_l1"):
        me[k] = rng.integers(-300, 301, size=len(me))
    v = me.view(np.uint8).reshape(len(me), 40)
    v[:, 12] = rng.integers(0, 3 if num_lists == 2 else 1, size=len(me))   # cand[0].direction (distortion u32 at 8, direction u32 at 12)
    return me


def test_oracle_open_loop_partition_geometry():
    """32x32 blocks + 16x16 strips cover every 16x16 cell of the picture exactly once; vectors and lists follow the candidate"""
    rng = np.random.default_rng(3)
    O = H.oracle()
    O.svtb200_oracle_me_best_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for w, h in ((3840, 2160), (208, 144), (64, 48), (80, 80)):
        me = _synthetic_me(rng, w, h)
        blocks = np.zeros((w // 16) * (h // 16), dtype=api.INTER_BLOCK_DTYPE)
        n = O.svtb200_oracle_me_best_blocks(me.ctypes.data, w, h, blocks.ctypes.data)
        blocks = blocks[:n]
        cover = np.zeros((h // 16, w // 16), int)
        for b in blocks:
            cover[b["y"] // 16:(b["y"] + b["bh"]) // 16, b["x"] // 16:(b["x"] + b["bw"]) // 16] += 1
        assert (cover == 1).all()
        if (w, h) == (3840, 2160):
            assert n == 120 * 67 + 240
        b = blocks[0]   # the 32x32 block at (0,0): PU 1 of super-block 0
        r = me[1]
        d = int(me[1:2].view(np.uint8)[12])
        assert int(b["list"][1]) == (1 if d == 2 else 0xff) and int(b["list"][0]) == (1 if d == 1 else 0)
        assert int(b["mv"][0][1]) == 2 * int(r["x_mv_l1"] if d == 1 else r["x_mv_l0"])


@pytest.mark.parametrize("subpel", [1, 0])
@pytest.mark.parametrize("size", [(200, 136), (640, 384), (1920, 1080)])
def test_kernel_source_block_walk_on_the_cpu(size, subpel):
    """the per-block arithmetic the two kernels of csrc/blocks.cu evaluate (csrc/blocks_device.cuh), on the CPU: descriptors of every block and
    plane -- every clamp side, both use_subpel_flag values, compound blocks -- and the open-loop partition, against the oracle"""
    w, h = size
    rng = np.random.default_rng(w + subpel)
    refs = [_planes(rng, w, h), _planes(rng, w, h)]
    dst = [np.zeros(((h + 63) & ~63, (w + 63) & ~63), np.uint8), np.zeros((((h + 63) & ~63) // 2, ((w + 63) & ~63) // 2), np.uint8),
           np.zeros((((h + 63) & ~63) // 2, ((w + 63) & ~63) // 2), np.uint8)]
    pic = _picture(w, h, refs, dst, lambda a: a.ctypes.data, subpel)
    blocks = _random_blocks(rng, w, h, 400)
    want = _oracle_expand(blocks, pic)
    E = H.emu()
    E.blocks_emu_expand.argtypes = [C.c_void_p, C.c_int, C.POINTER(api.PredPicture), C.c_void_p]
    got = (api.PredBlock * (3 * len(blocks)))()
    E.blocks_emu_expand(blocks.ctypes.data, len(blocks), C.byref(pic), C.addressof(got))
    assert bytes(got) == bytes(want)
    if w % 16 == 0 and h % 16 == 0:
        O = H.oracle()
        O.svtb200_oracle_me_best_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        E.blocks_emu_me_best.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        me = _synthetic_me(rng, w, h)
        a, b = np.zeros((w // 16) * (h // 16), dtype=api.INTER_BLOCK_DTYPE), np.zeros((w // 16) * (h // 16), dtype=api.INTER_BLOCK_DTYPE)
        n = O.svtb200_oracle_me_best_blocks(me.ctypes.data, w, h, a.ctypes.data)
        assert E.blocks_emu_me_best(me.ctypes.data, w, h, b.ctypes.data) == n and n > 0
        assert a[:n].tobytes() == b[:n].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("subpel", [1, 0])
def test_cuda_block_walk_descriptors_and_prediction(subpel):
    w, h = 640, 384
    rng = np.random.default_rng(17 + subpel)
    refs = [_planes(rng, w, h), _planes(rng, w, h)]
    blocks = _random_blocks(rng, w, h, 500)
    want, wdesc = _oracle_chain(blocks, w, h, refs, subpel)
    ctx = api.Context(0)
    L, hd = ctx.L, ctx.h
    dev = []

    def up(a):
        a = np.ascontiguousarray(a)
        d = L.svtb200_dev_alloc(hd, max(a.nbytes, 16))
        ctx.check(L.svtb200_memcpy_h2d(hd, d, a.ctypes.data, a.nbytes), "h2d")
        dev.append(d)
        return d
    try:
        addr = {}
        for l in range(2):
            for p in range(3):
                addr[id(refs[l][p])] = up(refs[l][p])
        got = [np.zeros_like(x) for x in want]
        for g in got:
            addr[id(g)] = up(g)
        pic = _picture(w, h, refs, got, lambda a: addr[id(a)], subpel)
        d_blk = up(blocks)
        d_desc = L.svtb200_dev_alloc(hd, 64 * 3 * len(blocks)); dev.append(d_desc)
        ctx.check(L.svtb200_inter_blocks_expand_dev(hd, d_blk, len(blocks), C.byref(pic), d_desc), "expand")
        ctx.check(L.svtb200_inter_predict_batch_dev(hd, d_desc, 3 * len(blocks)), "predict")
        ctx.sync()
        gdesc = (api.PredBlock * (3 * len(blocks)))()
        ctx.check(L.svtb200_memcpy_d2h(hd, C.addressof(gdesc), d_desc, C.sizeof(gdesc)), "d2h")
        wpic = _picture(w, h, refs, want, lambda a: a.ctypes.data, subpel)
        for i, (a, b) in enumerate(zip(gdesc, wdesc)):
            p = i % 3
            for r in range(2):
                l = [k for k in range(2) if pic.ref[k][p] == a.ref[r]]
                assert l and wpic.ref[l[0]][p] == b.ref[r] or pic.ref[0][p] == pic.ref[1][p]
                assert (a.x[r], a.y[r], a.subpel_x[r], a.subpel_y[r], a.ref_stride[r]) == (b.x[r], b.y[r], b.subpel_x[r], b.subpel_y[r], b.ref_stride[r]), i
            assert (a.w, a.h, a.compound, a.dst_stride, a.dst - pic.dst[p]) == (b.w, b.h, b.compound, b.dst_stride, b.dst - wpic.dst[p]), i
        for p in range(3):
            ctx.check(L.svtb200_memcpy_d2h(hd, got[p].ctypes.data, addr[id(got[p])], got[p].nbytes), "d2h")
            assert (got[p] == want[p]).all(), p
    finally:
        for d in dev:
            L.svtb200_dev_free(hd, d)
        ctx.close()


@pytest.mark.gpu
def test_cuda_open_loop_partition():
    rng = np.random.default_rng(23)
    O = H.oracle()
    O.svtb200_oracle_me_best_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    ctx = api.Context(0)
    L, hd = ctx.L, ctx.h
    try:
        for w, h in ((3840, 2160), (208, 144), (80, 80)):
            me = _synthetic_me(rng, w, h)
            want = np.zeros((w // 16) * (h // 16), dtype=api.INTER_BLOCK_DTYPE)
            n = O.svtb200_oracle_me_best_blocks(me.ctypes.data, w, h, want.ctypes.data)
            d_me = L.svtb200_dev_alloc(hd, me.nbytes)
            d_b = L.svtb200_dev_alloc(hd, want.nbytes)
            ctx.check(L.svtb200_memcpy_h2d(hd, d_me, me.ctypes.data, me.nbytes), "h2d")
            assert L.svtb200_me_best_blocks_dev(hd, d_me, w, h, d_b) == n
            got = np.zeros_like(want)
            ctx.check(L.svtb200_memcpy_d2h(hd, got.ctypes.data, d_b, n * 16), "d2h")
            assert (got[:n].view(np.uint8) == want[:n].view(np.uint8)).all()
            L.svtb200_dev_free(hd, d_me); L.svtb200_dev_free(hd, d_b)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_inter_chain_equals_single_calls():
    """svtb200_inter_chain_dev (partition -> block walk -> prediction -> transform batches -> packing -> rate in ONE call, on a chosen
    queue behind another) against the same stages through their own entry points: every output buffer identical."""
    import test_rate as TR
    w, h = 384, 256
    rng = np.random.default_rng(41)
    T = H.vp9_tables()
    refs = [_planes(rng, w, h, "smooth"), _planes(rng, w, h, "smooth")]
    src = [np.ascontiguousarray(r[PAD if p == 0 else PAD // 2:(PAD if p == 0 else PAD // 2) + (h if p == 0 else h // 2),
                                  PAD if p == 0 else PAD // 2:(PAD if p == 0 else PAD // 2) + (w if p == 0 else w // 2)]) for p, r in enumerate(refs[0])]
    src = [np.clip(s.astype(int) + rng.integers(-12, 13, size=s.shape), 0, 255).astype(np.uint8) for s in src]
    me = _synthetic_me(rng, w, h)
    for k in ("x_mv_l0", "y_mv_l0", "x_mv_l1", "y_mv_l1"):
        me[k] = rng.integers(-40, 41, size=len(me))
    n_blk = (w // 32) * (h // 32)
    rows = H.quant_rows([100])
    ctx = api.Context(0)
    ctx.register_scans(T)
    ctx.register_rate_tables(T)
    api._bind_txq(ctx.L)
    api._bind_rate(ctx.L)
    L, hd = ctx.L, ctx.h
    dev = []

    def up(a):
        a = np.ascontiguousarray(a)
        d = L.svtb200_dev_alloc(hd, max(a.nbytes, 16))
        ctx.check(L.svtb200_memcpy_h2d(hd, d, a.ctypes.data, a.nbytes), "h2d")
        dev.append(d)
        return d

    def down(d, like):
        out = np.zeros_like(like)
        ctx.check(L.svtb200_memcpy_d2h(hd, out.ctypes.data, d, out.nbytes), "d2h")
        return out
    try:
        d_ref = {id(a): up(a) for l in refs for a in l}
        d_src = [up(s) for s in src]
        d_me, d_rows = up(me), up(rows)
        d_tc = [up(T["token_costs_default"][ts]) for ts in range(4)]
        passes = [(3, 32, 0), (2, 16, 1), (2, 16, 2)]
        jobs, rblk = [], []
        for ts, N, p in passes:
            pw, ph = src[p].shape[1], src[p].shape[0]
            by, bx = np.meshgrid(np.arange(0, ph, N), np.arange(0, pw, N), indexing="ij")
            j = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
            j["src_offset"] = j["pred_offset"] = j["recon_offset"] = by.ravel() * pw + bx.ravel()
            j["flags"] = 1
            rb = np.zeros(by.size, dtype=api.RATE_BLOCK_DTYPE)
            rb["plane_type"], rb["is_inter"] = (1 if p else 0), 1
            jobs.append(j); rblk.append(rb)
        d_jobs, d_rblk = [up(j) for j in jobs], [up(r) for r in rblk]
        n_tu, n_coef = sum(len(j) for j in jobs), sum(len(j) * N * N for j, (_, N, _) in zip(jobs, passes))
        results = []
        for variant in ("single", "chain"):
            bufs = {"pred": [up(np.zeros_like(s)) for s in src], "rec": [up(np.zeros_like(s)) for s in src],
                    "q": [up(np.zeros(len(j) * N * N, np.int16)) for j, (_, N, _) in zip(jobs, passes)],
                    "dq": [up(np.zeros(len(j) * N * N, np.int16)) for j, (_, N, _) in zip(jobs, passes)],
                    "eob": [up(np.zeros(len(j), np.uint16)) for j in jobs], "rate": [up(np.zeros(len(j), np.int32)) for j in jobs],
                    "blk": up(np.zeros(n_blk, api.INTER_BLOCK_DTYPE)), "desc": up(np.zeros(3 * n_blk * 64, np.uint8)),
                    "stream": up(np.zeros(n_coef, np.int16)), "off": up(np.zeros(n_tu + 1, np.uint32))}
            pic = api.PredPicture()
            pic.width, pic.height, pic.use_subpel_flag = w, h, 1
            for l in range(2):
                for p in range(3):
                    a = refs[l][p]
                    pic.ref[l][p] = d_ref[id(a)] + _pad_of(p) * a.shape[1] + _pad_of(p)
                    pic.ref_stride[l][p] = a.shape[1]
            for p in range(3):
                pic.dst[p], pic.dst_stride[p] = bufs["pred"][p], src[p].shape[1]
            tx = (api.TxBatch * 3)()
            rt = (api.RateBatch * 3)()
            for k, (ts, N, p) in enumerate(passes):
                b = tx[k]
                b.tx_size, b.n, b.jobs = ts, len(jobs[k]), d_jobs[k]
                b.planes.src, b.planes.src_stride, b.planes.src_bytes = d_src[p], src[p].shape[1], src[p].nbytes
                b.planes.pred, b.planes.pred_stride, b.planes.pred_bytes = bufs["pred"][p], src[p].shape[1], src[p].nbytes
                b.planes.recon, b.planes.recon_stride, b.planes.recon_bytes = bufs["rec"][p], src[p].shape[1], src[p].nbytes
                b.coeff, b.qcoeff, b.dqcoeff, b.eob = None, bufs["q"][k], bufs["dq"][k], bufs["eob"][k]
                r = rt[k]
                r.tx_size, r.n, r.qcoeff, r.eob, r.blocks, r.token_costs, r.rate = ts, len(jobs[k]), bufs["q"][k], bufs["eob"][k], d_rblk[k], d_tc[ts], bufs["rate"][k]
            if variant == "single":
                assert L.svtb200_me_best_blocks_dev(hd, d_me, w, h, bufs["blk"]) == n_blk
                ctx.check(L.svtb200_inter_blocks_expand_dev(hd, bufs["blk"], n_blk, C.byref(pic), bufs["desc"]), "expand")
                ctx.check(L.svtb200_inter_predict_batch_dev(hd, bufs["desc"], 3 * n_blk), "predict")
                ctx.check(L.svtb200_txq_multi_dev(hd, tx, 3, d_rows, 1), "txq")
                ctx.check(L.svtb200_coeff_pack_dev(hd, tx, 3, bufs["stream"], n_coef, bufs["off"]), "pack")
                ctx.check(L.svtb200_coeff_rate_multi_dev(hd, rt, 3), "rate")
            else:
                ch = api.InterChain()
                ch.after_queue, ch.on_queue = api.COMPUTE, api.COMPUTE2   # the chain on the second queue, behind what the first holds
                ch.me, ch.blocks, ch.n_blocks, ch.pic, ch.desc = d_me, bufs["blk"], n_blk, pic, bufs["desc"]
                ch.tx, ch.n_tx, ch.qrows, ch.n_qrows = tx, 3, d_rows, 1
                ch.stream, ch.capacity, ch.offsets = bufs["stream"], n_coef, bufs["off"]
                ch.rate, ch.n_rate, ch.n_reads = rt, 3, 0
                l0 = ctx.launches()
                ctx.check(L.svtb200_inter_chain_dev(hd, C.addressof(ch)), "chain")
                assert ctx.launches() - l0 == 7
            ctx.sync()
            out = [down(bufs["pred"][p], src[p]) for p in range(3)] + [down(bufs["rec"][p], src[p]) for p in range(3)]
            out += [down(bufs["q"][k], np.zeros(len(jobs[k]) * N * N, np.int16)) for k, (_, N, _) in enumerate(passes)]
            out += [down(bufs["eob"][k], np.zeros(len(jobs[k]), np.uint16)) for k in range(3)] + [down(bufs["rate"][k], np.zeros(len(jobs[k]), np.int32)) for k in range(3)]
            out += [down(bufs["off"], np.zeros(n_tu + 1, np.uint32)), down(bufs["stream"], np.zeros(n_coef, np.int16))]
            results.append(out)
        for a, b in zip(*results):
            assert (a == b).all()
        assert results[0][0].any() and results[0][9].any() and results[0][12].any()
        # ... and the rates are the oracle's
        for k, (ts, N, p) in enumerate(passes):
            q = results[0][6 + k].reshape(len(jobs[k]), N * N)
            assert (results[0][12 + k] == TR._oracle_rate(ts, np.ascontiguousarray(q), results[0][9 + k], rblk[k], T["token_costs_default"])).all()
    finally:
        for d in dev:
            L.svtb200_dev_free(hd, d)
        ctx.close()
