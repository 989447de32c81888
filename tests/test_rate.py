"""Coefficient rate estimation (coeff_rate_estimate, Source/Lib/Codec/EbRateDistortionCost.c:55, use_fast_coef_costing == 0):
the oracle restatement (oracle/rate_port.c) against the reference's own function through the path driver, the CUDA kernel
(csrc/rate.cu, one warp per block, position-parallel) against both."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

api = H.api


def _blocks(rng, tx_size, n, kind):
    """n blocks as the quantiser leaves them: zero beyond eob in scan order, the coefficient at eob - 1 non-zero."""
    N = 4 << tx_size
    nn = N * N
    T = H.vp9_tables()
    blk = np.zeros(n, dtype=api.RATE_BLOCK_DTYPE)
    blk["plane_type"] = rng.integers(0, 2, n)
    blk["is_inter"] = rng.integers(0, 2, n)
    # the reference derives the transform type from the block's intra mode: luma intra blocks of 4x4 .. 16x16 only
    tt = rng.integers(0, 4, n)
    tt[(blk["plane_type"] == 1) | (blk["is_inter"] == 1)] = 0
    if tx_size == 3:
        tt[:] = 0
    blk["tx_type"] = tt
    blk["coeff_ctx"] = rng.integers(0, 3, n)  # combine_entropy_contexts: 0, 1 or 2
    q = np.zeros((n, nn), dtype=np.int16)
    eob = np.zeros(n, dtype=np.uint16)
    for b in range(n):
        if kind == "empty":
            e = 0
        elif kind == "full":
            e = nn
        elif kind == "dc":
            e = 1
        else:
            e = int(min(nn, rng.geometric(1.0 / (1 + nn // 6))))
            if b % 7 == 0:
                e = int(rng.integers(0, nn + 1))
        eob[b] = e
        if e == 0:
            continue
        scan = T["scan_%d_%d" % (tx_size, int(tt[b]))]
        # category-6 tokens and the value-table edges (+-66 / +-67); |v| - 67 < 2^14 is what eb_vp9_cat6_high_cost[64] covers
        # (8-bit coefficients never exceed it; beyond it the reference reads past its table)
        if kind == "large":
            vals = rng.choice(np.array([0, 1, -1, 2, 4, 5, 10, 34, 66, -66, 67, -67, 68, 300, -2000, 15000, -16450, 16450], dtype=np.int64), size=e)
        elif kind == "sparse":  # runs of zeros: the skip-EOB-node cost tables and the energy-class contexts
            vals = rng.integers(-3, 4, size=e) * (rng.integers(0, 3, size=e) == 0)
        else:
            mag = np.maximum(1, 40 // (1 + np.arange(e) // 2))
            vals = rng.integers(-mag, mag + 1)
        vals = np.asarray(vals, dtype=np.int64)
        if vals[e - 1] == 0:
            vals[e - 1] = 1 if b & 1 else -1
        q[b, scan[:e]] = vals.astype(np.int16)
    return q, eob, blk


def _token_costs(rng, kind):
    T = H.vp9_tables()
    if kind == "default":
        return np.ascontiguousarray(T["token_costs_default"])
    # every entry distinct in its low bits: a wrong band / context / skip / token index cannot cancel out
    return rng.integers(1, 1 << 14, size=(4, 2, 2, 6, 2, 6, 12)).astype(np.uint32)


def _oracle_rate(tx_size, q, eob, blk, tc_all):
    L = H.oracle()
    T = H.vp9_tables()
    scans = [np.ascontiguousarray(T["scan_%d_%d" % (tx_size, 0 if tx_size == 3 else t)], dtype=np.int16) for t in range(4)]
    nbs = [np.ascontiguousarray(T["nb_%d_%d" % (tx_size, 0 if tx_size == 3 else t)], dtype=np.int16) for t in range(4)]
    sp = (C.c_void_p * 4)(*[s.ctypes.data for s in scans])
    npp = (C.c_void_p * 4)(*[s.ctypes.data for s in nbs])
    tabs = [np.ascontiguousarray(T[k]) for k in ("rate_vtok", "rate_vcost", "rate_c6lo", "rate_c6hi")]
    tc = np.ascontiguousarray(tc_all[tx_size])
    rate = np.zeros(len(eob), dtype=np.int32)
    L.svtb200_oracle_coeff_rate.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 11
    L.svtb200_oracle_coeff_rate.restype = None
    L.svtb200_oracle_coeff_rate(tx_size, len(eob), q.ctypes.data, eob.ctypes.data, blk.ctypes.data, tc.ctypes.data, sp, npp,
                                *(t.ctypes.data for t in tabs), rate.ctypes.data)
    return rate


def _ref_rate(tx_size, q, eob, blk, tc_all):
    L = H.refdrv()
    L.refdrv_coeff_rate.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5
    rate = np.zeros(len(eob), dtype=np.int32)
    tc_all = np.ascontiguousarray(tc_all)
    assert L.refdrv_coeff_rate(tx_size, len(eob), q.ctypes.data, eob.ctypes.data, blk.ctypes.data, tc_all.ctypes.data, rate.ctypes.data) == 0
    return rate


KINDS = ["typical", "sparse", "large", "empty", "full", "dc"]


def test_tables_match_reference():
    """the committed rate tables are the reference's (regenerate with oracle/extract_vp9_tables.py)"""
    if not H.have_refdrv():
        pytest.skip("oracle/_ref not built")
    T = H.vp9_tables()
    L = H.refdrv()
    tc = np.zeros((4, 2, 2, 6, 2, 6, 12), dtype=np.uint32)
    assert L.refdrv_token_costs(C.c_void_p(tc.ctypes.data)) == tc.nbytes
    assert (tc == T["token_costs_default"]).all()
    for ts in range(4):
        for tt in range(4):
            nb = np.zeros(2 * ((4 << ts) ** 2 + 1), dtype=np.int16)
            L.refdrv_scan_neighbors(ts, tt, C.c_void_p(nb.ctypes.data))
            assert (nb == T["nb_%d_%d" % (ts, tt)]).all()


@pytest.mark.parametrize("tx_size", [0, 1, 2, 3])
@pytest.mark.parametrize("costs", ["default", "random"])
def test_oracle_vs_reference(tx_size, costs):
    if not H.have_refdrv():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(100 + tx_size)
    tc = _token_costs(rng, costs)
    for kind in KINDS:
        q, eob, blk = _blocks(rng, tx_size, 600 if tx_size < 3 else 200, kind)
        want = _ref_rate(tx_size, q, eob, blk, tc)
        got = _oracle_rate(tx_size, q, eob, blk, tc)
        assert (got == want).all(), (kind, np.flatnonzero(got != want)[:5])


def test_golden_rates():
    """fixture made from the reference function (tests/golden/make_rate_golden.py): travels to boxes without oracle/_ref"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rate_golden.npz"))
    tc = H.vp9_tables()["token_costs_default"]
    for ts in range(4):
        got = _oracle_rate(ts, np.ascontiguousarray(g["q%d" % ts]), np.ascontiguousarray(g["eob%d" % ts]),
                           np.ascontiguousarray(g["blk%d" % ts].view(api.RATE_BLOCK_DTYPE).reshape(-1)), tc)
        assert (got == g["rate%d" % ts]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tx_size", [0, 1, 2, 3])
def test_cuda_vs_oracle(tx_size):
    rng = np.random.default_rng(200 + tx_size)
    ctx = api.Context(0)
    ctx.register_scans(H.vp9_tables())
    ctx.register_rate_tables(H.vp9_tables())
    try:
        for costs in ("default", "random"):
            tc = _token_costs(rng, costs)
            for kind in KINDS:
                q, eob, blk = _blocks(rng, tx_size, 3000 if tx_size < 3 else 700, kind)
                want = _oracle_rate(tx_size, q, eob, blk, tc)
                got = ctx.coeff_rate(tx_size, q, eob, blk, tc[tx_size])
                assert (got == want).all(), (kind, costs, np.flatnonzero(got != want)[:5])
                if H.have_refdrv():
                    assert (got == _ref_rate(tx_size, q, eob, blk, tc)).all()
        # empty batch and the golden fixture
        assert len(ctx.coeff_rate(tx_size, np.zeros((0, (4 << tx_size) ** 2), np.int16), np.zeros(0, np.uint16), np.zeros(0, api.RATE_BLOCK_DTYPE),
                                  tc[tx_size])) == 0
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_rate_of_txq_output():
    """the rate of the coefficients the fused transform kernel leaves on the device (the chain perform_dist_rate_calc runs,
    EbEncDecProcess.c:723-745): 4K-picture-sized batch, device-resident"""
    rng = np.random.default_rng(9)
    T = H.vp9_tables()
    ctx = api.Context(0)
    ctx.register_scans(T)
    ctx.register_rate_tables(T)
    try:
        src = rng.integers(0, 256, size=(512, 1024), dtype=np.uint8)
        pred = np.clip(src.astype(int) + rng.integers(-25, 26, size=src.shape), 0, 255).astype(np.uint8)
        rows = H.quant_rows([45, 120])
        for ts in (3, 2, 1, 0):
            N = 4 << ts
            by, bx = np.meshgrid(np.arange(0, 512, N), np.arange(0, 1024, N), indexing="ij")
            jobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
            jobs["src_offset"] = jobs["pred_offset"] = jobs["recon_offset"] = by.ravel() * 1024 + bx.ravel()
            jobs["qrow"] = np.arange(by.size) & 1
            jobs["tx_type"] = 0 if ts == 3 else np.arange(by.size) & 3
            jobs["flags"] = 1
            coeff, q, dq, eob, rec = ctx.txq(ts, jobs, src, pred, np.zeros_like(src), rows)
            blk = np.zeros(len(jobs), dtype=api.RATE_BLOCK_DTYPE)
            blk["tx_type"] = jobs["tx_type"]
            blk["coeff_ctx"] = np.arange(len(jobs)) % 3
            q2 = np.ascontiguousarray(q.reshape(len(jobs), -1))
            e2 = np.ascontiguousarray(eob, dtype=np.uint16)
            got = ctx.coeff_rate(ts, q2, e2, blk, T["token_costs_default"][ts])
            assert (got == _oracle_rate(ts, q2, e2, blk, T["token_costs_default"])).all()
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_rate_multi_launch():
    """svtb200_coeff_rate_multi_dev: batches of all four sizes in ONE launch == the batches one by one (and the oracle)"""
    rng = np.random.default_rng(31)
    T = H.vp9_tables()
    ctx = api.Context(0)
    ctx.register_scans(T)
    ctx.register_rate_tables(T)
    L, h = ctx.L, ctx.h
    api._bind_rate(L)
    dev = []

    def up(a):
        a = np.ascontiguousarray(a)
        d = L.svtb200_dev_alloc(h, max(a.nbytes, 16))
        ctx.check(L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes), "h2d")
        dev.append(d)
        return d
    try:
        tc = _token_costs(rng, "random")
        arr = (api.RateBatch * 6)()
        want, outs = [], []
        for k, (ts, kind, n) in enumerate(((3, "typical", 700), (2, "sparse", 900), (1, "typical", 1500), (0, "typical", 2000), (2, "dc", 300), (3, "large", 100))):
            q, eob, blk = _blocks(rng, ts, n, kind)
            want.append(_oracle_rate(ts, q, eob, blk, tc))
            b = arr[k]
            b.tx_size, b.n, b.qcoeff, b.eob, b.blocks, b.token_costs = ts, n, up(q), up(eob), up(blk), up(tc[ts])
            b.rate = up(np.zeros(n, np.int32))
            outs.append((b.rate, n))
        l0 = ctx.launches()
        ctx.check(L.svtb200_coeff_rate_multi_dev(h, arr, 6), "rate_multi")
        assert ctx.launches() - l0 == 1
        ctx.sync()
        for (d, n), w in zip(outs, want):
            got = np.zeros(n, np.int32)
            ctx.check(L.svtb200_memcpy_d2h(h, got.ctypes.data, d, got.nbytes), "d2h")
            assert (got == w).all()
    finally:
        for d in dev:
            L.svtb200_dev_free(h, d)
        ctx.close()


# ---------------------------------------------------------------------------------------------- kernel sources on the CPU
class _RateEmuBatch(C.Structure):
    _fields_ = [("tx_size", C.c_int32), ("n", C.c_int32), ("qcoeff", C.c_void_p), ("eob", C.c_void_p), ("blocks", C.c_void_p), ("token_costs", C.c_void_p),
                ("scan", C.c_void_p * 4), ("nb", C.c_void_p * 4), ("rate", C.c_void_p)]


class _PackEmuBatch(C.Structure):
    _fields_ = [("tx_size", C.c_int32), ("n", C.c_int32), ("jobs", C.c_void_p), ("qcoeff", C.c_void_p), ("eob", C.c_void_p), ("scan", C.c_void_p * 4)]


def _tables(tx_size):
    T = H.vp9_tables()
    scans = [np.ascontiguousarray(T["scan_%d_%d" % (tx_size, 0 if tx_size == 3 else t)], dtype=np.int16) for t in range(4)]
    nbs = [np.ascontiguousarray(T["nb_%d_%d" % (tx_size, 0 if tx_size == 3 else t)], dtype=np.int16) for t in range(4)]
    return scans, nbs


@pytest.mark.parametrize("costs", ["default", "random"])
def test_kernel_source_rate_on_the_cpu(costs):
    """rate_multi_kernel (csrc/rate_kernel.cuh) under the CPU emulation: all four transform sizes as the batches of ONE launch, every kind of
    block, against the oracle (which the tests above pin against the reference's coeff_rate_estimate)"""
    rng = np.random.default_rng(7)
    tc = _token_costs(rng, costs)
    T = H.vp9_tables()
    tabs = [np.ascontiguousarray(T[k]) for k in ("rate_vtok", "rate_vcost", "rate_c6lo", "rate_c6hi")]
    keep, batches, want = [], (_RateEmuBatch * 4)(), []
    for ts in range(4):
        parts = [_blocks(rng, ts, 40 if ts < 3 else 12, kind) for kind in KINDS]
        q, eob, blk = (np.ascontiguousarray(np.concatenate([p[i] for p in parts])) for i in range(3))
        scans, nbs = _tables(ts)
        tcs, rate = np.ascontiguousarray(tc[ts]), np.full(len(eob), -7, np.int32)
        keep += [q, eob, blk, scans, nbs, tcs, rate]
        b = batches[ts]
        b.tx_size, b.n, b.qcoeff, b.eob, b.blocks, b.token_costs, b.rate = ts, len(eob), q.ctypes.data, eob.ctypes.data, blk.ctypes.data, tcs.ctypes.data, rate.ctypes.data
        for t in range(4):
            b.scan[t], b.nb[t] = scans[t].ctypes.data, nbs[t].ctypes.data
        want.append((_oracle_rate(ts, q, eob, blk, tc), rate))
    E = H.emu()
    E.rate_emu_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert E.rate_emu_run(C.addressof(batches), 4, *[t.ctypes.data for t in tabs]) == 0
    for ts, (w, got) in enumerate(want):
        assert (got == w).all(), (ts, np.flatnonzero(got != w)[:5])


def test_kernel_source_pack_on_the_cpu():
    """pack_sum_kernel + pack_kernel (csrc/pack_kernel.cuh) under the CPU emulation: the scan-order prefix (eob entries, padded with a zero to an
    even count) of every block of three batches, back to back, against a direct restatement of what tokenize_b reads (vp9_tokenize.c:275)"""
    rng = np.random.default_rng(9)
    T = H.vp9_tables()
    keep, batches, runs = [], (_PackEmuBatch * 3)(), []
    for k, ts in enumerate((3, 1, 0)):
        parts = [_blocks(rng, ts, 50 if ts < 3 else 20, kind) for kind in ("typical", "sparse", "empty", "full", "dc")]
        q, eob, blk = (np.ascontiguousarray(np.concatenate([p[i] for p in parts])) for i in range(3))
        jobs = np.zeros(len(eob), dtype=api.TX_JOB_DTYPE)
        jobs["tx_type"] = blk["tx_type"]
        scans, _ = _tables(ts)
        keep += [q, eob, jobs, scans]
        b = batches[k]
        b.tx_size, b.n, b.jobs, b.qcoeff, b.eob = ts, len(eob), jobs.ctypes.data, q.ctypes.data, eob.ctypes.data
        for t in range(4):
            b.scan[t] = scans[t].ctypes.data
        for j in range(len(eob)):
            e = int(eob[j])
            run = q[j][scans[int(jobs["tx_type"][j]) & 3][:e]].astype(np.int16)
            runs.append(np.concatenate([run, np.zeros(e & 1, np.int16)]))
    want_stream = np.concatenate(runs)
    want_off = np.concatenate([[0], np.cumsum([len(r) for r in runs])]).astype(np.uint32)
    stream, off = np.full(len(want_stream) + 64, 0x7777, np.int16), np.zeros(len(runs) + 1, np.uint32)
    E = H.emu()
    E.pack_emu_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
    assert E.pack_emu_run(C.addressof(batches), 3, stream.ctypes.data, len(stream), off.ctypes.data) == 0
    assert (off == want_off).all()
    assert (stream[:len(want_stream)] == want_stream).all() and (stream[len(want_stream):] == 0x7777).all()
    # capacity smaller than the stream: the total is still reported, blocks that do not fit are left out, nothing is written past the capacity
    cap = int(want_off[len(runs) // 2])
    stream2, off2 = np.full(len(want_stream) + 64, 0x7777, np.int16), np.zeros(len(runs) + 1, np.uint32)
    assert E.pack_emu_run(C.addressof(batches), 3, stream2.ctypes.data, cap, off2.ctypes.data) == 0
    assert off2[-1] == want_off[-1] and (stream2[:cap] == want_stream[:cap]).all() and (stream2[cap:] == 0x7777).all()
