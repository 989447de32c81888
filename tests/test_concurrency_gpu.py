"""One context used from several host threads at once (what the drop-in shim does: up to 20 ME threads, the deblocking
thread and the upload path share a context) and argument checking of the C ABI."""
import ctypes as C
import importlib
import threading

import numpy as np
import pytest

import helpers as H

api = H.api
frames = importlib.import_module("svt-vp9_b200.frames")
presets = importlib.import_module("svt-vp9_b200.presets")


@pytest.mark.gpu
def test_concurrent_calls_on_one_context_give_the_sequential_results():
    w, h, n_pic = 448, 320, 7
    clip = H.synth.make_clip(w, h, n_pic, seed=21, k=3, g=4, a=5, b=3, m=32, noise=2)
    pl = [frames.pa_planes(f[0]) for f in clip]
    st = [frames.plane_structs(p) for p in pl]
    ctx = api.Context(0)
    for i, s in enumerate(st):
        ctx.picture_upload(i, s[0], s[1], s[2])
    jobs = [(presets.me_params_4k_m8(w, h, 4, 2, (i - 1, i + 1)), i, [i - 1, i + 1]) for i in range(1, n_pic - 1)]
    jobs += [(presets.me_params_1080p_m8(w, h, 2, 2, (0, n_pic - 1)), 3, [0, n_pic - 1])]
    seq = [ctx.me_picture(P, cur, refs) for P, cur, refs in jobs]
    rng = np.random.default_rng(3)
    P_lf, y, u, v, lfm, stride = H.synthetic_lf_case(rng, 456, 264, 28)
    thr = H.lf_thresholds()
    lf_seq = ctx.deblock(P_lf, y, u, v, lfm, stride, thr)
    ctx.register_scans(H.vp9_tables())
    src = rng.integers(0, 256, size=(128, 256), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-25, 26, size=src.shape), 0, 255).astype(np.uint8)
    by, bx = np.meshgrid(np.arange(0, 128, 16), np.arange(0, 256, 16), indexing="ij")
    tjobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
    tjobs["src_offset"] = tjobs["pred_offset"] = tjobs["recon_offset"] = by.ravel() * 256 + bx.ravel()
    tjobs["tx_type"] = np.arange(by.size) & 3
    tjobs["flags"] = 1
    rows = H.quant_rows([40, 120])
    tx_seq = ctx.txq(2, tjobs, src, pred, np.zeros_like(src), rows)
    out, lf_out, tx_out, errs = [None] * len(jobs), [None] * 3, [None] * 3, []

    def tx_worker(k):
        try:
            for _ in range(6):
                tx_out[k] = ctx.txq(2, tjobs, src, pred, np.zeros_like(src), rows)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def me_worker(k):
        try:
            for _ in range(4):
                out[k] = ctx.me_picture(*jobs[k])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def lf_worker(k):
        try:
            for _ in range(4):
                lf_out[k] = ctx.deblock(P_lf, y, u, v, lfm, stride, thr)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=me_worker, args=(k,)) for k in range(len(jobs))] + [threading.Thread(target=lf_worker, args=(k,)) for k in range(3)]
    ts += [threading.Thread(target=tx_worker, args=(k,)) for k in range(3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    ctx.close()
    assert not errs, errs
    for a, b, (P, _, _) in zip(out, seq, jobs):
        assert H.me_equal(a, b, P.num_lists).all()
    for got in lf_out:
        for a, b in zip(got, lf_seq):
            assert (a == b).all()
    for got in tx_out:
        for a, b in zip(got, tx_seq):
            assert (a == b).all()


@pytest.mark.gpu
def test_bad_arguments_are_rejected_not_executed():
    ctx = api.Context(0)
    L = ctx.L
    P = presets.me_params_4k_m8(256, 192, 3, 2, (0, 2))
    ids = (C.c_uint64 * 2)(0, 2)
    out = np.zeros(12 * 85, dtype=api.ME_RESULT_DTYPE)
    assert L.svtb200_me_picture(ctx.h, C.byref(P), 1, ids, out.ctypes.data, None) == -5          # unknown picture id
    assert b"picture" in L.svtb200_last_error(ctx.h)
    assert L.svtb200_me_picture(ctx.h, None, 1, ids, out.ctypes.data, None) == -2                # null parameters
    P.num_lists = 3
    assert L.svtb200_me_picture(ctx.h, C.byref(P), 1, ids, out.ctypes.data, None) == -2
    assert L.svtb200_picture_release(ctx.h, 99) == -5
    assert L.svtb200_stream_wait(ctx.h, 0, 12345) == -2                                           # a ticket that was never issued
    assert L.svtb200_select_compute(ctx.h, 1) == -2                                               # only the two compute queues
    jobs = np.zeros(4, dtype=api.TX_JOB_DTYPE)
    pl = api.TxPlanes()
    q = np.zeros((4, 32, 32), dtype=np.int16)
    eob = np.zeros(4, dtype=np.uint16)
    rows = H.quant_rows([60])
    assert ctx.L.svtb200_fwd_txfm_batch(ctx.h, 7, 0, 0, 4, q.ctypes.data, q.ctypes.data) == -2    # tx_size out of range
    ctx.close()
