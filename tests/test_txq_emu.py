"""The transform / quantiser kernel sources (csrc/txq_kernel.cuh, csrc/txfm_device.cuh) run on the CPU under the emulation header
(tests/emu/txq_emu.cpp): what the GPU tests of tests/test_txq_gpu.py check on the device, on a box without one -- every stage, every size /
type, the fused kernel from planes at unaligned block origins, and the several-batches-in-one-launch kernel -- against the oracle that
tests/test_txfm_oracle.py pins to the reference's own C kernels."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

api = H.api
SIZES = [(0, tt) for tt in range(4)] + [(1, tt) for tt in range(4)] + [(2, tt) for tt in range(4)] + [(3, 0)]


class _Batch(C.Structure):
    _fields_ = [("tx_size", C.c_int32), ("n", C.c_int32), ("jobs", C.c_void_p), ("planes", api.TxPlanes), ("iscan", C.c_void_p * 4), ("coeff", C.c_void_p),
                ("qcoeff", C.c_void_p), ("dqcoeff", C.c_void_p), ("eob", C.c_void_p)]


def _E():
    E = H.emu()
    E.txq_emu_fwd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    E.txq_emu_quant.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    E.txq_emu_inv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    E.txq_emu_fused.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return E


def _iscans(tx_size):
    arrs = [np.ascontiguousarray(H.scan_of(tx_size, 0 if tx_size == 3 else t)[1], dtype=np.int16) for t in range(4)]
    return arrs, (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])


@pytest.mark.parametrize("tx_size,tx_type", SIZES)
def test_kernel_source_stages_on_the_cpu(tx_size, tx_type):
    rng = np.random.default_rng(500 + 10 * tx_size + tx_type)
    N = 4 << tx_size
    rows = H.quant_rows([0, 45, 120, 255])
    E = _E()
    isc, iscp = _iscans(tx_size)
    for kind, n in (("uniform", 37 if tx_size < 3 else 9), ("smooth", 21 if tx_size < 3 else 6), ("extreme", 9 if tx_size < 3 else 4)):
        x = np.ascontiguousarray(H.residual_blocks(rng, n, N, kind))
        c = np.full_like(x, 0x1111)
        E.txq_emu_fwd(tx_size, tx_type, 0, n, x.ctypes.data, c.ctypes.data)
        c0 = H.oracle_fwd(tx_size, tx_type, 0, x)
        assert (c == c0).all(), "forward %s" % kind
        ridx = rng.integers(0, len(rows), size=n).astype(np.uint16)
        q, dq, eob = np.full_like(x, 0x1111), np.full_like(x, 0x1111), np.full(n, 0x1111, np.uint16)
        E.txq_emu_quant(tx_size, tx_type, n, c0.ctypes.data, rows.ctypes.data, ridx.ctypes.data, iscp, q.ctypes.data, dq.ctypes.data, eob.ctypes.data)
        q0, dq0, eob0 = H.oracle_quant(tx_size, tx_type, c0, rows, ridx)
        assert (q == q0).all() and (dq == dq0).all() and (eob == eob0).all(), "quantiser %s" % kind
        pred = rng.integers(0, 256, size=(n, N, N), dtype=np.uint8)
        r = np.ascontiguousarray(pred.copy())
        dq0c, eob0c = np.ascontiguousarray(dq0), np.ascontiguousarray(eob0, dtype=np.uint16)
        E.txq_emu_inv(tx_size, tx_type, n, dq0c.ctypes.data, eob0c.ctypes.data, r.ctypes.data)
        assert (r == H.oracle_inv_add(tx_size, tx_type, dq0, eob0, pred)).all(), "inverse %s" % kind
    # raw int16 input to the inverse: the wrap behaviour of the reference network
    c = np.ascontiguousarray(rng.integers(-32768, 32768, size=(8, N, N), dtype=np.int16))
    pred = rng.integers(0, 256, size=(8, N, N), dtype=np.uint8)
    e = np.full(8, N * N, dtype=np.uint16)
    r = np.ascontiguousarray(pred.copy())
    E.txq_emu_inv(tx_size, tx_type, 8, c.ctypes.data, e.ctypes.data, r.ctypes.data)
    assert (r == H.oracle_inv_add(tx_size, tx_type, c, e, pred)).all()


def test_kernel_source_partial_fdct32_on_the_cpu():
    rng = np.random.default_rng(5)
    x = np.ascontiguousarray(H.residual_blocks(rng, 7, 32, "uniform"))
    c = np.full_like(x, 0x1111)
    _E().txq_emu_fwd(3, 0, 1, 7, x.ctypes.data, c.ctypes.data)
    assert (c == H.oracle_fwd(3, 0, 1, x)).all()


def _fused_case(rng, ts, shift, stride, rows_n, n_max):
    N = 4 << ts
    Hh = 64
    by, bx = np.meshgrid(np.arange(0, Hh - N + 1, N), np.arange(0, 128, N), indexing="ij")
    sel = rng.permutation(by.size)[:n_max]
    jobs = np.zeros(len(sel), dtype=api.TX_JOB_DTYPE)
    jobs["src_offset"] = by.ravel()[sel] * stride + bx.ravel()[sel] + shift
    jobs["pred_offset"] = jobs["src_offset"] + 3 * (shift & 1)        # the two planes need not share an alignment
    jobs["recon_offset"] = by.ravel()[sel] * stride + bx.ravel()[sel]
    jobs["qrow"] = rng.integers(0, rows_n, size=len(sel))
    jobs["tx_type"] = 0 if ts == 3 else rng.integers(0, 4, size=len(sel))
    jobs["flags"] = rng.integers(0, 2, size=len(sel)) | (2 * rng.integers(0, 2, size=len(sel)) if ts == 3 else 0)
    return jobs


def _check_fused(ts, jobs, src, pred, rows, coeff, q, dq, eob, rec):
    N, stride = 4 << ts, src.shape[1]
    exp = np.zeros_like(rec)
    for i, j in enumerate(jobs):
        sy, sx = divmod(int(j["src_offset"]), stride)
        py, px = divmod(int(j["pred_offset"]), stride)
        ry, rx = divmod(int(j["recon_offset"]), stride)
        resid = (src[sy:sy + N, sx:sx + N].astype(np.int16) - pred[py:py + N, px:px + N].astype(np.int16))[None]
        tt, part = int(j["tx_type"]), int(j["flags"]) >> 1 & 1
        c0 = H.oracle_fwd(ts, tt, part, resid)
        q0, d0, e0 = H.oracle_quant(ts, tt, c0, rows, [int(j["qrow"])])
        assert (coeff[i] == c0[0]).all() and (q[i] == q0[0]).all() and (dq[i] == d0[0]).all() and eob[i] == e0[0], "job %d" % i
        if j["flags"] & 1:
            exp[ry:ry + N, rx:rx + N] = H.oracle_inv_add(ts, tt, d0, e0, pred[py:py + N, px:px + N][None])[0]
    assert (rec == exp).all()


@pytest.mark.parametrize("multi", [0, 1], ids=["txq_kernel", "txq_multi_kernel"])
@pytest.mark.parametrize("shift", [0, 5])
def test_kernel_source_fused_from_planes_on_the_cpu(shift, multi):
    """perform_coding_loop on blocks scattered over planes (all four sizes as four batches): residual -> transform (type per job) -> quantise ->
    inverse + reconstruction; block origins at a 16-byte boundary (vector loads) and 5 bytes off it (re-aligned words)"""
    rng = np.random.default_rng(30 + shift)
    stride = 252           # rows do not share an alignment
    src = rng.integers(0, 256, size=(72, stride), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-40, 41, size=src.shape), 0, 255).astype(np.uint8)
    rows = H.quant_rows([20, 45, 160])
    batches, keep, per = (_Batch * 4)(), [], []
    for k, ts in enumerate((3, 2, 1, 0)):
        N = 4 << ts
        jobs = _fused_case(rng, ts, shift, stride, len(rows), 14 if ts == 3 else 40)
        n = len(jobs)
        rec = np.zeros_like(src)
        coeff, q, dq, eob = (np.full((n, N, N), 0x1111, np.int16) for _ in range(3)), None, None, np.full(n, 0x1111, np.uint16)
        coeff, q, dq = list(coeff)
        isc, _ = _iscans(ts)
        keep += [jobs, rec, coeff, q, dq, eob, isc]
        b = batches[k]
        b.tx_size, b.n, b.jobs = ts, n, jobs.ctypes.data
        b.planes.src, b.planes.src_stride, b.planes.src_bytes = src.ctypes.data, stride, src.nbytes
        b.planes.pred, b.planes.pred_stride, b.planes.pred_bytes = pred.ctypes.data, stride, pred.nbytes
        b.planes.recon, b.planes.recon_stride, b.planes.recon_bytes = rec.ctypes.data, stride, rec.nbytes
        for t in range(4):
            b.iscan[t] = isc[t].ctypes.data
        b.coeff, b.qcoeff, b.dqcoeff, b.eob = coeff.ctypes.data, q.ctypes.data, dq.ctypes.data, eob.ctypes.data
        per.append((ts, jobs, coeff, q, dq, eob, rec))
    assert _E().txq_emu_fused(C.addressof(batches), 4, rows.ctypes.data, multi) == 0
    for ts, jobs, coeff, q, dq, eob, rec in per:
        _check_fused(ts, jobs, src, pred, rows, coeff, q, dq, eob, rec)
