"""GPU parity tests for the transform / quantisation kernels (through the C ABI), against the CPU oracle that is pinned
to the reference's own C kernels (tests/test_txfm_oracle.py), and directly against those kernels where oracle/_ref is present."""
import numpy as np
import pytest

import helpers as H

api = H.api
pytestmark = pytest.mark.gpu
SIZES = [(0, tt) for tt in range(4)] + [(1, tt) for tt in range(4)] + [(2, tt) for tt in range(4)] + [(3, 0)]


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    c.register_scans(H.vp9_tables())
    yield c
    c.close()


@pytest.mark.parametrize("tx_size,tx_type", SIZES)
def test_forward_quant_inverse_stagewise(ctx, tx_size, tx_type):
    rng = np.random.default_rng(1000 + 10 * tx_size + tx_type)
    N = 4 << tx_size
    rows = H.quant_rows([0, 5, 45, 120, 200, 255])
    for kind, n in (("uniform", 257), ("smooth", 130), ("extreme", 33)):
        x = H.residual_blocks(rng, n, N, kind)
        c_gpu = ctx.fwd_txfm(tx_size, tx_type, x)
        c_ref = H.oracle_fwd(tx_size, tx_type, 0, x)
        assert (c_gpu == c_ref).all(), "forward %s differs on %d coefficients" % (kind, (c_gpu != c_ref).sum())
        ridx = rng.integers(0, len(rows), size=n)
        q, dq, eob = ctx.quantize(tx_size, tx_type, c_ref, rows, ridx)
        q0, dq0, eob0 = H.oracle_quant(tx_size, tx_type, c_ref, rows, ridx)
        assert (q == q0).all() and (dq == dq0).all() and (eob == eob0).all()
        pred = rng.integers(0, 256, size=(n, N, N), dtype=np.uint8)
        r = ctx.inv_txfm_add(tx_size, tx_type, dq0, eob0, pred)
        r0 = H.oracle_inv_add(tx_size, tx_type, dq0, eob0, pred)
        assert (r == r0).all(), "inverse %s differs on %d samples" % (kind, (r != r0).sum())
    # inverse on raw int16 input: the wrap behaviour of the reference network
    c = rng.integers(-32768, 32768, size=(40, N, N), dtype=np.int16)
    pred = rng.integers(0, 256, size=(40, N, N), dtype=np.uint8)
    e = np.full(40, N * N, dtype=np.uint16)
    assert (ctx.inv_txfm_add(tx_size, tx_type, c, e, pred) == H.oracle_inv_add(tx_size, tx_type, c, e, pred)).all()
    if H.have_ref():
        x = H.residual_blocks(rng, 16, N, "uniform")
        assert (ctx.fwd_txfm(tx_size, tx_type, x) == H.ref_fwd(tx_size, tx_type, 0, x)).all()


def test_partial_fdct32(ctx):
    rng = np.random.default_rng(5)
    x = H.residual_blocks(rng, 50, 32, "uniform")
    assert (ctx.fwd_txfm(3, 0, x, partial32=1) == H.oracle_fwd(3, 0, 1, x)).all()


def test_empty_and_single(ctx):
    rows = H.quant_rows([45])
    for ts in range(4):
        N = 4 << ts
        assert ctx.fwd_txfm(ts, 0, np.zeros((0, N, N), np.int16)).shape[0] == 0
        x = np.zeros((1, N, N), np.int16)
        assert (ctx.fwd_txfm(ts, 0, x) == 0).all()
        q, dq, eob = ctx.quantize(ts, 0, x, rows)
        assert (q == 0).all() and eob[0] == 0


@pytest.mark.parametrize("tx_size", [0, 1, 2, 3])
def test_fused_txq_from_planes(ctx, tx_size):
    """perform_coding_loop on blocks scattered over planes: residual -> transform (type per job) -> quantise -> recon."""
    rng = np.random.default_rng(77 + tx_size)
    N = 4 << tx_size
    H_, W_ = 192, 320
    src = rng.integers(0, 256, size=(H_, W_), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-40, 41, size=src.shape), 0, 255).astype(np.uint8)
    recon = np.zeros_like(src)
    rows = H.quant_rows([20, 45, 90, 160])
    by, bx = np.meshgrid(np.arange(0, H_, N), np.arange(0, W_, N), indexing="ij")
    sel = rng.permutation(by.size)[: min(by.size, 777)]
    jobs = np.zeros(len(sel), dtype=api.TX_JOB_DTYPE)
    jobs["src_offset"] = by.ravel()[sel] * W_ + bx.ravel()[sel]
    jobs["pred_offset"] = jobs["src_offset"]
    jobs["recon_offset"] = jobs["src_offset"]
    jobs["qrow"] = rng.integers(0, len(rows), size=len(sel))
    jobs["tx_type"] = 0 if tx_size == 3 else rng.integers(0, 4, size=len(sel))
    jobs["flags"] = rng.integers(0, 2, size=len(sel)) | (2 * rng.integers(0, 2, size=len(sel)) if tx_size == 3 else 0)
    coeff, q, dq, eob, rec = ctx.txq(tx_size, jobs, src, pred, recon, rows)
    exp_rec = recon.copy()
    for i, j in enumerate(jobs):
        y, x = divmod(int(j["src_offset"]), W_)
        resid = (src[y:y + N, x:x + N].astype(np.int16) - pred[y:y + N, x:x + N].astype(np.int16))[None]
        tt, part = int(j["tx_type"]), int(j["flags"]) >> 1 & 1
        c0 = H.oracle_fwd(tx_size, tt, part, resid)
        q0, d0, e0 = H.oracle_quant(tx_size, tt, c0, rows, [int(j["qrow"])])
        assert (coeff[i] == c0[0]).all() and (q[i] == q0[0]).all() and (dq[i] == d0[0]).all() and eob[i] == e0[0], "job %d" % i
        if j["flags"] & 1:
            exp_rec[y:y + N, x:x + N] = H.oracle_inv_add(tx_size, tt, d0, e0, pred[y:y + N, x:x + N][None])[0]
    assert (rec == exp_rec).all()


def test_large_batch_properties(ctx):
    """2^16 32x32 blocks: linear structure checks that do not need the oracle at full size."""
    rng = np.random.default_rng(9)
    n = 1 << 16
    x = rng.integers(-255, 256, size=(n, 32, 32), dtype=np.int16)
    c = ctx.fwd_txfm(3, 0, x)
    # DC coefficient of the 32x32 DCT is round-ish(sum / 8): |dc - sum/8| small
    dc_err = np.abs(c[:, 0, 0].astype(np.int64) - (x.sum(axis=(1, 2)) // 8))
    assert dc_err.max() <= 4
    # a sample against the oracle
    idx = rng.choice(n, 64, replace=False)
    assert (c[idx] == H.oracle_fwd(3, 0, 0, x[idx])).all()
