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


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [0, 4, 1, 7])
def test_txq_rows_at_any_byte_alignment(shift):
    """Source / prediction blocks are fetched as 16-byte vectors when their address allows it and as re-aligned words otherwise:
    every alignment of the block origin (plane stride 1004 keeps rows from sharing an alignment) against the oracle chain."""
    rng = np.random.default_rng(60 + shift)
    Hh, W, pad = 128, 1004 - 12, 12
    src = rng.integers(0, 256, size=(Hh, W + pad), dtype=np.uint8)
    pred = np.clip(src.astype(int) + rng.integers(-40, 41, size=src.shape), 0, 255).astype(np.uint8)
    rows = H.quant_rows([45, 120])
    ctx = api.Context(0)
    ctx.register_scans(H.vp9_tables())
    for ts in (3, 2, 1, 0):
        N = 4 << ts
        by, bx = np.meshgrid(np.arange(0, Hh - N + 1, N), np.arange(0, 256, N), indexing="ij")
        jobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
        jobs["src_offset"] = by.ravel() * src.shape[1] + bx.ravel() + shift
        jobs["pred_offset"] = jobs["src_offset"] + 3 * ((shift & 1) and 1)      # the two planes need not share an alignment
        jobs["recon_offset"] = by.ravel() * src.shape[1] + bx.ravel()
        jobs["qrow"] = np.arange(by.size) & 1
        jobs["tx_type"] = 0 if ts == 3 else np.arange(by.size) & 3
        jobs["flags"] = 1
        c1, q1, d1, e1, r1 = ctx.txq(ts, jobs, src, pred, np.zeros_like(src), rows)
        for i, j in enumerate(jobs):
            sy, sx = divmod(int(j["src_offset"]), src.shape[1])
            py, px = divmod(int(j["pred_offset"]), src.shape[1])
            ry, rx = divmod(int(j["recon_offset"]), src.shape[1])
            resid = (src[sy:sy + N, sx:sx + N].astype(np.int16) - pred[py:py + N, px:px + N].astype(np.int16))[None]
            c0 = H.oracle_fwd(ts, int(j["tx_type"]), 0, resid)
            q0, d0, e0 = H.oracle_quant(ts, int(j["tx_type"]), c0, rows, [int(j["qrow"])])
            r0 = H.oracle_inv_add(ts, int(j["tx_type"]), d0, e0, pred[py:py + N, px:px + N][None])[0]
            assert (c1[i] == c0[0]).all() and (q1[i] == q0[0]).all() and (d1[i] == d0[0]).all() and e1[i] == e0[0], (ts, i)
            assert (r1[ry:ry + N, rx:rx + N] == r0).all(), (ts, i)
    ctx.close()


@pytest.mark.gpu
def test_txq_multi_launch_equals_single_batches():
    """svtb200_txq_multi_dev: batches of all four transform sizes on different planes in ONE launch == the same batches one by one."""
    rng = np.random.default_rng(77)
    rows = H.quant_rows([30, 180])
    batches = []
    for ts, (Hh, W) in zip((3, 2, 1, 0, 2), ((160, 416), (80, 208), (40, 200), (36, 100), (16, 2000))):
        N = 4 << ts
        src = rng.integers(0, 256, size=(Hh, W), dtype=np.uint8)
        pred = np.clip(src.astype(int) + rng.integers(-25, 26, size=src.shape), 0, 255).astype(np.uint8)
        by, bx = np.meshgrid(np.arange(0, Hh - N + 1, N), np.arange(0, W - N + 1, N), indexing="ij")
        jobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
        jobs["src_offset"] = jobs["pred_offset"] = jobs["recon_offset"] = by.ravel() * W + bx.ravel()
        jobs["qrow"] = np.arange(by.size) % 2
        jobs["tx_type"] = 0 if ts == 3 else np.arange(by.size) & 3
        jobs["flags"] = 1
        batches.append((ts, jobs, src, pred, np.zeros_like(src)))
    ctx = api.Context(0)
    ctx.register_scans(H.vp9_tables())
    launches0 = ctx.launches()
    got = ctx.txq_multi(batches, rows)
    assert ctx.launches() - launches0 == 1
    for (ts, jobs, src, pred, rec), (q, dq, eob, r) in zip(batches, got):
        _, q0, d0, e0, r0 = ctx.txq(ts, jobs, src, pred, rec, rows)
        assert (q == q0).all() and (dq == d0).all() and (eob == e0).all() and (r == r0).all(), ts
    ctx.close()


def unpack_coeffs(batches, eobs, stream, offsets):
    """host-side reader of svtb200_coeff_pack_dev's stream: dense (n, N, N) qcoeff arrays per batch."""
    T = H.vp9_tables()
    out, t = [], 0
    for (ts, jobs, *_), eob in zip(batches, eobs):
        N = 4 << ts
        q = np.zeros((len(jobs), N * N), np.int16)
        for j in range(len(jobs)):
            e, o = int(eob[j]), int(offsets[t])
            assert int(offsets[t + 1]) - o == (e + 1) & ~1
            scan = T["scan_%d_%d" % (ts, 0 if ts == 3 else int(jobs["tx_type"][j]) & 3)]
            q[j, scan[:e]] = stream[o:o + e]
            if e & 1:
                assert stream[o + e] == 0
            t += 1
        out.append(q.reshape(len(jobs), N, N))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True])
def test_coeff_pack_roundtrip(host):
    """svtb200_coeff_pack_dev: scan-order prefixes of every block's coefficients, into device memory and straight into
    page-locked host memory; unpacked on the host they are the dense arrays (tokenize_b reads nothing else)."""
    rng = np.random.default_rng(78)
    rows = H.quant_rows([20, 200])
    batches = []
    for ts, (Hh, W) in zip((3, 2, 1, 0, 2, 3), ((160, 416), (80, 208), (40, 200), (36, 100), (16, 2000), (64, 64))):
        N = 4 << ts
        src = rng.integers(0, 256, size=(Hh, W), dtype=np.uint8)
        spread = 3 if ts == 3 and Hh == 64 else 25   # the last batch: (almost) empty blocks
        pred = np.clip(src.astype(int) + rng.integers(-spread, spread + 1, size=src.shape), 0, 255).astype(np.uint8)
        by, bx = np.meshgrid(np.arange(0, Hh - N + 1, N), np.arange(0, W - N + 1, N), indexing="ij")
        jobs = np.zeros(by.size, dtype=api.TX_JOB_DTYPE)
        jobs["src_offset"] = jobs["pred_offset"] = jobs["recon_offset"] = by.ravel() * W + bx.ravel()
        jobs["qrow"] = np.arange(by.size) % 2
        jobs["tx_type"] = 0 if ts == 3 else np.arange(by.size) & 3
        jobs["flags"] = 1
        batches.append((ts, jobs, src, pred, np.zeros_like(src)))
    ctx = api.Context(0)
    ctx.register_scans(H.vp9_tables())
    pack = {"host": host}
    got = ctx.txq_multi(batches, rows, pack=pack)
    eobs = [g[2] for g in got]
    n_total = sum(len(b[1]) for b in batches)
    assert int(pack["offsets"][n_total]) == sum(int(((e.astype(np.int64) + 1) & ~1).sum()) for e in eobs)
    assert any((e == 0).any() for e in eobs) and any((e & 1).any() for e in eobs)
    dense = unpack_coeffs(batches, eobs, pack["stream"], pack["offsets"])
    for (q, dq, eob, r), d in zip(got, dense):
        assert (q == d).all()
    # a stream that is too small: the total still reports the need, the blocks that fit are intact
    small = {"host": host, "capacity": int(pack["offsets"][n_total]) // 2}
    ctx.txq_multi(batches, rows, pack=small)
    assert int(small["offsets"][n_total]) == int(pack["offsets"][n_total])
    fit = np.flatnonzero(small["offsets"][1:] <= small["capacity"])
    k = int(fit[-1]) + 1
    assert (small["stream"][:int(small["offsets"][k])] == pack["stream"][:int(small["offsets"][k])]).all()
    ctx.close()
