"""Device times of the transform stages on all 32x32 TUs of a 4K luma plane: fused kernel vs forward / quantise / inverse alone."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
api = H.api
ctx = api.Context(0); L, h = ctx.L, ctx.h
ctx.register_scans(H.vp9_tables())
W, Hh, N = 3840, 2144, 32
n = (W // N) * (Hh // N)
rng = np.random.default_rng(1)
src = rng.integers(0, 256, size=(Hh, W), dtype=np.uint8)
pred = np.clip(src.astype(int) + rng.integers(-30, 31, size=src.shape), 0, 255).astype(np.uint8)
def dev(a):
    d = L.svtb200_dev_alloc(h, a.nbytes); L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes); return d
d_src, d_pred, d_rec = dev(src), dev(pred), L.svtb200_dev_alloc(h, src.nbytes)
by, bx = np.meshgrid(np.arange(0, Hh, N), np.arange(0, W, N), indexing="ij")
jobs = np.zeros(n, dtype=api.TX_JOB_DTYPE)
jobs["src_offset"] = jobs["pred_offset"] = jobs["recon_offset"] = by.ravel() * W + bx.ravel(); jobs["flags"] = 1
d_jobs, d_rows = dev(jobs), dev(H.quant_rows([180]))
d_c, d_q, d_dq, d_e = (L.svtb200_dev_alloc(h, n * 2048) for _ in range(3)), None, None, None
d_c = list(d_c); d_coef, d_q, d_dq = d_c; d_e = L.svtb200_dev_alloc(h, n * 2)
resid = (src.astype(np.int16) - pred.astype(np.int16)).reshape(Hh // N, N, W // N, N).transpose(0, 2, 1, 3).copy()
d_res = dev(resid)
pl = api.TxPlanes(); pl.src, pl.src_stride, pl.src_bytes = d_src, W, src.nbytes; pl.pred, pl.pred_stride, pl.pred_bytes = d_pred, W, src.nbytes
pl.recon, pl.recon_stride, pl.recon_bytes = d_rec, W, src.nbytes
stream = torch.cuda.ExternalStream(L.svtb200_stream(h))
def timed(name, f, reps=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): f()
    e1.record(stream); ctx.sync()
    print("%-28s %7.1f us" % (name, 1000.0 * e0.elapsed_time(e1) / reps))
timed("fused (coeff not stored)", lambda: L.svtb200_txq_batch_dev(h, 3, n, d_jobs, C.byref(pl), d_rows, 1, None, d_q, d_dq, d_e))
jobs2 = jobs.copy(); jobs2["flags"] = 0; d_jobs2 = dev(jobs2)
timed("fused, no reconstruction", lambda: L.svtb200_txq_batch_dev(h, 3, n, d_jobs2, C.byref(pl), d_rows, 1, None, d_q, d_dq, d_e))
timed("forward only", lambda: L.svtb200_fwd_txfm_batch_dev(h, 3, 0, 0, n, d_res, d_coef))
timed("quantise only", lambda: L.svtb200_quantize_batch_dev(h, 3, 0, n, d_coef, d_rows, 1, None, d_q, d_dq, d_e))
timed("inverse + add only", lambda: L.svtb200_inv_txfm_add_batch_dev(h, 3, 0, n, d_dq, d_e, d_rec))
