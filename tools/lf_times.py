"""Device time of the deblocking kernel on pictures of 1 super-block row, 1 super-block column and full 4K: separates the
per-SB step time from the row-to-row lag of the wavefront."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
import bench
api = H.api
ctx = api.Context(0); L, h = ctx.L, ctx.h
ctx.L.svtb200_deblock_picture_dev.argtypes = [C.c_void_p, C.POINTER(api.LfParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
stream = torch.cuda.ExternalStream(L.svtb200_stream(h))
def dev(a):
    d = L.svtb200_dev_alloc(h, a.nbytes); L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes); return d
rng = np.random.default_rng(3)
full = bench.lf_masks_32x32()
thr = dev(H.lf_thresholds())
for (w, hh) in ((3840, 64), (64, 2176), (1920, 1088), (3840, 2160)):
    sbs_w, sbs_h = (w + 63) // 64, (hh + 63) // 64
    m = np.ascontiguousarray(full.reshape(-1)[: sbs_w * sbs_h]) if full.size >= sbs_w * sbs_h else full
    y = rng.integers(90, 110, size=(hh, w), dtype=np.uint8); u = rng.integers(90, 110, size=(hh // 2, w // 2), dtype=np.uint8)
    dy, du, dv, dm = dev(y), dev(u), dev(u), dev(m)
    P = api.LfParams(); P.width, P.height, P.mi_rows, P.mi_cols, P.y_only, P.filter_level = w, hh, hh // 8, w // 8, 0, 30
    f = lambda: L.svtb200_deblock_picture_dev(h, C.byref(P), dy, w, du, dv, w // 2, dm, sbs_w, thr)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): f()
    e1.record(stream); ctx.sync()
    t = 100.0 * e0.elapsed_time(e1)
    print("%4dx%4d  %3d x %3d SBs  %8.1f us   %.2f us per (cols + 2*(rows-1)) step" % (w, hh, sbs_w, sbs_h, t, t / (sbs_w + 2 * (sbs_h - 1))))
