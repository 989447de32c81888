"""Per-step SM-clock stamps of the deblocking kernel on a 4K picture (diagnosis of the wavefront)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
import bench
api = H.api
ctx = api.Context(0); L, h = ctx.L, ctx.h
L.svtb200_deblock_picture_dev.argtypes = [C.c_void_p, C.POINTER(api.LfParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
L.svtb200_lf_set_trace.argtypes = [C.c_void_p]
def dev(a):
    d = L.svtb200_dev_alloc(h, a.nbytes); L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes); return d
rng = np.random.default_rng(3)
w, hh = 3840, 2160
sbs_w, sbs_h = 60, 34
y = rng.integers(90, 110, size=(hh, w), dtype=np.uint8); u = rng.integers(90, 110, size=(hh // 2, w // 2), dtype=np.uint8)
dy, du, dv, dm, thr = dev(y), dev(u), dev(u), dev(bench.lf_masks_32x32()), dev(H.lf_thresholds())
P = api.LfParams(); P.width, P.height, P.mi_rows, P.mi_cols, P.y_only, P.filter_level = w, hh, hh // 8, w // 8, 0, 30
f = lambda: L.svtb200_deblock_picture_dev(h, C.byref(P), dy, w, du, dv, w // 2, dm, sbs_w, thr)
for _ in range(3): f()
ctx.sync()
tr = np.zeros((sbs_h, sbs_w, 8), dtype=np.uint64)
d_tr = dev(tr)
L.svtb200_lf_set_trace(d_tr)
f(); ctx.sync()
L.svtb200_lf_set_trace(None)
L.svtb200_memcpy_d2h(h, tr.ctypes.data, d_tr, tr.nbytes)
t = tr.astype(np.int64)
names = ["V phase (filters)", "S1 wait (filters)", "H phase (filters)", "step total (filters)", "movers: issue+put_body", "movers: poll wait", "movers: strip fetch", "movers: S1 -> publish"]
for rows, label in (([0], "row 0"), ([1, 2, 3], "rows 1-3"), (list(range(16, 20)), "rows 16-19"), ([33], "row 33")):
    a = t[rows][:, 5:55]
    nxt = t[rows][:, 6:56, 0]
    vals = [a[..., 1] - a[..., 0], a[..., 2] - a[..., 1], a[..., 3] - a[..., 2], nxt - a[..., 0], a[..., 4] - a[..., 0], a[..., 5] - a[..., 4], a[..., 6] - a[..., 5], a[..., 7] - a[..., 2]]
    print(label + ": " + "; ".join("%s %.0f" % (n, v.mean()) for n, v in zip(names, vals)) + "  [SM clocks, ~1.9 GHz]")
# start of every row relative to the previous row's start, and where a row stands when the row below starts its first tile
t0 = t[:, 0, 0].astype(np.float64)
print("clock stamps are per SM: only durations inside a row are comparable; per-row total (first stamp of SB 0 -> S1 of the last SB):",
      ["%.0f" % ((t[r, -1, 2] - t[r, 0, 0]) / 1.9e3) for r in (0, 1, 16, 33)], "us")
for rows, label in (([1], "row 1"), ([16], "row 16")):
    a = t[rows][0]
    print(label, "first 6 steps, step time us:", ["%.1f" % ((a[c + 1, 0] - a[c, 0]) / 1.9e3) for c in range(6)],
          " poll wait us:", ["%.1f" % ((a[c, 5] - a[c, 4]) / 1.9e3) for c in range(6)])
