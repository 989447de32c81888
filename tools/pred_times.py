"""Device time of svtb200_inter_predict_batch on one 4K picture's worth of blocks (luma 32x32 + both chroma 16x16), by
filter case: copy, horizontal, vertical, both, both + compound.  Algorithmic bytes: every predicted sample is read once
from the reference (per reference) and written once."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
api = H.api
ctx = api.Context(0); L, h = ctx.L, ctx.h
stream = torch.cuda.ExternalStream(L.svtb200_stream(h))
W, HH, PAD = 3840, 2160, 80
rng = np.random.default_rng(1)
def plane(w, hh):
    a = rng.integers(0, 256, size=(hh + 2 * PAD, w + 2 * PAD), dtype=np.uint8)
    d = L.svtb200_dev_alloc(h, a.nbytes); L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes); return d, w + 2 * PAD
refs = {0: [plane(W, HH), plane(W, HH)], 1: [plane(W // 2, HH // 2), plane(W // 2, HH // 2)], 2: [plane(W // 2, HH // 2), plane(W // 2, HH // 2)]}
dst = {0: L.svtb200_dev_alloc(h, W * HH), 1: L.svtb200_dev_alloc(h, W * HH // 4), 2: L.svtb200_dev_alloc(h, W * HH // 4)}
def batch(sx, sy, compound):
    blocks = []
    for pl, (w, hh, b) in enumerate(((W, HH, 32), (W // 2, HH // 2, 16), (W // 2, HH // 2, 16))):
        for by in range(0, hh - b + 1, b):
            for bx in range(0, w - b + 1, b):
                blocks.append((pl, bx, by, b, w))
    arr = (api.PredBlock * len(blocks))()
    mv = rng.integers(-48, 49, size=(len(blocks), 2, 2))
    for d, (pl, bx, by, b, w), m in zip(arr, blocks, mv):
        for r in range(2):
            base, stride = refs[pl][r]
            d.ref[r] = base + PAD * stride + PAD; d.ref_stride[r] = stride
            d.x[r], d.y[r] = bx + int(m[r][0]), by + int(m[r][1])
            d.subpel_x[r], d.subpel_y[r] = sx, sy
        d.dst = dst[pl] + by * w + bx; d.dst_stride = w; d.w = d.h = b; d.compound = compound
    return arr
samples = W * HH * 3 // 2
for name, sx, sy, comp in (("copy", 0, 0, 0), ("horizontal", 6, 0, 0), ("vertical", 0, 10, 0), ("both", 6, 10, 0), ("both, compound", 6, 10, 1)):
    arr = batch(sx, sy, comp)
    for _ in range(3): ctx.inter_predict(arr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): ctx.inter_predict(arr)
    e1.record(stream); ctx.sync()
    t = 100.0 * e0.elapsed_time(e1)
    nb = C.sizeof(arr)
    d_arr = L.svtb200_dev_alloc(h, nb); L.svtb200_memcpy_h2d(h, d_arr, C.addressof(arr), nb)
    for _ in range(3): L.svtb200_inter_predict_batch_dev(h, d_arr, len(arr))
    e0.record(stream)
    for _ in range(10): L.svtb200_inter_predict_batch_dev(h, d_arr, len(arr))
    e1.record(stream); ctx.sync()
    td = 100.0 * e0.elapsed_time(e1)
    L.svtb200_dev_free(h, d_arr)
    alg = samples * (2 + comp)
    print("%-16s %6d blocks  host descriptors %7.1f us per picture   descriptors on the device %7.1f us  %7.1f GB/s algorithmic" % (name, len(arr), t, td, alg / td / 1e3))

# ---- intra predictors: every 32x32 luma block and 16x16 chroma block of a 4K picture, one mode per run
n32, n16 = (W // 32) * (HH // 32), 2 * (W // 32) * (HH // 32)
pool = rng.integers(0, 256, size=(n32 + n16) * 128, dtype=np.uint8)
d_pool = L.svtb200_dev_alloc(h, pool.nbytes); L.svtb200_memcpy_h2d(h, d_pool, pool.ctypes.data, pool.nbytes)
d_out = L.svtb200_dev_alloc(h, W * HH * 3 // 2)
for mode, name in ((0, "DC"), (1, "V"), (9, "TM"), (3, "D45"), (4, "D135"), (5, "D117"), (7, "D207")):
    arr = (api.IntraBlock * (n32 + n16))()
    off = 0
    for k, d in enumerate(arr):
        ts = 3 if k < n32 else 2
        d.above, d.left, d.dst, d.dst_stride = d_pool + k * 128 + 48, d_pool + k * 128, d_out + off, 4 << ts
        d.mode, d.tx_size, d.have_left, d.have_top = mode, ts, 1, 1
        off += (4 << ts) ** 2
    for _ in range(3): ctx.intra_predict(arr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): ctx.intra_predict(arr)
    e1.record(stream); ctx.sync()
    t = 100.0 * e0.elapsed_time(e1)
    print("intra %-5s %6d blocks  %7.1f us per picture (host descriptors)  %7.1f GB/s written" % (name, len(arr), t, samples / t / 1e3))
