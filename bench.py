#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200.

A "step" is one pass of the hot path over one 3840x2160 picture of BASELINE.json's configs[2] (4K 8-bit 4:2:0,
enc-mode 8): the pictures of a 16-picture hierarchical mini-GOP are processed in encode order with their real reference
structure and per-temporal-layer search parameters:
    motion estimation (2 lists)  ->  transform / quantise / reconstruct every TU once (luma 32x32, chroma 16x16,
    prediction = list-0 reference at zero motion)  ->  deblocking of the base-layer picture (enc-mode 8 filters
    temporal layer 0 only).

  value         pictures / s, planes already resident in HBM, results left in HBM (CUDA events on the context stream)
  e2e           pictures / s through the host-pointer side of the C ABI: pinned-host upload of the picture's planes,
                kernels, download of ME results, quantised coefficients + eob and (base layer) the filtered planes
  roofline      algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured HBM peak
  cpu_baseline  the same pictures through the CPU implementation on the host cores: the reference's own functions
                (oracle/_ref, kind "reference") or the oracle port (kind "port")

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

api = importlib.import_module("svt-vp9_b200.api")
frames = importlib.import_module("svt-vp9_b200.frames")
presets = importlib.import_module("svt-vp9_b200.presets")
synth = importlib.import_module("svt-vp9_b200.synth")
sharding = importlib.import_module("svt-vp9_b200.sharding")

W4K, H4K = 3840, 2160
N_SB = 60 * 34
N_FRAMES = 17
ENCODE_ORDER = [16, 8, 4, 2, 1, 3, 6, 5, 7, 12, 10, 9, 11, 14, 13, 15]
QINDEX = 180          # app default q 45 -> qindex (quantizer_to_qindex, VPX/vp9_quantize.c)
LF_LEVEL = 30
TU_ROWS, TU_COLS = H4K // 32, W4K // 32  # 67 x 120 luma 32x32 TUs; the last 16 luma / 8 chroma rows of the 2160 take smaller TUs
N_TU_Y = TU_ROWS * TU_COLS
N_TU_C = TU_ROWS * TU_COLS               # per chroma plane, 16x16
# every sample of the picture belongs to exactly one TU: (plane 0 Y / 1 U / 2 V, tx_size, TU size, TUs per row, TU rows, first sample row)
TU_PASSES = [(0, 3, 32, 120, 67, 0), (0, 2, 16, 240, 1, 2144), (1, 2, 16, 120, 67, 0), (1, 1, 8, 240, 1, 1072),
             (2, 2, 16, 120, 67, 0), (2, 1, 8, 240, 1, 1072)]
METRIC = "encoded fps (hot path: motion estimation + transform/quant/recon + deblocking, 4K enc-mode 8)"


def refs_of(n):
    if n == 16:
        return (0, 0)
    step = n & -n
    return (n - step, n + step)


def me_params(n):
    r0, r1 = refs_of(n)
    return presets.me_params_4k_m8(W4K, H4K, presets.MINIGOP16_TL[n - 1], 2, (r0, r1)), (r0, r1)


def me_bytes(P):
    """(1+L)*(W*H + W*H/16 [+ W*H/4]) + nSB*85*40"""
    wh = P.luma_width * P.luma_height
    per = wh + wh // 16 + (wh // 4 if P.enable_hme_level1 else 0)
    return (1 + P.num_lists) * per + N_SB * 85 * 40


def txq_bytes():
    """6N^2+2 per block + N^2 for the reconstruction"""
    return sum(cols * rows * (7 * size * size + 2) for _, _, size, cols, rows, _ in TU_PASSES)


def lf_bytes():
    return 2 * 3 * W4K * H4K // 2 + N_SB * 160 + 64 * 3


def make_clip(seed):
    w, h, sd, k, g, a, b, m, noise, _ = synth.CONFIGS["C"]
    return synth.make_clip(w, h, N_FRAMES, seed, k, g, a, b, m, noise)


def tu_jobs(stride_y, ox, oy, n, size, stride_out, cols=None, row0=0):
    """One job per TU on a regular grid (cols TUs per row, first sample row row0): src/pred offsets into the (padded) plane,
    recon offsets into a packed plane."""
    cols = cols or TU_COLS
    jobs = np.zeros(n, dtype=api.TX_JOB_DTYPE)
    r, c = np.divmod(np.arange(n), cols)
    jobs["src_offset"] = (oy + row0 + r * size) * stride_y + ox + c * size
    jobs["pred_offset"] = jobs["src_offset"]
    jobs["recon_offset"] = (row0 + r * size) * stride_out + c * size
    jobs["qrow"] = 0
    jobs["tx_type"] = 0
    jobs["flags"] = 1
    return jobs


def lf_masks_32x32():
    """Edge masks of a picture made of 32x32 luma / 16x16 chroma transform blocks (as eb_vp9_build_mask would set them)."""
    m = np.zeros(N_SB, dtype=api.LF_MASK_DTYPE)
    ly = ay = 0
    for q in range(4):
        qr, qc = (q >> 1) * 4, (q & 1) * 4
        for k in range(4):
            ly |= 1 << ((qr + k) * 8 + qc)
            ay |= 1 << (qr * 8 + qc + k)
    lu = au = 0
    for q in range(4):
        qr, qc = (q >> 1) * 2, (q & 1) * 2
        for k in range(2):
            lu |= 1 << ((qr + k) * 4 + qc)
            au |= 1 << (qr * 4 + qc + k)
    m["left_y"][:, 3], m["above_y"][:, 3] = ly, ay
    m["left_uv"][:, 2], m["above_uv"][:, 2] = lu, au
    m["lfl_y"][:] = LF_LEVEL
    return m


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        def num(s):
            try:
                return float(s)
            except ValueError:
                return None
        sm = [num(r[0]) for r in self.rows if len(r) >= 6 and num(r[0]) is not None]
        mx = [num(r[1]) for r in self.rows if len(r) >= 6 and num(r[1]) is not None]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own functions (oracle/_ref) or the oracle port, all host threads
# ---------------------------------------------------------------------------------------------------------------------
def aligned(shape, dtype, align=64):
    """numpy array whose data pointer is 'align'-byte aligned (the reference's SIMD kernels use aligned loads on
    prediction / coefficient buffers, as its own pooled buffers are)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def aligned_copy(a):
    b = aligned(a.shape, a.dtype)
    b[...] = a
    return b


class CpuPath:
    def __init__(self, clip, threads):
        import helpers as H
        from concurrent.futures import ThreadPoolExecutor
        self.H, self.threads = H, threads
        self.pool = ThreadPoolExecutor(threads)
        self.kind = "reference" if H.have_refdrv() else "port"
        self.pics = [frames.pa_planes(f[0]) for f in clip]
        self.structs = [frames.plane_structs(p) for p in self.pics]
        self.luma = [aligned_copy(f[0]) for f in clip]  # unpadded, aligned: source / prediction planes of the transform stage
        self.chroma = [(aligned_copy(f[1]), aligned_copy(f[2])) for f in clip]
        self.rows = H.quant_rows([QINDEX])
        self.passes = []
        for plane, ts, size, cols, rows, row0 in TU_PASSES:
            w = W4K if plane == 0 else W4K // 2
            n = cols * rows
            self.passes.append((plane, ts, size, tu_jobs(w, 0, 0, n, size, w, cols, row0), [aligned((n, size, size), np.int16) for _ in range(3)],
                                np.zeros(n, np.uint16)))
        self.lfm, self.thr = lf_masks_32x32(), H.lf_thresholds()
        self.recon = [aligned((H4K, W4K), np.uint8), aligned((H4K // 2, W4K // 2), np.uint8), aligned((H4K // 2, W4K // 2), np.uint8)]
        self.out = np.zeros(N_SB * 85, dtype=api.ME_RESULT_DTYPE)
        if self.kind == "reference":
            self.L = H.refdrv(1)
            self.me_handles = [self.L.refdrv_me_new() for _ in range(threads)]
        else:
            self.L = H.oracle()
        self.LfP = api.LfParams()
        self.LfP.width, self.LfP.height, self.LfP.mi_rows, self.LfP.mi_cols, self.LfP.y_only, self.LfP.filter_level = W4K, H4K, H4K // 8, W4K // 8, 0, LF_LEVEL

    def _split(self, n):
        chunk = (n + self.threads - 1) // self.threads
        return [(i, s, min(n, s + chunk)) for i, s in enumerate(range(0, n, chunk))]

    def step(self, n):
        H, L = self.H, self.L
        P, (r0, r1) = me_params(n)
        cur, s0, s1 = self.structs[n], self.structs[r0], self.structs[r1]
        if self.kind == "reference":
            list(self.pool.map(lambda a: L.refdrv_me_range(self.me_handles[a[0]], C.byref(P), cur, s0, s1, a[1], a[2], self.out.ctypes.data), self._split(N_SB)))
        else:
            list(self.pool.map(lambda a: L.svtb200_oracle_me_picture(C.byref(P), cur, s0, s1, self.out.ctypes.data, None, a[1], a[2]), self._split(N_SB)))
        # transform / quantise / reconstruct: every TU of the three planes
        for plane, ts, size, jobs, (c, q, dq), eob in self.passes:
            src = self.luma[n] if plane == 0 else self.chroma[n][plane - 1]
            pred = self.luma[r0] if plane == 0 else self.chroma[r0][plane - 1]
            rec = self.recon[plane]
            pl = api.TxPlanes()
            pl.src, pl.src_stride, pl.src_bytes = src.ctypes.data, src.shape[1], src.nbytes
            pl.pred, pl.pred_stride, pl.pred_bytes = pred.ctypes.data, pred.shape[1], pred.nbytes
            pl.recon, pl.recon_stride, pl.recon_bytes = rec.ctypes.data, rec.shape[1], rec.nbytes
            if self.kind == "reference":
                if len(jobs) < 2048:   # a few hundred small blocks: one call, the thread pool's dispatch would cost more than the work
                    L.refdrv_txq(ts, 0, len(jobs), jobs.ctypes.data, C.byref(pl), self.rows.ctypes.data, c.ctypes.data, q.ctypes.data, dq.ctypes.data,
                                 eob.ctypes.data)
                else:
                    list(self.pool.map(lambda a: L.refdrv_txq(ts, a[1], a[2], jobs.ctypes.data, C.byref(pl), self.rows.ctypes.data, c.ctypes.data,
                                                              q.ctypes.data, dq.ctypes.data, eob.ctypes.data), self._split(len(jobs))))
            else:
                self._port_txq(ts, jobs, src, pred, rec, c, q, dq, eob)
        if n == 16:
            if self.kind == "reference":
                L.refdrv_deblock(C.byref(self.LfP), self.recon[0].ctypes.data, W4K, self.recon[1].ctypes.data, self.recon[2].ctypes.data, W4K // 2,
                                 self.lfm.ctypes.data, 60, self.thr.ctypes.data)
            else:
                m = self.lfm.copy()
                L.svtb200_oracle_deblock_picture(C.byref(self.LfP), self.recon[0].ctypes.data, W4K, self.recon[1].ctypes.data, self.recon[2].ctypes.data,
                                                 W4K // 2, m.ctypes.data, 60, self.thr.ctypes.data)

    def _port_txq(self, ts, jobs, src, pred, rec, c, q, dq, eob):
        H = self.H
        N = 4 << ts
        scan, _ = H.scan_of(ts, 0)
        scan = np.ascontiguousarray(scan)
        Lt = H.oracle_tx()

        def work(a):
            for i in range(a[1], a[2]):
                so, ro = int(jobs["src_offset"][i]), int(jobs["recon_offset"][i])
                sy, sx = divmod(so, src.shape[1])
                resid = np.ascontiguousarray(src[sy:sy + N, sx:sx + N].astype(np.int16) - pred[sy:sy + N, sx:sx + N].astype(np.int16))
                Lt.svtb200_oracle_fwd_txfm(ts, 0, 0, resid.ctypes.data, N, c[i].ctypes.data)
                Lt.svtb200_oracle_quantize(ts, c[i].ctypes.data, self.rows.ctypes.data, scan.ctypes.data, q[i].ctypes.data, dq[i].ctypes.data, eob[i:].ctypes.data)
                ry, rx = divmod(ro, rec.shape[1])
                blk = np.ascontiguousarray(pred[sy:sy + N, sx:sx + N])
                Lt.svtb200_oracle_inv_txfm_add(ts, 0, dq[i].ctypes.data, int(eob[i]), blk.ctypes.data, N)
                rec[ry:ry + N, rx:rx + N] = blk
        list(self.pool.map(work, self._split(len(jobs))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=320)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--verify", action="store_true", help="after the end-to-end leg, recompute its last pictures synchronously on one queue and compare every downloaded byte")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    config = {"workload": "configs[2]: 4K (3840x2160) 8-bit yuv420p, enc-mode 8 (tune 1 parameter set): per picture ME (2 lists) + one "
                          "transform/quantise/reconstruct pass over every TU (luma 32x32, chroma 16x16; 16x16 / 8x8 for the last 16 rows; q 45) + deblocking on the "
                          "base-layer picture; 16-picture hierarchical mini-GOP in encode order",
              "pictures_per_step": 1,
              "queues": "motion estimation (followed by the three small edge-row transform launches) on the context's first compute queue, the "
                        "picture-sized transform launches + deblocking on its second (they belong to different pictures in the encoder); value = "
                        "CUDA events around the K steps with both queues joined inside the timed region",
              "l2": "17 distinct pictures (165 MB of luma planes + 105 MB chroma) cycled: working set larger than the 126 MB L2",
              "parallelism": "pictures sharded across GPUs (independent mini-GOP streams per rank), no collective on the data path"}

    if args.impl == "reference":
        if rank != 0:
            return
        clip = make_clip(11)
        cpu = CpuPath(clip, ncores)
        times, t_all = [], time.perf_counter()
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            cpu.step(ENCODE_ORDER[i % 16])
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > 200 and len(times) >= 3:
                break
            if time.perf_counter() - t_all > 200 and i < args.warmup:
                args.warmup = i  # bounded run: cut the warm-up short
        steps = len(times)
        v = steps / sum(times) if times else 0.0
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "fps", "n_gpus": args.gpus, "steps": steps,
                          "warmup": args.warmup, "ms_per_step": 1000.0 / v if v else None, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "fps", "cores": ncores, "kind": cpu.kind,
                                           "sample": "one whole 4K picture per step (all 2040 SBs, all TUs), mini-GOP encode order, %d host threads" % ncores},
                          "e2e": {"value": v, "unit": "fps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    import helpers as H
    ctx = api.Context(local_rank)
    ctx.register_scans(H.vp9_tables())
    L, h = ctx.L, ctx.h
    stream = torch.cuda.ExternalStream(L.svtb200_stream(h), device=torch.device("cuda", local_rank))

    clip = make_clip(11 + rank)
    pics = [frames.pa_planes(f[0]) for f in clip]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pfull = [pin(p["full"]) for p in pics]
    psix = [pin(p["sixteenth"]) for p in pics]
    pu = [pin(f[1]) for f in clip]
    pv = [pin(f[2]) for f in clip]

    def pstructs(i):
        arr = (api.Plane * 3)()
        arr[0] = api.plane_of(pfull[i].numpy(), frames.FULL_PAD, frames.FULL_PAD, W4K, H4K)
        arr[2] = api.plane_of(psix[i].numpy(), frames.SIXTEENTH_PAD, frames.SIXTEENTH_PAD, W4K // 4, H4K // 4)
        return arr
    PS = [pstructs(i) for i in range(N_FRAMES)]
    csize = (W4K // 2) * (H4K // 2)
    d_u = [L.svtb200_dev_alloc(h, csize) for _ in range(N_FRAMES)]
    d_v = [L.svtb200_dev_alloc(h, csize) for _ in range(N_FRAMES)]
    for i in range(N_FRAMES):
        ctx.picture_upload(i, PS[i][0], None, PS[i][2])
        L.svtb200_memcpy_h2d(h, d_u[i], pu[i].data_ptr(), csize)
        L.svtb200_memcpy_h2d(h, d_v[i], pv[i].data_ptr(), csize)
    dplane = []
    for i in range(N_FRAMES):
        p = api.Plane()
        ctx.check(L.svtb200_picture_plane_dev(h, i, 0, C.byref(p)), "plane_dev")
        dplane.append(p)
    sy = dplane[0].stride

    def to_dev(arr):
        d = L.svtb200_dev_alloc(h, arr.nbytes)
        L.svtb200_memcpy_h2d(h, d, arr.ctypes.data, arr.nbytes)
        return d
    # one job table per TU pass: luma addresses the padded resident plane, chroma the packed buffers
    pass_n = [cols * rows for _, _, _, cols, rows, _ in TU_PASSES]
    pass_cb = [n * size * size * 2 for n, (_, _, size, _, _, _) in zip(pass_n, TU_PASSES)]   # bytes of one coefficient array
    d_jobs = [to_dev(tu_jobs(sy, frames.FULL_PAD, frames.FULL_PAD, n, size, W4K, cols, row0) if plane == 0 else
                     tu_jobs(W4K // 2, 0, 0, n, size, W4K // 2, cols, row0))
              for n, (plane, _, size, cols, rows, row0) in zip(pass_n, TU_PASSES)]
    d_rows = to_dev(H.quant_rows([QINDEX]))
    d_lfm, d_thr = to_dev(lf_masks_32x32()), to_dev(H.lf_thresholds())
    # result buffers exist NSLOT times: the end-to-end leg downloads slot s while the kernels of the next pictures fill the others
    NSLOT = 3
    d_rec = [[L.svtb200_dev_alloc(h, W4K * H4K), L.svtb200_dev_alloc(h, csize), L.svtb200_dev_alloc(h, csize)] for _ in range(NSLOT)]
    d_me = [L.svtb200_dev_alloc(h, N_SB * 85 * 40) for _ in range(NSLOT)]
    d_q = [[L.svtb200_dev_alloc(h, b) for b in pass_cb] for _ in range(NSLOT)]
    d_dq = [L.svtb200_dev_alloc(h, b) for b in pass_cb]
    d_eob = [[L.svtb200_dev_alloc(h, n * 2) for n in pass_n] for _ in range(NSLOT)]
    LfP = api.LfParams()
    LfP.width, LfP.height, LfP.mi_rows, LfP.mi_cols, LfP.y_only, LfP.filter_level = W4K, H4K, H4K // 8, W4K // 8, 0, LF_LEVEL
    # pinned host result buffers for the e2e leg
    pinned = lambda n: torch.empty(n, dtype=torch.uint8).pin_memory()
    h_me = [pinned(N_SB * 85 * 40) for _ in range(NSLOT)]
    h_q = [[pinned(b) for b in pass_cb] for _ in range(NSLOT)]
    h_eob = [[pinned(n * 2) for n in pass_n] for _ in range(NSLOT)]
    h_rec = [[pinned(W4K * H4K), pinned(csize), pinned(csize)] for _ in range(NSLOT)]

    def txq_planes(src, sstride, sbytes, pred, rec, rstride, rbytes):
        pl = api.TxPlanes()
        pl.src, pl.src_stride, pl.src_bytes = src, sstride, sbytes
        pl.pred, pl.pred_stride, pl.pred_bytes = pred, sstride, sbytes
        pl.recon, pl.recon_stride, pl.recon_bytes = rec, rstride, rbytes
        return pl

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {"me_sb_kernel": 0.0, "txq_kernel": 0.0, "lf_row_kernel": 0.0}
    alg = {"me_sb_kernel": 0, "txq_kernel": 0, "lf_row_kernel": 0}
    marks = []

    # The transform pass and the deblocking of a picture go to the context's second compute queue: in the encoder they belong to
    # a different (earlier) picture than the one being motion-estimated, and their CTAs fill the tail of the ME grid.
    ctx.select_compute(api.COMPUTE2)
    stream2 = torch.cuda.ExternalStream(L.svtb200_stream(h), device=torch.device("cuda", local_rank))
    # Motion estimation of consecutive pictures is independent (the reference runs it several pictures ahead on separate
    # threads): with SVTB200_BENCH_ME_QUEUES=2 it alternates between the first and the third compute queue, so that the next
    # picture's CTAs could fill the tail of the current picture's grid (2040 CTAs = 3.45 waves of 4 x 148).  Measured: no gain
    # (0.403 vs 0.400 ms per step) -- with the transform queue beside it the SMs are already busy, the step is bound by the
    # instruction work of the kernels, not by their tails -- so one motion-estimation queue is the default.
    ctx.select_compute(api.COMPUTE3)
    stream3 = torch.cuda.ExternalStream(L.svtb200_stream(h), device=torch.device("cuda", local_rank))
    ctx.select_compute(api.COMPUTE)
    ME_QUEUES = int(os.environ.get("SVTB200_BENCH_ME_QUEUES", "1"))

    _txq_cache = {}
    EDGE_PASSES = tuple(k for k, (_, _, _, cols, rows, _) in enumerate(TU_PASSES) if rows == 1)

    def txq_args(slot, n, r0):
        """ctypes argument tuples of the six transform launches of picture n into result slot 'slot' (built once)"""
        key = (slot, n)
        if key not in _txq_cache:
            ybytes = sy * dplane[n].padded_height
            rec, q, eob = d_rec[slot], d_q[slot], d_eob[slot]
            lst = []
            for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
                if plane == 0:
                    pl = txq_planes(dplane[n].buf, sy, ybytes, dplane[r0].buf, rec[0], W4K, W4K * H4K)
                else:
                    dc = d_u if plane == 1 else d_v
                    pl = txq_planes(dc[n], W4K // 2, csize, dc[r0], rec[plane], W4K // 2, csize)
                lst.append((h, ts, pass_n[k], d_jobs[k], C.byref(pl), d_rows, 1, None, q[k], d_dq[k], eob[k], pl))
            _txq_cache[key] = [a[:-1] for a in lst], [a[-1] for a in lst]   # keep the structs alive
        return _txq_cache[key][0]

    def kernels(i, timed, slot=0, overlap=True, sink=None):
        n = ENCODE_ORDER[i % 16]
        s2 = stream2 if overlap else stream
        q1, s1 = (api.COMPUTE3, stream3) if overlap and ME_QUEUES == 2 and (i & 1) else (api.COMPUTE, stream)
        P, (r0, r1) = me_params(n)
        ids = (C.c_uint64 * 2)(r0, r1)
        e = [ev() for _ in range(6)] if timed else None
        ctx.select_compute(q1)
        if timed:
            e[0].record(s1)
        ctx.check(L.svtb200_me_picture_dev(h, C.byref(P), n, ids, d_me[slot], None), "me_picture_dev")
        if timed:
            e[1].record(s1)
        targs = txq_args(slot, n, r0)
        if overlap:
            # the three edge passes (240 small blocks each, latency-bound 5-7 us launches) ride in the tail of the ME grid on the
            # first queue; the picture-sized passes go to the second queue
            for k in EDGE_PASSES:
                rc = L.svtb200_txq_batch_dev(*targs[k])
                if rc:
                    ctx.check(rc, "txq_batch_dev")
            L.svtb200_picture_touch(h, n)
            L.svtb200_picture_touch(h, r0)
            if n == 16:
                edge_done = ctx.marker(q1)
            ctx.select_compute(api.COMPUTE2)
        if timed:
            e[2].record(s2)
        for k, a in enumerate(targs):
            if overlap and k in EDGE_PASSES:
                continue
            rc = L.svtb200_txq_batch_dev(*a)
            if rc:
                ctx.check(rc, "txq_batch_dev")
        if timed:
            e[3].record(s2)
        if n == 16:
            if overlap:
                ctx.stream_wait(api.COMPUTE2, edge_done)   # deblocking needs the reconstruction of the edge rows too
            rec = d_rec[slot]
            ctx.check(L.svtb200_deblock_picture_dev(h, C.byref(LfP), rec[0], W4K, rec[1], rec[2], W4K // 2, d_lfm, 60, d_thr), "deblock")
        if timed:
            e[4].record(s2)
            (marks if sink is None else sink).append((e, P, n))
        L.svtb200_picture_touch(h, n)      # the transform kernels read the planes of pictures n and r0 on this queue
        L.svtb200_picture_touch(h, r0)
        ctx.select_compute(api.COMPUTE)
        return n, r0

    # ---- end-to-end leg: host buffers in, host buffers out, through the C ABI's three queues.  Picture i+1 uploads and
    # picture i-1 downloads while picture i computes; the host collects picture i-3's results (the previous user of the slot)
    # before it queues picture i, i.e. three pictures are in flight (upload / kernels / download).
    downloaded = [0] * NSLOT   # marker after the last download out of each result slot
    chroma_used = {}      # marker after the last kernel that read the chroma buffers of a picture

    skip = os.environ.get("SVTB200_E2E_SKIP", "").split(",")  # diagnosis only: "h2d", "d2h", "kernels"

    def step_e2e(i):
        n = ENCODE_ORDER[i % 16]
        slot = i % NSLOT
        if downloaded[slot]:
            ctx.wait(downloaded[slot])                                          # picture i-NSLOT is complete in host memory; its slot is free
        if "h2d" not in skip:
            ctx.picture_upload_luma(n, PS[n][0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, False)   # H2D luma picture area; padding + 1/16 on the device
            if n in chroma_used:
                for m in chroma_used[n]:
                    ctx.stream_wait(api.UPLOAD, m)
            ctx.check(L.svtb200_memcpy_h2d_async(h, d_u[n], pu[n].data_ptr(), csize), "h2d u")     # H2D chroma
            ctx.check(L.svtb200_memcpy_h2d_async(h, d_v[n], pv[n].data_ptr(), csize), "h2d v")
            ctx.fence(api.UPLOAD, api.COMPUTE)
            ctx.fence(api.UPLOAD, api.COMPUTE2)
            ctx.fence(api.UPLOAD, api.COMPUTE3)
        if "kernels" in skip:
            r0 = n
        else:
            _, r0 = kernels(i, False, slot)
        chroma_used[n] = chroma_used[r0] = (ctx.marker(api.COMPUTE2), ctx.marker(api.COMPUTE), ctx.marker(api.COMPUTE3))   # every compute queue reads the chroma buffers
        for m in chroma_used[n]:
            ctx.stream_wait(api.DOWNLOAD, m)
        rec, q, eob = d_rec[slot], d_q[slot], d_eob[slot]
        d2h = lambda hb, db, nb: ("d2h" in skip) or ctx.check(L.svtb200_memcpy_d2h_async(h, hb.data_ptr(), db, nb), "d2h")
        d2h(h_me[slot], d_me[slot], N_SB * 85 * 40)                             # D2H ME results
        for k in range(len(TU_PASSES)):                                         # D2H quantised coefficients + eob of every TU pass
            d2h(h_q[slot][k], q[k], pass_cb[k])
            d2h(h_eob[slot][k], eob[k], pass_n[k] * 2)
        if n == 16:                                                            # D2H filtered reference picture
            d2h(h_rec[slot][0], rec[0], W4K * H4K)
            d2h(h_rec[slot][1], rec[1], csize)
            d2h(h_rec[slot][2], rec[2], csize)
        downloaded[slot] = ctx.marker(api.DOWNLOAD)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L._bound = True
    ctx.L.svtb200_deblock_picture_dev.argtypes = [C.c_void_p, C.POINTER(api.LfParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    sampler = ClockSampler(local_rank)
    for i in range(args.warmup):
        kernels(i, False, i % NSLOT)
    ctx.sync()
    barrier()
    sampler.start()
    e0, e1 = ev(), ev()
    launches0 = ctx.launches()
    e0.record(stream)
    ctx.fence(api.COMPUTE, api.COMPUTE2)   # the other queues start inside the timed region ...
    ctx.fence(api.COMPUTE, api.COMPUTE3)
    t_host = time.perf_counter()
    for i in range(args.steps):
        kernels(args.warmup + i, True, (args.warmup + i) % NSLOT)
    t_host = time.perf_counter() - t_host
    if os.environ.get("SVTB200_BENCH_DEBUG"):
        sys.stderr.write("value leg: host issue time %.3f ms/step\n" % (1000.0 * t_host / args.steps))
    ctx.fence(api.COMPUTE2, api.COMPUTE)   # ... and end inside it
    ctx.fence(api.COMPUTE3, api.COMPUTE)
    e1.record(stream)
    ctx.sync()
    barrier()
    ms_res = e0.elapsed_time(e1)
    launches = ctx.launches() - launches0
    n_me = 0
    for e, P, n in marks:
        stage_ms["me_sb_kernel"] += e[0].elapsed_time(e[1]); alg["me_sb_kernel"] += me_bytes(P); n_me += 1
        stage_ms["txq_kernel"] += e[2].elapsed_time(e[3]); alg["txq_kernel"] += txq_bytes()
        if n == 16:
            stage_ms["lf_row_kernel"] += e[3].elapsed_time(e[4]); alg["lf_row_kernel"] += lf_bytes()
    # the same kernels one at a time (single queue): durations without the other queue's interference
    iso_marks, iso_ms = [], {"me_sb_kernel": 0.0, "txq_kernel": 0.0, "lf_row_kernel": 0.0}
    iso_n = {"me_sb_kernel": 0, "txq_kernel": 0, "lf_row_kernel": 0}
    for i in range(16):
        kernels(i, True, 0, overlap=False, sink=iso_marks)
    ctx.sync()
    for e, P, n in iso_marks:
        iso_ms["me_sb_kernel"] += e[0].elapsed_time(e[1]); iso_n["me_sb_kernel"] += 1
        iso_ms["txq_kernel"] += e[2].elapsed_time(e[3]); iso_n["txq_kernel"] += 1
        if n == 16:
            iso_ms["lf_row_kernel"] += e[3].elapsed_time(e[4]); iso_n["lf_row_kernel"] += 1
    for i in range(min(args.warmup, 4)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    t_issue = 0.0
    for i in range(args.steps):
        ti = time.perf_counter()
        step_e2e(args.warmup + i)
        t_issue += time.perf_counter() - ti
    ctx.sync()
    if os.environ.get("SVTB200_BENCH_DEBUG"):
        sys.stderr.write("e2e host time in step_e2e (issue + waits): %.3f ms/step\n" % (1000.0 * t_issue / args.steps))
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1000.0
    sampler.stop()

    if args.verify:
        # the pipelined leg (4 queues, 3 slots) must have produced exactly what a synchronous single-queue run produces
        checked = 0
        for i in range(args.warmup + args.steps - NSLOT, args.warmup + args.steps):
            slot, n = i % NSLOT, ENCODE_ORDER[i % 16]
            got = [h_me[slot].numpy().copy()] + [h_q[slot][k].numpy().copy() for k in range(len(TU_PASSES))]
            got += [h_eob[slot][k].numpy().copy() for k in range(len(TU_PASSES))]
            if n == 16:
                got += [h_rec[slot][k].numpy().copy() for k in range(3)]
            kernels(i, False, slot, overlap=False)
            ctx.sync()
            devs = [(d_me[slot], N_SB * 85 * 40)] + [(d_q[slot][k], pass_cb[k]) for k in range(len(TU_PASSES))]
            devs += [(d_eob[slot][k], pass_n[k] * 2) for k in range(len(TU_PASSES))]
            if n == 16:
                devs += [(d_rec[slot][0], W4K * H4K), (d_rec[slot][1], csize), (d_rec[slot][2], csize)]
            for g, (dptr, nbytes) in zip(got, devs):
                want = np.zeros(nbytes, dtype=np.uint8)
                ctx.check(L.svtb200_memcpy_d2h(h, want.ctypes.data, dptr, nbytes), "verify d2h")
                if not np.array_equal(g, want):
                    sys.stderr.write("verify: MISMATCH at step %d (picture %d), buffer of %d bytes\n" % (i, n, nbytes))
                    raise SystemExit(2)
                checked += nbytes
        sys.stderr.write("verify: ok (%d result bytes of the last %d pipelined pictures equal the synchronous run)\n" % (checked, NSLOT))

    # every rank processed its own pictures (svt-vp9_b200/sharding.py): time = max over ranks, units = sum over ranks
    ms_res, units = sharding.agree_step(ms_res, args.steps, device="cuda")
    ms_e2e, _ = sharding.agree_step(ms_e2e, args.steps, device="cuda")
    value = units / (ms_res / 1000.0)
    e2e = units / (ms_e2e / 1000.0)
    h2d = W4K * H4K + 2 * csize  # the picture area of the three planes (padding and the 1/16 plane are made on the device)
    d2h = N_SB * 85 * 40 + sum(pass_cb) + 2 * sum(pass_n) + (W4K * H4K * 3 // 2) // 16  # filtered planes on 1 of 16 pictures

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
    except Exception:
        pass
    # the dominant kernel is the one with the most work per step when run alone: inside the timed region the two queues
    # overlap, which stretches the event-to-event time of whichever kernel shares the SMs (ncu, which serialises the
    # launches, agrees with "share_alone")
    per_step_alone = {k: (iso_ms[k] / max(1, iso_n[k])) * (1.0 / 16 if k == "lf_row_kernel" else 1.0) for k in stage_ms}
    dom = max(per_step_alone, key=lambda k: per_step_alone[k])
    per_kernel = {}
    for k in stage_ms:
        if stage_ms[k] > 0:
            a = alg[k] / (stage_ms[k] / 1000.0) / 1e9
            per_pic = alg[k] / {"me_sb_kernel": n_me, "txq_kernel": n_me, "lf_row_kernel": max(1, sum(1 for m in marks if m[2] == 16))}[k]
            iso = iso_ms[k] / max(1, iso_n[k])
            per_kernel[k] = {"ms_total": stage_ms[k], "share": stage_ms[k] / sum(stage_ms.values()), "achieved_gbs": a, "frac": a / peak,
                             "traffic": traffic.get(k), "ms_per_picture_alone": iso, "achieved_gbs_alone": per_pic / (iso / 1000.0) / 1e9 if iso else None,
                             "frac_alone": per_pic / (iso / 1000.0) / 1e9 / peak if iso else None,
                             "share_alone": per_step_alone[k] / sum(per_step_alone.values())}
    launches_dom = n_me if dom == "me_sb_kernel" else (len(TU_PASSES) * n_me if dom == "txq_kernel" else max(1, args.steps // 16))
    roof = {"bound": "hbm", "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per_kernel[dom]["frac"],
            "traffic": traffic.get(dom), "kernel": dom, "alg_bytes_per_launch": alg[dom] / launches_dom,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)", "per_kernel": per_kernel}

    if rank == 0:
        cpu_line = None
        if world == 1:
            cpu = CpuPath(clip, ncores)
            sample = [16, 8, 4, 2, 1, 3]
            for n in sample[:2]:          # thread pool / page-fault warm-up, untimed
                cpu.step(n)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                for n in sample:
                    cpu.step(n)
            dt = time.perf_counter() - t0
            cpu_line = {"value": reps * len(sample) / dt, "unit": "fps", "cores": ncores, "kind": cpu.kind,
                        "sample": "%d whole 4K pictures (3 x temporal layers 0,1,2,3,4,4) through ME + transform/quant/recon (+ deblock on the "
                                  "base-layer picture), %d host threads, %.1f s" % (reps * len(sample), ncores, dt)}
        print(json.dumps({"metric": METRIC, "value": value, "unit": "fps", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic", "config": config, "clocks": sampler.summary(),
                          "e2e": {"value": e2e, "unit": "fps", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                          "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu_line}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
