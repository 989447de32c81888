#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200.

A "step" is one pass of the hot path over one 16-picture hierarchical mini-GOP of 3840x2160 pictures of BASELINE.json's
configs[2] (4K 8-bit 4:2:0, enc-mode 8), processed in encode order with the real reference structure and
per-temporal-layer search parameters; per picture:
    motion estimation (2 lists)  ->  transform / quantise / reconstruct every TU once (luma 32x32, chroma 16x16,
    prediction = list-0 reference at zero motion)  ->  deblocking of the base-layer picture (enc-mode 8 filters
    temporal layer 0 only).

  value         pictures / s, planes already resident in HBM, results left in HBM (CUDA events on the context stream)
  e2e           pictures / s through the host-pointer side of the C ABI: pinned-host upload of the picture's planes,
                kernels, download of ME results, quantised coefficients + eob and (base layer) the filtered planes
  roofline      algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured HBM peak
  cpu_baseline  the same pictures through the CPU implementation on the host cores: the reference's own functions
                (oracle/_ref, kind "reference") on a native pinned thread pool, or the oracle port (kind "port")
  check         two pictures of the end-to-end leg (one of them deblocked) compared, buffer by buffer, with the CPU arm
  encoder       BASELINE.json's own metric: the reference application against the same objects with the GPU hot path
                bound in (oracle/_ref/SvtVp9EncApp_ref vs SvtVp9EncApp_b200): fps as the application reports it
                (first send_picture -> last packet, Source/App/EbAppProcessCmd.c:669-672), bitstream md5s

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-encoder] [--no-cpu]
"""
import argparse
import ctypes as C
import hashlib
import importlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
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
PICS_PER_STEP = 16
ENCODE_ORDER = [16, 8, 4, 2, 1, 3, 6, 5, 7, 12, 10, 9, 11, 14, 13, 15]
QINDEX = 180          # app default q 45 -> qindex (quantizer_to_qindex, VPX/vp9_quantize.c)
LF_LEVEL = 30
TU_ROWS, TU_COLS = H4K // 32, W4K // 32  # 67 x 120 luma 32x32 TUs; the last 16 luma / 8 chroma rows of the 2160 take smaller TUs
# every sample of the picture belongs to exactly one TU: (plane 0 Y / 1 U / 2 V, tx_size, TU size, TUs per row, TU rows, first sample row)
TU_PASSES = [(0, 3, 32, 120, 67, 0), (0, 2, 16, 240, 1, 2144), (1, 2, 16, 120, 67, 0), (1, 1, 8, 240, 1, 1072),
             (2, 2, 16, 120, 67, 0), (2, 1, 8, 240, 1, 1072)]
METRIC = "pictures/s through the encode hot path (motion estimation + transform/quant/recon + deblocking; 4K, enc-mode 8 parameter set)"


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
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
# CPU arm: the reference's own functions (oracle/_ref) on a native thread pool, or the oracle port
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
    """The bench workload on the host cores.  kind "reference": one call per mini-GOP into oracle/ref_path_driver.c's pinned
    thread pool (super-block / transform-block granular tasks of all 16 pictures in one queue, deblocking as one task per
    picture, as the reference runs it) -- every picture has its own result buffers, so the pictures of a mini-GOP overlap
    freely.  kind "port" (no oracle/_ref): the oracle restatement, picture by picture, from a Python thread pool."""

    def __init__(self, clip, threads, n_buffers=PICS_PER_STEP):
        import helpers as H
        self.H, self.threads = H, threads
        self.kind = "reference" if H.have_refdrv() else "port"
        self.pics = [frames.pa_planes(f[0]) for f in clip]
        self.structs = [frames.plane_structs(p) for p in self.pics]
        self.luma = [aligned_copy(f[0]) for f in clip]  # unpadded, aligned: source / prediction planes of the transform stage
        self.chroma = [(aligned_copy(f[1]), aligned_copy(f[2])) for f in clip]
        self.rows = H.quant_rows([QINDEX])
        self.lfm, self.thr = lf_masks_32x32(), H.lf_thresholds()
        self.LfP = api.LfParams()
        self.LfP.width, self.LfP.height, self.LfP.mi_rows, self.LfP.mi_cols, self.LfP.y_only, self.LfP.filter_level = W4K, H4K, H4K // 8, W4K // 8, 0, LF_LEVEL
        self.jobs = []
        for plane, ts, size, cols, rows, row0 in TU_PASSES:
            w = W4K if plane == 0 else W4K // 2
            self.jobs.append(tu_jobs(w, 0, 0, cols * rows, size, w, cols, row0))
        # result buffers of picture slot b
        self.n_buffers = n_buffers
        self.me_out = [np.zeros(N_SB * 85, dtype=api.ME_RESULT_DTYPE) for _ in range(n_buffers)]
        self.q = [[aligned((cols * rows, size, size), np.int16) for _, _, size, cols, rows, _ in TU_PASSES] for _ in range(n_buffers)]
        self.dq = [[aligned((cols * rows, size, size), np.int16) for _, _, size, cols, rows, _ in TU_PASSES] for _ in range(n_buffers)]
        self.eob = [[np.zeros(cols * rows, np.uint16) for _, _, _, cols, rows, _ in TU_PASSES] for _ in range(n_buffers)]
        self.recon = [[aligned((H4K, W4K), np.uint8), aligned((H4K // 2, W4K // 2), np.uint8), aligned((H4K // 2, W4K // 2), np.uint8)]
                      for _ in range(n_buffers)]
        self.stage_cpu_s = [0.0, 0.0, 0.0]
        if self.kind == "reference":
            self.L = H.refdrv(1)
            self.pool = self.L.refdrv_pool_new(threads, 1)
            if not self.pool:
                raise RuntimeError("refdrv_pool_new failed")
        else:
            from concurrent.futures import ThreadPoolExecutor
            self.L = H.oracle()
            self.tpool = ThreadPoolExecutor(threads)

    def close(self):
        if self.kind == "reference" and self.pool:
            self.L.refdrv_pool_free(self.pool)
            self.pool = None

    def _picture_desc(self, d, n, b):
        """fill RefdrvPicture d: picture n (display offset in the mini-GOP) into result slot b"""
        P, (r0, r1) = me_params(n)
        d.P = P
        d.cur, d.ref0, d.ref1 = self.structs[n], self.structs[r0], self.structs[r1]
        d.me_out = self.me_out[b].ctypes.data
        d.n_pass = len(TU_PASSES)
        for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
            src = self.luma[n] if plane == 0 else self.chroma[n][plane - 1]
            pred = self.luma[r0] if plane == 0 else self.chroma[r0][plane - 1]
            rec = self.recon[b][plane]
            tp = d.passes[k]
            tp.tx_size, tp.n, tp.jobs = ts, cols * rows, self.jobs[k].ctypes.data
            tp.pl.src, tp.pl.src_stride, tp.pl.src_bytes = src.ctypes.data, src.shape[1], src.nbytes
            tp.pl.pred, tp.pl.pred_stride, tp.pl.pred_bytes = pred.ctypes.data, pred.shape[1], pred.nbytes
            tp.pl.recon, tp.pl.recon_stride, tp.pl.recon_bytes = rec.ctypes.data, rec.shape[1], rec.nbytes
            tp.coeff, tp.qcoeff, tp.dqcoeff, tp.eob = None, self.q[b][k].ctypes.data, self.dq[b][k].ctypes.data, self.eob[b][k].ctypes.data
        d.qrows = self.rows.ctypes.data
        d.deblock = 1 if n == 16 else 0
        d.lf = self.LfP
        d.y, d.u, d.v = (r.ctypes.data for r in self.recon[b])
        d.ys, d.cs = W4K, W4K // 2
        d.lfm, d.lfm_stride, d.thr = self.lfm.ctypes.data, 60, self.thr.ctypes.data

    def run(self, order):
        """the pictures of 'order' (display offsets, encode order), picture j into result slot j"""
        assert len(order) <= self.n_buffers
        if self.kind != "reference":
            for b, n in enumerate(order):
                self._port_step(n, b)
            return
        H = self.H
        arr = (H.RefdrvPicture * len(order))()
        for b, n in enumerate(order):
            self._picture_desc(arr[b], n, b)
        cpu_s = (C.c_double * 3)()
        rc = self.L.refdrv_pool_run(self.pool, arr, len(order), 4, 64, cpu_s)
        if rc:
            raise RuntimeError("refdrv_pool_run: %d" % rc)
        for k in range(3):
            self.stage_cpu_s[k] += cpu_s[k]

    def minigop(self):
        self.run(ENCODE_ORDER)

    def _split(self, n):
        chunk = (n + self.threads - 1) // self.threads
        return [(i, s, min(n, s + chunk)) for i, s in enumerate(range(0, n, chunk))]

    def _port_step(self, n, b):
        H, L = self.H, self.L
        P, (r0, r1) = me_params(n)
        cur, s0, s1 = self.structs[n], self.structs[r0], self.structs[r1]
        out = self.me_out[b]
        list(self.tpool.map(lambda a: L.svtb200_oracle_me_picture(C.byref(P), cur, s0, s1, out.ctypes.data, None, a[1], a[2]), self._split(N_SB)))
        Lt = H.oracle_tx()
        for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
            src = self.luma[n] if plane == 0 else self.chroma[n][plane - 1]
            pred = self.luma[r0] if plane == 0 else self.chroma[r0][plane - 1]
            rec, jobs, q, dq, eob = self.recon[b][plane], self.jobs[k], self.q[b][k], self.dq[b][k], self.eob[b][k]
            N = size
            scan = np.ascontiguousarray(H.scan_of(ts, 0)[0])

            def work(a):
                c = np.zeros((N, N), np.int16)
                for i in range(a[1], a[2]):
                    so, ro = int(jobs["src_offset"][i]), int(jobs["recon_offset"][i])
                    sy, sx = divmod(so, src.shape[1])
                    resid = np.ascontiguousarray(src[sy:sy + N, sx:sx + N].astype(np.int16) - pred[sy:sy + N, sx:sx + N].astype(np.int16))
                    Lt.svtb200_oracle_fwd_txfm(ts, 0, 0, resid.ctypes.data, N, c.ctypes.data)
                    Lt.svtb200_oracle_quantize(ts, c.ctypes.data, self.rows.ctypes.data, scan.ctypes.data, q[i].ctypes.data, dq[i].ctypes.data, eob[i:].ctypes.data)
                    ry, rx = divmod(ro, rec.shape[1])
                    blk = np.ascontiguousarray(pred[sy:sy + N, sx:sx + N])
                    Lt.svtb200_oracle_inv_txfm_add(ts, 0, dq[i].ctypes.data, int(eob[i]), blk.ctypes.data, N)
                    rec[ry:ry + N, rx:rx + N] = blk
            list(self.tpool.map(work, self._split(len(jobs))))
        if n == 16:
            m = self.lfm.copy()
            L.svtb200_oracle_deblock_picture(C.byref(self.LfP), self.recon[b][0].ctypes.data, W4K, self.recon[b][1].ctypes.data, self.recon[b][2].ctypes.data,
                                             W4K // 2, m.ctypes.data, 60, self.thr.ctypes.data)

    def stage_share(self):
        t = sum(self.stage_cpu_s)
        return None if t <= 0 else {"motion_estimation": self.stage_cpu_s[0] / t, "transform_chain": self.stage_cpu_s[1] / t, "deblocking": self.stage_cpu_s[2] / t}


# ---------------------------------------------------------------------------------------------------------------------
# whole-encoder record: BASELINE.json's metric through the reference's own application
# ---------------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
APP_REF, APP_B200 = os.path.join(REF_DIR, "SvtVp9EncApp_ref"), os.path.join(REF_DIR, "SvtVp9EncApp_b200")


def run_encoder(app, yuv, out, w, h, n_frames, n_buffered, mode, tune, env_extra=None, timeout=600):
    """One run of the reference application (or of the same objects with the GPU path bound in).  Frames are preloaded into
    RAM by the application itself (-nb), fps is its own "Average Speed" (frames / time from the first send_picture to the last
    packet); stdout is unbuffered because at 3840x2160 the reference aborts inside its own teardown after the summary."""
    import helpers as H
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(p for p in (os.path.dirname(H.cudart_path()), env.get("LD_LIBRARY_PATH", "")) if p)
    env.update(env_extra or {})
    cmd = [app, "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n_frames), "-nb", str(n_buffered), "-enc-mode", str(mode), "-tune", str(tune),
           "-q", "45", "-fps", "60", "-b", out]
    if shutil.which("stdbuf"):
        cmd = ["stdbuf", "-o0"] + cmd
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
        txt, rc = r.stdout.decode(errors="replace"), r.returncode
    except subprocess.TimeoutExpired as e:
        txt, rc = (e.stdout or b"").decode(errors="replace"), -9
    wall = time.perf_counter() - t0
    m = re.search(r"Average Speed:\s*([0-9.]+)\s*fps", txt)
    m2 = re.search(r"Total Encoding Time:\s*([0-9.]+)\s*ms", txt)
    md5 = hashlib.md5(open(out, "rb").read()).hexdigest() if os.path.exists(out) and os.path.getsize(out) else None
    return {"fps": float(m.group(1)) if m else None, "encode_ms": float(m2.group(1)) if m2 else None, "wall_s": round(wall, 2), "rc": rc, "md5": md5,
            "tail": None if m else txt[-300:]}


def gpu_report(path):
    try:
        return dict((l.split()[0], int(l.split()[1])) for l in open(path) if len(l.split()) == 2)
    except Exception:
        return None


def encoder_record(local_rank, world, clip4k, budget_s=120.0):
    """configs[1] (1080p, enc-mode 8, 300 frames) and configs[2] (4K, enc-mode 8; a bounded number of frames) through the
    application, reference build against drop-in build, one process per rank when world > 1 (configs[4]: concurrent streams,
    one per GPU).  Returns the per-rank record; the caller aggregates over ranks."""
    if not (os.path.exists(APP_REF) and os.path.exists(APP_B200)):
        return {"unavailable": "oracle/_ref applications not built"}
    t_start = time.perf_counter()
    d = tempfile.mkdtemp(prefix="svtb200_enc_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    rec = {"host_threads": len(os.sched_getaffinity(0))}
    try:
        w, h, sd, k, g, a, b, m, noise, distinct = synth.CONFIGS["B"]
        clipb = synth.make_clip(w, h, distinct, sd + local_rank, k, g, a, b, m, noise)
        yuv_b = os.path.join(d, "b.yuv")
        synth.write_yuv(yuv_b, clipb)
        del clipb
        cases = [("1080p_m8", yuv_b, 1920, 1080, 300, 40)]
        if world == 1 and clip4k is not None:
            yuv_c = os.path.join(d, "c.yuv")
            synth.write_yuv(yuv_c, clip4k[:16])
            cases.append(("4k_m8", yuv_c, 3840, 2160, int(os.environ.get("SVTB200_BENCH_4K_FRAMES", "100")), 16))
        for name, yuv, cw, ch, n_frames, n_buf in cases:
            if time.perf_counter() - t_start > budget_s:
                rec[name] = {"skipped": "time budget of the encoder record used up"}
                continue
            rep = os.path.join(d, "rep.txt")
            # alternate the two builds (page cache, clocks and the pools' first touches favour whoever runs later); the better run of each counts
            refs, gpus = [], []
            for r_i in range(2 if name == "1080p_m8" else 1):
                refs.append(run_encoder(APP_REF, yuv, os.path.join(d, "ref.ivf"), cw, ch, n_frames, n_buf, 8, 1))
                gpus.append(run_encoder(APP_B200, yuv, os.path.join(d, "b200.ivf"), cw, ch, n_frames, n_buf, 8, 1,
                                        {"SVT_VP9_GPU": str(local_rank), "SVT_VP9_GPU_REPORT": rep}))
            if name != "1080p_m8":   # one more reference run after the drop-in run, so that the reference is never the only cold one
                refs.append(run_encoder(APP_REF, yuv, os.path.join(d, "ref.ivf"), cw, ch, n_frames, n_buf, 8, 1))
            best = lambda runs: max(runs, key=lambda r: r["fps"] or 0.0)
            ref, gpu = best(refs), best(gpus)
            one = {"frames": n_frames, "distinct_frames": n_buf, "reference": ref, "b200": gpu, "b200_gpu_report": gpu_report(rep),
                   "reference_fps_runs": [r["fps"] for r in refs], "b200_fps_runs": [r["fps"] for r in gpus],
                   "bitstream_identical": bool(ref["md5"]) and all(r["md5"] == ref["md5"] for r in refs + gpus),
                   "speedup": (gpu["fps"] / ref["fps"]) if ref["fps"] and gpu["fps"] else None}
            if name == "1080p_m8" and world == 1 and time.perf_counter() - t_start < budget_s:
                # host-bound ceiling: the same encode with motion estimation served from a table recorded by a previous run
                # (zero-cost ME: no GPU, no CPU kernel) -- what the pipeline around the seam allows at best
                tab = os.path.join(d, "me_table.bin")
                run_encoder(APP_B200, yuv, os.path.join(d, "rec.ivf"), cw, ch, n_frames, n_buf, 8, 1, {"SVT_VP9_GPU": str(local_rank), "SVT_VP9_ME_RECORD": tab})
                if os.path.exists(tab):
                    rp = run_encoder(APP_B200, yuv, os.path.join(d, "rep.ivf"), cw, ch, n_frames, n_buf, 8, 1, {"SVT_VP9_GPU": "off", "SVT_VP9_ME_REPLAY": tab})
                    one["me_free_ceiling"] = {"fps": rp["fps"], "bitstream_identical": rp["md5"] == ref["md5"],
                                              "note": "motion estimation replayed from a table (no ME work anywhere): the bound the host pipeline around seam S1 sets"}
            rec[name] = one
    finally:
        shutil.rmtree(d, ignore_errors=True)
    rec["seconds"] = round(time.perf_counter() - t_start, 1)
    return rec


def numa_bind(local_rank):
    """Run this rank (and place its pinned buffers, first-touch) on the host cores next to its GPU; returns the affinity before."""
    before = os.sched_getaffinity(0)
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return before, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
            return before, node
    except Exception:
        pass
    return before, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--verify", action="store_true", help="after the end-to-end leg, recompute its last pictures synchronously on one queue and compare every downloaded byte")
    ap.add_argument("--no-encoder", action="store_true", help="skip the whole-encoder record")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and check legs")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    config = {"workload": "configs[2]: 4K (3840x2160) 8-bit yuv420p, enc-mode 8 (tune 1 parameter set): per picture ME (2 lists) + one "
                          "transform/quantise/reconstruct pass over every TU (luma 32x32, chroma 16x16; 16x16 / 8x8 for the last 16 rows; q 45) + deblocking on the "
                          "base-layer picture; step = one 16-picture hierarchical mini-GOP in encode order",
              "pictures_per_step": PICS_PER_STEP,
              "queues": "motion estimation on the context's first compute queue, the picture's transform launch (all six TU batches in one "
                        "kernel) + deblocking on its second (they belong to different pictures in the encoder); value = "
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
            cpu.minigop()
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
            else:
                cpu.stage_cpu_s = [0.0, 0.0, 0.0]
            if time.perf_counter() - t_all > 200 and len(times) >= 2:
                break
            if time.perf_counter() - t_all > 200 and i < args.warmup:
                args.warmup = i  # bounded run: cut the warm-up short
        steps = len(times)
        v = PICS_PER_STEP * steps / sum(times) if times else 0.0
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "fps", "n_gpus": args.gpus, "steps": steps,
                          "warmup": args.warmup, "ms_per_step": 1000.0 * PICS_PER_STEP / v if v else None, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "fps", "cores": ncores, "kind": cpu.kind, "stage_share_of_cpu_time": cpu.stage_share(),
                                           "sample": "whole 4K mini-GOPs (16 pictures, all 2040 SBs and all TUs each), one native call per mini-GOP, %d pinned host threads" % ncores},
                          "e2e": {"value": v, "unit": "fps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        cpu.close()
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    affinity_before, numa_node = numa_bind(local_rank) if world > 1 else (os.sched_getaffinity(0), None)
    config["host_binding"] = ("rank bound to NUMA node %d (the GPU's), pinned buffers allocated after binding" % numa_node) if numa_node is not None else "none"
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
        ctx.picture_upload_luma(i, PS[i][0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, False)   # aligned device geometry, padding + 1/16 plane made on the device
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
    d_jobs = [to_dev(tu_jobs(sy, dplane[0].origin_x, dplane[0].origin_y, n, size, W4K, cols, row0) if plane == 0 else
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
    ctx.select_compute(api.COMPUTE3)
    stream3 = torch.cuda.ExternalStream(L.svtb200_stream(h), device=torch.device("cuda", local_rank))
    ctx.select_compute(api.COMPUTE)
    ME_QUEUES = int(os.environ.get("SVTB200_BENCH_ME_QUEUES", "1"))

    _txq_cache = {}
    api._bind_txq(L)

    def txq_batches(slot, n, r0):
        """the SvtB200TxBatch array (six batches: every TU of the three planes) of picture n into result slot 'slot' (built once)"""
        key = (slot, n)
        if key not in _txq_cache:
            ybytes = sy * dplane[n].padded_height
            rec, q, eob = d_rec[slot], d_q[slot], d_eob[slot]
            arr = (api.TxBatch * len(TU_PASSES))()
            for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
                if plane == 0:
                    pl = txq_planes(dplane[n].buf, sy, ybytes, dplane[r0].buf, rec[0], W4K, W4K * H4K)
                else:
                    dc = d_u if plane == 1 else d_v
                    pl = txq_planes(dc[n], W4K // 2, csize, dc[r0], rec[plane], W4K // 2, csize)
                b = arr[k]
                b.tx_size, b.n, b.jobs, b.planes = ts, pass_n[k], d_jobs[k], pl
                b.coeff, b.qcoeff, b.dqcoeff, b.eob = None, q[k], d_dq[k], eob[k]
            _txq_cache[key] = arr
        return _txq_cache[key]

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
        if overlap:
            ctx.select_compute(api.COMPUTE2)
        if timed:
            e[2].record(s2)
        # every TU of the picture (luma 32x32 + the 16x16 edge row, chroma 16x16 + 8x8 edge rows) in ONE launch
        rc = L.svtb200_txq_multi_dev(h, txq_batches(slot, n, r0), len(TU_PASSES), d_rows, 1)
        if rc:
            ctx.check(rc, "txq_multi_dev")
        if timed:
            e[3].record(s2)
        if n == 16:
            rec = d_rec[slot]
            ctx.check(L.svtb200_deblock_picture_dev(h, C.byref(LfP), rec[0], W4K, rec[1], rec[2], W4K // 2, d_lfm, 60, d_thr), "deblock")
        if timed:
            e[4].record(s2)
            (marks if sink is None else sink).append((e, P, n))
        L.svtb200_picture_touch(h, n)      # the transform kernel reads the planes of pictures n and r0 on this queue
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
    n_timed, n_warm = args.steps * PICS_PER_STEP, args.warmup * PICS_PER_STEP
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(n_warm):
        kernels(i, False, i % NSLOT)
    ctx.sync()
    barrier()
    del sampler.rows[:]   # keep only the samples taken under load: the timed legs from here on
    e0, e1 = ev(), ev()
    launches0 = ctx.launches()
    e0.record(stream)
    ctx.fence(api.COMPUTE, api.COMPUTE2)   # the other queues start inside the timed region ...
    ctx.fence(api.COMPUTE, api.COMPUTE3)
    t_host = time.perf_counter()
    for i in range(n_timed):
        kernels(n_warm + i, True, (n_warm + i) % NSLOT)
    t_host = time.perf_counter() - t_host
    if os.environ.get("SVTB200_BENCH_DEBUG"):
        sys.stderr.write("value leg: host issue time %.3f ms/picture\n" % (1000.0 * t_host / n_timed))
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
    for i in range(min(n_warm, 8)):
        step_e2e(i)
    ctx.sync()
    barrier()
    t0 = time.perf_counter()
    t_issue = 0.0
    for i in range(n_timed):
        ti = time.perf_counter()
        step_e2e(n_warm + i)
        t_issue += time.perf_counter() - ti
    ctx.sync()
    if os.environ.get("SVTB200_BENCH_DEBUG"):
        sys.stderr.write("e2e host time in step_e2e (issue + waits): %.3f ms/picture\n" % (1000.0 * t_issue / n_timed))
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1000.0
    sampler.stop()

    if args.verify:
        # the pipelined leg (4 queues, 3 slots) must have produced exactly what a synchronous single-queue run produces
        checked = 0
        for i in range(n_warm + n_timed - NSLOT, n_warm + n_timed):
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
    ms_res, units = sharding.agree_step(ms_res, n_timed, device="cuda")
    ms_e2e, _ = sharding.agree_step(ms_e2e, n_timed, device="cuda")
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
    launches_dom = n_me if dom in ("me_sb_kernel", "txq_kernel") else max(1, n_timed // 16)
    roof = {"bound": "hbm", "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per_kernel[dom]["frac"],
            "traffic": traffic.get(dom), "kernel": dom, "alg_bytes_per_launch": alg[dom] / launches_dom,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)", "per_kernel": per_kernel}

    # ---- check leg: two more pictures through the SAME end-to-end path (picture 16 -- deblocked -- and picture 8), then every
    # downloaded buffer against the CPU arm (the reference's own functions when oracle/_ref is present)
    check, cpu_line = None, None
    os.sched_setaffinity(0, affinity_before)
    if rank == 0 and not args.no_cpu:
        cpu = CpuPath(clip, ncores, n_buffers=PICS_PER_STEP if world == 1 else 2)
        base = ((n_warm + n_timed + 15) // 16) * 16        # i = base -> picture 16, base + 1 -> picture 8
        got = []
        for i in (base, base + 1):
            step_e2e(i)
            ctx.sync()
            slot, n = i % NSLOT, ENCODE_ORDER[i % 16]
            got.append((n, h_me[slot].numpy().copy(), [h_q[slot][k].numpy().copy() for k in range(len(TU_PASSES))],
                        [h_eob[slot][k].numpy().copy() for k in range(len(TU_PASSES))], [h_rec[slot][k].numpy().copy() for k in range(3)] if n == 16 else None))
        cpu.run([g[0] for g in got])
        bad, compared = [], 0
        for b, (n, me, q, eob, rec) in enumerate(got):
            ok = H.me_equal(me.view(api.ME_RESULT_DTYPE), cpu.me_out[b], 2)
            compared += me.nbytes
            if not ok.all():
                bad.append("picture %d: %d ME partitions differ" % (n, int((~ok).sum())))
            for k in range(len(TU_PASSES)):
                compared += q[k].nbytes + eob[k].nbytes
                if not np.array_equal(q[k].view(np.int16), cpu.q[b][k].reshape(-1)):
                    bad.append("picture %d: quantised coefficients of TU pass %d differ" % (n, k))
                if not np.array_equal(eob[k].view(np.uint16), cpu.eob[b][k]):
                    bad.append("picture %d: eob of TU pass %d differ" % (n, k))
            if rec is not None:
                for k in range(3):
                    compared += rec[k].nbytes
                    if not np.array_equal(rec[k], cpu.recon[b][k].reshape(-1)):
                        bad.append("picture %d: filtered plane %d differs" % (n, k))
        check = {"pictures": [g[0] for g in got], "bytes_compared": compared, "against": cpu.kind, "identical": not bad, "mismatches": bad[:8]}
        if bad:
            sys.stderr.write("check: MISMATCH " + "; ".join(bad[:8]) + "\n")
        if world == 1:
            cpu.minigop()          # page-fault / cache warm-up, untimed
            cpu.stage_cpu_s = [0.0, 0.0, 0.0]
            t0, reps = time.perf_counter(), 0
            while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 64):
                cpu.minigop()
                reps += 1
            dt = time.perf_counter() - t0
            cpu_line = {"value": reps * PICS_PER_STEP / dt, "unit": "fps", "cores": ncores, "kind": cpu.kind, "stage_share_of_cpu_time": cpu.stage_share(),
                        "sample": "%d whole 4K mini-GOPs (16 pictures each: ME + transform/quant/recon, deblocking of the base-layer picture), one native "
                                  "call per mini-GOP on %d pinned host threads, %.1f s" % (reps, ncores, dt)}
        cpu.close()
        del cpu

    # ---- BASELINE.json's own metric through the application (every rank encodes its own stream when world > 1)
    encoder = None
    if not args.no_encoder:
        enc = encoder_record(local_rank, world, clip if world == 1 else None)
        if world > 1:
            allenc = [None] * world
            dist.all_gather_object(allenc, enc)
        else:
            allenc = [enc]
        if rank == 0:
            encoder = {"streams": world, "per_stream": allenc[0] if world == 1 else None}
            if world > 1:
                agg = {}
                for arm in ("reference", "b200"):
                    fps = [(e.get("1080p_m8") or {}).get(arm, {}).get("fps") for e in allenc]
                    agg[arm + "_fps_sum"] = sum(fps) if all(fps) else None
                    agg[arm + "_fps_each"] = fps
                agg["bitstreams_identical"] = all((e.get("1080p_m8") or {}).get("bitstream_identical") for e in allenc)
                agg["speedup"] = (agg["b200_fps_sum"] / agg["reference_fps_sum"]) if agg["b200_fps_sum"] and agg["reference_fps_sum"] else None
                agg["note"] = "configs[4]: %d concurrent 1080p enc-mode 8 streams, one per GPU, reference run then drop-in run" % world
                encoder["concurrent_1080p_m8"] = agg

    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": value, "unit": "fps", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic", "config": config, "clocks": sampler.summary(),
                          "e2e": {"value": e2e, "unit": "fps", "h2d_bytes_per_step": h2d * PICS_PER_STEP, "d2h_bytes_per_step": d2h * PICS_PER_STEP},
                          "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu_line, "check": check, "encoder": encoder}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
