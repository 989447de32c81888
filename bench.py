#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200.

A "step" is one pass of the hot path over one 16-picture hierarchical mini-GOP of 3840x2160 pictures of BASELINE.json's
configs[2] (4K 8-bit 4:2:0, enc-mode 8), processed in encode order with the real reference structure and
per-temporal-layer search parameters; per picture:
    motion estimation (2 lists)  ->  inter prediction of an open-loop partition (every 32x32 block -- 16x16 in the last 16 rows --
    motion compensated with the best candidate motion estimation found for it: the encode pass's block walk + 8-tap predictors,
    three planes)  ->  transform / quantise / reconstruct every TU once (luma 32x32, chroma 16x16)  ->  the coefficients packed as
    the entropy coder reads them + their rate (coeff_rate_estimate)  ->  deblocking of the base-layer picture (enc-mode 8 filters
    temporal layer 0 only).

  value         pictures / s, planes already resident in HBM, results left in HBM (CUDA events on the context stream)
  e2e           pictures / s through the host-pointer side of the C ABI: pinned-host upload of the picture's planes,
                kernels, download of ME results, quantised coefficients (scan-order prefixes) + eob + rates and (base layer) the
                filtered planes
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
N_BLK = (W4K // 32) * (H4K // 32) + (H4K // 16 - 2 * (H4K // 32)) * (W4K // 16)   # blocks of the open-loop partition: 8040 32x32 + 240 16x16
CPAD = 40             # padding of the chroma reference planes (the reference pads chroma by half the luma padding)
METRIC = "pictures/s through the encode hot path (motion estimation + inter prediction + transform/quant/recon + coefficient packing / rate + deblocking; 4K, enc-mode 8 parameter set)"


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


def pred_bytes(P):
    """lower bound: every predicted sample read from one reference (compound blocks read two) and written once; per block the
    motion-estimation result (40 B), the block record (16 B) and three descriptors (64 B each)"""
    return (W4K * H4K * 3 // 2) * 2 + N_BLK * (40 + 16 + 3 * 64)


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
        self.luma = [aligned_copy(f[0]) for f in clip]  # unpadded, aligned: source planes of the transform stage
        self.chroma = [(aligned_copy(f[1]), aligned_copy(f[2])) for f in clip]
        # reference planes of the inter prediction: the padded luma of the picture-analysis stage, chroma padded by CPAD
        self.chroma_pad = [(aligned_copy(np.pad(f[1], CPAD, mode="edge")), aligned_copy(np.pad(f[2], CPAD, mode="edge"))) for f in clip]
        self.token_costs = np.ascontiguousarray(H.vp9_tables()["token_costs_default"])
        self.rblk = []
        for plane, ts, size, cols, rows, row0 in TU_PASSES:
            rb = np.zeros(cols * rows, dtype=api.RATE_BLOCK_DTYPE)
            rb["plane_type"], rb["is_inter"] = (1 if plane else 0), 1
            self.rblk.append(rb)
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
        self.pred = [[aligned((H4K, W4K), np.uint8), aligned((H4K // 2, W4K // 2), np.uint8), aligned((H4K // 2, W4K // 2), np.uint8)]
                     for _ in range(n_buffers)]
        self.rate = [[np.zeros(cols * rows, np.int32) for _, _, _, cols, rows, _ in TU_PASSES] for _ in range(n_buffers)]
        self.stage_cpu_s = [0.0, 0.0, 0.0, 0.0]
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
            pred = self.pred[b][plane]
            rec = self.recon[b][plane]
            tp = d.passes[k]
            tp.tx_size, tp.n, tp.jobs = ts, cols * rows, self.jobs[k].ctypes.data
            tp.pl.src, tp.pl.src_stride, tp.pl.src_bytes = src.ctypes.data, src.shape[1], src.nbytes
            tp.pl.pred, tp.pl.pred_stride, tp.pl.pred_bytes = pred.ctypes.data, pred.shape[1], pred.nbytes
            tp.pl.recon, tp.pl.recon_stride, tp.pl.recon_bytes = rec.ctypes.data, rec.shape[1], rec.nbytes
            tp.coeff, tp.qcoeff, tp.dqcoeff, tp.eob = None, self.q[b][k].ctypes.data, self.dq[b][k].ctypes.data, self.eob[b][k].ctypes.data
            tp.rblk, tp.rate = self.rblk[k].ctypes.data, self.rate[b][k].ctypes.data
        d.n_blocks = N_BLK
        d.pred = self._pred_picture(b, r0, r1)
        d.token_costs_all = self.token_costs.ctypes.data
        d.qrows = self.rows.ctypes.data
        d.deblock = 1 if n == 16 else 0
        d.lf = self.LfP
        d.y, d.u, d.v = (r.ctypes.data for r in self.recon[b])
        d.ys, d.cs = W4K, W4K // 2
        d.lfm, d.lfm_stride, d.thr = self.lfm.ctypes.data, 60, self.thr.ctypes.data

    def _pred_picture(self, b, r0, r1):
        pp = api.PredPicture()
        pp.width, pp.height, pp.use_subpel_flag = W4K, H4K, 1
        for l, r in enumerate((r0, r1)):
            full = self.pics[r]["full"]
            pp.ref[l][0] = full.ctypes.data + frames.FULL_PAD * full.shape[1] + frames.FULL_PAD
            pp.ref_stride[l][0] = full.shape[1]
            for c in range(2):
                a = self.chroma_pad[r][c]
                pp.ref[l][1 + c] = a.ctypes.data + CPAD * a.shape[1] + CPAD
                pp.ref_stride[l][1 + c] = a.shape[1]
        for k in range(3):
            pp.dst[k], pp.dst_stride[k] = self.pred[b][k].ctypes.data, self.pred[b][k].shape[1]
        return pp

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
        cpu_s = (C.c_double * 4)()
        rc = self.L.refdrv_pool_run(self.pool, arr, len(order), 4, 64, cpu_s)
        if rc:
            raise RuntimeError("refdrv_pool_run: %d" % rc)
        for k in range(4):
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
        # open-loop partition -> descriptors -> prediction planes (oracle/blocks_port.c, pred_port.c)
        blocks = np.zeros(N_BLK, dtype=api.INTER_BLOCK_DTYPE)
        L.svtb200_oracle_me_best_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.svtb200_oracle_inter_expand.argtypes = [C.c_void_p, C.c_int, C.POINTER(api.PredPicture), C.c_void_p]
        L.svtb200_oracle_inter_predict.argtypes = [C.c_void_p, C.c_int]
        assert L.svtb200_oracle_me_best_blocks(out.ctypes.data, W4K, H4K, blocks.ctypes.data) == N_BLK
        pp = self._pred_picture(b, r0, r1)
        desc = (api.PredBlock * (3 * N_BLK))()
        L.svtb200_oracle_inter_expand(blocks.ctypes.data, N_BLK, C.byref(pp), C.addressof(desc))
        list(self.tpool.map(lambda a: L.svtb200_oracle_inter_predict(C.addressof(desc) + a[1] * C.sizeof(api.PredBlock), a[2] - a[1]), self._split(3 * N_BLK)))
        Lt = H.oracle_tx()
        for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
            src = self.luma[n] if plane == 0 else self.chroma[n][plane - 1]
            pred = self.pred[b][plane]
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
            import test_rate as TR
            self.rate[b][k][:] = TR._oracle_rate(ts, np.ascontiguousarray(q.reshape(len(jobs), -1)), eob, self.rblk[k], self.token_costs)
        if n == 16:
            m = self.lfm.copy()
            L.svtb200_oracle_deblock_picture(C.byref(self.LfP), self.recon[b][0].ctypes.data, W4K, self.recon[b][1].ctypes.data, self.recon[b][2].ctypes.data,
                                             W4K // 2, m.ctypes.data, 60, self.thr.ctypes.data)

    def stage_share(self):
        t = sum(self.stage_cpu_s)
        return None if t <= 0 else {"motion_estimation": self.stage_cpu_s[0] / t, "inter_prediction": self.stage_cpu_s[3] / t,
                                    "transform_chain_and_rate": self.stage_cpu_s[1] / t, "deblocking": self.stage_cpu_s[2] / t}


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
            "bytes": os.path.getsize(out) if os.path.exists(out) else 0, "tail": None if m else txt[-300:]}


def bitstream_difference(a, b):
    """None when the two files hold the same bytes; otherwise their sizes and the offset of the first differing byte (a file cut short by
    the application's abort in its own teardown -- 3840x2160, SURVEY.md section 5 -- is a prefix of the complete one)."""
    try:
        x, y = np.fromfile(a, dtype=np.uint8), np.fromfile(b, dtype=np.uint8)
    except OSError:
        return {"missing": True}
    k = min(x.size, y.size)
    nz = np.flatnonzero(x[:k] != y[:k])
    if x.size == y.size and not nz.size:
        return None
    return {"bytes": [int(x.size), int(y.size)], "first_difference_at": int(nz[0]) if nz.size else int(k)}


def gpu_report(path):
    try:
        return dict((l.split()[0], int(l.split()[1])) for l in open(path) if len(l.split()) == 2)
    except Exception:
        return None


def encoder_record(local_rank, world, clip4k, budget_s=120.0, sync=None):
    """configs[1] (1080p, enc-mode 8, 300 frames) and configs[2] (4K, enc-mode 8; a bounded number of frames) through the
    application, reference build against drop-in build, one process per rank when world > 1 (configs[4]: concurrent streams,
    one per GPU; sync() lines the ranks up before every run so that the N reference encoders run together and the N drop-in encoders
    run together).  Returns the per-rank record; the caller aggregates over ranks."""
    sync = sync or (lambda: None)
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
            if world == 1 and time.perf_counter() - t_start > budget_s:   # (with several ranks every rank must make the same calls to sync)
                rec[name] = {"skipped": "time budget of the encoder record used up"}
                continue
            rep = os.path.join(d, "rep.txt")
            # alternate the two builds (page cache, clocks and the pools' first touches favour whoever runs later); the better run of each counts
            refs, gpus = [], []
            for r_i in range(2 if name == "1080p_m8" else 1):
                sync()
                refs.append(run_encoder(APP_REF, yuv, os.path.join(d, "ref.ivf"), cw, ch, n_frames, n_buf, 8, 1))
                sync()
                gpus.append(run_encoder(APP_B200, yuv, os.path.join(d, "b200.ivf"), cw, ch, n_frames, n_buf, 8, 1,
                                        {"SVT_VP9_GPU": str(local_rank), "SVT_VP9_GPU_REPORT": rep}))
                if gpus[-1]["md5"] != refs[-1]["md5"]:
                    gpus[-1]["difference"] = bitstream_difference(os.path.join(d, "ref.ivf"), os.path.join(d, "b200.ivf"))
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


def encoder_all_gpus_record(world, clip4k):
    """ONE encoder handle on all the GPUs of the launch (SVT_VP9_GPU=0,1,..: pictures dealt to the devices mini-GOP by mini-GOP): configs[2]
    through the application, reference build against drop-in build.  Runs on rank 0 while the other ranks wait."""
    if not (os.path.exists(APP_REF) and os.path.exists(APP_B200)):
        return {"unavailable": "oracle/_ref applications not built"}
    d = tempfile.mkdtemp(prefix="svtb200_encN_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        yuv = os.path.join(d, "c.yuv")
        synth.write_yuv(yuv, clip4k[:16])
        n_frames, rep = int(os.environ.get("SVTB200_BENCH_4K_FRAMES", "100")), os.path.join(d, "rep.txt")
        ref = run_encoder(APP_REF, yuv, os.path.join(d, "ref.ivf"), 3840, 2160, n_frames, 16, 8, 1)
        gpu = run_encoder(APP_B200, yuv, os.path.join(d, "b200.ivf"), 3840, 2160, n_frames, 16, 8, 1,
                          {"SVT_VP9_GPU": ",".join(str(i) for i in range(world)), "SVT_VP9_GPU_REPORT": rep})
        if gpu["md5"] != ref["md5"]:
            gpu["difference"] = bitstream_difference(os.path.join(d, "ref.ivf"), os.path.join(d, "b200.ivf"))
        return {"frames": n_frames, "gpus": world, "reference": ref, "b200": gpu, "b200_gpu_report": gpu_report(rep),
                "bitstream_identical": bool(ref["md5"]) and ref["md5"] == gpu["md5"], "speedup": (gpu["fps"] / ref["fps"]) if ref["fps"] and gpu["fps"] else None,
                "note": "one EbEncHandle pipeline, motion estimation and deblocking of its pictures dealt to %d GPUs mini-GOP by mini-GOP (no exchange between devices)" % world}
    finally:
        shutil.rmtree(d, ignore_errors=True)


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
                          "transform/quantise/reconstruct pass over every TU (luma 32x32, chroma 16x16; 16x16 / 8x8 for the last 16 rows; q 45; prediction = every "
                          "32x32 block motion compensated with its best motion-estimation candidate, 3 planes) + scan-order packing and rate of the coefficients + "
                          "deblocking on the base-layer picture; step = one 16-picture hierarchical mini-GOP in encode order",
              "pictures_per_step": PICS_PER_STEP,
              "queues": "motion estimation on the context's first compute queue; the rest of the picture's chain (open-loop partition, block walk, "
                        "inter prediction, the transform launch with all six TU batches, coefficient packing, rate, deblocking) on its second, "
                        "behind the picture's motion estimation and beside the next picture's; value = "
                        "CUDA events around the K steps with both queues joined inside the timed region",
              "coefficients": "quantised coefficients leave the device as the scan-order prefix (eob entries) of every TU -- what the entropy coder reads "
                              "(svtb200_coeff_pack_dev); in the end-to-end leg the block offsets (with the stream's length) come down with the other results and the exact-size copy "
                              "of the stream is queued two steps later (SVTB200_E2E_PACK=mapped: stored by the kernel straight into page-locked host memory, measured slower)",
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
                cpu.stage_cpu_s = [0.0, 0.0, 0.0, 0.0]
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
                                           "sample": "whole 4K mini-GOPs (16 pictures, all 2040 SBs, all prediction blocks and all TUs each), one native call per mini-GOP, %d pinned host threads" % ncores},
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
    # chroma planes live padded on the device (they are reference planes of the inter prediction): the picture area is uploaded into
    # the interior, the padding is made on the device (svtb200_pad_plane_dev = eb_vp9_generate_padding)
    cw, chh = W4K // 2, H4K // 2
    cstride, cbytes = cw + 2 * CPAD, (cw + 2 * CPAD) * (chh + 2 * CPAD)
    corg = CPAD * cstride + CPAD
    d_u = [L.svtb200_dev_alloc(h, cbytes) for _ in range(N_FRAMES)]
    d_v = [L.svtb200_dev_alloc(h, cbytes) for _ in range(N_FRAMES)]

    def upload_chroma(n):
        for d, src in ((d_u[n], pu[n]), (d_v[n], pv[n])):
            ctx.check(L.svtb200_memcpy_h2d_2d_async(h, d + corg, cstride, src.data_ptr(), cw, cw, chh), "h2d chroma")

    def pad_chroma(n):
        """on the current compute queue"""
        for d in (d_u[n], d_v[n]):
            ctx.check(L.svtb200_pad_plane_dev(h, d, cstride, cw, chh, CPAD, CPAD), "pad chroma")
    for i in range(N_FRAMES):
        ctx.picture_upload_luma(i, PS[i][0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, False)   # aligned device geometry, padding + 1/16 plane made on the device
        upload_chroma(i)
        ctx.sync()
        pad_chroma(i)
    ctx.sync()
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
    def gpu_jobs(plane, n, size, cols, row0):
        """source offsets into the padded resident planes; prediction and reconstruction offsets into packed planes"""
        j = tu_jobs(sy, dplane[0].origin_x, dplane[0].origin_y, n, size, W4K, cols, row0) if plane == 0 else tu_jobs(cstride, CPAD, CPAD, n, size, cw, cols, row0)
        j["pred_offset"] = j["recon_offset"]
        return j
    d_jobs = [to_dev(gpu_jobs(plane, n, size, cols, row0)) for n, (plane, _, size, cols, rows, row0) in zip(pass_n, TU_PASSES)]
    d_rows = to_dev(H.quant_rows([QINDEX]))
    d_lfm, d_thr = to_dev(lf_masks_32x32()), to_dev(H.lf_thresholds())
    # result buffers exist NSLOT times: the end-to-end leg downloads slot s while the kernels of the next pictures fill the others
    NSLOT = int(os.environ.get("SVTB200_BENCH_SLOTS", "4"))
    # the end-to-end leg queues a picture's exact-size coefficient-stream copy STREAM_LAG steps after the picture itself (the host must have
    # seen the stream's length by then): with a lag of one the host could never run more than one picture ahead of the download queue --
    # picture i + 1's upload was not even queued before picture i - 1 had come down -- and the copies did not overlap the kernels
    STREAM_LAG = max(1, min(NSLOT - 2, int(os.environ.get("SVTB200_E2E_STREAM_LAG", "2"))))
    # eob and rate arrays of the six passes are carved from one allocation per slot: one download each
    al = lambda v: (v + 255) & ~255
    eob_off = [sum(al(m * 2) for m in pass_n[:k]) for k in range(len(pass_n) + 1)]
    rate_off = [sum(al(m * 4) for m in pass_n[:k]) for k in range(len(pass_n) + 1)]
    d_rec = [[L.svtb200_dev_alloc(h, W4K * H4K), L.svtb200_dev_alloc(h, csize), L.svtb200_dev_alloc(h, csize)] for _ in range(NSLOT)]
    d_pred = [[L.svtb200_dev_alloc(h, W4K * H4K), L.svtb200_dev_alloc(h, csize), L.svtb200_dev_alloc(h, csize)] for _ in range(NSLOT)]
    d_blk = [L.svtb200_dev_alloc(h, N_BLK * 16) for _ in range(NSLOT)]
    d_desc = [L.svtb200_dev_alloc(h, 3 * N_BLK * C.sizeof(api.PredBlock)) for _ in range(NSLOT)]
    d_rate_all = [L.svtb200_dev_alloc(h, rate_off[-1]) for _ in range(NSLOT)]
    d_rate = [[d_rate_all[s_] + rate_off[k] for k in range(len(pass_n))] for s_ in range(NSLOT)]
    rblk = []
    for plane, _, _, cols, rows, _ in TU_PASSES:
        rb = np.zeros(cols * rows, dtype=api.RATE_BLOCK_DTYPE)
        rb["plane_type"], rb["is_inter"] = (1 if plane else 0), 1
        rblk.append(to_dev(rb))
    d_tc = [to_dev(np.ascontiguousarray(H.vp9_tables()["token_costs_default"][ts])) for ts in range(4)]
    ctx.register_rate_tables(H.vp9_tables())
    d_me = [L.svtb200_dev_alloc(h, N_SB * 85 * 40) for _ in range(NSLOT)]
    d_q = [[L.svtb200_dev_alloc(h, b) for b in pass_cb] for _ in range(NSLOT)]
    d_dq = [L.svtb200_dev_alloc(h, b) for b in pass_cb]
    d_eob_all = [L.svtb200_dev_alloc(h, eob_off[-1]) for _ in range(NSLOT)]
    d_eob = [[d_eob_all[s_] + eob_off[k] for k in range(len(pass_n))] for s_ in range(NSLOT)]
    LfP = api.LfParams()
    LfP.width, LfP.height, LfP.mi_rows, LfP.mi_cols, LfP.y_only, LfP.filter_level = W4K, H4K, H4K // 8, W4K // 8, 0, LF_LEVEL
    # pinned host result buffers for the e2e leg
    pinned = lambda n: torch.empty(n, dtype=torch.uint8).pin_memory()
    h_me = [pinned(N_SB * 85 * 40) for _ in range(NSLOT)]
    # quantised coefficients leave the device as scan-order prefixes (svtb200_coeff_pack_dev): one stream + offsets per picture.  The
    # value leg keeps them in HBM; the end-to-end leg lets the pack kernel store straight into these page-locked host buffers
    n_tu, n_coef = sum(pass_n), sum(pass_cb) // 2
    d_stream = [L.svtb200_dev_alloc(h, n_coef * 2) for _ in range(NSLOT)]
    d_off = [L.svtb200_dev_alloc(h, (n_tu + 1) * 4) for _ in range(NSLOT)]
    h_stream = [pinned(n_coef * 2) for _ in range(NSLOT)]
    h_off = [pinned((n_tu + 1) * 4) for _ in range(NSLOT)]
    h_off_np = [x.numpy().view(np.uint32) for x in h_off]
    h_eob_all = [pinned(eob_off[-1]) for _ in range(NSLOT)]
    h_rate_all = [pinned(rate_off[-1]) for _ in range(NSLOT)]
    h_eob = [[h_eob_all[s_][eob_off[k]:eob_off[k] + pass_n[k] * 2] for k in range(len(pass_n))] for s_ in range(NSLOT)]
    h_rate = [[h_rate_all[s_][rate_off[k]:rate_off[k] + pass_n[k] * 4] for k in range(len(pass_n))] for s_ in range(NSLOT)]
    h_rec = [[pinned(W4K * H4K), pinned(csize), pinned(csize)] for _ in range(NSLOT)]

    def txq_planes(src, sstride, sbytes, pred, rec, rstride, rbytes):
        """prediction and reconstruction planes share their (packed) geometry"""
        pl = api.TxPlanes()
        pl.src, pl.src_stride, pl.src_bytes = src, sstride, sbytes
        pl.pred, pl.pred_stride, pl.pred_bytes = pred, rstride, rbytes
        pl.recon, pl.recon_stride, pl.recon_bytes = rec, rstride, rbytes
        return pl

    ev = lambda: torch.cuda.Event(enable_timing=True)
    KERNELS = ("me_sb_kernel", "pred_kernel", "txq_kernel", "pack_kernel", "rate_kernel", "lf_row_kernel")
    stage_ms = {k: 0.0 for k in KERNELS}
    alg = {k: 0 for k in KERNELS}
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

    _pp_cache, _rb_cache = {}, {}

    def pred_picture(slot, r0, r1):
        """reference planes of both lists (resident, padded), prediction planes of the slot"""
        key = (slot, r0, r1)
        if key not in _pp_cache:
            pp = api.PredPicture()
            pp.width, pp.height, pp.use_subpel_flag = W4K, H4K, 1
            for l, r in enumerate((r0, r1)):
                pp.ref[l][0] = dplane[r].buf + dplane[r].origin_y * sy + dplane[r].origin_x
                pp.ref[l][1], pp.ref[l][2] = d_u[r] + corg, d_v[r] + corg
                pp.ref_stride[l][0], pp.ref_stride[l][1], pp.ref_stride[l][2] = sy, cstride, cstride
            for k in range(3):
                pp.dst[k], pp.dst_stride[k] = d_pred[slot][k], (W4K if k == 0 else cw)
            _pp_cache[key] = pp
        return _pp_cache[key]

    def rate_batches(slot):
        if slot not in _rb_cache:
            arr = (api.RateBatch * len(TU_PASSES))()
            for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
                b = arr[k]
                b.tx_size, b.n, b.qcoeff, b.eob, b.blocks, b.token_costs, b.rate = ts, pass_n[k], d_q[slot][k], d_eob[slot][k], rblk[k], d_tc[ts], d_rate[slot][k]
            _rb_cache[slot] = arr
        return _rb_cache[slot]

    def txq_batches(slot, n, r0=None):
        """the SvtB200TxBatch array (six batches: every TU of the three planes) of picture n into result slot 'slot' (built once)"""
        key = (slot, n)
        if key not in _txq_cache:
            ybytes = sy * dplane[n].padded_height
            rec, q, eob = d_rec[slot], d_q[slot], d_eob[slot]
            arr = (api.TxBatch * len(TU_PASSES))()
            for k, (plane, ts, size, cols, rows, row0) in enumerate(TU_PASSES):
                if plane == 0:
                    pl = txq_planes(dplane[n].buf, sy, ybytes, d_pred[slot][0], rec[0], W4K, W4K * H4K)
                else:
                    dc = d_u if plane == 1 else d_v
                    pl = txq_planes(dc[n], cstride, cbytes, d_pred[slot][plane], rec[plane], cw, csize)
                b = arr[k]
                b.tx_size, b.n, b.jobs, b.planes = ts, pass_n[k], d_jobs[k], pl
                b.coeff, b.qcoeff, b.dqcoeff, b.eob = None, q[k], d_dq[k], eob[k]
            _txq_cache[key] = arr
        return _txq_cache[key]

    _chain_cache = {}

    def chain_of(slot, n, r0, r1, pack_host, behind, on):
        """the SvtB200InterChain of picture n into result slot 'slot' (built once)"""
        key = (slot, n, pack_host, behind, on)
        if key not in _chain_cache:
            ch = api.InterChain()
            ch.after_queue, ch.on_queue = behind, on
            ch.me, ch.blocks, ch.n_blocks, ch.pic, ch.desc = d_me[slot], d_blk[slot], N_BLK, pred_picture(slot, r0, r1), d_desc[slot]
            ch.tx, ch.n_tx, ch.qrows, ch.n_qrows = txq_batches(slot, n), len(TU_PASSES), d_rows, 1
            ch.stream, ch.offsets = (h_stream[slot].data_ptr(), h_off[slot].data_ptr()) if pack_host else (d_stream[slot], d_off[slot])
            ch.capacity = n_coef
            ch.rate, ch.n_rate = rate_batches(slot), len(TU_PASSES)
            rd = sorted({n, r0, r1})
            for k, r in enumerate(rd):
                ch.reads[k] = r
            ch.n_reads = len(rd)
            _chain_cache[key] = ch
        return _chain_cache[key]

    def kernels(i, timed, slot=0, overlap=True, sink=None, pack_host=False, detail=False):
        """the kernels of step i: motion estimation on the first compute queue, the rest of the picture's chain on the second (one
        svtb200_inter_chain_dev call; detail=True: the chain's stages through their own entry points, an event after each)"""
        n = ENCODE_ORDER[i % 16]
        s2 = stream2 if overlap else stream
        q1, s1 = (api.COMPUTE3, stream3) if overlap and ME_QUEUES == 2 and (i & 1) else (api.COMPUTE, stream)
        P, (r0, r1) = me_params(n)
        ids = (C.c_uint64 * 2)(r0, r1)
        e = ([ev() for _ in range(9)] if detail else [ev(), ev()]) if timed else None
        if q1 != api.COMPUTE or detail:
            ctx.select_compute(q1)          # (the first compute queue is the selected one between calls)
        if timed:
            e[0].record(s1)
        ctx.check(L.svtb200_me_picture_dev(h, C.byref(P), n, ids, d_me[slot], None), "me_picture_dev")
        if timed:
            e[1].record(s1)
        if not detail:
            # (inside the timed region only the dominant kernel -- motion estimation -- carries events of its own: every event between
            # two kernels of a queue, and every extra host call, costs overlap)
            # open-loop partition -> block walk -> inter prediction (3 planes) -> every TU of the picture -> packing -> rate: one call,
            # on the second compute queue, queued behind the picture's motion estimation
            rc = L.svtb200_inter_chain_dev(h, C.addressof(chain_of(slot, n, r0, r1, pack_host, q1 if overlap else -1, api.COMPUTE2 if overlap else q1)))
            if rc:
                ctx.check(rc, "inter_chain_dev")
            if n == 16 and overlap:
                ctx.select_compute(api.COMPUTE2)
        else:
            if overlap:
                ctx.select_compute(api.COMPUTE2)
                ctx.stream_wait(api.COMPUTE2, ctx.marker(q1))
            e[2].record(s2)
            rc = L.svtb200_me_best_blocks_dev(h, d_me[slot], W4K, H4K, d_blk[slot])
            if rc != N_BLK:
                ctx.check(rc if rc < 0 else -1, "me_best_blocks_dev")
            ctx.check(L.svtb200_inter_blocks_expand_dev(h, d_blk[slot], N_BLK, C.byref(pred_picture(slot, r0, r1)), d_desc[slot]), "inter_blocks_expand_dev")
            ctx.check(L.svtb200_inter_predict_batch_dev(h, d_desc[slot], 3 * N_BLK), "inter_predict_batch_dev")
            e[6].record(s2)
            ctx.check(L.svtb200_txq_multi_dev(h, txq_batches(slot, n), len(TU_PASSES), d_rows, 1), "txq_multi_dev")
            e[3].record(s2)
            ctx.check(L.svtb200_coeff_pack_dev(h, txq_batches(slot, n), len(TU_PASSES), d_stream[slot], n_coef, d_off[slot]), "coeff_pack_dev")
            e[5].record(s2)
            ctx.check(L.svtb200_coeff_rate_multi_dev(h, rate_batches(slot), len(TU_PASSES)), "coeff_rate_multi_dev")
            e[7].record(s2)
            for r in {n, r0, r1}:
                L.svtb200_picture_touch(h, r)
        if n == 16:
            rec = d_rec[slot]
            ctx.check(L.svtb200_deblock_picture_dev(h, C.byref(LfP), rec[0], W4K, rec[1], rec[2], W4K // 2, d_lfm, 60, d_thr), "deblock")
        if timed:
            if detail:
                e[4].record(s2)
            (marks if sink is None else sink).append((e, P, n))
        if detail or q1 != api.COMPUTE or (n == 16 and overlap):
            ctx.select_compute(api.COMPUTE)
        return n, r0, r1

    # ---- end-to-end leg: host buffers in, host buffers out, through the C ABI's three queues.  Picture i+1 uploads and
    # picture i-1 downloads while picture i computes; the host collects picture i-3's results (the previous user of the slot)
    # before it queues picture i, i.e. NSLOT pictures are in flight (upload / kernels / download).
    downloaded = [0] * NSLOT   # marker after the last download out of each result slot
    stream_coefs = [0, 0]      # coefficients the pack kernel stored into host memory, pictures counted
    chroma_used = {}      # marker after the last kernel that read the chroma buffers of a picture

    # how the packed coefficients reach the host: "copy" (default): the pack kernel writes HBM, the offsets (with the stream length) come
    # down with the other results, and the exact-size copy of picture i's stream is queued one step later, when its length is known;
    # "mapped": the pack kernel stores straight into page-locked host memory (no second phase; measured slower: the stores hold the queue)
    PACK_MODE = os.environ.get("SVTB200_E2E_PACK", "copy")
    pending_stream = []        # [(slot, marker after the download of its offsets)]: pictures whose stream copy is still to be queued
    skip = os.environ.get("SVTB200_E2E_SKIP", "").split(",")  # diagnosis only: "h2d", "d2h", "kernels", "pack" (stream stays on the device), "me" (no ME download)

    def step_e2e(i):
        n = ENCODE_ORDER[i % 16]
        slot = i % NSLOT
        if downloaded[slot]:
            ctx.wait(downloaded[slot])                                          # picture i-NSLOT is complete in host memory; its slot is free
            stream_coefs[0] += int(h_off_np[slot][n_tu]); stream_coefs[1] += 1  # what the pack kernel stored for it
        if "h2d" not in skip:
            ctx.picture_upload_luma(n, PS[n][0], frames.QUARTER_PAD, frames.SIXTEENTH_PAD, False)   # H2D luma picture area; padding + 1/16 on the device
            if n in chroma_used:
                for m in chroma_used[n]:
                    ctx.stream_wait(api.UPLOAD, m)
            upload_chroma(n)                                                    # H2D chroma into the interior of the padded planes
            ctx.fence(api.UPLOAD, api.COMPUTE2)                                 # the chain reads the planes by address (motion estimation is
            ctx.select_compute(api.COMPUTE2)                                    # ordered behind the luma upload by the library); chroma padding
            pad_chroma(n)                                                       # on the queue that reads chroma
            ctx.select_compute(api.COMPUTE)
        if "kernels" in skip:
            r0 = r1 = n
        else:
            _, r0, r1 = kernels(i, False, slot, pack_host=PACK_MODE == "mapped" and "d2h" not in skip and "pack" not in skip)
        m2 = ctx.marker(api.COMPUTE2)                                           # the end of the picture's chain (which started behind its ME)
        chroma_used[n] = chroma_used[r0] = chroma_used[r1] = (m2,)              # the second compute queue reads the chroma planes
        ctx.stream_wait(api.DOWNLOAD, m2)
        rec = d_rec[slot]
        d2h = lambda hb, db, nb: ("d2h" in skip) or ctx.check(L.svtb200_memcpy_d2h_async(h, hb.data_ptr(), db, nb), "d2h")
        if "me" not in skip:
            d2h(h_me[slot], d_me[slot], N_SB * 85 * 40)                         # D2H ME results
        d2h(h_eob_all[slot], d_eob_all[slot], eob_off[-1])                      # D2H eob of every TU (the coefficients themselves were stored
        d2h(h_rate_all[slot], d_rate_all[slot], rate_off[-1])                   # into h_stream / h_off by the pack kernel) and their rates
        if n == 16:                                                            # D2H filtered reference picture
            d2h(h_rec[slot][0], rec[0], W4K * H4K)
            d2h(h_rec[slot][1], rec[1], csize)
            d2h(h_rec[slot][2], rec[2], csize)
        if PACK_MODE != "mapped" and "d2h" not in skip and "pack" not in skip:
            d2h(h_off[slot], d_off[slot], (n_tu + 1) * 4)                       # D2H block offsets + stream length
        downloaded[slot] = ctx.marker(api.DOWNLOAD)
        if PACK_MODE != "mapped" and "d2h" not in skip and "pack" not in skip:
            pending_stream.append((slot, downloaded[slot]))
            finish_streams(STREAM_LAG)                                          # the stream of picture i - STREAM_LAG: its length has arrived by now

    def finish_streams(keep=0):
        """queue the exact-size stream copies of the pictures whose offsets are (being) downloaded, all but the newest `keep`"""
        while len(pending_stream) > keep:
            sl, mk = pending_stream.pop(0)
            ctx.wait(mk)
            nbytes = 2 * int(h_off_np[sl][n_tu])
            if nbytes:
                ctx.check(L.svtb200_memcpy_d2h_async(h, h_stream[sl].data_ptr(), d_stream[sl], nbytes), "d2h stream")
            downloaded[sl] = ctx.marker(api.DOWNLOAD)

    def unpack_stream(stream, offsets, eobs):
        """the host-side consumer of the packed stream: dense per-pass coefficient arrays (all TUs of the bench are DCT_DCT)"""
        T, out, t = H.vp9_tables(), [], 0
        for k, (_, ts, size, _, _, _) in enumerate(TU_PASSES):
            scan = T["scan_%d_0" % ts].astype(np.int64)
            q = np.zeros((pass_n[k], size * size), dtype=np.int16)
            e = eobs[k].view(np.uint16).astype(np.int64)
            o = offsets[t:t + pass_n[k]].astype(np.int64)
            for j in np.flatnonzero(e):
                q[j, scan[:e[j]]] = stream[o[j]:o[j] + e[j]]
            out.append(q.reshape(-1).view(np.uint8))
            t += pass_n[k]
        return out

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
    NOEV = bool(os.environ.get("SVTB200_BENCH_NOEVENTS"))   # diagnosis: the timed region without the per-launch events of the dominant kernel
    # at most DEPTH pictures queued ahead of the GPU (as the end-to-end leg, which collects picture i-3 before it queues picture i):
    # with the launch queue filled to the brim the two compute queues overlap worse (measured: 1.72 k against 2.0 k pictures/s)
    DEPTH = int(os.environ.get("SVTB200_BENCH_DEPTH", "4"))
    inflight = []
    for i in range(n_timed):
        if DEPTH and len(inflight) >= DEPTH:
            ctx.wait(inflight.pop(0))
        kernels(n_warm + i, not NOEV, (n_warm + i) % NSLOT)
        if DEPTH:
            inflight.append(ctx.marker(api.COMPUTE2))
    t_host = time.perf_counter() - t_host
    if os.environ.get("SVTB200_BENCH_DEBUG"):
        sys.stderr.write("value leg: host issue time %.3f ms/picture\n" % (1000.0 * t_host / n_timed))
    ctx.fence(api.COMPUTE2, api.COMPUTE)   # ... and end inside it
    ctx.fence(api.COMPUTE3, api.COMPUTE)
    e1.record(stream)
    ctx.sync()
    barrier()
    ms_res = e0.elapsed_time(e1)
    ms_res_local = ms_res
    launches = ctx.launches() - launches0
    if NOEV:
        sys.stderr.write("no per-launch events: %.1f pictures/s\n" % (n_timed / (ms_res / 1000.0)))
        ctx.close()
        return
    n_me = 0
    for e, P, n in marks:
        stage_ms["me_sb_kernel"] += e[0].elapsed_time(e[1]); alg["me_sb_kernel"] += me_bytes(P); n_me += 1
        alg["pred_kernel"] += pred_bytes(P); alg["txq_kernel"] += txq_bytes()
        if n == 16:
            alg["lf_row_kernel"] += lf_bytes()
    # the same kernels one at a time (single queue): durations without the other queue's interference
    iso_marks, iso_ms = [], {k: 0.0 for k in KERNELS}
    iso_n = {k: 0 for k in KERNELS}
    for i in range(16):
        kernels(i, True, 0, overlap=False, sink=iso_marks, detail=True)
    ctx.sync()
    for e, P, n in iso_marks:
        iso_ms["me_sb_kernel"] += e[0].elapsed_time(e[1]); iso_n["me_sb_kernel"] += 1
        iso_ms["pred_kernel"] += e[2].elapsed_time(e[6]); iso_n["pred_kernel"] += 1
        iso_ms["txq_kernel"] += e[6].elapsed_time(e[3]); iso_n["txq_kernel"] += 1
        iso_ms["pack_kernel"] += e[3].elapsed_time(e[5]); iso_n["pack_kernel"] += 1
        iso_ms["rate_kernel"] += e[5].elapsed_time(e[7]); iso_n["rate_kernel"] += 1
        if n == 16:
            iso_ms["lf_row_kernel"] += e[7].elapsed_time(e[4]); iso_n["lf_row_kernel"] += 1
    for i in range(min(n_warm, 8)):
        step_e2e(i)
    finish_streams()
    ctx.sync()
    barrier()
    t0 = time.perf_counter()
    t_issue = 0.0
    for i in range(n_timed):
        ti = time.perf_counter()
        step_e2e(n_warm + i)
        t_issue += time.perf_counter() - ti
    finish_streams()
    ctx.sync()
    if os.environ.get("SVTB200_BENCH_DEBUG"):
        sys.stderr.write("e2e host time in step_e2e (issue + waits): %.3f ms/picture\n" % (1000.0 * t_issue / n_timed))
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1000.0
    sampler.stop()
    for slot in range(NSLOT):
        if downloaded[slot]:
            stream_coefs[0] += int(h_off_np[slot][n_tu]); stream_coefs[1] += 1
    stream_bytes = 2.0 * stream_coefs[0] / max(1, stream_coefs[1])   # mean packed stream of a picture
    alg["pack_kernel"] = int((2 * stream_bytes + n_tu * 6) * n_me)      # coefficients read + written once, eob in, offset out
    alg["rate_kernel"] = int((stream_bytes + n_tu * 10) * n_me)         # coefficients up to eob read once; eob + descriptor in, rate out

    if args.verify:
        # the pipelined leg (4 queues, NSLOT slots) must have produced exactly what a synchronous single-queue run produces
        checked = 0
        for i in range(n_warm + n_timed - NSLOT, n_warm + n_timed):
            slot, n = i % NSLOT, ENCODE_ORDER[i % 16]
            ncoef = int(h_off_np[slot][n_tu])
            got = [h_me[slot].numpy().copy(), h_off[slot].numpy().copy(), h_stream[slot].numpy()[:2 * ncoef].copy()]
            got += [h_eob[slot][k].numpy().copy() for k in range(len(TU_PASSES))]
            got += [h_rate[slot][k].numpy().copy() for k in range(len(TU_PASSES))]
            if n == 16:
                got += [h_rec[slot][k].numpy().copy() for k in range(3)]
            kernels(i, False, slot, overlap=False)
            ctx.sync()
            devs = [(d_me[slot], N_SB * 85 * 40), (d_off[slot], (n_tu + 1) * 4), (d_stream[slot], 2 * ncoef)]
            devs += [(d_eob[slot][k], pass_n[k] * 2) for k in range(len(TU_PASSES))]
            devs += [(d_rate[slot][k], pass_n[k] * 4) for k in range(len(TU_PASSES))]
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
    # ME results, the packed coefficient stream + its offsets (stored by the pack kernel), eob + rate, filtered planes on 1 of 16 pictures
    d2h = N_SB * 85 * 40 + int(stream_bytes) + (n_tu + 1) * 4 + (2 + 4) * n_tu + (W4K * H4K * 3 // 2) // 16

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
        per_pic = alg[k] / (max(1, sum(1 for m in marks if m[2] == 16)) if k == "lf_row_kernel" else n_me)
        iso = iso_ms[k] / max(1, iso_n[k])
        row = {"traffic": traffic.get(k), "ms_per_picture_alone": iso, "achieved_gbs_alone": per_pic / (iso / 1000.0) / 1e9 if iso else None,
               "frac_alone": per_pic / (iso / 1000.0) / 1e9 / peak if iso else None, "share_alone": per_step_alone[k] / sum(per_step_alone.values())}
        if stage_ms[k] > 0:   # the dominant kernel carries events of its own inside the timed region (the other queue's kernels overlap it)
            a = alg[k] / (stage_ms[k] / 1000.0) / 1e9
            row.update({"ms_total": stage_ms[k], "ms_per_launch_in_timed_region": stage_ms[k] / n_me, "share_of_timed_region": stage_ms[k] / ms_res_local,
                        "achieved_gbs": a, "frac": a / peak})
        per_kernel[k] = row
    launches_dom = n_me if dom != "lf_row_kernel" else max(1, n_timed // 16)
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
            finish_streams()
            ctx.sync()
            slot, n = i % NSLOT, ENCODE_ORDER[i % 16]
            eobs = [h_eob[slot][k].numpy().copy() for k in range(len(TU_PASSES))]
            got.append((n, h_me[slot].numpy().copy(), unpack_stream(h_stream[slot].numpy().view(np.int16), h_off_np[slot], eobs), eobs,
                        [h_rec[slot][k].numpy().copy() for k in range(3)] if n == 16 else None, [h_rate[slot][k].numpy().copy() for k in range(len(TU_PASSES))]))
        cpu.run([g[0] for g in got])
        bad, compared = [], 0
        for b, (n, me, q, eob, rec, rate) in enumerate(got):
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
                compared += rate[k].nbytes
                if not np.array_equal(rate[k].view(np.int32), cpu.rate[b][k]):
                    bad.append("picture %d: rates of TU pass %d differ" % (n, k))
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
            cpu.stage_cpu_s = [0.0, 0.0, 0.0, 0.0]
            t0, reps = time.perf_counter(), 0
            while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 64):
                cpu.minigop()
                reps += 1
            dt = time.perf_counter() - t0
            cpu_line = {"value": reps * PICS_PER_STEP / dt, "unit": "fps", "cores": ncores, "kind": cpu.kind, "stage_share_of_cpu_time": cpu.stage_share(),
                        "sample": "%d whole 4K mini-GOPs (16 pictures each: ME + inter prediction + transform/quant/recon + rate, deblocking of the base-layer picture), one native "
                                  "call per mini-GOP on %d pinned host threads, %.1f s" % (reps, ncores, dt)}
        cpu.close()
        del cpu

    # ---- BASELINE.json's own metric through the application (every rank encodes its own stream when world > 1)
    encoder = None
    if not args.no_encoder:
        enc = encoder_record(local_rank, world, clip if world == 1 else None, sync=(dist.barrier if world > 1 else None))
        enc_all = None
        if world > 1:
            allenc = [None] * world
            dist.all_gather_object(allenc, enc)
            if rank == 0:
                try:
                    enc_all = encoder_all_gpus_record(world, clip)
                except Exception as ex:   # never lose the bench line to the extra record
                    enc_all = {"error": repr(ex)}
            dist.barrier()
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
                encoder["one_encoder_all_gpus_4k_m8"] = enc_all

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
