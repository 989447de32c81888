#!/usr/bin/env python
"""bench.py -- hot-path throughput on B200 ("Measurement").

A "step" is one pass of the hot path over one 3840x2160 picture of BASELINE.json's configs[2]
(4K 8-bit 4:2:0, enc-mode 8): the pictures of a 16-picture hierarchical mini-GOP are processed in
encode order with their real reference structure and per-temporal-layer search parameters.

  value      pictures / s with every plane already resident in HBM and results left in HBM
  e2e        pictures / s through the host-pointer C ABI: pinned-host upload of the picture's planes + kernel +
             download of the results, every step
  roofline   algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured HBM peak
  cpu_baseline  the CPU oracle on the host cores over a bounded sample of the same pictures

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
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

W4K, H4K = 3840, 2160
N_SB = 60 * 34
# encode order of a 16-picture mini-GOP (4 hierarchical levels) and the two references of each picture
ENCODE_ORDER = [16, 8, 4, 2, 1, 3, 6, 5, 7, 12, 10, 9, 11, 14, 13, 15]


def refs_of(n):
    if n == 16:
        return (0, 0)
    step = n & -n  # distance to the enclosing references
    return (n - step, n + step)


def me_bytes_per_picture(P):
    """(1+L)*(W*H + W*H/16 [+ W*H/4]) + nSB*85*40"""
    wh = P.luma_width * P.luma_height
    per = wh + wh // 16 + (wh // 4 if P.enable_hme_level1 else 0)
    return (1 + P.num_lists) * per + N_SB * 85 * 40


def make_pictures(n_frames, seed=11):
    w, h, sd, k, g, a, b, m, noise, _ = synth.CONFIGS["C"]
    clip = synth.make_clip(w, h, n_frames, seed, k, g, a, b, m, noise)
    return [frames.pa_planes(f[0]) for f in clip]


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_cpu_baseline(pics, sample_pictures, threads):
    """The CPU oracle (oracle/me_port.c) over whole pictures, SB ranges spread over host threads."""
    import helpers as H
    from concurrent.futures import ThreadPoolExecutor
    L = H.oracle()
    structs = [frames.plane_structs(p) for p in pics]
    t0 = time.perf_counter()
    done = 0
    for n in sample_pictures:
        r0, r1 = refs_of(n)
        tl = presets.MINIGOP16_TL[n - 1]
        P = presets.me_params_4k_m8(W4K, H4K, tl, 2, (r0, r1))
        out = np.zeros(N_SB * 85, dtype=api.ME_RESULT_DTYPE)
        chunk = (N_SB + threads * 4 - 1) // (threads * 4)
        ranges = [(s, min(N_SB, s + chunk)) for s in range(0, N_SB, chunk)]

        def work(rg):
            L.svtb200_oracle_me_picture(P, structs[n % len(pics)], structs[r0 % len(pics)], structs[r1 % len(pics)],
                                        out.ctypes.data, None, rg[0], rg[1])
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, ranges))
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    config = {"workload": "configs[2]: 4K (3840x2160) 8-bit yuv420p, enc-mode 8 (tune 1 parameter set), hot path = motion "
                          "estimation of every picture of a 16-picture hierarchical mini-GOP",
              "pictures_per_step": 1, "l2": "17 distinct pictures (165 MB of planes) cycled: working set larger than the 126 MB L2",
              "parallelism": "pictures sharded across GPUs (independent mini-GOP streams), no collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        n_frames = 17
        pics = make_pictures(n_frames)
        # bounded sample: one picture per step, cycling through representative temporal layers
        order = [1, 2, 4, 8, 16]
        fps_list = []
        t_all = time.perf_counter()
        for i in range(args.warmup + args.steps):
            fps, dt = run_cpu_baseline(pics, [order[i % len(order)]], ncores)
            if i >= args.warmup:
                fps_list.append(dt)
            if time.perf_counter() - t_all > 240:
                break
        steps = len(fps_list)
        v = steps / sum(fps_list) if fps_list else 0.0
        print(json.dumps({"impl": "reference", "metric": "encoded fps (hot path: motion estimation, 4K enc-mode 8)", "value": v, "unit": "fps",
                          "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v if v else None,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": v, "unit": "fps", "cores": ncores, "kind": "port",
                                           "sample": "one 4K picture (2040 SBs, 2 lists) per step, temporal layers cycled 4,3,2,1,0"},
                          "e2e": {"value": v, "unit": "fps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = api.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.L.svtb200_stream(ctx.h), device=torch.device("cuda", local_rank))

    n_frames = 17
    pics = make_pictures(n_frames, seed=11 + rank)
    structs = [frames.plane_structs(p) for p in pics]
    # pinned copies for the e2e leg
    pinned = []
    for p in pics:
        d = {}
        for k in ("full", "sixteenth"):
            t = torch.from_numpy(p[k]).pin_memory()
            d[k] = t
        pinned.append(d)

    def pinned_structs(i):
        p = pics[i]
        arr = (api.Plane * 3)()
        arr[0] = api.plane_of(pinned[i]["full"].numpy(), frames.FULL_PAD, frames.FULL_PAD, p["w"], p["h"])
        arr[2] = api.plane_of(pinned[i]["sixteenth"].numpy(), frames.SIXTEENTH_PAD, frames.SIXTEENTH_PAD, (p["w"] + 3) // 4, (p["h"] + 3) // 4)
        return arr

    pstructs = [pinned_structs(i) for i in range(n_frames)]
    for i in range(n_frames):
        ctx.picture_upload(i, pstructs[i][0], None, pstructs[i][2])
    ctx.sync()
    out_dev = ctx.L.svtb200_dev_alloc(ctx.h, N_SB * 85 * 40)
    out_host = np.zeros(N_SB * 85, dtype=api.ME_RESULT_DTYPE)

    def params_for(n):
        r0, r1 = refs_of(n)
        return presets.me_params_4k_m8(W4K, H4K, presets.MINIGOP16_TL[n - 1], 2, (r0, r1)), (r0, r1)

    import ctypes as C

    def step_resident(i):
        n = ENCODE_ORDER[i % 16]
        P, (r0, r1) = params_for(n)
        ids = (C.c_uint64 * 2)(r0, r1)
        rc = ctx.L.svtb200_me_picture_dev(ctx.h, C.byref(P), n, ids, out_dev, None)
        assert rc == 0, ctx.L.svtb200_last_error(ctx.h)
        return P

    def step_e2e(i):
        n = ENCODE_ORDER[i % 16]
        P, (r0, r1) = params_for(n)
        ctx.picture_upload(n, pstructs[n][0], None, pstructs[n][2])  # H2D of this picture's planes from pinned memory
        ids = (C.c_uint64 * 2)(r0, r1)
        rc = ctx.L.svtb200_me_picture(ctx.h, C.byref(P), n, ids, out_host.ctypes.data, None)  # kernel + D2H of the results
        assert rc == 0, ctx.L.svtb200_last_error(ctx.h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    # ---- device-resident leg
    for i in range(args.warmup):
        step_resident(i)
    ctx.sync()
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launches()
    alg_bytes = 0
    ev0.record(stream)
    for i in range(args.steps):
        P = step_resident(args.warmup + i)
        alg_bytes += me_bytes_per_picture(P)
    ev1.record(stream)
    ctx.sync()
    barrier()
    ms_res = ev0.elapsed_time(ev1)
    launches = ctx.launches() - launches0
    # ---- e2e leg (host buffers, copies inside the timed region)
    for i in range(min(args.warmup, 4)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    e1.record(stream)
    ctx.sync()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0)
    sampler.stop()

    t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t[0]), float(t[1])
    value = world * args.steps / (ms_res / 1000.0)
    e2e = world * args.steps / (ms_e2e / 1000.0)
    h2d = pics[0]["full"].nbytes + pics[0]["sixteenth"].nbytes
    d2h = N_SB * 85 * 40

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (ms_res / 1000.0) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel": "me_sb_kernel", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
            "alg_bytes_per_launch": alg_bytes / max(1, args.steps)}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        roof["traffic"] = prof.get("me_sb_kernel")
    except Exception:
        pass

    if rank == 0:
        cpu = None
        if world == 1:
            fps, dt = run_cpu_baseline(pics, [1, 2, 4, 8], ncores)
            cpu = {"value": fps, "unit": "fps", "cores": ncores, "kind": "port",
                   "sample": "4 whole 4K pictures (temporal layers 4,3,2,1), %d host threads, %.1f s" % (ncores, dt)}
        print(json.dumps({"metric": "encoded fps (hot path: motion estimation, 4K enc-mode 8)", "value": value, "unit": "fps",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": config, "clocks": sampler.summary(),
                          "e2e": {"value": e2e, "unit": "fps", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                          "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}))
    ctx.L.svtb200_dev_free(ctx.h, out_dev)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
