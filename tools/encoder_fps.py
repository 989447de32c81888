"""Whole-encoder speed of the reference application and of the same objects with the GPU hot path bound in
(oracle/_ref/SvtVp9EncApp_ref vs SvtVp9EncApp_b200), on a synthetic clip, run the way bench.py's encoder record runs them
(frames preloaded by the application, its own "Average Speed").  Prints one JSON line per run.
usage: python tools/encoder_fps.py [width height frames enc_mode [distinct_frames]]
ENCODER_FPS_APPS: comma-separated applications, each optionally followed by @NAME=value;NAME=value (environment of that run)."""
import json, os, shutil, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers as H
import bench

w, h, n, mode = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (3840, 2160, 33, 8)))
distinct = int(sys.argv[5]) if len(sys.argv) >= 6 else 16
d = tempfile.mkdtemp(prefix="svtb200_fps_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    yuv = os.path.join(d, "clip.yuv")
    H.synth.write_yuv(yuv, H.synth.make_clip(w, h, distinct, seed=5, k=3, g=4, a=5, b=3, m=32, noise=2))
    first = None
    for spec in os.environ.get("ENCODER_FPS_APPS", "SvtVp9EncApp_ref,SvtVp9EncApp_b200,SvtVp9EncApp_b200").split(","):
        name, _, extra = spec.partition("@")
        rep = os.path.join(d, "rep.txt")
        if os.path.exists(rep):
            os.remove(rep)
        env = {"SVT_VP9_GPU_REPORT": rep}
        env.update(kv.split("=", 1) for kv in extra.split(";") if "=" in kv)
        line = bench.run_encoder(os.path.join(bench.REF_DIR, name), yuv, os.path.join(d, name + ".ivf"), w, h, n, distinct, mode, 1, env)
        line = dict({"app": spec, "w": w, "h": h, "frames": n, "enc_mode": mode}, **line)
        if "b200" in name:
            line["gpu"] = bench.gpu_report(rep)
        # where a bitstream that differs from the first run's starts to differ (a file cut short by the application's abort in teardown
        # is a prefix of the complete one)
        cur = open(os.path.join(d, name + ".ivf"), "rb").read() if line["md5"] else b""
        if first is None:
            first = cur
        elif cur != first:
            import numpy as np
            k = min(len(cur), len(first))
            a, b = np.frombuffer(cur[:k], np.uint8), np.frombuffer(first[:k], np.uint8)
            nz = np.flatnonzero(a != b)
            line["first_difference_at"] = int(nz[0]) if nz.size else k
            line["bytes_of_first_run"] = len(first)
        print(json.dumps(line), flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
