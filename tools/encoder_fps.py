"""Whole-encoder speed of the reference application and of the same objects with the GPU hot path bound in
(oracle/_ref/SvtVp9EncApp_ref vs SvtVp9EncApp_b200), on a synthetic clip.  Prints one JSON line per run.
usage: python tools/encoder_fps.py [width height frames enc_mode [host_cores]]   (host_cores: run the applications under taskset on that many cores,
the share of the host one GPU of an 8-GPU box gets)"""
import hashlib, json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import helpers as H

w, h, n, mode = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (3840, 2160, 33, 8)))
cores = int(sys.argv[5]) if len(sys.argv) >= 6 else 0
prefix = ["taskset", "-c", "0-%d" % (cores - 1)] if cores else []
d = tempfile.mkdtemp()
yuv = os.path.join(d, "clip.yuv")
base = H.synth.make_clip(w, h, 9, seed=5, k=3, g=4, a=5, b=3, m=32, noise=2)
with open(yuv, "wb") as f:
    for i in range(n):
        y, u, v = base[i % 9] if (i // 9) % 2 == 0 else base[8 - i % 9]
        f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())
env = dict(os.environ)
env["LD_LIBRARY_PATH"] = os.path.dirname(H.cudart_path()) + ":" + env.get("LD_LIBRARY_PATH", "")
for name in os.environ.get("ENCODER_FPS_APPS", "SvtVp9EncApp_ref,SvtVp9EncApp_b200,SvtVp9EncApp_b200").split(","):
    out = os.path.join(d, name + ".ivf")
    rep = os.path.join(d, "rep.txt")
    e = dict(env, SVT_VP9_GPU_REPORT=rep)
    t0 = time.time()
    r = subprocess.run(prefix + [os.path.join(ROOT, "oracle", "_ref", name), "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n), "-enc-mode", str(mode),
                        "-tune", "1", "-q", "45", "-b", out], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    dt = time.time() - t0
    txt = r.stdout.decode(errors="replace")
    m = re.search(r"Average Speed:\s*([0-9.]+)\s*fps", txt)
    m2 = re.search(r"Total Encoding Time:\s*([0-9.]+)\s*ms", txt)
    line = {"app": name, "host_cores": cores or os.cpu_count(), "rc": r.returncode, "enc_ms": float(m2.group(1)) if m2 else None, "tail": None if m else txt[-200:], "w": w, "h": h, "frames": n, "enc_mode": mode, "wall_s": round(dt, 2), "app_fps": float(m.group(1)) if m else None,
            "md5": hashlib.md5(open(out, "rb").read()).hexdigest() if os.path.exists(out) else None}
    if "b200" in name and os.path.exists(rep):
        line["gpu"] = dict((l.split()[0], int(l.split()[1])) for l in open(rep) if len(l.split()) == 2)
    print(json.dumps(line), flush=True)
