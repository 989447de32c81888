"""Reads the resolved motion-estimation parameter sets back from the UNMODIFIED reference (oracle/_ref, seam-dump flavour):
runs a short encode of the given size / enc-mode / tune and prints, per temporal layer and slice type, the SvtB200MeParams the
reference-side glue (svt-vp9_b200/shim/svtb200_ref_glue.h) translated from the reference's own structures.  This is how the
tables of svt-vp9_b200/presets.py were obtained.  Test infrastructure only.

usage: python oracle/dump_me_params.py width height enc_mode tune [frames]"""
import glob
import importlib
import os
import shutil
import struct
import subprocess
import sys
import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
api = importlib.import_module("svt-vp9_b200.api")
synth = importlib.import_module("svt-vp9_b200.synth")


def main():
    w, h, mode, tune = (int(x) for x in sys.argv[1:5])
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 18
    d = "/tmp/svtb200_params_%dx%d_%d_%d" % (w, h, mode, tune)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    yuv = os.path.join(d, "clip.yuv")
    base = synth.make_clip(w, h, 3, seed=9, k=3, g=3, a=5, b=3, m=32, noise=2)
    with open(yuv, "wb") as f:
        for i in range(n):
            for p in base[i % 3]:
                f.write(p.tobytes())
    env = dict(os.environ, SVTB200_SEAM_MODE="dump", SVTB200_SEAM_DIR=d, SVTB200_SEAM_MAXPICS=str(n))
    subprocess.run(["./SvtVp9EncApp_seam", "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n), "-enc-mode", str(mode), "-tune", str(tune),
                    "-b", os.path.join(d, "out.ivf")], cwd=os.path.join(ROOT, "oracle", "_ref"), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    seen = {}
    for path in sorted(glob.glob(os.path.join(d, "me_pic*.bin"))):
        with open(path, "rb") as f:
            head = f.read(16 + C.sizeof(api.MeParams))
        magic, ver, psize, n_sb = struct.unpack_from("<4I", head, 0)
        P = api.MeParams.from_buffer_copy(head[16:16 + psize])
        key = (P.temporal_layer_index, P.num_lists)
        if key in seen:
            continue
        seen[key] = path
        vals = {}
        for name, _ in api.MeParams._fields_:
            v = getattr(P, name)
            vals[name] = list(v) if hasattr(v, "__len__") else v
        print(os.path.basename(path), vals)
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
