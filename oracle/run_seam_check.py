"""Runs the seam-check flavour of the UNMODIFIED reference (oracle/_ref/SvtVp9EncApp_seam, oracle/ref_seam_check.c) in check mode on a
synthetic clip and prints its summary: every super-block's motion estimation, block statistics, luma / chroma histograms and chroma
block means, the edge maps and the noise detection of the oracle restatements against what the reference computed inside the real encode.
usage: run_seam_check.py <width> <height> <enc-mode> <tune> [frames [noise amplitude of the clip [blur passes [contrast gain]]]]"""
import importlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("svt-vp9_b200.synth")


def main():
    w, h, mode, tune = (int(x) for x in sys.argv[1:5])
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 18
    noise = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    k = int(sys.argv[7]) if len(sys.argv) > 7 else 3      # more passes / less gain: smoother content (the noise classes the decimated
    g = int(sys.argv[8]) if len(sys.argv) > 8 else 3      # detection methods can be checked on)
    d = "/tmp/svtb200_seam_%dx%d_%d_%d" % (w, h, mode, tune)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    yuv = os.path.join(d, "clip.yuv")
    base = synth.make_clip(w, h, 5, seed=21, k=k, g=g, a=5, b=3, m=32, noise=noise)
    with open(yuv, "wb") as f:
        for i in range(n):
            for p in base[i % 5]:
                f.write(p.tobytes())
    env = dict(os.environ, SVTB200_SEAM_MODE="check", SVTB200_SEAM_DIR=d)
    subprocess.run(["./SvtVp9EncApp_seam", "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n), "-enc-mode", str(mode), "-tune", str(tune),
                    "-b", os.path.join(d, "out.ivf")], cwd=os.path.join(ROOT, "oracle", "_ref"), env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print(open(os.path.join(d, "summary.txt")).read())
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
