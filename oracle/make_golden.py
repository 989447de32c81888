"""Regenerates tests/golden/me_*.bin.gz from the UNMODIFIED reference (oracle/_ref, seam-dump flavour).

Run in the container that has /root/reference (after oracle/build_ref.sh).  The dumps are written by
oracle/ref_seam_check.c in dump mode: per picture the translated SvtB200MeParams, the padded
full / quarter / sixteenth planes of the current and reference pictures, and the reference's own
me_results for every super-block.  Test infrastructure only.
"""
import glob
import gzip
import importlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
synth = importlib.import_module("svt-vp9_b200.synth")

SELECT = {(8, 1): [3, 8, 16, 17], (3, 1): [2, 16, 18], (5, 0): [1, 4, 16]}


def main():
    w, h, n = 192, 128, 20
    yuv = "/tmp/svtb200_golden.yuv"
    synth.write_yuv(yuv, synth.make_clip(w, h, n, seed=9, k=3, g=3, a=5, b=3, m=32, noise=2))
    for (mode, tune), pics in SELECT.items():
        d = "/tmp/svtb200_gold_%d_%d" % (mode, tune)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        env = dict(os.environ, SVTB200_SEAM_MODE="dump", SVTB200_SEAM_DIR=d, SVTB200_SEAM_MAXPICS="32")
        subprocess.run(["./SvtVp9EncApp_seam", "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n), "-enc-mode", str(mode),
                        "-tune", str(tune), "-b", "/tmp/svtb200_golden.ivf"], cwd=os.path.join(ROOT, "oracle", "_ref"), env=env,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for p in pics:
            dst = os.path.join(ROOT, "tests", "golden", "me_%dx%d_m%dt%d_pic%02d.bin.gz" % (w, h, mode, tune, p))
            with open(os.path.join(d, "me_pic%04d.bin" % p), "rb") as f, gzip.open(dst, "wb", 9) as g:
                shutil.copyfileobj(f, g)
            print("wrote", dst)


if __name__ == "__main__":
    main()
    make_lf_golden()


def make_lf_golden():
    """tests/golden/lf_*.bin.gz: loop-filtered pictures (inputs, masks, thresholds and the reference's output) of a smooth
    192x128 clip at enc-mode 3, q 50 -- the textured clip above barely triggers the filter."""
    w, h, n = 192, 128, 20
    yuv = "/tmp/svtb200_golden_smooth.yuv"
    synth.write_yuv(yuv, synth.make_clip(w, h, n, seed=21, k=10, g=2.0, a=2, b=1, m=32, noise=0))
    d = "/tmp/svtb200_gold_lf"
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    env = dict(os.environ, SVTB200_SEAM_MODE="dump", SVTB200_SEAM_DIR=d, SVTB200_SEAM_MAXPICS="12")
    subprocess.run(["./SvtVp9EncApp_seam", "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n), "-enc-mode", "3", "-tune", "1",
                    "-q", "50", "-b", "/tmp/svtb200_golden.ivf"], cwd=os.path.join(ROOT, "oracle", "_ref"), env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for p in (0, 2, 9):
        dst = os.path.join(ROOT, "tests", "golden", "lf_%dx%d_smooth_m3_%02d.bin.gz" % (w, h, p))
        with open(os.path.join(d, "lf_pic%04d.bin" % p), "rb") as f, gzip.open(dst, "wb", 9) as g:
            shutil.copyfileobj(f, g)
        print("wrote", dst)
