"""Build steps for the in-tree native artefacts (all outputs are git-ignored .so files).

  build_product()  nvcc -> svt-vp9_b200/libsvtvp9_b200.so   (the sm_100a kernels + C ABI)
  build_oracle()   gcc  -> oracle/liboracle.so              (CPU restatement: checker only)
  build_emu()      g++  -> tests/emu/libme_emu.so           (kernel sources under the CPU emulation header)
  build_ref()      oracle/build_ref.sh -> oracle/_ref/*     (the unmodified reference; only where /root/reference exists)
"""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "svt-vp9_b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def product_lib():
    return os.path.join(PKG, "libsvtvp9_b200.so")


def build_product(force=False):
    src = sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))
    deps = src + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    out = product_lib()
    if force or _newer(out, deps):
        _run([NVCC] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
                              "-o", out] + src)
    return out


def oracle_lib():
    return os.path.join(ROOT, "oracle", "liboracle.so")


def build_oracle(force=False):
    src = sorted(glob.glob(os.path.join(ROOT, "oracle", "*_port.c")))
    deps = src + [os.path.join(ROOT, "oracle", "oracle.h"), os.path.join(ROOT, "include", "svtb200.h")]
    out = oracle_lib()
    if force or _newer(out, deps):
        _run(["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-pthread", "-o", out] + src + ["-lm"])
    return out


def emu_lib():
    return os.path.join(ROOT, "tests", "emu", "libme_emu.so")


def build_emu(force=False):
    src = sorted(glob.glob(os.path.join(ROOT, "tests", "emu", "*.cpp")))
    deps = src + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + glob.glob(os.path.join(PKG, "csrc", "emu", "*.h"))
    out = emu_lib()
    if force or _newer(out, deps):
        _run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-o", out] + src)
    return out


def ref_dir():
    return os.path.join(ROOT, "oracle", "_ref")


def build_ref():
    """Compile the unmodified reference where it lies (only in the container that has it)."""
    if not os.path.isdir(os.environ.get("SVT_REF", "/root/reference")):
        return None
    _run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    return ref_dir()


def build_all():
    build_product()
    build_oracle()
    build_emu()
    build_ref()


if __name__ == "__main__":
    build_all()
    print("ok")
