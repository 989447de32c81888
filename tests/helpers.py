"""Shared test plumbing: native library loaders, the ME fixture format, oracle / emulation bindings."""
import ctypes as C
import importlib
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

api = importlib.import_module("svt-vp9_b200.api")
build = importlib.import_module("svt-vp9_b200.build")
synth = importlib.import_module("svt-vp9_b200.synth")

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libSvtVp9Enc_ref.so"))


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(build.build_oracle())
        L.svtb200_oracle_me_sb.argtypes = [C.POINTER(api.MeParams), C.POINTER(api.Plane), C.POINTER(api.Plane),
                                           C.POINTER(api.Plane), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svtb200_oracle_me_picture.argtypes = [C.POINTER(api.MeParams), C.POINTER(api.Plane), C.POINTER(api.Plane),
                                                C.POINTER(api.Plane), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _oracle = L
    return _oracle


_emu = None


def emu():
    global _emu
    if _emu is None:
        L = C.CDLL(build.build_emu())
        L.me_emu_run.argtypes = [C.POINTER(api.MeParams), C.POINTER(api.Plane), C.POINTER(api.Plane), C.POINTER(api.Plane),
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _emu = L
    return _emu


class MePicture:
    """One picture's ME inputs + the reference's results (format written by oracle/ref_seam_check.c, dump mode)."""

    def __init__(self, path):
        if path.endswith(".gz"):
            import gzip
            with gzip.open(path, "rb") as f:
                data = f.read()
        else:
            with open(path, "rb") as f:
                data = f.read()
        magic, ver, psize, n_sb = struct.unpack_from("<4I", data, 0)
        assert magic == 0x42305653 and ver == 1 and psize == C.sizeof(api.MeParams), (hex(magic), ver, psize)
        off = 16
        self.params = api.MeParams.from_buffer_copy(data[off:off + psize])
        off += psize
        self.n_sb = n_sb

        def rd_plane(off):
            stride, ox, oy, w, h, ph = struct.unpack_from("<6i", data, off)
            off += 24
            arr = None
            if ph > 0 and stride > 0:
                arr = np.frombuffer(data, dtype=np.uint8, count=stride * ph, offset=off).reshape(ph, stride).copy()
                off += stride * ph
            return (arr, ox, oy, w, h), off

        self.cur, self.ref = [], [[], []]
        for _ in range(3):
            p, off = rd_plane(off)
            self.cur.append(p)
        for l in range(self.params.num_lists):
            for _ in range(3):
                p, off = rd_plane(off)
                self.ref[l].append(p)
        self.expected = np.frombuffer(data, dtype=api.ME_RESULT_DTYPE, count=n_sb * 85, offset=off).copy()

    @staticmethod
    def _planes(lst):
        arr = (api.Plane * 3)()
        for i, (a, ox, oy, w, h) in enumerate(lst):
            if a is not None:
                arr[i] = api.plane_of(a, ox, oy, w, h)
        return arr

    def cur_planes(self):
        return self._planes(self.cur)

    def ref_planes(self, l):
        return self._planes(self.ref[l])


def me_equal(a, b, num_lists):
    """Compare ME results field by field; on P pictures the list-1 MV fields are undefined in the reference."""
    ok = (a["x_mv_l0"] == b["x_mv_l0"]) & (a["y_mv_l0"] == b["y_mv_l0"]) & (a["total"] == b["total"])
    if num_lists == 2:
        ok &= (a["x_mv_l1"] == b["x_mv_l1"]) & (a["y_mv_l1"] == b["y_mv_l1"])
    for i in range(3):
        used = a["total"] > i
        ok &= ~used | ((a["cand"]["distortion"][:, i] == b["cand"]["distortion"][:, i]) &
                       (a["cand"]["direction"][:, i] == b["cand"]["direction"][:, i]))
    return ok


def oracle_me_picture(pic, sb_begin=0, sb_end=None, trace=False):
    sb_end = pic.n_sb if sb_end is None else sb_end
    out = np.zeros(pic.n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    rcme = np.zeros(pic.n_sb, dtype=np.uint32)
    cur, r0 = pic.cur_planes(), pic.ref_planes(0)
    r1 = pic.ref_planes(1) if pic.params.num_lists == 2 else None
    tr = np.zeros(pic.n_sb, dtype=api.ME_TRACE_DTYPE)
    sbs_w = (pic.params.luma_width + 63) // 64
    L = oracle()
    for sb in range(sb_begin, sb_end):
        L.svtb200_oracle_me_sb(C.byref(pic.params), cur, r0, r1, sb % sbs_w, sb // sbs_w,
                               out[sb * 85:].ctypes.data, rcme[sb:].ctypes.data, tr[sb:].ctypes.data if trace else None)
    return (out, rcme, tr) if trace else (out, rcme)


def emu_me_picture(pic, sb_begin=0, sb_end=None, win_bytes=65536, force_global_interp=0):
    sb_end = pic.n_sb if sb_end is None else sb_end
    out = np.zeros(pic.n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    rcme = np.zeros(pic.n_sb, dtype=np.uint32)
    tr = np.zeros(pic.n_sb, dtype=api.ME_TRACE_DTYPE)
    cur, r0 = pic.cur_planes(), pic.ref_planes(0)
    r1 = pic.ref_planes(1) if pic.params.num_lists == 2 else None
    emu().me_emu_run(C.byref(pic.params), cur, r0, r1, out.ctypes.data, rcme.ctypes.data, tr.ctypes.data, sb_begin, sb_end,
                     win_bytes, force_global_interp)
    return out, rcme, tr
