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
        L.svtb200_oracle_zz_sad_picture.argtypes = [C.POINTER(api.Plane), C.POINTER(api.Plane), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
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


# ---------------------------------------------------------------------------------------------
# transform / quantisation helpers
# ---------------------------------------------------------------------------------------------
_tables = None


def vp9_tables():
    global _tables
    if _tables is None:
        _tables = np.load(os.path.join(ROOT, "svt-vp9_b200", "data", "vp9_tables.npz"))
    return _tables


def scan_of(tx_size, tx_type):
    t = vp9_tables()
    return t["scan_%d_%d" % (tx_size, tx_type)], t["iscan_%d_%d" % (tx_size, tx_type)]


_oracle_tx_ready = False


def oracle_tx():
    global _oracle_tx_ready
    L = oracle()
    if not _oracle_tx_ready:
        L.svtb200_oracle_fwd_txfm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.svtb200_oracle_quantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svtb200_oracle_quant_row.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.svtb200_oracle_inv_txfm_add.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _oracle_tx_ready = True
    return L


def oracle_fwd(tx_size, tx_type, partial32, resid):
    """resid: (n, N, N) int16 -> (n, N, N) int16"""
    L = oracle_tx()
    n, N = resid.shape[0], resid.shape[1]
    resid = np.ascontiguousarray(resid)
    out = np.zeros_like(resid)
    for i in range(n):
        L.svtb200_oracle_fwd_txfm(tx_size, tx_type, partial32, resid[i].ctypes.data, N, out[i].ctypes.data)
    return out


def quant_rows(qindices):
    t = vp9_tables()
    rows = np.zeros(len(qindices), dtype=api.QUANT_ROW_DTYPE)
    L = oracle_tx()
    for i, q in enumerate(qindices):
        L.svtb200_oracle_quant_row(int(q), int(t["dc_q"][q]), int(t["ac_q"][q]), rows[i:].ctypes.data)
    return rows


def oracle_quant(tx_size, tx_type, coeff, rows, row_idx):
    L = oracle_tx()
    scan, _ = scan_of(tx_size, tx_type)
    scan = np.ascontiguousarray(scan)
    n = coeff.shape[0]
    coeff = np.ascontiguousarray(coeff)
    q, dq, eob = np.zeros_like(coeff), np.zeros_like(coeff), np.zeros(n, dtype=np.uint16)
    for i in range(n):
        L.svtb200_oracle_quantize(tx_size, coeff[i].ctypes.data, rows[int(row_idx[i]):].ctypes.data, scan.ctypes.data,
                                  q[i].ctypes.data, dq[i].ctypes.data, eob[i:].ctypes.data)
    return q, dq, eob


def oracle_inv_add(tx_size, tx_type, dq, eob, dest):
    L = oracle_tx()
    N = dq.shape[1]
    dq = np.ascontiguousarray(dq)
    out = np.ascontiguousarray(dest.copy())
    for i in range(dq.shape[0]):
        L.svtb200_oracle_inv_txfm_add(tx_size, tx_type, dq[i].ctypes.data, int(eob[i]), out[i].ctypes.data, N)
    return out


# ---- the unmodified reference's own C kernels (oracle/_ref), for pinning the oracle
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(REF_DIR, "libSvtVp9Enc_ref.so"))
    return _ref


def ref_fwd(tx_size, tx_type, partial32, resid):
    R = ref_lib()
    n, N = resid.shape[0], resid.shape[1]
    resid = np.ascontiguousarray(resid)
    out = np.zeros_like(resid)
    for i in range(n):
        src, dst = C.c_void_p(resid[i].ctypes.data), C.c_void_p(out[i].ctypes.data)
        if tx_size == 3:
            (R.eb_vpx_partial_fdct32x32_c if partial32 else R.eb_vp9_fdct32x32_c)(src, dst, N)
        else:
            [R.eb_vp9_fht4x4_c, R.eb_vp9_fht8x8_c, R.eb_vp9_fht16x16_c][tx_size](src, dst, N, tx_type)
    return out


def ref_quant(tx_size, tx_type, coeff, rows, row_idx):
    R = ref_lib()
    scan, iscan = scan_of(tx_size, tx_type)
    scan, iscan = np.ascontiguousarray(scan), np.ascontiguousarray(iscan)
    n, nn = coeff.shape[0], coeff.shape[1] * coeff.shape[2]
    coeff = np.ascontiguousarray(coeff)
    q, dq, eob = np.zeros_like(coeff), np.zeros_like(coeff), np.zeros(n, dtype=np.uint16)
    f = R.eb_vp9_quantize_b_32x32_c if tx_size == 3 else R.eb_vp9_quantize_b_c
    f.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int] + [C.c_void_p] * 10
    for i in range(n):
        r = rows[int(row_idx[i]):int(row_idx[i]) + 1]
        f(coeff[i].ctypes.data, nn, 0, r["zbin"].ctypes.data, r["round"].ctypes.data, r["quant"].ctypes.data,
          r["quant_shift"].ctypes.data, q[i].ctypes.data, dq[i].ctypes.data, r["dequant"].ctypes.data, eob[i:].ctypes.data,
          scan.ctypes.data, iscan.ctypes.data)
    return q, dq, eob


def ref_inv_add(tx_size, tx_type, dq, eob, dest):
    """Dispatch by eob exactly as eb_vp9_idct*_add / eb_vp9_iht*_add (vp9_idct.c:111-190), calling the _c kernels."""
    R = ref_lib()
    N = dq.shape[1]
    dq = np.ascontiguousarray(dq)
    out = np.ascontiguousarray(dest.copy())
    for i in range(dq.shape[0]):
        e, a, d = int(eob[i]), C.c_void_p(dq[i].ctypes.data), C.c_void_p(out[i].ctypes.data)
        if tx_type != 0 and tx_size < 3:
            [R.eb_vp9_iht4x4_16_add_c, R.eb_vp9_iht8x8_64_add_c, R.eb_vp9_iht16x16_256_add_c][tx_size](a, d, N, tx_type)
        elif tx_size == 0:
            (R.eb_vp9_idct4x4_16_add_c if e > 1 else R.eb_vp9_idct4x4_1_add_c)(a, d, N)
        elif tx_size == 1:
            (R.eb_vp9_idct8x8_1_add_c if e == 1 else R.eb_vp9_idct8x8_12_add_c if e <= 12 else R.eb_vp9_idct8x8_64_add_c)(a, d, N)
        elif tx_size == 2:
            (R.eb_vp9_idct16x16_1_add_c if e == 1 else R.eb_vp9_idct16x16_10_add_c if e <= 10 else
             R.eb_vp9_idct16x16_38_add_c if e <= 38 else R.eb_vp9_idct16x16_256_add_c)(a, d, N)
        else:
            (R.eb_vp9_idct32x32_1_add_c if e == 1 else R.eb_vp9_idct32x32_34_add_c if e <= 34 else
             R.eb_vp9_idct32x32_135_add_c if e <= 135 else R.eb_vp9_idct32x32_1024_add_c)(a, d, N)
    return out


def residual_blocks(rng, n, N, kind):
    """Test residuals: 'uniform' in [-255,255], 'smooth' (low-pass, small eob after quantisation), 'extreme' (+-255 patterns)."""
    if kind == "uniform":
        return rng.integers(-255, 256, size=(n, N, N), dtype=np.int16)
    if kind == "extreme":
        b = rng.choice(np.array([-255, 255, 0], dtype=np.int16), size=(n, N, N))
        b[0] = 255
        b[1 % n] = -255
        return b
    x = rng.normal(0, 1, size=(n, N, N))
    for _ in range(3):
        x = (x + np.roll(x, 1, 1) + np.roll(x, 1, 2) + np.roll(x, -1, 1) + np.roll(x, -1, 2)) / 5
    x = x / (np.abs(x).max() + 1e-9) * rng.integers(4, 200, size=(n, 1, 1))
    return np.clip(np.rint(x), -255, 255).astype(np.int16)


# ---------------------------------------------------------------------------------------------
# deblocking helpers
# ---------------------------------------------------------------------------------------------
class LfPicture:
    """One loop-filtered picture dumped from the reference (oracle/ref_seam_check.c, dump mode): inputs + the
    reference's filtered planes."""

    def __init__(self, path):
        import gzip
        with (gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")) as f:
            data = f.read()
        magic, ver, self.lfm_stride, n_lfm = struct.unpack_from("<4I", data, 0)
        assert magic == 0x4c305653 and ver == 1
        off = 16
        self.params = api.LfParams.from_buffer_copy(data[off:off + C.sizeof(api.LfParams)])
        off += C.sizeof(api.LfParams)
        self.thr = np.frombuffer(data, dtype=api.LF_THRESH_DTYPE, count=64, offset=off).copy()
        off += 64 * 3
        self.lfm = np.frombuffer(data, dtype=api.LF_MASK_DTYPE, count=n_lfm, offset=off).copy()
        off += 160 * n_lfm
        W, Hh = self.params.width, self.params.height
        planes = []
        for n, shape in [(W * Hh, (Hh, W)), (W * Hh // 4, (Hh // 2, W // 2)), (W * Hh // 4, (Hh // 2, W // 2))] * 2:
            planes.append(np.frombuffer(data, dtype=np.uint8, count=n, offset=off).reshape(shape).copy())
            off += n
        self.y, self.u, self.v, self.ey, self.eu, self.ev = planes


def oracle_deblock(params, y, u, v, lfm, lfm_stride, thr):
    L = oracle()
    L.svtb200_oracle_deblock_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    y, u, v, m = y.copy(), u.copy(), v.copy(), lfm.copy()
    L.svtb200_oracle_deblock_picture(C.byref(params), y.ctypes.data, y.shape[1], u.ctypes.data, v.ctypes.data, u.shape[1],
                                     m.ctypes.data, lfm_stride, np.ascontiguousarray(thr).ctypes.data)
    return y, u, v


def emu_deblock(params, y, u, v, lfm, lfm_stride, thr):
    L = emu()
    L.lf_emu_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    y, u, v = y.copy(), u.copy(), v.copy()
    L.lf_emu_run(C.byref(params), y.ctypes.data, y.shape[1], u.ctypes.data, v.ctypes.data, u.shape[1],
                 np.ascontiguousarray(lfm).ctypes.data, lfm_stride, np.ascontiguousarray(thr).ctypes.data)
    return y, u, v


def lf_thresholds(sharpness=0):
    """update_sharpness + eb_vp9_loop_filter_init (vp9_loopfilter.c:221-262) for the 64 levels."""
    thr = np.zeros(64, dtype=api.LF_THRESH_DTYPE)
    for lvl in range(64):
        lim = lvl >> ((sharpness > 0) + (sharpness > 4))
        if sharpness > 0:
            lim = min(lim, 9 - sharpness)
        lim = max(lim, 1)
        thr[lvl] = (2 * (lvl + 2) + lim, lim, lvl >> 4)
    return thr


def synthetic_lf_case(rng, width, height, level):
    """A picture with a random but self-consistent transform-block structure: per 64x64 SB a random mix of 32x32, 16x16,
    8x8 and 4x4 transform blocks (chroma one size down), edge masks set the way eb_vp9_build_mask does (before
    eb_vp9_adjust_mask), some blocks left unfiltered (skip)."""
    mi_cols, mi_rows = width // 8, height // 8
    sbs_w, sbs_h = (mi_cols + 7) // 8, (mi_rows + 7) // 8
    lfm = np.zeros(sbs_w * sbs_h, dtype=api.LF_MASK_DTYPE)
    for sb in range(sbs_w * sbs_h):
        m = lfm[sb]
        m["lfl_y"][:] = level
        ly, ay, lu, au = [0] * 4, [0] * 4, [0] * 4, [0] * 4
        i4y = i4u = 0
        for q in range(4):  # 32x32 quadrants
            qr, qc = (q >> 1) * 4, (q & 1) * 4
            kind = rng.integers(0, 4)  # 0: one 32x32, 1: 16x16s, 2: 8x8s, 3: 4x4s
            if rng.random() < 0.15:
                continue  # skipped inter block: no edges at all
            size = [4, 2, 1, 1][kind]
            tx = [3, 2, 1, 0][kind]
            for br in range(qr, qr + 4, size):
                for bc in range(qc, qc + 4, size):
                    for k in range(size):
                        ly[tx] |= 1 << ((br + k) * 8 + bc)
                        ay[tx] |= 1 << (br * 8 + bc + k)
            if kind == 3:
                for br in range(qr, qr + 4):
                    for bc in range(qc, qc + 4):
                        i4y |= 1 << (br * 8 + bc)
            # chroma: 16x16 chroma region of this quadrant = 2x2 chroma blocks; transform one size down
            ctx_ = max(tx - 1, 0)
            csize = {3: 2, 2: 2, 1: 1, 0: 1}[tx]
            cr0, cc0 = qr // 2, qc // 2
            for br in range(cr0, cr0 + 2, csize):
                for bc in range(cc0, cc0 + 2, csize):
                    for k in range(csize):
                        lu[ctx_] |= 1 << ((br + k) * 4 + bc)
                        au[ctx_] |= 1 << (br * 4 + bc + k)
            if tx <= 1 and kind == 3:
                for br in range(cr0, cr0 + 2):
                    for bc in range(cc0, cc0 + 2):
                        i4u |= 1 << (br * 4 + bc)
        for k in range(4):
            m["left_y"][k], m["above_y"][k], m["left_uv"][k], m["above_uv"][k] = ly[k], ay[k], lu[k], au[k]
        m["int_4x4_y"], m["int_4x4_uv"] = i4y, i4u
    # blocky content so that all filter flavours trigger: smooth areas with small steps + textured areas
    base = rng.integers(0, 256, size=(height // 8 + 1, width // 8 + 1)).astype(np.float32)
    yimg = np.kron(base, np.ones((8, 8), np.float32))[:height, :width]
    yimg = yimg * 0.15 + 100 + rng.integers(-2, 3, size=yimg.shape) * (rng.random(yimg.shape) < 0.5)
    tex = rng.integers(0, 256, size=yimg.shape)
    sel = np.kron(rng.random((height // 32 + 1, width // 32 + 1)) < 0.3, np.ones((32, 32), bool))[:height, :width]
    yimg = np.where(sel, tex, yimg)
    y = np.clip(np.rint(yimg), 0, 255).astype(np.uint8)
    u = np.ascontiguousarray(y[::2, ::2] // 2 + 60).astype(np.uint8)
    v = np.ascontiguousarray(255 - u)
    P = api.LfParams()
    P.width, P.height, P.mi_rows, P.mi_cols, P.y_only, P.filter_level = width, height, mi_rows, mi_cols, 0, level
    return P, y, u, v, lfm, sbs_w


# ---------------------------------------------------------------------------------------------
# the reference's own hot-path functions (oracle/_ref/libSvtVp9Enc_refdrv.so, oracle/ref_path_driver.c)
# ---------------------------------------------------------------------------------------------
_refdrv = None


class RefdrvTuPass(C.Structure):
    """oracle/ref_path_driver.c: one transform pass of a picture (job table + planes + result arrays)"""
    _fields_ = [("tx_size", C.c_int32), ("n", C.c_int32), ("jobs", C.c_void_p), ("pl", api.TxPlanes), ("coeff", C.c_void_p), ("qcoeff", C.c_void_p),
                ("dqcoeff", C.c_void_p), ("eob", C.c_void_p), ("rblk", C.c_void_p), ("rate", C.c_void_p)]


class RefdrvPicture(C.Structure):
    """oracle/ref_path_driver.c: the hot-path work of one picture for refdrv_pool_run"""
    _fields_ = [("P", api.MeParams), ("cur", C.POINTER(api.Plane)), ("ref0", C.POINTER(api.Plane)), ("ref1", C.POINTER(api.Plane)), ("me_out", C.c_void_p),
                ("n_pass", C.c_int32), ("passes", RefdrvTuPass * 8), ("qrows", C.c_void_p), ("deblock", C.c_int32), ("lf", api.LfParams),
                ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("ys", C.c_int32), ("cs", C.c_int32), ("lfm", C.c_void_p), ("lfm_stride", C.c_int32),
                ("thr", C.c_void_p), ("n_blocks", C.c_int32), ("pred", api.PredPicture), ("token_costs_all", C.c_void_p)]


def have_refdrv():
    return os.path.exists(os.path.join(REF_DIR, "libSvtVp9Enc_refdrv.so"))


def refdrv(use_simd=1):
    global _refdrv
    if _refdrv is None:
        L = C.CDLL(os.path.join(REF_DIR, "libSvtVp9Enc_refdrv.so"))
        L.refdrv_init.argtypes = [C.c_int]
        L.refdrv_me_new.restype = C.c_void_p
        L.refdrv_me_range.argtypes = [C.c_void_p, C.POINTER(api.MeParams), C.POINTER(api.Plane), C.POINTER(api.Plane), C.POINTER(api.Plane),
                                      C.c_int, C.c_int, C.c_void_p]
        L.refdrv_txq.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(api.TxPlanes), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refdrv_zz_sad.argtypes = [C.c_void_p, C.POINTER(api.Plane), C.POINTER(api.Plane), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.refdrv_deblock.argtypes = [C.POINTER(api.LfParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.refdrv_pool_new.argtypes = [C.c_int, C.c_int]
        L.refdrv_pool_new.restype = C.c_void_p
        L.refdrv_pool_run.argtypes = [C.c_void_p, C.POINTER(RefdrvPicture), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.refdrv_pool_free.argtypes = [C.c_void_p]
        L.refdrv_init(use_simd)
        _refdrv = L
    return _refdrv


def refdrv_me_picture(P, cur, r0, r1, n_sb, threads=1):
    """The reference's motion_estimate_sb over a whole picture; cur/r0/r1 = (api.Plane * 3)."""
    from concurrent.futures import ThreadPoolExecutor
    L = refdrv()
    out = np.zeros(n_sb * 85, dtype=api.ME_RESULT_DTYPE)
    chunk = max(1, (n_sb + threads * 4 - 1) // (threads * 4))
    ranges = [(s, min(n_sb, s + chunk)) for s in range(0, n_sb, chunk)]

    def work(rg):
        h = L.refdrv_me_new()
        L.refdrv_me_range(h, C.byref(P), cur, r0, r1, rg[0], rg[1], out.ctypes.data)

    if threads == 1:
        for rg in ranges:
            work(rg)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, ranges))
    return out


def aligned_zeros(shape, dtype, align=64):
    """The reference's SIMD kernels use aligned loads/stores on coefficient buffers and block rows."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    raw = np.zeros(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dt).reshape(shape)


def aligned_copy(a, align=64):
    out = aligned_zeros(a.shape, a.dtype, align)
    out[...] = a
    return out


def refdrv_txq(tx_size, jobs, src, pred, recon, rows):
    L = refdrv()
    n, N = len(jobs), 4 << tx_size
    jobs = np.ascontiguousarray(jobs)
    src, pred = aligned_copy(src), aligned_copy(pred)
    rec = None if recon is None else aligned_copy(recon)
    pl = api.TxPlanes()
    pl.src, pl.src_stride, pl.src_bytes = src.ctypes.data, src.shape[1], src.nbytes
    pl.pred, pl.pred_stride, pl.pred_bytes = pred.ctypes.data, pred.shape[1], pred.nbytes
    if rec is not None:
        pl.recon, pl.recon_stride, pl.recon_bytes = rec.ctypes.data, rec.shape[1], rec.nbytes
    coeff, q, dq = (aligned_zeros((n, N, N), np.int16) for _ in range(3))
    eob = np.zeros(n, dtype=np.uint16)
    rows = np.ascontiguousarray(rows)
    L.refdrv_txq(tx_size, 0, n, jobs.ctypes.data, C.byref(pl), rows.ctypes.data, coeff.ctypes.data, q.ctypes.data, dq.ctypes.data, eob.ctypes.data)
    return coeff, q, dq, eob, rec


def refdrv_deblock(P, y, u, v, lfm, lfm_stride, thr, pad=16):
    """Runs the reference filter on padded copies (it also filters the half block row/column below/right of odd-sized
    pictures, inside its padding)."""
    L = refdrv()
    yy = np.ascontiguousarray(np.pad(y, ((0, pad), (0, pad)), mode="edge"))
    uu = np.ascontiguousarray(np.pad(u, ((0, pad), (0, pad)), mode="edge"))
    vv = np.ascontiguousarray(np.pad(v, ((0, pad), (0, pad)), mode="edge"))
    L.refdrv_deblock(C.byref(P), yy.ctypes.data, yy.shape[1], uu.ctypes.data, vv.ctypes.data, uu.shape[1],
                     np.ascontiguousarray(lfm).ctypes.data, lfm_stride, np.ascontiguousarray(thr).ctypes.data)
    return yy[:y.shape[0], :y.shape[1]], uu[:u.shape[0], :u.shape[1]], vv[:v.shape[0], :v.shape[1]]


def cudart_path():
    """libcudart.so.12 that libsvtvp9_b200.so resolves to in this image (for child processes that are not python)."""
    import glob
    for pat in ("/usr/local/cuda/lib64/libcudart.so.12*", os.path.join(os.path.dirname(os.__file__), "site-packages", "nvidia", "cuda_runtime", "lib", "libcudart.so.12*")):
        hits = sorted(glob.glob(pat))
        if hits:
            return hits[0]
    try:
        import nvidia.cuda_runtime as m
        return os.path.join(os.path.dirname(m.__file__), "lib", "libcudart.so.12")
    except Exception:
        return ""


def oracle_zz_sad(cur_structs, prev_structs, w, h, th_shift):
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    scores, sads = np.zeros(n_sb, dtype=np.uint8), np.zeros(n_sb, dtype=np.uint32)
    oracle().svtb200_oracle_zz_sad_picture(C.byref(cur_structs[2]), C.byref(prev_structs[0]), w, h, th_shift, scores.ctypes.data, sads.ctypes.data)
    return scores, sads


def refdrv_zz_sad(cur_structs, prev_structs, w, h, input_resolution):
    """The reference's compute_zz_sad; input_resolution 0..3 -> non_moving_th_shift {4, 2, 0, 0}."""
    L = refdrv()
    n_sb = ((w + 63) // 64) * ((h + 63) // 64)
    scores = np.zeros(n_sb, dtype=np.uint8)
    L.refdrv_zz_sad(L.refdrv_me_new(), C.byref(cur_structs[2]), C.byref(prev_structs[0]), w, h, input_resolution, scores.ctypes.data)
    return scores
