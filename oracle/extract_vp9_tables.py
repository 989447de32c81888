"""Extracts the VP9 bitstream-constant tables the transform/quantisation path needs (scan orders per transform size and
type, DC/AC quantiser step lookups) from the reference library built by oracle/build_ref.sh and stores them in
svt-vp9_b200/data/vp9_tables.npz.  These are constants of the VP9 format (the reference keeps them in
Source/Lib/VPX/vp9_scan.c and vp9_quant_common.c); at run time the reference-side shim passes its own copies through
svtb200_set_scan() / SvtB200QuantRow, the npz only feeds tests and bench.  Run where /root/reference exists."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ScanOrder(C.Structure):
    _fields_ = [("scan", C.POINTER(C.c_int16)), ("iscan", C.POINTER(C.c_int16)), ("neighbors", C.POINTER(C.c_int16))]


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libSvtVp9Enc_ref.so"))
    so = ((ScanOrder * 4) * 4).in_dll(L, "eb_vp9_scan_orders")
    out = {}
    for ts in range(4):
        n = (4 << ts) ** 2
        for tt in range(4):
            out["scan_%d_%d" % (ts, tt)] = np.array([so[ts][tt].scan[i] for i in range(n)], dtype=np.int16)
            out["iscan_%d_%d" % (ts, tt)] = np.array([so[ts][tt].iscan[i] for i in range(n)], dtype=np.int16)
    L.eb_vp9_dc_quant.restype = C.c_int16
    L.eb_vp9_ac_quant.restype = C.c_int16
    out["dc_q"] = np.array([L.eb_vp9_dc_quant(q, 0, 8) for q in range(256)], dtype=np.int16)
    out["ac_q"] = np.array([L.eb_vp9_ac_quant(q, 0, 8) for q in range(256)], dtype=np.int16)
    # rate estimation (coeff_rate_estimate): scan neighbours, value -> token / cost tables, the token costs of the default
    # coefficient probabilities -- through the path driver, which reads the reference's own tables (ref_path_driver.c)
    D = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libSvtVp9Enc_refdrv.so"))
    for ts in range(4):
        n = (4 << ts) ** 2
        for tt in range(4):
            nb = np.zeros(2 * (n + 1), dtype=np.int16)
            D.refdrv_scan_neighbors(ts, tt, C.c_void_p(nb.ctypes.data))
            out["nb_%d_%d" % (ts, tt)] = nb
    vtok, vcost = np.zeros(133, dtype=np.int16), np.zeros(133, dtype=np.int32)
    c6lo, c6hi = np.zeros(256, dtype=np.int16), np.zeros(64, dtype=np.uint16)
    D.refdrv_value_tables(*(C.c_void_p(a.ctypes.data) for a in (vtok, vcost, c6lo, c6hi)))
    out.update(rate_vtok=vtok, rate_vcost=vcost, rate_c6lo=c6lo, rate_c6hi=c6hi)
    tc = np.zeros((4, 2, 2, 6, 2, 6, 12), dtype=np.uint32)
    assert D.refdrv_token_costs(C.c_void_p(tc.ctypes.data)) == tc.nbytes
    out["token_costs_default"] = tc
    dst = os.path.join(ROOT, "svt-vp9_b200", "data", "vp9_tables.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
