"""ctypes mirror of include/svtb200.h -- the host-side entry points the tests and bench call.

The CUDA library is the only implementation behind these calls: there is no CPU fallback, and
loading fails loudly when libsvtvp9_b200.so has not been built (see svt-vp9_b200/build.py).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVTB200_LIB") or os.path.join(_HERE, "libsvtvp9_b200.so")   # SVTB200_LIB: an experimental build of the same sources (tools)


class Plane(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("stride", C.c_int32), ("origin_x", C.c_int32), ("origin_y", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("padded_height", C.c_int32)]


class MeParams(C.Structure):
    _fields_ = [("luma_width", C.c_int32), ("luma_height", C.c_int32), ("num_lists", C.c_int32),
                ("temporal_layer_index", C.c_int32), ("hierarchical_levels", C.c_int32),
                ("ref_poc", C.c_uint64 * 2),
                ("enable_hme_flag", C.c_int32), ("enable_hme_level0", C.c_int32), ("enable_hme_level1", C.c_int32),
                ("enable_hme_level2", C.c_int32), ("single_hme_quadrant", C.c_int32),
                ("number_hme_search_region_in_width", C.c_int32), ("number_hme_search_region_in_height", C.c_int32),
                ("hme_level0_total_search_area_width", C.c_int32), ("hme_level0_total_search_area_height", C.c_int32),
                ("hme_level0_search_area_in_width", C.c_int32 * 2), ("hme_level0_search_area_in_height", C.c_int32 * 2),
                ("hme_level1_search_area_in_width", C.c_int32 * 2), ("hme_level1_search_area_in_height", C.c_int32 * 2),
                ("hme_level2_search_area_in_width", C.c_int32 * 2), ("hme_level2_search_area_in_height", C.c_int32 * 2),
                ("hme_l0_multiplier_x", C.c_int32), ("hme_l0_multiplier_y", C.c_int32),
                ("search_area_width", C.c_int32), ("search_area_height", C.c_int32),
                ("fractional_search_method", C.c_int32), ("fractional_search_model", C.c_int32),
                ("fractional_search64x64", C.c_int32), ("cu8x8_mode", C.c_int32), ("cu16x16_mode", C.c_int32),
                ("rate_control_mode", C.c_int32)]


ME_RESULT_DTYPE = np.dtype([("x_mv_l0", "<i2"), ("y_mv_l0", "<i2"), ("x_mv_l1", "<i2"), ("y_mv_l1", "<i2"),
                            ("cand", [("distortion", "<u4"), ("direction", "<u4")], 3),
                            ("total", "u1"), ("pad", "u1", 7)])
assert ME_RESULT_DTYPE.itemsize == 40

ME_TRACE_DTYPE = np.dtype([("hme_center_x", "<i2", 2), ("hme_center_y", "<i2", 2), ("sa_ox", "<i2", 2), ("sa_oy", "<i2", 2),
                           ("sa_w", "<i2", 2), ("sa_h", "<i2", 2), ("fullpel_sad", "<u4", (2, 85)),
                           ("fullpel_mv", "<u4", (2, 85)), ("subpel_enable", "u1", 2), ("pad", "u1", 2)])
assert ME_TRACE_DTYPE.itemsize == 1388


class QuantRow(C.Structure):
    _fields_ = [("zbin", C.c_int16 * 2), ("round", C.c_int16 * 2), ("quant", C.c_int16 * 2),
                ("quant_shift", C.c_int16 * 2), ("dequant", C.c_int16 * 2)]


QUANT_ROW_DTYPE = np.dtype([("zbin", "<i2", 2), ("round", "<i2", 2), ("quant", "<i2", 2), ("quant_shift", "<i2", 2),
                            ("dequant", "<i2", 2)])

TX_JOB_DTYPE = np.dtype([("src_offset", "<u4"), ("pred_offset", "<u4"), ("recon_offset", "<u4"), ("qrow", "<u2"),
                         ("tx_type", "u1"), ("flags", "u1")])
assert TX_JOB_DTYPE.itemsize == 16


class TxPlanes(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_stride", C.c_int32), ("src_bytes", C.c_size_t),
                ("pred", C.c_void_p), ("pred_stride", C.c_int32), ("pred_bytes", C.c_size_t),
                ("recon", C.c_void_p), ("recon_stride", C.c_int32), ("recon_bytes", C.c_size_t)]


class TxBatch(C.Structure):
    """SvtB200TxBatch: one fused batch of svtb200_txq_multi_dev (device pointers)"""
    _fields_ = [("tx_size", C.c_int32), ("n", C.c_int32), ("jobs", C.c_void_p), ("planes", TxPlanes), ("coeff", C.c_void_p), ("qcoeff", C.c_void_p),
                ("dqcoeff", C.c_void_p), ("eob", C.c_void_p)]


LF_MASK_DTYPE = np.dtype([("left_y", "<u8", 4), ("above_y", "<u8", 4), ("int_4x4_y", "<u8"), ("left_uv", "<u2", 4),
                          ("above_uv", "<u2", 4), ("int_4x4_uv", "<u2"), ("lfl_y", "u1", 64), ("pad", "u1", 6)])
assert LF_MASK_DTYPE.itemsize == 160
LF_MODE_INFO_DTYPE = np.dtype([("sb_type", "u1"), ("tx_size", "u1"), ("skip_inter", "u1"), ("filter_level", "u1")])
LF_THRESH_DTYPE = np.dtype([("mblim", "u1"), ("lim", "u1"), ("hev_thr", "u1")])


class LfParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mi_rows", C.c_int32), ("mi_cols", C.c_int32),
                ("y_only", C.c_int32), ("filter_level", C.c_int32)]


class PredBlock(C.Structure):
    """SvtB200PredBlock: one inter_predictor() call (device pointers for the product, host pointers for the oracle)."""
    _fields_ = [("ref", C.c_void_p * 2), ("dst", C.c_void_p), ("ref_stride", C.c_int32 * 2), ("x", C.c_int32 * 2), ("y", C.c_int32 * 2),
                ("dst_stride", C.c_int32), ("subpel_x", C.c_uint8 * 2), ("subpel_y", C.c_uint8 * 2), ("w", C.c_uint8), ("h", C.c_uint8),
                ("compound", C.c_uint8), ("reserved", C.c_uint8)]


INTER_BLOCK_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("bw", "u1"), ("bh", "u1"), ("list", "u1", 2), ("mv", "<i2", (2, 2))])
assert INTER_BLOCK_DTYPE.itemsize == 16


class PredPicture(C.Structure):
    """SvtB200PredPicture: reference planes of both lists, prediction planes (device pointers for the product, host for the oracle)."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ref", (C.c_void_p * 3) * 2), ("ref_stride", (C.c_int32 * 3) * 2),
                ("dst", C.c_void_p * 3), ("dst_stride", C.c_int32 * 3), ("use_subpel_flag", C.c_int32)]


class IntraBlock(C.Structure):
    """SvtB200IntraBlock: one intra predictor call (device pointers for the product, host pointers for the oracle)."""
    _fields_ = [("above", C.c_void_p), ("left", C.c_void_p), ("dst", C.c_void_p), ("dst_stride", C.c_int32), ("mode", C.c_uint8), ("tx_size", C.c_uint8),
                ("have_left", C.c_uint8), ("have_top", C.c_uint8)]


def plane_of(arr, origin_x, origin_y, width, height):
    """Describe a padded 2-D uint8 numpy array (C-contiguous) as an SvtB200Plane."""
    assert arr.dtype == np.uint8 and arr.flags["C_CONTIGUOUS"] and arr.ndim == 2
    p = Plane()
    p.buf = arr.ctypes.data
    p.stride = arr.shape[1]
    p.origin_x, p.origin_y, p.width, p.height, p.padded_height = origin_x, origin_y, width, height, arr.shape[0]
    return p


_lib = None


def lib():
    """Load libsvtvp9_b200.so; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA extension %s is not built; run python -c 'import __graft_entry__ as g; g.build()'" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.svtb200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.svtb200_destroy.argtypes = [C.c_void_p]
        L.svtb200_last_error.argtypes = [C.c_void_p]
        L.svtb200_last_error.restype = C.c_char_p
        L.svtb200_sync.argtypes = [C.c_void_p]
        L.svtb200_launch_count.argtypes = [C.c_void_p]
        L.svtb200_launch_count.restype = C.c_uint64
        L.svtb200_stream.argtypes = [C.c_void_p]
        L.svtb200_stream.restype = C.c_void_p
        L.svtb200_host_alloc.argtypes = [C.c_size_t]
        L.svtb200_host_alloc.restype = C.c_void_p
        L.svtb200_host_free.argtypes = [C.c_void_p]
        L.svtb200_dev_alloc.argtypes = [C.c_void_p, C.c_size_t]
        L.svtb200_dev_alloc.restype = C.c_void_p
        L.svtb200_dev_free.argtypes = [C.c_void_p, C.c_void_p]
        L.svtb200_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.svtb200_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.svtb200_picture_upload.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Plane), C.POINTER(Plane), C.POINTER(Plane)]
        L.svtb200_picture_upload_luma.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Plane), C.c_int32, C.c_int32, C.c_int]
        L.svtb200_picture_upload_luma_staged.argtypes = L.svtb200_picture_upload_luma.argtypes
        L.svtb200_picture_release.argtypes = [C.c_void_p, C.c_uint64]
        L.svtb200_picture_touch.argtypes = [C.c_void_p, C.c_uint64]
        L.svtb200_memcpy_h2d_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.svtb200_memcpy_d2h_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.svtb200_memcpy_h2d_2d_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.svtb200_pad_plane_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.svtb200_select_compute.argtypes = [C.c_void_p, C.c_int]
        L.svtb200_marker.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.svtb200_stream_wait.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.svtb200_fence.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.svtb200_wait.argtypes = [C.c_void_p, C.c_uint64]
        L.svtb200_picture_plane_dev.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Plane)]
        L.svtb200_me_picture.argtypes = [C.c_void_p, C.POINTER(MeParams), C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
        L.svtb200_me_picture_dev.argtypes = [C.c_void_p, C.POINTER(MeParams), C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p]
        L.svtb200_zz_sad_picture.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.svtb200_zz_sad_picture_dev.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.svtb200_me_enable_trace.argtypes = [C.c_void_p, C.c_int]
        L.svtb200_me_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.svtb200_inter_predict_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.svtb200_inter_predict_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.svtb200_intra_predict_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.svtb200_inter_blocks_expand_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(PredPicture), C.c_void_p]
        L.svtb200_me_best_blocks_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.svtb200_inter_chain_dev.argtypes = [C.c_void_p, C.c_void_p]
        L.svtb200_noise_filter_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.svtb200_noise_detect_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svtb200_edge_detection_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svtb200_edge_cu_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p]
        L.svtb200_chroma_block_mean_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.svtb200_chroma_histogram_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.svtb200_full_distortion_batch_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


COMPUTE, UPLOAD, DOWNLOAD, COMPUTE2, COMPUTE3 = 0, 1, 2, 3, 4


class Context:
    """One SvtB200Ctx (one CUDA device + stream)."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.svtb200_create(device, C.byref(h))
        if rc != 0:
            raise RuntimeError("svtb200_create(device=%d) failed: %d" % (device, rc))
        self.h = h

    def close(self):
        if self.h:
            self.L.svtb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %d (%s)" % (what, rc, self.L.svtb200_last_error(self.h).decode()))

    def sync(self):
        self.check(self.L.svtb200_sync(self.h), "sync")

    def launches(self):
        return int(self.L.svtb200_launch_count(self.h))

    # ---- S1
    def picture_upload(self, pic_id, full, quarter, sixteenth):
        q = C.byref(quarter) if quarter is not None else None
        s = C.byref(sixteenth) if sixteenth is not None else None
        self.check(self.L.svtb200_picture_upload(self.h, pic_id, C.byref(full), q, s), "picture_upload")

    def picture_upload_luma(self, pic_id, full, quarter_pad=32, sixteenth_pad=16, make_quarter=True, staged=False):
        """Only the picture area of the padded luma crosses PCIe; padding and both decimations are made on the device.
        staged: through the context's ring of page-locked slots (pageable caller memory)."""
        f = self.L.svtb200_picture_upload_luma_staged if staged else self.L.svtb200_picture_upload_luma
        self.check(f(self.h, pic_id, C.byref(full), quarter_pad, sixteenth_pad, int(make_quarter)), "picture_upload_luma")

    def picture_plane(self, pic_id, which):
        """Download one resident plane (0 full, 1 quarter, 2 sixteenth) -> (array (rows, stride), Plane descriptor)."""
        d = Plane()
        self.check(self.L.svtb200_picture_plane_dev(self.h, pic_id, which, C.byref(d)), "picture_plane_dev")
        out = np.zeros((d.padded_height, d.stride), dtype=np.uint8)
        self.check(self.L.svtb200_memcpy_d2h(self.h, out.ctypes.data, d.buf, out.nbytes), "memcpy_d2h")
        return out, d

    # ---- the three queues of a context (include/svtb200.h): ordering between them
    def marker(self, stream):
        t = C.c_uint64()
        self.check(self.L.svtb200_marker(self.h, stream, C.byref(t)), "marker")
        return t.value

    def stream_wait(self, stream, ticket):
        self.check(self.L.svtb200_stream_wait(self.h, stream, ticket), "stream_wait")

    def select_compute(self, which):
        self.check(self.L.svtb200_select_compute(self.h, which), "select_compute")

    def fence(self, src, dst):
        self.check(self.L.svtb200_fence(self.h, src, dst), "fence")

    def wait(self, ticket):
        self.check(self.L.svtb200_wait(self.h, ticket), "wait")

    def picture_release(self, pic_id):
        self.L.svtb200_picture_release(self.h, pic_id)

    def me_picture(self, params, cur_id, ref_ids, want_rcme=False):
        n_sb = ((params.luma_width + 63) // 64) * ((params.luma_height + 63) // 64)
        out = np.zeros(n_sb * 85, dtype=ME_RESULT_DTYPE)
        rcme = np.zeros(n_sb, dtype=np.uint32)
        ids = (C.c_uint64 * 2)(ref_ids[0], ref_ids[1] if len(ref_ids) > 1 else ref_ids[0])
        self.check(self.L.svtb200_me_picture(self.h, C.byref(params), cur_id, ids, out.ctypes.data,
                                             rcme.ctypes.data if want_rcme else None), "me_picture")
        return (out, rcme) if want_rcme else out

    def zz_sad_picture(self, cur_id, prev_id, width, height, th_shift):
        """compute_zz_sad for a whole picture -> (scores u8[n_sb], sads u32[n_sb])."""
        n_sb = ((width + 63) // 64) * ((height + 63) // 64)
        scores, sads = np.zeros(n_sb, dtype=np.uint8), np.zeros(n_sb, dtype=np.uint32)
        self.check(self.L.svtb200_zz_sad_picture(self.h, cur_id, prev_id, width, height, th_shift, scores.ctypes.data, sads.ctypes.data), "zz_sad_picture")
        return scores, sads

    def block_stats_picture(self, pic_id, width, height):
        """compute_block_mean_compute_variance for every SB -> (y_mean u8[n_sb, 85], variance u16[n_sb, 85])."""
        n_sb = ((width + 63) // 64) * ((height + 63) // 64)
        ym, var = np.zeros((n_sb, 85), dtype=np.uint8), np.zeros((n_sb, 85), dtype=np.uint16)
        self.L.svtb200_block_stats_picture.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        self.check(self.L.svtb200_block_stats_picture(self.h, pic_id, width, height, ym.ctypes.data, var.ctypes.data), "block_stats_picture")
        return ym, var

    def luma_histogram_picture(self, pic_id, regions_w=4, regions_h=4):
        """-> (hist u32[rw, rh, 256], avg u8[rw, rh], sum16) of the resident picture's 1/16 plane"""
        hist = np.zeros((regions_w, regions_h, 256), dtype=np.uint32)
        avg = np.zeros((regions_w, regions_h), dtype=np.uint8)
        s16 = C.c_uint64()
        self.L.svtb200_luma_histogram_picture.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        self.check(self.L.svtb200_luma_histogram_picture(self.h, pic_id, regions_w, regions_h, hist.ctypes.data, avg.ctypes.data, C.byref(s16)), "luma_histogram_picture")
        return hist, avg, s16.value

    def me_trace(self, n_sb):
        tr = np.zeros(n_sb, dtype=ME_TRACE_DTYPE)
        self.check(self.L.svtb200_me_read_trace(self.h, tr.ctypes.data, tr.nbytes), "me_read_trace")
        return tr


# ---- S3: transform / quantisation -------------------------------------------------------------------------------
def _bind_txq(L):
    if getattr(L, "_txq_bound", False):
        return
    vp = C.c_void_p
    L.svtb200_set_scan.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.svtb200_fwd_txfm_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.svtb200_fwd_txfm_batch_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.svtb200_quantize_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    L.svtb200_quantize_batch_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    L.svtb200_inv_txfm_add_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.svtb200_inv_txfm_add_batch_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.svtb200_txq_batch.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(TxPlanes), vp, C.c_int, vp, vp, vp, vp]
    L.svtb200_txq_batch_dev.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(TxPlanes), vp, C.c_int, vp, vp, vp, vp]
    L.svtb200_txq_multi_dev.argtypes = [vp, C.POINTER(TxBatch), C.c_int, vp, C.c_int]
    L.svtb200_coeff_pack_dev.argtypes = [vp, C.POINTER(TxBatch), C.c_int, vp, C.c_uint32, vp]
    L._txq_bound = True


def _ctx_register_scans(self, tables):
    """tables: mapping with scan_<ts>_<tt> / iscan_<ts>_<tt> int16 arrays (svt-vp9_b200/data/vp9_tables.npz)."""
    _bind_txq(self.L)
    for ts in range(4):
        for tt in range(4):
            s = np.ascontiguousarray(tables["scan_%d_%d" % (ts, tt)], dtype=np.int16)
            i = np.ascontiguousarray(tables["iscan_%d_%d" % (ts, tt)], dtype=np.int16)
            self.check(self.L.svtb200_set_scan(self.h, ts, tt, s.ctypes.data, i.ctypes.data), "set_scan")


def _ctx_fwd_txfm(self, tx_size, tx_type, resid, partial32=0):
    _bind_txq(self.L)
    resid = np.ascontiguousarray(resid, dtype=np.int16)
    out = np.zeros_like(resid)
    self.check(self.L.svtb200_fwd_txfm_batch(self.h, tx_size, tx_type, partial32, resid.shape[0], resid.ctypes.data, out.ctypes.data), "fwd_txfm_batch")
    return out


def _ctx_quantize(self, tx_size, tx_type, coeff, rows, row_idx=None):
    _bind_txq(self.L)
    coeff = np.ascontiguousarray(coeff, dtype=np.int16)
    n = coeff.shape[0]
    q, dq, eob = np.zeros_like(coeff), np.zeros_like(coeff), np.zeros(n, dtype=np.uint16)
    rows = np.ascontiguousarray(rows)
    idx = None if row_idx is None else np.ascontiguousarray(row_idx, dtype=np.uint16)
    self.check(self.L.svtb200_quantize_batch(self.h, tx_size, tx_type, n, coeff.ctypes.data, rows.ctypes.data, len(rows),
                                             None if idx is None else idx.ctypes.data, q.ctypes.data, dq.ctypes.data, eob.ctypes.data), "quantize_batch")
    return q, dq, eob


def _ctx_inv_add(self, tx_size, tx_type, dq, eob, dest):
    _bind_txq(self.L)
    dq = np.ascontiguousarray(dq, dtype=np.int16)
    out = np.ascontiguousarray(dest.copy(), dtype=np.uint8)
    eob = np.ascontiguousarray(eob, dtype=np.uint16)
    self.check(self.L.svtb200_inv_txfm_add_batch(self.h, tx_size, tx_type, dq.shape[0], dq.ctypes.data, eob.ctypes.data, out.ctypes.data), "inv_txfm_add_batch")
    return out


def _ctx_txq(self, tx_size, jobs, src, pred, recon, rows, want_coeff=True):
    """jobs: TX_JOB_DTYPE array; src/pred/recon: 2-D uint8 planes (recon may be None). Returns coeff, q, dq, eob, recon."""
    _bind_txq(self.L)
    n, N = len(jobs), 4 << tx_size
    jobs = np.ascontiguousarray(jobs)
    src, pred = np.ascontiguousarray(src), np.ascontiguousarray(pred)
    rec = None if recon is None else np.ascontiguousarray(recon.copy())
    pl = TxPlanes()
    pl.src, pl.src_stride, pl.src_bytes = src.ctypes.data, src.shape[1], src.nbytes
    pl.pred, pl.pred_stride, pl.pred_bytes = pred.ctypes.data, pred.shape[1], pred.nbytes
    if rec is not None:
        pl.recon, pl.recon_stride, pl.recon_bytes = rec.ctypes.data, rec.shape[1], rec.nbytes
    coeff = np.zeros((n, N, N), dtype=np.int16)
    q, dq, eob = np.zeros_like(coeff), np.zeros_like(coeff), np.zeros(n, dtype=np.uint16)
    rows = np.ascontiguousarray(rows)
    self.check(self.L.svtb200_txq_batch(self.h, tx_size, n, jobs.ctypes.data, C.byref(pl), rows.ctypes.data, len(rows),
                                        coeff.ctypes.data if want_coeff else None, q.ctypes.data, dq.ctypes.data, eob.ctypes.data), "txq_batch")
    return coeff, q, dq, eob, rec


def _ctx_txq_multi(self, batches, rows, pack=None):
    """batches: list of (tx_size, jobs, src, pred, recon) with 2-D uint8 planes (each batch its own planes); ONE launch through
    svtb200_txq_multi_dev.  Returns a list of (q, dq, eob, recon) per batch.  pack: a dict that receives "stream" / "offsets" of
    svtb200_coeff_pack_dev over the same batches ("host": True -> written by the kernel into page-locked host memory)."""
    _bind_txq(self.L)
    L, h = self.L, self.h
    rows = np.ascontiguousarray(rows)
    dev, arr, outs = [], (TxBatch * len(batches))(), []

    def up(a):
        d = L.svtb200_dev_alloc(h, max(a.nbytes, 16))
        self.check(L.svtb200_memcpy_h2d(h, d, a.ctypes.data, a.nbytes), "h2d")
        dev.append(d)
        return d
    try:
        d_rows = up(rows)
        for k, (ts, jobs, src, pred, recon) in enumerate(batches):
            N, n = 4 << ts, len(jobs)
            jobs, src, pred, recon = (np.ascontiguousarray(a) for a in (jobs, src, pred, recon))
            b = arr[k]
            b.tx_size, b.n, b.jobs = ts, n, up(jobs)
            b.planes.src, b.planes.src_stride, b.planes.src_bytes = up(src), src.shape[1], src.nbytes
            b.planes.pred, b.planes.pred_stride, b.planes.pred_bytes = up(pred), pred.shape[1], pred.nbytes
            b.planes.recon, b.planes.recon_stride, b.planes.recon_bytes = up(recon), recon.shape[1], recon.nbytes
            q, dq, eob = np.zeros((n, N, N), np.int16), np.zeros((n, N, N), np.int16), np.zeros(n, np.uint16)
            b.coeff, b.qcoeff, b.dqcoeff, b.eob = None, up(q), up(dq), up(eob)
            outs.append((q, dq, eob, recon.copy(), b))
        self.check(L.svtb200_txq_multi_dev(h, arr, len(batches), d_rows, len(rows)), "txq_multi_dev")
        if pack is not None:
            n_total = sum(len(b[1]) for b in batches)
            cap = pack.get("capacity", sum(len(b[1]) * (4 << b[0]) ** 2 for b in batches))
            stream, offsets = np.zeros(max(cap, 2), np.int16), np.zeros(n_total + 1, np.uint32)
            if pack.get("host"):
                hs, ho = L.svtb200_host_alloc(stream.nbytes), L.svtb200_host_alloc(offsets.nbytes)
                C.memset(hs, 0, stream.nbytes)
                self.check(L.svtb200_coeff_pack_dev(h, arr, len(batches), hs, cap, ho), "coeff_pack_dev")
                self.sync()
                C.memmove(stream.ctypes.data, hs, stream.nbytes)
                C.memmove(offsets.ctypes.data, ho, offsets.nbytes)
                L.svtb200_host_free(hs)
                L.svtb200_host_free(ho)
            else:
                ds, do = up(stream), up(offsets)
                self.check(L.svtb200_coeff_pack_dev(h, arr, len(batches), ds, cap, do), "coeff_pack_dev")
                self.sync()
                self.check(L.svtb200_memcpy_d2h(h, stream.ctypes.data, ds, stream.nbytes), "d2h")
                self.check(L.svtb200_memcpy_d2h(h, offsets.ctypes.data, do, offsets.nbytes), "d2h")
            pack["stream"], pack["offsets"] = stream, offsets
        self.sync()
        res = []
        for q, dq, eob, rec, b in outs:
            for a, d in ((q, b.qcoeff), (dq, b.dqcoeff), (eob, b.eob), (rec, b.planes.recon)):
                self.check(L.svtb200_memcpy_d2h(h, a.ctypes.data, d, a.nbytes), "d2h")
            res.append((q, dq, eob, rec))
        return res
    finally:
        for d in dev:
            L.svtb200_dev_free(h, d)


Context.txq_multi = _ctx_txq_multi
Context.register_scans = _ctx_register_scans
Context.fwd_txfm = _ctx_fwd_txfm
Context.quantize = _ctx_quantize
Context.inv_txfm_add = _ctx_inv_add
Context.txq = _ctx_txq


# ---- S2: deblocking -------------------------------------------------------------------------------------------------
def _ctx_build_masks(self, mi, lfm_stride=None):
    """mi: (mi_rows, mi_cols) array of LF_MODE_INFO_DTYPE -> (sbs_h * lfm_stride) LF_MASK_DTYPE (before eb_vp9_adjust_mask)."""
    mi = np.ascontiguousarray(mi)
    rows, cols = mi.shape
    sbs_w, sbs_h = (cols + 7) // 8, (rows + 7) // 8
    stride = lfm_stride or sbs_w
    out = np.zeros(sbs_h * stride, dtype=LF_MASK_DTYPE)
    self.L.svtb200_build_masks.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    self.check(self.L.svtb200_build_masks(self.h, rows, cols, mi.ctypes.data, cols, out.ctypes.data, stride), "build_masks")
    return out


Context.build_masks = _ctx_build_masks


def _ctx_deblock(self, params, y, u, v, lfm, lfm_stride, thr):
    """In-place eb_vp9_loop_filter_frame on host planes (2-D uint8 arrays at picture origin). Returns (y, u, v) copies."""
    L = self.L
    if not getattr(L, "_lf_bound", False):
        vp = C.c_void_p
        L.svtb200_deblock_picture.argtypes = [vp, C.POINTER(LfParams), vp, C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, vp]
        L.svtb200_deblock_picture_dev.argtypes = [vp, C.POINTER(LfParams), vp, C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, vp]
        L._lf_bound = True
    y, u, v = np.ascontiguousarray(y.copy()), np.ascontiguousarray(u.copy()), np.ascontiguousarray(v.copy())
    lfm = np.ascontiguousarray(lfm)
    thr = np.ascontiguousarray(thr)
    self.check(L.svtb200_deblock_picture(self.h, C.byref(params), y.ctypes.data, y.shape[1], u.ctypes.data, v.ctypes.data, u.shape[1],
                                         lfm.ctypes.data, lfm_stride, thr.ctypes.data), "deblock_picture")
    return y, u, v


Context.deblock = _ctx_deblock


# ---- next to S3: prediction of the encode pass, full-loop distortion -------------------------------------------------
def _ctx_inter_predict(self, blocks):
    """blocks: ctypes array of PredBlock with DEVICE pointers; asynchronous on the current compute queue."""
    self.check(self.L.svtb200_inter_predict_batch(self.h, C.addressof(blocks), len(blocks)), "inter_predict_batch")


def _ctx_full_distortion(self, tx_size, coeff, recon_coeff, eob):
    """coeff / recon_coeff: (n, N*N) int16 host arrays, eob: (n,) uint16 -> (n, 2) uint64 [residual, prediction] distortion."""
    L = self.L
    coeff, recon_coeff, eob = (np.ascontiguousarray(a) for a in (coeff, recon_coeff, eob))
    n = coeff.shape[0]
    out = np.zeros((n, 2), dtype=np.uint64)
    d = [L.svtb200_dev_alloc(self.h, max(a.nbytes, 8)) for a in (coeff, recon_coeff, eob, out)]
    try:
        for p, a in zip(d[:3], (coeff, recon_coeff, eob)):
            self.check(L.svtb200_memcpy_h2d(self.h, p, a.ctypes.data, a.nbytes), "memcpy_h2d")
        self.check(L.svtb200_full_distortion_batch_dev(self.h, tx_size, n, d[0], d[1], d[2], d[3]), "full_distortion_batch_dev")
        self.check(L.svtb200_memcpy_d2h(self.h, out.ctypes.data, d[3], out.nbytes), "memcpy_d2h")
    finally:
        for p in d:
            L.svtb200_dev_free(self.h, p)
    return out


def _ctx_intra_predict(self, blocks):
    """blocks: ctypes array of IntraBlock with DEVICE pointers; asynchronous on the current compute queue."""
    self.check(self.L.svtb200_intra_predict_batch(self.h, C.addressof(blocks), len(blocks)), "intra_predict_batch")


class RateBatch(C.Structure):
    """SvtB200RateBatch (device pointers)"""
    _fields_ = [("tx_size", C.c_int32), ("n", C.c_int32), ("qcoeff", C.c_void_p), ("eob", C.c_void_p), ("blocks", C.c_void_p),
                ("token_costs", C.c_void_p), ("rate", C.c_void_p)]


class InterChain(C.Structure):
    """SvtB200InterChain: the inter chain of one picture for svtb200_inter_chain_dev"""
    _fields_ = [("after_queue", C.c_int32), ("on_queue", C.c_int32), ("me", C.c_void_p), ("blocks", C.c_void_p), ("n_blocks", C.c_int32), ("pic", PredPicture), ("desc", C.c_void_p),
                ("tx", C.POINTER(TxBatch)), ("n_tx", C.c_int32), ("qrows", C.c_void_p), ("n_qrows", C.c_int32), ("stream", C.c_void_p), ("capacity", C.c_uint32),
                ("offsets", C.c_void_p), ("rate", C.POINTER(RateBatch)), ("n_rate", C.c_int32), ("reads", C.c_uint64 * 3), ("n_reads", C.c_int32)]


RATE_BLOCK_DTYPE = np.dtype([("plane_type", "u1"), ("is_inter", "u1"), ("tx_type", "u1"), ("coeff_ctx", "u1")])


def _bind_rate(L):
    if getattr(L, "_rate_bound", False):
        return
    vp = C.c_void_p
    L.svtb200_set_scan_neighbors.argtypes = [vp, C.c_int, C.c_int, vp]
    L.svtb200_set_rate_tables.argtypes = [vp, vp, vp, vp, vp]
    L.svtb200_coeff_rate_batch.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.svtb200_coeff_rate_batch_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.svtb200_coeff_rate_multi_dev.argtypes = [vp, C.POINTER(RateBatch), C.c_int]
    L._rate_bound = True


def _ctx_register_rate_tables(self, tables):
    """tables: svt-vp9_b200/data/vp9_tables.npz (nb_<ts>_<tt>, rate_vtok / rate_vcost / rate_c6lo / rate_c6hi); the scans
    themselves come from register_scans."""
    _bind_rate(self.L)
    self._rate_keep = []
    for ts in range(4):
        for tt in range(4):
            nb = np.ascontiguousarray(tables["nb_%d_%d" % (ts, tt)], dtype=np.int16)
            self.check(self.L.svtb200_set_scan_neighbors(self.h, ts, tt, nb.ctypes.data), "set_scan_neighbors")
    a = [np.ascontiguousarray(tables[k], dtype=d) for k, d in (("rate_vtok", np.int16), ("rate_vcost", np.int32), ("rate_c6lo", np.int16), ("rate_c6hi", np.uint16))]
    self.check(self.L.svtb200_set_rate_tables(self.h, *(x.ctypes.data for x in a)), "set_rate_tables")


def _ctx_coeff_rate(self, tx_size, qcoeff, eob, blocks, token_costs):
    """qcoeff (n, N*N) int16, eob (n,) uint16, blocks (n,) RATE_BLOCK_DTYPE, token_costs (2, 2, 6, 2, 6, 12) uint32 of this
    transform size -> (n,) int32 rates."""
    _bind_rate(self.L)
    qcoeff = np.ascontiguousarray(qcoeff, dtype=np.int16)
    eob = np.ascontiguousarray(eob, dtype=np.uint16)
    blocks = np.ascontiguousarray(blocks, dtype=RATE_BLOCK_DTYPE)
    token_costs = np.ascontiguousarray(token_costs, dtype=np.uint32)
    assert token_costs.size == 2 * 2 * 6 * 2 * 6 * 12
    n = len(eob)
    rate = np.zeros(n, dtype=np.int32)
    self.check(self.L.svtb200_coeff_rate_batch(self.h, tx_size, n, qcoeff.ctypes.data, eob.ctypes.data, blocks.ctypes.data,
                                               token_costs.ctypes.data, rate.ctypes.data), "coeff_rate_batch")
    return rate


Context.register_rate_tables = _ctx_register_rate_tables
Context.coeff_rate = _ctx_coeff_rate
Context.inter_predict = _ctx_inter_predict
Context.intra_predict = _ctx_intra_predict
Context.full_distortion = _ctx_full_distortion
