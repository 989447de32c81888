"""Edge-mask construction (eb_vp9_build_mask_frame -> eb_vp9_setup_mask): random valid partition trees.
oracle restatement vs the reference's own function; kernel source under the CPU emulation vs the oracle; CUDA vs both;
and masks built on the device feeding the deblocking kernel."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

api = H.api
W8 = [1, 1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8]
H8 = [1, 1, 1, 1, 2, 1, 2, 4, 2, 4, 8, 4, 8]
MAX_TX = [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3]   # largest luma transform of a block size
REFDRV_MI = np.dtype([("sb_type", "u1"), ("tx_size", "u1"), ("skip", "u1"), ("ref", "u1"), ("cls", "u1")])


def random_grid(rng, mi_rows, mi_cols, levels, p_split=0.55):
    """A valid partition of every 64x64 (blocks may hang over the picture edge, as in the encoder); every unit of a block
    carries the block's values."""
    rows8, cols8 = (mi_rows + 7) // 8 * 8, (mi_cols + 7) // 8 * 8
    g = np.zeros((rows8, cols8), dtype=REFDRV_MI)

    def put(r, c, bs):
        tx = int(rng.integers(0, MAX_TX[bs] + 1))
        ref = int(rng.integers(0, 4))
        v = (bs, tx, int(rng.integers(0, 2)), ref, int(rng.integers(0, 2)) if ref else 0)
        g[r:r + H8[bs], c:c + W8[bs]] = v

    def part(r, c, n):  # n = size in units: 8, 4, 2
        sq, hz, vt = {8: (12, 11, 10), 4: (9, 8, 7), 2: (6, 5, 4)}[n]
        k = rng.random()
        if k > p_split:
            t = int(rng.integers(0, 3))
            if t == 0:
                put(r, c, sq)
            elif t == 1:
                put(r, c, hz); put(r + n // 2, c, hz)
            else:
                put(r, c, vt); put(r, c + n // 2, vt)
        elif n > 2:
            for a in range(4):
                part(r + (a >> 1) * (n // 2), c + (a & 1) * (n // 2), n // 2)
        else:
            for a in range(4):
                put(r + (a >> 1), c + (a & 1), int(rng.integers(0, 4)))
    for r in range(0, rows8, 8):
        for c in range(0, cols8, 8):
            part(r, c, 8)
    g = np.ascontiguousarray(g[:mi_rows, :mi_cols])
    mi = np.zeros(g.shape, dtype=api.LF_MODE_INFO_DTYPE)
    mi["sb_type"], mi["tx_size"] = g["sb_type"], g["tx_size"]
    mi["skip_inter"] = (g["skip"] != 0) & (g["ref"] > 0)
    mi["filter_level"] = np.asarray(levels, dtype=np.uint8)[g["ref"] * 2 + g["cls"]]
    return g, mi


def oracle_masks(mi):
    rows, cols = mi.shape
    sbs_w, sbs_h = (cols + 7) // 8, (rows + 7) // 8
    out = np.zeros(sbs_w * sbs_h, dtype=api.LF_MASK_DTYPE)
    L = H.oracle()
    L.svtb200_oracle_build_masks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.svtb200_oracle_build_masks(rows, cols, mi.ctypes.data, cols, out.ctypes.data, sbs_w)
    return out


CASES = [(34, 60, 1), (17, 25, 2), (8, 8, 3), (135, 240, 4), (9, 15, 5)]
LEVELS = [[12, 0, 20, 33, 5, 63, 0, 41], [30] * 8]


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
@pytest.mark.parametrize("rows,cols,seed", CASES)
def test_oracle_masks_match_reference_function(rows, cols, seed):
    rng = np.random.default_rng(seed)
    for levels in LEVELS:
        g, mi = random_grid(rng, rows, cols, levels)
        sbs_w, sbs_h = (cols + 7) // 8, (rows + 7) // 8
        ref = np.zeros(sbs_w * sbs_h, dtype=api.LF_MASK_DTYPE)
        L = H.refdrv()
        L.refdrv_build_masks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lv = np.asarray(levels, dtype=np.uint8)
        L.refdrv_build_masks(rows, cols, g.ctypes.data, cols, lv.ctypes.data, ref.ctypes.data, sbs_w)
        got = oracle_masks(mi)
        assert ref.tobytes() == got.tobytes()
        assert got["left_y"].any() and got["int_4x4_uv"].any()


@pytest.mark.parametrize("rows,cols,seed", CASES[:3])
def test_emulated_kernel_masks_match_oracle(rows, cols, seed):
    rng = np.random.default_rng(100 + seed)
    g, mi = random_grid(rng, rows, cols, LEVELS[0])
    sbs_w, sbs_h = (cols + 7) // 8, (rows + 7) // 8
    out = np.zeros(sbs_w * sbs_h, dtype=api.LF_MASK_DTYPE)
    E = H.emu()
    E.lfmask_emu_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    E.lfmask_emu_run(rows, cols, mi.ctypes.data, cols, out.ctypes.data, sbs_w)
    assert out.tobytes() == oracle_masks(mi).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,seed", CASES + [(270, 480, 9)])
def test_cuda_masks_match_oracle(rows, cols, seed):
    rng = np.random.default_rng(200 + seed)
    g, mi = random_grid(rng, rows, cols, LEVELS[0])
    ctx = api.Context(0)
    got = ctx.build_masks(mi)
    ctx.close()
    assert got.tobytes() == oracle_masks(mi).tobytes()


@pytest.mark.gpu
def test_device_built_masks_drive_the_deblocking_kernel():
    """mode info -> masks (device) -> deblocking (device) equals oracle masks -> oracle deblocking"""
    rng = np.random.default_rng(77)
    w, h = 456, 264
    g, mi = random_grid(rng, h // 8, w // 8, [24, 0, 30, 18, 9, 40, 0, 27])
    y = rng.integers(96, 132, size=(h, w), dtype=np.uint8)
    u = rng.integers(100, 128, size=(h // 2, w // 2), dtype=np.uint8)
    v = rng.integers(100, 128, size=(h // 2, w // 2), dtype=np.uint8)
    P = api.LfParams()
    P.width, P.height, P.mi_rows, P.mi_cols, P.y_only, P.filter_level = w, h, h // 8, w // 8, 0, 24
    thr = H.lf_thresholds()
    sbs_w = (w // 8 + 7) // 8
    ctx = api.Context(0)
    lfm = ctx.build_masks(mi)
    got = ctx.deblock(P, y, u, v, lfm, sbs_w, thr)
    ctx.close()
    want = H.oracle_deblock(P, y, u, v, oracle_masks(mi), sbs_w, thr)
    for a, b in zip(got, want):
        assert (a == b).all()
    assert (want[0] != y).any()


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref path driver not built")
def test_pick_filter_level_matches_reference_for_every_qindex():
    """row d1: svtb200_pick_filter_level vs eb_vp9_pick_filter_level(LPF_PICK_FROM_Q); loads the product library without a GPU"""
    P = C.CDLL(H.os.path.join(H.ROOT, "svt-vp9_b200", "libsvtvp9_b200.so"))
    R = H.refdrv()
    tab = H.vp9_tables()["ac_q"]
    for q in range(256):
        ac = R.refdrv_ac_quant(q)
        assert ac == int(tab[q])
        for key in (0, 1):
            assert P.svtb200_pick_filter_level(ac, key) * 256 == R.refdrv_pick_filter_level(q, key), (q, key)
