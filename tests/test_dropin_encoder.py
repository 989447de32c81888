"""Encoder-level drop-in check: the UNMODIFIED reference encoder objects with motion_estimate_sb() and
eb_vp9_loop_filter_frame() bound to libsvtvp9_b200.so (svt-vp9_b200/shim/svtb200_wrap.c, linked with -Wl,--wrap by
oracle/build_ref.sh into oracle/_ref/SvtVp9EncApp_b200) must write the same bitstream, byte for byte, as the reference
application built from the same objects (oracle/_ref/SvtVp9EncApp_ref).  Both binaries are built where /root/reference
exists and travel to the GPU box; the test is skipped where they are absent."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

import helpers as H

REF_DIR = os.path.join(H.ROOT, "oracle", "_ref")
APP_REF, APP_B200 = os.path.join(REF_DIR, "SvtVp9EncApp_ref"), os.path.join(REF_DIR, "SvtVp9EncApp_b200")
pytestmark = pytest.mark.skipif(not (os.path.exists(APP_REF) and os.path.exists(APP_B200)), reason="oracle/_ref applications not built")


def _write_yuv(path, w, h, n, seed):
    clip = H.synth.make_clip(w, h, n, seed=seed, k=3, g=4, a=5, b=3, m=32, noise=2)
    with open(path, "wb") as f:
        for y, u, v in clip:
            f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())


def _encode(app, yuv, out, w, h, n, mode, extra_env=None, extra_args=()):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(p for p in (os.path.dirname(H.cudart_path()), env.get("LD_LIBRARY_PATH", "")) if p)
    env.update(extra_env or {})
    cmd = [app, "-i", yuv, "-w", str(w), "-h", str(h), "-n", str(n), "-enc-mode", str(mode), "-tune", "1", "-q", "45", "-b", out] + list(extra_args)
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    assert os.path.exists(out) and os.path.getsize(out) > 0, r.stdout.decode(errors="replace")[-2000:]
    return hashlib.md5(open(out, "rb").read()).hexdigest()


def _report(path):
    return dict((l.split()[0], int(l.split()[1])) for l in open(path) if len(l.split()) == 2)


def test_wrapper_falls_back_to_reference_kernels_without_a_gpu():
    """SVT_VP9_GPU=off: every call goes to __real_*; the bitstream must equal the reference application's."""
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        _write_yuv(yuv, 192, 128, 10, 7)
        a = _encode(APP_REF, yuv, os.path.join(d, "ref.ivf"), 192, 128, 10, 8)
        b = _encode(APP_B200, yuv, os.path.join(d, "b200.ivf"), 192, 128, 10, 8, {"SVT_VP9_GPU": "off", "SVT_VP9_GPU_REPORT": os.path.join(d, "rep.txt")})
        assert a == b
        assert _report(os.path.join(d, "rep.txt"))["me_pictures"] == 0


def test_recorded_motion_estimation_table_replays_to_the_same_bitstream():
    """SVT_VP9_ME_RECORD / SVT_VP9_ME_REPLAY (the "zero-cost motion estimation" ceiling measurement of bench.py): a table recorded
    from the CPU path and replayed yields the reference bitstream, with every ME super-block served from the table."""
    with tempfile.TemporaryDirectory() as d:
        yuv, tab = os.path.join(d, "clip.yuv"), os.path.join(d, "tab.bin")
        _write_yuv(yuv, 320, 192, 20, 5)
        a = _encode(APP_REF, yuv, os.path.join(d, "ref.ivf"), 320, 192, 20, 8)
        b = _encode(APP_B200, yuv, os.path.join(d, "rec.ivf"), 320, 192, 20, 8, {"SVT_VP9_GPU": "off", "SVT_VP9_ME_RECORD": tab})
        c = _encode(APP_B200, yuv, os.path.join(d, "rep.ivf"), 320, 192, 20, 8, {"SVT_VP9_GPU": "off", "SVT_VP9_ME_REPLAY": tab, "SVT_VP9_GPU_REPORT": os.path.join(d, "rep.txt")})
        assert a == b == c
        assert _report(os.path.join(d, "rep.txt"))["replayed_sbs"] == 19 * 15


def test_dropin_library_has_the_shape_of_the_reference_library():
    """SONAME libSvtVp9Enc.so.1 and the reference's symbol map (Source/API/EbApi.version: Eb*, eb* only)."""
    lib = os.path.join(REF_DIR, "dropin", "libSvtVp9Enc.so.1")
    if not os.path.exists(lib):
        pytest.skip("drop-in library not built")
    dyn = subprocess.run(["readelf", "-d", lib], stdout=subprocess.PIPE).stdout.decode()
    assert "libSvtVp9Enc.so.1" in [l.split("[")[-1].rstrip("]") for l in dyn.splitlines() if "SONAME" in l]
    syms = [l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", lib], stdout=subprocess.PIPE).stdout.decode().splitlines() if l.strip()]
    names = [s.split("@")[0] for s in syms]
    assert all(n.startswith(("eb", "Eb")) or n == "SVT_VP9_0" for n in names), [n for n in names if not n.startswith(("eb", "Eb"))][:5]
    for api_fn in ("eb_vp9_svt_init_handle", "eb_vp9_svt_enc_set_parameter", "eb_vp9_init_encoder", "eb_vp9_svt_enc_stream_header", "eb_vp9_svt_enc_eos_nal",
                   "eb_vp9_svt_enc_send_picture", "eb_vp9_svt_get_packet", "eb_vp9_svt_release_out_buffer", "eb_vp9_svt_get_recon", "eb_vp9_deinit_encoder",
                   "eb_vp9_deinit_handle"):
        assert api_fn in names, api_fn


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,mode", [(640, 384, 34, 8), (448, 256, 20, 3), (1920, 1088, 18, 8), (1920, 1080, 18, 8), (640, 480, 34, 9), (3840, 2160, 34, 8)],
                         ids=["640x384-m8", "448x256-m3", "1920x1088-m8", "1920x1080-m8", "640x480-m9", "3840x2160-m8"])
def test_encoder_with_gpu_hot_path_writes_the_reference_bitstream(w, h, n, mode):
    """(at 3840x2160 the reference aborts in its own teardown after the last packet -- both builds do; the bitstream is complete)"""
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        _write_yuv(yuv, w, h, n, 11 + mode)
        a = _encode(APP_REF, yuv, os.path.join(d, "ref.ivf"), w, h, n, mode)
        rep = os.path.join(d, "rep.txt")
        b = _encode(APP_B200, yuv, os.path.join(d, "b200.ivf"), w, h, n, mode, {"SVT_VP9_GPU": "0", "SVT_VP9_GPU_REPORT": rep, "SVT_VP9_GPU_VERBOSE": "1"})
        r = _report(rep)
        assert r["state"] == 1 and r["me_pictures"] >= n - 2 and r["fallbacks"] == 0 and r["launches"] > 0, r   # the GPU did the work
        assert a == b, "bitstreams differ (%s)" % r


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,mode", [(320, 192, 10, 8), (256, 128, 6, 3)], ids=["320x192-m8", "256x128-m3"])
def test_encoder_with_every_transform_call_on_the_gpu_writes_the_reference_bitstream(w, h, n, mode):
    """SVT_VP9_GPU_ENCODE_PASS=1: seam S3 bound call by call (svt-vp9_b200/shim/svtb200_wrap_tx.c) -- every forward transform, quantiser
    and inverse transform + reconstruction call of mode decision and of the encode pass (perform_coding_loop / perform_inv_trans_add,
    Source/Lib/Codec/EbEncDecProcess.c:365-655) is one device batch of one block: all sizes, the ADST types of intra blocks (enc-mode 3), the
    partial 32x32 transform, every eob the encoder produces.  Slow by construction (a PCIe round trip per call); the bitstream must not change."""
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        _write_yuv(yuv, w, h, n, 23 + mode)
        a = _encode(APP_REF, yuv, os.path.join(d, "ref.ivf"), w, h, n, mode)
        rep = os.path.join(d, "rep.txt")
        b = _encode(APP_B200, yuv, os.path.join(d, "b200.ivf"), w, h, n, mode,
                    {"SVT_VP9_GPU": "0", "SVT_VP9_GPU_ENCODE_PASS": "1", "SVT_VP9_GPU_REPORT": rep, "SVT_VP9_GPU_VERBOSE": "1"})
        r = _report(rep)
        assert r["state"] == 1 and r["fallbacks"] == 0 and r["tx_fallbacks"] == 0, r
        assert r["tx_forward"] > 1000 and r["tx_quantize"] > 1000 and r["tx_inverse"] > 100, r   # the device served the transform calls
        assert a == b, "bitstreams differ (%s)" % r


def test_encode_pass_binding_without_a_gpu_leaves_the_reference_kernels_in_place():
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        _write_yuv(yuv, 192, 128, 6, 9)
        a = _encode(APP_REF, yuv, os.path.join(d, "ref.ivf"), 192, 128, 6, 8)
        rep = os.path.join(d, "rep.txt")
        b = _encode(APP_B200, yuv, os.path.join(d, "b200.ivf"), 192, 128, 6, 8, {"SVT_VP9_GPU": "off", "SVT_VP9_GPU_ENCODE_PASS": "1", "SVT_VP9_GPU_REPORT": rep})
        assert a == b
        r = _report(rep)
        assert r["tx_forward"] == 0 and r["tx_quantize"] == 0 and r["tx_inverse"] == 0


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,mode", [(1920, 1080, 34, 8), (3840, 2160, 34, 8)], ids=["1920x1080-m8", "3840x2160-m8"])
def test_one_encoder_on_several_gpus_writes_the_reference_bitstream(w, h, n, mode):
    """SVT_VP9_GPU=0,1[,2,3]: ONE encoder handle, its pictures dealt to the devices mini-GOP by mini-GOP (each device gets the source
    planes it needs; no exchange between devices).  Same bitstream as the reference; every device took its share of the pictures."""
    g = _n_gpus()
    if g < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    g = min(g, 4)
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        _write_yuv(yuv, w, h, n, 3 + mode)
        a = _encode(APP_REF, yuv, os.path.join(d, "ref.ivf"), w, h, n, mode)
        rep = os.path.join(d, "rep.txt")
        b = _encode(APP_B200, yuv, os.path.join(d, "b200.ivf"), w, h, n, mode,
                    {"SVT_VP9_GPU": ",".join(str(i) for i in range(g)), "SVT_VP9_GPU_REPORT": rep, "SVT_VP9_GPU_VERBOSE": "1"})
        r = _report(rep)
        assert r["state"] == 1 and r["devices"] == g and r["fallbacks"] == 0, r
        per_dev = [r["me_pictures_dev%d" % i] for i in range(g)]
        assert sum(per_dev) == r["me_pictures"] and sum(1 for v in per_dev if v > 0) >= 2, r   # mini-GOPs dealt to the devices
        assert a == b, "bitstreams differ (%s)" % r
