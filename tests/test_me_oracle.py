"""CPU-side checks of the motion-estimation oracle and of the CUDA kernel's control logic.

  * the oracle port (oracle/me_port.c) must reproduce the reference encoder's own me_results on every golden
    picture under tests/golden (dumped from the unmodified reference by oracle/make_golden.py);
  * the CUDA kernel source, executed under the CPU emulation header (no GPU), must do the same on a sample of
    super-blocks -- this covers the host-visible logic of the kernel (search windows, tie-breaking, PU ordering).
"""
import glob
import os

import numpy as np
import pytest

import helpers as H

FIXTURES = sorted(glob.glob(os.path.join(H.GOLDEN, "me_*.bin.gz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 8


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-7] for p in FIXTURES])
def test_oracle_matches_reference_dump(path):
    pic = H.MePicture(path)
    out, _ = H.oracle_me_picture(pic)
    ok = H.me_equal(out, pic.expected, pic.params.num_lists)
    assert ok.all(), "oracle differs from the reference on %d PUs" % (~ok).sum()


def test_emulated_kernel_tables():
    assert H.emu().me_emu_check_tables() == 0


# (fixture, SBs, window-buffer bytes: 0 = the size the host code plans, 8400 = just above the full-pel window of the
#  16x9 search area so that the HME areas are searched in several stripes, 98304 = everything in one stripe; global interp)
EMU_CASES = [("me_192x128_m8t1_pic16", [0, 2, 5], 98304, 0), ("me_192x128_m8t1_pic16", [1, 4], 8400, 0),
             ("me_192x128_m8t1_pic17", [1, 4], 0, 0), ("me_192x128_m8t1_pic08", [3], 8400, 1),
             ("me_192x128_m5t0_pic01", [3], 0, 0), ("me_192x128_m3t1_pic18", [2], 0, 0), ("me_192x128_m3t1_pic02", [5], 0, 1)]


@pytest.mark.parametrize("name,sbs,win,gl", EMU_CASES, ids=["%s-w%d-g%d" % (c[0], c[2], c[3]) for c in EMU_CASES])
def test_emulated_kernel_matches_reference_dump(name, sbs, win, gl):
    pic = H.MePicture(os.path.join(H.GOLDEN, name + ".bin.gz"))
    for sb in sbs:
        out, _, _ = H.emu_me_picture(pic, sb, sb + 1, win_bytes=win, force_global_interp=gl)
        ok = H.me_equal(out[sb * 85:(sb + 1) * 85], pic.expected[sb * 85:(sb + 1) * 85], pic.params.num_lists)
        assert ok.all(), "emulated kernel differs on SB %d (%d PUs)" % (sb, (~ok).sum())
