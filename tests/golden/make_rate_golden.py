"""Generates tests/golden/rate_golden.npz: rates of the REFERENCE's coeff_rate_estimate (through oracle/_ref's path driver,
oracle/ref_path_driver.c:refdrv_coeff_rate) for seeded random blocks of every transform size, with the token costs of the default
coefficient probabilities.  Run where /root/reference exists (oracle/build_ref.sh first): python tests/golden/make_rate_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
import test_rate as R  # noqa: E402


def main():
    rng = np.random.default_rng(20260924)
    tc = H.vp9_tables()["token_costs_default"]
    out = {}
    for ts in range(4):
        qs, es, bs = [], [], []
        for kind in R.KINDS:
            q, eob, blk = R._blocks(rng, ts, 60 if ts < 3 else 24, kind)
            qs.append(q); es.append(eob); bs.append(blk)
        q, eob, blk = np.concatenate(qs), np.concatenate(es), np.concatenate(bs)
        out["q%d" % ts], out["eob%d" % ts], out["blk%d" % ts] = q, eob, blk.view(np.uint8).reshape(-1, 4)
        out["rate%d" % ts] = R._ref_rate(ts, q, eob, blk, tc)
    dst = os.path.join(HERE, "rate_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
