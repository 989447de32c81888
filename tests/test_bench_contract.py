"""bench.py --impl reference (the CPU arm: runs without a GPU): the JSON line carries the contract's keys, and under a
2-process launch only rank 0 prints."""
import json
import os
import subprocess
import sys

import pytest

import helpers as H

ROOT = H.ROOT
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "cpu_baseline", "e2e"}


def _lines(out):
    return [json.loads(l) for l in out.decode().splitlines() if l.startswith("{")]


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref not built")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-400:]
    lines = _lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "fps" and d["higher_is_better"] is True and d["dtype"] == "u8" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "fps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0 * d["config"]["pictures_per_step"]) < 1.0


@pytest.mark.skipif(not H.have_refdrv(), reason="oracle/_ref not built")
def test_reference_arm_two_ranks_print_once():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-400:]
    lines = _lines(r.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
