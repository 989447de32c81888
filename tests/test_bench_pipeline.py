"""The end-to-end leg of bench.py drives the C ABI's four queues with several result slots; `--verify` recomputes its last
pictures synchronously on one queue and compares every downloaded byte (ME results, quantised coefficients, eob, filtered
planes).  35 steps so that the base-layer picture (deblocking + plane download) is among the verified ones."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_pipelined_end_to_end_leg_equals_synchronous_run():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "30", "--warmup", "5", "--verify"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    err = r.stderr.decode(errors="replace")
    assert r.returncode == 0, err[-2000:]
    assert "verify: ok" in err, err[-2000:]
    assert b'"e2e"' in r.stdout
