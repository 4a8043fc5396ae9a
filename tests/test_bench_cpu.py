"""bench.py pieces that run without a GPU: the reference arm on the tiny workload, the byte accounting."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    hops = np.array([2, 3]); evals = np.array([11, 21]); w = {"k": 10}
    # per query: 4*(hops + evals-1) + rows*row_bytes + qbytes + k*8   
    total, mean = bench.algorithmic_bytes(hops, evals, evals, w, 384, 384)
    want = [4 * (2 + 10) + 11 * 384 + 384 + 80, 4 * (3 + 20) + 21 * 384 + 384 + 80]
    assert total == sum(want) and mean == sum(want) / 2
    total_f, _ = bench.algorithmic_bytes(hops, evals, np.array([5, 6]), w, 384, 384)
    assert total_f == sum(want) - (6 + 15) * 384
    assert bench.effective_cpus() >= 1


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line(reflib, tmp_path):
    """`bench.py --impl reference` (the reference's own CPU path via oracle/_ref) on the tiny workload."""
    env = dict(os.environ, SVSB200_CACHE=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "tiny-100kx96-f32-L2-w128", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "QPS" and d["unit"] == "queries/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "tiny-100kx96-f32-L2-w128" and d["higher_is_better"] is True
