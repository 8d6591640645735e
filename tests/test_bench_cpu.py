"""bench.py's host-side logic and its reference arm (the one place besides the tests where oracle/ is executed), without a
GPU: the layer-range steps of `--impl reference` add up to whole forwards, the arm prints ONE JSON line with the keys the
driver reads, and the BLAS thread count survives torchrun's OMP_NUM_THREADS=1."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_steps_partition_whole_forwards():
    sys.path.insert(0, str(ROOT))
    import bench
    for layers, k in [(62, 20), (62, 1), (62, 100), (62, 62), (62, 125), (45, 7)]:
        passes = bench._chunks(layers, k)
        assert sum(len(p) for p in passes) == k
        for p in passes:
            assert p[0][0] == 0 and p[-1][1] == layers - 1
            assert all(a[1] >= a[0] for a in p)
            assert all(b[0] == a[1] + 1 for a, b in zip(p, p[1:]))


def test_reference_arm_prints_the_contract_line():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built here")
    env = dict(os.environ, MSCNN_BENCH_TEST_HW="96x320", OMP_NUM_THREADS="1", RANK="0")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "7",
                        "--warmup", "1"], capture_output=True, text=True, cwd=str(ROOT), env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["steps"] == 7 and d["warmup"] == 1 and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - 1.0 / d["seconds_per_image"]) < 1e-9
    assert abs(d["ms_per_step"] * 7 / 1e3 - d["seconds_per_image"]) < 1e-6          # the 7 steps ARE one forward
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"]
    assert cb["cores"] == (os.cpu_count() or 1) or cb["cores"] > 1                   # not torchrun's single thread
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["proposals_per_image"] >= 1 and "TEST OVERRIDE" in d["config"]["workload"]
    # the other ranks of a torchrun launch exit quietly
    r2 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, cwd=str(ROOT), env=dict(env, RANK="1"), timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == ""
