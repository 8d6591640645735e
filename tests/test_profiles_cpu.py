"""The numbers bench.py quotes from profiles/ are reproducible from the committed raw data: `roofline.traffic` comes from
profiles/r02_launches_summary.json, which must be exactly what tools/summarize_launches.py makes of the committed ncu
launch list profiles/r02_launches.csv."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_launch_summary_matches_the_committed_launch_list():
    r = subprocess.run([sys.executable, str(ROOT / "tools/summarize_launches.py"), "profiles/r02_launches.csv"],
                       capture_output=True, text=True, cwd=str(ROOT), timeout=120)
    assert r.returncode == 0, r.stderr
    fresh = json.loads(r.stdout)
    committed = json.loads((ROOT / "profiles/r02_launches_summary.json").read_text())
    assert fresh == committed
    k = committed["kernels"]["conv_igemm_kernel<256, true>"]
    assert k["launches"] == 84 and 0.5 < k["share"] < 0.7


def test_bench_reads_traffic_from_that_summary():
    sys.path.insert(0, str(ROOT))
    import bench
    traffic, source = bench.launch_traffic("conv_igemm_kernel<256, true>")
    k = json.loads((ROOT / "profiles/r02_launches_summary.json").read_text())["kernels"]["conv_igemm_kernel<256, true>"]
    assert traffic == k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"]
    assert "r02_launches" in source
    assert bench.launch_traffic("no_such_kernel") == (None, None)
