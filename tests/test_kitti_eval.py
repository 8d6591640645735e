"""KITTI result writer / evaluator glue, host-only C-ABI entries mscnn_kitti_*.

Oracle: the reference's own tool examples/kitti_result/eval/evaluate_object.cpp compiled VERBATIM
(oracle/build_ref.py -> oracle/_ref/evaluate_object).  Where it is present the statistics files of both are
compared byte for byte on seeded synthetic label sets; the same comparison against files the reference tool
wrote here is committed under tests/golden/kitti_eval/ (tests/golden/make_kitti_golden.py) for boxes without
the reference.  The two MATLAB writer steps (dlmwrite / writeLabels record formats) are restated from the
scripts and are parity-unpinned beyond the format checks below."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from kitti_synth import make_dataset, rows_to_padded

ROOT = Path(__file__).resolve().parent.parent
REF_TOOL = ROOT / "oracle" / "_ref" / "evaluate_object"
GOLDEN = Path(__file__).resolve().parent / "golden" / "kitti_eval"


def _write_results(root, rows, n_images, comp="res"):
    from mscnn_b200 import kitti
    files = {}
    for cls, r in rows.items():
        dets, counts = rows_to_padded(r, n_images)
        files[cls] = root / f"{comp}_{cls.lower()}.txt"
        kitti.write_det_file(files[cls], dets, counts)
    res = root / comp
    kitti.write_labels(root / "val.txt", res / "data", car=files.get("Car"), ped=files.get("Pedestrian"),
                       cyc=files.get("Cyclist"))
    return res


def test_det_file_and_label_formats(tmp_path):
    from mscnn_b200 import kitti
    dets = np.zeros((2, 3, 5), np.float32)
    dets[0, 0] = [10.123456, 20.5, 30.25, 40.0, 0.987654]
    dets[1, 0] = [1234.5678, 0.0, 7.0, 8.0, 1e-5]
    dets[1, 1] = [5.0, 6.0, 7.0, 8.0, 0.5]
    f = tmp_path / "d.txt"
    kitti.write_det_file(f, dets, np.array([1, 2], np.int32))
    assert f.read_text().splitlines() == ["1,10.123,20.5,30.25,40,0.98765", "2,1234.6,0,7,8,1e-05", "2,5,6,7,8,0.5"]
    (tmp_path / "val.txt").write_text("000007\n000123\n")
    kitti.write_labels(tmp_path / "val.txt", tmp_path / "r" / "data", car=f)
    a = (tmp_path / "r" / "data" / "000007.txt").read_text()
    assert a == "Car -1 -1 -10 10.12 20.50 40.37 60.50 -1 -1 -1 -1000 -1000 -1000 -10 987.65 \n"
    b = (tmp_path / "r" / "data" / "000123.txt").read_text().splitlines()
    assert len(b) == 2 and b[1].startswith("Car -1 -1 -10 5.00 6.00 12.00 14.00 ")


@pytest.mark.parametrize("seed,n", [(7, 24), (11, 40), (3, 5)])
@pytest.mark.skipif(not REF_TOOL.exists(), reason="oracle/_ref/evaluate_object not built (needs /root/reference)")
def test_evaluate_byte_identical_to_reference_tool(tmp_path, seed, n):
    from mscnn_b200 import kitti
    ids, rows = make_dataset(tmp_path, n_images=n, seed=seed)
    ours = _write_results(tmp_path, rows, n, "ours")
    theirs = tmp_path / "theirs"
    shutil.copytree(ours / "data", theirs / "data")
    ap = kitti.evaluate(tmp_path / "label_2", ours, tmp_path / "val.txt")
    r = subprocess.run([str(REF_TOOL), str(tmp_path / "label_2"), str(theirs), str(tmp_path / "val.txt")],
                       capture_output=True, text=True, timeout=300)
    assert "done" in r.stdout, r.stdout + r.stderr
    for cls in ("car", "pedestrian", "cyclist"):
        for rel in (f"stats_{cls}_detection.txt", f"plot/{cls}_detection.txt"):
            assert (ours / rel).read_bytes() == (theirs / rel).read_bytes(), rel
        assert not (ours / f"stats_{cls}_orientation.txt").exists()
        assert not (theirs / f"stats_{cls}_orientation.txt").exists()
        tab = np.loadtxt(ours / f"plot/{cls}_detection.txt")
        assert np.allclose(ap[cls], 100 * tab[0:41:4, 1:4].mean(0), atol=1e-4)


@pytest.mark.skipif(not REF_TOOL.exists(), reason="oracle/_ref/evaluate_object not built (needs /root/reference)")
def test_evaluate_with_orientation_and_missing_classes(tmp_path):
    """Result files that carry a valid alpha switch the orientation statistics on; classes never detected are
    not evaluated (evaluate_object.cpp:124-134)."""
    from mscnn_b200 import kitti
    ids, rows = make_dataset(tmp_path, n_images=16, seed=21)
    rng = np.random.default_rng(0)
    for comp in ("ours", "theirs"):
        (tmp_path / comp / "data").mkdir(parents=True)
    for pos, img in enumerate(ids, start=1):
        lines = []
        for x in rows["Car"][rows["Car"][:, 0] == pos]:
            lines.append(f"Car -1 -1 {rng.uniform(-3, 3):.2f} {x[1]:.2f} {x[2]:.2f} {x[1] + x[3]:.2f} {x[2] + x[4]:.2f} "
                         f"-1 -1 -1 -1000 -1000 -1000 -10 {x[5] * 1000:.2f} \n")
        for comp in ("ours", "theirs"):
            (tmp_path / comp / "data" / f"{img:06d}.txt").write_text("".join(lines))
    ap = kitti.evaluate(tmp_path / "label_2", tmp_path / "ours", tmp_path / "val.txt")
    subprocess.run([str(REF_TOOL), str(tmp_path / "label_2"), str(tmp_path / "theirs"), str(tmp_path / "val.txt")],
                   capture_output=True, text=True, timeout=300, check=True)
    for rel in ("stats_car_detection.txt", "stats_car_orientation.txt", "plot/car_detection.txt", "plot/car_orientation.txt"):
        assert (tmp_path / "ours" / rel).read_bytes() == (tmp_path / "theirs" / rel).read_bytes(), rel
    assert ap["pedestrian"] is None and ap["cyclist"] is None and ap["car"] is not None
    assert not (tmp_path / "ours" / "stats_pedestrian_detection.txt").exists()
    assert not (tmp_path / "theirs" / "stats_pedestrian_detection.txt").exists()


def test_evaluate_against_committed_reference_output(tmp_path):
    """Same comparison without the reference tool: its output for seed 7 / 24 images is committed."""
    from mscnn_b200 import kitti
    ids, rows = make_dataset(tmp_path, n_images=24, seed=7)
    ours = _write_results(tmp_path, rows, 24, "ours")
    kitti.evaluate(tmp_path / "label_2", ours, tmp_path / "val.txt")
    for f in sorted(GOLDEN.glob("*.txt")):
        rel = f.name.replace("plot_", "plot/") if f.name.startswith("plot_") else f.name
        assert (ours / rel).read_bytes() == f.read_bytes(), rel
    assert len(list(GOLDEN.glob("*.txt"))) == 6


def test_evaluate_perfect_and_empty(tmp_path):
    """Known answers: 80 ground-truth cars (>= 2 per recall step, so all 41 recall points are reached), every one
    detected exactly -> precision 1 everywhere, AP 100; a missing result file is an error, not a crash."""
    from mscnn_b200 import capi, kitti
    gt = tmp_path / "label_2"
    gt.mkdir()
    (tmp_path / "val.txt").write_text("".join(f"{i:06d}\n" for i in range(4)))
    rows = []
    for img in range(4):
        lines = []
        for k in range(20):
            x, y = 10 + 120 * (k % 10), 20 + 150 * (k // 10)
            lines.append(f"Car 0.00 0 1.00 {x:.2f} {y:.2f} {x + 100:.2f} {y + 80:.2f} 1.5 1.6 3.9 1 1.5 20 1.0\n")
            rows.append([img + 1, x, y, 100, 80, 0.5 + 0.005 * (20 * img + k)])
        (gt / f"{img:06d}.txt").write_text("".join(lines))
    dets, counts = rows_to_padded(np.array(rows, dtype=np.float64), 4)
    kitti.write_det_file(tmp_path / "car.txt", dets, counts)
    kitti.write_labels(tmp_path / "val.txt", tmp_path / "res" / "data", car=tmp_path / "car.txt")
    ap = kitti.evaluate(gt, tmp_path / "res", tmp_path / "val.txt")
    assert ap["car"] == (100.0, 100.0, 100.0) and ap["pedestrian"] is None
    with pytest.raises(capi.MscnnError):
        kitti.evaluate(gt, tmp_path / "nowhere", tmp_path / "val.txt")
