"""KITTI result writer / evaluator glue over the host-only C-ABI entries
mscnn_kitti_* (mscnn_b200/csrc/kitti_eval.cpp): the counterpart of
examples/kitti_car/run_mscnn_detection.m:150-161 (detection list file),
examples/kitti_result/writeDetForEval.m (per-image label files) and
examples/kitti_result/eval/evaluate_object.cpp (the benchmark's precision/recall evaluation)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import capi


def _b(p) -> bytes | None:
    return None if p is None else str(p).encode()


def write_det_file(path, dets: np.ndarray, counts: np.ndarray, first_image_index: int = 1, append: bool = False) -> None:
    """dets [N][max_rois][5] = [x y w h prob], counts [N]: host copies of Net.detect's outputs."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    assert dets.ndim == 3 and dets.shape[2] == 5 and counts.shape == (dets.shape[0],)
    capi.check(capi.lib().mscnn_kitti_write_det_file(_b(path), dets.shape[0], dets.ctypes.data, counts.ctypes.data,
                                                     dets.shape[1], first_image_index, int(append)), "kitti_write_det_file")


def write_labels(list_path, save_dir, car=None, ped=None, cyc=None, score_scale: float = 1000.0) -> None:
    Path(save_dir).parent.mkdir(parents=True, exist_ok=True)
    capi.check(capi.lib().mscnn_kitti_write_labels(_b(car), _b(ped), _b(cyc), _b(list_path), _b(save_dir), score_scale),
               "kitti_write_labels")


def evaluate(gt_dir, result_dir, list_path) -> dict[str, tuple[float, float, float] | None]:
    """Writes stats_*_detection.txt and plot/*_detection.txt under result_dir; returns the 11-point AP (percent)
    per class as (easy, moderate, hard), None for a class without detections."""
    ap = np.zeros(9, dtype=np.float64)
    capi.check(capi.lib().mscnn_kitti_evaluate(_b(gt_dir), _b(result_dir), _b(list_path), ap.ctypes.data), "kitti_evaluate")
    out = {}
    for i, name in enumerate(("car", "pedestrian", "cyclist")):
        v = ap[3 * i: 3 * i + 3]
        out[name] = None if v[0] == -1 else tuple(float(x) for x in v)
    return out
