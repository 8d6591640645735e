"""SURVEY.md 8(a) row a15 (the MATLAB post-process) against HAND-DERIVED known answers
(tests/golden/a15_postprocess_cases.json, written by tests/golden/make_a15_vectors.py with the derivation of every number
from /root/reference/examples/kitti_car/run_mscnn_detection.m:75-120 and /root/reference/utils/bbNms.m:75-126): greedy
'maxg' order, overlap exactly 0.5 (strict >), score ties (stable sort), boxes of negative width produced by the border
clipping, zero-extent / low-score proposals dropped at :82, the '>=' at the threshold, delta decode, softmax, ratios; and
the cascade driver's variant (/root/reference/examples/kitti_car/run_cascademscnn.m:96-123: rescale, clip, the "+ 1"
extent convention, proposals of zero extent dropped at :113).
CPU: the restatement (oracle/port.py, oracle/mscnn_oracle.c).  GPU: the device kernels through the C ABI."""
import json
from pathlib import Path

import numpy as np
import pytest

CASES = json.loads((Path(__file__).resolve().parent / "golden" / "a15_postprocess_cases.json").read_text())["cases"]


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "bbnms"], ids=lambda c: c["name"])
def test_bbnms_hand_vectors_oracle(case):
    from oracle import port
    keep = port.bbnms_maxg(np.array(case["bbs"], dtype=np.float64), case["overlap"])
    assert keep.tolist() == case["keep"], case["derivation"]


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "postprocess"], ids=lambda c: c["name"])
def test_postprocess_hand_vectors_oracle(case):
    from oracle import port
    got = port.detect_postprocess(np.array(case["proposals_score"], np.float32), np.array(case["cls_pred"], np.float32),
                                  np.array(case["bbox_pred"], np.float32), cls_id=case["cls_id"],
                                  ratios=tuple(case["ratios"]), net_hw=tuple(case["net_hw"]),
                                  org_hw=tuple(case["org_hw"]) if "org_hw" in case else None)
    want = np.array(case["dets"], dtype=np.float64)
    assert got.shape == want.shape, (got, case["derivation"])
    np.testing.assert_allclose(got, want, rtol=case["tol"], atol=case["tol"], err_msg=case["derivation"])


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "cascade"], ids=lambda c: c["name"])
def test_cascade_postprocess_hand_vectors_oracle(case):
    from oracle import port
    got = port.cascade_detect_postprocess(np.array(case["proposals"], np.float32), np.array(case["cls_prob"], np.float32),
                                          np.array(case["output_bbox"], np.float32), cls_id=case["cls_id"],
                                          ratios=tuple(case["ratios"]), org_hw=tuple(case["org_hw"]),
                                          net_hw=tuple(case["net_hw"]))
    want = np.array(case["dets"], dtype=np.float64)
    assert got.shape == want.shape, (got, case["derivation"])
    np.testing.assert_allclose(got, want, rtol=case["tol"], atol=case["tol"], err_msg=case["derivation"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "cascade"], ids=lambda c: c["name"])
def test_cascade_postprocess_hand_vectors_device(cuda, case):
    import torch
    from mscnn_b200 import capi, ops
    prop = np.array(case["proposals"], np.float32)
    cfg = capi.DetectCfg()
    cfg.num_cls, cfg.cls_id = 2, case["cls_id"]
    cfg.nms_overlap = 0.5
    cfg.ratio_h, cfg.ratio_w = case["ratios"]
    cfg.org_h, cfg.org_w = case["org_hw"]
    cfg.max_rois_per_image = 64
    n = len(prop)
    num_out = torch.tensor([n, n, n], dtype=torch.int32, device=cuda)
    dets, cnt = ops.cascade_detect_postprocess(cfg, 1, torch.from_numpy(prop).to(cuda),
                                               torch.from_numpy(np.array(case["cls_prob"], np.float32)).to(cuda),
                                               torch.from_numpy(np.array(case["output_bbox"], np.float32)).to(cuda), num_out)
    torch.cuda.synchronize()
    got = dets[0, :int(cnt[0].item())].cpu().numpy()
    want = np.array(case["dets"], dtype=np.float64)
    assert got.shape == want.shape, (got, case["derivation"])
    np.testing.assert_allclose(got, want, rtol=case["tol"], atol=case["tol"], err_msg=case["derivation"])


def _device_detect(cuda, prop, cls, bbox, cls_id, net_hw, ratios, org_hw, overlap=0.5, thr=-10.0):
    import torch
    from mscnn_b200 import capi, ops
    cfg = capi.DetectCfg()
    cfg.num_cls, cfg.cls_id = cls.shape[1], cls_id
    for k, v in enumerate([0.1, 0.1, 0.2, 0.2]):
        cfg.bbox_std[k], cfg.bbox_mean[k] = v, 0.0
    cfg.proposal_thr, cfg.nms_overlap = thr, overlap
    cfg.ratio_h, cfg.ratio_w = ratios
    cfg.org_h, cfg.org_w = org_hw
    cfg.max_rois_per_image = 64
    n = len(prop)
    num_out = torch.tensor([max(n, 1), n, n], dtype=torch.int32, device=cuda)
    dets, cnt = ops.detect_postprocess(cfg, 1, torch.from_numpy(prop).to(cuda), torch.from_numpy(cls).to(cuda),
                                       torch.from_numpy(bbox).to(cuda), num_out)
    torch.cuda.synchronize()
    k = int(cnt[0].item())
    return dets[0, :k].cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "postprocess"], ids=lambda c: c["name"])
def test_postprocess_hand_vectors_device(cuda, case):
    org_hw = tuple(case.get("org_hw", case["net_hw"]))
    got = _device_detect(cuda, np.array(case["proposals_score"], np.float32), np.array(case["cls_pred"], np.float32),
                         np.array(case["bbox_pred"], np.float32), case["cls_id"], tuple(case["net_hw"]),
                         tuple(case["ratios"]), org_hw)
    want = np.array(case["dets"], dtype=np.float64)
    assert got.shape == want.shape, (got, case["derivation"])
    np.testing.assert_allclose(got, want, rtol=case["tol"], atol=case["tol"], err_msg=case["derivation"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "bbnms"], ids=lambda c: c["name"])
def test_bbnms_hand_vectors_device(cuda, case):
    """The NMS cases through the device post-process: each box is fed as a proposal with zero deltas, unit ratios and a
    class logit chosen so that softmax gives the case's score (two classes: p = 1 / (1 + exp(-z)) -> z = logit(p)), on an
    image large enough that the clipping is inactive (negative-width rows come in as x2 < x1 proposals)."""
    bbs = np.array(case["bbs"], dtype=np.float64)
    prop = np.zeros((len(bbs), 6), np.float32)
    prop[:, 1], prop[:, 2] = bbs[:, 0], bbs[:, 1]
    prop[:, 3], prop[:, 4] = bbs[:, 0] + bbs[:, 2], bbs[:, 1] + bbs[:, 3]
    z = np.log(bbs[:, 4] / (1.0 - bbs[:, 4]))
    cls = np.stack([np.zeros(len(bbs)), z], axis=1).astype(np.float32)
    bbox = np.zeros((len(bbs), 8), np.float32)
    got = _device_detect(cuda, prop, cls, bbox, 2, (1000, 1000), (1.0, 1.0), (1000.0, 1000.0), overlap=case["overlap"])
    want = bbs[case["keep"]]
    assert len(got) == len(want), (got, case["derivation"])
    np.testing.assert_allclose(got[:, :4], want[:, :4], rtol=1e-6, atol=1e-6, err_msg=case["derivation"])
    np.testing.assert_allclose(got[:, 4], want[:, 4], rtol=2e-6, err_msg=case["derivation"])
