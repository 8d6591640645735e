"""End-to-end parity of the B200 forward path with the reference (golden vectors generated from
oracle/_ref, tests/golden/make_golden.py), through the Caffe-API mirror's Net (C facade).

north_star bar: proposal boxes, class scores and final detections within 1e-3 relative on
identical synthetic inputs (fp32-faithful path).  Because ranking / NMS / round() are discrete, a
last-bit difference in a score can swap or evict a box; rows are therefore matched by content and
the test demands that (a) >= 99.7 % of the reference rows have a partner within 1e-3, (b) the row counts agree
within 0.5 %, (c) >= 99 % of the rows find THE SAME box within 32 places of their own position (adjacent rows whose
scores are closer than their 1e-4 error swap; an inserted / deleted row shifts what follows) and (d) the head outputs
of those pairs are within 1e-3.  (Measured on B200: 99.9-100 % matched.  The residual is inherent to
any implementation that is not bit-identical in the conv sums: the split-bf16 products carry a
2^-18 relative error, which moves an IoU by ~1e-4; with ~4e4 IoU evaluations per image a handful
land that close to the 0.65 threshold, and each flipped suppression adds or removes a row.)  Stage-isolated tests (BoxOutput, ROIPooling fed with the
reference's own inputs) are bit-exact and live in tests/test_detect_gpu.py.
"""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SUB = 7919


def _build(proto, n, h, w, precision="fp32", pool_fusion=True):
    """pool_fusion=False keeps the un-pooled conv1_2 / conv2_2 / conv3_3 blobs materialised (with the
    fusion on they are never written: only the pooling layers read them)."""
    import os
    from mscnn_b200 import net as mnet, synth
    mnet.set_precision(precision)
    if not pool_fusion:
        os.environ["MSCNN_NO_POOL_FUSION"] = "1"
    try:
        net = mnet.Net(proto)
    finally:
        os.environ.pop("MSCNN_NO_POOL_FUSION", None)
    net.set_params(synth.make_weights(net.layers()))
    net.set_input("data", synth.make_images(n, h, w))
    return net


def _rel_ok(got, ref, tol, m2):
    return np.abs(got - ref) <= tol * (np.abs(ref) + np.sqrt(m2))


def _match_rows(got, ref, tol):
    """Fraction of reference rows [img x1 y1 x2 y2 score] with a partner in `got`: same image, every
    corner within tol x the box extent (max(w, h): coordinates near 0 have no meaningful relative
    scale of their own), score within tol x max(|score|, rms(scores))."""
    if len(ref) == 0:
        return 1.0
    srms = float(np.sqrt(np.mean(ref[:, 5].astype(np.float64) ** 2)))
    hit = 0
    for r in ref:
        ext = max(r[3] - r[1], r[4] - r[2], 1.0)
        ok = (got[:, 0] == r[0]) & (np.abs(got[:, 1:5] - r[None, 1:5]).max(axis=1) <= tol * ext) & \
             (np.abs(got[:, 5] - r[5]) <= tol * max(abs(r[5]), srms))
        hit += bool(ok.any())
    return hit / len(ref)


def _row_same(a, r, tol):
    ext = max(r[3] - r[1], r[4] - r[2], 1.0)
    return a[0] == r[0] and np.abs(a[1:5] - r[1:5]).max() <= tol * ext


def _align_rows(got, ref, tol, window=32):
    """Pair every reference proposal with THE SAME box in `got` at (nearly) the same position.  Both lists are in
    descending score order per image; scores carry a ~1e-4 relative error, and with ~1,600 rows over a score range of a
    few units neighbouring scores are often closer than that, so adjacent rows may swap (measured on B200: 6-10 % of
    the rows sit one or two places away); a flipped NMS decision inserts or deletes a row and shifts what follows.  A
    row therefore counts as aligned when its partner (same image, corners within tol of the box extent) sits within
    `window` places; each got row is used once.  Returns index arrays (gi, ri) and the largest displacement."""
    gi, ri, used, worst = [], [], set(), 0
    for j, r in enumerate(ref):
        best = None
        for d in range(window + 1):
            for i in ((j - d, j + d) if d else (j,)):
                if 0 <= i < len(got) and i not in used and _row_same(got[i], r, tol):
                    best = i
                    break
            if best is not None:
                break
        if best is not None:
            used.add(best)
            gi.append(best)
            ri.append(j)
            worst = max(worst, abs(best - j))
    return np.array(gi, dtype=np.int64), np.array(ri, dtype=np.int64), worst


def _check_rows_and_head(out, g, tol, min_match, min_aligned, min_head, tag):
    """Proposals row by row and the detection head on the rows aligned in order.  Thresholds are the figures measured
    on B200 (see the prints), north_star tolerance 1e-3 relative."""
    ref_ps = g["proposals_score"].reshape(-1, 6)
    got_ps = out["proposals_score"].reshape(-1, 6)
    assert abs(len(got_ps) - len(ref_ps)) <= max(2, 0.005 * len(ref_ps)), (len(got_ps), len(ref_ps))
    frac = _match_rows(got_ps, ref_ps, tol)
    gi, ri, disp = _align_rows(got_ps, ref_ps, tol)
    aligned = len(gi) / max(len(got_ps), len(ref_ps))
    in_place = float((gi == ri).mean()) if len(gi) else 0.0
    srms = float(np.sqrt(np.mean(ref_ps[:, 5].astype(np.float64) ** 2)))
    score_ok = float((np.abs(got_ps[gi, 5] - ref_ps[ri, 5]) <= tol * np.maximum(np.abs(ref_ps[ri, 5]), srms)).mean())
    head_ok = {}
    for name in ("cls_pred", "bbox_pred"):
        a, r = out[name].reshape(len(got_ps), -1)[gi], g[name].reshape(len(ref_ps), -1)[ri]
        m2 = float(np.mean(r.astype(np.float64) ** 2))
        # ROIPooling's round() is one more discrete decision: a 0.01-pixel difference in a proposal corner flips it
        # for a fraction of a percent of the ROIs, and such a row then pools different cells.  The head itself is
        # checked element by element on identical proposals in the stage-isolated tests.
        head_ok[name] = float(_rel_ok(a, r, 1e-3, m2).all(axis=1).mean())
    print(f"[{tag}] proposals {len(got_ps)} vs {len(ref_ps)}: matched {frac:.4f}, same box within 32 places {aligned:.4f} "
          f"(at the identical index {in_place:.4f}, largest displacement {disp}), scores ok {score_ok:.4f}, "
          f"head rows within 1e-3 {head_ok}")
    assert frac >= min_match, f"only {frac:.4f} of the reference proposals matched"
    assert aligned >= min_aligned, aligned
    assert score_ok >= min_match
    for name, v in head_ok.items():
        assert v >= min_head, (name, v)


@pytest.mark.parametrize("precision,feat_tol,row_tol,min_match", [("fp32", 1e-3, 1e-3, 0.997), ("bf16", 6e-2, 5e-2, 0.5)])
def test_e2e_kitti_7s_vs_reference(cuda, precision, feat_tol, row_tol, min_match):
    from mscnn_b200 import models
    g = np.load(GOLD / "e2e_7s_192x640.npz")
    net = _build(models.kitti(192, 640, 7, False, batch=2), 2, 192, 640, precision, pool_fusion=False)
    out = net.forward()
    # ---- trunk features (subsampled) --------------------------------------------------------
    worst = {}
    for b in ["conv1_2", "conv2_2", "conv3_3", "conv4_3", "conv5_3", "conv6_1"]:
        x = net.blob(b)
        assert tuple(g[b + "__shape"]) == x.shape
        sub, ref, m2 = x.reshape(-1)[::SUB], g[b + "__sub"], float(g[b + "__m2"][0])
        ok = _rel_ok(sub, ref, feat_tol, m2)
        worst[b] = float(np.max(np.abs(sub - ref) / (np.abs(ref) + np.sqrt(m2))))
        assert ok.mean() >= (0.999 if precision == "fp32" else 0.98), (b, worst[b])
    # ---- proposal heads ---------------------------------------------------------------------
    for b in ["LFCN_1_5x5", "LFCN_1_7x7", "LFCN_2_5x5", "LFCN_2_7x7", "LFCN_3_5x5", "LFCN_3_7x7", "LFCN_4_5x5"]:
        x, ref = net.blob(b), g[b]
        m2 = float(np.mean(ref.astype(np.float64) ** 2))
        assert _rel_ok(x, ref, feat_tol, m2).mean() >= (0.999 if precision == "fp32" else 0.97), b
    # ---- proposals --------------------------------------------------------------------------
    ref_ps = g["proposals_score"].reshape(-1, 6)
    got_ps = out["proposals_score"].reshape(-1, 6)
    assert abs(len(got_ps) - len(ref_ps)) <= max(2, 0.01 * len(ref_ps) if precision == "fp32" else 0.3 * len(ref_ps))
    frac = _match_rows(got_ps, ref_ps, row_tol)
    assert frac >= min_match, f"only {frac:.4f} of the reference proposals matched"
    if precision == "fp32":
        _check_rows_and_head(out, g, row_tol, min_match, 0.99, 0.98, "7s 192x640")
    print(f"[{precision}] worst trunk rel err {worst}; proposals {len(got_ps)} vs {len(ref_ps)}, matched {frac:.4f}")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pool_fusion_is_bit_identical(cuda, precision):
    """Conv epilogue with fused 2x2 max pooling (pool1/2/3) vs the separate pooling kernel: the pooled
    planes and therefore every downstream blob must be bit-identical."""
    from mscnn_b200 import models
    a = _build(models.kitti(192, 640, 8, False, batch=2), 2, 192, 640, precision, pool_fusion=True)
    b = _build(models.kitti(192, 640, 8, False, batch=2), 2, 192, 640, precision, pool_fusion=False)
    oa, ob = a.forward(), b.forward()
    for blob in ["pool1", "pool2", "pool3", "conv4_3", "conv6_1"]:
        assert np.array_equal(a.blob(blob), b.blob(blob)), blob
    for k in oa:
        assert np.array_equal(oa[k], ob[k]), k


def test_head_taps_path_equals_direct_conv(cuda, monkeypatch):
    """The tap-as-N formulation of the LFCN heads (1x1 GEMM + gather) against the direct k x k
    implicit-GEMM of the same layer (MSCNN_NO_HEAD_TAPS=1): same values up to fp32 summation order."""
    from mscnn_b200 import models
    outs = {}
    for mode in ("taps", "direct"):
        if mode == "direct":
            monkeypatch.setenv("MSCNN_NO_HEAD_TAPS", "1")
        net = _build(models.kitti(96, 320, 8, False, batch=2), 2, 96, 320)
        net.forward_only(end="LFCN_4_7x7")
        outs[mode] = {n: net.blob(n) for n in net.layer_names if n.startswith("LFCN")}
    for n in outs["taps"]:
        a, b = outs["taps"][n], outs["direct"][n]
        rms = float(np.sqrt(np.mean(b.astype(np.float64) ** 2)))
        # the direct path accumulates 3*k*k*8 k-blocks into one TMEM accumulator, whose fp32 adds truncate
        # (a systematic ~2^-24 per step: measured 1.4e-4 over 2400 steps); the taps path sums 25/49 short chains
        assert np.abs(a - b).max() <= 5e-4 * rms + 1e-6, (n, float(np.abs(a - b).max()), rms)


def test_e2e_kitti_7s_2x_vs_reference(cuda):
    """The -2x variant: Deconvolution upsampling of conv4_3 and ROI pooling at scale 1/4."""
    from mscnn_b200 import models
    g = np.load(GOLD / "e2e_7s2x_96x320.npz")
    net = _build(models.kitti(96, 320, 7, True, batch=1), 1, 96, 320)
    out = net.forward()
    for b in ["conv4_3", "conv4_3_2x", "roi_pool", "fc6"]:
        x = net.blob(b)
        if b in ("roi_pool", "fc6") and tuple(g[b + "__shape"]) != x.shape:
            continue   # a discrete flip changed R; the row-matched checks below still apply
        sub, ref, m2 = x.reshape(-1)[::SUB], g[b + "__sub"], float(g[b + "__m2"][0])
        assert _rel_ok(sub, ref, 1e-3, m2).mean() >= 0.998, b
    ref_ps, got_ps = g["proposals_score"].reshape(-1, 6), out["proposals_score"].reshape(-1, 6)
    assert abs(len(got_ps) - len(ref_ps)) <= 3
    assert _match_rows(got_ps, ref_ps, 1e-3) >= 0.97


def test_head_stage_isolated_vs_reference(cuda):
    """ROIPooling x2 + Concat + roi_c1 + fc6 + cls/bbox fed with the REFERENCE's proposals (no
    discrete decision upstream differs): every output row must be within 1e-3."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    from mscnn_b200 import models, synth
    proto = models.kitti(96, 320, 7, False, batch=1)
    rnet = ref.RefNet(proto, is_path=False)
    layers = [(n, t, rnet.param_shapes(n)) for n, t in zip(rnet.layer_names, rnet.layer_types)]
    w = synth.make_weights(layers)
    rnet.set_params(w)
    img = synth.make_images(1, 96, 320)
    rnet.set_blob("data", img)
    rnet.forward()
    net = _build(proto, 1, 96, 320)
    net.forward_only()                                   # our own trunk
    props = rnet.blob("proposals")
    conv4_3 = rnet.blob("conv4_3")
    # inject the reference's conv4_3 and proposals into the blobs the head reads
    for k in (2, 3):
        net.set_input(f"conv4_3_relu4_3_0_split_{k}", conv4_3)
    for k in (0, 1):
        net.set_input(f"proposals_proposals_0_split_{k}", props)
    net.forward_only(start="roi_pool_org")
    for name in ("roi_c1", "fc6", "cls_pred", "bbox_pred"):
        a, r = net.blob(name), rnet.blob(name)
        assert a.shape == r.shape, name
        m2 = float(np.mean(r.astype(np.float64) ** 2))
        err = np.abs(a - r) / (np.abs(r) + np.sqrt(m2))
        assert err.max() <= 1e-3, (name, float(err.max()))


def test_final_detections_vs_oracle(cuda):
    """net outputs -> mscnn_net_detect (device) == the restated MATLAB post-process on the SAME
    net outputs (isolates the post-process kernels end to end)."""
    import torch
    from mscnn_b200 import models, net as mnet
    from oracle import port
    net = _build(models.kitti(192, 640, 7, False, batch=2), 2, 192, 640)
    out = net.forward()
    cfg = mnet.kitti_detect_cfg(192, 640)
    dets = torch.zeros((2, cfg.max_rois_per_image, 5), device=cuda)
    cnt = torch.zeros(2, dtype=torch.int32, device=cuda)
    net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
    torch.cuda.synchronize()
    dets, cnt = dets.cpu().numpy(), cnt.cpu().numpy()
    ps = out["proposals_score"].reshape(-1, 6)
    start = 0
    for i in range(2):
        n_i = net.num_proposals(i)
        sl = slice(start, start + n_i)
        ref = port.detect_postprocess(ps[sl], out["cls_pred"][sl], out["bbox_pred"][sl], cls_id=2, net_hw=(192, 640))
        assert cnt[i] == len(ref)
        np.testing.assert_allclose(dets[i, : cnt[i]], ref, rtol=1e-5, atol=1e-5)
        start += n_i
    assert net.num_proposals() == len(ps)


def test_empty_image_yields_dummy_roi(cuda):
    """All-background image: BoxOutput emits the dummy ROI and the head still runs (R = 1)."""
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    net = mnet.Net(models.kitti(96, 320, 7, False, batch=1))
    w = synth.make_weights(net.layers())
    for k in w:
        if k.startswith("LFCN"):
            w[k][1][0] = 1e4          # background bias dominates every anchor
    net.set_params(w)
    out = net.forward(data=synth.make_images(1, 96, 320))
    assert net.num_proposals() == 0
    assert out["proposals_score"].shape[0] == 1 and not out["proposals_score"].any()
    assert net.blob("proposals").reshape(-1).tolist() == [0.0, 1.0, 1.0, 10.0, 10.0]
    assert out["cls_pred"].shape == (1, 5) and np.isfinite(out["cls_pred"]).all()


# ------------------------------------------------------------------------------ cascade nets
def _stage_rows(blobs, stage, cls_col):
    """[img x1 y1 x2 y2 prob] rows of one cascade stage (what run_cascademscnn.m:99-126 consumes)."""
    bb = blobs[f"output_bbox_{stage}"].reshape(-1, 5)
    pr = blobs[f"cls_prob_{stage}"].reshape(len(bb), -1)[:, cls_col]
    return np.concatenate([bb, pr[:, None]], axis=1)


@pytest.mark.parametrize("kind", ["kitti", "wider"])
def test_e2e_cascade_vs_reference(cuda, kind):
    """Cascade deploy nets end to end vs golden vectors from the verbatim reference build: three
    detection stages chained by DecodeBBox on the device; (wider) ROIAlign + AVE pooling, third-stage
    heads sharing weights by ParamSpec name, Eltwise average of the three probabilities."""
    from mscnn_b200 import models
    if kind == "kitti":
        g = np.load(GOLD / "e2e_cascade_kitti_96x320.npz")
        net = _build(models.kitti_cascade(96, 320, batch=2), 2, 96, 320)
        feats, cls_col = ["conv4_3_2x", "roi_pool", "fc6"], 1
    else:
        g = np.load(GOLD / "e2e_cascade_wider_128x192.npz")
        net = _build(models.widerface_cascade(128, 192, batch=2), 2, 128, 192)
        feats, cls_col = ["conv4_3", "roi_grid_org", "roi_grid_ctx", "roi_pool", "fc6"], 1
    net.forward_only()
    same_rows = net.blob("proposals").shape == g["proposals"].shape
    for b in feats:
        x = net.blob(b)
        if tuple(g[b + "__shape"]) != x.shape:
            assert not same_rows
            continue
        sub, ref, m2 = x.reshape(-1)[::SUB], g[b + "__sub"], float(g[b + "__m2"][0])
        frac = _rel_ok(sub, ref, 1e-3, m2).mean()
        # trunk features are row-independent; ROI features only line up when the proposal lists do
        assert frac >= (0.998 if b.startswith("conv") else 0.97), (b, frac)
    ref_ps, got_ps = g["proposals_score"].reshape(-1, 6), net.blob("proposals_score").reshape(-1, 6)
    assert abs(len(got_ps) - len(ref_ps)) <= max(3, len(ref_ps) // 100)
    assert _match_rows(got_ps, ref_ps, 1e-3) >= 0.97
    blobs = {k: net.blob(k) for k in ["output_bbox_1st", "output_bbox_2nd", "output_bbox_3rd", "cls_prob_1st",
                                      "cls_prob_2nd", "cls_prob_3rd"]}
    # ROIPooling's round() makes stage k+1 features a discontinuous function of stage k's boxes, so a
    # growing (small) share of rows legitimately lands on a neighbouring bin layout; ROIAlign does not.
    floors = {"kitti": (0.95, 0.90, 0.85), "wider": (0.97, 0.97, 0.97)}[kind]
    for stage, floor in zip(["1st", "2nd", "3rd"], floors):
        m = _match_rows(_stage_rows(blobs, stage, cls_col), _stage_rows(g, stage, cls_col), 1e-3)
        assert m >= floor, (stage, m)
    if kind == "wider":
        avg = net.blob("cls_prob_3rd_avg")
        manual = (net.blob("cls_prob_1st_3rd") * np.float32(0.33333333) + net.blob("cls_prob_2nd_3rd") * np.float32(0.33333333)
                  + net.blob("cls_prob_3rd") * np.float32(0.33333333))
        np.testing.assert_allclose(avg, manual, rtol=1e-6)
        np.testing.assert_allclose(avg.sum(1), 1.0, rtol=1e-5)


@pytest.mark.parametrize("kind", ["kitti", "wider"])
def test_cascade_stages_isolated_vs_reference(cuda, kind):
    """The three detection stages fed with the REFERENCE's feature map and proposals (run live through
    oracle/_ref): no discrete decision upstream differs, so rows line up one to one."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not present")
    from mscnn_b200 import models, synth
    if kind == "kitti":
        proto, (n, h, w), feat, first = models.kitti_cascade(96, 320, batch=2), (2, 96, 320), "conv4_3_2x", "roi_pool_org"
    else:
        proto, (n, h, w), feat, first = models.widerface_cascade(128, 192, batch=2), (2, 128, 192), "conv4_3", "roi_grid_org"
    rnet = ref.RefNet(proto, is_path=False)
    layers = [(nm, t, rnet.param_shapes(nm)) for nm, t in zip(rnet.layer_names, rnet.layer_types)]
    rnet.set_params(synth.make_weights(layers))
    rnet.set_blob("data", synth.make_images(n, h, w))
    rnet.forward()
    net = _build(proto, n, h, w)
    net.forward_only()
    fmap, props = rnet.blob(feat), rnet.blob("proposals")
    for b in net.blob_names:      # every Split top of the feature map / the proposals that the stages read
        if b.startswith(feat + "_") and "_split_" in b:
            net.set_input(b, fmap)
        if b.startswith("proposals_proposals_0_split_"):
            net.set_input(b, props)
    net.forward_only(start=first)
    names = ["cls_pred", "bbox_pred", "proposals_2nd", "cls_pred_2nd", "bbox_pred_2nd", "proposals_3rd", "cls_pred_3rd",
             "bbox_pred_3rd", "output_bbox_1st", "output_bbox_2nd", "output_bbox_3rd", "cls_prob_1st", "cls_prob_2nd",
             "cls_prob_3rd"] + (["cls_prob_1st_3rd", "cls_prob_2nd_3rd", "cls_prob_3rd_avg"] if kind == "wider" else [])
    worst = {}
    for name in names:
        a, r = net.blob(name), rnet.blob(name)
        assert a.shape == r.shape, name
        a, r = a.reshape(a.shape[0], -1), r.reshape(r.shape[0], -1)
        if name.startswith(("proposals", "output_bbox")):     # boxes: relative to the box extent
            ext = np.maximum(np.maximum(r[:, 3] - r[:, 1], r[:, 4] - r[:, 2]), 1.0)[:, None]
            err = np.abs(a - r) / ext
        else:
            m2 = float(np.mean(r.astype(np.float64) ** 2))
            err = np.abs(a - r) / (np.abs(r) + np.sqrt(m2))
        ok_rows = (err.max(axis=1) <= 1e-3).mean()
        worst[name] = (float(ok_rows), float(err.max()))
        first_stage = name in ("cls_pred", "bbox_pred", "proposals_2nd", "output_bbox_1st", "cls_prob_1st")
        if kind == "wider" or first_stage:
            # ROIAlign is continuous in the boxes, and the first stage pools the reference's own ROIs:
            # every row within 1e-3 (measured on B200: <= 1.1e-4)
            assert ok_rows == 1.0, (name, worst[name])
        else:
            # later ROIPooling stages round() boxes that carry a 1e-5 error: measured 99.4-99.7 % of rows
            assert ok_rows >= 0.98, (name, worst[name])
    print(kind, worst)


def test_cascade_final_detections_vs_oracle(cuda):
    """net outputs -> mscnn_net_detect_cascade (device) == the restated run_cascademscnn.m post-process on
    the SAME net outputs, for the third stage and (wider) the averaged probabilities."""
    import torch
    from mscnn_b200 import capi, models
    from oracle import port
    for kind in ("kitti", "wider"):
        if kind == "kitti":
            net, hw, ncls, prob_blob = _build(models.kitti_cascade(96, 320, batch=2), 2, 96, 320), (96, 320), 5, None
        else:
            net, hw, ncls, prob_blob = _build(models.widerface_cascade(128, 192, batch=2), 2, 128, 192), (128, 192), 2, "cls_prob_3rd_avg"
        net.forward_only()
        cfg = capi.DetectCfg()
        cfg.num_cls, cfg.cls_id = ncls, 2
        cfg.nms_overlap = 0.5 if kind == "kitti" else 0.3
        cfg.ratio_h, cfg.ratio_w = 1.25, 0.8
        cfg.org_h, cfg.org_w = hw[0] / 1.25, hw[1] / 0.8
        cfg.max_rois_per_image = 3000
        dets = torch.zeros((2, cfg.max_rois_per_image, 5), device=cuda)
        cnt = torch.zeros(2, dtype=torch.int32, device=cuda)
        net.detect_cascade(cfg, dets.data_ptr(), cnt.data_ptr(), stage="3rd", cls_prob=prob_blob)
        torch.cuda.synchronize()
        dets, cnt = dets.cpu().numpy(), cnt.cpu().numpy()
        props = net.blob("proposals_3rd").reshape(-1, 5)
        prob = net.blob(prob_blob or "cls_prob_3rd").reshape(len(props), -1)
        outb = net.blob("output_bbox_3rd").reshape(-1, 5)
        start = 0
        for i in range(2):
            n_i = net.num_proposals(i)
            sl = slice(start, start + n_i)
            ref = port.cascade_detect_postprocess(props[sl], prob[sl], outb[sl], cls_id=2, overlap=cfg.nms_overlap,
                                                  ratios=(1.25, 0.8), org_hw=(cfg.org_h, cfg.org_w))
            assert cnt[i] == len(ref), (kind, i)
            assert np.array_equal(dets[i, : cnt[i]], ref), (kind, i)
            start += n_i


def test_e2e_kitti_8s_vs_reference(cuda):
    """The bench architecture itself (mscnn-8s-768: eight heads, the last pair on pool6) at 2x3x192x640, default net
    configuration (pooling fused into conv1_2 / conv2_2 / conv3_3, so those three blobs are never materialised)."""
    from mscnn_b200 import models
    g = np.load(GOLD / "e2e_8s_192x640.npz")
    net = _build(models.kitti(192, 640, 8, False, batch=2), 2, 192, 640, "fp32")
    out = net.forward()
    for b in ["conv4_3", "conv5_3", "conv6_1"]:
        x = net.blob(b)
        assert tuple(g[b + "__shape"]) == x.shape
        sub, ref, m2 = x.reshape(-1)[::SUB], g[b + "__sub"], float(g[b + "__m2"][0])
        assert _rel_ok(sub, ref, 1e-3, m2).mean() >= 0.999, b
    heads = [k for k in g.files if k.startswith("LFCN_")]
    assert len(heads) == 8
    for b in heads:
        x, ref = net.blob(b), g[b]
        m2 = float(np.mean(ref.astype(np.float64) ** 2))
        assert _rel_ok(x, ref, 1e-3, m2).mean() >= 0.999, b
    _check_rows_and_head(out, g, 1e-3, 0.997, 0.99, 0.98, "8s 192x640")


def test_e2e_widerface_vs_reference(cuda):
    """WIDER FACE mscnn-12s-2x geometry (BASELINE.json configs[4]): twelve 1x1 heads of 6 channels, AVE pool6, bbox
    normalisation inside BoxOutput, Deconvolution 2x, 5x5 ROI pooling, fc6 2048 -- against the reference's own CPU
    layers (tests/golden/e2e_wider_128x192.npz), same acceptance rules as the KITTI nets."""
    from mscnn_b200 import models
    g = np.load(GOLD / "e2e_wider_128x192.npz")
    net = _build(models.widerface(128, 192, batch=2), 2, 128, 192, "fp32")
    out = net.forward()
    for b in ["conv4_3", "conv4_3_2x", "conv5_3", "pool6"]:
        x = net.blob(b)
        assert tuple(g[b + "__shape"]) == x.shape
        sub, ref, m2 = x.reshape(-1)[::SUB], g[b + "__sub"], float(g[b + "__m2"][0])
        assert _rel_ok(sub, ref, 1e-3, m2).mean() >= 0.999, b
    heads = [k for k in g.files if k.startswith("LFCN_")]
    assert len(heads) == 12
    for b in heads:
        x, ref = net.blob(b), g[b]
        m2 = float(np.mean(ref.astype(np.float64) ** 2))
        assert _rel_ok(x, ref, 1e-3, m2).mean() >= 0.999, b
    _check_rows_and_head(out, g, 1e-3, 0.997, 0.99, 0.98, "wider 128x192")


def test_async_input_upload_pipelines_correctly(cuda):
    """mscnn_net_set_blob_async: the upload for the NEXT forward is issued while the current forward is still
    running; every forward must see exactly the batch uploaded for it (two alternating batches, several rounds),
    and the outputs must equal those of the synchronous set_input path bit for bit."""
    import torch
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    h, w = 96, 320
    net = mnet.Net(models.kitti(h, w, 7, False, batch=2))
    net.set_params(synth.make_weights(net.layers()))
    batches = [torch.from_numpy(synth.make_images(2, h, w, first_index=i * 2)).pin_memory() for i in range(2)]
    want = []
    for b in batches:
        out = net.forward(data=b.numpy())
        want.append({k: v.copy() for k, v in out.items()})
    assert not np.array_equal(want[0]["cls_pred"], want[1]["cls_pred"])
    net.set_input_async("data", batches[0])
    for step in range(6):
        net.forward_only()
        net.set_input_async("data", batches[(step + 1) % 2])     # while this forward's ROI head is still queued
        got = {o: net.blob(o) for o in net.outputs}
        for k in got:
            assert np.array_equal(got[k], want[step % 2][k]), (step, k)
