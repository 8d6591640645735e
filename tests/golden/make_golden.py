"""Generate the golden vectors under tests/golden/ from oracle/_ref -- the reference's own CPU
layer code compiled verbatim (oracle/build_ref.py).  Run where /root/reference is mounted:

    python tests/golden/make_golden.py

Outputs (committed):
  e2e_7s_192x640.npz   mscnn-7s-576 geometry, input 2x3x192x640 (synthetic seeded image + weights,
                       mscnn_b200/synth.py): the net outputs, all LFCN maps, the proposals, a fixed
                       subsample of conv4_3 / roi_c1 / fc6, per-blob second moments.
  e2e_7s2x_96x320.npz  mscnn-7s-576-2x geometry (Deconvolution + ROI scale 1/4), 1x3x96x320.
  layers.npz           single-layer known-answer vectors for BoxOutput / ROIPooling edge cases.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np

from mscnn_b200 import models, synth
from oracle import ref

OUT = Path(__file__).resolve().parent
SUB = 7919  # subsample stride (prime)


def run_net(proto: str, n: int, h: int, w: int, keep: list[str], sub: list[str]):
    net = ref.RefNet(proto, is_path=False)
    layers = [(nm, t, net.param_shapes(nm)) for nm, t in zip(net.layer_names, net.layer_types)]
    net.set_params(synth.make_weights(layers))
    net.set_blob("data", synth.make_images(n, h, w))
    net.forward()
    out = {}
    for b in keep:
        out[b] = net.blob(b)
    for b in sub:
        x = net.blob(b)
        out[b + "__sub"] = x.reshape(-1)[::SUB].copy()
        out[b + "__m2"] = np.array([np.mean(x.astype(np.float64) ** 2)], dtype=np.float64)
        out[b + "__shape"] = np.array(x.shape, dtype=np.int64)
    return out


def main():
    heads7 = ["LFCN_1_5x5", "LFCN_1_7x7", "LFCN_2_5x5", "LFCN_2_7x7", "LFCN_3_5x5", "LFCN_3_7x7", "LFCN_4_5x5"]
    g = run_net(models.kitti(192, 640, 7, False, batch=2), 2, 192, 640,
                keep=heads7 + ["proposals", "proposals_score", "cls_pred", "bbox_pred"],
                sub=["conv1_2", "conv2_2", "conv3_3", "conv4_3", "conv5_3", "conv6_1", "roi_pool", "roi_c1", "fc6"])
    np.savez_compressed(OUT / "e2e_7s_192x640.npz", **g)
    print("e2e_7s_192x640: proposals", g["proposals"].shape)
    g = run_net(models.kitti(96, 320, 7, True, batch=1), 1, 96, 320,
                keep=["proposals", "proposals_score", "cls_pred", "bbox_pred"],
                sub=["conv4_3", "conv4_3_2x", "roi_pool", "fc6"])
    np.savez_compressed(OUT / "e2e_7s2x_96x320.npz", **g)
    print("e2e_7s2x_96x320: proposals", g["proposals"].shape)

    # ---- single-layer vectors ---------------------------------------------------------------
    rng = np.random.default_rng(1706)
    vec = {}
    # BoxOutput: 2 scales, ties, clamps, border clipping, min_size
    box_proto = '''input: "a" input_dim: 2 input_dim: 9 input_dim: 6 input_dim: 10
input: "b" input_dim: 2 input_dim: 9 input_dim: 3 input_dim: 5
layer { bottom: "a" bottom: "b" top: "rois" top: "rois_score" name: "p" type: "BoxOutput"
  box_output_param { fg_thr: -1 iou_thr: 0.5 nms_type: "IOU" field_w: 40 field_w: 80 field_h: 40 field_h: 80
    downsample_rate: 8 downsample_rate: 16 field_whr: 2 field_xyr: 2 max_nms_num: 40 min_size: 20 } }'''
    net = ref.RefNet(box_proto, is_path=False)
    a = rng.standard_normal((2, 9, 6, 10)).astype(np.float32)
    b = rng.standard_normal((2, 9, 3, 5)).astype(np.float32)
    a[:, :5] = np.round(a[:, :5] * 2) / 2   # score ties
    b[:, :5] = np.round(b[:, :5] * 2) / 2
    a[:, 5:] *= 0.8
    b[:, 5:] *= 0.8
    net.set_blob("a", a)
    net.set_blob("b", b)
    net.forward()
    vec.update(box_a=a, box_b=b, box_rois=net.blob("rois"), box_rois_score=net.blob("rois_score"))
    # ROIPooling with pad_ratio: malformed / outside / half-pixel ROIs
    roi_proto = '''input: "x" input_dim: 2 input_dim: 8 input_dim: 12 input_dim: 20
input: "r" input_dim: 6 input_dim: 5 input_dim: 1 input_dim: 1
layer { bottom: "x" bottom: "r" top: "o" name: "o" type: "ROIPooling" roi_pooling_param { pooled_w: 7 pooled_h: 7 spatial_scale: 0.125 pad_ratio: 0 } }
layer { bottom: "x" bottom: "r" top: "c" name: "c" type: "ROIPooling" roi_pooling_param { pooled_w: 7 pooled_h: 7 spatial_scale: 0.125 pad_ratio: 0.25 } }'''
    net = ref.RefNet(roi_proto, is_path=False)
    x = rng.standard_normal((2, 8, 12, 20)).astype(np.float32)
    r = np.array([[0, 1, 1, 10, 10], [1, 50, 40, 30, 20], [0, 300, 300, 400, 400], [1, 3.5, 4.5, 4.5, 5.5],
                  [0, 0, 0, 159, 95], [1, 20.4, 11.6, 77.5, 60.5]], dtype=np.float32)
    net.set_blob("x", x)
    net.set_blob("r", r.reshape(6, 5, 1, 1))
    net.forward()
    vec.update(roi_x=x, roi_r=r, roi_org=net.blob("o"), roi_ctx=net.blob("c"))
    np.savez_compressed(OUT / "layers.npz", **vec)
    print("layers.npz:", {k: v.shape for k, v in vec.items()})


if __name__ == "__main__":
    main()
