"""Calibration run for mscnn_b200/synth.py (CPU, needs oracle/_ref): prints activation second
moments, proposal statistics and prediction-head spreads for the synthetic init."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import ref
from mscnn_b200 import synth

proto = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples/kitti_car/mscnn-7s-576/mscnn_deploy.prototxt"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (192, 640)
txt = open(proto).read()
import re
dims = re.findall(r"input_dim:\s*(\d+)", txt)
txt = txt.replace(f"input_dim: {dims[2]}", f"input_dim: {H}", 1).replace(f"input_dim: {dims[3]}", f"input_dim: {W}", 1)
net = ref.RefNet(txt, is_path=False)
layers = [(n, t, net.param_shapes(n)) for n, t in zip(net.layer_names, net.layer_types)]
net.set_params(synth.make_weights(layers))
net.set_blob("data", synth.make_images(1, H, W))
t = time.time(); ms = net.forward(); print("forward %.1f s" % (time.time() - t))
for b in ["conv1_1", "conv4_3", "loss1_conv1", "rpn_1_conv", "conv5_3", "conv6_1", "pool6", "roi_pool", "roi_c1", "fc6"]:
    try:
        x = net.blob(b).astype(np.float64); print(f"{b:12s} {x.shape} m2={np.mean(x*x):.3f} max={x.max():.2f}")
    except KeyError:
        pass
for n, t in zip(net.layer_names, net.layer_types):
    if n.startswith("LFCN"):
        x = net.blob(n)[0]; cls = x.shape[0] - 4
        s = x[1:cls].max(0) - x[0]
        print(f"{n:14s} cls std {x[1:cls].std():.2f} score mean {s.mean():.2f} std {s.std():.2f} pass {np.mean(s >= -5):.2f} "
              f"box std {x[cls:].std():.3f} |dxy|>0.5 {np.mean(np.abs(x[cls:cls+2]) > 0.5):.3f} |dwh|>ln2 {np.mean(np.abs(x[cls+2:]) > 0.693):.3f}")
p = net.blob("proposals_score"); print("proposals", p.shape, "score range", p[:, 5].min(), p[:, 5].max())
w = p[:, 3] - p[:, 1]; h = p[:, 4] - p[:, 2]; print("w mean %.1f h mean %.1f" % (w.mean(), h.mean()))
c = net.blob("cls_pred"); bb = net.blob("bbox_pred"); print("cls_pred std %.3f bbox_pred std %.3f" % (c.std(), bb.std()))
top = sorted(ms.items(), key=lambda kv: -kv[1])[:8]; print([(k, round(v)) for k, v in top])
