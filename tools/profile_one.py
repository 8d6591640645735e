"""Small, deterministic workloads for `ncu --set full` captures of ONE kernel at a time (B200_PROFILING.md: keep the
profiled command short).  Each mode launches the kernel of interest a few times on the bench's real layer geometry.

    python tools/profile_one.py conv3_2      # conv_igemm_kernel<256, true>: 8 x 256 x 192 x 640 -> 256, fp32-faithful (CTA pairs)
    python tools/profile_one.py conv1_2      # conv_igemm_kernel<64, false>: 8 x 64 x 768 x 2560 -> 64 with the pooling fused (vpool)
    python tools/profile_one.py conv2_2      # conv_igemm_kernel<128, false>: 8 x 128 x 384 x 1280 -> 128, pooled
    python tools/profile_one.py net8s [n]    # n (default 2) x (forward + detect) of the bench net, batch 8, 768 x 2560
    python tools/profile_one.py net7s2x [n]  # the -2x net (Deconvolution), batch 1, 576 x 1920
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch


def conv(n, cin, h, w, cout, pool):
    from mscnn_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn((n, cin, h, w), generator=g).to(dev)
    wt = (torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5).to(dev)
    b = torch.zeros(cout, device=dev)
    xp = ops.nchw_to_planes(x, True)
    del x
    wp = ops.pack_conv_weights(wt, b, True)
    for _ in range(4):
        y = ops.conv_forward(xp, wp, 1, relu=True, pool="only" if pool else None)
    torch.cuda.synchronize()
    return y


def net(kind, reps):
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    if kind == "net8s":
        b, h, w, proto = 8, 768, 2560, models.kitti(768, 2560, 8, False, batch=8)
    else:
        b, h, w, proto = 1, 576, 1920, models.kitti(576, 1920, 7, True, batch=1)
    n = mnet.Net(proto)
    n.set_params(synth.make_weights(n.layers()))
    img = torch.from_numpy(synth.make_images(b, h, w)).cuda()
    cfg = mnet.kitti_detect_cfg(h, w)
    dets = torch.zeros((b, cfg.max_rois_per_image, 5), device="cuda")
    cnt = torch.zeros(b, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        n.set_input("data", img)
        n.forward_only()
        n.detect(cfg, dets.data_ptr(), cnt.data_ptr())
    torch.cuda.synchronize()
    print(kind, "proposals", n.num_proposals(), "detections", int(cnt.sum().item()))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "conv3_2":
        conv(8, 256, 192, 640, 256, False)
    elif what == "conv1_2":
        conv(8, 64, 768, 2560, 64, True)
    elif what == "conv2_2":
        conv(8, 128, 384, 1280, 128, True)
    elif what == "conv2_1":
        conv(8, 64, 384, 1280, 128, False)
    else:
        net(what, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
