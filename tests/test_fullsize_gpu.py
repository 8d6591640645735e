"""BASELINE.json's FULL sizes (mscnn-8s-768: 3 x 768 x 2560 images) through size-independent properties, where a full
oracle run would take minutes:

  * exact scaling: a power-of-two factor on the input commutes with every rounding of the fp32-faithful
    convolution (bf16 splits, fp32 accumulation), so conv(2 x) == 2 conv(x) bit for bit (bias-free, ReLU keeps it);
  * strips: a few rows of the full-size output (top border, interior tile boundaries, bottom border) against an fp64
    convolution of exactly those rows;
  * batch-order independence: images A, B forwarded as (A, B) and as (B, A) give each image the same proposals and
    head outputs bit for bit, and every image's proposals come out in non-increasing score order (the reference's
    std::sort, box_output_layer.cpp:166-179);
  * fused == unfused at full size for the row-pair / register-pooling path of conv1_2.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
H, W = 768, 2560


def _strip_ref(x, wt, b, rows, relu=True):
    """fp64 conv of output rows [r0, r1) only (3x3, pad 1)."""
    r0, r1 = rows
    lo, hi = max(r0 - 1, 0), min(r1 + 1, x.shape[2])
    xs = x[:, :, lo:hi].double()
    pad_top, pad_bot = (1 if r0 == 0 else 0), (1 if r1 == x.shape[2] else 0)
    xs = F.pad(xs, (1, 1, pad_top, pad_bot))
    y = F.conv2d(xs, wt.double(), None if b is None else b.double())
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("cin,cout", [(64, 64), (64, 128)])
def test_conv_full_size_scaling_and_strips(cuda, cin, cout):
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    h, w = (H, W) if cout == 64 else (H // 2, W // 2)          # conv1_2 / conv2_1 geometry
    x = torch.randn((1, cin, h, w), generator=g).to(cuda)
    wt = (torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5).to(cuda)
    wp = ops.pack_conv_weights(wt, None, True)
    y1 = ops.conv_forward(ops.nchw_to_planes(x, True), wp, 1, relu=True)
    y2 = ops.conv_forward(ops.nchw_to_planes(x * 2.0, True), wp, 1, relu=True)
    torch.cuda.synchronize()
    assert torch.equal(y2.hi.float(), y1.hi.float() * 2.0) and torch.equal(y2.lo.float(), y1.lo.float() * 2.0)
    got = ops.planes_to_nchw(y1)
    for rows in [(0, 3), (h // 2 - 1, h // 2 + 2), (h - 3, h)]:
        ref = _strip_ref(x, wt, None, rows)
        err = (got[:, :, rows[0]:rows[1]].double() - ref).abs()
        assert float(err.max()) <= 1e-4 * float(ref.abs().max()), (rows, float(err.max()))
    # fused 2x2 pooling (row-pair tiles, pooling in registers) == conv -> pool, bit for bit, at full size
    p_only = ops.conv_forward(ops.nchw_to_planes(x, True), wp, 1, relu=True, pool="only")
    p_ref = ops.pool_forward(y1, 2, 2)
    torch.cuda.synchronize()
    assert torch.equal(p_only.hi, p_ref.hi) and torch.equal(p_only.lo, p_ref.lo)


def test_conv1_1_full_size_strips(cuda):
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(6)
    x = (torch.randint(0, 256, (2, 3, H, W), generator=g).float()
         - torch.tensor([104.0, 117.0, 123.0]).view(1, 3, 1, 1)).to(cuda).contiguous()
    wt = (torch.randn((64, 3, 3, 3), generator=g) * (2.0 / 27) ** 0.5 / 64).to(cuda)
    b = (torch.randn((64,), generator=g) * 0.1).to(cuda)
    got = ops.planes_to_nchw(ops.conv1_tc_forward(x, wt, b, relu=True, split=True))
    torch.cuda.synchronize()
    for rows in [(0, 2), (383, 386), (H - 2, H)]:
        ref = _strip_ref(x, wt, b, rows)
        err = (got[:, :, rows[0]:rows[1]].double() - ref).abs()
        assert float(err.max()) <= 2e-5 * float(ref.abs().max()), (rows, float(err.max()))
    # columns across every 128-pixel tile boundary of one row
    ref = _strip_ref(x, wt, b, (100, 101))[0, :, 0]
    cols = torch.tensor([c for k in range(1, W // 128) for c in (128 * k - 1, 128 * k)], device=cuda)
    assert float((got[0, :, 100][:, cols].double() - ref[:, cols]).abs().max()) <= 2e-5 * float(ref.abs().max())


def test_net_full_size_batch_order_and_sorted_proposals(cuda):
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    net = mnet.Net(models.kitti(H, W, 8, False, batch=2))
    net.set_params(synth.make_weights(net.layers()))
    imgs = synth.make_images(2, H, W)
    res = []
    for order in ([0, 1], [1, 0]):
        out = net.forward(data=np.ascontiguousarray(imgs[order]))
        ps = out["proposals_score"].reshape(-1, 6)
        per = {}
        for slot, img in enumerate(order):
            sel = ps[:, 0] == slot
            per[img] = (ps[sel][:, 1:], out["cls_pred"].reshape(len(ps), -1)[sel], out["bbox_pred"].reshape(len(ps), -1)[sel])
            sc = ps[sel][:, 5]
            assert len(sc) > 100 and np.all(sc[:-1] >= sc[1:])          # descending score order per image
        res.append(per)
    for img in (0, 1):
        for a, b in zip(res[0][img], res[1][img]):
            assert np.array_equal(a, b)
