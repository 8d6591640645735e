"""Pre-processing in front of the path (examples/kitti_car/run_mscnn_detection.m:64-69,
examples/widerface/run_mscnn_detection.m:70-86): imresize (bicubic, antialiased, uint8) + BGR + mean + CHW.

imresize is MathWorks code outside the reference repository: the restatement (oracle.port.imresize*) follows its
published algorithm and is PARITY-UNPINNED against MATLAB itself.  What is pinned here:
  * CPU: hand-derived known answers of the algorithm (identity at scale 1, the cubic kernel's partition of
    unity, the 2x tap phases 0.25 / 0.75, mirror padding), agreement with torch's antialiased bicubic (same
    a = -0.5 kernel and support widening) away from the borders, and the library's fp64 tap tables bit-exact
    against the restatement (host-only C-ABI call, no GPU);
  * GPU: the device path bit-exact against the restatement (upscale as KITTI, downscale as WIDER, both orders
    of passes, identity dimension), through the C ABI and through Net.set_input_images."""
import numpy as np
import pytest
import torch

from oracle import port

SIZES = [(375, 768), (1242, 2560), (1024, 768), (683, 512), (100, 100), (7, 33), (500, 31), (1, 5), (5, 1)]


def test_contributions_known_answers():
    # scale 1: one tap of weight 1 on the sample itself
    w, idx = port.imresize_contributions(9, 9)
    assert w.shape == (9, 1) and np.array_equal(w[:, 0], np.ones(9)) and np.array_equal(idx[:, 0], np.arange(9))
    # exact 2x upscale: output 2j+1 / 2j+2 sit at phase -0.25 / +0.25 around input j+1: taps cubic(1.25, .25, .75, 1.75)
    w, idx = port.imresize_contributions(8, 16)
    c = lambda t: float(port._cubic(np.array([t]))[0])
    ph = np.array([c(1.25), c(0.25), c(0.75), c(1.75)])
    assert abs(ph.sum() - 1) < 1e-15
    o = 6                                     # 1-based x = 7: u = 3.75, taps 2,3,4,5 (1-based) -> 1..4 (0-based)
    nz = w[o] != 0
    assert np.allclose(w[o][nz], ph[::-1], atol=1e-15) and list(idx[o][nz]) == [1, 2, 3, 4]
    # every row sums to 1 and mirror padding reflects about the border (index -1 -> 0, in -> in-1)
    for i, o in SIZES:
        w, idx = port.imresize_contributions(i, o)
        assert np.allclose(w.sum(1), 1.0, atol=1e-14) and idx.min() >= 0 and idx.max() <= i - 1
    w, idx = port.imresize_contributions(8, 16)
    assert list(idx[0][w[0] != 0]) == [1, 0, 0, 1][-int((w[0] != 0).sum()):]


def test_imresize_matches_torch_antialias_bicubic_in_the_interior():
    rng = np.random.default_rng(3)
    img = rng.random((60, 80, 3))
    for out in [(30, 47), (45, 80), (90, 133), (60, 80), (123, 40)]:
        a = port.imresize(img, out)
        t = torch.nn.functional.interpolate(torch.from_numpy(img).permute(2, 0, 1)[None], size=out, mode="bicubic",
                                            antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
        assert np.abs(a - t)[8:-8, 8:-8].max() < 1e-13


def test_imresize_matches_pillow_bicubic_in_the_interior():
    """A second independent implementation of the same published algorithm: Pillow's BICUBIC resize (a = -0.5, kernel
    support scaled by the reduction factor when shrinking) on float images agrees with the restatement to float32
    precision away from the borders, for enlarging, shrinking and mixed size pairs incl. the KITTI one.  (At the borders
    Pillow renormalises the truncated window where imresize mirrors the image; uint8 images additionally differ by the
    pass order and the per-pass uint8 rounding, which are covered by the hand-derived cases below.)"""
    Image = pytest.importorskip("PIL.Image")
    from oracle import port
    rng = np.random.default_rng(0)
    for (h, w), (H, W) in [((37, 53), (80, 120)), ((90, 130), (40, 50)), ((375, 1242), (768, 2560)), ((64, 64), (31, 97))]:
        img = rng.uniform(0, 255, (h, w)).astype(np.float32)
        ref = port.imresize(img[:, :, None].astype(np.float64), (H, W))[:, :, 0]
        pil = np.asarray(Image.fromarray(img, mode="F").resize((W, H), Image.Resampling.BICUBIC), dtype=np.float64)
        m = 8
        assert np.abs(ref - pil)[m:-m, m:-m].max() < 1e-4, ((h, w), (H, W))


def test_imresize_u8_rounds_and_saturates_per_pass():
    img = np.zeros((8, 8, 3), np.uint8)
    img[:, 4:] = 255                           # a step edge overshoots with a cubic kernel: must clamp to 0..255
    r = port.imresize(img, (16, 16))
    assert r.dtype == np.uint8 and r.min() == 0 and r.max() == 255
    assert np.array_equal(port.imresize(img, (8, 8)), img)
    const = np.full((5, 7, 3), 137, np.uint8)
    assert np.all(port.imresize(const, (11, 3)) == 137)
    assert np.array_equal(port._round_half_away(np.array([0.5, 1.5, 2.4999, 0.49999999999999994, 254.5])),
                          np.array([1.0, 2.0, 2.0, 0.0, 255.0]))


def test_preprocess_layout_and_mean():
    img = np.zeros((4, 6, 3), np.uint8)
    img[..., 0], img[..., 1], img[..., 2] = 10, 20, 30            # R, G, B
    x = port.preprocess(img, (4, 6))
    assert x.shape == (3, 4, 6) and x.dtype == np.float32
    assert np.all(x[0] == 30 - 104) and np.all(x[1] == 20 - 117) and np.all(x[2] == 10 - 123)


def test_library_tap_tables_bit_exact_without_gpu():
    from mscnn_b200 import ops
    for i, o in SIZES:
        w, idx = port.imresize_contributions(i, o)
        w2, idx2 = ops.imresize_contributions(i, o)
        assert w.shape == w2.shape and np.array_equal(w, w2) and np.array_equal(idx, idx2), (i, o)


def test_widerface_net_size():
    from mscnn_b200 import ops
    for args in [(683, 1024, 0, 0, 3072), (4000, 3001, 0, 0, 3072), (100, 50, 0, 0, 3072), (720, 1280, 768, 1024, 3072),
                 (3500, 3500, 0, 0, 2048), (48, 80, 0, 0, 3072)]:
        assert port.widerface_net_size(*args) == ops.widerface_net_size(*args)
    assert port.widerface_net_size(683, 1024, 0, 0, 3072) == (672, 1024)
    assert port.widerface_net_size(4000, 3001, 0, 0, 3072) == (3072, 2304)


def _images(rng, n, h, w):
    img = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    img[0, : h // 2, : w // 2] = 255           # saturating overshoot region
    img[0, h // 2:, : w // 2] = 0
    return img


@pytest.mark.gpu
@pytest.mark.parametrize("in_hw,out_hw", [((94, 311), (192, 640)),      # KITTI-like upscale, width scale larger
                                          ((120, 100), (96, 128)),        # down in h, up in w
                                          ((100, 120), (128, 96)),        # up in h, down in w (width pass first)
                                          ((211, 333), (64, 96)),         # WIDER-like antialiased downscale
                                          ((64, 96), (64, 96)),           # identity
                                          ((40, 64), (40, 128))])
def test_preprocess_gpu_bit_exact(cuda, in_hw, out_hw):
    from mscnn_b200 import ops
    rng = np.random.default_rng(11)
    imgs = _images(rng, 2, *in_hw)
    ref = np.stack([port.preprocess(im, out_hw) for im in imgs])
    pre = ops.Preprocess(in_hw, out_hw)
    got_dev = pre(torch.from_numpy(imgs).cuda()).cpu().numpy()
    got_host = pre(imgs).cpu().numpy()
    assert np.array_equal(got_dev, ref)
    assert np.array_equal(got_host, ref)


@pytest.mark.gpu
def test_preprocess_gpu_full_size_properties(cuda):
    """BASELINE size (375x1242 KITTI frame -> 768x2560): oracle on one image, constant-image and
    channel-permutation properties on the batch."""
    from mscnn_b200 import ops
    rng = np.random.default_rng(5)
    imgs = _images(rng, 2, 375, 1242)
    pre = ops.Preprocess((375, 1242), (768, 2560))
    got = pre(imgs).cpu().numpy()
    assert np.array_equal(got[1], port.preprocess(imgs[1], (768, 2560)))
    const = np.full((1, 375, 1242, 3), 77, np.uint8)
    c = pre(const).cpu().numpy()
    assert np.all(c[0, 0] == 77 - 104) and np.all(c[0, 1] == 77 - 117) and np.all(c[0, 2] == 77 - 123)
    noswap = ops.Preprocess((375, 1242), (768, 2560), swap_rb=False)
    g2 = noswap(np.ascontiguousarray(imgs[..., ::-1])).cpu().numpy()
    assert np.array_equal(g2, got)


@pytest.mark.gpu
def test_net_set_input_images(cuda):
    from mscnn_b200 import models, net as mnet, ops, synth
    mnet.set_precision("fp32")
    h, w = 96, 320
    net = mnet.Net(models.kitti(h, w, 7, False, batch=2))
    net.set_params(synth.make_weights(net.layers()))
    rng = np.random.default_rng(2)
    imgs = _images(rng, 2, 47, 155)
    pre = ops.Preprocess((47, 155), (h, w))
    net.set_input_images("data", pre, imgs)
    torch.cuda.synchronize()
    ref = np.stack([port.preprocess(im, (h, w)) for im in imgs])
    assert np.array_equal(net.blob("data"), ref)
    a = net.forward()
    b = net.forward(data=ref)
    for k in a:
        assert np.array_equal(a[k], b[k])
