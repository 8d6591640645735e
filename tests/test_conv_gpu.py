"""Parity of the tcgen05 implicit-GEMM convolution (mscnn_conv_forward) against an fp64
convolution of the same operands.

Reference semantics: ConvolutionLayer::Forward_cpu = im2col + sgemm + bias
(/root/reference/src/caffe/layers/conv_layer.cpp:25-40, base_conv_layer.cpp:257-280) and the
reference's own test tolerance for it, 1e-4 absolute on O(1) data
(src/caffe/test/test_convolution_layer.cpp:231-265).  Tolerances used here:
  split-bf16 ("fp32-faithful") path: |err| <= 1e-4 * (|ref| + rms(ref)).  The 3-term split has a
  per-product relative error of ~2^-16; over K random-sign terms that is ~1e-5 rms(ref) typical,
  ~5e-5 rms(ref) at the 4.5-sigma tail of a 32k-element tensor (measured on B200).
  plain bf16 path: exact products of bf16-rounded operands, so vs fp64 conv of the ROUNDED
  operands |err| <= 1e-5 scale for fp32 output, one bf16 ulp (2^-8 rel) for bf16 planes.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_conv(x, w, b, pad, relu):
    y = F.conv2d(x.double(), w.double(), None if b is None else b.double(), padding=pad)
    return torch.relu(y) if relu else y


def _report(name, got, ref, tol_rel, tol_abs):
    err = (got.double() - ref).abs()
    lim = tol_rel * ref.abs() + tol_abs
    bad = err > lim
    if bad.any():
        idx = bad.nonzero()
        msg = [f"{name}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; "
               f"max err {float(err.max()):.3e}, ref rms {float(ref.pow(2).mean().sqrt()):.3e}"]
        for d in range(idx.shape[1]):
            u = idx[:, d].unique()
            msg.append(f"  dim{d}: {u.numel()} distinct bad indices, first {u[:16].tolist()}")
        for i in idx[:8]:
            t = tuple(i.tolist())
            msg.append(f"  at {t}: got {float(got[t]):.6f} ref {float(ref[t]):.6f}")
        pytest.fail("\n".join(msg))


CASES = [
    # name,            N, Cin, H,  W, Cout, k, pad
    ("trunk64",        1, 64, 16, 32, 64, 3, 1),
    ("trunk128",       2, 64, 12, 40, 128, 3, 1),
    ("trunk256",       1, 128, 24, 20, 256, 3, 1),
    ("trunk512_2nt",   1, 128, 8, 16, 512, 3, 1),
    ("ragged",         3, 64, 9, 30, 64, 3, 1),
    ("roi_c1_like",    7, 128, 7, 7, 128, 3, 0),
    ("conv1x1",        1, 64, 8, 24, 64, 1, 0),
    ("deepK",          1, 512, 6, 40, 128, 3, 1),
    # enough M tiles that a CTA tile holds several 128-pixel sub-tiles (mt = 2 / 4 path)
    ("mt_n64",         2, 64, 128, 160, 64, 3, 1),
    ("mt_n128",        2, 64, 128, 320, 128, 3, 1),
    # W a multiple of 128 with narrow N: 128x1 boxes -> row-share mode (one 130-pixel activation tile per dy,
    # the dx taps are descriptor row shifts), one and two channel chunks, more tiles than ring slots
    ("rowshare_n64",   1, 64, 6, 256, 64, 3, 1),
    ("rowshare_n128",  2, 128, 5, 384, 128, 3, 1),
    ("rowshare_many",  2, 64, 40, 512, 64, 3, 1),
    # BLOCK_N = 256 on 128x1 boxes: two-ring engine with the row-share halo (a_taps = 3), single-tile weight slots
    # (b_split) and, by default, CTA pairs -- the variant conv3_x takes at 768x2560 (conv_igemm.cu "BLOCK_N = 256,
    # fp32-faithful").  Even / odd M-tile counts (odd: the pair build's phantom M tile), one and two n tiles,
    # 4 and 8 channel chunks, more pair tiles than clusters.
    ("ring256_halo",   1, 256, 6, 256, 256, 3, 1),
    ("ring256_odd",    1, 256, 5, 384, 256, 3, 1),
    ("ring256_2nt",    2, 256, 5, 640, 512, 3, 1),
    ("ring256_c512",   1, 512, 3, 384, 512, 3, 1),
    ("ring256_many",   2, 256, 40, 640, 256, 3, 1),
    # BLOCK_N = 256 on 2-D boxes (a_taps = 1): conv4_x / conv5_x geometry, odd tile count
    ("ring256_box2d",  1, 256, 24, 40, 512, 3, 1),
    ("ring256_box2d_odd", 3, 256, 12, 40, 256, 3, 1),
]


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_planes(cuda, case, split):
    from mscnn_b200 import ops
    name, n, cin, h, w, cout, k, pad = case
    g = torch.Generator(device="cpu").manual_seed(1706)
    x = torch.randn((n, cin, h, w), generator=g).to(cuda)
    wt = (torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(cuda)
    b = (torch.randn((cout,), generator=g) * 0.1).to(cuda)
    xp = ops.nchw_to_planes(x, split)
    wp = ops.pack_conv_weights(wt, b, split)
    yp = ops.conv_forward(xp, wp, pad, relu=True)
    torch.cuda.synchronize()
    got = ops.planes_to_nchw(yp)
    torch.cuda.synchronize()
    if split:
        ref = _ref_conv(x, wt, b, pad, True)
        rms = float(ref.pow(2).mean().sqrt())
        _report(name, got, ref, 1e-4, 1e-4 * rms)
    else:
        ref = _ref_conv(x.bfloat16().float(), wt.bfloat16().float(), b, pad, True)
        rms = float(ref.pow(2).mean().sqrt())
        _report(name, got, ref, 2.0 ** -8, 1e-5 * rms)


VARIANT_CASES = [c for c in CASES if c[0].startswith("ring256") or c[0] in ("trunk256", "trunk512_2nt", "rowshare_n128")]


@pytest.mark.parametrize("case", VARIANT_CASES, ids=[c[0] for c in VARIANT_CASES])
def test_conv_variants_bit_identical(cuda, case, monkeypatch):
    """Operand-delivery variants of one layer.  The two-ring engine accumulates a k-block's three products in the order
    (lo*hi, hi*hi, hi*lo) per (dy, channel chunk, dx) whether it runs on CTA pairs or single CTAs, with or without the
    row-share halo: default == MSCNN_NO_2CTA == MSCNN_NO_ROWSHARE bit for bit.  MSCNN_NO_RING256 (BLOCK_N = 256) falls
    back to the term-major loop (all of hi*hi, then hi*lo, then lo*hi): another summation order, held to fp64 only.
    Every variant must be within 1e-4 of the fp64 convolution."""
    from mscnn_b200 import ops
    name, n, cin, h, w, cout, k, pad = case
    g = torch.Generator(device="cpu").manual_seed(1707)
    x = torch.randn((n, cin, h, w), generator=g).to(cuda)
    wt = (torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(cuda)
    b = (torch.randn((cout,), generator=g) * 0.1).to(cuda)
    xp = ops.nchw_to_planes(x, True)
    wp = ops.pack_conv_weights(wt, b, True)
    ref = _ref_conv(x, wt, b, pad, True)
    rms = float(ref.pow(2).mean().sqrt())

    def run(**env):
        for kk, v in env.items():
            monkeypatch.setenv(kk, v)
        ops.reload_config()
        try:
            y = ops.conv_forward(xp, wp, pad, relu=True)
            got = ops.planes_to_nchw(y)
            torch.cuda.synchronize()
            _report(f"{name} {env}", got, ref, 1e-4, 1e-4 * rms)
            return y.hi.clone(), y.lo.clone()
        finally:
            for kk in env:
                monkeypatch.delenv(kk)
            ops.reload_config()

    base = run()
    for env in ({"MSCNN_NO_2CTA": "1"}, {"MSCNN_NO_ROWSHARE": "1"}, {"MSCNN_NO_2CTA": "1", "MSCNN_NO_ROWSHARE": "1"}):
        other = run(**env)
        same = torch.equal(base[0], other[0]) and torch.equal(base[1], other[1])
        if not same:
            d = (base[0] != other[0]) | (base[1] != other[1])
            pytest.fail(f"{name}: {env} differs from the default build in {int(d.sum())}/{d.numel()} elements, "
                        f"first at {d.nonzero()[:4].tolist()}")
    run(MSCNN_NO_RING256="1")


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("k,pad", [(5, 2), (7, 3), (1, 0)])
def test_conv_head_f32_nchw(cuda, k, pad, split):
    """Narrow proposal heads: Cout = 9 (LFCN_*), fp32 NCHW output."""
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7)
    n, cin, h, w, cout = 2, 128, 12, 40, 9
    x = torch.randn((n, cin, h, w), generator=g).to(cuda)
    wt = (torch.randn((cout, cin, k, k), generator=g) * 0.02).to(cuda)
    b = (torch.randn((cout,), generator=g) * 0.1).to(cuda)
    xp = ops.nchw_to_planes(x, split)
    wp = ops.pack_conv_weights(wt, b, split)
    got = ops.conv_forward(xp, wp, pad, relu=False, out_f32=True)
    torch.cuda.synchronize()
    if split:
        ref = _ref_conv(x, wt, b, pad, False)
        _report("head", got, ref, 1e-4, 1e-4 * float(ref.pow(2).mean().sqrt()))
    else:
        ref = _ref_conv(x.bfloat16().float(), wt.bfloat16().float(), b, pad, False)
        _report("head", got, ref, 1e-4, 1e-4 * float(ref.pow(2).mean().sqrt()))  # fp32 accumulation over K=6272


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
def test_inner_product(cuda, split):
    """InnerProduct (inner_product_layer.cpp:84-97): Y = X W^T + b with X the NCHW flattening of
    a [R,C,H,W] bottom; weights are re-ordered to the NHWC flattening at pack time."""
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11)
    r, c, h, w, nout = 300, 64, 5, 5, 128
    x = torch.randn((r, c, h, w), generator=g).to(cuda)
    wt = (torch.randn((nout, c * h * w), generator=g) * 0.02).to(cuda)
    b = (torch.randn((nout,), generator=g) * 0.1).to(cuda)
    xp = ops.nchw_to_planes(x, split)
    flat = ops.Planes(xp.hi.view(r, 1, 1, -1), None if xp.lo is None else xp.lo.view(r, 1, 1, -1),
                      h * w * 64)
    wp = ops.pack_fc_weights(wt, b, split, c, h, w)
    yp = ops.conv_forward(flat, wp, 0, relu=True)
    got = ops.planes_to_nchw(yp).view(r, nout)
    # second, narrow fc with fp32 output (cls_pred / bbox_pred shape class)
    wt2 = (torch.randn((25, nout), generator=g) * 0.05).to(cuda)
    b2 = (torch.randn((25,), generator=g) * 0.1).to(cuda)
    wp2 = ops.pack_fc_weights(wt2, b2, split, nout, 1, 1)
    got2 = ops.conv_forward(yp, wp2, 0, relu=False, out_f32=True).view(r, 25)
    torch.cuda.synchronize()
    xin, win = (x, wt) if split else (x.bfloat16().float(), wt.bfloat16().float())
    ref = torch.relu(xin.view(r, -1).double() @ win.double().t() + b.double())
    rms = float(ref.pow(2).mean().sqrt())
    if split:
        _report("fc6", got, ref, 1e-4, 1e-4 * rms)
        ref2 = ref @ wt2.double().t() + b2.double()
        _report("fc_narrow", got2, ref2, 2e-4, 2e-4 * float(ref2.pow(2).mean().sqrt()))
    else:
        _report("fc6", got, ref, 2.0 ** -8, 1e-5 * rms)
        ref2 = got.double() @ wt2.bfloat16().double().t() + b2.double()
        _report("fc_narrow", got2, ref2, 1e-5, 1e-5 * float(ref2.pow(2).mean().sqrt()))


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
def test_conv1_1_im2col(cuda, split):
    """conv1_1: 3 -> 64, 3x3 pad 1 as a 1x1 GEMM over the 27-tap patch planes."""
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    n, h, w = 2, 20, 36
    x = (torch.randint(0, 256, (n, 3, h, w), generator=g).float()
         - torch.tensor([104.0, 117.0, 123.0]).view(1, 3, 1, 1)).to(cuda).contiguous()
    wt = (torch.randn((64, 3, 3, 3), generator=g) * (2.0 / 27) ** 0.5).to(cuda)
    b = (torch.randn((64,), generator=g) * 0.1).to(cuda)
    patches = ops.im2col3x3_c3(x, split)
    wp = ops.pack_conv_weights(wt.view(64, 27, 1, 1).contiguous(), b, split)
    yp = ops.conv_forward(patches, wp, 0, relu=True)
    got = ops.planes_to_nchw(yp)
    torch.cuda.synchronize()
    if split:
        ref = _ref_conv(x, wt, b, 1, True)
        _report("conv1_1", got, ref, 1e-4, 1e-4 * float(ref.pow(2).mean().sqrt()))
    else:
        ref = _ref_conv(x.bfloat16().float(), wt.bfloat16().float(), b, 1, True)
        _report("conv1_1", got, ref, 2.0 ** -8, 1e-5 * float(ref.pow(2).mean().sqrt()))


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
def test_conv1_1_direct(cuda, split):
    """conv1_1 as the direct exact-fp32 kernel (the path the Net uses)."""
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(4)
    n, h, w = 2, 21, 150        # W crosses a 128-pixel tile boundary
    x = (torch.randint(0, 256, (n, 3, h, w), generator=g).float()
         - torch.tensor([104.0, 117.0, 123.0]).view(1, 3, 1, 1)).to(cuda).contiguous()
    wt = (torch.randn((64, 3, 3, 3), generator=g) * (2.0 / 27) ** 0.5 / 64).to(cuda)
    b = (torch.randn((64,), generator=g) * 0.1).to(cuda)
    yp = ops.conv3x3_c3_forward(x, wt, b, relu=True, split=split)
    got = ops.planes_to_nchw(yp)
    torch.cuda.synchronize()
    ref = _ref_conv(x, wt, b, 1, True)
    rms = float(ref.pow(2).mean().sqrt())
    if split:
        _report("conv1_1_direct", got, ref, 2e-5, 2e-5 * rms)      # fp32 FMA + 2^-18 split of the output
    else:
        _report("conv1_1_direct", got, ref, 2.0 ** -8, 1e-5 * rms)


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("shape", [(2, 21, 150), (1, 5, 40), (3, 40, 384), (1, 3, 129), (2, 160, 640)])
def test_conv1_1_tensor_core(cuda, split, shape):
    """conv1_1 as the single tensor-core kernel (descriptor-shifted pixel rows; the path the Net uses).
    Widths below / across / at multiples of the 128-pixel tile, more tiles than SMs in the last case."""
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(4)
    n, h, w = shape
    x = (torch.randint(0, 256, (n, 3, h, w), generator=g).float()
         - torch.tensor([104.0, 117.0, 123.0]).view(1, 3, 1, 1)).to(cuda).contiguous()
    x[0, :, h // 2, w // 3] += 0.37          # not bf16-representable: exercises the lo plane of the input
    wt = (torch.randn((64, 3, 3, 3), generator=g) * (2.0 / 27) ** 0.5 / 64).to(cuda)
    b = (torch.randn((64,), generator=g) * 0.1).to(cuda)
    for relu in (True, False):
        yp = ops.conv1_tc_forward(x, wt, b, relu=relu, split=split)
        got = ops.planes_to_nchw(yp)
        torch.cuda.synchronize()
        if split:
            ref = _ref_conv(x, wt, b, 1, relu)
            rms = float(ref.pow(2).mean().sqrt())
            _report("conv1_1_tc", got, ref, 2e-5, 2e-5 * rms)
        else:
            ref = _ref_conv(x.bfloat16().float(), wt.bfloat16().float(), b, 1, relu)
            rms = float(ref.pow(2).mean().sqrt())
            _report("conv1_1_tc", got, ref, 2.0 ** -8, 1e-5 * rms)


@pytest.mark.parametrize("split", [True, False], ids=["split", "bf16"])
@pytest.mark.parametrize("shape", [(2, 64, 64, 192, 64), (1, 64, 128, 320, 128), (2, 128, 24, 80, 256), (3, 64, 10, 12, 64),
                                   # W % 128 == 0, narrow N: row-pair tiles with the pooling done in registers
                                   (1, 64, 8, 256, 64), (2, 128, 6, 128, 128), (2, 64, 64, 512, 64), (1, 64, 32, 384, 128)])
def test_conv_fused_pool(cuda, split, shape):
    """mscnn_conv_forward with pool_hi set == conv followed by mscnn_pool_forward, bit for bit; the
    pool-only form (no un-pooled store) gives the same pooled planes."""
    from mscnn_b200 import ops
    n, cin, h, w, cout = shape
    g = torch.Generator(device="cpu").manual_seed(21)
    x = torch.randn((n, cin, h, w), generator=g).to(cuda)
    wt = (torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5).to(cuda)
    b = (torch.randn((cout,), generator=g) * 0.1).to(cuda)
    xp = ops.nchw_to_planes(x, split)
    wp = ops.pack_conv_weights(wt, b, split)
    y_ref = ops.conv_forward(xp, wp, 1, relu=True)
    p_ref = ops.pool_forward(y_ref, 2, 2)
    y_both, p_both = ops.conv_forward(xp, wp, 1, relu=True, pool="both")
    p_only = ops.conv_forward(xp, wp, 1, relu=True, pool="only")
    torch.cuda.synchronize()
    assert torch.equal(y_both.hi, y_ref.hi)
    for cand in (p_both, p_only):
        assert torch.equal(cand.hi, p_ref.hi)
        if split:
            assert torch.equal(cand.lo, p_ref.lo)


def test_layout_roundtrip(cuda):
    from mscnn_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn((2, 70, 5, 37), generator=g).to(cuda)
    p = ops.nchw_to_planes(x, True)
    y = ops.planes_to_nchw(p)
    torch.cuda.synchronize()
    assert float((y - x).abs().max()) <= 2.0 ** -16 * float(x.abs().max())
    assert float(p.hi[..., 70:].float().abs().max()) == 0.0
    p1 = ops.nchw_to_planes(x, False)
    assert torch.equal(p1.hi[..., :70], x.permute(0, 2, 3, 1).bfloat16())
