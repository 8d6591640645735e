"""Host-side pin of the kernel variant every convolution of the BASELINE configurations takes (mscnn_conv_plan_describe:
the plan mscnn_conv_forward builds, without touching a device).  The full-size net reaches variants the small fixtures
never do (BLOCK_N = 256 two-ring engine with the row-share halo and CTA pairs on conv3_x, row pairs with register pooling
on conv1_2 / conv2_2); the GPU tests check their VALUES (tests/test_conv_gpu.py::test_conv_variants_bit_identical,
tests/test_fullsize_parity_gpu.py), this file checks that those are the variants the bench really runs and that
bench.py's roofline grouping names the same instantiations."""
import ctypes as C

import pytest


def _describe(n, cin, h, w, cout, k, pad, split=True, pooled=False, out_f32=False, dyn=False):
    from mscnn_b200 import capi
    L = capi.lib()
    d = capi.ConvDesc()
    fake = 0x10000                                   # never dereferenced: the plan query makes no CUDA call
    cpad = (cin + 63) // 64 * 64
    opad = (cout + 63) // 64 * 64
    d.x_hi, d.x_lo = fake, (fake if split else None)
    d.N, d.H, d.W, d.C = n, h, w, cpad
    d.w_hi, d.w_lo, d.bias = fake, (fake if split else None), fake
    d.Cout, d.Cout_pad, d.KH, d.KW, d.pad_h, d.pad_w = cout, opad, k, k, pad, pad
    d.relu = 1
    if out_f32:
        d.out_mode, d.y_f32 = capi.OUT_NCHW_F32, fake
    else:
        d.out_mode = capi.OUT_NHWC_BF16
        if pooled:
            d.pool_hi, d.pool_lo = fake, (fake if split else None)
        else:
            d.y_hi, d.y_lo = fake, (fake if split else None)
    if dyn:
        d.dyn_n = fake
    buf = C.create_string_buffer(512)
    rc = L.mscnn_conv_plan_describe(C.byref(d), buf, 512)
    assert rc == 0, rc
    import re
    return dict(kv.split("=", 1) for kv in re.split(r" (?=[a-z_]+=)", buf.value.decode()))


# mscnn-8s-768, batch 8, 3x768x2560 (BASELINE configs[2]); conv1_1 has its own kernel
FULL = {
    "conv1_2": (8, 64, 768, 2560, 64, True), "conv2_1": (8, 64, 384, 1280, 128, False), "conv2_2": (8, 128, 384, 1280, 128, True),
    "conv3_1": (8, 128, 192, 640, 256, False), "conv3_2": (8, 256, 192, 640, 256, False), "conv3_3": (8, 256, 192, 640, 256, True),
    "conv4_1": (8, 256, 96, 320, 512, False), "conv4_2": (8, 512, 96, 320, 512, False), "conv5_1": (8, 512, 48, 160, 512, False),
    "conv6_1": (8, 512, 24, 80, 512, False),
}


def test_full_size_trunk_variants():
    p = {k: _describe(n, cin, h, w, cout, 3, 1, pooled=pool) for k, (n, cin, h, w, cout, pool) in FULL.items()}
    # narrow N, pooled, W % 128 == 0: row pairs with the 2x2 max pooling in registers
    for name, bn in (("conv1_2", 64), ("conv2_2", 128)):
        assert p[name]["kernel"] == f"conv_igemm_kernel<{bn}, false>" and p[name]["vpool"] == "1" and p[name]["a_taps"] == "3"
        assert p[name]["wide"] == "1" and p[name]["box"] == "128x1x1" and p[name]["mt"] == "2"
    assert p["conv2_2"]["acc_sets"] == "1" and p["conv1_2"]["acc_sets"] == "2"
    # narrow N, not pooled: wide-B with the row-share halo
    assert p["conv2_1"]["kernel"] == "conv_igemm_kernel<128, false>" and p["conv2_1"]["a_taps"] == "3" and p["conv2_1"]["vpool"] == "0"
    # BLOCK_N = 256 on 128x1 boxes: two rings, halo, single-tile weight slots, CTA pairs -- the variant that carries the bench
    for name in ("conv3_1", "conv3_2"):
        assert p[name]["kernel"] == "conv_igemm_kernel<256, true>", p[name]
        assert p[name]["box"] == "128x1x1" and p[name]["a_taps"] == "3" and p[name]["b_split"] == "1" and p[name]["mt"] == "1"
    # the pooled conv3_3: even-sized 2-D boxes (a 2x2 window never straddles tiles), pooled in the epilogue per CTA, pairs too
    assert p["conv3_3"]["kernel"] == "conv_igemm_kernel<256, true>" and p["conv3_3"]["pool"] == "1" and p["conv3_3"]["b_split"] == "1"
    assert p["conv3_3"]["box"] == "64x2x1" and p["conv3_3"]["a_taps"] == "1"
    # 2-D boxes further down (W = 320 / 160 / 80): pairs without the halo
    for name in ("conv4_1", "conv4_2", "conv5_1", "conv6_1"):
        assert p[name]["kernel"] == "conv_igemm_kernel<256, true>" and p[name]["a_taps"] == "1" and p[name]["b_split"] == "1", p[name]
        assert p[name]["box"] != "128x1x1"


def test_head_variants_and_dynamic_rows():
    # roi_c1: 3x3 without padding on R x 1024 x 7 x 7 -> R x 512 x 5 x 5; an M tile = ONE output pixel of 128 consecutive
    # ROIs (128 of 128 MMA rows used; a 5x5x5 box would use 125); with the device-side row count MT stays 1
    r = _describe(16000, 1024, 7, 7, 512, 3, 0, dyn=True)
    assert r["kernel"] == "conv_igemm_kernel<256, true>" and r["box"] == "1x1x128" and r["dyn"] == "1" and r["mt"] == "1"
    fc6 = _describe(16000, 12800, 1, 1, 4096, 1, 0, dyn=True)
    assert fc6["kernel"] == "conv_igemm_kernel<256, true>" and fc6["box"] == "1x1x128"
    cls = _describe(16000, 4096, 1, 1, 5, 1, 0, out_f32=True, dyn=True)
    assert cls["kernel"] == "conv_igemm_kernel<64, false>" and cls["mt"] == "1" and cls["dyn"] == "1"
    # plain bf16: no CTA pairs, same rings
    b = _describe(8, 256, 192, 640, 256, 3, 1, split=False)
    assert b["kernel"] == "conv_igemm_kernel<256, false>" and b["terms"] == "1" and b["a_taps"] == "3"


def test_switches_change_the_plan(monkeypatch):
    from mscnn_b200 import ops
    try:
        monkeypatch.setenv("MSCNN_NO_2CTA", "1")
        ops.reload_config()
        assert _describe(8, 256, 192, 640, 256, 3, 1)["kernel"] == "conv_igemm_kernel<256, false>"
        monkeypatch.setenv("MSCNN_NO_RING256", "1")
        ops.reload_config()
        assert _describe(8, 256, 192, 640, 256, 3, 1)["rings"].startswith("0")
    finally:
        monkeypatch.delenv("MSCNN_NO_2CTA", raising=False)
        monkeypatch.delenv("MSCNN_NO_RING256", raising=False)
        ops.reload_config()
    assert _describe(8, 256, 192, 640, 256, 3, 1)["kernel"] == "conv_igemm_kernel<256, true>"


def test_bench_roofline_grouping_names_the_same_kernels():
    """bench.py's `roofline.by_kernel` derives the instantiation of every layer from its shape; it must agree with the
    library's own plan for the layers of the bench net."""
    import bench
    for name, (n, cin, h, w, cout, pool) in FULL.items():
        lib = _describe(n, cin, h, w, cout, 3, 1, pooled=pool)["kernel"]
        cpad = (cout + 63) // 64 * 64
        bn = 256 if cpad % 256 == 0 else 128 if cpad % 128 == 0 else 64
        pair = bn == 256
        assert lib == f"conv_igemm_kernel<{bn}, {'true' if pair else 'false'}>", (name, lib)
