"""Thin torch-tensor conveniences over the C ABI (used by tests, bench.py and the Python
mirror of the Caffe Net).  torch is used for device memory and streams only; every function
here ends in exactly one C-ABI call and raises if the native library is missing.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import capi


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def reload_config() -> None:
    """Re-read the MSCNN_* environment switches (they are read once per process / per Net otherwise)."""
    capi.lib().mscnn_config_reload()


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


@dataclass
class Planes:
    """NHWC bf16 activation planes [N][H][W][Cpad]; ``lo`` is None on the plain-bf16 path."""
    hi: torch.Tensor
    lo: torch.Tensor | None
    channels: int  # logical channel count (<= Cpad)

    @property
    def shape(self):
        return tuple(self.hi.shape)

    def float(self) -> torch.Tensor:
        v = self.hi.float()
        if self.lo is not None:
            v = v + self.lo.float()
        return v[..., : self.channels]


@dataclass
class PackedWeights:
    hi: torch.Tensor          # [Cout_pad][KH][KW][Cin_pad] bf16
    lo: torch.Tensor | None
    bias: torch.Tensor        # fp32 [Cout_pad]
    cout: int
    kh: int
    kw: int


def nchw_to_planes(x: torch.Tensor, split: bool) -> Planes:
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, c, h, w = x.shape
    cp = pad64(c)
    hi = torch.empty((n, h, w, cp), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi) if split else None
    capi.check(capi.lib().mscnn_nchw_f32_to_planes(capi.ptr(x), capi.ptr(hi), capi.ptr(lo),
                                                   n, c, h, w, cp, _stream()), "nchw_f32_to_planes")
    return Planes(hi, lo, c)


def planes_to_nchw(p: Planes) -> torch.Tensor:
    n, h, w, cp = p.hi.shape
    y = torch.empty((n, p.channels, h, w), dtype=torch.float32, device=p.hi.device)
    capi.check(capi.lib().mscnn_planes_to_nchw_f32(capi.ptr(p.hi), capi.ptr(p.lo), capi.ptr(y),
                                                   n, p.channels, h, w, cp, _stream()),
               "planes_to_nchw_f32")
    return y


def im2col3x3_c3(x: torch.Tensor, split: bool) -> Planes:
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    n, _, h, w = x.shape
    hi = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi) if split else None
    capi.check(capi.lib().mscnn_im2col3x3_c3_to_planes(capi.ptr(x), capi.ptr(hi), capi.ptr(lo),
                                                       n, h, w, _stream()), "im2col3x3_c3")
    return Planes(hi, lo, 27)


def pack_conv_weights(w: torch.Tensor, b: torch.Tensor | None, split: bool,
                      cout_pad: int | None = None, cin_pad: int | None = None) -> PackedWeights:
    """w: [Cout][Cin][KH][KW] fp32 (Caffe blob 0), b: [Cout] fp32 (blob 1) or None."""
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 4
    cout, cin, kh, kw = w.shape
    cout_pad = cout_pad or (pad64(cout) if cout > 32 else 32)
    cin_pad = cin_pad or pad64(cin)
    hi = torch.empty((cout_pad, kh, kw, cin_pad), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi) if split else None
    capi.check(capi.lib().mscnn_pack_conv_weights(capi.ptr(w), capi.ptr(hi), capi.ptr(lo), cout, cin,
                                                  kh, kw, cout_pad, cin_pad, _stream()),
               "pack_conv_weights")
    bias = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    if b is not None:
        bias[:cout] = b
    return PackedWeights(hi, lo, bias, cout, kh, kw)


def pack_fc_weights(w: torch.Tensor, b: torch.Tensor | None, split: bool, c: int, h: int, wd: int,
                    nout_pad: int | None = None) -> PackedWeights:
    """w: [Nout][c*h*wd] fp32 in Caffe's NCHW flattening of the bottom blob."""
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 2
    nout = w.shape[0]
    assert w.shape[1] == c * h * wd
    nout_pad = nout_pad or (pad64(nout) if nout > 32 else 32)
    cp = pad64(c)
    hi = torch.empty((nout_pad, 1, 1, h * wd * cp), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi) if split else None
    capi.check(capi.lib().mscnn_pack_fc_weights(capi.ptr(w), capi.ptr(hi), capi.ptr(lo), nout, c, h,
                                                wd, nout_pad, cp, _stream()), "pack_fc_weights")
    bias = torch.zeros(nout_pad, dtype=torch.float32, device=w.device)
    if b is not None:
        bias[:nout] = b
    return PackedWeights(hi, lo, bias, nout, 1, 1)


def conv_forward(x: Planes, w: PackedWeights, pad: int, relu: bool, out_f32: bool = False,
                 out_split: bool | None = None, pool: str | None = None):
    """Convolution (stride 1) + bias (+ReLU).  Returns Planes, or an NCHW fp32 tensor when
    ``out_f32``.  An InnerProduct is the KH=KW=H=W=1 case."""
    n, h, wd, c = x.hi.shape
    assert w.hi.shape[3] == c, (w.hi.shape, x.hi.shape)
    ho, wo = h + 2 * pad - w.kh + 1, wd + 2 * pad - w.kw + 1
    cout_pad = w.hi.shape[0]
    d = capi.ConvDesc()
    d.x_hi, d.x_lo = capi.ptr(x.hi), capi.ptr(x.lo)
    d.N, d.H, d.W, d.C = n, h, wd, c
    d.w_hi, d.w_lo, d.bias = capi.ptr(w.hi), capi.ptr(w.lo), capi.ptr(w.bias)
    d.Cout, d.Cout_pad, d.KH, d.KW, d.pad_h, d.pad_w = w.cout, cout_pad, w.kh, w.kw, pad, pad
    d.relu = int(relu)
    if out_f32:
        y = torch.empty((n, w.cout, ho, wo), dtype=torch.float32, device=x.hi.device)
        d.out_mode, d.y_f32 = capi.OUT_NCHW_F32, capi.ptr(y)
        out = y
    else:
        if out_split is None:
            out_split = x.lo is not None
        y_hi = torch.empty((n, ho, wo, cout_pad), dtype=torch.bfloat16, device=x.hi.device)
        y_lo = torch.empty_like(y_hi) if out_split else None
        d.out_mode, d.y_hi, d.y_lo = capi.OUT_NHWC_BF16, capi.ptr(y_hi), capi.ptr(y_lo)
        out = Planes(y_hi, y_lo, w.cout)
        if pool:   # "both": conv output and its fused 2x2 max pool; "only": pooled tensor alone
            p_hi = torch.empty((n, ho // 2, wo // 2, cout_pad), dtype=torch.bfloat16, device=x.hi.device)
            p_lo = torch.empty_like(p_hi) if out_split else None
            d.pool_hi, d.pool_lo = capi.ptr(p_hi), capi.ptr(p_lo)
            pooled = Planes(p_hi, p_lo, w.cout)
            if pool == "only":
                d.y_hi, d.y_lo = None, None
                out = pooled
            else:
                out = (out, pooled)
    capi.check(capi.lib().mscnn_conv_forward(d, _stream()), "conv_forward")
    return out


def pool_forward(x: Planes, kernel: int, stride: int, mode: int = capi.POOL_MAX) -> Planes:
    n, h, w, c = x.hi.shape
    ho = -(-(h - kernel) // stride) + 1
    wo = -(-(w - kernel) // stride) + 1
    y_hi = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.hi.device)
    y_lo = torch.empty_like(y_hi) if x.lo is not None else None
    capi.check(capi.lib().mscnn_pool_forward(capi.ptr(x.hi), capi.ptr(x.lo), capi.ptr(y_hi), capi.ptr(y_lo),
                                             n, h, w, c, kernel, stride, mode, _stream()), "pool_forward")
    return Planes(y_hi, y_lo, x.channels)


def deconv2x_forward(x: Planes, w: torch.Tensor) -> Planes:
    """w: fp32 [C][1][4][4] (group == channels, stride 2, pad 1, no bias)."""
    n, h, wd, c = x.hi.shape
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape[1:]) == (1, 4, 4)
    y_hi = torch.empty((n, 2 * h, 2 * wd, c), dtype=torch.bfloat16, device=x.hi.device)
    y_lo = torch.empty_like(y_hi) if x.lo is not None else None
    capi.check(capi.lib().mscnn_deconv2x_forward(capi.ptr(x.hi), capi.ptr(x.lo), capi.ptr(w), capi.ptr(y_hi),
                                                 capi.ptr(y_lo), n, h, wd, c, w.shape[0], _stream()),
               "deconv2x_forward")
    return Planes(y_hi, y_lo, x.channels)


def make_box_cfg(maps_shapes, field_w, field_h, rate, fg_thr, iou_thr, nms_type="IOU", field_whr=2.0,
                 field_xyr=2.0, min_size=15.0, max_nms_num=2000, max_post_nms_num=0, bbox_mean=None,
                 bbox_std=None) -> capi.BoxOutputCfg:
    """maps_shapes: [(C, H, W)] per bottom, in bottom order (BoxOutputParameter, caffe.proto:1315-1329)."""
    cfg = capi.BoxOutputCfg()
    cfg.num_scales = len(maps_shapes)
    cfg.channels = maps_shapes[0][0]
    for j, (c, h, w) in enumerate(maps_shapes):
        assert c == cfg.channels
        cfg.height[j], cfg.width[j] = h, w
        cfg.field_w[j], cfg.field_h[j], cfg.downsample_rate[j] = field_w[j], field_h[j], rate[j]
    cfg.fg_thr, cfg.iou_thr = fg_thr, iou_thr
    cfg.nms_type = {"IOU": capi.NMS_IOU, "IOMU": capi.NMS_IOMU, "IOFU": capi.NMS_IOFU}.get(nms_type, capi.NMS_IOU)
    cfg.field_whr, cfg.field_xyr, cfg.min_size = field_whr, field_xyr, min_size
    cfg.max_nms_num, cfg.max_post_nms_num = max_nms_num, max_post_nms_num
    cfg.do_bbox_norm = int(bool(bbox_mean) and bool(bbox_std))
    for k in range(4):
        cfg.bbox_mean[k] = bbox_mean[k] if cfg.do_bbox_norm else 0.0
        cfg.bbox_std[k] = bbox_std[k] if cfg.do_bbox_norm else 1.0
    return cfg


def box_output_forward(cfg: capi.BoxOutputCfg, maps: list[torch.Tensor]):
    """maps: fp32 NCHW score/delta maps.  Returns (proposals[cap,5], proposals_score[cap,6],
    num_out int32[2+N]) device tensors; rows beyond num_out[0] are undefined."""
    import ctypes as C
    n = maps[0].shape[0]
    dev = maps[0].device
    for m in maps:
        assert m.is_cuda and m.dtype == torch.float32 and m.is_contiguous() and m.shape[0] == n
    nbytes = C.c_size_t(0)
    capi.check(capi.lib().mscnn_box_output_workspace_bytes(cfg, n, C.byref(nbytes)), "box_output_workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    cap = max(n * cfg.max_nms_num, 1)
    rois = torch.empty((cap, 5), dtype=torch.float32, device=dev)
    rois_score = torch.empty((cap, 6), dtype=torch.float32, device=dev)
    num_out = torch.zeros(2 + n, dtype=torch.int32, device=dev)
    ptrs = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
    capi.check(capi.lib().mscnn_box_output_forward(cfg, n, ptrs, capi.ptr(ws), nbytes.value, capi.ptr(rois),
                                                   capi.ptr(rois_score), capi.ptr(num_out), _stream()),
               "box_output_forward")
    return rois, rois_score, num_out


def roi_pool_forward(x: Planes, rois: torch.Tensor, num_rois: int, pooled: int, scale: float,
                     pad_ratio: float, out: Planes | None = None, channel_offset: int = 0,
                     out_channels: int | None = None) -> Planes:
    n, h, w, c = x.hi.shape
    ctot = out_channels or c
    if out is None:
        y_hi = torch.zeros((num_rois, pooled, pooled, ctot), dtype=torch.bfloat16, device=x.hi.device)
        y_lo = torch.zeros_like(y_hi) if x.lo is not None else None
        out = Planes(y_hi, y_lo, ctot)
    capi.check(capi.lib().mscnn_roi_pool_forward(capi.ptr(x.hi), capi.ptr(x.lo), n, h, w, c, capi.ptr(rois),
                                                 num_rois, pooled, pooled, scale, pad_ratio, capi.ptr(out.hi),
                                                 capi.ptr(out.lo), ctot, channel_offset, _stream()),
               "roi_pool_forward")
    return out


def roi_align_forward(x: Planes, rois: torch.Tensor, num_rois: int, pooled: int, scale: float,
                      pad_ratio: float) -> Planes:
    """ROIAlign grid [R, pooled+1, pooled+1, C] (roi_align_layer.cpp:49-139)."""
    n, h, w, c = x.hi.shape
    y_hi = torch.zeros((num_rois, pooled + 1, pooled + 1, c), dtype=torch.bfloat16, device=x.hi.device)
    y_lo = torch.zeros_like(y_hi) if x.lo is not None else None
    out = Planes(y_hi, y_lo, c)
    capi.check(capi.lib().mscnn_roi_align_forward(capi.ptr(x.hi), capi.ptr(x.lo), n, h, w, c, capi.ptr(rois),
                                                  num_rois, pooled, pooled, scale, pad_ratio, capi.ptr(out.hi),
                                                  capi.ptr(out.lo), c, 0, _stream()), "roi_align_forward")
    return out


def decode_bbox_forward(bbox_pred: torch.Tensor, prior: torch.Tensor, mean=None, std=None) -> torch.Tensor:
    import ctypes as C
    r = prior.shape[0]
    out = torch.empty((r, 5), dtype=torch.float32, device=prior.device)
    m = (C.c_float * 4)(*mean) if mean is not None else None
    s = (C.c_float * 4)(*std) if std is not None else None
    capi.check(capi.lib().mscnn_decode_bbox_forward(capi.ptr(bbox_pred), capi.ptr(prior), r, bbox_pred.shape[1],
                                                    C.cast(m, C.c_void_p), C.cast(s, C.c_void_p), capi.ptr(out),
                                                    _stream()), "decode_bbox_forward")
    return out


def softmax_forward(x: torch.Tensor, axis: int = 1) -> torch.Tensor:
    outer = int(np.prod(x.shape[:axis])) if axis > 0 else 1
    inner = int(np.prod(x.shape[axis + 1:])) if axis + 1 < x.dim() else 1
    y = torch.empty_like(x)
    capi.check(capi.lib().mscnn_softmax_forward(capi.ptr(x), outer, x.shape[axis], inner, capi.ptr(y), _stream()),
               "softmax_forward")
    return y


def eltwise_forward(bottoms: list[torch.Tensor], op: int, coeffs=None) -> torch.Tensor:
    import ctypes as C
    y = torch.empty_like(bottoms[0])
    ptrs = (C.c_void_p * len(bottoms))(*[b.data_ptr() for b in bottoms])
    cf = (C.c_float * len(bottoms))(*coeffs) if coeffs is not None else None
    capi.check(capi.lib().mscnn_eltwise_forward(ptrs, len(bottoms), op, C.cast(cf, C.c_void_p), y.numel(),
                                                capi.ptr(y), _stream()), "eltwise_forward")
    return y


def cascade_detect_postprocess(cfg: capi.DetectCfg, n: int, proposals, cls_prob, output_bbox, num_out):
    import ctypes as C
    dev = proposals.device
    nbytes = C.c_size_t(0)
    capi.check(capi.lib().mscnn_detect_workspace_bytes(cfg, n, C.byref(nbytes)), "detect_workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    dets = torch.zeros((n, cfg.max_rois_per_image, 5), dtype=torch.float32, device=dev)
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    capi.check(capi.lib().mscnn_cascade_detect_postprocess(cfg, n, capi.ptr(proposals), capi.ptr(cls_prob),
                                                           capi.ptr(output_bbox), capi.ptr(num_out), capi.ptr(ws),
                                                           nbytes.value, capi.ptr(dets), capi.ptr(counts), _stream()),
               "cascade_detect_postprocess")
    return dets, counts


def detect_postprocess(cfg: capi.DetectCfg, n: int, proposals_score, cls_pred, bbox_pred, num_out):
    import ctypes as C
    dev = proposals_score.device
    nbytes = C.c_size_t(0)
    capi.check(capi.lib().mscnn_detect_workspace_bytes(cfg, n, C.byref(nbytes)), "detect_workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    dets = torch.zeros((n, cfg.max_rois_per_image, 5), dtype=torch.float32, device=dev)
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    capi.check(capi.lib().mscnn_detect_postprocess(cfg, n, capi.ptr(proposals_score), capi.ptr(cls_pred),
                                                   capi.ptr(bbox_pred), capi.ptr(num_out), capi.ptr(ws),
                                                   nbytes.value, capi.ptr(dets), capi.ptr(counts), _stream()),
               "detect_postprocess")
    return dets, counts


def conv3x3_c3_forward(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, relu: bool, split: bool) -> Planes:
    """conv1_1 shape: x [N,3,H,W] fp32 NCHW, w [Cout,3,3,3] fp32 -> planes (direct fp32 kernel)."""
    n, _, h, wd = x.shape
    cout = w.shape[0]
    cp = pad64(cout)
    y_hi = torch.empty((n, h, wd, cp), dtype=torch.bfloat16, device=x.device)
    y_lo = torch.empty_like(y_hi) if split else None
    capi.check(capi.lib().mscnn_conv3x3_c3_forward(capi.ptr(x), capi.ptr(w), capi.ptr(b), capi.ptr(y_hi),
                                                   capi.ptr(y_lo), n, h, wd, cout, cp, int(relu), _stream()),
               "conv3x3_c3_forward")
    return Planes(y_hi, y_lo, cout)


def conv1_tc_forward(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, relu: bool, split: bool) -> Planes:
    """conv1_1 (3 -> 64, 3x3, pad 1) as the single tensor-core kernel: x [N,3,H,W] fp32 NCHW -> planes."""
    n, c, h, wd = x.shape
    assert c == 3 and tuple(w.shape) == (64, 3, 3, 3) and x.is_contiguous() and w.is_contiguous()
    L = capi.lib()
    packed = torch.empty(L.mscnn_conv1_tc_packed_bytes(), dtype=torch.uint8, device=x.device)
    capi.check(L.mscnn_pack_conv1_tc_weights(capi.ptr(w), capi.ptr(packed), int(split), _stream()), "pack_conv1_tc")
    bias = b if b is not None else torch.zeros(64, device=x.device)
    y_hi = torch.empty((n, h, wd, 64), dtype=torch.bfloat16, device=x.device)
    y_lo = torch.empty_like(y_hi) if split else None
    capi.check(L.mscnn_conv1_tc_forward(capi.ptr(x), capi.ptr(packed), capi.ptr(bias), capi.ptr(y_hi), capi.ptr(y_lo),
                                        n, h, wd, int(relu), _stream()), "conv1_tc_forward")
    return Planes(y_hi, y_lo, 64)


# ---- pre-processing ------------------------------------------------------------
def imresize_contributions(in_len: int, out_len: int) -> tuple[np.ndarray, np.ndarray]:
    """The library's imresize tap tables for one dimension (host-only call, works without a GPU)."""
    L = capi.lib()
    taps = L.mscnn_imresize_taps(in_len, out_len)
    capi.check(min(taps, 0), "imresize_taps")
    w = np.zeros((out_len, taps), dtype=np.float64)
    idx = np.zeros((out_len, taps), dtype=np.int32)
    rc = L.mscnn_imresize_contributions(in_len, out_len, w.ctypes.data, idx.ctypes.data, taps)
    capi.check(min(rc, 0), "imresize_contributions")
    return w, idx


def widerface_net_size(org_h: int, org_w: int, img_h: int = 0, img_w: int = 0, max_size: int = 2048) -> tuple[int, int]:
    import ctypes as C
    h, w = C.c_int(), C.c_int()
    capi.check(capi.lib().mscnn_widerface_net_size(org_h, org_w, img_h, img_w, max_size, C.byref(h), C.byref(w)),
               "widerface_net_size")
    return h.value, w.value


class Preprocess:
    """imresize + BGR + mean subtraction + CHW on the device for images of one size
    (examples/kitti_car/run_mscnn_detection.m:64-69)."""

    def __init__(self, in_hw, out_hw, mean=(104.0, 117.0, 123.0), swap_rb=True):
        import ctypes as C
        d = capi.PreprocessDesc(in_hw[0], in_hw[1], out_hw[0], out_hw[1], (C.c_float * 3)(*mean), int(swap_rb))
        h = C.c_void_p()
        capi.check(capi.lib().mscnn_preprocess_create(C.byref(d), C.byref(h)), "preprocess_create")
        self.handle, self.in_hw, self.out_hw = h, tuple(in_hw), tuple(out_hw)

    def __del__(self):
        if getattr(self, "handle", None):
            capi.lib().mscnn_preprocess_destroy(self.handle)
            self.handle = None

    def __call__(self, images, out: torch.Tensor | None = None) -> torch.Tensor:
        """images: uint8 [N][h][w][3], a torch.cuda tensor, or a host numpy array / CPU tensor (H2D inside)."""
        n = images.shape[0]
        assert tuple(images.shape[1:]) == (*self.in_hw, 3) and images.dtype in (torch.uint8, np.uint8)
        if out is None:
            out = torch.empty((n, 3, *self.out_hw), dtype=torch.float32, device="cuda")
        L = capi.lib()
        if isinstance(images, torch.Tensor) and images.is_cuda:
            assert images.is_contiguous()
            capi.check(L.mscnn_preprocess_forward(self.handle, n, capi.ptr(images), capi.ptr(out), _stream()),
                       "preprocess_forward")
        else:
            capi.check(L.mscnn_preprocess_forward_host(self.handle, n, capi.ptr(images), capi.ptr(out), _stream()),
                       "preprocess_forward_host")
        return out
