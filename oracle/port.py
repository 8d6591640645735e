"""numpy / C restatement ("port") of the MS-CNN forward-path layers -- TEST INFRASTRUCTURE
(see oracle/__init__.py).  Each function cites the reference file:line it follows; the loops
that numpy cannot express efficiently live in oracle/mscnn_oracle.c.

Pinned by tests/test_oracle.py against (a) the upstream known-answer tests the reference ships
for conv / pooling / inner-product / deconv (golden matrices restated from
src/caffe/test/test_pooling_layer.cpp etc.), (b) oracle/_ref -- the reference's own code
compiled verbatim -- on random and edge-case inputs, and (c) golden vectors generated from
oracle/_ref (tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SRC = _DIR / "mscnn_oracle.c"
_LIB_PATH = _DIR / "libmscnn_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < _SRC.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(_LIB_PATH),
                        str(_SRC), "-lm"], check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.oracle_box_iou.restype = C.c_float
        L.oracle_box_iou.argtypes = [C.c_float] * 8 + [C.c_int]
        L.oracle_box_output.restype = C.c_int
        L.oracle_bbnms_maxg.restype = C.c_int
        L.oracle_bbnms_maxg.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ---------------------------------------------------------------------------------------------
def conv2d(x, w, b=None, pad=0, stride=1, group=1) -> np.ndarray:
    """ConvolutionLayer::Forward_cpu = per-image im2col + sgemm + bias
    (src/caffe/layers/conv_layer.cpp:25-40, base_conv_layer.cpp:257-280, util/im2col.cpp:19-55).
    x [N,C,H,W], w [Cout,C/group,kh,kw] fp32 -> [N,Cout,Ho,Wo] (output size conv_layer.cpp:8-22)."""
    x, w = _f32(x), _f32(w)
    n, c, h, wd = x.shape
    cout, cg, kh, kw = w.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (wd + 2 * pad - kw) // stride + 1
    xp = np.zeros((n, c, h + 2 * pad, wd + 2 * pad), dtype=np.float32)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    y = np.empty((n, cout, ho, wo), dtype=np.float32)
    og = cout // group
    for i in range(n):
        for g in range(group):
            col = np.empty((cg, kh, kw, ho, wo), dtype=np.float32)
            for dy in range(kh):
                for dx in range(kw):
                    col[:, dy, dx] = xp[i, g * cg:(g + 1) * cg, dy:dy + stride * ho:stride, dx:dx + stride * wo:stride]
            y[i, g * og:(g + 1) * og] = (w[g * og:(g + 1) * og].reshape(og, -1) @ col.reshape(cg * kh * kw, -1)
                                          ).reshape(og, ho, wo)
    if b is not None:
        y += _f32(b).reshape(1, -1, 1, 1)
    return y


def relu(x) -> np.ndarray:
    """ReLULayer::Forward_cpu with negative_slope 0 (src/caffe/layers/relu_layer.cpp:9-19)."""
    return np.maximum(_f32(x), np.float32(0))


def inner_product(x, w, b=None) -> np.ndarray:
    """InnerProductLayer::Forward_cpu, Y = X W^T + 1 b^T (inner_product_layer.cpp:84-97)."""
    x = _f32(x)
    y = x.reshape(x.shape[0], -1) @ _f32(w).T
    if b is not None:
        y = y + _f32(b)[None, :]
    return y.astype(np.float32)


def concat_channels(*xs) -> np.ndarray:
    """ConcatLayer::Forward_cpu on axis 1 (src/caffe/layers/concat_layer.cpp:57-74)."""
    return np.concatenate([_f32(x) for x in xs], axis=1)


def pool(x, kernel, stride, pad=0, mode="MAX") -> np.ndarray:
    """PoolingLayer::Forward_cpu (pooling_layer.cpp:79-123 shape, :128-220 MAX / AVE)."""
    x = _f32(x)
    n, c, h, w = x.shape
    oh, ow = C.c_int(0), C.c_int(0)
    m = 0 if mode == "MAX" else 1
    lib().oracle_pool(C.c_void_p(x.ctypes.data), n, c, h, w, kernel, stride, pad, m, None, C.byref(oh), C.byref(ow))
    y = np.empty((n, c, oh.value, ow.value), dtype=np.float32)
    lib().oracle_pool(C.c_void_p(x.ctypes.data), n, c, h, w, kernel, stride, pad, m, C.c_void_p(y.ctypes.data),
                      C.byref(oh), C.byref(ow))
    return y


def deconv_depthwise(x, w, stride=2, pad=1) -> np.ndarray:
    """DeconvolutionLayer::Forward_cpu with group == channels, no bias (deconv_layer.cpp:8-40)."""
    x, w = _f32(x), _f32(w)
    n, c, h, wd = x.shape
    k = w.shape[-1]
    ho, wo = stride * (h - 1) + k - 2 * pad, stride * (wd - 1) + k - 2 * pad
    y = np.empty((n, c, ho, wo), dtype=np.float32)
    lib().oracle_deconv_depthwise(C.c_void_p(x.ctypes.data), n, c, h, wd, C.c_void_p(w.ctypes.data), k, stride,
                                  pad, C.c_void_p(y.ctypes.data))
    return y


def box_iou(b1, b2, mode="IOU") -> float:
    """BoxIOU on (x, y, w, h) boxes (util/math_functions.cpp:13-35)."""
    m = {"IOU": 0, "IOMU": 1, "IOFU": 2}.get(mode, 0)
    return float(lib().oracle_box_iou(*[float(np.float32(v)) for v in (*b1, *b2)], m))


def box_output(maps, field_w, field_h, downsample_rate, fg_thr=0.0, iou_thr=0.5, nms_type="IOU",
               field_whr=2.0, field_xyr=2.0, min_size=15.0, max_nms_num=0, max_post_nms_num=0,
               bbox_mean=None, bbox_std=None):
    """BoxOutputLayer::Forward_cpu (box_output_layer.cpp:66-234).  maps: list of [N,C,Hj,Wj].
    Returns (proposals [R,5], proposals_score [R,6], per_image_counts [N], true_count)."""
    maps = [_f32(m) for m in maps]
    n, ch = maps[0].shape[:2]
    j = len(maps)
    total = sum(m.shape[2] * m.shape[3] for m in maps)
    cap = max(1, n * (min(max_nms_num, total) if max_nms_num > 0 else total))
    rois = np.zeros((cap, 5), dtype=np.float32)
    rois_score = np.zeros((cap, 6), dtype=np.float32)
    per_image = np.zeros(n, dtype=np.int32)
    true_count = C.c_int(0)
    ptrs = (C.c_void_p * j)(*[m.ctypes.data for m in maps])
    hs = (C.c_int * j)(*[m.shape[2] for m in maps])
    ws = (C.c_int * j)(*[m.shape[3] for m in maps])
    fw = (C.c_uint * j)(*[int(v) for v in field_w])
    fh = (C.c_uint * j)(*[int(v) for v in field_h])
    dr = (C.c_uint * j)(*[int(v) for v in downsample_rate])
    do_norm = int(bool(bbox_mean) and bool(bbox_std))
    mean = (C.c_float * 4)(*(bbox_mean if do_norm else [0, 0, 0, 0]))
    std = (C.c_float * 4)(*(bbox_std if do_norm else [1, 1, 1, 1]))
    mode = {"IOU": 0, "IOMU": 1, "IOFU": 2}.get(nms_type, 0)
    rows = lib().oracle_box_output(n, ch, j, ptrs, hs, ws, fw, fh, dr, C.c_float(fg_thr), C.c_float(iou_thr), mode,
                                   C.c_float(field_whr), C.c_float(field_xyr), C.c_float(min_size),
                                   int(max_nms_num), int(max_post_nms_num), do_norm, mean, std,
                                   C.c_void_p(rois.ctypes.data), C.c_void_p(rois_score.ctypes.data), cap,
                                   C.c_void_p(per_image.ctypes.data), C.byref(true_count))
    return rois[:rows].copy(), rois_score[:rows].copy(), per_image, true_count.value


def roi_pool(x, rois, pooled_h, pooled_w, spatial_scale, pad_ratio=0.0) -> np.ndarray:
    """ROIPoolingLayer::Forward_cpu incl. pad_ratio (roi_pooling_layer.cpp:49-139)."""
    x, rois = _f32(x), _f32(rois).reshape(-1, 5)
    n, c, h, w = x.shape
    r = rois.shape[0]
    y = np.empty((r, c, pooled_h, pooled_w), dtype=np.float32)
    lib().oracle_roi_pool(C.c_void_p(x.ctypes.data), n, c, h, w, C.c_void_p(rois.ctypes.data), r, pooled_h,
                          pooled_w, C.c_float(spatial_scale), C.c_float(pad_ratio), C.c_void_p(y.ctypes.data))
    return y


def roi_align(x, rois, pooled_h, pooled_w, spatial_scale, pad_ratio=0.0) -> np.ndarray:
    """ROIAlignLayer::Forward_cpu (roi_align_layer.cpp:49-139) -> [R, C, pooled_h+1, pooled_w+1]."""
    x, rois = _f32(x), _f32(rois).reshape(-1, 5)
    n, c, h, w = x.shape
    r = rois.shape[0]
    y = np.empty((r, c, pooled_h + 1, pooled_w + 1), dtype=np.float32)
    lib().oracle_roi_align(C.c_void_p(x.ctypes.data), n, c, h, w, C.c_void_p(rois.ctypes.data), r, pooled_h,
                           pooled_w, C.c_float(spatial_scale), C.c_float(pad_ratio), C.c_void_p(y.ctypes.data))
    return y


def decode_bbox(bbox_pred, prior, bbox_mean=(0, 0, 0, 0), bbox_std=(1, 1, 1, 1)) -> np.ndarray:
    """DecodeBBoxLayer::Forward_cpu, TEST phase (decode_bbox_layer.cpp:53-124, math_functions.cpp:46-77)."""
    b, q = _f32(bbox_pred), _f32(prior).reshape(-1, 5)
    b = b.reshape(len(q), -1)
    out = np.empty((len(q), 5), dtype=np.float32)
    m, sd = _f32(bbox_mean), _f32(bbox_std)
    lib().oracle_decode_bbox(C.c_void_p(b.ctypes.data), C.c_void_p(q.ctypes.data), len(q), b.shape[1],
                             C.c_void_p(m.ctypes.data), C.c_void_p(sd.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def softmax(x, axis=1) -> np.ndarray:
    """SoftmaxLayer::Forward_cpu (softmax_layer.cpp:28-62): subtract the max, exp (single), divide by the
    sum.  The reference sums with cblas_sgemv, i.e. in BLAS-defined order; here channel order."""
    x = _f32(x)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp((x - m).astype(np.float64)).astype(np.float32)
    s = np.zeros_like(m)
    for k in range(x.shape[axis]):
        s = (s + np.take(e, [k], axis=axis)).astype(np.float32)
    return (e / s).astype(np.float32)


def eltwise(bottoms, op="SUM", coeffs=None) -> np.ndarray:
    """EltwiseLayer::Forward_cpu (eltwise_layer.cpp:46-96)."""
    bs = [_f32(b) for b in bottoms]
    if op == "PROD":
        r = bs[0] * bs[1]
        for b in bs[2:]:
            r = r * b
        return r
    if op == "SUM":
        cs = [np.float32(1)] * len(bs) if coeffs is None else [np.float32(c) for c in coeffs]
        r = np.zeros_like(bs[0])
        for c, b in zip(cs, bs):
            r = (r + c * b).astype(np.float32)      # caffe_axpy per bottom (:66-68)
        return r
    r = np.where(bs[0] > bs[1], bs[0], bs[1])       # :73-80
    for b in bs[2:]:
        r = np.where(b > r, b, r)
    return r


def bbnms_maxg(bbs, overlap=0.5) -> np.ndarray:
    """bbNms(bbs,'type','maxg','overlap',o,'ovrDnm','union') -> kept row indices in kept order
    (utils/bbNms.m:75-126)."""
    b = np.ascontiguousarray(bbs, dtype=np.float64).reshape(-1, 5)
    order = np.zeros(max(len(b), 1), dtype=np.int32)
    m = lib().oracle_bbnms_maxg(b.ctypes.data, len(b), float(overlap), order.ctypes.data)
    return order[:m].copy()


def detect_postprocess(proposals_score, cls_pred, bbox_pred, cls_id=2, bbox_mean=(0, 0, 0, 0),
                       bbox_std=(0.1, 0.1, 0.2, 0.2), proposal_thr=-10.0, overlap=0.5, ratios=(1.0, 1.0),
                       org_hw=None, net_hw=None) -> np.ndarray:
    """The MATLAB post-process of ONE image (examples/kitti_car/run_mscnn_detection.m:75-120):
    rows [img x1 y1 x2 y2 score] -> final detections [K,5] = [x y w h prob] for class `cls_id`
    (1-based).  MATLAB computes the decode in single precision and bbNms in double."""
    f = np.float32
    p = _f32(proposals_score).reshape(-1, 6)[:, 1:].copy()          # :77
    if len(p) == 0:
        return np.zeros((0, 5), dtype=np.float32)
    p[:, 2] = p[:, 2] - p[:, 0]
    p[:, 3] = p[:, 3] - p[:, 1]                                      # :78
    cls_pred = _f32(cls_pred).reshape(len(p), -1)
    bbox_pred = _f32(bbox_pred).reshape(len(p), -1)
    keep = (p[:, 4] >= f(proposal_thr)) & (p[:, 2] != 0) & (p[:, 3] != 0)   # :82
    p, cls_pred, bbox_pred = p[keep], cls_pred[keep], bbox_pred[keep]
    ratio_h, ratio_w = f(ratios[0]), f(ratios[1])
    if org_hw is None:
        org_hw = net_hw
    org_h, org_w = f(org_hw[0]), f(org_hw[1])
    bb = bbox_pred[:, 4 * cls_id - 4:4 * cls_id] * _f32(bbox_std)[None, :] + _f32(bbox_mean)[None, :]   # :95-99
    e = np.exp(cls_pred.astype(np.float64)).astype(np.float32)      # single exp, correctly rounded
    s = np.zeros(len(p), dtype=np.float32)
    for k in range(e.shape[1]):
        s = (s + e[:, k]).astype(np.float32)
    prob = (e[:, cls_id - 1] / s).astype(np.float32)                # :101-103
    cx = p[:, 0] + f(0.5) * p[:, 2]
    cy = p[:, 1] + f(0.5) * p[:, 3]
    tx = bb[:, 0] * p[:, 2] + cx
    ty = bb[:, 1] * p[:, 3] + cy
    tw = p[:, 2] * np.exp(bb[:, 2].astype(np.float64)).astype(np.float32)
    th = p[:, 3] * np.exp(bb[:, 3].astype(np.float64)).astype(np.float32)
    tx = tx - tw / f(2)
    ty = ty - th / f(2)
    tx, tw = tx / ratio_w, tw / ratio_w
    ty, th = ty / ratio_h, th / ratio_h                             # :110-111
    tx = np.maximum(f(0), tx)
    ty = np.maximum(f(0), ty)
    tw = np.minimum(tw, org_w - tx)
    th = np.minimum(th, org_h - ty)                                  # :114-115
    bbs = np.stack([tx, ty, tw, th, prob], axis=1).astype(np.float64)
    bbs = bbs[~np.isnan(bbs[:, 4])]                                  # bbNms.m:82 kp = score > -inf
    return bbs[bbnms_maxg(bbs, overlap)].astype(np.float32)


def cascade_detect_postprocess(proposals, cls_prob, output_bbox, cls_id=2, overlap=0.5, ratios=(1.0, 1.0),
                               org_hw=None, net_hw=None) -> np.ndarray:
    """The MATLAB post-process of ONE image of the cascade driver
    (examples/kitti_car/run_cascademscnn.m:99-126): stage blobs proposals [R,5], cls_prob [R,C],
    output_bbox [R,5] -> final detections [K,5] = [x y w h prob] for class `cls_id` (1-based)."""
    f = np.float32
    q = _f32(proposals).reshape(-1, 5)[:, 1:].copy()
    if len(q) == 0:
        return np.zeros((0, 5), dtype=np.float32)
    t = _f32(output_bbox).reshape(len(q), 5)[:, 1:].copy()           # :99-100
    cls_prob = _f32(cls_prob).reshape(len(q), -1)
    if org_hw is None:
        org_hw = net_hw
    ratio_h, ratio_w = f(ratios[0]), f(ratios[1])
    org_h, org_w = f(org_hw[0]), f(org_hw[1])
    t[:, [0, 2]] = t[:, [0, 2]] / ratio_w                             # :101-102
    t[:, [1, 3]] = t[:, [1, 3]] / ratio_h
    t[:, [0, 1]] = np.maximum(f(0), t[:, [0, 1]])                     # :104
    t[:, 2] = np.minimum(t[:, 2], org_w)                              # :105
    t[:, 3] = np.minimum(t[:, 3], org_h)
    t[:, [2, 3]] = t[:, [2, 3]] - t[:, [0, 1]] + f(1)                 # :106
    q[:, [2, 3]] = q[:, [2, 3]] - q[:, [0, 1]] + f(1)                 # :114
    keep = (q[:, 2] != 0) & (q[:, 3] != 0)                            # :117
    t, prob = t[keep], cls_prob[keep][:, cls_id - 1]
    bbs = np.concatenate([t, prob[:, None]], axis=1).astype(np.float64)   # :124, det_thr = -1: no threshold
    bbs = bbs[~np.isnan(bbs[:, 4])]
    return bbs[bbnms_maxg(bbs, overlap)].astype(np.float32)


# ---------------------------------------------------------------------------------------------
# Pre-processing in front of net.forward: examples/kitti_car/run_mscnn_detection.m:64-69,
# examples/widerface/run_mscnn_detection.m:70-86.  `imresize` is MathWorks code that is NOT part of the reference
# repository; what follows restates its published algorithm (imresize.m: `contributions`, `cubic`; resize order by
# ascending scale; per-pass saturating round for integer images).  PARITY UNPINNED against MATLAB itself (no
# MATLAB / Octave here); tests/test_oracle.py cross-checks the tap tables against torch's antialiased bicubic
# (same a = -0.5 kernel, same support widening) on float data.
def _cubic(x: np.ndarray) -> np.ndarray:
    a = np.abs(x)
    a2 = a * a
    a3 = a2 * a
    f1 = (1.5 * a3 - 2.5 * a2) + 1.0
    f2 = ((-0.5 * a3 + 2.5 * a2) - 4.0 * a) + 2.0
    return np.where(a <= 1.0, f1, np.where(a <= 2.0, f2, 0.0))


def imresize_contributions(in_len: int, out_len: int) -> tuple[np.ndarray, np.ndarray]:
    """(weights fp64 [out][P], indices int [out][P], 0-based) of imresize.m's `contributions` for the bicubic
    kernel with antialiasing (the defaults of `imresize(img, [H W])`)."""
    scale = float(out_len) / float(in_len)
    aa = scale < 1.0
    kw = 4.0 / scale if aa else 4.0
    x = np.arange(1, out_len + 1, dtype=np.float64)
    u = x / scale + 0.5 * (1.0 - 1.0 / scale)
    left = np.floor(u - kw / 2.0)
    P = int(np.ceil(kw)) + 2
    ind = left[:, None] + np.arange(P, dtype=np.float64)[None, :]          # 1-based positions
    d = u[:, None] - ind
    w = scale * _cubic(scale * d) if aa else _cubic(d)
    s = np.zeros(out_len)
    for k in range(P):                                                       # sequential row sum
        s = s + w[:, k]
    w = w / s[:, None]
    m = np.mod(ind.astype(np.int64) - 1, 2 * in_len)                          # aux = [1:in, in:-1:1]
    idx = np.where(m < in_len, m, 2 * in_len - 1 - m)
    keep = np.any(w != 0.0, axis=0)
    return np.ascontiguousarray(w[:, keep]), np.ascontiguousarray(idx[:, keep]).astype(np.int32)


def _round_half_away(v: np.ndarray) -> np.ndarray:
    f = np.floor(v)
    return np.where(v - f >= 0.5, f + 1.0, f)


def _resize_along(img: np.ndarray, dim: int, w: np.ndarray, idx: np.ndarray, to_u8: bool) -> np.ndarray:
    src = img.astype(np.float64)
    out_len, P = w.shape
    shape = list(src.shape)
    shape[dim] = out_len
    acc = np.zeros(shape)
    for k in range(P):                                                       # tap order = table order
        taps = np.take(src, idx[:, k], axis=dim)
        wk = w[:, k].reshape([-1 if a == dim else 1 for a in range(src.ndim)])
        acc = acc + wk * taps
    if to_u8:
        acc = _round_half_away(np.clip(acc, 0.0, 255.0)).astype(np.uint8)
    return acc


def imresize(img: np.ndarray, out_hw: tuple[int, int]) -> np.ndarray:
    """imresize(img, [H W]) for an H x W x C image: uint8 in -> uint8 out (rounded after each pass); float in ->
    float64 out (no rounding)."""
    to_u8 = img.dtype == np.uint8
    tabs = [imresize_contributions(img.shape[k], out_hw[k]) for k in range(2)]
    scales = [out_hw[k] / img.shape[k] for k in range(2)]
    order = [0, 1] if scales[0] <= scales[1] else [1, 0]                    # sort(scale), stable
    out = img
    for dim in order:
        out = _resize_along(out, dim, tabs[dim][0], tabs[dim][1], to_u8)
    return out


def preprocess(img_rgb_u8: np.ndarray, net_hw: tuple[int, int], mean_bgr=(104.0, 117.0, 123.0)) -> np.ndarray:
    """run_mscnn_detection.m:64-69: uint8 H x W x 3 RGB image -> fp32 3 x netH x netW (BGR, mean-subtracted)."""
    r = imresize(img_rgb_u8, net_hw)
    bgr = r[:, :, ::-1].astype(np.float32)
    bgr = bgr - np.asarray(mean_bgr, dtype=np.float32)[None, None, :]
    return np.ascontiguousarray(bgr.transpose(2, 0, 1))


def widerface_net_size(org_h: int, org_w: int, img_h: int = 0, img_w: int = 0, max_size: int = 2048) -> tuple[int, int]:
    """examples/widerface/run_mscnn_detection.m:72-80."""
    def rnd(v):
        return float(np.floor(v + 0.5))
    w = float(org_w if img_w == 0 else img_w)
    h = float(org_h if img_h == 0 else img_h)
    w = rnd(w / 32.0) * 32.0
    h = rnd(h / 32.0) * 32.0
    if h > max_size or w > max_size:
        r = max_size / max(h, w)
        h = rnd(h * r / 32.0) * 32.0
        w = rnd(w * r / 32.0) * 32.0
    return int(h), int(w)
