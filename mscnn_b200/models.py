"""Deploy-net definitions of the MS-CNN model zoo as generated .prototxt text.

The reference ships these as examples/*/*/mscnn_deploy.prototxt; any of those files loads
unchanged (tests/test_net_cpu.py checks that when /root/reference is mounted).  They are not
copied into this repository; bench.py and the tests need the same nets where the reference
tree is absent (the GPU box), so this module re-generates them from a compact description and
tests/test_net_cpu.py verifies layer-by-layer equality with the shipped files.

    kitti(576, 1920, scales=7)            -> examples/kitti_car/mscnn-7s-576/mscnn_deploy.prototxt
    kitti(576, 1920, scales=7, up2x=True) -> .../mscnn-7s-576-2x/mscnn_deploy.prototxt
    kitti(768, 2560, scales=8)            -> .../mscnn-8s-768-trainval/mscnn_deploy.prototxt
    widerface(512, 512)                   -> examples/widerface/mscnn-12s-2x/mscnn_deploy.prototxt
    kitti_cascade(576, 1920)              -> examples/kitti_car/cascade-mscnn-7s-576-2x/mscnn_deploy.prototxt
    widerface_cascade(512, 512)           -> examples/widerface/cascade-mscnn-12s-align/mscnn_deploy.prototxt
"""
from __future__ import annotations

VGG = [("conv1_1", 64), ("conv1_2", 64), "pool1", ("conv2_1", 128), ("conv2_2", 128), "pool2",
       ("conv3_1", 256), ("conv3_2", 256), ("conv3_3", 256), "pool3",
       ("conv4_1", 512), ("conv4_2", 512), ("conv4_3", 512)]


def _conv(name, bottom, top, cout, k, pad=None, relu=None):
    pad_s = f" pad: {pad}" if pad is not None else ""
    s = (f'layer {{ bottom: "{bottom}" top: "{top}" name: "{name}" type: "Convolution" '
         f'convolution_param {{ num_output: {cout}{pad_s} kernel_size: {k} }} }}\n')
    if relu:
        s += f'layer {{ bottom: "{top}" top: "{top}" name: "{relu}" type: "ReLU" }}\n'
    return s


def _pool(name, bottom, mode="MAX"):
    return (f'layer {{ bottom: "{bottom}" top: "{name}" name: "{name}" type: "Pooling" '
            f'pooling_param {{ pool: {mode} kernel_size: 2 stride: 2 }} }}\n')


def _trunk(net_name, n, h, w):
    s = f'name: "{net_name}"\ninput: "data"\ninput_dim: {n}\ninput_dim: 3\ninput_dim: {h}\ninput_dim: {w}\n'
    prev = "data"
    for item in VGG:
        if isinstance(item, str):
            s += _pool(item, prev)
            prev = item
        else:
            name, c = item
            s += _conv(name, prev, name, c, 3, 1, relu="relu" + name[4:])
            prev = name
    return s


def _head_kitti(up2x: bool, pooled: int = 7):
    feat, scale = ("conv4_3_2x", 0.25) if up2x else ("conv4_3", 0.125)
    s = ""
    if up2x:
        s += ('layer { bottom: "conv4_3" top: "conv4_3_2x" name: "conv4_3_2x" type: "Deconvolution" '
              'convolution_param { kernel_size: 4 stride: 2 num_output: 512 group: 512 pad: 1 '
              'weight_filler: { type: "bilinear" } bias_term: false } param { lr_mult: 0 decay_mult: 0 } }\n')
    for nm, pr in (("roi_pool_org", 0), ("roi_pool_ctx", 0.25)):
        s += (f'layer {{ name: "{nm}" type: "ROIPooling" bottom: "{feat}" bottom: "proposals" top: "{nm}" '
              f'roi_pooling_param {{ pooled_w: {pooled} pooled_h: {pooled} spatial_scale: {scale} pad_ratio: {pr} }} }}\n')
    s += 'layer { name: "roi_pool" type: "Concat" bottom: "roi_pool_org" bottom: "roi_pool_ctx" top: "roi_pool" }\n'
    return s


def _fc_head(fc6, ncls, roi_c1_pad=None):
    s = _conv("roi_c1", "roi_pool", "roi_c1", 512, 3, roi_c1_pad, relu="roi_c1_relu")
    s += f'layer {{ name: "fc6" type: "InnerProduct" bottom: "roi_c1" top: "fc6" inner_product_param {{ num_output: {fc6} }} }}\n'
    s += 'layer { name: "relu6" type: "ReLU" bottom: "fc6" top: "fc6" }\n'
    s += 'layer { name: "drop6" type: "Dropout" bottom: "fc6" top: "fc6" dropout_param { dropout_ratio: 0.5 } }\n'
    s += f'layer {{ name: "cls_pred" type: "InnerProduct" bottom: "fc6" top: "cls_pred" inner_product_param {{ num_output: {ncls} }} }}\n'
    s += f'layer {{ name: "bbox_pred" type: "InnerProduct" bottom: "fc6" top: "bbox_pred" inner_product_param {{ num_output: {4 * ncls} }} }}\n'
    return s


def kitti(h: int, w: int, scales: int = 8, up2x: bool = False, batch: int = 1, max_nms_num: int = 2000) -> str:
    """KITTI-car MS-CNN deploy net with 7 or 8 proposal scales."""
    assert scales in (7, 8)
    s = _trunk("VGG_ILSVRC_16_layers", batch, h, w)
    s += _conv("loss1_conv1", "conv4_3", "loss1_conv1", 512, 3, 1, relu="loss_relu1")
    s += _conv("LFCN_1_5x5", "loss1_conv1", "LFCN_1_5x5", 9, 5, 2)
    s += _conv("LFCN_1_7x7", "loss1_conv1", "LFCN_1_7x7", 9, 7, 3)
    s += _pool("pool4", "conv4_3")
    prev = "pool4"
    for nm in ("conv5_1", "conv5_2", "conv5_3"):
        s += _conv(nm, prev, nm, 512, 3, 1, relu="relu" + nm[4:])
        prev = nm
    s += _conv("LFCN_2_5x5", "conv5_3", "LFCN_2_5x5", 9, 5, 2)
    s += _conv("LFCN_2_7x7", "conv5_3", "LFCN_2_7x7", 9, 7, 3)
    s += _pool("pool5", "conv5_3")
    s += _conv("conv6_1", "pool5", "conv6_1", 512, 3, 1, relu="relu6_1")
    s += _conv("LFCN_3_5x5", "conv6_1", "LFCN_3_5x5", 9, 5, 2)
    s += _conv("LFCN_3_7x7", "conv6_1", "LFCN_3_7x7", 9, 7, 3)
    s += _pool("pool6", "conv6_1")
    s += _conv("LFCN_4_5x5", "pool6", "LFCN_4_5x5", 9, 5, 2)
    bottoms = ["LFCN_1_5x5", "LFCN_1_7x7", "LFCN_2_5x5", "LFCN_2_7x7", "LFCN_3_5x5", "LFCN_3_7x7", "LFCN_4_5x5"]
    if scales == 8:
        s += _conv("LFCN_4_7x7", "pool6", "LFCN_4_7x7", 9, 7, 3)
        bottoms.append("LFCN_4_7x7")
    fields = [60, 84, 120, 168, 240, 336, 480, 672][:scales]
    rates = [8, 8, 16, 16, 32, 32, 64, 64][:scales]
    s += "layer { " + " ".join(f'bottom: "{b}"' for b in bottoms)
    s += ' top: "proposals" top: "proposals_score" name: "proposals" type: "BoxOutput" box_output_param { '
    s += 'fg_thr: -5 iou_thr: 0.65 nms_type: "IOU" '
    s += " ".join(f"field_w: {f}" for f in fields) + " " + " ".join(f"field_h: {f}" for f in fields) + " "
    s += " ".join(f"downsample_rate: {r}" for r in rates)
    s += f" field_whr: 2 field_xyr: 2 max_nms_num: {max_nms_num} }} }}\n"
    s += _head_kitti(up2x)
    s += _fc_head(4096, 5)
    return s


def widerface(h: int = 512, w: int = 512, batch: int = 1, max_nms_num: int = 3000) -> str:
    """WIDER FACE mscnn-12s-2x deploy net (12 proposal heads of 1x1 convs, AVE pool6, ROI 5x5)."""
    s = _trunk("VGG_ILSVRC_16_layers", batch, h, w)
    heads = []

    def rpn(idx, bottom, sizes):
        nonlocal s
        s += _conv(f"rpn_{idx}_conv", bottom, f"rpn_{idx}_conv", 512, 3, 1, relu=f"rpn_{idx}_relu")
        for z in sizes:
            nm = f"LFCN_{idx}_{z}x{z}"
            s += _conv(nm, f"rpn_{idx}_conv", nm, 6, 1, 0)
            heads.append((nm, z))

    rpn(1, "conv4_3", [12, 16, 24, 32, 48])
    s += _pool("pool4", "conv4_3")
    prev = "pool4"
    for nm in ("conv5_1", "conv5_2", "conv5_3"):
        s += _conv(nm, prev, nm, 512, 3, 1, relu="relu" + nm[4:])
        prev = nm
    rpn(2, "conv5_3", [64, 96])
    s += _pool("pool5", "conv5_3")
    rpn(3, "pool5", [128, 192])   # NB: the shipped file gives LFCN_3_192x192 a 196-pixel field
    s += _pool("pool6", "pool5", "AVE")
    rpn(4, "pool6", [256, 384, 480])
    rates = [8] * 5 + [16] * 2 + [32] * 2 + [64] * 3
    s += "layer { " + " ".join(f'bottom: "{b}"' for b, _ in heads)
    s += ' top: "proposals" top: "proposals_score" name: "proposals" type: "BoxOutput" box_output_param { '
    s += 'fg_thr: -3 iou_thr: 0.65 nms_type: "IOU" '
    fields = [196 if z == 192 else z for _, z in heads]
    s += " ".join(f"field_w: {z}" for z in fields) + " " + " ".join(f"field_h: {z}" for z in fields) + " "
    s += " ".join(f"downsample_rate: {r}" for r in rates)
    s += f" field_whr: 4 field_xyr: 1 min_size: 5 max_nms_num: {max_nms_num} }} "
    s += "bbox_reg_param { bbox_mean: 0 bbox_mean: 0 bbox_mean: 0 bbox_mean: 0 bbox_std: 0.1 bbox_std: 0.1 bbox_std: 0.2 bbox_std: 0.2 } }\n"
    s += _head_kitti(True, pooled=5)
    s += _fc_head(2048, 2, roi_c1_pad=1)
    return s


# ---------------------------------------------------------------------------- cascade nets
# Three detection stages; stage k+1 pools the boxes stage k regressed (DecodeBBox), with regression
# statistics that tighten from stage to stage (bbox_std below).  SURVEY.md section 8(f) rank 2.
CASCADE_STD = [(0.1, 0.1, 0.2, 0.2), (0.05, 0.05, 0.1, 0.1), (0.033, 0.033, 0.067, 0.067)]
ORD = ["1st", "2nd", "3rd"]


def _share(names):
    return "".join(f'param {{ name: "{n}" }} ' for n in names) if names else ""


def _conv_shared(name, bottom, cout, k, pad, relu, share=None):
    pad_s = f" pad: {pad}" if pad is not None else ""
    return (f'layer {{ bottom: "{bottom}" top: "{name}" name: "{name}" type: "Convolution" {_share(share)}'
            f'convolution_param {{ num_output: {cout} kernel_size: {k}{pad_s} }} }}\n'
            f'layer {{ bottom: "{name}" top: "{name}" name: "{relu}" type: "ReLU" }}\n')


def _ip(name, bottom, nout, share=None):
    return (f'layer {{ name: "{name}" type: "InnerProduct" bottom: "{bottom}" top: "{name}" {_share(share)}'
            f'inner_product_param {{ num_output: {nout} }} }}\n')


def _stage_head(sfx, pooled_blob, fc6, ncls, roi_c1_pad, share_tag=None, with_bbox=True):
    """roi_c1 -> fc6 -> cls_pred (-> bbox_pred) on `pooled_blob`; layer names carry `sfx`.  With share_tag the
    weights are shared by ParamSpec name with the stage that owns them (third-stage ensemble heads)."""
    sh = (lambda base: [f"{base}{share_tag}_w", f"{base}{share_tag}_b"]) if share_tag is not None else (lambda base: None)
    s = _conv_shared(f"roi_c1{sfx}", pooled_blob, 512, 3, roi_c1_pad, f"roi_c1_relu{sfx}", sh("roi_c1"))
    s += _ip(f"fc6{sfx}", f"roi_c1{sfx}", fc6, sh("fc6"))
    s += f'layer {{ name: "relu6{sfx}" type: "ReLU" bottom: "fc6{sfx}" top: "fc6{sfx}" }}\n'
    s += (f'layer {{ name: "drop6{sfx}" type: "Dropout" bottom: "fc6{sfx}" top: "fc6{sfx}" '
          f'dropout_param {{ dropout_ratio: 0.5 }} }}\n')
    s += _ip(f"cls_pred{sfx}", f"fc6{sfx}", ncls, sh("cls_pred"))
    if with_bbox:
        s += _ip(f"bbox_pred{sfx}", f"fc6{sfx}", 8)
    return s


def _decode(name, bbox, prior, std, no_grad=False):
    st = " ".join(f"bbox_mean: 0" for _ in range(4)) + " " + " ".join(f"bbox_std: {v}" for v in std)
    pd = " propagate_down: 0 propagate_down: 0" if no_grad else ""
    return (f'layer {{ name: "{name}" type: "DecodeBBox" bottom: "{bbox}" bottom: "{prior}" top: "{name}" '
            f'bbox_reg_param {{ {st} }}{pd} }}\n')


def _softmax(name, bottom):
    return f'layer {{ name: "{name}" type: "Softmax" bottom: "{bottom}" top: "{name}" softmax_param {{ axis: 1 }} }}\n'


def _roi_stage(sfx, feat, rois, scale, pooled, align):
    """org + ctx pooling of `rois` on `feat`, concatenated to roi_pool{sfx}."""
    s = ""
    for kind, pr in (("org", 0), ("ctx", 0.25)):
        if align:   # ROIAlign grid of (pooled+1)^2 corner samples, averaged 2x2 -> pooled^2
            s += (f'layer {{ name: "roi_grid_{kind}{sfx}" type: "ROIAlign" bottom: "{feat}" bottom: "{rois}" '
                  f'top: "roi_grid_{kind}{sfx}" roi_pooling_param {{ pooled_w: {pooled} pooled_h: {pooled} '
                  f'spatial_scale: {scale} pad_ratio: {pr} }} }}\n')
            s += (f'layer {{ name: "roi_pool_{kind}{sfx}" type: "Pooling" bottom: "roi_grid_{kind}{sfx}" '
                  f'top: "roi_pool_{kind}{sfx}" pooling_param {{ pool: AVE kernel_size: 2 stride: 1 }} }}\n')
        else:
            s += (f'layer {{ name: "roi_pool_{kind}{sfx}" type: "ROIPooling" bottom: "{feat}" bottom: "{rois}" '
                  f'top: "roi_pool_{kind}{sfx}" roi_pooling_param {{ pooled_w: {pooled} pooled_h: {pooled} '
                  f'spatial_scale: {scale} pad_ratio: {pr} }} }}\n')
    s += (f'layer {{ name: "roi_pool{sfx}" type: "Concat" bottom: "roi_pool_org{sfx}" bottom: "roi_pool_ctx{sfx}" '
          f'top: "roi_pool{sfx}" }}\n')
    return s


def kitti_cascade(h: int, w: int, batch: int = 1, max_nms_num: int = 2000) -> str:
    """KITTI-car cascade-mscnn-7s-576-2x deploy net: the 7-scale "-2x" proposal net followed by three
    detection stages (examples/kitti_car/cascade-mscnn-7s-576-2x/mscnn_deploy.prototxt:440-949)."""
    base = kitti(h, w, 7, True, batch, max_nms_num)
    s = base[:base.index('layer { name: "roi_pool_org"')]            # trunk + proposals + conv4_3_2x
    rois = "proposals"
    for k, sfx in enumerate(["", "_2nd", "_3rd"]):
        s += _roi_stage(sfx, "conv4_3_2x", rois, 0.25, 7, align=False)
        s += _stage_head(sfx, f"roi_pool{sfx}", 4096, 5, None)
        if k < 2:
            nxt = f"proposals_{ORD[k + 1]}"
            s += _decode(nxt, f"bbox_pred{sfx}", rois, CASCADE_STD[k], no_grad=True)
            rois = nxt
    priors = ["proposals", "proposals_2nd", "proposals_3rd"]
    for k, sfx in enumerate(["", "_2nd", "_3rd"]):
        s += _decode(f"output_bbox_{ORD[k]}", f"bbox_pred{sfx}", priors[k], CASCADE_STD[k])
    for k, sfx in enumerate(["", "_2nd", "_3rd"]):
        s += _softmax(f"cls_prob_{ORD[k]}", f"cls_pred{sfx}")
    return s


def widerface_cascade(h: int = 512, w: int = 512, batch: int = 1, max_nms_num: int = 3000) -> str:
    """WIDER FACE cascade-mscnn-12s-align deploy net: ROIAlign + 2x2 AVE pooling instead of ROIPooling, no
    2x upsampling, and a third stage that also runs the first- and second-stage heads (weights shared by
    ParamSpec name) and averages the three class probabilities
    (examples/widerface/cascade-mscnn-12s-align/mscnn_deploy.prototxt:811-1674)."""
    base = widerface(h, w, batch, max_nms_num).replace('name: "VGG_ILSVRC_16_layers"', 'name: "MSCNN"')
    cut = base.index('layer { bottom: "conv4_3" top: "conv4_3_2x"')
    s = base[:cut]
    rois = "proposals"
    owners = ["", "_2nd"]
    for k, sfx in enumerate(["", "_2nd"]):
        s += _roi_stage(sfx, "conv4_3", rois, 0.125, 5, align=True)
        s += _stage_head(sfx, f"roi_pool{sfx}", 2048, 2, 1, share_tag=owners[k])
        nxt = f"proposals_{ORD[k + 1]}"
        s += _decode(nxt, f"bbox_pred{sfx}", rois, CASCADE_STD[k], no_grad=True)
        rois = nxt
    s += _roi_stage("_3rd", "conv4_3", rois, 0.125, 5, align=True)
    s += _stage_head("_1st_3rd", "roi_pool_3rd", 2048, 2, 1, share_tag="", with_bbox=False)
    s += _stage_head("_2nd_3rd", "roi_pool_3rd", 2048, 2, 1, share_tag="_2nd", with_bbox=False)
    s += _stage_head("_3rd", "roi_pool_3rd", 2048, 2, 1)
    priors = ["proposals", "proposals_2nd", "proposals_3rd"]
    for k, sfx in enumerate(["", "_2nd", "_3rd"]):
        s += _decode(f"output_bbox_{ORD[k]}", f"bbox_pred{sfx}", priors[k], CASCADE_STD[k])
    s += _softmax("cls_prob_1st", "cls_pred") + _softmax("cls_prob_2nd", "cls_pred_2nd")
    s += _softmax("cls_prob_1st_3rd", "cls_pred_1st_3rd") + _softmax("cls_prob_2nd_3rd", "cls_pred_2nd_3rd")
    s += _softmax("cls_prob_3rd", "cls_pred_3rd")
    s += ('layer { name: "cls_prob_3rd_avg" type: "Eltwise" bottom: "cls_prob_1st_3rd" bottom: "cls_prob_2nd_3rd" '
          'bottom: "cls_prob_3rd" top: "cls_prob_3rd_avg" eltwise_param { operation: SUM coeff: 0.33333333 '
          'coeff: 0.33333333 coeff: 0.33333333 } }\n')
    return s


CONFIGS = {
    "mscnn-7s-576": lambda batch=1: kitti(576, 1920, 7, False, batch),
    "mscnn-7s-576-2x": lambda batch=1: kitti(576, 1920, 7, True, batch),
    "mscnn-8s-768": lambda batch=1: kitti(768, 2560, 8, False, batch),
    "widerface-12s-2x": lambda batch=1, h=768, w=1024: widerface(h, w, batch),
    "cascade-mscnn-7s-576-2x": lambda batch=1: kitti_cascade(576, 1920, batch),
    "cascade-widerface-12s-align": lambda batch=1, h=768, w=1024: widerface_cascade(h, w, batch),
}
