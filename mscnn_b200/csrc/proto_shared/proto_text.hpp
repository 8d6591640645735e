// Typed view of a parsed .prototxt: fills the caffe.pb.h message classes from the schema-less
// tree produced by prototxt.hpp.  Together the two files play the role of
// ReadProtoFromTextFile (/root/reference/src/caffe/util/io.cpp:34-44) for the messages the
// MS-CNN deploy nets use.  Unknown fields inside known messages are rejected (like protobuf's
// text parser) except the training-only ones the deploy files are known to carry.
#pragma once
#include <set>
#include <stdexcept>
#include <string>

#include "caffe/proto/caffe.pb.h"
#include "prototxt.hpp"

namespace caffe {
namespace proto_text {

using prototxt::Field;
using prototxt::Node;

inline void check_known(const Node& n, const char* msg, const std::set<std::string>& known) {
  for (const Field& f : n.fields)
    if (!known.count(f.name))
      throw std::runtime_error(std::string("prototxt: line ") + std::to_string(f.line) + ": message " +
                               msg + " has no field named \"" + f.name + "\" (or it is outside the "
                               "subset mscnn_b200 supports)");
}

inline void fill(const Node* n, FillerParameter* f) {
  if (!n) return;
  check_known(*n, "FillerParameter", {"type", "value", "min", "max", "mean", "std", "sparse", "variance_norm"});
  if (n->has("type")) f->set_type(n->str("type"));
  if (n->has("value")) f->set_value((float)n->num("value", 0));
  if (n->has("min")) f->set_min((float)n->num("min", 0));
  if (n->has("max")) f->set_max((float)n->num("max", 1));
  if (n->has("mean")) f->set_mean((float)n->num("mean", 0));
  if (n->has("std")) f->set_std((float)n->num("std", 1));
  if (n->has("sparse")) f->set_sparse((int)n->num("sparse", -1));
  if (n->has("variance_norm")) {
    const std::string v = n->str("variance_norm");
    f->set_variance_norm(v == "FAN_OUT" ? FillerParameter_VarianceNorm_FAN_OUT
                         : v == "AVERAGE" ? FillerParameter_VarianceNorm_AVERAGE
                                          : FillerParameter_VarianceNorm_FAN_IN);
  }
}

inline void fill(const Node* n, BlobShape* s) {
  if (!n) return;
  for (double d : n->nums("dim")) s->add_dim((int64_t)d);
}

inline void fill(const Node& n, LayerParameter* lp) {
  check_known(n, "LayerParameter",
              {"name", "type", "bottom", "top", "phase", "loss_weight", "param", "propagate_down", "include",
               "exclude", "convolution_param", "pooling_param", "inner_product_param", "input_param",
               "dropout_param", "concat_param", "relu_param", "roi_pooling_param", "box_output_param",
               "bbox_reg_param", "decode_bbox_param", "softmax_param", "eltwise_param"});
  lp->set_name(n.str("name"));
  lp->set_type(n.str("type"));
  for (const std::string& s : n.strs("bottom")) lp->add_bottom(s);
  for (const std::string& s : n.strs("top")) lp->add_top(s);
  if (n.has("phase")) lp->set_phase(n.str("phase") == "TRAIN" ? TRAIN : TEST);
  for (double v : n.nums("loss_weight")) lp->add_loss_weight((float)v);
  for (const Field* f : n.all("propagate_down")) lp->add_propagate_down(f->scalar == "true" || f->scalar == "1");
  for (const Field* f : n.all("param")) {
    ParamSpec* ps = lp->add_param();
    if (!f->is_message) continue;
    if (f->message->has("name")) ps->set_name(f->message->str("name"));
    if (f->message->has("lr_mult")) ps->set_lr_mult((float)f->message->num("lr_mult", 1));
    if (f->message->has("decay_mult")) ps->set_decay_mult((float)f->message->num("decay_mult", 1));
  }
  if (const Node* c = n.child("convolution_param")) {
    check_known(*c, "ConvolutionParameter",
                {"num_output", "bias_term", "pad", "kernel_size", "stride", "dilation", "pad_h", "pad_w",
                 "kernel_h", "kernel_w", "stride_h", "stride_w", "group", "weight_filler", "bias_filler",
                 "engine", "axis", "force_nd_im2col"});
    ConvolutionParameter* p = lp->mutable_convolution_param();
    if (c->has("num_output")) p->set_num_output((uint32_t)c->num("num_output", 0));
    if (c->has("bias_term")) p->set_bias_term(c->boolean("bias_term", true));
    for (double v : c->nums("pad")) p->add_pad((uint32_t)v);
    for (double v : c->nums("kernel_size")) p->add_kernel_size((uint32_t)v);
    for (double v : c->nums("stride")) p->add_stride((uint32_t)v);
    for (double v : c->nums("dilation")) p->add_dilation((uint32_t)v);
    if (c->has("pad_h")) p->set_pad_h((uint32_t)c->num("pad_h", 0));
    if (c->has("pad_w")) p->set_pad_w((uint32_t)c->num("pad_w", 0));
    if (c->has("kernel_h")) p->set_kernel_h((uint32_t)c->num("kernel_h", 0));
    if (c->has("kernel_w")) p->set_kernel_w((uint32_t)c->num("kernel_w", 0));
    if (c->has("stride_h")) p->set_stride_h((uint32_t)c->num("stride_h", 0));
    if (c->has("stride_w")) p->set_stride_w((uint32_t)c->num("stride_w", 0));
    if (c->has("group")) p->set_group((uint32_t)c->num("group", 1));
    if (c->has("axis")) p->set_axis((int)c->num("axis", 1));
    if (c->has("force_nd_im2col")) p->set_force_nd_im2col(c->boolean("force_nd_im2col", false));
    if (c->child("weight_filler")) fill(c->child("weight_filler"), p->mutable_weight_filler());
    if (c->child("bias_filler")) fill(c->child("bias_filler"), p->mutable_bias_filler());
  }
  if (const Node* c = n.child("pooling_param")) {
    check_known(*c, "PoolingParameter",
                {"pool", "pad", "pad_h", "pad_w", "kernel_size", "kernel_h", "kernel_w", "stride", "stride_h",
                 "stride_w", "engine", "global_pooling"});
    PoolingParameter* p = lp->mutable_pooling_param();
    if (c->has("pool")) {
      const std::string m = c->str("pool");
      p->set_pool(m == "AVE" ? PoolingParameter_PoolMethod_AVE
                  : m == "STOCHASTIC" ? PoolingParameter_PoolMethod_STOCHASTIC
                                      : PoolingParameter_PoolMethod_MAX);
    }
    if (c->has("pad")) p->set_pad((uint32_t)c->num("pad", 0));
    if (c->has("pad_h")) p->set_pad_h((uint32_t)c->num("pad_h", 0));
    if (c->has("pad_w")) p->set_pad_w((uint32_t)c->num("pad_w", 0));
    if (c->has("kernel_size")) p->set_kernel_size((uint32_t)c->num("kernel_size", 0));
    if (c->has("kernel_h")) p->set_kernel_h((uint32_t)c->num("kernel_h", 0));
    if (c->has("kernel_w")) p->set_kernel_w((uint32_t)c->num("kernel_w", 0));
    if (c->has("stride")) p->set_stride((uint32_t)c->num("stride", 1));
    if (c->has("stride_h")) p->set_stride_h((uint32_t)c->num("stride_h", 0));
    if (c->has("stride_w")) p->set_stride_w((uint32_t)c->num("stride_w", 0));
    if (c->has("global_pooling")) p->set_global_pooling(c->boolean("global_pooling", false));
  }
  if (const Node* c = n.child("inner_product_param")) {
    check_known(*c, "InnerProductParameter",
                {"num_output", "bias_term", "weight_filler", "bias_filler", "axis", "transpose"});
    InnerProductParameter* p = lp->mutable_inner_product_param();
    if (c->has("num_output")) p->set_num_output((uint32_t)c->num("num_output", 0));
    if (c->has("bias_term")) p->set_bias_term(c->boolean("bias_term", true));
    if (c->has("axis")) p->set_axis((int)c->num("axis", 1));
    if (c->has("transpose")) p->set_transpose(c->boolean("transpose", false));
    if (c->child("weight_filler")) fill(c->child("weight_filler"), p->mutable_weight_filler());
    if (c->child("bias_filler")) fill(c->child("bias_filler"), p->mutable_bias_filler());
  }
  if (const Node* c = n.child("input_param")) {
    check_known(*c, "InputParameter", {"shape"});
    for (const Field* f : c->all("shape"))
      fill(f->is_message ? f->message.get() : nullptr, lp->mutable_input_param()->add_shape());
  }
  if (const Node* c = n.child("dropout_param")) {
    check_known(*c, "DropoutParameter", {"dropout_ratio"});
    if (c->has("dropout_ratio")) lp->mutable_dropout_param()->set_dropout_ratio((float)c->num("dropout_ratio", 0.5));
  }
  if (const Node* c = n.child("concat_param")) {
    check_known(*c, "ConcatParameter", {"axis", "concat_dim"});
    if (c->has("axis")) lp->mutable_concat_param()->set_axis((int)c->num("axis", 1));
    if (c->has("concat_dim")) lp->mutable_concat_param()->set_concat_dim((uint32_t)c->num("concat_dim", 1));
  }
  if (const Node* c = n.child("relu_param")) {
    check_known(*c, "ReLUParameter", {"negative_slope", "engine"});
    if (c->has("negative_slope")) lp->mutable_relu_param()->set_negative_slope((float)c->num("negative_slope", 0));
  }
  if (const Node* c = n.child("roi_pooling_param")) {
    check_known(*c, "ROIPoolingParameter", {"pooled_h", "pooled_w", "spatial_scale", "pad_ratio"});
    ROIPoolingParameter* p = lp->mutable_roi_pooling_param();
    if (c->has("pooled_h")) p->set_pooled_h((uint32_t)c->num("pooled_h", 0));
    if (c->has("pooled_w")) p->set_pooled_w((uint32_t)c->num("pooled_w", 0));
    if (c->has("spatial_scale")) p->set_spatial_scale((float)c->num("spatial_scale", 1));
    if (c->has("pad_ratio")) p->set_pad_ratio((float)c->num("pad_ratio", 0));
  }
  if (const Node* c = n.child("box_output_param")) {
    check_known(*c, "BoxOutputParameter",
                {"fg_thr", "iou_thr", "nms_type", "field_h", "field_w", "downsample_rate", "field_whr",
                 "field_xyr", "max_nms_num", "max_post_nms_num", "min_size"});
    BoxOutputParameter* p = lp->mutable_box_output_param();
    if (c->has("fg_thr")) p->set_fg_thr((float)c->num("fg_thr", 0));
    if (c->has("iou_thr")) p->set_iou_thr((float)c->num("iou_thr", 0.5));
    if (c->has("nms_type")) p->set_nms_type(c->str("nms_type"));
    for (double v : c->nums("field_h")) p->add_field_h((uint32_t)v);
    for (double v : c->nums("field_w")) p->add_field_w((uint32_t)v);
    for (double v : c->nums("downsample_rate")) p->add_downsample_rate((uint32_t)v);
    if (c->has("field_whr")) p->set_field_whr((float)c->num("field_whr", 2));
    if (c->has("field_xyr")) p->set_field_xyr((float)c->num("field_xyr", 2));
    if (c->has("max_nms_num")) p->set_max_nms_num((uint32_t)c->num("max_nms_num", 0));
    if (c->has("max_post_nms_num")) p->set_max_post_nms_num((uint32_t)c->num("max_post_nms_num", 0));
    if (c->has("min_size")) p->set_min_size((float)c->num("min_size", 15));
  }
  if (const Node* c = n.child("bbox_reg_param")) {
    check_known(*c, "BBoxRegParameter", {"bbox_mean", "bbox_std", "cls_aware"});
    BBoxRegParameter* p = lp->mutable_bbox_reg_param();
    for (double v : c->nums("bbox_mean")) p->add_bbox_mean((float)v);
    for (double v : c->nums("bbox_std")) p->add_bbox_std((float)v);
    if (c->has("cls_aware")) p->set_cls_aware(c->boolean("cls_aware", true));
  }
  if (const Node* c = n.child("decode_bbox_param")) {
    check_known(*c, "DecodeBBoxParameter", {"gt_iou_thr"});
    if (c->has("gt_iou_thr")) lp->mutable_decode_bbox_param()->set_gt_iou_thr((float)c->num("gt_iou_thr", 0.95));
  }
  if (const Node* c = n.child("softmax_param")) {
    check_known(*c, "SoftmaxParameter", {"engine", "axis"});
    if (c->has("axis")) lp->mutable_softmax_param()->set_axis((int)c->num("axis", 1));
  }
  if (const Node* c = n.child("eltwise_param")) {
    check_known(*c, "EltwiseParameter", {"operation", "coeff", "stable_prod_grad"});
    EltwiseParameter* p = lp->mutable_eltwise_param();
    if (c->has("operation")) {
      const std::string m = c->str("operation");
      p->set_operation(m == "PROD" ? EltwiseParameter_EltwiseOp_PROD
                       : m == "MAX" ? EltwiseParameter_EltwiseOp_MAX
                                    : EltwiseParameter_EltwiseOp_SUM);
    }
    for (double v : c->nums("coeff")) p->add_coeff((float)v);
    if (c->has("stable_prod_grad")) p->set_stable_prod_grad(c->boolean("stable_prod_grad", true));
  }
}

inline void fill(const Node& root, NetParameter* np) {
  check_known(root, "NetParameter",
              {"name", "input", "input_shape", "input_dim", "force_backward", "state", "debug_info", "layer"});
  if (root.has("name")) np->set_name(root.str("name"));
  for (const std::string& s : root.strs("input")) np->add_input(s);
  for (const Field* f : root.all("input_shape")) fill(f->is_message ? f->message.get() : nullptr, np->add_input_shape());
  for (double d : root.nums("input_dim")) np->add_input_dim((int32_t)d);
  if (root.has("force_backward")) np->set_force_backward(root.boolean("force_backward", false));
  if (root.has("debug_info")) np->set_debug_info(root.boolean("debug_info", false));
  for (const Field* f : root.all("layer")) {
    if (!f->is_message) throw std::runtime_error("prototxt: 'layer' must be a message");
    fill(*f->message, np->add_layer());
  }
}

// ReadProtoFromTextFile / ReadNetParamsFromTextFileOrDie (util/io.cpp:34-44, upgrade_proto.cpp:1051)
inline void ReadNetParamsFromTextFile(const std::string& path, NetParameter* np) {
  fill(*prototxt::parse_file(path), np);
}
inline void ReadNetParamsFromString(const std::string& text, NetParameter* np) {
  fill(*prototxt::parse_string(text), np);
}

// UpgradeNetInput (util/upgrade_proto.cpp:966-1000): legacy `input:` / `input_dim:` /
// `input_shape` fields become an "Input" layer that is placed first.
inline void UpgradeNetInput(NetParameter* np) {
  if (np->input_size() == 0) return;
  NetParameter out;
  out.CopyFrom(*np);
  out.clear_layer();
  LayerParameter* lp = out.add_layer();
  lp->set_name("input");
  lp->set_type("Input");
  const bool has_shape = np->input_shape_size() > 0;
  for (int i = 0; i < np->input_size(); ++i) {
    lp->add_top(np->input(i));
    BlobShape* s = lp->mutable_input_param()->add_shape();
    if (has_shape) {
      for (int d = 0; d < np->input_shape(i).dim_size(); ++d) s->add_dim(np->input_shape(i).dim(d));
    } else {
      for (int d = 0; d < 4 && i * 4 + d < np->input_dim_size(); ++d) s->add_dim(np->input_dim(i * 4 + d));
    }
  }
  for (int i = 0; i < np->layer_size(); ++i) out.add_layer()->CopyFrom(np->layer(i));
  out.clear_input();
  out.clear_input_shape();
  out.clear_input_dim();
  np->CopyFrom(out);
}

}  // namespace proto_text
}  // namespace caffe
