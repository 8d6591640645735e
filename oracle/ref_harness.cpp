// oracle/_ref harness -- TEST INFRASTRUCTURE, never linked into or called by the product.
//
// Drives the reference's own layer implementations, compiled VERBATIM from
// /root/reference/src/caffe (see oracle/build_ref.py for the file list) against the shim
// headers in oracle/ref_shim/, through a small C ABI that tests / bench.py's cpu_baseline leg
// load with ctypes.  What is reference code: every Layer<float>::{LayerSetUp,Reshape,Forward_cpu},
// Blob, SyncedMemory, im2col, caffe_cpu_gemm, BoxIOU.  What is NOT reference code: this graph
// walker (the reference's net.cpp needs protobuf reflection, HDF5 and upgrade_proto and is not
// compiled); it wires blobs by name exactly as Net::Init does for a linear deploy net
// (net.cpp:384-446: in-place tops reuse the bottom blob, other tops get new blobs) and runs
// layers in file order like Net::ForwardFromTo (net.cpp:544-555).  Split layers are not
// inserted: for the forward pass they only share data (split_layer.cpp:26-31).
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/layer.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/layers/conv_layer.hpp"
#include "caffe/layers/pooling_layer.hpp"
#include "caffe/layers/relu_layer.hpp"
#include "caffe/proto/caffe.pb.h"
#include "cblas.h"
#include "prototxt.hpp"

namespace caffe {
// The reference registers these three through layer_factory.cpp (engine dispatch to cuDNN,
// layer_factory.cpp:36-74,76-112,152-174); with CPU_ONLY the dispatch always lands on the
// Caffe-engine classes, which is what these creators return.
template <typename Dtype>
shared_ptr<Layer<Dtype> > MscnnRefGetConvolutionLayer(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new ConvolutionLayer<Dtype>(p));
}
template <typename Dtype>
shared_ptr<Layer<Dtype> > MscnnRefGetPoolingLayer(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new PoolingLayer<Dtype>(p));
}
template <typename Dtype>
shared_ptr<Layer<Dtype> > MscnnRefGetReLULayer(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new ReLULayer<Dtype>(p));
}
REGISTER_LAYER_CREATOR(Convolution, MscnnRefGetConvolutionLayer);
REGISTER_LAYER_CREATOR(Pooling, MscnnRefGetPoolingLayer);
REGISTER_LAYER_CREATOR(ReLU, MscnnRefGetReLULayer);
}  // namespace caffe

namespace {

using caffe::Blob;
using caffe::Layer;
using caffe::LayerParameter;
using prototxt::Node;

void fill_filler(const Node* n, caffe::FillerParameter* f) {
  if (!n) return;
  if (n->has("type")) f->set_type(n->str("type"));
  if (n->has("value")) f->set_value((float)n->num("value", 0));
  if (n->has("min")) f->set_min((float)n->num("min", 0));
  if (n->has("max")) f->set_max((float)n->num("max", 1));
  if (n->has("mean")) f->set_mean((float)n->num("mean", 0));
  if (n->has("std")) f->set_std((float)n->num("std", 1));
  if (n->has("sparse")) f->set_sparse((int)n->num("sparse", -1));
}

void fill_layer_param(const Node& n, LayerParameter* lp) {
  lp->set_name(n.str("name"));
  lp->set_type(n.str("type"));
  for (const std::string& s : n.strs("bottom")) lp->add_bottom(s);
  for (const std::string& s : n.strs("top")) lp->add_top(s);
  lp->set_phase(caffe::TEST);
  for (const prototxt::Field* f : n.all("param")) {
    caffe::ParamSpec* ps = lp->add_param();
    if (f->is_message) {
      if (f->message->has("lr_mult")) ps->set_lr_mult((float)f->message->num("lr_mult", 1));
      if (f->message->has("decay_mult")) ps->set_decay_mult((float)f->message->num("decay_mult", 1));
    }
  }
  if (const Node* c = n.child("convolution_param")) {
    caffe::ConvolutionParameter* p = lp->mutable_convolution_param();
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
    if (c->child("weight_filler")) fill_filler(c->child("weight_filler"), p->mutable_weight_filler());
    if (c->child("bias_filler")) fill_filler(c->child("bias_filler"), p->mutable_bias_filler());
  }
  if (const Node* c = n.child("pooling_param")) {
    caffe::PoolingParameter* p = lp->mutable_pooling_param();
    if (c->has("pool")) {
      const std::string m = c->str("pool");
      p->set_pool(m == "AVE" ? caffe::PoolingParameter_PoolMethod_AVE
                  : m == "STOCHASTIC" ? caffe::PoolingParameter_PoolMethod_STOCHASTIC
                                      : caffe::PoolingParameter_PoolMethod_MAX);
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
    caffe::InnerProductParameter* p = lp->mutable_inner_product_param();
    if (c->has("num_output")) p->set_num_output((uint32_t)c->num("num_output", 0));
    if (c->has("bias_term")) p->set_bias_term(c->boolean("bias_term", true));
    if (c->has("axis")) p->set_axis((int)c->num("axis", 1));
    if (c->has("transpose")) p->set_transpose(c->boolean("transpose", false));
    if (c->child("weight_filler")) fill_filler(c->child("weight_filler"), p->mutable_weight_filler());
    if (c->child("bias_filler")) fill_filler(c->child("bias_filler"), p->mutable_bias_filler());
  }
  if (const Node* c = n.child("input_param")) {
    for (const prototxt::Field* f : c->all("shape")) {
      caffe::BlobShape* s = lp->mutable_input_param()->add_shape();
      if (f->is_message)
        for (double d : f->message->nums("dim")) s->add_dim((int64_t)d);
    }
  }
  if (const Node* c = n.child("dropout_param")) {
    if (c->has("dropout_ratio"))
      lp->mutable_dropout_param()->set_dropout_ratio((float)c->num("dropout_ratio", 0.5));
  }
  if (const Node* c = n.child("concat_param")) {
    if (c->has("axis")) lp->mutable_concat_param()->set_axis((int)c->num("axis", 1));
    if (c->has("concat_dim")) lp->mutable_concat_param()->set_concat_dim((uint32_t)c->num("concat_dim", 1));
  }
  if (const Node* c = n.child("relu_param")) {
    if (c->has("negative_slope"))
      lp->mutable_relu_param()->set_negative_slope((float)c->num("negative_slope", 0));
  }
  if (const Node* c = n.child("roi_pooling_param")) {
    caffe::ROIPoolingParameter* p = lp->mutable_roi_pooling_param();
    if (c->has("pooled_h")) p->set_pooled_h((uint32_t)c->num("pooled_h", 0));
    if (c->has("pooled_w")) p->set_pooled_w((uint32_t)c->num("pooled_w", 0));
    if (c->has("spatial_scale")) p->set_spatial_scale((float)c->num("spatial_scale", 1));
    if (c->has("pad_ratio")) p->set_pad_ratio((float)c->num("pad_ratio", 0));
  }
  if (const Node* c = n.child("box_output_param")) {
    caffe::BoxOutputParameter* p = lp->mutable_box_output_param();
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
    caffe::BBoxRegParameter* p = lp->mutable_bbox_reg_param();
    for (double v : c->nums("bbox_mean")) p->add_bbox_mean((float)v);
    for (double v : c->nums("bbox_std")) p->add_bbox_std((float)v);
  }
}

struct RefLayer {
  LayerParameter param;
  boost::shared_ptr<Layer<float> > layer;
  std::vector<Blob<float>*> bottom, top;
};

struct RefNet {
  std::vector<RefLayer> layers;
  std::map<std::string, std::shared_ptr<Blob<float> > > blobs;
  std::string error;

  Blob<float>* blob(const std::string& name) {
    auto it = blobs.find(name);
    return it == blobs.end() ? nullptr : it->second.get();
  }
  int layer_index(const std::string& name) const {
    for (size_t i = 0; i < layers.size(); ++i)
      if (layers[i].param.name() == name) return (int)i;
    return -1;
  }

  void build(const Node& root, int n_override) {
    std::vector<LayerParameter> params;
    // legacy `input:` + `input_dim:` x4 -> Input layer (upgrade_proto.cpp:966-1000)
    const std::vector<std::string> inputs = root.strs("input");
    if (!inputs.empty()) {
      const std::vector<double> dims = root.nums("input_dim");
      LayerParameter lp;
      lp.set_name("input");
      lp.set_type("Input");
      lp.set_phase(caffe::TEST);
      for (size_t i = 0; i < inputs.size(); ++i) {
        lp.add_top(inputs[i]);
        caffe::BlobShape* s = lp.mutable_input_param()->add_shape();
        for (int d = 0; d < 4 && i * 4 + d < dims.size(); ++d) s->add_dim((int64_t)dims[i * 4 + d]);
      }
      for (const prototxt::Field* f : root.all("input_shape")) {
        caffe::BlobShape* s = lp.mutable_input_param()->add_shape();
        if (f->is_message)
          for (double d : f->message->nums("dim")) s->add_dim((int64_t)d);
      }
      params.push_back(lp);
    }
    for (const prototxt::Field* f : root.all("layer")) {
      if (!f->is_message) continue;
      LayerParameter lp;
      fill_layer_param(*f->message, &lp);
      params.push_back(lp);
    }
    if (n_override > 0) {
      for (LayerParameter& lp : params) {
        if (lp.type() != "Input") continue;
        for (int i = 0; i < lp.input_param().shape_size(); ++i) {
          caffe::BlobShape* s = lp.mutable_input_param()->mutable_shape(i);
          if (s->dim_size() > 0) s->set_dim(0, n_override);
        }
      }
    }
    layers.resize(params.size());
    for (size_t i = 0; i < params.size(); ++i) {
      RefLayer& L = layers[i];
      L.param = params[i];
      for (int b = 0; b < L.param.bottom_size(); ++b) {
        Blob<float>* bl = blob(L.param.bottom(b));
        CHECK(bl != nullptr) << "unknown bottom blob " << L.param.bottom(b) << " in layer "
                             << L.param.name();
        L.bottom.push_back(bl);
      }
      for (int t = 0; t < L.param.top_size(); ++t) {
        const std::string& tn = L.param.top(t);
        const bool in_place = t < L.param.bottom_size() && L.param.bottom(t) == tn;
        if (!in_place) blobs[tn] = std::make_shared<Blob<float> >();
        L.top.push_back(blob(tn));
      }
      L.layer = caffe::LayerRegistry<float>::CreateLayer(L.param);
      L.layer->SetUp(L.bottom, L.top);
    }
  }
};

}  // namespace

extern "C" {

void* mscnn_ref_net_create(const char* prototxt_or_path, int is_path, int n_override) {
  try {
    caffe::Caffe::set_mode(caffe::Caffe::CPU);
    std::shared_ptr<Node> root = is_path ? prototxt::parse_file(prototxt_or_path)
                                         : prototxt::parse_string(prototxt_or_path);
    RefNet* net = new RefNet();
    net->build(*root, n_override);
    return net;
  } catch (const std::exception& e) {
    fprintf(stderr, "mscnn_ref_net_create: %s\n", e.what());
    return nullptr;
  }
}
void mscnn_ref_net_destroy(void* h) { delete static_cast<RefNet*>(h); }
int mscnn_ref_net_num_layers(void* h) { return (int)static_cast<RefNet*>(h)->layers.size(); }
const char* mscnn_ref_net_layer_name(void* h, int i) {
  return static_cast<RefNet*>(h)->layers[i].param.name().c_str();
}
const char* mscnn_ref_net_layer_type(void* h, int i) {
  return static_cast<RefNet*>(h)->layers[i].param.type().c_str();
}
int mscnn_ref_net_num_params(void* h, const char* layer) {
  RefNet* n = static_cast<RefNet*>(h);
  const int i = n->layer_index(layer);
  return i < 0 ? -1 : (int)n->layers[i].layer->blobs().size();
}
// shape4 receives up to 4 dims; returns the number of axes (or -1)
int mscnn_ref_net_param_shape(void* h, const char* layer, int idx, int* shape4) {
  RefNet* n = static_cast<RefNet*>(h);
  const int i = n->layer_index(layer);
  if (i < 0 || idx >= (int)n->layers[i].layer->blobs().size()) return -1;
  const std::vector<int>& s = n->layers[i].layer->blobs()[idx]->shape();
  for (size_t d = 0; d < s.size() && d < 4; ++d) shape4[d] = s[d];
  return (int)s.size();
}
int mscnn_ref_net_set_param(void* h, const char* layer, int idx, const float* data, long count) {
  RefNet* n = static_cast<RefNet*>(h);
  const int i = n->layer_index(layer);
  if (i < 0 || idx >= (int)n->layers[i].layer->blobs().size()) return -1;
  Blob<float>* b = n->layers[i].layer->blobs()[idx].get();
  if (b->count() != count) return -2;
  memcpy(b->mutable_cpu_data(), data, sizeof(float) * count);
  return 0;
}
int mscnn_ref_net_blob_shape(void* h, const char* name, int* shape4) {
  Blob<float>* b = static_cast<RefNet*>(h)->blob(name);
  if (!b) return -1;
  const std::vector<int>& s = b->shape();
  for (size_t d = 0; d < s.size() && d < 4; ++d) shape4[d] = s[d];
  return (int)s.size();
}
int mscnn_ref_net_reshape_blob(void* h, const char* name, int n, int c, int hh, int w) {
  Blob<float>* b = static_cast<RefNet*>(h)->blob(name);
  if (!b) return -1;
  b->Reshape(n, c, hh, w);
  return 0;
}
int mscnn_ref_net_set_blob(void* h, const char* name, const float* data, long count) {
  Blob<float>* b = static_cast<RefNet*>(h)->blob(name);
  if (!b) return -1;
  if (b->count() != count) return -2;
  memcpy(b->mutable_cpu_data(), data, sizeof(float) * count);
  return 0;
}
int mscnn_ref_net_get_blob(void* h, const char* name, float* out, long count) {
  Blob<float>* b = static_cast<RefNet*>(h)->blob(name);
  if (!b) return -1;
  if (b->count() != count) return -2;
  memcpy(out, b->cpu_data(), sizeof(float) * count);
  return 0;
}
// Forward layers [from, to] (inclusive, -1 = last) in file order; per_layer_ms may be NULL.
// Timing protocol = `caffe time` forward loop (tools/caffe.cpp:380-389): wall clock around
// each Layer::Forward.  Returns the total in milliseconds.
double mscnn_ref_net_forward(void* h, int from, int to, double* per_layer_ms) {
  RefNet* n = static_cast<RefNet*>(h);
  if (to < 0 || to >= (int)n->layers.size()) to = (int)n->layers.size() - 1;
  if (from < 0) from = 0;
  double total = 0;
  for (int i = from; i <= to; ++i) {
    RefLayer& L = n->layers[i];
    const auto t0 = std::chrono::steady_clock::now();
    L.layer->Forward(L.bottom, L.top);
    const auto t1 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (per_layer_ms) per_layer_ms[i] = ms;
    total += ms;
  }
  return total;
}
}
