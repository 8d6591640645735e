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
#include "caffe/layers/softmax_layer.hpp"
#include "caffe/proto/caffe.pb.h"
#include "cblas.h"
#include "proto_text.hpp"
#include "prototxt.hpp"

namespace caffe {
// The reference registers these four through layer_factory.cpp (engine dispatch to cuDNN,
// layer_factory.cpp:36-74,76-112,152-174,199-221); with CPU_ONLY the dispatch always lands on the
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
template <typename Dtype>
shared_ptr<Layer<Dtype> > MscnnRefGetSoftmaxLayer(const LayerParameter& p) {
  return shared_ptr<Layer<Dtype> >(new SoftmaxLayer<Dtype>(p));
}
REGISTER_LAYER_CREATOR(ReLU, MscnnRefGetReLULayer);
REGISTER_LAYER_CREATOR(Softmax, MscnnRefGetSoftmaxLayer);
}  // namespace caffe

namespace {

using caffe::Blob;
using caffe::Layer;
using caffe::LayerParameter;
using prototxt::Node;

struct RefLayer {
  LayerParameter param;
  boost::shared_ptr<Layer<float> > layer;
  std::vector<Blob<float>*> bottom, top;
};

struct RefNet {
  std::vector<RefLayer> layers;
  std::map<std::string, std::shared_ptr<Blob<float> > > blobs;
  std::map<std::string, Blob<float>*> shared_params;
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
    caffe::NetParameter np;
    caffe::proto_text::fill(root, &np);
    caffe::proto_text::UpgradeNetInput(&np);  // upgrade_proto.cpp:966-1000 semantics
    std::vector<LayerParameter> params;
    for (int i = 0; i < np.layer_size(); ++i) {
      params.push_back(np.layer(i));
      params.back().set_phase(caffe::TEST);
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
      // Net::AppendParam (net.cpp:448-538): same non-empty ParamSpec name => ShareData with the owner
      for (int j = 0; j < (int)L.layer->blobs().size() && j < L.param.param_size(); ++j) {
        const std::string& pname = L.param.param(j).name();
        if (pname.empty()) continue;
        auto owner = shared_params.find(pname);
        if (owner == shared_params.end()) {
          shared_params[pname] = L.layer->blobs()[j].get();
        } else {
          CHECK(owner->second->shape() == L.layer->blobs()[j]->shape()) << "shared param shape mismatch " << pname;
          L.layer->blobs()[j]->ShareData(*owner->second);
        }
      }
    }
  }
};

}  // namespace

#pragma GCC visibility push(default)
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
#pragma GCC visibility pop
