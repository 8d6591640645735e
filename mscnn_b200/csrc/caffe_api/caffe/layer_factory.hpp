// Layer-type registry with the interface of /root/reference/include/caffe/layer_factory.hpp:56-137: a type string maps
// to a creator function; REGISTER_LAYER_CLASS / REGISTER_LAYER_CREATOR add entries at static-initialisation time and
// Net::Init asks CreateLayer for every LayerParameter.
//
// Unlike the reference (header-only, one table per binary that instantiates it), the table and its accessors live
// INSIDE libmscnn_b200.so (caffe_api.cu) and are exported: layer types registered by a host program and the library's
// own layers meet in one table (examples/caffe_driver.cpp registers "HostPass").
#pragma once
#include <map>
#include <string>
#include <vector>

#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

template <typename Dtype>
class Layer;

template <typename Dtype>
class CAFFE_API LayerRegistry {
 public:
  typedef shared_ptr<Layer<Dtype> > (*Creator)(const LayerParameter&);
  typedef std::map<string, Creator> CreatorRegistry;

  static CreatorRegistry& Registry();
  static void AddCreator(const string& type, Creator creator);                  // aborts on a duplicate type
  static shared_ptr<Layer<Dtype> > CreateLayer(const LayerParameter& param);    // aborts on an unknown type
  static vector<string> LayerTypeList();

 private:
  LayerRegistry();  // static interface only
};

// Registers a creator from its constructor (one static object per registered type).
template <typename Dtype>
struct CAFFE_API LayerRegisterer {
  LayerRegisterer(const string& type, typename LayerRegistry<Dtype>::Creator creator);
};

namespace detail {
// creator of a class template `SomeLayer<Dtype>` with the usual (const LayerParameter&) constructor
template <template <typename> class LayerT, typename Dtype>
shared_ptr<Layer<Dtype> > make_layer(const LayerParameter& param) {
  return shared_ptr<Layer<Dtype> >(new LayerT<Dtype>(param));
}
}  // namespace detail

// mscnn_b200 instantiates float only (see INSTANTIATE_CLASS in common.hpp).
#define REGISTER_LAYER_CREATOR(type, creator) \
  static ::caffe::LayerRegisterer<float> g_creator_f_##type(#type, creator<float>)

#define REGISTER_LAYER_CLASS(type) \
  static ::caffe::LayerRegisterer<float> g_creator_f_##type(#type, &::caffe::detail::make_layer<type##Layer, float>)

}  // namespace caffe
