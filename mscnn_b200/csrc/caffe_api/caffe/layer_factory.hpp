// LayerRegistry / REGISTER_LAYER_CLASS -- same interface as
// /root/reference/include/caffe/layer_factory.hpp:56-137 (type string -> creator).
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

  // Defined once, inside the library (caffe_api.cu): layers registered by the host program and the library's own
  // layers must meet in ONE table (an inline function-local static would exist once per binary).
  static CreatorRegistry& Registry();
  static void AddCreator(const string& type, Creator creator) {
    CreatorRegistry& registry = Registry();
    CHECK_EQ(registry.count(type), 0) << "Layer type " << type << " already registered.";
    registry[type] = creator;
  }
  static shared_ptr<Layer<Dtype> > CreateLayer(const LayerParameter& param) {
    const string& type = param.type();
    CreatorRegistry& registry = Registry();
    CHECK_EQ(registry.count(type), 1) << "Unknown layer type: " << type
                                      << " (known types: " << LayerTypeListString() << ")";
    return registry[type](param);
  }
  static vector<string> LayerTypeList() {
    vector<string> layer_types;
    for (typename CreatorRegistry::iterator it = Registry().begin(); it != Registry().end(); ++it)
      layer_types.push_back(it->first);
    return layer_types;
  }

 private:
  LayerRegistry() {}
  static string LayerTypeListString() {
    string s;
    for (const string& t : LayerTypeList()) s += (s.empty() ? "" : ", ") + t;
    return s;
  }
};

template <typename Dtype>
class CAFFE_API LayerRegisterer {
 public:
  LayerRegisterer(const string& type, shared_ptr<Layer<Dtype> > (*creator)(const LayerParameter&)) {
    LayerRegistry<Dtype>::AddCreator(type, creator);
  }
};

#define REGISTER_LAYER_CREATOR(type, creator) \
  static LayerRegisterer<float> g_creator_f_##type(#type, creator<float>)

#define REGISTER_LAYER_CLASS(type)                                              \
  template <typename Dtype>                                                     \
  shared_ptr<Layer<Dtype> > Creator_##type##Layer(const LayerParameter& param) { \
    return shared_ptr<Layer<Dtype> >(new type##Layer<Dtype>(param));            \
  }                                                                             \
  REGISTER_LAYER_CREATOR(type, Creator_##type##Layer)

}  // namespace caffe
