// A C++ host written against the Caffe API only -- the calls tools/caffe.cpp, matcaffe (matlab/+caffe/private/caffe_.cpp)
// and pycaffe make on caffe::Net (/root/reference/include/caffe/net.hpp:23-120) -- compiled with plain g++ against the
// Caffe-API mirror headers and linked against libmscnn_b200.so instead of libcaffe.so:
//
//   g++ -std=c++17 -I include -I mscnn_b200/csrc/caffe_api -I mscnn_b200/csrc/proto_shared -I /usr/local/cuda/include \
//       examples/caffe_driver.cpp -L mscnn_b200 -lmscnn_b200 -Wl,-rpath,$PWD/mscnn_b200 -o caffe_driver
//   ./caffe_driver mscnn_deploy.prototxt              # structure only (works without a GPU)
//   ./caffe_driver mscnn_deploy.prototxt --forward    # fill parameters and input deterministically, Forward(), checksums
//
// It also registers a layer type of its own with REGISTER_LAYER_CLASS to show that host code and library share one
// LayerRegistry (tests/test_cpp_driver.py).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "caffe/caffe.hpp"

namespace caffe {
// A host-side layer: top = bottom (shares the data), like the reference's SplitLayer with one top.
template <typename Dtype>
class HostPassLayer : public Layer<Dtype> {
 public:
  explicit HostPassLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    top[0]->ReshapeLike(*bottom[0]);
    top[0]->ShareData(*bottom[0]);
  }
  virtual inline const char* type() const { return "HostPass"; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {}
};
REGISTER_LAYER_CLASS(HostPass);
}  // namespace caffe

using namespace caffe;

// Deterministic integer-hash fill in [-0.5, 0.5) * scale (the same formula in tests/test_cpp_driver.py).
static void fill(float* p, int count, unsigned salt, float scale) {
  for (int e = 0; e < count; ++e) {
    const unsigned h = (static_cast<unsigned>(e) * 2654435761u + salt * 40503u) >> 8;
    p[e] = (static_cast<float>(h & 0xFFFFu) / 65536.0f - 0.5f) * scale;
  }
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s deploy.prototxt [--forward]\n", argv[0]);
    return 2;
  }
  const bool forward = argc > 2 && !std::strcmp(argv[2], "--forward");
  Caffe::set_mode(Caffe::GPU);
  Net<float> net(argv[1], TEST);
  std::printf("net %s: %zu layers, %zu blobs, %d inputs, %d outputs\n", net.name().c_str(), net.layers().size(),
              net.blobs().size(), net.num_inputs(), net.num_outputs());
  for (size_t i = 0; i < net.layers().size(); ++i)
    std::printf("layer %zu %s %s params=%zu\n", i, net.layer_names()[i].c_str(), net.layers()[i]->type(),
                net.layers()[i]->blobs().size());
  const std::vector<std::string> types = LayerRegistry<float>::LayerTypeList();
  bool seen_host = false, seen_lib = false;
  for (const std::string& t : types) {
    seen_host |= (t == "HostPass");
    seen_lib |= (t == "BoxOutput");
  }
  std::printf("registry: %zu types, HostPass=%d BoxOutput=%d\n", types.size(), (int)seen_host, (int)seen_lib);
  if (!forward) return (seen_host && seen_lib) ? 0 : 1;

  for (size_t i = 0; i < net.layers().size(); ++i) {
    vector<shared_ptr<Blob<float> > >& blobs = net.layers()[i]->blobs();
    for (size_t j = 0; j < blobs.size(); ++j) {
      const int fan = blobs[j]->count() / blobs[j]->shape(0);
      const float scale = j == 0 ? 3.4641f / std::sqrt(static_cast<float>(fan > 0 ? fan : 1)) : 0.2f;
      fill(blobs[j]->mutable_cpu_data(), blobs[j]->count(), static_cast<unsigned>(i * 8 + j), scale);
    }
  }
  Blob<float>* in = net.input_blobs()[0];
  fill(in->mutable_cpu_data(), in->count(), 9999u, 200.0f);
  const vector<Blob<float>*>& out = net.Forward();
  for (size_t k = 0; k < out.size(); ++k) {
    const float* v = out[k]->cpu_data();
    double s = 0, a = 0;
    for (int e = 0; e < out[k]->count(); ++e) {
      s += v[e];
      a += std::fabs(v[e]);
    }
    std::printf("output %s shape %s sum %.9e abs %.9e\n", net.blob_names()[net.output_blob_indices()[k]].c_str(),
                out[k]->shape_string().c_str(), s, a);
  }
  return 0;
}
