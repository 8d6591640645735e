// caffe::Blob<Dtype> -- source-compatible subset of /root/reference/include/caffe/blob.hpp:24-277
// (forward path: no diff storage), plus the mscnn_b200 extension that lets the hot path keep a
// tensor as NHWC bf16 "planes" on the device without ever materialising the NCHW fp32 array.
//
// A blob has up to two device representations of the same logical N x C x H x W tensor:
//   * the Caffe one: NCHW fp32 in a SyncedMemory (cpu_data()/gpu_data());
//   * planes: NHWC bf16, channels padded to 64, hi (+ lo in fp32-faithful mode).
// `layout_head_` says which one is current.  cpu_data()/gpu_data() convert from planes on
// demand (mscnn_planes_to_nchw_f32), planes() converts from NCHW on demand
// (mscnn_nchw_f32_to_planes), so user code that inspects any blob by name sees ordinary Caffe
// data while layers hand planes to each other with no conversion.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"
#include "caffe/syncedmem.hpp"

const int kMaxBlobAxes = 32;

namespace caffe {

// Device buffer pair for the planes representation (grow-only, like Blob::Reshape).
struct CAFFE_API PlaneStore {
  void* hi = nullptr;
  void* lo = nullptr;
  size_t capacity = 0;     // bytes per plane
  size_t lo_capacity = 0;
  ~PlaneStore();
  void reserve(size_t bytes, bool need_lo);
};

template <typename Dtype>
class CAFFE_API Blob {
 public:
  Blob() : data_(), count_(0), capacity_(0), layout_head_(HEAD_NCHW), planes_split_(false), version_(0) {}
  explicit Blob(const int num, const int channels, const int height, const int width);
  explicit Blob(const vector<int>& shape);

  void Reshape(const int num, const int channels, const int height, const int width);
  void Reshape(const vector<int>& shape);
  void Reshape(const BlobShape& shape);
  void ReshapeLike(const Blob& other);
  inline string shape_string() const {
    std::ostringstream stream;
    for (size_t i = 0; i < shape_.size(); ++i) stream << shape_[i] << " ";
    stream << "(" << count_ << ")";
    return stream.str();
  }
  inline const vector<int>& shape() const { return shape_; }
  inline int shape(int index) const { return shape_[CanonicalAxisIndex(index)]; }
  inline int num_axes() const { return (int)shape_.size(); }
  inline int count() const { return count_; }
  inline int count(int start_axis, int end_axis) const {
    CHECK_LE(start_axis, end_axis);
    CHECK_GE(start_axis, 0);
    CHECK_LE(end_axis, num_axes());
    int count = 1;
    for (int i = start_axis; i < end_axis; ++i) count *= shape(i);
    return count;
  }
  inline int count(int start_axis) const { return count(start_axis, num_axes()); }
  inline int CanonicalAxisIndex(int axis_index) const {
    CHECK_GE(axis_index, -num_axes()) << "axis " << axis_index << " out of range";
    CHECK_LT(axis_index, num_axes()) << "axis " << axis_index << " out of range";
    if (axis_index < 0) return axis_index + num_axes();
    return axis_index;
  }
  inline int num() const { return LegacyShape(0); }
  inline int channels() const { return LegacyShape(1); }
  inline int height() const { return LegacyShape(2); }
  inline int width() const { return LegacyShape(3); }
  inline int LegacyShape(int index) const {
    CHECK_LE(num_axes(), 4) << "Cannot use legacy accessors on Blobs with > 4 axes.";
    CHECK_LT(index, 4);
    CHECK_GE(index, -4);
    if (index >= num_axes() || index < -num_axes()) return 1;
    return shape(index);
  }
  inline int offset(const int n, const int c = 0, const int h = 0, const int w = 0) const {
    return ((n * channels() + c) * height() + h) * width() + w;
  }
  void CopyFrom(const Blob<Dtype>& source, bool copy_diff = false, bool reshape = false);
  inline Dtype data_at(const int n, const int c, const int h, const int w) const {
    return const_cast<Blob*>(this)->cpu_data()[offset(n, c, h, w)];
  }

  const Dtype* cpu_data();
  void set_cpu_data(Dtype* data);
  const Dtype* gpu_data();
  Dtype* mutable_cpu_data();
  Dtype* mutable_gpu_data();
  void FromProto(const BlobProto& proto, bool reshape = true);
  void ToProto(BlobProto* proto, bool write_diff = false);
  Dtype asum_data();
  Dtype sumsq_data();
  void ShareData(const Blob& other);
  bool ShapeEquals(const BlobProto& other);

  // ---- mscnn_b200 extension: planes -----------------------------------------------------
  struct Planes {
    void* hi;
    void* lo;   // NULL on the plain-bf16 path
    int n, h, w, cpad;
  };
  // Current contents as planes (converted from NCHW fp32 if that is where the head is).
  Planes planes(bool split);
  // Storage for a producer to fill: sets the head to planes and invalidates the NCHW copy.
  Planes mutable_planes(bool split);
  bool head_is_planes() const { return layout_head_ == HEAD_PLANES; }
  // Bumped by every mutable_* accessor; layers use it to re-pack cached weights lazily.
  unsigned long version() const { return version_; }

 protected:
  enum LayoutHead { HEAD_NCHW, HEAD_PLANES, HEAD_BOTH };
  void planes_dims(int* n, int* c, int* h, int* w) const;
  void sync_to_nchw();
  shared_ptr<SyncedMemory> data_;
  shared_ptr<PlaneStore> planes_;
  vector<int> shape_;
  int count_;
  int capacity_;
  LayoutHead layout_head_;
  bool planes_split_;
  unsigned long version_;
  DISABLE_COPY_AND_ASSIGN(Blob);
};

}  // namespace caffe
