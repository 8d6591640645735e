// Hand-written equivalent of the protoc-generated caffe.pb.h (no protoc / libprotobuf in the
// image).  It declares, with the field names, types and DEFAULTS of
// /root/reference/src/caffe/proto/caffe.proto, the messages and accessors that the MS-CNN
// deploy nets and the hot-path layers use (census in SURVEY.md section 8(c)), with the same
// generated-code API surface (name(), has_name(), set_name(), name_size(), add_name(),
// mutable_name(), clear_name()).  Used by the Caffe-API mirror (csrc/caffe_api) and, so that the
// verbatim reference build sees identical parameters, by oracle/_ref.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace caffe {

#define PB_OPT(type, name, dflt)                                  \
 private:                                                         \
  type name##_ = dflt;                                            \
  bool has_##name##_ = false;                                     \
                                                                  \
 public:                                                          \
  type name() const { return name##_; }                           \
  bool has_##name() const { return has_##name##_; }               \
  void set_##name(type v) { name##_ = v; has_##name##_ = true; }  \
  void clear_##name() { name##_ = dflt; has_##name##_ = false; }

#define PB_STR(name, dflt)                                                     \
 private:                                                                      \
  std::string name##_ = dflt;                                                  \
  bool has_##name##_ = false;                                                  \
                                                                               \
 public:                                                                       \
  const std::string& name() const { return name##_; }                          \
  bool has_##name() const { return has_##name##_; }                            \
  void set_##name(const std::string& v) { name##_ = v; has_##name##_ = true; } \
  std::string* mutable_##name() { has_##name##_ = true; return &name##_; }

template <typename T>
class RepeatedField : public std::vector<T> {
 public:
  const T& Get(int i) const { return (*this)[i]; }
  void Add(const T& v) { this->push_back(v); }
  void Clear() { this->clear(); }
};

#define PB_REP(type, name)                                        \
 private:                                                         \
  RepeatedField<type> name##_;                                    \
                                                                  \
 public:                                                          \
  int name##_size() const { return (int)name##_.size(); }         \
  type name(int i) const { return name##_[i]; }                   \
  const RepeatedField<type>& name() const { return name##_; }     \
  RepeatedField<type>* mutable_##name() { return &name##_; }      \
  void add_##name(type v) { name##_.push_back(v); }               \
  void set_##name(int i, type v) { name##_[i] = v; }              \
  void clear_##name() { name##_.clear(); }

#define PB_REP_STR(name)                                                  \
 private:                                                                 \
  RepeatedField<std::string> name##_;                                     \
                                                                          \
 public:                                                                  \
  int name##_size() const { return (int)name##_.size(); }                 \
  const std::string& name(int i) const { return name##_[i]; }             \
  const RepeatedField<std::string>& name() const { return name##_; }      \
  void add_##name(const std::string& v) { name##_.push_back(v); }         \
  void set_##name(int i, const std::string& v) { name##_[i] = v; }        \
  void clear_##name() { name##_.clear(); }

#define PB_MSG(Type, name)                                              \
 private:                                                               \
  Type name##_;                                                         \
  bool has_##name##_ = false;                                           \
                                                                        \
 public:                                                                \
  const Type& name() const { return name##_; }                          \
  bool has_##name() const { return has_##name##_; }                     \
  Type* mutable_##name() { has_##name##_ = true; return &name##_; }     \
  void clear_##name() { name##_ = Type(); has_##name##_ = false; }

#define PB_REP_MSG(Type, name)                                          \
 private:                                                               \
  std::vector<Type> name##_;                                            \
                                                                        \
 public:                                                                \
  int name##_size() const { return (int)name##_.size(); }               \
  const Type& name(int i) const { return name##_[i]; }                  \
  Type* mutable_##name(int i) { return &name##_[i]; }                   \
  const std::vector<Type>& name() const { return name##_; }             \
  Type* add_##name() { name##_.emplace_back(); return &name##_.back(); } \
  void clear_##name() { name##_.clear(); }

enum Phase { TRAIN = 0, TEST = 1 };

class BlobShape {  // caffe.proto:6-8
  PB_REP(int64_t, dim)
};

class BlobProto {  // caffe.proto:10-22
  PB_MSG(BlobShape, shape)
  PB_REP(float, data)
  PB_REP(float, diff)
  PB_REP(double, double_data)
  PB_REP(double, double_diff)
  PB_OPT(int32_t, num, 0)
  PB_OPT(int32_t, channels, 0)
  PB_OPT(int32_t, height, 0)
  PB_OPT(int32_t, width, 0)
};

enum FillerParameter_VarianceNorm {
  FillerParameter_VarianceNorm_FAN_IN = 0,
  FillerParameter_VarianceNorm_FAN_OUT = 1,
  FillerParameter_VarianceNorm_AVERAGE = 2
};
class FillerParameter {  // caffe.proto:43-62
 public:
  typedef FillerParameter_VarianceNorm VarianceNorm;
  static const VarianceNorm FAN_IN = FillerParameter_VarianceNorm_FAN_IN;
  static const VarianceNorm FAN_OUT = FillerParameter_VarianceNorm_FAN_OUT;
  static const VarianceNorm AVERAGE = FillerParameter_VarianceNorm_AVERAGE;
  PB_STR(type, "constant")
  PB_OPT(float, value, 0.f)
  PB_OPT(float, min, 0.f)
  PB_OPT(float, max, 1.f)
  PB_OPT(float, mean, 0.f)
  PB_OPT(float, std, 1.f)
  PB_OPT(int32_t, sparse, -1)
  PB_OPT(FillerParameter_VarianceNorm, variance_norm, FillerParameter_VarianceNorm_FAN_IN)
};

class ParamSpec {  // caffe.proto:283-305 (only the fields deploy nets carry)
  PB_STR(name, "")
  PB_OPT(float, lr_mult, 1.f)
  PB_OPT(float, decay_mult, 1.f)
};

enum ConvolutionParameter_Engine {
  ConvolutionParameter_Engine_DEFAULT = 0,
  ConvolutionParameter_Engine_CAFFE = 1,
  ConvolutionParameter_Engine_CUDNN = 2
};
class ConvolutionParameter {  // caffe.proto:567-618
 public:
  typedef ConvolutionParameter_Engine Engine;
  PB_OPT(uint32_t, num_output, 0)
  PB_OPT(bool, bias_term, true)
  PB_REP(uint32_t, pad)
  PB_REP(uint32_t, kernel_size)
  PB_REP(uint32_t, stride)
  PB_REP(uint32_t, dilation)
  PB_OPT(uint32_t, pad_h, 0)
  PB_OPT(uint32_t, pad_w, 0)
  PB_OPT(uint32_t, kernel_h, 0)
  PB_OPT(uint32_t, kernel_w, 0)
  PB_OPT(uint32_t, stride_h, 0)
  PB_OPT(uint32_t, stride_w, 0)
  PB_OPT(uint32_t, group, 1)
  PB_MSG(FillerParameter, weight_filler)
  PB_MSG(FillerParameter, bias_filler)
  PB_OPT(ConvolutionParameter_Engine, engine, ConvolutionParameter_Engine_DEFAULT)
  PB_OPT(int32_t, axis, 1)
  PB_OPT(bool, force_nd_im2col, false)
};

enum PoolingParameter_PoolMethod {
  PoolingParameter_PoolMethod_MAX = 0,
  PoolingParameter_PoolMethod_AVE = 1,
  PoolingParameter_PoolMethod_STOCHASTIC = 2
};
enum PoolingParameter_Engine {
  PoolingParameter_Engine_DEFAULT = 0,
  PoolingParameter_Engine_CAFFE = 1,
  PoolingParameter_Engine_CUDNN = 2
};
class PoolingParameter {  // caffe.proto:893-920
 public:
  typedef PoolingParameter_PoolMethod PoolMethod;
  static const PoolMethod MAX = PoolingParameter_PoolMethod_MAX;
  static const PoolMethod AVE = PoolingParameter_PoolMethod_AVE;
  static const PoolMethod STOCHASTIC = PoolingParameter_PoolMethod_STOCHASTIC;
  PB_OPT(PoolingParameter_PoolMethod, pool, PoolingParameter_PoolMethod_MAX)
  PB_OPT(uint32_t, pad, 0)
  PB_OPT(uint32_t, pad_h, 0)
  PB_OPT(uint32_t, pad_w, 0)
  PB_OPT(uint32_t, kernel_size, 0)
  PB_OPT(uint32_t, kernel_h, 0)
  PB_OPT(uint32_t, kernel_w, 0)
  PB_OPT(uint32_t, stride, 1)
  PB_OPT(uint32_t, stride_h, 0)
  PB_OPT(uint32_t, stride_w, 0)
  PB_OPT(PoolingParameter_Engine, engine, PoolingParameter_Engine_DEFAULT)
  PB_OPT(bool, global_pooling, false)
};

class InnerProductParameter {  // caffe.proto:817-832
  PB_OPT(uint32_t, num_output, 0)
  PB_OPT(bool, bias_term, true)
  PB_MSG(FillerParameter, weight_filler)
  PB_MSG(FillerParameter, bias_filler)
  PB_OPT(int32_t, axis, 1)
  PB_OPT(bool, transpose, false)
};

class InputParameter {  // caffe.proto:834-840
  PB_REP_MSG(BlobShape, shape)
};

class DropoutParameter {  // caffe.proto:672-674
  PB_OPT(float, dropout_ratio, 0.5f)
};

class ConcatParameter {  // caffe.proto:500-509
  PB_OPT(int32_t, axis, 1)
  PB_OPT(uint32_t, concat_dim, 1)
};

enum ReLUParameter_Engine {
  ReLUParameter_Engine_DEFAULT = 0,
  ReLUParameter_Engine_CAFFE = 1,
  ReLUParameter_Engine_CUDNN = 2
};
class ReLUParameter {  // caffe.proto:992-1005
 public:
  typedef ReLUParameter_Engine Engine;
  PB_OPT(float, negative_slope, 0.f)
  PB_OPT(ReLUParameter_Engine, engine, ReLUParameter_Engine_DEFAULT)
};

class ROIPoolingParameter {  // caffe.proto:1257-1266
  PB_OPT(uint32_t, pooled_h, 0)
  PB_OPT(uint32_t, pooled_w, 0)
  PB_OPT(float, spatial_scale, 1.f)
  PB_OPT(float, pad_ratio, 0.f)
};

class BoxOutputParameter {  // caffe.proto:1315-1329
  PB_OPT(float, fg_thr, 0.f)
  PB_OPT(float, iou_thr, 0.5f)
  PB_STR(nms_type, "IOU")
  PB_REP(uint32_t, field_h)
  PB_REP(uint32_t, field_w)
  PB_REP(uint32_t, downsample_rate)
  PB_OPT(float, field_whr, 2.f)
  PB_OPT(float, field_xyr, 2.f)
  PB_OPT(uint32_t, max_nms_num, 0)
  PB_OPT(uint32_t, max_post_nms_num, 0)
  PB_OPT(float, min_size, 15.f)
};

class BBoxRegParameter {  // caffe.proto:1346-1350
  PB_REP(float, bbox_mean)
  PB_REP(float, bbox_std)
  PB_OPT(bool, cls_aware, true)
};

class DecodeBBoxParameter {  // caffe.proto:1353-1355
  PB_OPT(float, gt_iou_thr, 0.95f)
};

enum SoftmaxParameter_Engine {
  SoftmaxParameter_Engine_DEFAULT = 0,
  SoftmaxParameter_Engine_CAFFE = 1,
  SoftmaxParameter_Engine_CUDNN = 2
};
class SoftmaxParameter {  // caffe.proto:1129-1141
 public:
  typedef SoftmaxParameter_Engine Engine;
  PB_OPT(SoftmaxParameter_Engine, engine, SoftmaxParameter_Engine_DEFAULT)
  PB_OPT(int32_t, axis, 1)
};

enum EltwiseParameter_EltwiseOp {
  EltwiseParameter_EltwiseOp_PROD = 0,
  EltwiseParameter_EltwiseOp_SUM = 1,
  EltwiseParameter_EltwiseOp_MAX = 2
};
class EltwiseParameter {  // caffe.proto:695-707
 public:
  typedef EltwiseParameter_EltwiseOp EltwiseOp;
  PB_OPT(EltwiseParameter_EltwiseOp, operation, EltwiseParameter_EltwiseOp_SUM)
  PB_REP(float, coeff)
  PB_OPT(bool, stable_prod_grad, true)
};

class LayerParameter {  // caffe.proto:310-414 (fields used by the deploy nets' layer types)
 public:
  void Clear() { *this = LayerParameter(); }
  void CopyFrom(const LayerParameter& o) { *this = o; }
  PB_STR(name, "")
  PB_STR(type, "")
  PB_REP_STR(bottom)
  PB_REP_STR(top)
  PB_OPT(Phase, phase, TEST)
  PB_REP(float, loss_weight)
  PB_REP_MSG(ParamSpec, param)
  PB_REP_MSG(BlobProto, blobs)
  PB_REP(bool, propagate_down)
  PB_MSG(ConvolutionParameter, convolution_param)
  PB_MSG(PoolingParameter, pooling_param)
  PB_MSG(InnerProductParameter, inner_product_param)
  PB_MSG(InputParameter, input_param)
  PB_MSG(DropoutParameter, dropout_param)
  PB_MSG(ConcatParameter, concat_param)
  PB_MSG(ReLUParameter, relu_param)
  PB_MSG(ROIPoolingParameter, roi_pooling_param)
  PB_MSG(BoxOutputParameter, box_output_param)
  PB_MSG(BBoxRegParameter, bbox_reg_param)
  PB_MSG(DecodeBBoxParameter, decode_bbox_param)
  PB_MSG(SoftmaxParameter, softmax_param)
  PB_MSG(EltwiseParameter, eltwise_param)
};

class NetState {  // caffe.proto:257-261
  PB_OPT(Phase, phase, TEST)
  PB_OPT(int32_t, level, 0)
  PB_REP_STR(stage)
};

class NetParameter {  // caffe.proto:64-100
 public:
  void Clear() { *this = NetParameter(); }
  void CopyFrom(const NetParameter& o) { *this = o; }
  PB_STR(name, "")
  PB_REP_STR(input)
  PB_REP_MSG(BlobShape, input_shape)
  PB_REP(int32_t, input_dim)
  PB_OPT(bool, force_backward, false)
  PB_MSG(NetState, state)
  PB_OPT(bool, debug_info, false)
  PB_REP_MSG(LayerParameter, layer)
};

}  // namespace caffe
