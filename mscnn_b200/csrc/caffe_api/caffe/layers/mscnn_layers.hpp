// The layer types the MS-CNN and cascade deploy nets instantiate (census: SURVEY.md section 2 and 8(f)), as
// caffe::Layer<Dtype> subclasses whose Forward_gpu calls the C ABI in include/mscnn_b200.h.
// Class names, type() strings, blob layouts and parameter semantics are the reference's
// (/root/reference/include/caffe/layers/*.hpp); the per-layer headers next to this file
// (conv_layer.hpp, ...) only include this one so that existing #include lines keep working.
#pragma once
#include <string>
#include <vector>

#include "caffe/blob.hpp"
#include "caffe/layer.hpp"
#include "caffe/proto/caffe.pb.h"

namespace caffe {

// Fill a parameter blob according to a FillerParameter (include/caffe/filler.hpp).  Supported:
// constant, gaussian, bilinear (the only ones the MS-CNN prototxts name); anything else aborts.
CAFFE_API void FillBlob(const FillerParameter& filler, Blob<float>* blob);

// Device copies of packed weights (planes) owned by Convolution / InnerProduct layers.
struct PackedParam {
  PlaneStore w;          // [Cout_pad][KH][KW][Cin_pad] bf16 hi (+lo)
  float* bias = nullptr; // [Cout_pad]
  int cout_pad = 0, cin_pad = 0;
  unsigned long w_version = ~0ul, b_version = ~0ul;
  bool split = false;
  long key = -1;         // layout key (fc: bottom C*H*W arrangement)
  ~PackedParam();
};

/// InputLayer: /root/reference/src/caffe/layers/input_layer.cpp:8-22
template <typename Dtype>
class CAFFE_API InputLayer : public Layer<Dtype> {
 public:
  explicit InputLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {}
  virtual inline const char* type() const { return "Input"; }
  virtual inline int ExactNumBottomBlobs() const { return 0; }
  virtual inline int MinTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {}
};

template <typename Dtype> class PoolingLayer;

/// ConvolutionLayer: conv_layer.cpp:8-40 / base_conv_layer.cpp:15-254 (stride 1, group 1, dilation 1:
/// every Convolution in the shipped deploy nets).  blobs_[0] = [Cout, Cin, kh, kw], blobs_[1] = [Cout].
template <typename Dtype>
class CAFFE_API ConvolutionLayer : public Layer<Dtype> {
 public:
  explicit ConvolutionLayer(const LayerParameter& param) : Layer<Dtype>(param), fuse_relu_(false) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Convolution"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
  // mscnn_b200 extension: fold the in-place ReLU that follows into the epilogue (set by Net).
  void set_fuse_relu(bool f) { fuse_relu_ = f; }
  bool fuse_relu() const { return fuse_relu_; }
  // mscnn_b200 extension (set by Net): the 2x2/2 MAX PoolingLayer that follows is computed in this
  // layer's epilogue and written to `pool_top`; with keep_full == false nothing else reads this
  // layer's own top, which is then never written (its blob holds no data).
  void set_fused_pool(PoolingLayer<Dtype>* pool, Blob<Dtype>* pool_top, bool keep_full) {
    fused_pool_ = pool; fused_pool_top_ = pool_top; fused_keep_full_ = keep_full;
  }
  int num_output() const { return num_output_; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  PoolingLayer<Dtype>* fused_pool_ = nullptr;
  Blob<Dtype>* fused_pool_top_ = nullptr;
  bool fused_keep_full_ = true;
  int num_output_, channels_, kernel_h_, kernel_w_, pad_h_, pad_w_;
  bool bias_term_, fuse_relu_;
  PackedParam packed_;
  PlaneStore patches_;  // conv1_1: 27-tap patch planes
  float* head_bias_ = nullptr;        // narrow k x k heads: bias applied by the tap gather
  unsigned long head_bias_version_ = ~0ul;
 public:
  virtual ~ConvolutionLayer();
};

/// DeconvolutionLayer: deconv_layer.cpp:8-40, restricted to the depthwise 4/2/1 bilinear-upsampling
/// shape of the "-2x" nets.
template <typename Dtype>
class CAFFE_API DeconvolutionLayer : public Layer<Dtype> {
 public:
  explicit DeconvolutionLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Deconvolution"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int channels_;
};

/// ReLULayer: relu_layer.cpp:9-19 (negative_slope 0), in place.
template <typename Dtype>
class CAFFE_API ReLULayer : public Layer<Dtype> {
 public:
  explicit ReLULayer(const LayerParameter& param) : Layer<Dtype>(param), fused_(false) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    top[0]->ReshapeLike(*bottom[0]);
  }
  virtual inline const char* type() const { return "ReLU"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
  void set_fused(bool f) { fused_ = f; }   // producer already applied it
  bool fused() const { return fused_; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  bool fused_;
};

/// PoolingLayer: pooling_layer.cpp:16-220 (MAX / AVE, pad 0, ceil-mode output size).
template <typename Dtype>
class CAFFE_API PoolingLayer : public Layer<Dtype> {
 public:
  explicit PoolingLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Pooling"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int MinTopBlobs() const { return 1; }
  virtual inline int MaxTopBlobs() const { return 1; }
  int kernel() const { return kernel_; }
  int stride() const { return stride_; }
  int mode() const { return mode_; }
  void mark_done_by_producer() { done_by_producer_ = true; }  // the convolution's epilogue pooled already
  virtual void ResetFusedState() { done_by_producer_ = false; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int kernel_, stride_, mode_, pooled_h_, pooled_w_;
  bool done_by_producer_ = false;
};

/// SplitLayer: split_layer.cpp:9-31 (forward = share data).
template <typename Dtype>
class CAFFE_API SplitLayer : public Layer<Dtype> {
 public:
  explicit SplitLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Split"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int MinTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
};

/// ConcatLayer: concat_layer.cpp:11-74, channel axis.
template <typename Dtype>
class CAFFE_API ConcatLayer : public Layer<Dtype> {
 public:
  explicit ConcatLayer(const LayerParameter& param) : Layer<Dtype>(param), fused_(false) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Concat"; }
  virtual inline int MinBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
  void set_fused(bool f) { fused_ = f; }   // the producers write straight into top[0]
  bool fused() const { return fused_; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  bool fused_;
};

/// InnerProductLayer: inner_product_layer.cpp:10-97 (axis 1, no transpose).
/// blobs_[0] = [N_out, K] with K in the bottom's NCHW flattening, blobs_[1] = [N_out].
template <typename Dtype>
class CAFFE_API InnerProductLayer : public Layer<Dtype> {
 public:
  explicit InnerProductLayer(const LayerParameter& param) : Layer<Dtype>(param), fuse_relu_(false) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "InnerProduct"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
  void set_fuse_relu(bool f) { fuse_relu_ = f; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int M_, K_, N_;
  bool bias_term_, fuse_relu_;
  PackedParam packed_;
};

/// DropoutLayer: dropout_layer.cpp:31-46 -- TEST phase is the identity.
template <typename Dtype>
class CAFFE_API DropoutLayer : public Layer<Dtype> {
 public:
  explicit DropoutLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    if (top[0] != bottom[0]) top[0]->ReshapeLike(*bottom[0]);
  }
  virtual inline const char* type() const { return "Dropout"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
};

/// BoxOutputLayer: box_output_layer.cpp:19-234.  top[0] = [R,5,1,1] ROIs, top[1] = [R,6,1,1] with score.
/// Unlike the reference (CPU only) this runs on the device; one small D2H copy of the row count
/// is needed because the Caffe API exposes R as the blobs' shape.
template <typename Dtype>
class CAFFE_API BoxOutputLayer : public Layer<Dtype> {
 public:
  explicit BoxOutputLayer(const LayerParameter& param)
      : Layer<Dtype>(param), workspace_(nullptr), workspace_bytes_(0), num_out_dev_(nullptr),
        num_out_host_(nullptr), num_out_cap_(0) {}
  virtual ~BoxOutputLayer();
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "BoxOutput"; }
  virtual inline int MinBottomBlobs() const { return 1; }
  virtual inline int MinTopBlobs() const { return 1; }
  virtual inline int MaxTopBlobs() const { return 2; }
  // mscnn_b200 extensions: results of the last Forward
  const int* num_out_device() const { return num_out_dev_; }   // int[2+N], see mscnn_box_output_forward
  int num_proposals() const { return num_out_host_ ? num_out_host_[1] : 0; }
  int image_proposals(int n) const { return num_out_host_[2 + n]; }
  // upper bound of the proposals one image can contribute (box_output_layer.cpp:176-186)
  int max_rows_per_image() const {
    return (cfg_.max_post_nms_num > 0 && cfg_.max_post_nms_num < cfg_.max_nms_num) ? cfg_.max_post_nms_num : cfg_.max_nms_num;
  }
  // Deferred rows (set by Net): Forward leaves the tops at cap rows and does NOT wait for the device; ResolveRows
  // waits for the 12-byte copy of the counts and trims the tops.  Without it (a layer driven directly) Forward
  // synchronises and trims itself, which is the reference's observable behaviour (box_output_layer.cpp:201).
  void set_defer_rows(bool on) { defer_rows_ = on; }
  const DynRows* dyn_rows() const { return &dyn_; }
  bool rows_pending() const { return dyn_.pending; }
  void ResolveRows(const vector<Blob<Dtype>*>& top);
  // after a CUDA-graph replay of this layer's launches (Net::GraphForward): the counts are pending again
  void RearmRows();
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  mscnn_box_output_cfg cfg_;
  bool output_proposal_with_score_;
  void* workspace_;
  size_t workspace_bytes_;
  int* num_out_dev_;
  int* num_out_host_;
  int num_out_cap_;
  Blob<Dtype> scratch_score_;
  bool defer_rows_ = false;
  DynRows dyn_ = {nullptr, false};
  cudaEvent_t rows_event_ = nullptr;
};

/// ROIPoolingLayer: roi_pooling_layer.cpp:22-139 with the MS-CNN pad_ratio extension.
template <typename Dtype>
class CAFFE_API ROIPoolingLayer : public Layer<Dtype> {
 public:
  explicit ROIPoolingLayer(const LayerParameter& param)
      : Layer<Dtype>(param), concat_top_(nullptr), concat_offset_(0), concat_channels_(0), rows_source_(nullptr) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "ROIPooling"; }
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int MaxBottomBlobs() const { return 2; }
  virtual inline int MinTopBlobs() const { return 1; }
  virtual inline int MaxTopBlobs() const { return 1; }
  // mscnn_b200 extension (set by Net): the BoxOutput layer this layer's ROI list descends from (per-image row counts
  // for the gather schedule; NULL = unknown origin, plain schedule).
  void set_rows_source(const BoxOutputLayer<Dtype>* box) { rows_source_ = box; }
  // mscnn_b200 extension (set by Net): write into `target` (the Concat top) at a channel offset.
  void set_concat_target(Blob<Dtype>* target, int channel_offset, int total_channels) {
    concat_top_ = target; concat_offset_ = channel_offset; concat_channels_ = total_channels;
  }
  // mscnn_b200 extension (set by Net): the sibling ROIPooling layers that feed the same Concat.  The
  // first one pools all siblings that read the same feature map / ROIs in ONE launch; the others
  // then find their work done (done_by_leader_) and return.
  struct Sibling {
    ROIPoolingLayer<Dtype>* layer;
    Blob<Dtype>* feature;
    Blob<Dtype>* rois;
  };
  void set_siblings(const vector<Sibling>& s) { siblings_ = s; }
  float pad_ratio() const { return pad_ratio_; }
  float spatial_scale() const { return spatial_scale_; }
  int pooled_h() const { return pooled_height_; }
  int pooled_w() const { return pooled_width_; }
  int concat_offset() const { return concat_offset_; }
  void mark_done_by_leader() { done_by_leader_ = true; }
  virtual void ResetFusedState() { done_by_leader_ = false; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int channels_, height_, width_, pooled_height_, pooled_width_;
  float spatial_scale_, pad_ratio_;
  Blob<Dtype>* concat_top_;
  int concat_offset_, concat_channels_;
  const BoxOutputLayer<Dtype>* rows_source_;
  vector<Sibling> siblings_;
  bool done_by_leader_ = false;
};

/// ROIAlignLayer: roi_align_layer.cpp:22-139 (cascade WIDER-face net).  top = [R, C, pooled_h+1, pooled_w+1].
template <typename Dtype>
class CAFFE_API ROIAlignLayer : public Layer<Dtype> {
 public:
  explicit ROIAlignLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "ROIAlign"; }
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int MaxBottomBlobs() const { return 2; }
  virtual inline int MinTopBlobs() const { return 1; }
  virtual inline int MaxTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int channels_, height_, width_, pooled_height_, pooled_width_, grid_height_, grid_width_;
  float spatial_scale_, pad_ratio_;
};

/// DecodeBBoxLayer: decode_bbox_layer.cpp:17-124, TEST phase (2 bottoms: bbox_pred [R,8], prior ROIs [R,5]).
/// CPU-only in the reference; here it stays on the device, so the cascade stages do not sync with the host.
template <typename Dtype>
class CAFFE_API DecodeBBoxLayer : public Layer<Dtype> {
 public:
  explicit DecodeBBoxLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "DecodeBBox"; }
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int MaxBottomBlobs() const { return 3; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  float bbox_mean_[4], bbox_std_[4];
};

/// SoftmaxLayer: softmax_layer.cpp:10-62.
template <typename Dtype>
class CAFFE_API SoftmaxLayer : public Layer<Dtype> {
 public:
  explicit SoftmaxLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Softmax"; }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int outer_num_, inner_num_, softmax_axis_;
};

/// EltwiseLayer: eltwise_layer.cpp:10-96 (PROD / SUM with coefficients / MAX).
template <typename Dtype>
class CAFFE_API EltwiseLayer : public Layer<Dtype> {
 public:
  explicit EltwiseLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  virtual inline const char* type() const { return "Eltwise"; }
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top);
  int op_;
  vector<float> coeffs_;
};

}  // namespace caffe
