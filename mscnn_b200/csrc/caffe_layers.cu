// Layer implementations of the Caffe-API mirror: parameter parsing and shapes follow the
// reference layer sources (cited per method); the arithmetic is delegated to the C ABI.
#include <cstring>
#include <random>
#include <string>

#include "caffe/layers/mscnn_layers.hpp"
#include "config.h"

namespace caffe {

static inline int pad64(int c) { return (c + 63) / 64 * 64; }
static inline int out_pad(int c) { return c > 32 ? pad64(c) : 32; }

PackedParam::~PackedParam() {
  if (bias) cudaFree(bias);
}

void FillBlob(const FillerParameter& filler, Blob<float>* blob) {
  const std::string& type = filler.type();
  float* data = blob->mutable_cpu_data();
  const int n = blob->count();
  if (type == "constant") {  // filler.hpp:36-51
    for (int i = 0; i < n; ++i) data[i] = filler.value();
  } else if (type == "gaussian") {  // filler.hpp:70-107 (without sparsity)
    CHECK_EQ(filler.sparse(), -1) << "sparse gaussian filler is not supported";
    std::mt19937 gen(1706);
    std::normal_distribution<float> dist(filler.mean(), filler.std());
    for (int i = 0; i < n; ++i) data[i] = dist(gen);
  } else if (type == "bilinear") {  // filler.hpp:244-262
    CHECK_EQ(blob->num_axes(), 4) << "Blob must be 4 dim.";
    CHECK_EQ(blob->width(), blob->height()) << "Filter must be square";
    const int f = (int)std::ceil(blob->width() / 2.);
    const float c = (2 * f - 1 - f % 2) / (2. * f);
    for (int i = 0; i < n; ++i) {
      const float x = i % blob->width();
      const float y = (i / blob->width()) % blob->height();
      data[i] = (1 - std::fabs(x / f - c)) * (1 - std::fabs(y / f - c));
    }
  } else {
    LOG(FATAL) << "Unknown or unsupported filler name: " << type;
  }
}

// ---------------------------------------------------------------------------------- Input
template <typename Dtype>
void InputLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const int num_top = (int)top.size();
  const InputParameter& param = this->layer_param_.input_param();
  const int num_shape = param.shape_size();
  CHECK(num_shape == 0 || num_shape == 1 || num_shape == num_top)
      << "Must specify 'shape' once, once per top blob, or not at all: " << num_top << " tops vs. "
      << num_shape << " shapes.";
  if (num_shape > 0) {
    for (int i = 0; i < num_top; ++i) top[i]->Reshape(param.shape(num_shape == 1 ? 0 : i));
  }
}

// ---------------------------------------------------------------------------- Convolution
template <typename Dtype>
void ConvolutionLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  // base_conv_layer.cpp:15-180, 2-D case
  const ConvolutionParameter& p = this->layer_param_.convolution_param();
  CHECK_EQ(bottom[0]->num_axes(), 4) << "mscnn_b200 Convolution takes 4-D bottoms";
  if (p.has_kernel_h() || p.has_kernel_w()) {
    CHECK_EQ(0, p.kernel_size_size()) << "Either kernel_size or kernel_h/w should be specified; not both.";
    kernel_h_ = p.kernel_h();
    kernel_w_ = p.kernel_w();
  } else {
    CHECK(p.kernel_size_size() == 1 || p.kernel_size_size() == 2) << "kernel_size must be specified once or twice";
    kernel_h_ = p.kernel_size(0);
    kernel_w_ = p.kernel_size(p.kernel_size_size() == 1 ? 0 : 1);
  }
  CHECK_GT(kernel_h_, 0) << "Filter dimensions must be nonzero.";
  CHECK_GT(kernel_w_, 0) << "Filter dimensions must be nonzero.";
  if (p.has_pad_h() || p.has_pad_w()) {
    pad_h_ = p.pad_h();
    pad_w_ = p.pad_w();
  } else {
    pad_h_ = p.pad_size() > 0 ? p.pad(0) : 0;
    pad_w_ = p.pad_size() > 1 ? p.pad(1) : pad_h_;
  }
  int stride_h = 1, stride_w = 1;
  if (p.has_stride_h() || p.has_stride_w()) {
    stride_h = p.stride_h();
    stride_w = p.stride_w();
  } else if (p.stride_size() > 0) {
    stride_h = p.stride(0);
    stride_w = p.stride_size() > 1 ? p.stride(1) : stride_h;
  }
  CHECK(stride_h == 1 && stride_w == 1) << "mscnn_b200 Convolution supports stride 1 only (layer "
                                        << this->layer_param_.name() << ")";
  for (int i = 0; i < p.dilation_size(); ++i) CHECK_EQ(p.dilation(i), 1u) << "dilation is not supported";
  CHECK_EQ(p.group(), 1u) << "mscnn_b200 Convolution supports group 1 only";
  channels_ = bottom[0]->channels();
  num_output_ = p.num_output();
  CHECK_GT(num_output_, 0);
  bias_term_ = p.bias_term();
  if (this->blobs_.size() > 0) {
    CHECK_EQ((int)this->blobs_.size(), 1 + (bias_term_ ? 1 : 0)) << "Incorrect number of weight blobs.";
  } else {
    this->blobs_.resize(bias_term_ ? 2 : 1);
    this->blobs_[0].reset(new Blob<Dtype>(num_output_, channels_, kernel_h_, kernel_w_));
    FillBlob(p.weight_filler(), this->blobs_[0].get());
    if (bias_term_) {
      this->blobs_[1].reset(new Blob<Dtype>(vector<int>(1, num_output_)));
      FillBlob(p.bias_filler(), this->blobs_[1].get());
    }
  }
}

template <typename Dtype>
void ConvolutionLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom[0]->channels(), channels_) << "Input size incompatible with convolution kernel.";
  // conv_layer.cpp:8-22 with stride 1, dilation 1
  const int ho = bottom[0]->height() + 2 * pad_h_ - kernel_h_ + 1;
  const int wo = bottom[0]->width() + 2 * pad_w_ - kernel_w_ + 1;
  top[0]->Reshape(bottom[0]->num(), num_output_, ho, wo);
}

// (re)pack weights when the fp32 blobs or the precision changed
static void ensure_packed_conv(PackedParam* pk, Blob<float>* w, Blob<float>* b, bool split, int cout, int cin,
                               int kh, int kw, int cin_pad_override) {
  const int cout_pad = out_pad(cout);
  const int cin_pad = cin_pad_override > 0 ? cin_pad_override : pad64(cin);
  const bool stale_w = pk->w_version != w->version() || pk->split != split || pk->cout_pad != cout_pad ||
                       pk->cin_pad != cin_pad;
  if (stale_w) {
    const size_t bytes = (size_t)cout_pad * kh * kw * cin_pad * 2;
    pk->w.reserve(bytes, split);
    MSCNN_CHECK(mscnn_pack_conv_weights(w->gpu_data(), pk->w.hi, split ? pk->w.lo : nullptr, cout, cin, kh, kw,
                                        cout_pad, cin_pad, Caffe::stream()));
    pk->w_version = w->version();
    pk->split = split;
  }
  const unsigned long bver = b ? b->version() : 0;
  if (!pk->bias || pk->cout_pad != cout_pad || pk->b_version != bver) {
    if (pk->bias && pk->cout_pad != cout_pad) { CUDA_CHECK(cudaFree(pk->bias)); pk->bias = nullptr; }
    if (!pk->bias) CUDA_CHECK(cudaMalloc(&pk->bias, sizeof(float) * cout_pad));
    CUDA_CHECK(cudaMemsetAsync(pk->bias, 0, sizeof(float) * cout_pad, Caffe::stream()));
    if (b)
      CUDA_CHECK(cudaMemcpyAsync(pk->bias, b->gpu_data(), sizeof(float) * cout, cudaMemcpyDeviceToDevice,
                                 Caffe::stream()));
    pk->b_version = bver;
  }
  pk->cout_pad = cout_pad;
  pk->cin_pad = cin_pad;
}

template <typename Dtype>
ConvolutionLayer<Dtype>::~ConvolutionLayer() {
  if (head_bias_) cudaFree(head_bias_);
}

template <typename Dtype>
void ConvolutionLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const bool split = Caffe::split();
  const int N = bottom[0]->num(), H = bottom[0]->height(), W = bottom[0]->width();
  Blob<Dtype>* bias = bias_term_ ? this->blobs_[1].get() : nullptr;
  mscnn_conv_desc d;
  memset(&d, 0, sizeof(d));
  // Narrow k x k proposal heads (LFCN_*: 9 or 6 outputs, 5x5 / 7x7, "same" padding): the horizontal
  // taps move into the GEMM's N dimension and a 1-D gather sums them (mscnn_head_gather in the C ABI).
  const bool head_path = (num_output_ == 9 || num_output_ == 6) && kernel_h_ == kernel_w_ && kernel_h_ > 1 &&
                         pad_h_ == pad_w_ && 2 * pad_h_ + 1 == kernel_h_ && !fuse_relu_ && channels_ % 64 == 0 &&
                         !mscnn::config().no_head_taps;
  if (head_path) {
    const int k = kernel_h_;
    const int n_pad = (k * num_output_ + 63) / 64 * 64;
    Blob<Dtype>* wb = this->blobs_[0].get();
    if (packed_.w_version != wb->version() || packed_.split != split || packed_.cout_pad != n_pad) {
      packed_.w.reserve((size_t)n_pad * k * channels_ * 2, split);
      MSCNN_CHECK(mscnn_pack_head_weights(wb->gpu_data(), packed_.w.hi, split ? packed_.w.lo : nullptr, num_output_,
                                          channels_, k, n_pad, channels_, Caffe::stream()));
      if (packed_.bias) CUDA_CHECK(cudaFree(packed_.bias));
      CUDA_CHECK(cudaMalloc(&packed_.bias, sizeof(float) * n_pad));
      CUDA_CHECK(cudaMemsetAsync(packed_.bias, 0, sizeof(float) * n_pad, Caffe::stream()));  // bias is added later
      packed_.w_version = wb->version();
      packed_.split = split;
      packed_.cout_pad = n_pad;
    }
    const unsigned long bver = bias ? bias->version() : 0;
    if (!head_bias_ || head_bias_version_ != bver) {
      if (!head_bias_) CUDA_CHECK(cudaMalloc(&head_bias_, sizeof(float) * num_output_));
      CUDA_CHECK(cudaMemsetAsync(head_bias_, 0, sizeof(float) * num_output_, Caffe::stream()));
      if (bias)
        CUDA_CHECK(cudaMemcpyAsync(head_bias_, bias->gpu_data(), sizeof(float) * num_output_,
                                   cudaMemcpyDeviceToDevice, Caffe::stream()));
      head_bias_version_ = bver;
    }
    typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
    float* P = static_cast<float*>(Caffe::scratch((size_t)N * H * W * n_pad * sizeof(float)));
    d.x_hi = x.hi; d.x_lo = x.lo;
    d.N = N; d.H = H; d.W = W; d.C = x.cpad;
    d.w_hi = packed_.w.hi; d.w_lo = split ? packed_.w.lo : nullptr;
    d.bias = packed_.bias;
    d.Cout = n_pad; d.Cout_pad = n_pad;
    d.KH = k; d.KW = 1;        // vertical taps stay in K ...
    d.pad_h = pad_h_; d.pad_w = 0;
    d.out_mode = MSCNN_OUT_NHWC_F32;  // ... horizontal taps are columns of P
    d.y_f32 = P;
    d.dyn_n = this->dyn_rows_device();
  MSCNN_CHECK(mscnn_conv_forward(&d, Caffe::stream()));
    MSCNN_CHECK(mscnn_head_gather(P, n_pad, head_bias_, top[0]->mutable_gpu_data(), N, H, W, num_output_, k, pad_h_,
                                  Caffe::stream()));
    return;
  }
  // conv1_1-style layer (3 input channels, 3x3, pad 1): K = 27 is far below one 64-channel
  // tensor-core k-block; it runs as a direct exact-fp32 kernel straight from the NCHW input blob.
  const bool first_path = (channels_ == 3 && kernel_h_ == 3 && kernel_w_ == 3 && pad_h_ == 1 && pad_w_ == 1);
  const int conv1_mode = mscnn::config().conv1_mode;  // MSCNN_CONV1: 1 "pair" | 2 "direct" | 3 "patch" | 0 default = single tensor-core kernel
  if (first_path && num_output_ == 64 && !conv1_mode) {
    // One kernel from the fp32 NCHW blob to planes: pixel rows staged once, the horizontal taps are
    // descriptor-shifted views of them (mscnn_conv1_tc_forward, conv_c3_tc.cu).
    Blob<Dtype>* wb = this->blobs_[0].get();
    constexpr long kTcKey = -64;
    if (packed_.w_version != wb->version() || packed_.key != kTcKey) {
      packed_.w.reserve((size_t)mscnn_conv1_tc_packed_bytes(), false);
      MSCNN_CHECK(mscnn_pack_conv1_tc_weights(wb->gpu_data(), packed_.w.hi, 1, Caffe::stream()));
      packed_.w_version = wb->version();
      packed_.key = kTcKey;
      packed_.b_version = ~0ul;
      if (packed_.bias) { CUDA_CHECK(cudaFree(packed_.bias)); packed_.bias = nullptr; }
    }
    const unsigned long bver = bias ? bias->version() : 0;
    if (!packed_.bias || packed_.b_version != bver) {
      if (!packed_.bias) CUDA_CHECK(cudaMalloc(&packed_.bias, sizeof(float) * 64));
      CUDA_CHECK(cudaMemsetAsync(packed_.bias, 0, sizeof(float) * 64, Caffe::stream()));
      if (bias)
        CUDA_CHECK(cudaMemcpyAsync(packed_.bias, bias->gpu_data(), sizeof(float) * 64, cudaMemcpyDeviceToDevice,
                                   Caffe::stream()));
      packed_.b_version = bver;
    }
    packed_.cout_pad = 64;
    typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
    MSCNN_CHECK(mscnn_conv1_tc_forward(bottom[0]->gpu_data(), packed_.w.hi, packed_.bias, y.hi, split ? y.lo : nullptr,
                                       N, H, W, fuse_relu_ ? 1 : 0, Caffe::stream()));
    return;
  }
  if (first_path && (W % 2 == 0) && num_output_ % 64 == 0 && conv1_mode == 1) {
    // Pixel-pair GEMM: rows = two adjacent pixels (K = 54 of 64), block-diagonal weights, the output
    // [N][H][W/2][2*Cout] is the NHWC tensor [N][H][W][Cout] (see mscnn_im2col3x3_c3_pair_to_planes).
    const int cp = num_output_;  // multiple of 64
    Blob<Dtype>* wb = this->blobs_[0].get();
    if (packed_.w_version != wb->version() || packed_.split != split || packed_.cout_pad != 2 * cp || packed_.key != -2) {
      packed_.key = -2;
      packed_.w.reserve((size_t)2 * cp * 64 * 2, split);
      MSCNN_CHECK(mscnn_pack_conv1_pair_weights(wb->gpu_data(), packed_.w.hi, split ? packed_.w.lo : nullptr,
                                                num_output_, cp, Caffe::stream()));
      packed_.w_version = wb->version();
      packed_.split = split;
      packed_.b_version = ~0ul;
      if (packed_.bias) { CUDA_CHECK(cudaFree(packed_.bias)); packed_.bias = nullptr; }
    }
    const unsigned long bver = bias ? bias->version() : 0;
    if (!packed_.bias || packed_.b_version != bver) {
      if (!packed_.bias) CUDA_CHECK(cudaMalloc(&packed_.bias, sizeof(float) * 2 * cp));
      CUDA_CHECK(cudaMemsetAsync(packed_.bias, 0, sizeof(float) * 2 * cp, Caffe::stream()));
      if (bias) {
        CUDA_CHECK(cudaMemcpyAsync(packed_.bias, bias->gpu_data(), sizeof(float) * num_output_,
                                   cudaMemcpyDeviceToDevice, Caffe::stream()));
        CUDA_CHECK(cudaMemcpyAsync(packed_.bias + cp, bias->gpu_data(), sizeof(float) * num_output_,
                                   cudaMemcpyDeviceToDevice, Caffe::stream()));
      }
      packed_.b_version = bver;
    }
    packed_.cout_pad = 2 * cp;
    const size_t bytes = (size_t)N * H * (W / 2) * 64 * 2;
    patches_.reserve(bytes, split);
    MSCNN_CHECK(mscnn_im2col3x3_c3_pair_to_planes(bottom[0]->gpu_data(), patches_.hi,
                                                  split ? patches_.lo : nullptr, N, H, W, Caffe::stream()));
    typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
    d.x_hi = patches_.hi;
    d.x_lo = split ? patches_.lo : nullptr;
    d.N = N; d.H = H; d.W = W / 2; d.C = 64;
    d.w_hi = packed_.w.hi;
    d.w_lo = split ? packed_.w.lo : nullptr;
    d.bias = packed_.bias;
    d.Cout = 2 * cp; d.Cout_pad = 2 * cp;
    d.KH = d.KW = 1;
    d.relu = fuse_relu_ ? 1 : 0;
    d.out_mode = MSCNN_OUT_NHWC_BF16;
    d.y_hi = y.hi;
    d.y_lo = y.lo;
    d.dyn_n = this->dyn_rows_device();
  MSCNN_CHECK(mscnn_conv_forward(&d, Caffe::stream()));
    return;
  }
  if (first_path && conv1_mode != 3) {
    typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
    MSCNN_CHECK(mscnn_conv3x3_c3_forward(bottom[0]->gpu_data(), this->blobs_[0]->gpu_data(),
                                         bias ? bias->gpu_data() : nullptr, y.hi, y.lo, N, H, W, num_output_,
                                         y.cpad, fuse_relu_ ? 1 : 0, Caffe::stream()));
    return;
  }
  // (MSCNN_CONV1=patch: the same layer as a 1x1 GEMM over single-pixel 27-tap patch planes, kept for comparison)
  const bool patch_path = first_path;
  if (patch_path) {
    ensure_packed_conv(&packed_, this->blobs_[0].get(), bias, split, num_output_, 27, 1, 1, 64);
    const size_t bytes = (size_t)N * H * W * 64 * 2;
    patches_.reserve(bytes, split);
    MSCNN_CHECK(mscnn_im2col3x3_c3_to_planes(bottom[0]->gpu_data(), patches_.hi, split ? patches_.lo : nullptr, N,
                                             H, W, Caffe::stream()));
    d.x_hi = patches_.hi;
    d.x_lo = split ? patches_.lo : nullptr;
    d.C = 64;
    d.KH = d.KW = 1;
    d.pad_h = d.pad_w = 0;
  } else {
    ensure_packed_conv(&packed_, this->blobs_[0].get(), bias, split, num_output_, channels_, kernel_h_, kernel_w_,
                       0);
    typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
    d.x_hi = x.hi;
    d.x_lo = x.lo;
    d.C = x.cpad;
    d.KH = kernel_h_;
    d.KW = kernel_w_;
    d.pad_h = pad_h_;
    d.pad_w = pad_w_;
  }
  d.N = N; d.H = H; d.W = W;
  d.w_hi = packed_.w.hi;
  d.w_lo = split ? packed_.w.lo : nullptr;
  d.bias = packed_.bias;
  d.Cout = num_output_;
  d.Cout_pad = packed_.cout_pad;
  d.relu = fuse_relu_ ? 1 : 0;
  if (packed_.cout_pad % 64 == 0) {
    d.out_mode = MSCNN_OUT_NHWC_BF16;
    const int ho = top[0]->height(), wo = top[0]->width();
    const bool pool_here = fused_pool_ && !patch_path && (ho % 2 == 0) && (wo % 2 == 0);
    if (!pool_here || fused_keep_full_) {
      typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
      d.y_hi = y.hi;
      d.y_lo = y.lo;
    }
    if (pool_here) {
      fused_pool_top_->Reshape(top[0]->num(), num_output_, ho / 2, wo / 2);
      typename Blob<Dtype>::Planes q = fused_pool_top_->mutable_planes(split);
      d.pool_hi = q.hi;
      d.pool_lo = q.lo;
      fused_pool_->mark_done_by_producer();
    }
  } else {
    d.out_mode = MSCNN_OUT_NCHW_F32;  // narrow heads (LFCN_*): straight into the Caffe layout
    d.y_f32 = top[0]->mutable_gpu_data();
  }
  d.dyn_n = this->dyn_rows_device();
  MSCNN_CHECK(mscnn_conv_forward(&d, Caffe::stream()));
}

// -------------------------------------------------------------------------- Deconvolution
template <typename Dtype>
void DeconvolutionLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const ConvolutionParameter& p = this->layer_param_.convolution_param();
  channels_ = bottom[0]->channels();
  const int k = p.kernel_size_size() ? (int)p.kernel_size(0) : (int)p.kernel_h();
  const int s = p.stride_size() ? (int)p.stride(0) : (p.has_stride_h() ? (int)p.stride_h() : 1);
  const int pd = p.pad_size() ? (int)p.pad(0) : (int)p.pad_h();
  CHECK(k == 4 && s == 2 && pd == 1 && (int)p.group() == channels_ && (int)p.num_output() == channels_ &&
        !p.bias_term())
      << "mscnn_b200 Deconvolution supports the depthwise kernel 4 / stride 2 / pad 1 / no-bias shape of the "
         "MS-CNN -2x nets only (layer " << this->layer_param_.name() << ")";
  if (this->blobs_.empty()) {
    this->blobs_.resize(1);
    this->blobs_[0].reset(new Blob<Dtype>(channels_, 1, 4, 4));  // deconv: [Cin, Cout/group, kh, kw]
    FillBlob(p.weight_filler(), this->blobs_[0].get());
  }
}
template <typename Dtype>
void DeconvolutionLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  // deconv_layer.cpp:8-22: stride*(in-1) + kernel - 2*pad = 2*in
  top[0]->Reshape(bottom[0]->num(), channels_, 2 * bottom[0]->height(), 2 * bottom[0]->width());
}
template <typename Dtype>
void DeconvolutionLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const bool split = Caffe::split();
  typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
  typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
  MSCNN_CHECK(mscnn_deconv2x_forward(x.hi, x.lo, this->blobs_[0]->gpu_data(), y.hi, y.lo, x.n, x.h, x.w, x.cpad,
                                     channels_, Caffe::stream()));
}

// ----------------------------------------------------------------------------------- ReLU
template <typename Dtype>
void ReLULayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(this->layer_param_.relu_param().negative_slope(), 0.f) << "leaky ReLU is not supported";
}
template <typename Dtype>
void ReLULayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (fused_) return;  // the producing Convolution / InnerProduct already clamped at zero
  if (top[0] != bottom[0]) top[0]->CopyFrom(*bottom[0], false, true);
  if (top[0]->head_is_planes()) {
    typename Blob<Dtype>::Planes p = top[0]->planes(Caffe::split());
    MSCNN_CHECK(mscnn_relu_planes(p.hi, p.lo, (size_t)p.n * p.h * p.w * p.cpad, Caffe::stream()));
    top[0]->mutable_planes(p.lo != nullptr);  // mark planes as the (modified) head
  } else {
    MSCNN_CHECK(mscnn_relu_f32(top[0]->mutable_gpu_data(), top[0]->count(), Caffe::stream()));
  }
}

// -------------------------------------------------------------------------------- Pooling
template <typename Dtype>
void PoolingLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const PoolingParameter& p = this->layer_param_.pooling_param();  // pooling_layer.cpp:16-77
  CHECK(!p.global_pooling()) << "global pooling is not supported";
  CHECK(p.has_kernel_size() || (p.has_kernel_h() && p.has_kernel_w())) << "kernel size is required";
  const int kh = p.has_kernel_size() ? p.kernel_size() : p.kernel_h();
  const int kw = p.has_kernel_size() ? p.kernel_size() : p.kernel_w();
  const int sh = p.has_stride_h() ? p.stride_h() : p.stride();
  const int sw = p.has_stride_w() ? p.stride_w() : p.stride();
  CHECK(kh == kw && sh == sw) << "square kernel / stride only";
  CHECK(p.pad() == 0 && p.pad_h() == 0 && p.pad_w() == 0) << "mscnn_b200 Pooling supports pad 0 only";
  kernel_ = kh;
  stride_ = sh;
  CHECK(p.pool() == PoolingParameter_PoolMethod_MAX || p.pool() == PoolingParameter_PoolMethod_AVE)
      << "MAX and AVE pooling only";
  mode_ = p.pool() == PoolingParameter_PoolMethod_MAX ? MSCNN_POOL_MAX : MSCNN_POOL_AVE;
}
template <typename Dtype>
void PoolingLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(4, bottom[0]->num_axes()) << "Input must have 4 axes";
  // pooling_layer.cpp:90-93 (pad 0)
  pooled_h_ = (int)std::ceil((float)(bottom[0]->height() - kernel_) / stride_) + 1;
  pooled_w_ = (int)std::ceil((float)(bottom[0]->width() - kernel_) / stride_) + 1;
  top[0]->Reshape(bottom[0]->num(), bottom[0]->channels(), pooled_h_, pooled_w_);
}
template <typename Dtype>
void PoolingLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (done_by_producer_) {  // computed in the producing convolution's epilogue
    done_by_producer_ = false;
    return;
  }
  const bool split = Caffe::split();
  typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
  typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
  MSCNN_CHECK(mscnn_pool_forward_dyn(x.hi, x.lo, y.hi, y.lo, x.n, x.h, x.w, x.cpad, kernel_, stride_, mode_,
                                     this->dyn_rows_device(), Caffe::stream()));
}

// ---------------------------------------------------------------------------------- Split
template <typename Dtype>
void SplitLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  for (size_t i = 0; i < top.size(); ++i) {
    CHECK_NE(top[i], bottom[0]) << this->type() << " Layer does not allow in-place computation.";
    top[i]->ReshapeLike(*bottom[0]);
  }
}
template <typename Dtype>
void SplitLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  for (size_t i = 0; i < top.size(); ++i) top[i]->ShareData(*bottom[0]);
}

// --------------------------------------------------------------------------------- Concat
template <typename Dtype>
void ConcatLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const ConcatParameter& p = this->layer_param_.concat_param();
  CHECK(!(p.has_axis() && p.has_concat_dim())) << "Either axis or concat_dim should be specified; not both.";
  const int axis = p.has_concat_dim() ? (int)p.concat_dim() : bottom[0]->CanonicalAxisIndex(p.axis());
  CHECK_EQ(axis, 1) << "mscnn_b200 Concat supports the channel axis only";
}
template <typename Dtype>
void ConcatLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  vector<int> shape = bottom[0]->shape();  // concat_layer.cpp:28-55
  for (size_t i = 1; i < bottom.size(); ++i) {
    CHECK_EQ(bottom[0]->num_axes(), bottom[i]->num_axes()) << "All inputs must have the same #axes.";
    for (int j = 0; j < bottom[0]->num_axes(); ++j)
      if (j != 1) CHECK_EQ(shape[j], bottom[i]->shape(j)) << "All inputs must have the same shape, except at concat_axis.";
    shape[1] += bottom[i]->shape(1);
  }
  top[0]->Reshape(shape);
}
template <typename Dtype>
void ConcatLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (fused_) return;  // ROIPooling producers already wrote their channel ranges of top[0]
  if (bottom.size() == 1) { top[0]->ShareData(*bottom[0]); return; }
  const bool split = Caffe::split();
  bool aligned = true;
  for (size_t i = 0; i < bottom.size(); ++i) aligned = aligned && (bottom[i]->channels() % 64 == 0);
  if (aligned) {
    typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
    int off = 0;
    for (size_t i = 0; i < bottom.size(); ++i) {
      typename Blob<Dtype>::Planes x = bottom[i]->planes(split);
      const size_t pixels = (size_t)x.n * x.h * x.w;
      MSCNN_CHECK(mscnn_concat_planes(x.hi, y.hi, pixels, x.cpad, y.cpad, off, Caffe::stream()));
      if (split) MSCNN_CHECK(mscnn_concat_planes(x.lo, y.lo, pixels, x.cpad, y.cpad, off, Caffe::stream()));
      off += x.cpad;
    }
  } else {
    Dtype* dst = top[0]->mutable_gpu_data();
    const int num = top[0]->num();
    const size_t inner = (size_t)top[0]->count(2);
    const size_t top_row = (size_t)top[0]->channels() * inner;
    size_t off = 0;
    for (size_t i = 0; i < bottom.size(); ++i) {
      const size_t row = (size_t)bottom[i]->channels() * inner;
      CUDA_CHECK(cudaMemcpy2DAsync(dst + off, top_row * sizeof(Dtype), bottom[i]->gpu_data(), row * sizeof(Dtype),
                                   row * sizeof(Dtype), num, cudaMemcpyDeviceToDevice, Caffe::stream()));
      off += row;
    }
  }
}

// --------------------------------------------------------------------------- InnerProduct
template <typename Dtype>
void InnerProductLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const InnerProductParameter& p = this->layer_param_.inner_product_param();  // inner_product_layer.cpp:10-56
  N_ = p.num_output();
  bias_term_ = p.bias_term();
  CHECK(!p.transpose()) << "transpose is not supported";
  const int axis = bottom[0]->CanonicalAxisIndex(p.axis());
  CHECK_EQ(axis, 1) << "mscnn_b200 InnerProduct supports axis 1 only";
  K_ = bottom[0]->count(axis);
  if (this->blobs_.size() > 0) {
    CHECK_EQ((int)this->blobs_.size(), 1 + (bias_term_ ? 1 : 0));
  } else {
    this->blobs_.resize(bias_term_ ? 2 : 1);
    vector<int> ws(2);
    ws[0] = N_;
    ws[1] = K_;
    this->blobs_[0].reset(new Blob<Dtype>(ws));
    FillBlob(p.weight_filler(), this->blobs_[0].get());
    if (bias_term_) {
      this->blobs_[1].reset(new Blob<Dtype>(vector<int>(1, N_)));
      FillBlob(p.bias_filler(), this->blobs_[1].get());
    }
  }
}
template <typename Dtype>
void InnerProductLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const int new_K = bottom[0]->count(1);
  CHECK_EQ(K_, new_K) << "Input size incompatible with inner product parameters.";
  M_ = bottom[0]->count(0, 1);
  vector<int> ts(2);
  ts[0] = M_;
  ts[1] = N_;
  top[0]->Reshape(ts);  // inner_product_layer.cpp:58-74
}
template <typename Dtype>
void InnerProductLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const bool split = Caffe::split();
  const int C = bottom[0]->LegacyShape(1), H = bottom[0]->LegacyShape(2), W = bottom[0]->LegacyShape(3);
  typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
  const int nout_pad = out_pad(N_);
  const long key = ((long)C << 32) | ((long)H << 16) | W;
  Blob<Dtype>* wb = this->blobs_[0].get();
  if (packed_.w_version != wb->version() || packed_.split != split || packed_.key != key ||
      packed_.cout_pad != nout_pad) {
    const size_t bytes = (size_t)nout_pad * H * W * x.cpad * 2;
    packed_.w.reserve(bytes, split);
    MSCNN_CHECK(mscnn_pack_fc_weights(wb->gpu_data(), packed_.w.hi, split ? packed_.w.lo : nullptr, N_, C, H, W,
                                      nout_pad, x.cpad, Caffe::stream()));
    packed_.w_version = wb->version();
    packed_.split = split;
    packed_.key = key;
  }
  Blob<Dtype>* bb = bias_term_ ? this->blobs_[1].get() : nullptr;
  const unsigned long bver = bb ? bb->version() : 0;
  if (!packed_.bias || packed_.cout_pad != nout_pad || packed_.b_version != bver) {
    if (packed_.bias && packed_.cout_pad != nout_pad) { CUDA_CHECK(cudaFree(packed_.bias)); packed_.bias = nullptr; }
    if (!packed_.bias) CUDA_CHECK(cudaMalloc(&packed_.bias, sizeof(float) * nout_pad));
    CUDA_CHECK(cudaMemsetAsync(packed_.bias, 0, sizeof(float) * nout_pad, Caffe::stream()));
    if (bb)
      CUDA_CHECK(cudaMemcpyAsync(packed_.bias, bb->gpu_data(), sizeof(float) * N_, cudaMemcpyDeviceToDevice,
                                 Caffe::stream()));
    packed_.b_version = bver;
  }
  packed_.cout_pad = nout_pad;

  mscnn_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.x_hi = x.hi;
  d.x_lo = x.lo;
  d.N = M_; d.H = 1; d.W = 1;
  d.C = H * W * x.cpad;  // NHWC flattening of the bottom
  d.w_hi = packed_.w.hi;
  d.w_lo = split ? packed_.w.lo : nullptr;
  d.bias = packed_.bias;
  d.Cout = N_;
  d.Cout_pad = nout_pad;
  d.KH = d.KW = 1;
  d.relu = fuse_relu_ ? 1 : 0;
  if (nout_pad % 64 == 0) {
    typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
    d.out_mode = MSCNN_OUT_NHWC_BF16;
    d.y_hi = y.hi;
    d.y_lo = y.lo;
  } else {
    d.out_mode = MSCNN_OUT_NCHW_F32;  // cls_pred / bbox_pred: [R][N_] fp32, the Caffe layout
    d.y_f32 = top[0]->mutable_gpu_data();
  }
  d.dyn_n = this->dyn_rows_device();
  MSCNN_CHECK(mscnn_conv_forward(&d, Caffe::stream()));
}

// -------------------------------------------------------------------------------- Dropout
template <typename Dtype>
void DropoutLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK(this->phase_ == TEST) << "mscnn_b200 is forward-only: Dropout runs in the TEST phase";
  if (top[0] != bottom[0]) top[0]->ShareData(*bottom[0]);  // dropout_layer.cpp:43-45 copies; sharing is equivalent
}

// ------------------------------------------------------------------------------ BoxOutput
template <typename Dtype>
BoxOutputLayer<Dtype>::~BoxOutputLayer() {
  if (workspace_) cudaFree(workspace_);
  if (num_out_dev_) cudaFree(num_out_dev_);
  if (num_out_host_) cudaFreeHost(num_out_host_);
  if (rows_event_) cudaEventDestroy(rows_event_);
}
template <typename Dtype>
void BoxOutputLayer<Dtype>::ResolveRows(const vector<Blob<Dtype>*>& top) {
  if (!dyn_.pending) return;
  CUDA_CHECK(cudaEventSynchronize(rows_event_));
  dyn_.pending = false;
  const int rows = num_out_host_[0];
  top[0]->Reshape(rows, 5, 1, 1);
  if (output_proposal_with_score_) top[1]->Reshape(rows, 6, 1, 1);
}
template <typename Dtype>
void BoxOutputLayer<Dtype>::RearmRows() {
  CUDA_CHECK(cudaEventRecord(rows_event_, Caffe::stream()));
  dyn_.pending = true;
}
template <typename Dtype>
void BoxOutputLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const BoxOutputParameter& p = this->layer_param_.box_output_param();  // box_output_layer.cpp:19-26, 79-104
  const int J = (int)bottom.size();
  CHECK_LE(J, MSCNN_MAX_SCALES);
  CHECK_EQ(J, p.field_h_size());
  CHECK_EQ(J, p.field_w_size());
  CHECK_EQ(J, p.downsample_rate_size());
  memset(&cfg_, 0, sizeof(cfg_));
  cfg_.num_scales = J;
  for (int j = 0; j < J; ++j) {
    cfg_.field_w[j] = (float)p.field_w(j);
    cfg_.field_h[j] = (float)p.field_h(j);
    cfg_.downsample_rate[j] = (float)p.downsample_rate(j);
  }
  cfg_.fg_thr = p.fg_thr();
  cfg_.iou_thr = p.iou_thr();
  cfg_.nms_type = p.nms_type() == "IOMU" ? MSCNN_NMS_IOMU : p.nms_type() == "IOFU" ? MSCNN_NMS_IOFU : MSCNN_NMS_IOU;
  cfg_.field_whr = p.field_whr();
  cfg_.field_xyr = p.field_xyr();
  cfg_.min_size = p.min_size();
  cfg_.max_nms_num = p.max_nms_num();
  cfg_.max_post_nms_num = p.max_post_nms_num();
  CHECK(cfg_.max_nms_num > 0 && cfg_.max_nms_num <= 8192)
      << "mscnn_b200 BoxOutput needs 0 < max_nms_num <= 8192 (the reference's 0 = unbounded is not supported)";
  const BBoxRegParameter& r = this->layer_param_.bbox_reg_param();
  cfg_.do_bbox_norm = (r.bbox_mean_size() > 0 && r.bbox_std_size() > 0) ? 1 : 0;
  if (cfg_.do_bbox_norm) {
    CHECK_EQ(r.bbox_mean_size(), 4);
    CHECK_EQ(r.bbox_std_size(), 4);
    for (int k = 0; k < 4; ++k) { cfg_.bbox_mean[k] = r.bbox_mean(k); cfg_.bbox_std[k] = r.bbox_std(k); }
  }
  output_proposal_with_score_ = (top.size() == 2);
}
template <typename Dtype>
void BoxOutputLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  // "dummy reshape" (box_output_layer.cpp:29-36); the real shape is data dependent.  Rows still pending from the
  // last Forward (Net::ResolveRows re-runs Reshape down the net) keep their cap shape until resolved.
  if (dyn_.pending) return;
  top[0]->Reshape(1, 5, 1, 1);
  if (output_proposal_with_score_) top[1]->Reshape(1, 6, 1, 1);
}
template <typename Dtype>
void BoxOutputLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const int N = bottom[0]->num();
  cfg_.channels = bottom[0]->channels();
  const float* maps[MSCNN_MAX_SCALES];
  for (size_t j = 0; j < bottom.size(); ++j) {
    CHECK_EQ(bottom[j]->channels(), cfg_.channels);
    CHECK_EQ(bottom[j]->num(), N);
    cfg_.height[j] = bottom[j]->height();
    cfg_.width[j] = bottom[j]->width();
    maps[j] = bottom[j]->gpu_data();
  }
  size_t need = 0;
  MSCNN_CHECK(mscnn_box_output_workspace_bytes(&cfg_, N, &need));
  if (need > workspace_bytes_) {
    if (workspace_) CUDA_CHECK(cudaFree(workspace_));
    CUDA_CHECK(cudaMalloc(&workspace_, need));
    workspace_bytes_ = need;
  }
  if (2 + N > num_out_cap_) {
    if (num_out_dev_) CUDA_CHECK(cudaFree(num_out_dev_));
    if (num_out_host_) CUDA_CHECK(cudaFreeHost(num_out_host_));
    CUDA_CHECK(cudaMalloc(&num_out_dev_, sizeof(int) * (2 + N)));
    CUDA_CHECK(cudaMallocHost(&num_out_host_, sizeof(int) * (2 + N)));
    num_out_cap_ = 2 + N;
    dyn_.device_rows = num_out_dev_;  // [0] = rows in the top blobs (>= 1)
  }
  if (!rows_event_) CUDA_CHECK(cudaEventCreateWithFlags(&rows_event_, cudaEventDisableTiming));
  dyn_.pending = false;  // a new Forward supersedes counts nobody asked for
  // Size the tops for the cap, let the kernels write in place, then shrink to the true row
  // count (Blob::Reshape never reallocates when shrinking, blob.cpp:40-44).
  const int cap = N * cfg_.max_nms_num;
  top[0]->Reshape(cap, 5, 1, 1);
  Blob<Dtype>* score_blob = output_proposal_with_score_ ? top[1] : &scratch_score_;
  score_blob->Reshape(cap, 6, 1, 1);
  Dtype* rois = top[0]->mutable_gpu_data();
  Dtype* rois_score = output_proposal_with_score_ ? score_blob->mutable_gpu_data() : nullptr;
  MSCNN_CHECK(mscnn_box_output_forward(&cfg_, N, maps, workspace_, workspace_bytes_, rois, rois_score,
                                       num_out_dev_, Caffe::stream()));
  CUDA_CHECK(cudaMemcpyAsync(num_out_host_, num_out_dev_, sizeof(int) * (2 + N), cudaMemcpyDeviceToHost,
                             Caffe::stream()));
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(Caffe::stream(), &capturing);
  if (capturing == cudaStreamCaptureStatusNone)  // under capture the Net re-arms after the graph launch (RearmRows)
    CUDA_CHECK(cudaEventRecord(rows_event_, Caffe::stream()));
  dyn_.pending = true;  // tops have `cap` rows; the true count is in num_out_dev_[0] (and soon in num_out_host_[0])
  if (!defer_rows_) ResolveRows(top);
}

// ----------------------------------------------------------------------------- ROIPooling
template <typename Dtype>
void ROIPoolingLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const ROIPoolingParameter& p = this->layer_param_.roi_pooling_param();  // roi_pooling_layer.cpp:22-34
  CHECK_GT(p.pooled_h(), 0u) << "pooled_h must be > 0";
  CHECK_GT(p.pooled_w(), 0u) << "pooled_w must be > 0";
  pooled_height_ = p.pooled_h();
  pooled_width_ = p.pooled_w();
  spatial_scale_ = p.spatial_scale();
  pad_ratio_ = p.pad_ratio();
}
template <typename Dtype>
void ROIPoolingLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  channels_ = bottom[0]->channels();
  height_ = bottom[0]->height();
  width_ = bottom[0]->width();
  top[0]->Reshape(bottom[1]->num(), channels_, pooled_height_, pooled_width_);
}
template <typename Dtype>
void ROIPoolingLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  if (done_by_leader_) {  // a sibling already pooled this variant in its launch
    done_by_leader_ = false;
    return;
  }
  const bool split = Caffe::split();
  typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
  const int R = bottom[1]->num();
  const Dtype* rois = bottom[1]->gpu_data();
  if (concat_top_) {
    // fused ConcatLayer: the Net sized concat_top_ as [R, total, ph, pw]
    concat_top_->Reshape(R, concat_channels_, pooled_height_, pooled_width_);
    typename Blob<Dtype>::Planes y = concat_top_->mutable_planes(split);
    float ratios[4];
    int offs[4];
    int nv = 0;
    ratios[nv] = pad_ratio_;
    offs[nv++] = concat_offset_;
    for (size_t i = 0; i < siblings_.size() && nv < 4; ++i) {
      ROIPoolingLayer<Dtype>* sb = siblings_[i].layer;
      if (sb == this) continue;
      // same feature map (shared through Split), same ROIs, same geometry -> same launch
      typename Blob<Dtype>::Planes sx = siblings_[i].feature->planes(split);
      if (sx.hi == x.hi && siblings_[i].rois->gpu_data() == rois && sb->pooled_h() == pooled_height_ &&
          sb->pooled_w() == pooled_width_ && sb->spatial_scale() == spatial_scale_) {
        ratios[nv] = sb->pad_ratio();
        offs[nv++] = sb->concat_offset();
        sb->mark_done_by_leader();
      }
    }
    MSCNN_CHECK(mscnn_roi_pool_multi_forward_dyn(x.hi, x.lo, x.n, x.h, x.w, x.cpad, rois, R, pooled_height_,
                                                 pooled_width_, spatial_scale_, nv, ratios, offs, y.hi, y.lo, y.cpad,
                                                 this->dyn_rows_device(), Caffe::stream()));
  } else {
    typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
    const int off0 = 0;
    MSCNN_CHECK(mscnn_roi_pool_multi_forward_dyn(x.hi, x.lo, x.n, x.h, x.w, x.cpad, rois, R, pooled_height_,
                                                 pooled_width_, spatial_scale_, 1, &pad_ratio_, &off0, y.hi, y.lo,
                                                 y.cpad, this->dyn_rows_device(), Caffe::stream()));
  }
}

// ------------------------------------------------------------------------------- ROIAlign
template <typename Dtype>
void ROIAlignLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const ROIPoolingParameter& p = this->layer_param_.roi_pooling_param();  // roi_align_layer.cpp:22-37
  CHECK_GT(p.pooled_h(), 0u) << "pooled_h must be > 0";
  CHECK_GT(p.pooled_w(), 0u) << "pooled_w must be > 0";
  pooled_height_ = p.pooled_h();
  pooled_width_ = p.pooled_w();
  grid_height_ = pooled_height_ + 1;
  grid_width_ = pooled_width_ + 1;
  spatial_scale_ = p.spatial_scale();
  pad_ratio_ = p.pad_ratio();
}
template <typename Dtype>
void ROIAlignLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  channels_ = bottom[0]->channels();
  height_ = bottom[0]->height();
  width_ = bottom[0]->width();
  top[0]->Reshape(bottom[1]->num(), channels_, grid_height_, grid_width_);
}
template <typename Dtype>
void ROIAlignLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const bool split = Caffe::split();
  typename Blob<Dtype>::Planes x = bottom[0]->planes(split);
  typename Blob<Dtype>::Planes y = top[0]->mutable_planes(split);
  MSCNN_CHECK(mscnn_roi_align_forward_dyn(x.hi, x.lo, x.n, x.h, x.w, x.cpad, bottom[1]->gpu_data(), bottom[1]->num(),
                                          pooled_height_, pooled_width_, spatial_scale_, pad_ratio_, y.hi, y.lo,
                                          y.cpad, 0, this->dyn_rows_device(), Caffe::stream()));
}

// ----------------------------------------------------------------------------- DecodeBBox
template <typename Dtype>
void DecodeBBoxLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const BBoxRegParameter& r = this->layer_param_.bbox_reg_param();  // decode_bbox_layer.cpp:20-36
  if (r.bbox_mean_size() > 0 && r.bbox_std_size() > 0) {
    CHECK_EQ(r.bbox_mean_size(), 4);
    CHECK_EQ(r.bbox_std_size(), 4);
    for (int i = 0; i < 4; ++i) {
      bbox_mean_[i] = r.bbox_mean(i);
      bbox_std_[i] = r.bbox_std(i);
      CHECK_GT(bbox_std_[i], 0);
    }
  } else {
    for (int i = 0; i < 4; ++i) { bbox_mean_[i] = 0.f; bbox_std_[i] = 1.f; }
  }
}
template <typename Dtype>
void DecodeBBoxLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK_EQ(bottom[0]->num(), bottom[1]->num());  // decode_bbox_layer.cpp:42-50
  CHECK(bottom.size() < 3) << "mscnn_b200 is forward-only: DecodeBBox takes gt boxes only in the TRAIN phase";
  CHECK_EQ(bottom[0]->channels(), 8);
  CHECK_EQ(bottom[1]->channels(), 5);
  top[0]->ReshapeLike(*bottom[1]);
}
template <typename Dtype>
void DecodeBBoxLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  CHECK(this->phase_ == TEST) << "mscnn_b200 is forward-only: DecodeBBox runs in the TEST phase";
  const int num = bottom[0]->num();
  // TEST phase keeps every row (decode_bbox_layer.cpp:79-110), so the top has the priors' shape
  top[0]->Reshape(num, bottom[1]->channels(), 1, 1);
  MSCNN_CHECK(mscnn_decode_bbox_forward(bottom[0]->gpu_data(), bottom[1]->gpu_data(), num, bottom[0]->channels(),
                                        bbox_mean_, bbox_std_, top[0]->mutable_gpu_data(), Caffe::stream()));
}

// -------------------------------------------------------------------------------- Softmax
template <typename Dtype>
void SoftmaxLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  softmax_axis_ = bottom[0]->CanonicalAxisIndex(this->layer_param_.softmax_param().axis());  // softmax_layer.cpp:10-25
  top[0]->ReshapeLike(*bottom[0]);
  outer_num_ = bottom[0]->count(0, softmax_axis_);
  inner_num_ = bottom[0]->count(softmax_axis_ + 1);
}
template <typename Dtype>
void SoftmaxLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  MSCNN_CHECK(mscnn_softmax_forward(bottom[0]->gpu_data(), outer_num_, bottom[0]->shape(softmax_axis_), inner_num_,
                                    top[0]->mutable_gpu_data(), Caffe::stream()));
}

// -------------------------------------------------------------------------------- Eltwise
template <typename Dtype>
void EltwiseLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const EltwiseParameter& p = this->layer_param_.eltwise_param();  // eltwise_layer.cpp:10-28
  CHECK(p.coeff_size() == 0 || p.coeff_size() == (int)bottom.size())
      << "Eltwise Layer takes one coefficient per bottom blob.";
  CHECK(!(p.operation() == EltwiseParameter_EltwiseOp_PROD && p.coeff_size()))
      << "Eltwise layer only takes coefficients for summation.";
  CHECK_LE((int)bottom.size(), MSCNN_MAX_ELTWISE);
  op_ = p.operation() == EltwiseParameter_EltwiseOp_PROD ? MSCNN_ELTWISE_PROD
        : p.operation() == EltwiseParameter_EltwiseOp_MAX ? MSCNN_ELTWISE_MAX
                                                          : MSCNN_ELTWISE_SUM;
  coeffs_.assign(bottom.size(), 1.f);
  for (int i = 0; i < p.coeff_size(); ++i) coeffs_[i] = p.coeff(i);
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  for (size_t i = 1; i < bottom.size(); ++i) CHECK(bottom[i]->shape() == bottom[0]->shape());
  top[0]->ReshapeLike(*bottom[0]);
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
  const float* in[MSCNN_MAX_ELTWISE];
  for (size_t i = 0; i < bottom.size(); ++i) in[i] = bottom[i]->gpu_data();
  MSCNN_CHECK(mscnn_eltwise_forward(in, (int)bottom.size(), op_, coeffs_.data(), (size_t)top[0]->count(),
                                    top[0]->mutable_gpu_data(), Caffe::stream()));
}

INSTANTIATE_CLASS(InputLayer);
INSTANTIATE_CLASS(ConvolutionLayer);
INSTANTIATE_CLASS(DeconvolutionLayer);
INSTANTIATE_CLASS(ReLULayer);
INSTANTIATE_CLASS(PoolingLayer);
INSTANTIATE_CLASS(SplitLayer);
INSTANTIATE_CLASS(ConcatLayer);
INSTANTIATE_CLASS(InnerProductLayer);
INSTANTIATE_CLASS(DropoutLayer);
INSTANTIATE_CLASS(BoxOutputLayer);
INSTANTIATE_CLASS(ROIPoolingLayer);
INSTANTIATE_CLASS(ROIAlignLayer);
INSTANTIATE_CLASS(DecodeBBoxLayer);
INSTANTIATE_CLASS(SoftmaxLayer);
INSTANTIATE_CLASS(EltwiseLayer);
REGISTER_LAYER_CLASS(Input);
REGISTER_LAYER_CLASS(Convolution);
REGISTER_LAYER_CLASS(Deconvolution);
REGISTER_LAYER_CLASS(ReLU);
REGISTER_LAYER_CLASS(Pooling);
REGISTER_LAYER_CLASS(Split);
REGISTER_LAYER_CLASS(Concat);
REGISTER_LAYER_CLASS(InnerProduct);
REGISTER_LAYER_CLASS(Dropout);
REGISTER_LAYER_CLASS(BoxOutput);
REGISTER_LAYER_CLASS(ROIPooling);
REGISTER_LAYER_CLASS(ROIAlign);
REGISTER_LAYER_CLASS(DecodeBBox);
REGISTER_LAYER_CLASS(Softmax);
REGISTER_LAYER_CLASS(Eltwise);

}  // namespace caffe
