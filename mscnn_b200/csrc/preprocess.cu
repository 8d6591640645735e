// Image pre-processing on the device: the step in front of the path.
//   examples/kitti_car/run_mscnn_detection.m:64-69, examples/widerface/run_mscnn_detection.m:70-86:
//     test_image = imresize(test_image,[imgH imgW]);        uint8 in, uint8 out, bicubic + antialiasing
//     test_image = single(test_image(:,:,[3 2 1]));          RGB -> BGR, to fp32
//     test_image = bsxfun(@minus,test_image,mu);             mu = [104 117 123] per BGR channel
//     test_image = permute(test_image, [2 1 3]);             MATLAB [W H C] column-major = Caffe C x H x W
// The reference does this on the host in MATLAB and uploads 4 bytes per value of the RESIZED image
// (23.6 MB for 3x768x2560); here the ORIGINAL uint8 image (1.4 MB for a 375x1242 KITTI frame) is
// uploaded and both resize passes, the channel swap, the mean subtraction and the layout change run on
// the device, writing straight into the net's `data` blob.
//
// imresize itself is MathWorks code that is not part of /root/reference; its published algorithm
// (imresize.m `contributions`, cubic kernel a = -0.5, Keys 1981) is restated here:
//   scale = out / in per dimension;   antialiasing widens the kernel by 1/scale when scale < 1;
//   u = x/scale + 0.5 (1 - 1/scale), left = floor(u - kw/2), P = ceil(kw) + 2 taps, weights normalised per
//   output sample, out-of-range taps mirrored (symmetric padding), all-zero tap columns dropped;
//   dimensions are resized in ascending order of scale; each pass accumulates in fp64 and, for uint8
//   images, rounds (half away from zero) and saturates back to uint8.
// Tap tables are built on the host in fp64 (mscnn_imresize_contributions, testable without a GPU); the
// kernels multiply and add in the table's tap order without FMA contraction (-fmad=false), so the result
// is bit-identical to the numpy restatement in oracle/port.py.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include <vector>

#include "mscnn_b200.h"
#include "launch_count.h"

namespace mscnn {

namespace {

// Keys cubic convolution kernel, a = -0.5 (imresize.m `cubic`).
double cubic(double x) {
  const double a = fabs(x), a2 = a * a, a3 = a2 * a;
  if (a <= 1.0) return (1.5 * a3 - 2.5 * a2) + 1.0;
  if (a <= 2.0) return ((-0.5 * a3 + 2.5 * a2) - 4.0 * a) + 2.0;
  return 0.0;
}

struct Contrib {
  int taps = 0;
  std::vector<double> w;  // [out_len][taps]
  std::vector<int> idx;   // [out_len][taps], 0-based
};

Contrib contributions(int in_len, int out_len) {
  const double scale = (double)out_len / (double)in_len;
  const bool aa = scale < 1.0;
  const double kw = aa ? 4.0 / scale : 4.0;
  const int P = (int)ceil(kw) + 2;
  std::vector<double> w((size_t)out_len * P);
  std::vector<int> idx((size_t)out_len * P);
  std::vector<char> used(P, 0);
  const int period = 2 * in_len;
  for (int o = 0; o < out_len; ++o) {
    const double x = (double)(o + 1);
    const double u = x / scale + 0.5 * (1.0 - 1.0 / scale);
    const double left = floor(u - kw / 2.0);
    double sum = 0.0;
    for (int k = 0; k < P; ++k) {
      const double ind = left + (double)k;  // 1-based sample position
      const double d = u - ind;
      const double v = aa ? scale * cubic(scale * d) : cubic(d);
      w[(size_t)o * P + k] = v;
      sum += v;
      long m = ((long)ind - 1) % period;
      if (m < 0) m += period;
      idx[(size_t)o * P + k] = (int)(m < in_len ? m : period - 1 - m);
    }
    for (int k = 0; k < P; ++k) {
      w[(size_t)o * P + k] /= sum;
      if (w[(size_t)o * P + k] != 0.0) used[k] = 1;
    }
  }
  Contrib c;
  for (int k = 0; k < P; ++k) c.taps += used[k];
  c.w.resize((size_t)out_len * c.taps);
  c.idx.resize((size_t)out_len * c.taps);
  for (int o = 0; o < out_len; ++o) {
    int t = 0;
    for (int k = 0; k < P; ++k) {
      if (!used[k]) continue;
      c.w[(size_t)o * c.taps + t] = w[(size_t)o * P + k];
      c.idx[(size_t)o * c.taps + t] = idx[(size_t)o * P + k];
      ++t;
    }
  }
  return c;
}

struct DevTable {
  int taps = 0;
  double* w = nullptr;
  int* idx = nullptr;
};

}  // namespace

struct PreprocessPlan {
  mscnn_preprocess_desc d;
  int first_dim;  // 0: height pass first, 1: width pass first (ascending scale, ties -> height)
  DevTable th, tw;
  unsigned char* mid = nullptr;  // intermediate image after the first pass
  size_t mid_cap = 0;
  unsigned char* stage = nullptr;  // device copy of host images (mscnn_preprocess_forward_host)
  size_t stage_cap = 0;
};

// uint8 conversion of an fp64 accumulator: saturate, round half away from zero.
__device__ __forceinline__ double round_u8(double v) {
  v = fmin(fmax(v, 0.0), 255.0);
  return round(v);
}

// First pass: resize along one dimension, uint8 [N][H][W][3] -> uint8.  One thread per output byte
// (the three channels of a pixel are adjacent bytes, so a warp writes 32 consecutive bytes and reads
// `taps` rows or pixel neighbourhoods that its lanes share).
//   along = 0: out[n][y][x][c] = sum_k w[y][k] * in[n][idx[y][k]][x][c]   (out_h x in_w)
//   along = 1: out[n][y][x][c] = sum_k w[x][k] * in[n][y][idx[x][k]][c]   (in_h x out_w)
__global__ void imresize_pass_u8_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                        int N, int in_h, int in_w, int out_h, int out_w, int along, int taps,
                                        const double* __restrict__ w, const int* __restrict__ idx) {
  const size_t total = (size_t)N * out_h * out_w * 3;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % 3);
    size_t p = e / 3;
    const int x = (int)(p % out_w);
    p /= out_w;
    const int y = (int)(p % out_h);
    const int n = (int)(p / out_h);
    const unsigned char* img = in + (size_t)n * in_h * in_w * 3;
    const int o = along == 0 ? y : x;
    const double* wr = w + (size_t)o * taps;
    const int* ir = idx + (size_t)o * taps;
    double acc = 0.0;
    for (int k = 0; k < taps; ++k) {
      const int s = ir[k];
      const size_t off = along == 0 ? ((size_t)s * in_w + x) * 3 + c : ((size_t)y * in_w + s) * 3 + c;
      acc = __dadd_rn(acc, __dmul_rn(wr[k], (double)img[off]));
    }
    out[e] = (unsigned char)round_u8(acc);
  }
}

// Second pass + channel swap + mean subtraction + HWC -> CHW.  One thread per output pixel; a warp
// writes 32 consecutive floats of each of the three output planes.
__global__ void imresize_pass_to_blob_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int N,
                                             int in_h, int in_w, int out_h, int out_w, int along, int taps,
                                             const double* __restrict__ w, const int* __restrict__ idx,
                                             float mean0, float mean1, float mean2, int swap_rb) {
  const size_t total = (size_t)N * out_h * out_w;
  const size_t plane = (size_t)out_h * out_w;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % out_w);
    size_t p = e / out_w;
    const int y = (int)(p % out_h);
    const int n = (int)(p / out_h);
    const unsigned char* img = in + (size_t)n * in_h * in_w * 3;
    const int o = along == 0 ? y : x;
    const double* wr = w + (size_t)o * taps;
    const int* ir = idx + (size_t)o * taps;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < taps; ++k) {
      const int s = ir[k];
      const unsigned char* q = img + (along == 0 ? ((size_t)s * in_w + x) * 3 : ((size_t)y * in_w + s) * 3);
      const double wk = wr[k];
      a0 = __dadd_rn(a0, __dmul_rn(wk, (double)q[0]));
      a1 = __dadd_rn(a1, __dmul_rn(wk, (double)q[1]));
      a2 = __dadd_rn(a2, __dmul_rn(wk, (double)q[2]));
    }
    const float v0 = (float)round_u8(a0), v1 = (float)round_u8(a1), v2 = (float)round_u8(a2);
    float* dst = out + (size_t)n * 3 * plane + (size_t)y * out_w + x;
    // output channel order after test_image(:,:,[3 2 1]): plane 0 = input channel 2 when swapping
    dst[0] = (swap_rb ? v2 : v0) - mean0;
    dst[plane] = v1 - mean1;
    dst[2 * plane] = (swap_rb ? v0 : v2) - mean2;
  }
}

// No resize needed in a dimension (out == in): imresize still runs the pass, and the cubic kernel at
// integer offsets is exactly the identity, so the general kernels are used unchanged.

static int upload(const Contrib& c, DevTable* t) {
  t->taps = c.taps;
  if (cudaMalloc(&t->w, c.w.size() * sizeof(double)) != cudaSuccess) return MSCNN_ERR_NOMEM;
  if (cudaMalloc(&t->idx, c.idx.size() * sizeof(int)) != cudaSuccess) return MSCNN_ERR_NOMEM;
  if (cudaMemcpy(t->w, c.w.data(), c.w.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t->idx, c.idx.data(), c.idx.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
    return MSCNN_ERR_CUDA;
  return MSCNN_OK;
}

static int grow(unsigned char** buf, size_t* cap, size_t need) {
  if (need <= *cap) return MSCNN_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  if (cudaMalloc(buf, need) != cudaSuccess) return MSCNN_ERR_NOMEM;
  *cap = need;
  return MSCNN_OK;
}

}  // namespace mscnn

using namespace mscnn;

extern "C" {

int mscnn_imresize_taps(int in_len, int out_len) {
  if (in_len < 1 || out_len < 1) return MSCNN_ERR_INVALID;
  return contributions(in_len, out_len).taps;
}

int mscnn_imresize_contributions(int in_len, int out_len, double* host_weights, int* host_indices, int cap_taps) {
  if (in_len < 1 || out_len < 1 || !host_weights || !host_indices) return MSCNN_ERR_INVALID;
  const Contrib c = contributions(in_len, out_len);
  if (c.taps > cap_taps) return MSCNN_ERR_INVALID;
  for (int o = 0; o < out_len; ++o)
    for (int k = 0; k < c.taps; ++k) {
      host_weights[(size_t)o * cap_taps + k] = c.w[(size_t)o * c.taps + k];
      host_indices[(size_t)o * cap_taps + k] = c.idx[(size_t)o * c.taps + k];
    }
  return c.taps;
}

int mscnn_widerface_net_size(int org_h, int org_w, int img_h, int img_w, int max_size, int* rz_h, int* rz_w) {
  if (org_h < 1 || org_w < 1 || !rz_h || !rz_w || max_size < 32) return MSCNN_ERR_INVALID;
  // examples/widerface/run_mscnn_detection.m:72-80 (MATLAB round = half away from zero)
  double w = img_w == 0 ? org_w : img_w, h = img_h == 0 ? org_h : img_h;
  w = round(w / 32.0) * 32.0;
  h = round(h / 32.0) * 32.0;
  if (h > max_size || w > max_size) {
    const double r = (double)max_size / fmax(h, w);
    h = round(h * r / 32.0) * 32.0;
    w = round(w * r / 32.0) * 32.0;
  }
  *rz_h = (int)h;
  *rz_w = (int)w;
  return MSCNN_OK;
}

int mscnn_preprocess_create(const mscnn_preprocess_desc* d, void** plan) {
  if (!d || !plan || d->in_h < 1 || d->in_w < 1 || d->out_h < 1 || d->out_w < 1) return MSCNN_ERR_INVALID;
  PreprocessPlan* p = new PreprocessPlan();
  p->d = *d;
  const double sh = (double)d->out_h / d->in_h, sw = (double)d->out_w / d->in_w;
  p->first_dim = sw < sh ? 1 : 0;  // ascending scale, stable: height first on ties (imresize.m sort(scale))
  int rc = upload(contributions(d->in_h, d->out_h), &p->th);
  if (rc == MSCNN_OK) rc = upload(contributions(d->in_w, d->out_w), &p->tw);
  if (rc != MSCNN_OK) {
    mscnn_preprocess_destroy(p);
    return rc;
  }
  *plan = p;
  return MSCNN_OK;
}

int mscnn_preprocess_get_desc(void* plan, mscnn_preprocess_desc* d) {
  if (!plan || !d) return MSCNN_ERR_INVALID;
  *d = static_cast<PreprocessPlan*>(plan)->d;
  return MSCNN_OK;
}

int mscnn_preprocess_destroy(void* plan) {
  PreprocessPlan* p = static_cast<PreprocessPlan*>(plan);
  if (!p) return MSCNN_OK;
  cudaFree(p->th.w);
  cudaFree(p->th.idx);
  cudaFree(p->tw.w);
  cudaFree(p->tw.idx);
  cudaFree(p->mid);
  cudaFree(p->stage);
  delete p;
  return MSCNN_OK;
}

int mscnn_preprocess_forward(void* plan, int N, const unsigned char* images, float* data, void* stream) {
  PreprocessPlan* p = static_cast<PreprocessPlan*>(plan);
  if (!p || N < 1 || !images || !data) return MSCNN_ERR_INVALID;
  const mscnn_preprocess_desc& d = p->d;
  cudaStream_t st = (cudaStream_t)stream;
  // sizes after the first pass
  const int mid_h = p->first_dim == 0 ? d.out_h : d.in_h;
  const int mid_w = p->first_dim == 0 ? d.in_w : d.out_w;
  const size_t mid_bytes = (size_t)N * mid_h * mid_w * 3;
  int rc = grow(&p->mid, &p->mid_cap, mid_bytes);
  if (rc != MSCNN_OK) return rc;
  const int threads = 256;
  const DevTable& t1 = p->first_dim == 0 ? p->th : p->tw;
  const DevTable& t2 = p->first_dim == 0 ? p->tw : p->th;
  {
    const size_t blocks = (mid_bytes + threads - 1) / threads;
    imresize_pass_u8_kernel<<<(unsigned)(blocks > 148u * 64 ? 148u * 64 : blocks), threads, 0, st>>>(
        images, p->mid, N, d.in_h, d.in_w, mid_h, mid_w, p->first_dim, t1.taps, t1.w, t1.idx);
    note_launch();
  }
  {
    const size_t px = (size_t)N * d.out_h * d.out_w;
    const size_t blocks = (px + threads - 1) / threads;
    imresize_pass_to_blob_kernel<<<(unsigned)(blocks > 148u * 64 ? 148u * 64 : blocks), threads, 0, st>>>(
        p->mid, data, N, mid_h, mid_w, d.out_h, d.out_w, 1 - p->first_dim, t2.taps, t2.w, t2.idx, d.mean[0],
        d.mean[1], d.mean[2], d.swap_rb);
    note_launch();
  }
  return cudaPeekAtLastError() == cudaSuccess ? MSCNN_OK : MSCNN_ERR_CUDA;
}

int mscnn_preprocess_forward_host(void* plan, int N, const unsigned char* host_images, float* data, void* stream) {
  PreprocessPlan* p = static_cast<PreprocessPlan*>(plan);
  if (!p || N < 1 || !host_images || !data) return MSCNN_ERR_INVALID;
  const size_t bytes = (size_t)N * p->d.in_h * p->d.in_w * 3;
  int rc = grow(&p->stage, &p->stage_cap, bytes);
  if (rc != MSCNN_OK) return rc;
  if (cudaMemcpyAsync(p->stage, host_images, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess)
    return MSCNN_ERR_CUDA;
  return mscnn_preprocess_forward(plan, N, p->stage, data, stream);
}

}  // extern "C"
