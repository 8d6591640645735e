// The small per-ROI layers of the cascade deploy nets (SURVEY.md section 8(f) rank 2;
// examples/kitti_car/cascade-mscnn-7s-576-2x/mscnn_deploy.prototxt:563-949):
//   DecodeBBox  -- decode_bbox_layer.cpp:53-124 + DecodeBBoxesWithPrior (util/math_functions.cpp:46-77)
//   Softmax     -- softmax_layer.cpp:28-62
//   Eltwise     -- eltwise_layer.cpp:46-96
// All three work on R x (a few) fp32 values: they are launch-latency bound, so each is one kernel
// with one thread per row / element, kept on the device only so that the cascade stages never
// round-trip to the host (DecodeBBox is CPU-only in the reference: a D2H + H2D per stage).
// Arithmetic is fp32 in the reference's operation order (no FMA contraction: the library is
// built with -fmad=false); exp() goes through fp64 and is rounded once, see box_output.cu.
#include <cuda_runtime.h>
#include <stdio.h>

#include "mscnn_b200.h"
#include "launch_count.h"

namespace mscnn {

__device__ __forceinline__ float exp_f32(float x) { return static_cast<float>(exp(static_cast<double>(x))); }

struct Stat4 {
  float mean[4], stdv[4];
};

// One thread per ROI.  Output row = [img, x1, y1, x2, y2] of class 1 (decode_bbox_layer.cpp:113-121:
// base_index = keep_id * bbox_dim + 4), TEST phase: every row is kept (:79-106 only filter in TRAIN).
__global__ void decode_bbox_kernel(const float* __restrict__ bbox /*[R][dim]*/, const float* __restrict__ prior /*[R][5]*/,
                                   int R, int dim, const Stat4 st, float* __restrict__ out /*[R][5]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const float* q = prior + (size_t)i * 5;
  const float xmin = q[1], ymin = q[2], xmax = q[3], ymax = q[4];
  // math_functions.cpp:52-57
  const float pw = xmax - xmin + 1.f, ph = ymax - ymin + 1.f;
  const float cx = 0.5f * (xmax + xmin), cy = 0.5f * (ymax + ymin);
  const float* b = bbox + (size_t)i * dim + 4;
  // :60-64 de-normalisation
  const float bx = b[0] * st.stdv[0] + st.mean[0];
  const float by = b[1] * st.stdv[1] + st.mean[1];
  const float bw = b[2] * st.stdv[2] + st.mean[2];
  const float bh = b[3] * st.stdv[3] + st.mean[3];
  // :66-74
  float tx = bx * pw + cx, ty = by * ph + cy;
  const float tw = pw * exp_f32(bw), th = ph * exp_f32(bh);
  tx = tx - (tw - 1.f) / 2.f;
  ty = ty - (th - 1.f) / 2.f;
  float* o = out + (size_t)i * 5;
  o[0] = q[0];
  o[1] = tx;
  o[2] = ty;
  o[3] = tx + tw - 1.f;
  o[4] = ty + th - 1.f;
}

// One thread per (outer, inner) position; channels are walked with stride `inner`.
// softmax_layer.cpp:38-61: subtract the channel max, exp, divide by the channel sum (the
// reference sums with cblas_sgemv, so the summation order is BLAS-defined; here it is 0..C-1).
__global__ void softmax_kernel(const float* __restrict__ x, int outer, int channels, int inner,
                               float* __restrict__ y) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)outer * inner) return;
  const size_t o = idx / inner, k = idx % inner;
  const float* xp = x + o * channels * inner + k;
  float* yp = y + o * channels * inner + k;
  float m = xp[0];
  for (int c = 1; c < channels; ++c) m = fmaxf(m, xp[(size_t)c * inner]);
  float sum = 0.f;
  for (int c = 0; c < channels; ++c) {
    const float e = exp_f32(xp[(size_t)c * inner] - m);
    yp[(size_t)c * inner] = e;
    sum = sum + e;
  }
  for (int c = 0; c < channels; ++c) yp[(size_t)c * inner] = yp[(size_t)c * inner] / sum;
}

struct EltwiseArgs {
  const float* in[MSCNN_MAX_ELTWISE];
  float coeff[MSCNN_MAX_ELTWISE];
  int n;
};

// eltwise_layer.cpp:54-93.  SUM accumulates coeff[i] * bottom[i] onto 0 in bottom order (caffe_axpy).
__global__ void eltwise_kernel(const EltwiseArgs a, int op, size_t count, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    float r;
    if (op == MSCNN_ELTWISE_PROD) {
      r = a.in[0][i] * a.in[1][i];
      for (int b = 2; b < a.n; ++b) r = r * a.in[b][i];
    } else if (op == MSCNN_ELTWISE_SUM) {
      r = 0.f;
      for (int b = 0; b < a.n; ++b) r = r + a.coeff[b] * a.in[b][i];
    } else {
      const float v0 = a.in[0][i], v1 = a.in[1][i];
      r = (v0 > v1) ? v0 : v1;  // :73-80 (ties and NaN pick bottom 1)
      for (int b = 2; b < a.n; ++b) {
        const float v = a.in[b][i];
        if (v > r) r = v;
      }
    }
    y[i] = r;
  }
}

static int check(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "mscnn: %s launch failed: %s\n", what, cudaGetErrorString(e));
    return MSCNN_ERR_CUDA;
  }
  return MSCNN_OK;
}

}  // namespace mscnn

using namespace mscnn;

extern "C" int mscnn_decode_bbox_forward(const float* bbox_pred, const float* prior, int R, int bbox_dim,
                                         const float* mean4, const float* std4, float* out, void* stream) {
  if (!bbox_pred || !prior || !out || R < 0 || bbox_dim < 8 || bbox_dim % 4) return MSCNN_ERR_INVALID;
  if (R == 0) return MSCNN_OK;
  Stat4 st;
  for (int k = 0; k < 4; ++k) {
    st.mean[k] = mean4 ? mean4[k] : 0.f;  // decode_bbox_layer.cpp:33-35 defaults
    st.stdv[k] = std4 ? std4[k] : 1.f;
  }
  mscnn::note_launch();
  decode_bbox_kernel<<<(R + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bbox_pred, prior, R, bbox_dim, st, out);
  return check("decode_bbox");
}

extern "C" int mscnn_softmax_forward(const float* x, int outer, int channels, int inner, float* y, void* stream) {
  if (!x || !y || outer < 0 || channels <= 0 || inner <= 0) return MSCNN_ERR_INVALID;
  const size_t total = (size_t)outer * inner;
  if (total == 0) return MSCNN_OK;
  mscnn::note_launch();
  softmax_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, outer, channels, inner, y);
  return check("softmax");
}

extern "C" int mscnn_eltwise_forward(const float* const* bottoms, int num_bottoms, int op, const float* coeffs,
                                     size_t count, float* y, void* stream) {
  if (!bottoms || !y || num_bottoms < 2 || num_bottoms > MSCNN_MAX_ELTWISE) return MSCNN_ERR_INVALID;
  if (op != MSCNN_ELTWISE_PROD && op != MSCNN_ELTWISE_SUM && op != MSCNN_ELTWISE_MAX) return MSCNN_ERR_INVALID;
  if (count == 0) return MSCNN_OK;
  EltwiseArgs a;
  a.n = num_bottoms;
  for (int b = 0; b < num_bottoms; ++b) {
    if (!bottoms[b]) return MSCNN_ERR_INVALID;
    a.in[b] = bottoms[b];
    a.coeff[b] = coeffs ? coeffs[b] : 1.f;
  }
  size_t blocks = (count + 255) / 256;
  const size_t cap = (size_t)mscnn_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  mscnn::note_launch();
  eltwise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, op, count, y);
  return check("eltwise");
}
