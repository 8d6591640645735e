// HBM-bound plane kernels: Pooling, ROIPooling (+ fused Concat), depthwise 2x Deconvolution.
// All work on NHWC bf16 planes with 16-byte (8-channel) vector accesses, so a warp touches
// 512 contiguous bytes per pixel; values of split tensors are hi + lo evaluated in fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mscnn_b200.h"
#include "launch_count.h"

namespace mscnn {

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off) {
  Vec8 r;
  const uint4 a = *reinterpret_cast<const uint4*>(hi + off);
  const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __bfloat162float(ap[i].x);
    r.v[2 * i + 1] = __bfloat162float(ap[i].y);
  }
  if (lo) {
    const uint4 b = *reinterpret_cast<const uint4*>(lo + off);
    const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] += __bfloat162float(bp[i].x);
      r.v[2 * i + 1] += __bfloat162float(bp[i].y);
    }
  }
  return r;
}

// STREAM: st.global.cs (evict-first).  ROI pooling writes 2.6 GB per batch while gathering from a 63 MB per-image
// feature map that should stay in the 126 MB L2; with default stores the output stream evicted it and DRAM
// read 3.7 GB per launch instead of 0.5 GB (profiles/r01h_summary.md).
template <bool STREAM = false>
__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const Vec8& x) {
  uint4 a, b;
  __nv_bfloat162* ap = reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162* bp = reinterpret_cast<__nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x.v[2 * i]), h1 = __float2bfloat16_rn(x.v[2 * i + 1]);
    ap[i] = __nv_bfloat162(h0, h1);
    bp[i] = __nv_bfloat162(__float2bfloat16_rn(x.v[2 * i] - __bfloat162float(h0)),
                           __float2bfloat16_rn(x.v[2 * i + 1] - __bfloat162float(h1)));
  }
  if (STREAM) {
    __stcs(reinterpret_cast<uint4*>(hi + off), a);
    if (lo) __stcs(reinterpret_cast<uint4*>(lo + off), b);
  } else {
    *reinterpret_cast<uint4*>(hi + off) = a;
    if (lo) *reinterpret_cast<uint4*>(lo + off) = b;
  }
}

// ------------------------------------------------------------------------------- Pooling
// PoolingLayer::Forward_cpu (src/caffe/layers/pooling_layer.cpp:128-220) with pad = 0:
// output size in ceil mode (:90-93), windows clipped to the image, AVE divides by the clipped
// window size (pad 0 => pool_size = (hend-hstart)*(wend-wstart), :196-203).
__global__ void pool_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                            __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl, int N, int H,
                            int W, int C, int Ho, int Wo, int k, int s, int mode, const int* __restrict__ dyn_n) {
  const int cg = C / 8;
  if (dyn_n) N = max(0, min(N, *dyn_n));  // data-dependent row count (mscnn_pool_forward_dyn)
  const size_t total = (size_t)N * Ho * Wo * cg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int g = idx % cg;
    size_t r = idx / cg;
    const int ow = r % Wo; r /= Wo;
    const int oh = r % Ho;
    const int n = r / Ho;
    const int hs = oh * s, ws = ow * s;
    const int he = min(hs + k, H), we = min(ws + k, W);
    Vec8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = (mode == MSCNN_POOL_MAX) ? -3.402823466e+38f : 0.f;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const Vec8 v = load8(xh, xl, ((size_t)(n * H + h) * W + w) * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (mode == MSCNN_POOL_MAX) acc.v[i] = (v.v[i] > acc.v[i]) ? v.v[i] : acc.v[i];
          else acc.v[i] = acc.v[i] + v.v[i];
        }
      }
    if (mode == MSCNN_POOL_AVE) {
      const float sz = (float)((he - hs) * (we - ws));
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = acc.v[i] / sz;
    }
    store8(yh, yl, ((size_t)(n * Ho + oh) * Wo + ow) * C + g * 8, acc);
  }
}

// --------------------------------------------------------------------------- ROIPooling
// ROIPoolingLayer::Forward_cpu (src/caffe/layers/roi_pooling_layer.cpp:49-139) including the
// MS-CNN pad_ratio context extension (:66-72).  Output rows are written at channel offset
// `c_off` of a [R][P][P][Cout_total] tensor, which fuses ConcatLayer (concat_layer.cpp:57-74).
// One warp per (ROI, output bin): the bin geometry is computed once per warp (uniform), lanes
// stride over the 8-channel groups so that every load / store is a 512-byte coalesced row segment.
// Up to four pad_ratio variants of the same ROI set are pooled by the same warp (MS-CNN pools every
// ROI twice, object and 1.5x context window, and concatenates): the context window contains the
// object window, so the second variant mostly hits lines the first one just touched.
// best[i] = max(best[i], hi[i] + lo[i]) for the 8 bf16 channels packed in `a` (+ `l` when has_lo)
__device__ __forceinline__ void max8(Vec8& best, const uint4& a, const uint4& l, bool has_lo) {
  const uint32_t ah[4] = {a.x, a.y, a.z, a.w}, al[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v0 = __uint_as_float(ah[q] << 16), v1 = __uint_as_float(ah[q] & 0xFFFF0000u);
    if (has_lo) {
      v0 = v0 + __uint_as_float(al[q] << 16);
      v1 = v1 + __uint_as_float(al[q] & 0xFFFF0000u);
    }
    best.v[2 * q] = fmaxf(best.v[2 * q], v0);
    best.v[2 * q + 1] = fmaxf(best.v[2 * q + 1], v1);
  }
}

struct RoiVariants {
  int count;
  float pad_ratio[4];
  int c_off[4];
};

// Pools one (ROI, bin, variant) window for the channel groups g0 + lane (+ 32 when DUAL) and stores the result.
// The window is walked as one flattened pixel sequence in 16-byte units with 32-bit indices (the feature map has
// < 2^31 / 8 elements: checked by the launcher), two pixels per step, so up to eight 16-byte loads (hi, lo, second
// group) are in flight per lane: the kernel was bound by the latency of two (profiles/r01h_summary.md).  fmaxf gives the
// reference's `if (x > max) max = x` result for every non-NaN input.  Stores are streaming (st.global.cs).
template <bool DUAL>
__device__ __forceinline__ void roi_pool_window(const uint4* __restrict__ xh4, const uint4* __restrict__ xl4, int base_idx,
                                                int npx, int bw, int row_skip, int cg, bool empty,
                                                __nv_bfloat16* yh, __nv_bfloat16* yl, size_t out_off) {
  const bool has_lo = xl4 != nullptr;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  Vec8 best0, best1;
#pragma unroll
  for (int i = 0; i < 8; ++i) best0.v[i] = best1.v[i] = empty ? 0.f : -3.402823466e+38f;
  int idx = base_idx;
  int wpos = 0;
  int k = 0;
  for (; k + 2 <= npx; k += 2) {
    int ix[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ix[u] = idx;
      idx += cg;
      if (++wpos == bw) { wpos = 0; idx += row_skip; }
    }
    uint4 a[2], l[2], c[2], m[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      a[u] = __ldg(xh4 + ix[u]);
      l[u] = has_lo ? __ldg(xl4 + ix[u]) : z;
      if (DUAL) {
        c[u] = __ldg(xh4 + ix[u] + 32);
        m[u] = has_lo ? __ldg(xl4 + ix[u] + 32) : z;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      max8(best0, a[u], l[u], has_lo);
      if (DUAL) max8(best1, c[u], m[u], has_lo);
    }
  }
  if (k < npx) {
    max8(best0, __ldg(xh4 + idx), has_lo ? __ldg(xl4 + idx) : z, has_lo);
    if (DUAL) max8(best1, __ldg(xh4 + idx + 32), has_lo ? __ldg(xl4 + idx + 32) : z, has_lo);
  }
  store8<true>(yh, yl, out_off, best0);
  if (DUAL) store8<true>(yh, yl, out_off + 32 * 8, best1);
}

// One warp per (ROI, bin); a lane owns channel groups g and g + 32 (C = 512: all 64 groups in one visit).
// Schedules that were measured and dropped (profiles/r02_summary.md): per-image half-channel passes (halves the L2
// working set: DRAM reads 3.27 -> 2.83 GB but 1.74 -> 2.15 ms), block-per-ROI (L1 sharing between the bins of one ROI:
// 1.79 -> 2.02 ms), createpolicy evict_last loads / evict_first stores (no change).
__global__ void roi_pool_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                const float* __restrict__ rois, int R, int N, int H, int W, int C,
                                int PH, int PW, float scale, const RoiVariants var,
                                __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl, int Ctot,
                                const int* __restrict__ dyn_R) {
  const int cg = C / 8;
  const int lane = threadIdx.x & 31;
  const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
  if (dyn_R) R = max(0, min(R, *dyn_R));  // data-dependent ROI count (mscnn_roi_pool_multi_forward_dyn)
  const size_t bins = (size_t)R * PH * PW;
  const uint4* xh4 = reinterpret_cast<const uint4*>(xh);
  const uint4* xl4 = reinterpret_cast<const uint4*>(xl);
  for (size_t bin = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bin < bins; bin += warps_total) {
    const int pw = bin % PW;
    const int ph = (bin / PW) % PH;
    const int roi = bin / ((size_t)PW * PH);
    const float* q = rois + (size_t)roi * 5;
    int b = (int)q[0];
    b = min(max(b, 0), N - 1);  // the reference CHECKs the range (:63-64); never out of range here
    const float q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
    for (int vi = 0; vi < var.count; ++vi) {
      const float pad_w = (q3 - q1 + 1.f) * var.pad_ratio[vi];
      const float pad_h = (q4 - q2 + 1.f) * var.pad_ratio[vi];
      const int sw = (int)roundf((q1 - pad_w) * scale);
      const int sh = (int)roundf((q2 - pad_h) * scale);
      const int ew = (int)roundf((q3 + pad_w) * scale);
      const int eh = (int)roundf((q4 + pad_h) * scale);
      const int roi_h = max(eh - sh + 1, 1), roi_w = max(ew - sw + 1, 1);
      const float bin_h = (float)roi_h / (float)PH, bin_w = (float)roi_w / (float)PW;
      int hstart = (int)floorf((float)ph * bin_h);
      int wstart = (int)floorf((float)pw * bin_w);
      int hend = (int)ceilf((float)(ph + 1) * bin_h);
      int wend = (int)ceilf((float)(pw + 1) * bin_w);
      hstart = min(max(hstart + sh, 0), H);
      hend = min(max(hend + sh, 0), H);
      wstart = min(max(wstart + sw, 0), W);
      wend = min(max(wend + sw, 0), W);
      const bool empty = (hend <= hstart) || (wend <= wstart);
      const int bw = wend - wstart;
      const int npx = empty ? 0 : (hend - hstart) * bw;
      const int row_skip = (W - bw) * cg;
      const int base = ((b * H + hstart) * W + wstart) * cg;
      const size_t out = (size_t)bin * Ctot + var.c_off[vi];
      for (int g = lane; g < cg; g += 64) {
        if (g + 32 < cg) roi_pool_window<true>(xh4, xl4, base + g, npx, bw, row_skip, cg, empty, yh, yl, out + g * 8);
        else roi_pool_window<false>(xh4, xl4, base + g, npx, bw, row_skip, cg, empty, yh, yl, out + g * 8);
      }
    }
  }
}

// ------------------------------------------------------------------------------ ROIAlign
// ROIAlignLayer::Forward_cpu (src/caffe/layers/roi_align_layer.cpp:49-139; GPU twin
// roi_align_layer.cu:21-98): bilinear samples on the (PH+1) x (PW+1) GRID of bin corners of the
// (padded) ROI, in feature-map coordinates shifted by half a pixel.  Same work distribution as
// roi_pool_kernel: one warp per (ROI, grid point), lanes over 8-channel groups, so the four
// neighbour loads are coalesced 512-byte row segments.  fp32 arithmetic in the reference's order.
__global__ void roi_align_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                 const float* __restrict__ rois, int R, int N, int H, int W, int C, int PH,
                                 int PW, float scale, float pad_ratio, __nv_bfloat16* __restrict__ yh,
                                 __nv_bfloat16* __restrict__ yl, int Ctot, int c_off, const int* __restrict__ dyn_R) {
  const int cg = C / 8;
  const int GH = PH + 1, GW = PW + 1;
  const int lane = threadIdx.x & 31;
  const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
  if (dyn_R) R = max(0, min(R, *dyn_R));
  const size_t points = (size_t)R * GH * GW;
  for (size_t pt = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pt < points; pt += warps_total) {
    const int pw = pt % GW;
    const int ph = (pt / GW) % GH;
    const int roi = pt / ((size_t)GW * GH);
    const float* q = rois + (size_t)roi * 5;
    int b = (int)q[0];
    b = min(max(b, 0), N - 1);  // CHECKed by the reference (:63-64)
    // :67-81
    const float pad_w = (q[3] - q[1] + 1.f) * pad_ratio;
    const float pad_h = (q[4] - q[2] + 1.f) * pad_ratio;
    const float start_w = (q[1] - pad_w) * scale - 0.5f;
    const float start_h = (q[2] - pad_h) * scale - 0.5f;
    const float end_w = (q[3] + pad_w) * scale - 0.5f;
    const float end_h = (q[4] + pad_h) * scale - 0.5f;
    const float roi_h = end_h - start_h, roi_w = end_w - start_w;
    const float bin_h = roi_h / (float)PH, bin_w = roi_w / (float)PW;
    float hf = start_h + (float)ph * bin_h;
    float wf = start_w + (float)pw * bin_w;
    // malformed ROI (:94-97) or grid point outside the map (:104-108) -> 0
    const bool zero = (roi_h <= 0.f || roi_w <= 0.f) || hf < -0.5f || (double)hf > (double)H - 0.5 ||
                      wf < -0.5f || (double)wf > (double)W - 0.5;
    int h0 = (int)floorf(hf), w0 = (int)floorf(wf);
    int h1 = h0 + 1, w1 = w0 + 1;
    hf = fminf(fmaxf(hf, 0.f), (float)(H - 1));
    wf = fminf(fmaxf(wf, 0.f), (float)(W - 1));
    h0 = min(max(h0, 0), H - 1); w0 = min(max(w0, 0), W - 1);
    h1 = min(max(h1, 0), H - 1); w1 = min(max(w1, 0), W - 1);
    const float lh = hf - (float)h0, lw = wf - (float)w0;
    const float hh = 1.f - lh, hw = 1.f - lw;
    const float w00 = hw * hh, w10 = lw * hh, w01 = hw * lh, w11 = lw * lh;
    const size_t img = (size_t)b * H;
    for (int g = lane; g < cg; g += 32) {
      Vec8 r;
      if (zero) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
      } else {
        const Vec8 v00 = load8(xh, xl, ((img + h0) * W + w0) * C + g * 8);
        const Vec8 v10 = load8(xh, xl, ((img + h0) * W + w1) * C + g * 8);
        const Vec8 v01 = load8(xh, xl, ((img + h1) * W + w0) * C + g * 8);
        const Vec8 v11 = load8(xh, xl, ((img + h1) * W + w1) * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i)  // :131 left-to-right
          r.v[i] = ((w00 * v00.v[i] + w10 * v10.v[i]) + w01 * v01.v[i]) + w11 * v11.v[i];
      }
      store8(yh, yl, pt * Ctot + c_off + g * 8, r);
    }
  }
}

// ------------------------------------------------------------ depthwise Deconvolution 2x
// DeconvolutionLayer::Forward_cpu (src/caffe/layers/deconv_layer.cpp:25-40) for the shape the
// MS-CNN "-2x" nets use: group == channels, kernel 4, stride 2, pad 1, no bias
// (examples/kitti_car/mscnn-7s-576-2x/mscnn_deploy.prototxt:452-466):
//   y[n, c, 2*iy - 1 + ky, 2*ix - 1 + kx] += x[n, c, iy, ix] * w[c, ky, kx]
// Gather form: each output pixel has exactly 2 x 2 contributing taps.  Weights are read from
// the layer blob (the bilinear filler is just the usual content, filler.hpp:248-258).
__global__ void deconv2x_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                const float* __restrict__ wt /*[C][4][4]*/, __nv_bfloat16* __restrict__ yh,
                                __nv_bfloat16* __restrict__ yl, int N, int H, int W, int C, int Creal) {
  // The 16 taps of every channel, staged once per block as w_s[(tap * 8 + i) * cg + g] (channel c = 8 g + i): a warp
  // reads consecutive g, so the reads are conflict-free.  (Reading wt[(c * 4 + ky) * 4 + kx] from global memory cost 32
  // scattered 4-byte loads per thread: 3.2 ms for the 1 GB of planes this layer moves on the WIDER net, 20x its
  // HBM time.)
  extern __shared__ float w_s[];
  const int cg = C / 8;
  for (int j = threadIdx.x; j < 16 * C; j += blockDim.x) {
    const int g = j % cg, i = (j / cg) % 8, tap = j / (8 * cg);
    const int c = g * 8 + i;
    w_s[j] = (c < Creal) ? wt[c * 16 + tap] : 0.f;
  }
  __syncthreads();
  const int Ho = 2 * H, Wo = 2 * W;
  const size_t total = (size_t)N * Ho * Wo * cg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int g = idx % cg;
    size_t r = idx / cg;
    const int ox = r % Wo; r /= Wo;
    const int oy = r % Ho;
    const int n = r / Ho;
    Vec8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
    // ky has the parity of oy+1; iy = (oy + 1 - ky) / 2.  The reference's col2im adds the
    // kernel offsets in increasing (ky, kx) order (util/im2col.cpp:col2im_cpu), so do we.
    for (int ky = (oy + 1) & 1; ky < 4; ky += 2) {
      const int iy = (oy + 1 - ky) / 2;
      if (oy + 1 - ky < 0 || iy >= H) continue;
      for (int kx = (ox + 1) & 1; kx < 4; kx += 2) {
        const int ix = (ox + 1 - kx) / 2;
        if (ox + 1 - kx < 0 || ix >= W) continue;
        const Vec8 v = load8(xh, xl, ((size_t)(n * H + iy) * W + ix) * C + g * 8);
        const float* wp = w_s + (size_t)((ky * 4 + kx) * 8) * cg + g;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc.v[i] = acc.v[i] + v.v[i] * wp[i * cg];
      }
    }
    store8(yh, yl, ((size_t)(n * Ho + oy) * Wo + ox) * C + g * 8, acc);
  }
}

// ------------------------------------------------------------------ stand-alone ReLU / Concat
// Used only when a ReLU / Concat layer is run unfused (outside mscnn_b200's Net, which folds
// ReLU into the producing convolution and Concat into ROIPooling).
// relu(hi + lo): |lo| <= ulp(hi)/2, so the sign of the sum is the sign of hi.
__global__ void relu_planes_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 a = reinterpret_cast<uint4*>(hi)[i];
    uint4 b = lo ? reinterpret_cast<uint4*>(lo)[i] : make_uint4(0, 0, 0, 0);
    uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
    uint32_t* bp = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (ap[k] & 0x00008000u) { ap[k] &= 0xFFFF0000u; bp[k] &= 0xFFFF0000u; }   // low element negative
      if (ap[k] & 0x80000000u) { ap[k] &= 0x0000FFFFu; bp[k] &= 0x0000FFFFu; }   // high element negative
    }
    reinterpret_cast<uint4*>(hi)[i] = a;
    if (lo) reinterpret_cast<uint4*>(lo)[i] = b;
  }
}
__global__ void relu_f32_kernel(float* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = fmaxf(x[i], 0.f);
}
// y[p][off .. off+C) = x[p][0 .. C) for every pixel p (channel concat of NHWC planes)
__global__ void concat_planes_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     size_t pixels, int C, int Ctot, int off) {
  const int cg = C / 8;
  const size_t total = pixels * cg;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / cg;
    const int g = i % cg;
    *reinterpret_cast<uint4*>(y + p * Ctot + off + g * 8) = *reinterpret_cast<const uint4*>(x + p * C + g * 8);
  }
}

static int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "mscnn: %s launch failed: %s\n", what, cudaGetErrorString(e));
    return MSCNN_ERR_CUDA;
  }
  return MSCNN_OK;
}

static int grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = (size_t)mscnn_sm_count() * 16;  // 16 resident 256-thread CTAs per SM at most
  if (b > cap) b = cap;
  return (int)(b ? b : 1);
}

}  // namespace mscnn

using namespace mscnn;

extern "C" int mscnn_pool_forward_dyn(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, int N, int H,
                                      int W, int C, int kernel, int stride, int mode, const int* dyn_n, void* stream) {
  if (!x_hi || !y_hi || N <= 0 || H <= 0 || W <= 0 || C % 8 || kernel <= 0 || stride <= 0)
    return MSCNN_ERR_INVALID;
  if ((x_lo == nullptr) != (y_lo == nullptr)) return MSCNN_ERR_INVALID;
  if (mode != MSCNN_POOL_MAX && mode != MSCNN_POOL_AVE) return MSCNN_ERR_INVALID;
  // pooling_layer.cpp:90-93 with pad 0
  const int Ho = (H - kernel + stride - 1) / stride + 1;
  const int Wo = (W - kernel + stride - 1) / stride + 1;
  const size_t total = (size_t)N * Ho * Wo * (C / 8);
  mscnn::note_launch();
  pool_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo, N,
      H, W, C, Ho, Wo, kernel, stride, mode, dyn_n);
  return launch_check("pool");
}

extern "C" int mscnn_pool_forward(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, int N, int H,
                                  int W, int C, int kernel, int stride, int mode, void* stream) {
  return mscnn_pool_forward_dyn(x_hi, x_lo, y_hi, y_lo, N, H, W, C, kernel, stride, mode, nullptr, stream);
}

extern "C" int mscnn_roi_pool_multi_forward_dyn(const void* x_hi, const void* x_lo, int N, int H, int W, int C,
                                                const float* rois, int R, int pooled_h, int pooled_w,
                                                float spatial_scale, int num_variants, const float* pad_ratios,
                                                const int* out_channel_offsets, void* y_hi, void* y_lo,
                                                int out_channels_total, const int* dyn_R, void* stream) {
  if (!x_hi || !y_hi || !rois || !pad_ratios || !out_channel_offsets || N <= 0 || C % 8 || R < 0 ||
      pooled_h <= 0 || pooled_w <= 0 || out_channels_total % 8 || num_variants < 1 || num_variants > 4)
    return MSCNN_ERR_INVALID;
  RoiVariants var;
  var.count = num_variants;
  for (int i = 0; i < num_variants; ++i) {
    if (out_channel_offsets[i] % 8 || out_channel_offsets[i] + C > out_channels_total) return MSCNN_ERR_INVALID;
    var.pad_ratio[i] = pad_ratios[i];
    var.c_off[i] = out_channel_offsets[i];
  }
  if (R == 0) return MSCNN_OK;
  if ((size_t)N * H * W * (C / 8) >= (size_t)1 << 31) return MSCNN_ERR_INVALID;  // 32-bit 16-byte indices in the kernel
  const size_t total = (size_t)R * pooled_h * pooled_w * 32;  // one warp per (ROI, bin)
  mscnn::note_launch();
  roi_pool_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, rois, R, N, H, W, C, pooled_h, pooled_w,
      spatial_scale, var, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo, out_channels_total, dyn_R);
  return launch_check("roi_pool");
}

extern "C" int mscnn_roi_pool_multi_forward(const void* x_hi, const void* x_lo, int N, int H, int W, int C,
                                            const float* rois, int R, int pooled_h, int pooled_w,
                                            float spatial_scale, int num_variants, const float* pad_ratios,
                                            const int* out_channel_offsets, void* y_hi, void* y_lo,
                                            int out_channels_total, void* stream) {
  return mscnn_roi_pool_multi_forward_dyn(x_hi, x_lo, N, H, W, C, rois, R, pooled_h, pooled_w, spatial_scale,
                                          num_variants, pad_ratios, out_channel_offsets, y_hi, y_lo, out_channels_total,
                                          nullptr, stream);
}

extern "C" int mscnn_roi_pool_forward(const void* x_hi, const void* x_lo, int N, int H, int W, int C,
                                      const float* rois, int R, int pooled_h, int pooled_w,
                                      float spatial_scale, float pad_ratio, void* y_hi, void* y_lo,
                                      int out_channels_total, int out_channel_offset, void* stream) {
  return mscnn_roi_pool_multi_forward(x_hi, x_lo, N, H, W, C, rois, R, pooled_h, pooled_w, spatial_scale, 1,
                                      &pad_ratio, &out_channel_offset, y_hi, y_lo, out_channels_total, stream);
}

extern "C" int mscnn_roi_align_forward_dyn(const void* x_hi, const void* x_lo, int N, int H, int W, int C,
                                           const float* rois, int R, int pooled_h, int pooled_w,
                                           float spatial_scale, float pad_ratio, void* y_hi, void* y_lo,
                                           int out_channels_total, int out_channel_offset, const int* dyn_R,
                                           void* stream) {
  if (!x_hi || !y_hi || !rois || N <= 0 || H <= 0 || W <= 0 || C % 8 || R < 0 || pooled_h <= 0 || pooled_w <= 0 ||
      out_channels_total % 8 || out_channel_offset % 8 || out_channel_offset + C > out_channels_total)
    return MSCNN_ERR_INVALID;
  if ((x_lo == nullptr) != (y_lo == nullptr)) return MSCNN_ERR_INVALID;
  if (R == 0) return MSCNN_OK;
  const size_t total = (size_t)R * (pooled_h + 1) * (pooled_w + 1) * 32;  // one warp per (ROI, grid point)
  mscnn::note_launch();
  roi_align_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, rois, R, N, H, W, C, pooled_h, pooled_w,
      spatial_scale, pad_ratio, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo, out_channels_total,
      out_channel_offset, dyn_R);
  return launch_check("roi_align");
}

extern "C" int mscnn_roi_align_forward(const void* x_hi, const void* x_lo, int N, int H, int W, int C,
                                       const float* rois, int R, int pooled_h, int pooled_w,
                                       float spatial_scale, float pad_ratio, void* y_hi, void* y_lo,
                                       int out_channels_total, int out_channel_offset, void* stream) {
  return mscnn_roi_align_forward_dyn(x_hi, x_lo, N, H, W, C, rois, R, pooled_h, pooled_w, spatial_scale, pad_ratio,
                                     y_hi, y_lo, out_channels_total, out_channel_offset, nullptr, stream);
}

extern "C" int mscnn_deconv2x_forward(const void* x_hi, const void* x_lo, const float* w, void* y_hi,
                                      void* y_lo, int N, int H, int W, int C, int Creal, void* stream) {
  if (!x_hi || !y_hi || !w || N <= 0 || H <= 0 || W <= 0 || C % 8 || Creal > C) return MSCNN_ERR_INVALID;
  if ((x_lo == nullptr) != (y_lo == nullptr)) return MSCNN_ERR_INVALID;
  const size_t total = (size_t)N * 2 * H * 2 * W * (C / 8);
  mscnn::note_launch();
  const size_t w_smem = (size_t)16 * C * sizeof(float);
  if (w_smem > 200 * 1024) return MSCNN_ERR_INVALID;
  if (w_smem > 48 * 1024 &&
      cudaFuncSetAttribute(deconv2x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w_smem) != cudaSuccess)
    return MSCNN_ERR_CUDA;
  deconv2x_kernel<<<grid_for(total, 256), 256, w_smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, w, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo,
      N, H, W, C, Creal);
  return launch_check("deconv2x");
}

extern "C" int mscnn_relu_planes(void* hi, void* lo, size_t count, void* stream) {
  if (!hi || count % 8) return MSCNN_ERR_INVALID;
  if (count == 0) return MSCNN_OK;
  mscnn::note_launch();
  relu_planes_kernel<<<grid_for(count / 8, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)hi,
                                                                                 (__nv_bfloat16*)lo, count / 8);
  return launch_check("relu_planes");
}

extern "C" int mscnn_relu_f32(float* x, size_t count, void* stream) {
  if (!x) return MSCNN_ERR_INVALID;
  if (count == 0) return MSCNN_OK;
  mscnn::note_launch();
  relu_f32_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(x, count);
  return launch_check("relu_f32");
}

extern "C" int mscnn_concat_planes(const void* x, void* y, size_t pixels, int C, int Ctot, int offset,
                                   void* stream) {
  if (!x || !y || C % 8 || Ctot % 8 || offset % 8 || offset + C > Ctot) return MSCNN_ERR_INVALID;
  if (pixels == 0) return MSCNN_OK;
  mscnn::note_launch();
  concat_planes_kernel<<<grid_for(pixels * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, pixels, C, Ctot, offset);
  return launch_check("concat_planes");
}
