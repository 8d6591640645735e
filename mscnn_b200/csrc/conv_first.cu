// First convolution of the trunk (conv1_1: 3 -> 64 channels, 3x3, pad 1) as a direct fp32 kernel.
//
// K = 27 is far below one 64-channel tensor-core k-block (a GEMM formulation pads it to 64 and has
// to materialise 128 B of patch data per pixel and plane), and the layer is bound by its OUTPUT
// traffic (256 B per pixel in fp32-faithful mode) rather than by its 1,728 MACs per pixel.  So
// this one layer runs on the CUDA cores in exact fp32 FMA arithmetic: one thread per pixel,
// 64 accumulators in registers, the 3x3x3 input window and the 27x64 weights in shared memory
// (weights read as broadcast float4), fused bias + ReLU + bf16 split, 128-byte row stores.
// Replaces ConvolutionLayer::Forward_gpu for that layer shape (src/caffe/layers/conv_layer.cu:8-23).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mscnn_b200.h"

namespace mscnn {

constexpr int kFirstTile = 128;  // pixels (along W) per CTA

__global__ void __launch_bounds__(kFirstTile)
conv3x3_c3_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cout][3][3][3]*/,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ yh,
                  __nv_bfloat16* __restrict__ yl, int N, int H, int W, int Cout, int Cout_pad, int relu) {
  __shared__ float s_in[3][3][kFirstTile + 2];
  __shared__ __align__(16) float s_w[27][64];
  __shared__ float s_b[64];
  const int tid = threadIdx.x;
  const int w0 = blockIdx.x * kFirstTile, h = blockIdx.y;
  const int n = blockIdx.z / (Cout_pad / 64), cg = blockIdx.z % (Cout_pad / 64);
  for (int i = tid; i < 27 * 64; i += kFirstTile) {
    const int t = i / 64, co = cg * 64 + (i % 64);
    s_w[t][i % 64] = (co < Cout) ? w[(size_t)co * 27 + t] : 0.f;  // t = c*9 + dy*3 + dx: Caffe's own order
  }
  if (tid < 64) s_b[tid] = (bias && cg * 64 + tid < Cout) ? bias[cg * 64 + tid] : 0.f;
  for (int i = tid; i < 9 * (kFirstTile + 2); i += kFirstTile) {
    const int col = i % (kFirstTile + 2), r = i / (kFirstTile + 2);
    const int c = r / 3, dy = r % 3;
    const int hh = h + dy - 1, ww = w0 + col - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((size_t)(n * 3 + c) * H + hh) * W + ww];
    s_in[c][dy][col] = v;
  }
  __syncthreads();
  const int wpix = w0 + tid;
  if (wpix >= W) return;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const int c = t / 9, dy = (t % 9) / 3, dx = t % 3;
    const float xv = s_in[c][dy][tid + dx];
    const float4* wr = reinterpret_cast<const float4*>(&s_w[t][0]);
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      const float4 wv = wr[g];
      acc[4 * g + 0] = __fmaf_rn(xv, wv.x, acc[4 * g + 0]);
      acc[4 * g + 1] = __fmaf_rn(xv, wv.y, acc[4 * g + 1]);
      acc[4 * g + 2] = __fmaf_rn(xv, wv.z, acc[4 * g + 2]);
      acc[4 * g + 3] = __fmaf_rn(xv, wv.w, acc[4 * g + 3]);
    }
  }
  const size_t o = (((size_t)n * H + h) * W + wpix) * Cout_pad + cg * 64;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 a, b;
    uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
    uint32_t* bp = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f0 = acc[8 * j + 2 * q] + s_b[8 * j + 2 * q];
      float f1 = acc[8 * j + 2 * q + 1] + s_b[8 * j + 2 * q + 1];
      if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
      const __nv_bfloat16 h0 = __float2bfloat16_rn(f0), h1 = __float2bfloat16_rn(f1);
      ap[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(f0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(f1 - __bfloat162float(h1));
      bp[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(yh + o + 8 * j) = a;
    if (yl) *reinterpret_cast<uint4*>(yl + o + 8 * j) = b;
  }
}

}  // namespace mscnn

extern "C" int mscnn_conv3x3_c3_forward(const float* x, const float* w, const float* bias, void* y_hi, void* y_lo,
                                        int N, int H, int W, int Cout, int Cout_pad, int relu, void* stream) {
  if (!x || !w || !y_hi || N <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout_pad % 64 || Cout > Cout_pad)
    return MSCNN_ERR_INVALID;
  if (H > 65535 || (long)N * (Cout_pad / 64) > 65535) return MSCNN_ERR_INVALID;
  dim3 grid((W + mscnn::kFirstTile - 1) / mscnn::kFirstTile, H, N * (Cout_pad / 64));
  mscnn::conv3x3_c3_kernel<<<grid, mscnn::kFirstTile, 0, (cudaStream_t)stream>>>(
      x, w, bias, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo, N, H, W, Cout, Cout_pad, relu);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "mscnn_conv3x3_c3_forward: %s\n", cudaGetErrorString(e));
    return MSCNN_ERR_CUDA;
  }
  return MSCNN_OK;
}
