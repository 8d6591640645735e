// First convolution of the trunk (conv1_1: 3 -> 64 channels, 3x3, pad 1) as a direct fp32 kernel.
//
// K = 27 is far below one 64-channel tensor-core k-block (a GEMM formulation pads it to 64 and has
// to materialise 128 B of patch data per pixel and plane), and the layer is bound by its OUTPUT
// traffic (256 B per pixel in fp32-faithful mode) rather than by its 1,728 MACs per pixel.  So
// this one layer runs on the CUDA cores in exact fp32 FMA arithmetic: one thread per pixel, the
// 27-value input window in registers (staged through shared memory), weights as immediate
// constant-bank operands, fused bias + ReLU + bf16 split, 128-byte row stores.
// Replaces ConvolutionLayer::Forward_gpu for that layer shape (src/caffe/layers/conv_layer.cu:8-23).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mscnn_b200.h"
#include "launch_count.h"

namespace mscnn {

constexpr int kFirstTile = 128;  // pixels (along W) per CTA
constexpr int kFirstRows = 4;    // output rows per CTA

// Weights / bias of the 64-output group being computed live in constant memory: with the tap and
// output loops fully unrolled every FFMA takes its weight as an immediate constant-bank operand,
// so the inner loop has no load instructions at all (a shared-memory broadcast of the weights
// would cost 4 smem cycles per 4 FMAs and bound the kernel at 1/4 of the FMA rate; measured).
__constant__ float c_first_w[64 * 27];  // [co][c*9 + dy*3 + dx] = Caffe's blob order
__constant__ float c_first_b[64];

__global__ void __launch_bounds__(kFirstTile)
conv3x3_c3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl,
                  int N, int H, int W, int Cout_pad, int cg, int relu) {
  __shared__ float s_in[3][kFirstRows + 2][kFirstTile + 2];
  const int tid = threadIdx.x;
  const int w0 = blockIdx.x * kFirstTile, h0 = blockIdx.y * kFirstRows, n = blockIdx.z;
  for (int i = tid; i < 3 * (kFirstRows + 2) * (kFirstTile + 2); i += kFirstTile) {
    const int col = i % (kFirstTile + 2), r = i / (kFirstTile + 2);
    const int c = r / (kFirstRows + 2), dy = r % (kFirstRows + 2);
    const int hh = h0 + dy - 1, ww = w0 + col - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((size_t)(n * 3 + c) * H + hh) * W + ww];
    s_in[c][dy][col] = v;
  }
  __syncthreads();
  const int wpix = w0 + tid;
  if (wpix >= W) return;
#pragma unroll 1
  for (int r = 0; r < kFirstRows; ++r) {
    const int h = h0 + r;
    if (h >= H) break;
    float xin[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) xin[t] = s_in[t / 9][r + (t % 9) / 3][tid + t % 3];
    const size_t o = (((size_t)n * H + h) * W + wpix) * Cout_pad + cg * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 8 output channels at a time -> one 16-byte store per plane
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
      for (int t = 0; t < 27; ++t) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = __fmaf_rn(xin[t], c_first_w[(8 * j + q) * 27 + t], acc[q]);
      }
      uint4 a, b;
      uint32_t* ap = reinterpret_cast<uint32_t*>(&a);
      uint32_t* bp = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f0 = acc[2 * q] + c_first_b[8 * j + 2 * q];
        float f1 = acc[2 * q + 1] + c_first_b[8 * j + 2 * q + 1];
        if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
        const __nv_bfloat16 h0b = __float2bfloat16_rn(f0), h1b = __float2bfloat16_rn(f1);
        ap[q] = (uint32_t)__bfloat16_as_ushort(h0b) | ((uint32_t)__bfloat16_as_ushort(h1b) << 16);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(f0 - __bfloat162float(h0b));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(f1 - __bfloat162float(h1b));
        bp[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint4*>(yh + o + 8 * j) = a;
      if (yl) *reinterpret_cast<uint4*>(yl + o + 8 * j) = b;
    }
  }
}

}  // namespace mscnn

extern "C" int mscnn_conv3x3_c3_forward(const float* x, const float* w, const float* bias, void* y_hi, void* y_lo,
                                        int N, int H, int W, int Cout, int Cout_pad, int relu, void* stream) {
  if (!x || !w || !y_hi || N <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout_pad % 64 || Cout > Cout_pad)
    return MSCNN_ERR_INVALID;
  if (N > 65535 || (H + mscnn::kFirstRows - 1) / mscnn::kFirstRows > 65535) return MSCNN_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  static const float zeros[64 * 27] = {0};
  dim3 grid((W + mscnn::kFirstTile - 1) / mscnn::kFirstTile, (H + mscnn::kFirstRows - 1) / mscnn::kFirstRows, N);
  for (int cg = 0; cg < Cout_pad / 64; ++cg) {
    const int cnt = (Cout - cg * 64 < 64) ? (Cout - cg * 64 > 0 ? Cout - cg * 64 : 0) : 64;
    // stream-ordered refresh of the constant bank for this group of 64 outputs (zero-filled tail)
    if (cnt < 64) {
      if (cudaMemcpyToSymbolAsync(mscnn::c_first_w, zeros, sizeof(float) * 64 * 27, 0, cudaMemcpyHostToDevice, st) !=
              cudaSuccess ||
          cudaMemcpyToSymbolAsync(mscnn::c_first_b, zeros, sizeof(float) * 64, 0, cudaMemcpyHostToDevice, st) !=
              cudaSuccess)
        return MSCNN_ERR_CUDA;
    }
    if (cnt > 0) {
      if (cudaMemcpyToSymbolAsync(mscnn::c_first_w, w + (size_t)cg * 64 * 27, sizeof(float) * cnt * 27, 0,
                                  cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return MSCNN_ERR_CUDA;
      if (bias) {
        if (cudaMemcpyToSymbolAsync(mscnn::c_first_b, bias + cg * 64, sizeof(float) * cnt, 0,
                                    cudaMemcpyDeviceToDevice, st) != cudaSuccess)
          return MSCNN_ERR_CUDA;
      } else if (cudaMemcpyToSymbolAsync(mscnn::c_first_b, zeros, sizeof(float) * 64, 0, cudaMemcpyHostToDevice,
                                         st) != cudaSuccess) {
        return MSCNN_ERR_CUDA;
      }
    }
    mscnn::note_launch();
    mscnn::conv3x3_c3_kernel<<<grid, mscnn::kFirstTile, 0, st>>>(x, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo, N,
                                                                 H, W, Cout_pad, cg, relu);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "mscnn_conv3x3_c3_forward: %s\n", cudaGetErrorString(e));
    return MSCNN_ERR_CUDA;
  }
  return MSCNN_OK;
}
