// conv1_1 (3 -> 64 channels, 3x3, pad 1) as ONE kernel: fp32 NCHW image in, NHWC bf16 planes out.
//
// Replaces ConvolutionLayer::Forward_gpu + ReLULayer for the first layer of the trunk
// (src/caffe/layers/conv_layer.cpp:25-40, base_conv_layer.cpp:257-280; the reference builds a
// 27 x (H*W) im2col buffer per image and calls sgemm).  The layer is HBM-bound (K = 27): per pixel it
// must read 12 B and write 64 channels x 2 B (x2 for the (hi, lo) pair) = 256 B.  The previous formulation
// (im2col3x3_c3_pair_kernel + the generic implicit-GEMM kernel) wrote and re-read a 2 GB patch tensor and ran
// the generic epilogue with one staging buffer: 2.7 ms for 8 x 768 x 2560 pixels against a 0.65 ms HBM floor
// (profiles/r01h_summary.md).  Here:
//   * producer warps read the fp32 image rows themselves (coalesced along x, zero padding by predication),
//     split to bf16 (hi, lo) and write them to shared memory as PIXEL ROWS of 8 channels (3 real) = 16 B per
//     pixel.  Nothing else is staged: no im2col tensor, no padded copy of the image.
//   * the im2col is done by the UMMA shared-memory DESCRIPTOR: in the no-swizzle K-major canonical layout a core
//     matrix is 8 rows x 16 B with rows 16 B apart, the next 8-row group is SBO bytes further and the next
//     K chunk LBO bytes further.  With SBO = 128 B and LBO = 16 B over a pixel row, A[row r][K chunk j] is
//     pixel (r + j): the three horizontal taps are three overlapping views of the same 2 KB row segment.
//     K = 16 per MMA = two taps; (dx = 0,1) and (dx = 2, zero weights) -> 2 MMAs per image row dy, 6 per tile
//     and term, accumulated in TMEM.  Weights (64 x 96, 12 KB per plane) stay resident in shared memory.
//   * 8 epilogue warps (two per TMEM lane quarter) convert to (hi, lo), stage the 128 pixel x 64 channel tile
//     128B-swizzled and hand it to TMA stores, four staging buffers deep so that stores of earlier tiles drain
//     while later tiles are converted; 4 TMEM accumulators decouple the MMA warp from the epilogue.
// fp32-faithful mode accumulates hi*lo + lo*hi + hi*hi like conv_igemm.cu; bf16 mode one term.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "config.h"
#include "mscnn_b200.h"
#include "ptx_sm100.cuh"
#include "tmap.h"
#include "launch_count.h"

namespace mscnn {
namespace c3 {

constexpr int kTileW = 128;                  // output pixels per tile = MMA rows
constexpr int kCout = 64;
constexpr int kRowPx = 132;                  // staged pixels per image row: 128 + 3 halo/overreach, padded
constexpr int kRowBytes = kRowPx * 16;       // 2112
constexpr int kPlaneBytes = 3 * kRowBytes;   // three image rows (dy)
constexpr int kStageBytes = 2 * kPlaneBytes; // hi + lo, 12672 B (multiple of 128)
constexpr int kStages = 4;
constexpr int kChunks = 12;                  // K chunks of 8 elements: (dy, dx = 0..3), dx = 3 is all-zero
constexpr int kWPlaneBytes = kChunks * 8 * 128;  // [chunk][n8][8 couts][8 k] = 12 KB
constexpr int kEpiBufs = 4;
constexpr int kEpiPlane = kTileW * 128;      // 16 KB: 128 pixels x 64 channels bf16
constexpr int kAccs = 4;                     // TMEM accumulators of 64 columns
constexpr int kProducerWarps = 4, kEpiWarps = 8;
constexpr int kMmaWarp = kProducerWarps;     // warp 4
constexpr int kFirstEpiWarp = kProducerWarps + 1;
constexpr int kThreads = 32 * (kProducerWarps + 1 + kEpiWarps);  // 416
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kEpiBarId = 1;

struct Params {
  const float* x;     // [N][3][H][W] fp32
  const uint8_t* w;   // packed weights: hi plane then lo plane, kWPlaneBytes each (smem image)
  const float* bias;  // [64]
  int N, H, W, tiles_w;
  int split, relu;
  int swap;  // debug (MSCNN_C3_SWAP=1): exchange the LBO / SBO fields of the operand descriptors
};

// no-swizzle K-major descriptor: start >> 4, LBO (K-direction core-matrix stride) at [16,30),
// SBO (8-row group stride) at [32,46), version 1 at [46,48), layout type 0.
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_c3_tc_kernel(const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t sEpi = smem_base;                                   // 1024-aligned staging buffers
  const int epi_buf_bytes = kEpiPlane * (p.split ? 2 : 1);
  const uint32_t sW = sEpi + kEpiBufs * 2 * kEpiPlane;
  const uint32_t sA = sW + 2 * kWPlaneBytes;
  const uint32_t sMisc = sA + kStages * kStageBytes;
  uint8_t* gW = gen + (sW - smem_base);
  uint8_t* gA = gen + (sA - smem_base);
  float* bias_s = reinterpret_cast<float*>(gen + (sMisc - smem_base));
  const uint32_t sBar = sMisc + kCout * 4;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return sBar + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return sBar + 8u * (2 * kStages + kAccs + a); };
  const uint32_t sTmemPtr = sBar + 8u * (2 * kStages + 2 * kAccs);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(gen + (sTmemPtr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmemCols = kAccs * kCout;  // 256

  // resident weights + bias (generic-proxy writes, made visible to the tensor core's async proxy below)
  {
    const int n16 = (p.split ? 2 : 1) * kWPlaneBytes / 16;
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(gW);
    for (int i = threadIdx.x; i < n16; i += kThreads) dst[i] = src[i];
    if (threadIdx.x < kCout) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmO_hi);
    if (p.split) ptx::prefetch_tmap(&tmO_lo);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 32);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kAccs; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), kEpiThreads);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(sTmemPtr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int total_tiles = p.N * p.H * p.tiles_w;
  const size_t plane = static_cast<size_t>(p.H) * p.W;

  if (warp < kProducerWarps) {
    // ---------------------------------------------------------------- producers: image rows -> pixel rows
    // Producer warp w fills stage w with every kStages-th tile of this CTA (kProducerWarps == kStages), so four
    // tiles' global loads are in flight per SM; the MMA warp consumes the stages in tile order.
    static_assert(kProducerWarps == kStages, "one stage per producer warp");
    constexpr int kItems = 3 * 131;                 // (dy, px) pixels a tile needs
    constexpr int kIters = (kItems + 31) / 32;      // 13
    const int stage = warp;
    uint32_t phase = 0;
    for (int tile = blockIdx.x + warp * gridDim.x; tile < total_tiles; tile += kStages * gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int y = (tile / p.tiles_w) % p.H;
      const int n = tile / (p.tiles_w * p.H);
      const int x0 = tw * kTileW - 1;  // image x of staged pixel 0
      const float* img = p.x + static_cast<size_t>(n) * 3 * plane;
      // all global loads of the tile first ...
      float v[kIters][3];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int item = it * 32 + lane;
        const int dy = item / 131, px = item - dy * 131;
        const int yy = y + dy - 1, xx = x0 + px;
        const bool ok = item < kItems && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        const float* q = img + static_cast<size_t>(ok ? yy : 0) * p.W + (ok ? xx : 0);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[it][c] = ok ? __ldg(q + c * plane) : 0.f;
      }
      // ... then wait for the slot and fill it
      ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
      uint8_t* a_hi = gA + stage * kStageBytes;
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int item = it * 32 + lane;
        if (item < kItems) {
          const int dy = item / 131, px = item - dy * 131;
          uint32_t h[3], l[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(v[it][c]);
            const __nv_bfloat16 lb = __float2bfloat16_rn(v[it][c] - __bfloat162float(hb));
            h[c] = __bfloat16_as_ushort(hb);
            l[c] = __bfloat16_as_ushort(lb);
          }
          uint8_t* dst = a_hi + dy * kRowBytes + px * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(h[0] | (h[1] << 16), h[2], 0u, 0u);
          if (p.split) *reinterpret_cast<uint4*>(dst + kPlaneBytes) = make_uint4(l[0] | (l[1] << 16), l[2], 0u, 0u);
        }
      }
      ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma operand reads
      ptx::mbar_arrive(full_bar(stage));
      phase ^= 1u;
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------------------------------- MMA issuer
    constexpr uint32_t kIdesc = ptx::umma_idesc_bf16(kTileW, kCout);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      ptx::mbar_wait(full_bar(stage), phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kCout);
        const uint32_t a_hi = sA + stage * kStageBytes, a_lo = a_hi + kPlaneBytes;
        const uint32_t w_hi = sW, w_lo = sW + kWPlaneBytes;
        uint32_t accum = 0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int m = 0; m < 2; ++m) {  // K chunks (2m, 2m+1) = pixels r + 2m, r + 2m + 1
            const uint32_t aoff = dy * kRowBytes + m * 32;
            const uint32_t boff = (dy * 4 + 2 * m) * 1024;
            const uint32_t a_l = p.swap ? 128 : 16, a_s = p.swap ? 16 : 128, b_l = p.swap ? 128 : 1024, b_s = p.swap ? 1024 : 128;
            const uint64_t ah = desc_nosw(a_hi + aoff, a_l, a_s), al = desc_nosw(a_lo + aoff, a_l, a_s);
            const uint64_t bh = desc_nosw(w_hi + boff, b_l, b_s), bl = desc_nosw(w_lo + boff, b_l, b_s);
            if (p.split) {
              ptx::umma_bf16(d_tmem, ah, bl, kIdesc, accum);
              ptx::umma_bf16(d_tmem, al, bh, kIdesc, 1u);
              ptx::umma_bf16(d_tmem, ah, bh, kIdesc, 1u);
            } else {
              ptx::umma_bf16(d_tmem, ah, bh, kIdesc, accum);
            }
            accum = 1u;
          }
        }
        ptx::umma_commit(empty_bar(stage));
        ptx::umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1u;
      }
      if (++acc == kAccs) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue
    const int et = threadIdx.x - kFirstEpiWarp * 32;  // 0..255
    const int quarter = warp & 3;                      // TMEM lanes this warp may read
    const int half = (warp - kFirstEpiWarp) >> 2;      // columns [32 half, 32 half + 32)
    const int row = quarter * 32 + lane;               // pixel of the tile
    const bool issuer = (et == 0);
    int acc = 0, ebuf = 0;
    uint32_t acc_phase = 0;
    float bias_r[32];  // this thread's 32 output channels: constant over tiles
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_r[i] = bias_s[half * 32 + i];
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int y = (tile / p.tiles_w) % p.H;
      const int n = tile / (p.tiles_w * p.H);
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + static_cast<uint32_t>(acc * kCout + half * 32) + (static_cast<uint32_t>(quarter * 32) << 16), v);
      if (issuer) ptx::tma_store_wait_read<kEpiBufs - 1>();  // the buffer we are about to overwrite has been read
      ptx::named_bar_sync(kEpiBarId, kEpiThreads);
      ptx::tmem_ld_wait();
      const uint32_t buf = sEpi + ebuf * epi_buf_bytes;
      const uint32_t row_hi = buf + row * 128, row_lo = row_hi + kEpiPlane;
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // 4 x 16 B = 32 channels
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f0 = __uint_as_float(v[j * 8 + 2 * q]) + bias_r[j * 8 + 2 * q];
          float f1 = __uint_as_float(v[j * 8 + 2 * q + 1]) + bias_r[j * 8 + 2 * q + 1];
          if (p.relu) {
            f0 = fmaxf(f0, 0.f);
            f1 = fmaxf(f1, 0.f);
          }
          // packed conversions (cvt.rn.bf16x2.f32): same rounding as two scalar ones, half the F2F issue slots
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
          hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(f0 - __uint_as_float(hi[q] << 16),
                                                          f1 - __uint_as_float(hi[q] & 0xFFFF0000u));
          lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const uint32_t off = static_cast<uint32_t>(((half * 4 + j) ^ (row & 7)) << 4);  // 128B swizzle
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                     "r"(hi[3]) : "memory");
        if (p.split)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_lo + off), "r"(lo[0]), "r"(lo[1]),
                       "r"(lo[2]), "r"(lo[3]) : "memory");
      }
      // accumulator fully read -> back to the MMA warp
      ptx::tc_fence_before();
      ptx::mbar_arrive(tempty_bar(acc));
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(kEpiBarId, kEpiThreads);
      if (issuer) {
        ptx::tma_store_4d(&tmO_hi, buf, 0, tw * kTileW, y, n);
        if (p.split) ptx::tma_store_4d(&tmO_lo, buf + kEpiPlane, 0, tw * kTileW, y, n);
        ptx::tma_store_commit();
      }
      if (++ebuf == kEpiBufs) ebuf = 0;
      if (++acc == kAccs) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (issuer) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// fp32 [64][3][3][3] -> the kernel's shared-memory weight image, hi plane then lo plane:
// [chunk = dy*4 + dx][n8][8 couts][8 k = input channel], zero for dx = 3 and channels 3..7.
__global__ void pack_c3_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int split) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // element of the hi plane
  if (i >= kWPlaneBytes / 2) return;
  const int k = i & 7, r = (i >> 3) & 7, n8 = (i >> 6) & 7, chunk = i >> 9;
  const int dy = chunk >> 2, dx = chunk & 3, co = n8 * 8 + r;
  float v = 0.f;
  if (dx < 3 && k < 3) v = w[((co * 3 + k) * 3 + dy) * 3 + dx];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  out[i] = h;
  if (split) out[kWPlaneBytes / 2 + i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

}  // namespace c3
}  // namespace mscnn

using namespace mscnn;

extern "C" int mscnn_conv1_tc_packed_bytes(void) { return 2 * c3::kWPlaneBytes; }

extern "C" int mscnn_pack_conv1_tc_weights(const float* w_f32, void* packed, int split, void* stream) {
  if (!w_f32 || !packed) return MSCNN_ERR_INVALID;
  note_launch();
  c3::pack_c3_weights_kernel<<<(c3::kWPlaneBytes / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w_f32, (__nv_bfloat16*)packed, split ? 1 : 0);
  return cudaPeekAtLastError() == cudaSuccess ? MSCNN_OK : MSCNN_ERR_CUDA;
}

extern "C" int mscnn_conv1_tc_forward(const float* x, const void* packed_w, const float* bias64, void* y_hi, void* y_lo,
                                      int N, int H, int W, int relu, void* stream) {
  if (!x || !packed_w || !bias64 || !y_hi || N < 1 || H < 1 || W < 1) return MSCNN_ERR_INVALID;
  c3::Params p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.w = static_cast<const uint8_t*>(packed_w);
  p.bias = bias64;
  p.N = N;
  p.H = H;
  p.W = W;
  p.tiles_w = (W + c3::kTileW - 1) / c3::kTileW;
  p.split = y_lo ? 1 : 0;
  p.relu = relu ? 1 : 0;
  p.swap = mscnn::config().c3_swap ? 1 : 0;
  CUtensorMap maps[2];
  memset(maps, 0, sizeof(maps));
  const uint64_t odim[4] = {64u, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint32_t obox[4] = {64u, (uint32_t)(W < c3::kTileW ? W : c3::kTileW), 1u, 1u};
  int rc = tmap_nhwc_bf16(&maps[0], y_hi, odim, obox);
  if (rc) return rc;
  if (y_lo) {
    rc = tmap_nhwc_bf16(&maps[1], y_lo, odim, obox);
    if (rc) return rc;
  } else {
    maps[1] = maps[0];
  }
  const size_t smem = (size_t)c3::kEpiBufs * 2 * c3::kEpiPlane + 2 * c3::kWPlaneBytes + c3::kStages * c3::kStageBytes +
                      c3::kCout * 4 + 8 * (2 * c3::kStages + 2 * c3::kAccs) + 16 + 1024;
  cudaError_t e = cudaFuncSetAttribute(c3::conv_c3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return MSCNN_ERR_CUDA;
  const long total = (long)N * H * p.tiles_w;
  int grid = mscnn_sm_count();
  if (grid > total) grid = (int)total;
  note_launch();
  c3::conv_c3_tc_kernel<<<grid, c3::kThreads, smem, (cudaStream_t)stream>>>(maps[0], maps[1], p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "mscnn_conv1_tc_forward: launch failed: %s\n", cudaGetErrorString(e));
    return MSCNN_ERR_CUDA;
  }
  return MSCNN_OK;
}
