// Micro-experiment: can a K-major SWIZZLE_128B UMMA A-operand start at a row that is NOT a
// multiple of 8 (i.e. not 1024-byte aligned), as needed to reuse one TMA-loaded halo tile for the
// horizontally shifted taps of a convolution?  For each row shift s the MMA is issued with the
// descriptor start address advanced by s*128 B and (a) base_offset = 0, (b) base_offset = s & 7;
// the 128x64 result is compared with A[s:s+128] * B^T computed on the host.
// Second question (sbo_rows = 10, 12): may the stride between the 8-row groups (SBO) differ from 1024 B, so
// that the groups are the 8-pixel rows of a halo tile (8 + 2) pixels wide and ONE TMA-loaded halo tile
// (10 x 18 pixels) serves all nine taps of a 3x3 convolution over an 8 x 16 pixel output tile?
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mscnn_b200/csrc -I include tools/umma_shift_probe.cu \
//        mscnn_b200/csrc/tmap.cu -o gpurun_out/umma_shift_probe     (run on the GPU box)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ptx_sm100.cuh"
#include "tmap.h"

using namespace mscnn;

constexpr int kRowsA = 192;  // 16 groups x 10 rows + up to 22 rows of shift, <= 24 KB
constexpr int kN = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift,
             int base_offset, int sbo_rows, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 24 * 1024, sBar = sB + 8 * 1024, sTm = sBar + 16;
  volatile uint32_t* tm = reinterpret_cast<volatile uint32_t*>(smem_raw + (sTm - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(sBar, 1);
    ptx::mbar_init(sBar + 8, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(sTm, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tm;
  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(sBar, kRowsA * 128 + kN * 128);
      ptx::tma_load_2d(sA, &tmA, sBar, 0, 0);
      ptx::tma_load_2d(sB, &tmB, sBar, 0, 0);
    }
    __syncwarp();
    ptx::mbar_wait(sBar, 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      uint64_t a_desc = ptx::umma_desc_sw128(sA + shift * 128);
      a_desc |= static_cast<uint64_t>(base_offset & 7) << 49;
      // stride between 8-row groups: 1024 B in the canonical layout; (8 + 2) * 128 B when the groups are
      // the 8-pixel rows of a halo tile that is 10 pixels wide
      a_desc &= ~(static_cast<uint64_t>(0x3FFF) << 32);
      a_desc |= static_cast<uint64_t>((sbo_rows * 128) >> 4) << 32;
      const uint64_t b_desc = ptx::umma_desc_sw128(sB);
      const uint32_t idesc = ptx::umma_idesc_bf16(128, kN);
      for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, k ? 1u : 0u);
      ptx::umma_commit(sBar + 8);
    }
    __syncwarp();
  }
  ptx::mbar_wait(sBar + 8, 0);
  ptx::tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < kN / 32; ++c) {
    ptx::tmem_ld_32x32(tmem + c * 32 + (static_cast<uint32_t>(warp * 32) << 16), v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + c * 32 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__nv_bfloat16> hA(kRowsA * 64), hB(kN * 64);
  std::vector<float> fA(kRowsA * 64), fB(kN * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 13 - 6) / 8.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, 128 * kN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (tmap_2d_bf16(&tmA, dA, 64, kRowsA, 64, kRowsA) || tmap_2d_bf16(&tmB, dB, 64, kN, 64, kN)) return 1;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  std::vector<float> hO(128 * kN);
  for (int sbo_rows = 8; sbo_rows <= 12; sbo_rows += 2)
  for (int shift = 0; shift <= (sbo_rows == 8 ? 10 : 22); ++shift) {
    if (15 * sbo_rows + 7 + shift >= kRowsA) continue;
    for (int variant = 0; variant < 2; ++variant) {
      const int bo = variant ? (shift & 7) : 0;
      cudaMemset(dO, 0, 128 * kN * 4);
      probe_kernel<<<1, 128, 40 * 1024>>>(tmA, tmB, shift, bo, sbo_rows, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("sbo %d shift %d base_offset %d: CUDA error %s\n", sbo_rows, shift, bo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      double maxerr = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          const int src = (r / 8) * sbo_rows + (r % 8) + shift;
          for (int k = 0; k < 64; ++k) ref += (double)fA[src * 64 + k] * fB[n * 64 + k];
          const double err = fabs(ref - hO[r * kN + n]);
          if (err > 1e-3) ++bad;
          if (err > maxerr) maxerr = err;
        }
      printf("sbo_rows %2d shift %2d base_offset %d: %s (bad %d / %d, max err %.4f)\n", sbo_rows, shift, bo, bad ? "MISMATCH" : "ok", bad, 128 * kN, maxerr);
    }
  }
  return 0;
}
