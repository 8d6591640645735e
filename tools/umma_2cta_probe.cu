// Micro-experiment for the next round: the mechanics of a 2-CTA tcgen05 MMA (cta_group::2), in isolation.
//   D[256][256] (fp32) = A[256][64] * B[256][64]^T   (bf16, K-major, one 64-element k-block = 4 MMAs of K = 16)
// A cluster of two CTAs: CTA r holds A rows [128 r, 128 r + 128) and B rows (= output columns) [128 r, 128 r + 128) in its
// own shared memory, at the SAME offsets in both CTAs; the leader (rank 0) issues tcgen05.mma.cta_group::2 with M = 256,
// N = 256 once; each CTA's TMEM receives its 128 rows x 256 columns.  What the probe establishes:
//   * tcgen05.alloc / dealloc with cta_group::2, executed by one warp of EACH CTA;
//   * TMA loads in the peer CTA that land in the peer's shared memory but complete_tx on the LEADER's mbarrier
//     (cp.async.bulk.tensor ... .cta_group::2, barrier address from mapa);
//   * tcgen05.commit ... .multicast::cluster signalling the same barrier offset in both CTAs;
//   * the operand split: every SM reads its own A half and HALF of the weight tile.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mscnn_b200/csrc -I include tools/umma_2cta_probe.cu \
//        mscnn_b200/csrc/tmap.cu -o tools/umma_2cta_probe.bin -lcuda      (run on the GPU box under `timeout`)
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

constexpr int kM = 256, kN = 256, kK = 64;
constexpr int kTileBytes = 128 * kK * 2;  // one CTA's half of A or B: 128 rows x 128 B

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(local_addr), "r"(rank));
  return out;
}
// 2-D TMA load whose completion is counted on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + kTileBytes, sFull = sB + kTileBytes, sDone = sFull + 8, sTm = sDone + 8;
  volatile uint32_t* tm = reinterpret_cast<volatile uint32_t*>(smem_raw + (sTm - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    ptx::mbar_init(sFull, 1);
    ptx::mbar_init(sDone, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {  // one warp of EACH CTA allocates (and later frees) the pair's tensor memory
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTm), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tm;

  if (warp == 0) {
    const uint32_t leader_full = mapa(sFull, 0);
    if (ptx::elect_one()) {
      if (rank == 0) ptx::mbar_expect_tx(sFull, 4u * kTileBytes);  // both CTAs' A and B halves
      tma_load_2d_2sm(sA, &tmA, leader_full, 0, static_cast<int>(rank) * 128);
      tma_load_2d_2sm(sB, &tmB, leader_full, 0, static_cast<int>(rank) * 128);
    }
    __syncwarp();
    if (rank == 0) {
      ptx::mbar_wait(sFull, 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t a_desc = ptx::umma_desc_sw128(sA), b_desc = ptx::umma_desc_sw128(sB);
        const uint32_t idesc = ptx::umma_idesc_bf16(kM, kN);
        for (int k = 0; k < kK / 16; ++k) umma_bf16_2sm(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, k ? 1u : 0u);
        umma_commit_2sm(sDone, 0b11);  // arrives on sDone of BOTH CTAs
      }
      __syncwarp();
    }
  }
  ptx::mbar_wait(sDone, 0);
  ptx::tc_fence_after();
  uint32_t v[32];
  float* row = out + (static_cast<size_t>(rank) * 128 + warp * 32 + lane) * kN;
  for (int c = 0; c < kN / 32; ++c) {
    ptx::tmem_ld_32x32(tmem + c * 32 + (static_cast<uint32_t>(warp * 32) << 16), v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) row[c * 32 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(kM * kK), hB(kN * kK);
  std::vector<float> fA(kM * kK), fB(kN * kK);
  srand(3);
  for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 13 - 6) / 8.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, (size_t)kM * kN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xFF, (size_t)kM * kN * 4);
  CUtensorMap tmA, tmB;
  if (tmap_2d_bf16(&tmA, dA, kK, kM, 64, 128) || tmap_2d_bf16(&tmB, dB, kK, kN, 64, 128)) return 1;
  const int smem = 2 * kTileBytes + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<2, 128, smem>>>(tmA, tmB, dO);  // __cluster_dims__(2,1,1): one CTA pair
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> hO((size_t)kM * kN);
  cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  double maxerr = 0;
  for (int r = 0; r < kM; ++r)
    for (int n = 0; n < kN; ++n) {
      double ref = 0;
      for (int k = 0; k < kK; ++k) ref += (double)fA[r * kK + k] * fB[n * kK + k];
      const double err = fabs(ref - hO[(size_t)r * kN + n]);
      if (!(err <= 1e-3)) ++bad;
      if (err > maxerr) maxerr = err;
    }
  printf("2-CTA MMA M=256 N=256 K=64: %s (bad %d / %d, max err %.4f)\n", bad ? "MISMATCH" : "ok", bad, kM * kN, maxerr);
  // which quadrants are wrong tells which half of the operand split was misread
  for (int qr = 0; qr < 2; ++qr)
    for (int qc = 0; qc < 2; ++qc) {
      int b = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 128; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += (double)fA[(qr * 128 + r) * kK + k] * fB[(qc * 128 + n) * kK + k];
          if (!(fabs(ref - hO[(size_t)(qr * 128 + r) * kN + qc * 128 + n]) <= 1e-3)) ++b;
        }
      printf("  rows %3d.. cols %3d..: %d bad\n", qr * 128, qc * 128, b);
    }
  return bad ? 3 : 0;
}
