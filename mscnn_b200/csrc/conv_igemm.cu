// Implicit-GEMM convolution / inner product on the sm_100a tensor cores.
//
// Replaces, for the MS-CNN forward path, the reference's
//   ConvolutionLayer::Forward_{cpu,gpu}   (src/caffe/layers/conv_layer.cpp:25-40,
//                                          base_conv_layer.cpp:257-280: im2col + sgemm + bias gemm)
//   InnerProductLayer::Forward_{cpu,gpu}  (src/caffe/layers/inner_product_layer.cpp:84-97)
//   ReLULayer::Forward_*                  (src/caffe/layers/relu_layer.cpp:9-19, fused here)
//
// Design (B200-first, not a translation of im2col+sgemm):
//   * activations live in HBM as NHWC bf16 ("planes"); an fp32-faithful tensor is the pair
//     (hi, lo) with x ~= hi + lo, |lo| <= ulp_bf16(hi)/2;
//   * GEMM view: M = output pixels, N = output channels, K = taps x input channels.
//     An M tile is a TMA box {64 ch, bw, bh, bn} with bw*bh*bn <= 128 pixels; the tap
//     (dy,dx) operand is the same box shifted by (dx-pad, dy-pad): TMA zero-fills out-of-
//     bounds pixels, which *is* the convolution's zero padding.  No im2col buffer exists.
//   * tcgen05.mma (M=128, N=BLOCK_N, K=16, bf16 -> fp32) accumulates in TMEM; two TMEM
//     accumulators are double buffered so the epilogue of tile i overlaps the MMAs of i+1.
//   * fp32-faithful mode = three bf16 GEMM terms accumulated in the same TMEM tile:
//     hi*hi + hi*lo + lo*hi (the dropped lo*lo term is < 2^-16 relative).
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2-5 =
//     epilogue (TMEM -> regs -> bias/ReLU -> bf16 split -> swizzled smem -> TMA store, or
//     direct fp32 NCHW stores for the narrow proposal / prediction heads).
//   * persistent CTAs (one per SM), static round-robin tile schedule, n-tile fastest so
//     CTAs working on the same pixels share the activation tile through L2.
//   * operand delivery is what bounds these layers (TMA fills and UMMA operand reads share the 128 B/clk
//     shared-memory port; the N = 256 layers also run against the board's power cap), so the fp32-faithful
//     path moves every operand tile ONCE per (tap, channel chunk):
//       - "fat" stages / "wide-B" (BLOCK_N <= 128): A_hi, A_lo, B_hi, B_lo in one stage, A_hi x [B_hi | B_lo] as one
//         MMA of N = 2 BLOCK_N into two accumulators that the epilogue adds, plus A_lo x B_hi;
//       - two-ring engine: activation pairs and weight tiles in separate rings; "row-share" (128 x 1 pixel boxes,
//         three horizontal taps): one activation tile of 130 pixels per (dy, chunk), the dx taps are UMMA
//         descriptors that start dx rows further (tools/umma_shift_probe.cu); BLOCK_N = 256: single-tile weight
//         slots, B_hi -> lo*hi + hi*hi, B_lo -> hi*lo;
//       - "vpool": tiles of two image rows x 128 pixels whose 2x2 max pooling happens in registers (vertical: the
//         two accumulators, horizontal: lane ^ 1), bit-identical to conv -> store -> pool_kernel.
//       - CTA pairs (PAIR instantiation, BLOCK_N = 256): clusters of two CTAs, tcgen05 cta_group::2 MMAs of M = 256
//         issued by the leader; every SM loads and reads only half of each weight tile.
//     All paths accumulate K in the order (dy, channel chunk, dx).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "config.h"
#include "mscnn_b200.h"
#include "ptx_sm100.cuh"
#include "tmap.h"
#include "launch_count.h"

namespace mscnn {

struct IgemmParams {
  int num_terms;   // 1: bf16; 3: split-bf16 (hi*hi, hi*lo, lo*hi)
  int taps_h, taps_w, pad_h, pad_w;
  int cin_chunks;  // Cin_pad / 64
  int tiles_w, tiles_h, tiles_n, n_tiles;
  int box_w, box_h, box_n;
  int relu;
  int out_mode;    // MSCNN_OUT_NHWC_BF16, MSCNN_OUT_NCHW_F32 or MSCNN_OUT_NHWC_F32
  int has_lo_out;
  int pool;           // fused 2x2 / stride-2 MAX pooling of the output tile (box_w, box_h even)
  int store_full;     // also store the un-pooled tile (0 when only the pooling layer reads it)
  int stages, epi_bufs;
  int fat;            // split mode with all four operand tiles (A_hi, A_lo, B_hi, B_lo) in one stage
  int wide;           // fat mode, two MMAs per K step: A_hi x [B_hi | B_lo] (N = 2 BLOCK_N, two accumulators
                      // side by side) and A_lo x B_hi; the epilogue adds the two accumulators
  int rowshare;       // wide mode, 3 horizontal taps, 128x1 pixel boxes: ONE activation tile of (128 + 2) pixels per
                      // (dy, channel chunk) serves the three dx taps through descriptor row shifts; activation and
                      // weight tiles travel in separate rings (sa_slots x a_slot bytes, sb_slots x 2 weight tiles)
  int sa_slots, sb_slots, a_slot;
  int a_taps;         // two-ring engine: 3 = one activation slot serves the three dx taps (row shifts), 1 = one slot per tap
  int b_split;        // two-ring engine, BLOCK_N = 256: a weight slot holds ONE tile; B_hi and B_lo of a tap follow each other
  int b_slot;         // bytes per weight slot
  int pair;           // two-ring engine on CTA PAIRS (cluster of 2, tcgen05 cta_group::2): the pair computes two adjacent M
                      // tiles of one n tile with MMAs of M = 256; each CTA loads its own activation tile and HALF of
                      // every weight tile, the leader (cluster rank 0) issues the MMAs for both (tools/umma_2cta_probe.cu)
  int acc_sets;       // TMEM accumulator sets: 2 (epilogue of tile i overlaps the MMAs of tile i + 1) or, when two wide
                      // sub-tiles of BLOCK_N = 128 already fill the 512 columns, 1 (vpool mode only)
  int vpool;          // row-share mode over image-row PAIRS (mt = 2: sub-tile j = row 2 th + j) with the 2x2 MAX pooling
                      // done in registers by the epilogue (vertical max of the two accumulators, horizontal max by
                      // lane shuffle); only the pooled tile (64 pixels) is staged and stored
  int mt;             // M sub-tiles (128 pixels each) per CTA tile: one weight tile feeds mt activation tiles
  int tmem_cols;      // 2 * mt * BLOCK_N rounded to a power of two >= 32
  const float* bias;  // [Cout_pad]
  const int* dyn_n;   // optional device count: process min(*dyn_n, out_n) images only (mscnn_conv_desc.dyn_n)
  float* out_f32;     // NCHW fp32 (out_mode 1)
  int out_n, out_c, out_h, out_w;
  int out_ld;         // row length of the pixel-major fp32 output (MSCNN_OUT_NHWC_F32)
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 B = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kEpiBarId = 1;
constexpr int kMaxMt = 4;

// PAIR: the CTA-pair (cluster of 2, tcgen05 cta_group::2) build of the two-ring engine.  It is a separate instantiation:
// a kernel that contains cluster / cta_group::2 instructions can only be launched as a cluster.
template <int BLOCK_N, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA_hi,
                  const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi,
                  const __grid_constant__ CUtensorMap tmB_lo,
                  const __grid_constant__ CUtensorMap tmO_hi,
                  const __grid_constant__ CUtensorMap tmO_lo,
                  const __grid_constant__ CUtensorMap tmP_hi,
                  const __grid_constant__ CUtensorMap tmP_lo, const IgemmParams p) {
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr uint32_t kIdesc = ptx::umma_idesc_bf16(kBlockM, BLOCK_N);
  constexpr uint32_t kIdescWide = ptx::umma_idesc_bf16(kBlockM, BLOCK_N <= 128 ? 2 * BLOCK_N : BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int S = p.stages;  // row-share mode: stages = sa_slots + sb_slots, barriers [0, sa_slots) = A ring, rest = B ring
  const int MT = p.mt;
  const uint32_t a_sub = p.fat ? 2u * kABytes : kABytes;     // one sub-tile: [A_hi] or [A_hi][A_lo]
  const uint32_t a_stride = a_sub * MT;                      // one stage: MT sub-tiles
  const uint32_t b_stride = p.fat ? 2u * kBBytes : kBBytes;  //            [B_hi] or [B_hi][B_lo]
  const uint32_t sA = smem_base;
  const uint32_t sB = p.rowshare ? sA + p.sa_slots * p.a_slot : sA + S * a_stride;
  const uint32_t sEpi = p.rowshare ? sB + p.sb_slots * p.b_slot
                                   : sB + S * b_stride;  // 1024-aligned: kABytes, kBBytes, a_slot are multiples of 1024
  constexpr int kPoolBytes = kABytes / 4;  // pooled tile: 32 rows x 128 B
  constexpr int kVpPlane = 64 * 128;  // vpool: pooled tile of 64 pixels x 64 channels
  const int epi_buf_bytes = p.vpool ? kVpPlane * (p.has_lo_out ? 2 : 1)
                                    : (kABytes + (p.pool ? kPoolBytes : 0)) * (p.has_lo_out ? 2 : 1);
  const uint32_t sMisc = sEpi + p.epi_bufs * epi_buf_bytes;
  uint8_t* misc_gen = smem_gen + (sMisc - smem_base);
  float* bias_s = reinterpret_cast<float*>(misc_gen);  // BLOCK_N floats
  const uint32_t sBar = sMisc + BLOCK_N * 4;
  // barrier layout: full[S], empty[S], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return sBar + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return sBar + 8u * (2 * S + 2 + a); };
  const uint32_t sTmemPtr = sBar + 8u * (2 * S + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(misc_gen + BLOCK_N * 4 + 8 * (2 * S + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;  // position in the CTA pair
  const bool leader = (rank == 0u);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA_hi);
    ptx::prefetch_tmap(&tmB_hi);
    if (p.num_terms > 1) {
      ptx::prefetch_tmap(&tmA_lo);
      ptx::prefetch_tmap(&tmB_lo);
    }
    if (p.out_mode == MSCNN_OUT_NHWC_BF16) {
      ptx::prefetch_tmap(&tmO_hi);
      if (p.has_lo_out) ptx::prefetch_tmap(&tmO_lo);
      if (p.pool) {
        ptx::prefetch_tmap(&tmP_hi);
        if (p.has_lo_out) ptx::prefetch_tmap(&tmP_lo);
      }
    }
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), kEpiThreads * (PAIR ? 2 : 1));  // pair: both CTAs' epilogues release the leader
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      ptx::tmem_alloc_2sm(sTmemPtr, static_cast<uint32_t>(p.tmem_cols));
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(sTmemPtr, static_cast<uint32_t>(p.tmem_cols));
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // data-dependent batch size: every role derives the same tile count from the device-side count
  int tiles_n = p.tiles_n;
  if (p.dyn_n != nullptr) {
    int n_eff = *p.dyn_n;
    n_eff = n_eff < 0 ? 0 : (n_eff > p.out_n ? p.out_n : n_eff);
    tiles_n = (n_eff + p.box_n - 1) / p.box_n;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  // the host guarantees m_tiles % MT == 0 (MT = 1 with dyn_n); in vpool mode the tile grid is already one of row pairs
  const int total_tiles = PAIR ? ((m_tiles + 1) / 2) * p.n_tiles
                          : p.vpool ? m_tiles * p.n_tiles : (m_tiles / MT) * p.n_tiles;
  // pair mode: `tile` counts PAIR tiles, a cluster strides over them; this CTA's M tile is 2 * (tile / n_tiles) + rank
  // (an odd tile count leaves a phantom M tile: its loads are zero-filled and its stores clipped by TMA)
  const int tile_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_kb = (p.fat ? 1 : p.num_terms) * p.taps_h * p.taps_w * p.cin_chunks;
  const uint32_t a_box_bytes = static_cast<uint32_t>(p.box_w * p.box_h * p.box_n) * kBlockK * 2;
  const uint32_t stage_tx = (p.fat ? 2u : 1u) * (static_cast<uint32_t>(MT) * a_box_bytes + kBBytes);

  // CTA tile -> n tile and first M sub-tile; M sub-tile -> pixel-box origin
  auto m_coords = [&](int m, int& tw, int& th, int& tn) {
    tw = m % p.tiles_w;
    m /= p.tiles_w;
    th = m % p.tiles_h;
    tn = m / p.tiles_h;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loops (warp-uniform control flow); one elected lane issues.
    if (p.rowshare) {
      // Two-ring engine: activation pairs (A_hi, A_lo) and weight tiles travel in separate rings, each operand tile of
      // a (tap, channel chunk) is fetched ONCE for the three products hi*hi, hi*lo, lo*hi.  a_taps == 3 (128x1 boxes,
      // three horizontal taps): the activation slot holds box_w + 2 pixels and serves the dx taps by row shifts.
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int halo = p.a_taps == 3 ? 2 : 0;
      const bool single = (p.num_terms == 1);  // plain bf16 path: no lo planes anywhere
      const uint32_t a_tx = (single ? 1u : 2u) * static_cast<uint32_t>((p.box_w + halo) * p.box_h * p.box_n * (p.vpool ? 2 : 1)) * kBlockK * 2;
      const uint32_t b_tx = (p.b_split || single) ? kBBytes : 2u * kBBytes;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int n_tile = tile % p.n_tiles;
        int tw, th, tn;
        m_coords(PAIR ? 2 * (tile / p.n_tiles) + static_cast<int>(rank) : tile / p.n_tiles, tw, th, tn);
        const int wa = tw * p.box_w - p.pad_w, ha = th * (p.vpool ? 2 : p.box_h) - p.pad_h, na = tn * p.box_n;
        for (int dy = 0; dy < p.taps_h; ++dy) {
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            for (int dx = 0; dx < p.taps_w; ++dx) {
              if (p.a_taps == 1 || dx == 0) {
                ptx::mbar_wait(empty_bar(sa), pa ^ 1u);
                if (ptx::elect_one()) {
                  const uint32_t a0 = sA + sa * p.a_slot;
                  const int wx = wa + (p.a_taps == 1 ? dx : 0);
                  if (PAIR) {
                    // both CTAs' tiles are counted on the LEADER's barrier, which alone expects the bytes of both
                    const uint32_t lb = ptx::mapa(full_bar(sa), 0);
                    if (leader) ptx::mbar_expect_tx(full_bar(sa), 2u * a_tx);
                    ptx::tma_load_4d_2sm(a0, &tmA_hi, lb, cc * kBlockK, wx, ha + dy, na);
                    ptx::tma_load_4d_2sm(a0 + p.a_slot / 2, &tmA_lo, lb, cc * kBlockK, wx, ha + dy, na);
                  } else {
                    ptx::mbar_expect_tx(full_bar(sa), a_tx);
                    ptx::tma_load_4d(a0, &tmA_hi, full_bar(sa), cc * kBlockK, wx, ha + dy, na);
                    if (!single) ptx::tma_load_4d(a0 + p.a_slot / 2, &tmA_lo, full_bar(sa), cc * kBlockK, wx, ha + dy, na);
                  }
                }
                __syncwarp();
                if (++sa == p.sa_slots) { sa = 0; pa ^= 1u; }
              }
              const int kcol = ((dy * p.taps_w + dx) * p.cin_chunks + cc) * kBlockK;
              for (int half = 0; half < (p.b_split ? 2 : 1); ++half) {
                const int bi = p.sa_slots + sb;
                ptx::mbar_wait(empty_bar(bi), pb ^ 1u);
                if (ptx::elect_one()) {
                  const uint32_t b0 = sB + sb * p.b_slot;
                  if (PAIR) {
                    // this CTA's half of the weight tile: output columns [128 rank, 128 rank + 128) of the n tile
                    if (leader) ptx::mbar_expect_tx(full_bar(bi), kBBytes);
                    ptx::tma_load_2d_2sm(b0, half == 0 ? &tmB_hi : &tmB_lo, ptx::mapa(full_bar(bi), 0), kcol,
                                         n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
                  } else if (p.b_split) {
                    ptx::mbar_expect_tx(full_bar(bi), b_tx);
                    ptx::tma_load_2d(b0, half == 0 ? &tmB_hi : &tmB_lo, full_bar(bi), kcol, n_tile * BLOCK_N);
                  } else {
                    ptx::mbar_expect_tx(full_bar(bi), b_tx);
                    ptx::tma_load_2d(b0, &tmB_hi, full_bar(bi), kcol, n_tile * BLOCK_N);
                    if (!single) ptx::tma_load_2d(b0 + kBBytes, &tmB_lo, full_bar(bi), kcol, n_tile * BLOCK_N);
                  }
                }
                __syncwarp();
                if (++sb == p.sb_slots) { sb = 0; pb ^= 1u; }
              }
            }
          }
        }
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile_first; tile < total_tiles && !p.rowshare; tile += tile_step) {
      const int n_tile = tile % p.n_tiles, m0 = (tile / p.n_tiles) * MT;
      int w0[kMaxMt], h0[kMaxMt], n0[kMaxMt];
#pragma unroll
      for (int j = 0; j < kMaxMt; ++j) {
        int tw, th, tn;
        m_coords(m0 + (j < MT ? j : 0), tw, th, tn);
        w0[j] = tw * p.box_w - p.pad_w;
        h0[j] = th * p.box_h - p.pad_h;
        n0[j] = tn * p.box_n;
      }
      const int terms = p.fat ? 1 : p.num_terms;
      for (int term = 0; term < terms; ++term) {
        const CUtensorMap* mapA = (term == 2) ? &tmA_lo : &tmA_hi;
        const CUtensorMap* mapB = (term == 1) ? &tmB_lo : &tmB_hi;
        // K order: dy, channel chunk, dx -- the order of the row-share mode, so that every path of this kernel
        // accumulates a given output in the same sequence (fused / unfused variants stay bit-identical)
        for (int dy = 0; dy < p.taps_h; ++dy) {
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            for (int dx = 0; dx < p.taps_w; ++dx) {
              const int tap = dy * p.taps_w + dx;
              ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
              if (ptx::elect_one()) {
                ptx::mbar_expect_tx(full_bar(stage), stage_tx);
                const uint32_t a0 = sA + stage * a_stride, b0 = sB + stage * b_stride;
                const int kcol = (tap * p.cin_chunks + cc) * kBlockK;
                ptx::tma_load_2d(b0, mapB, full_bar(stage), kcol, n_tile * BLOCK_N);
                if (p.fat) ptx::tma_load_2d(b0 + kBBytes, &tmB_lo, full_bar(stage), kcol, n_tile * BLOCK_N);
#pragma unroll
                for (int j = 0; j < kMaxMt; ++j) {
                  if (j < MT) {
                    ptx::tma_load_4d(a0 + j * a_sub, mapA, full_bar(stage), cc * kBlockK, w0[j] + dx, h0[j] + dy,
                                     n0[j]);
                    if (p.fat)
                      ptx::tma_load_4d(a0 + j * a_sub + kABytes, &tmA_lo, full_bar(stage), cc * kBlockK,
                                       w0[j] + dx, h0[j] + dy, n0[j]);
                  }
                }
              }
              __syncwarp();
              if (++stage == S) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.rowshare && (leader || !PAIR)) {  // pair mode: the leader issues the MMAs of both CTAs
      constexpr uint32_t kIdescPair = ptx::umma_idesc_bf16(2 * kBlockM, BLOCK_N);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        if (PAIR) ptx::mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);  // arrivals come from both CTAs
        else ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t acc_w = static_cast<uint32_t>(p.wide ? 2 * BLOCK_N : BLOCK_N);
        const uint32_t d_base = tmem_base + static_cast<uint32_t>(acc * MT) * acc_w;
        const uint32_t sub_rows = static_cast<uint32_t>(p.box_w + 2) * 128u;  // vpool: sub-tile j starts j image rows further
        uint32_t first = 0u;  // the first MMA of the tile overwrites the accumulator
        uint32_t a_hi = 0, a_lo = 0;
        for (int dy = 0; dy < p.taps_h; ++dy) {
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            for (int dx = 0; dx < p.taps_w; ++dx) {
              if (p.a_taps == 1 || dx == 0) {
                ptx::mbar_wait(full_bar(sa), pa);
                a_hi = sA + sa * p.a_slot;
                a_lo = a_hi + p.a_slot / 2;
              }
              // tap dx = the activation rows shifted by dx pixels: a K-major SWIZZLE_128B operand may start at any
              // 128-byte row of a TMA-written tile (tools/umma_shift_probe.cu)
              const uint32_t shift = p.a_taps == 3 ? static_cast<uint32_t>(dx) * 128u : 0u;
              const bool a_done = (p.a_taps == 1) || (dx == p.taps_w - 1);
              const bool last = (dy == p.taps_h - 1) && (cc == p.cin_chunks - 1) && (dx == p.taps_w - 1);
              const int sa_now = sa;
              if (p.b_split) {
                // weight slot 1: B_hi -> lo*hi and hi*hi; weight slot 2: B_lo -> hi*lo
                for (int half = 0; half < 2; ++half) {
                  const int bi = p.sa_slots + sb;
                  ptx::mbar_wait(full_bar(bi), pb);
                  ptx::tc_fence_after();
                  if (ptx::elect_one()) {
                    const uint64_t b_desc = ptx::umma_desc_sw128(sB + sb * p.b_slot);
                    const uint64_t ah = ptx::umma_desc_sw128(a_hi + shift), al = ptx::umma_desc_sw128(a_lo + shift);
                    if (PAIR) {
                      // M = 256 across the pair: A = each CTA's own 128 rows, B = the two half tiles side by side
                      if (half == 0) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                          ptx::umma_bf16_2sm(d_base, al + 2u * k, b_desc + 2u * k, kIdescPair, (first | k) != 0 ? 1u : 0u);
                          ptx::umma_bf16_2sm(d_base, ah + 2u * k, b_desc + 2u * k, kIdescPair, 1u);
                        }
                      } else {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                          ptx::umma_bf16_2sm(d_base, ah + 2u * k, b_desc + 2u * k, kIdescPair, 1u);
                      }
                      ptx::umma_commit_2sm(empty_bar(bi));  // frees the slot in BOTH CTAs
                      if (half == 1 && a_done) ptx::umma_commit_2sm(empty_bar(sa_now));
                      if (half == 1 && last) ptx::umma_commit_2sm(tfull_bar(acc));
                    } else {
                    if (half == 0) {
#pragma unroll
                      for (int k = 0; k < kBlockK / 16; ++k) {
                        ptx::umma_bf16(d_base, al + 2u * k, b_desc + 2u * k, kIdesc, (first | k) != 0 ? 1u : 0u);
                        ptx::umma_bf16(d_base, ah + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                      }
                    } else {
#pragma unroll
                      for (int k = 0; k < kBlockK / 16; ++k) ptx::umma_bf16(d_base, ah + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                    }
                    ptx::umma_commit(empty_bar(bi));
                    if (half == 1 && a_done) ptx::umma_commit(empty_bar(sa_now));
                    if (half == 1 && last) ptx::umma_commit(tfull_bar(acc));
                    }
                  }
                  __syncwarp();
                  first = 1u;
                  if (++sb == p.sb_slots) { sb = 0; pb ^= 1u; }
                }
              } else {
                const int bi = p.sa_slots + sb;
                ptx::mbar_wait(full_bar(bi), pb);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                  const uint64_t b_desc = ptx::umma_desc_sw128(sB + sb * p.b_slot);
                  const uint64_t b_lo = ptx::umma_desc_sw128(sB + sb * p.b_slot + kBBytes);
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    if (j < MT) {
                      const uint32_t d_tmem = d_base + static_cast<uint32_t>(j) * acc_w;
                      const uint64_t ah = ptx::umma_desc_sw128(a_hi + j * sub_rows + shift);
                      const uint64_t al = ptx::umma_desc_sw128(a_lo + j * sub_rows + shift);
                      if (p.num_terms == 1) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                          ptx::umma_bf16(d_tmem, ah + 2u * k, b_desc + 2u * k, kIdesc, (first | k) != 0 ? 1u : 0u);
                      } else if (p.wide) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                          ptx::umma_bf16(d_tmem, ah + 2u * k, b_desc + 2u * k, kIdescWide, (first | k) != 0 ? 1u : 0u);
                          ptx::umma_bf16(d_tmem, al + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                        }
                      } else {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                          ptx::umma_bf16(d_tmem, ah + 2u * k, b_lo + 2u * k, kIdesc, (first | k) != 0 ? 1u : 0u);
                          ptx::umma_bf16(d_tmem, al + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                          ptx::umma_bf16(d_tmem, ah + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                        }
                      }
                    }
                  }
                  ptx::umma_commit(empty_bar(bi));
                  if (a_done) ptx::umma_commit(empty_bar(sa_now));
                  if (last) ptx::umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                first = 1u;
                if (++sb == p.sb_slots) { sb = 0; pb ^= 1u; }
              }
              if (a_done) {
                if (++sa == p.sa_slots) { sa = 0; pa ^= 1u; }
              }
            }
          }
        }
        if (++acc == p.acc_sets) acc = 0;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    for (int tile = tile_first; tile < total_tiles && !p.rowshare; tile += tile_step) {
      ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t acc_w = static_cast<uint32_t>(p.wide ? 2 * BLOCK_N : BLOCK_N);  // TMEM columns per sub-tile
      const uint32_t d_base = tmem_base + static_cast<uint32_t>(acc * MT) * acc_w;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t b_desc = ptx::umma_desc_sw128(sB + stage * b_stride);
          const uint64_t b_lo = ptx::umma_desc_sw128(sB + stage * b_stride + kBBytes);
#pragma unroll
          for (int j = 0; j < kMaxMt; ++j) {
            if (j < MT) {
              const uint32_t d_tmem = d_base + static_cast<uint32_t>(j) * acc_w;
              const uint64_t a_desc = ptx::umma_desc_sw128(sA + stage * a_stride + j * a_sub);
              if (p.wide) {
                // B_hi and B_lo are adjacent in the stage = one K-major tile of 2 BLOCK_N rows: A_hi meets both
                // in ONE MMA (columns [0, BN) = hi*hi, [BN, 2 BN) = hi*lo), then A_lo * B_hi adds to the first
                // half.  Per K step the tensor core reads 2 A tiles + 3 B tiles from shared memory instead of
                // 3 + 3: narrow-N layers are bound by exactly that read bandwidth (128 B/clk/SM).
                const uint64_t a_lo = ptx::umma_desc_sw128(sA + stage * a_stride + j * a_sub + kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  ptx::umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdescWide, (kb | k) != 0 ? 1u : 0u);
                  ptx::umma_bf16(d_tmem, a_lo + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                }
              } else if (p.fat) {
                // three products per K step from one stage: hi*lo, lo*hi, hi*hi
                const uint64_t a_lo = ptx::umma_desc_sw128(sA + stage * a_stride + j * a_sub + kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  ptx::umma_bf16(d_tmem, a_desc + 2u * k, b_lo + 2u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
                  ptx::umma_bf16(d_tmem, a_lo + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                  ptx::umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, 1u);
                }
              } else {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  // +32 B per UMMA_K step inside the 128 B swizzle atom -> +2 in the address field
                  ptx::umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
                }
              }
            }
          }
          ptx::umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs finish
          if (kb == num_kb - 1) ptx::umma_commit(tfull_bar(acc));  // accumulators complete -> epilogue
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int et = threadIdx.x - 64;         // 0..127
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;     // accumulator row == tile pixel
    const bool issuer = (et == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int ebuf = 0;
    const int hw_box = p.box_w * p.box_h;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int n_tile = tile % p.n_tiles;
      const int m0 = PAIR ? 2 * (tile / p.n_tiles) + static_cast<int>(rank) : (tile / p.n_tiles) * MT;
      const int n_base = n_tile * BLOCK_N;
      // bias slice for this n tile (visible after the first named barrier below)
      for (int j = et; j < BLOCK_N; j += kEpiThreads) bias_s[j] = p.bias[n_base + j];

      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      if constexpr (BLOCK_N >= 64 && BLOCK_N <= 128) {
        if (p.vpool) {
          // Fused PoolingLayer (MAX 2x2 / 2, pooling_layer.cpp:128-187) in registers.  Accumulator 0 holds image row
          // 2 th, accumulator 1 row 2 th + 1, thread `row` owns pixel x0 + row of both.  max commutes with the
          // (monotonic) bias add, ReLU and (hi, lo) rounding, so pooling the raw fp32 sums gives the values
          // pool_kernel computes from the stored planes.  Vertical max: the two accumulators; horizontal max: lane ^ 1.
          // Even lanes then convert channels [0, 32) of the pooled pixel row / 2, odd lanes channels [32, 64).
          int tw, th, tn;
          m_coords(tile / p.n_tiles, tw, th, tn);
          const uint32_t acc_w = static_cast<uint32_t>(p.wide ? 2 * BLOCK_N : BLOCK_N);
          const uint32_t t0 = tmem_base + static_cast<uint32_t>(acc * 2) * acc_w + (static_cast<uint32_t>(quarter * 32) << 16);
          const uint32_t t1 = t0 + acc_w;
          const int odd = lane & 1, prow = row >> 1;
#pragma unroll 1
          for (int chunk = 0; chunk < BLOCK_N / 64; ++chunk) {
            float mine[32];  // this lane's 32 channels of the pooled pixel
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {  // channels [32 hc, 32 hc + 32) of the chunk
              uint32_t a[32], b[32];
              ptx::tmem_ld_32x32(t0 + chunk * 64 + hc * 32, a);
              ptx::tmem_ld_32x32(t1 + chunk * 64 + hc * 32, b);
              if (p.wide) {
                uint32_t ua[32], ub[32];
                ptx::tmem_ld_32x32(t0 + BLOCK_N + chunk * 64 + hc * 32, ua);
                ptx::tmem_ld_32x32(t1 + BLOCK_N + chunk * 64 + hc * 32, ub);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  a[i] = __float_as_uint(__uint_as_float(a[i]) + __uint_as_float(ua[i]));
                  b[i] = __float_as_uint(__uint_as_float(b[i]) + __uint_as_float(ub[i]));
                }
              }
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float m = fmaxf(__uint_as_float(a[i]), __uint_as_float(b[i]));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                if (hc == odd) mine[i] = m;
              }
            }
            if (issuer) {
              if (p.epi_bufs == 1) ptx::tma_store_wait_read<0>();
              else ptx::tma_store_wait_read<1>();
            }
            ptx::named_bar_sync(kEpiBarId, kEpiThreads);  // staging buffer free; also publishes bias_s
            const uint32_t pbuf = sEpi + ebuf * epi_buf_bytes;
            const uint32_t prow_hi = pbuf + prow * 128, prow_lo = prow_hi + kVpPlane;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float f0 = mine[j * 8 + 2 * q] + bias_s[chunk * 64 + odd * 32 + j * 8 + 2 * q];
                float f1 = mine[j * 8 + 2 * q + 1] + bias_s[chunk * 64 + odd * 32 + j * 8 + 2 * q + 1];
                if (p.relu) {
                  f0 = fmaxf(f0, 0.f);
                  f1 = fmaxf(f1, 0.f);
                }
                {
                  // what pool_kernel sees is the STORED value hi + lo of the winning pixel; it then splits that sum
                  // again.  Reproduce both roundings so that the pooled planes are bit-identical to the unfused path
                  // (the two splits differ when lo is exactly half an ulp of hi).
                  const __nv_bfloat162 s2 = __floats2bfloat162_rn(f0, f1);
                  const uint32_t sh = *reinterpret_cast<const uint32_t*>(&s2);
                  const float g0 = __uint_as_float(sh << 16), g1 = __uint_as_float(sh & 0xFFFF0000u);
                  const __nv_bfloat162 t2 = __floats2bfloat162_rn(f0 - g0, f1 - g1);
                  const uint32_t sl = *reinterpret_cast<const uint32_t*>(&t2);
                  // (plain bf16 path: only hi is stored, so the stored value is g)
                  f0 = p.has_lo_out ? g0 + __uint_as_float(sl << 16) : g0;
                  f1 = p.has_lo_out ? g1 + __uint_as_float(sl & 0xFFFF0000u) : g1;
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(f0 - __uint_as_float(hi[q] << 16),
                                                                f1 - __uint_as_float(hi[q] & 0xFFFF0000u));
                lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              const uint32_t off = static_cast<uint32_t>(((odd * 4 + j) ^ (prow & 7)) << 4);  // 128B swizzle
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow_hi + off), "r"(hi[0]), "r"(hi[1]),
                           "r"(hi[2]), "r"(hi[3]) : "memory");
              if (p.has_lo_out)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow_lo + off), "r"(lo[0]), "r"(lo[1]),
                             "r"(lo[2]), "r"(lo[3]) : "memory");
            }
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(kEpiBarId, kEpiThreads);
            if (issuer) {
              const int c0 = n_base + chunk * 64;
              ptx::tma_store_4d(&tmP_hi, pbuf, c0, tw * (p.box_w >> 1), th, tn);
              if (p.has_lo_out) ptx::tma_store_4d(&tmP_lo, pbuf + kVpPlane, c0, tw * (p.box_w >> 1), th, tn);
              ptx::tma_store_commit();
            }
            if (++ebuf == p.epi_bufs) ebuf = 0;
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(tempty_bar(acc));
          if (++acc == p.acc_sets) acc = 0;
          if (acc == 0) acc_phase ^= 1u;
          continue;
        }
      }
      for (int sub = 0; sub < MT; ++sub) {
        int tw, th, tn;
        m_coords(m0 + sub, tw, th, tn);
        const uint32_t t_row = tmem_base + static_cast<uint32_t>((acc * MT + sub) * (p.wide ? 2 * BLOCK_N : BLOCK_N)) +
                               (static_cast<uint32_t>(quarter * 32) << 16);

        if (p.out_mode == MSCNN_OUT_NHWC_BF16) {
          if constexpr (BLOCK_N >= 64) {
#pragma unroll 1
            for (int chunk = 0; chunk < BLOCK_N / 64; ++chunk) {
              uint32_t v[64];
              ptx::tmem_ld_32x32(t_row + chunk * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
              ptx::tmem_ld_32x32(t_row + chunk * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
              uint32_t u[64];
              if (p.wide) {
                ptx::tmem_ld_32x32(t_row + BLOCK_N + chunk * 64, *reinterpret_cast<uint32_t(*)[32]>(&u[0]));
                ptx::tmem_ld_32x32(t_row + BLOCK_N + chunk * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&u[32]));
              }
              // the staging buffer we are about to overwrite must have been read by its TMA store
              if (issuer) {
                if (p.epi_bufs == 1) ptx::tma_store_wait_read<0>();
                else ptx::tma_store_wait_read<1>();
              }
              ptx::named_bar_sync(kEpiBarId, kEpiThreads);  // also publishes bias_s
              ptx::tmem_ld_wait();
              if (p.wide) {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
              }
              const uint32_t buf = sEpi + ebuf * epi_buf_bytes;
              const uint32_t row_hi = buf + row * 128;
              const uint32_t row_lo = row_hi + kABytes;
#pragma unroll
              for (int j = 0; j < 8; ++j) {  // 8 x 16 B = 64 channels
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float f0 = __uint_as_float(v[j * 8 + 2 * q]) + bias_s[chunk * 64 + j * 8 + 2 * q];
                  float f1 =
                      __uint_as_float(v[j * 8 + 2 * q + 1]) + bias_s[chunk * 64 + j * 8 + 2 * q + 1];
                  if (p.relu) {
                    f0 = fmaxf(f0, 0.f);
                    f1 = fmaxf(f1, 0.f);
                  }
                  // packed conversions (cvt.rn.bf16x2.f32): same rounding as scalar ones, half the issue slots
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                  hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
                  const __nv_bfloat162 l2 = __floats2bfloat162_rn(f0 - __uint_as_float(hi[q] << 16),
                                                                  f1 - __uint_as_float(hi[q] & 0xFFFF0000u));
                  lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                const uint32_t off = static_cast<uint32_t>((j ^ (row & 7)) << 4);  // 128B swizzle
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + off),
                             "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                             : "memory");
                if (p.has_lo_out)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_lo + off),
                               "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3])
                               : "memory");
              }
              const uint32_t pbuf = buf + kABytes * (p.has_lo_out ? 2 : 1);  // pooled staging [hi][lo]
              if (p.pool) {
                // Fused PoolingLayer (MAX, 2x2, stride 2; pooling_layer.cpp:128-187) on the staged tile:
                // exactly what pool_kernel would compute from the stored planes (value = hi + lo).
                ptx::named_bar_sync(kEpiBarId, kEpiThreads);
                const int pw_box = p.box_w >> 1, phw_box = pw_box * (p.box_h >> 1);
                const int prows = phw_box * p.box_n;  // <= 32
                for (int item = et; item < prows * 8; item += kEpiThreads) {
                  const int pr = item >> 3, j = item & 7;
                  const int dn = pr / phw_box, rem = pr - dn * phw_box;
                  const int ph = rem / pw_box, pwi = rem - ph * pw_box;
                  const int r00 = dn * hw_box + (2 * ph) * p.box_w + 2 * pwi;
                  float best[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) best[e] = -3.402823466e+38f;
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const int r = r00 + (c >> 1) * p.box_w + (c & 1);
                    const uint32_t a = buf + r * 128 + static_cast<uint32_t>((j ^ (r & 7)) << 4);
                    uint32_t xh[4], xl[4] = {0u, 0u, 0u, 0u};
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(xh[0]), "=r"(xh[1]), "=r"(xh[2]), "=r"(xh[3]) : "r"(a));
                    if (p.has_lo_out)
                      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                   : "=r"(xl[0]), "=r"(xl[1]), "=r"(xl[2]), "=r"(xl[3]) : "r"(a + kABytes));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      const float v0 = __uint_as_float(xh[q] << 16) + __uint_as_float(xl[q] << 16);
                      const float v1 = __uint_as_float(xh[q] & 0xFFFF0000u) + __uint_as_float(xl[q] & 0xFFFF0000u);
                      best[2 * q] = (v0 > best[2 * q]) ? v0 : best[2 * q];
                      best[2 * q + 1] = (v1 > best[2 * q + 1]) ? v1 : best[2 * q + 1];
                    }
                  }
                  uint32_t oh[4], ol[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(best[2 * q], best[2 * q + 1]);
                    oh[q] = *reinterpret_cast<const uint32_t*>(&h2);
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(best[2 * q] - __uint_as_float(oh[q] << 16),
                                                                    best[2 * q + 1] - __uint_as_float(oh[q] & 0xFFFF0000u));
                    ol[q] = *reinterpret_cast<const uint32_t*>(&l2);
                  }
                  const uint32_t pa = pbuf + pr * 128 + static_cast<uint32_t>((j ^ (pr & 7)) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa), "r"(oh[0]), "r"(oh[1]),
                               "r"(oh[2]), "r"(oh[3]) : "memory");
                  if (p.has_lo_out)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + kPoolBytes), "r"(ol[0]),
                                 "r"(ol[1]), "r"(ol[2]), "r"(ol[3]) : "memory");
                }
              }
              ptx::fence_proxy_async_smem();
              ptx::named_bar_sync(kEpiBarId, kEpiThreads);
              if (issuer) {
                const int c0 = n_base + chunk * 64;
                if (!p.pool || p.store_full) {
                  ptx::tma_store_4d(&tmO_hi, buf, c0, tw * p.box_w, th * p.box_h, tn * p.box_n);
                  if (p.has_lo_out)
                    ptx::tma_store_4d(&tmO_lo, buf + kABytes, c0, tw * p.box_w, th * p.box_h,
                                      tn * p.box_n);
                }
                if (p.pool) {
                  ptx::tma_store_4d(&tmP_hi, pbuf, c0, tw * (p.box_w >> 1), th * (p.box_h >> 1), tn * p.box_n);
                  if (p.has_lo_out)
                    ptx::tma_store_4d(&tmP_lo, pbuf + kPoolBytes, c0, tw * (p.box_w >> 1), th * (p.box_h >> 1),
                                      tn * p.box_n);
                }
                ptx::tma_store_commit();
              }
              if (++ebuf == p.epi_bufs) ebuf = 0;
            }
          }
        } else {
          // fp32 outputs, direct stores
          if (sub == 0) ptx::named_bar_sync(kEpiBarId, kEpiThreads);  // bias_s visible
          const int dn = row / hw_box, rem = row - dn * hw_box;
          const int dh = rem / p.box_w, dw = rem - dh * p.box_w;
          const int n = tn * p.box_n + dn, h = th * p.box_h + dh, w = tw * p.box_w + dw;
          const bool ok = (dn < p.box_n) && n < p.out_n && h < p.out_h && w < p.out_w;
          const size_t plane = static_cast<size_t>(p.out_h) * p.out_w;
          // NCHW: consecutive lanes are consecutive pixels of one channel
          float* obase = p.out_f32 + (static_cast<size_t>(n) * p.out_c) * plane +
                         static_cast<size_t>(h) * p.out_w + w;
          // pixel-major fp32 rows [pixel][ld]: each thread owns one 128-byte segment per chunk
          float* rbase = p.out_f32 + ((static_cast<size_t>(n) * p.out_h + h) * p.out_w + w) * p.out_ld + n_base;
#pragma unroll 1
          for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_row + chunk * 32, v);
            if (p.wide) {
              uint32_t u[32];
              ptx::tmem_ld_32x32(t_row + BLOCK_N + chunk * 32, u);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
            }
            ptx::tmem_ld_wait();
            if (p.out_mode == MSCNN_OUT_NHWC_F32) {
              if (ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 o;
                  o.x = __uint_as_float(v[j]) + bias_s[chunk * 32 + j];
                  o.y = __uint_as_float(v[j + 1]) + bias_s[chunk * 32 + j + 1];
                  o.z = __uint_as_float(v[j + 2]) + bias_s[chunk * 32 + j + 2];
                  o.w = __uint_as_float(v[j + 3]) + bias_s[chunk * 32 + j + 3];
                  if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                  *reinterpret_cast<float4*>(rbase + chunk * 32 + j) = o;
                }
              }
            } else if (ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int ch = n_base + chunk * 32 + j;
                if (ch < p.out_c) {
                  float f = __uint_as_float(v[j]) + bias_s[chunk * 32 + j];
                  if (p.relu) f = fmaxf(f, 0.f);
                  obase[static_cast<size_t>(ch) * plane] = f;
                }
              }
            }
          }
        }
      }
      if (p.out_mode != MSCNN_OUT_NHWC_BF16)
        ptx::named_bar_sync(kEpiBarId, kEpiThreads);  // bias_s reuse hazard for the next tile
      // all TMEM reads of these accumulators are done -> hand them back to the MMA warp (pair mode: the leader's)
      ptx::tc_fence_before();
      if (PAIR && !leader) ptx::mbar_arrive_cluster(ptx::mapa(tempty_bar(acc), 0));
      else ptx::mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (issuer) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all();  // the peer may still be arriving on this CTA's barriers / reading its weights
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ------------------------------------------------------------------ host side

static int pick_block_n(int cout_pad) {
  if (cout_pad % 256 == 0) return 256;
  if (cout_pad % 128 == 0) return 128;
  if (cout_pad % 64 == 0) return 64;
  return 32;
}

// Choose the pixel box (bw, bh, bn), bw*bh*bn <= 128, that wastes the fewest MMA rows.
static void pick_box(int N, int Ho, int Wo, bool even, int* bw_o, int* bh_o, int* bn_o) {
  double best = -1.0;
  int b_w = 1, b_h = 1, b_n = 1;
  for (int bw = 1; bw <= 128 && bw <= Wo; ++bw) {
    for (int bh = 1; bh * bw <= 128 && bh <= Ho; ++bh) {
      if (even && ((bw & 1) || (bh & 1))) continue;  // 2x2 pooling windows must not straddle tiles
      int bn = 128 / (bw * bh);
      if (bn > N) bn = N;
      // rows of a K-major SW128 tile come in groups of 8: any row count works for TMA, the
      // MMA simply ignores the tail rows.
      const long tiles = (long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * ((N + bn - 1) / bn);
      const double eff = (double)N * Ho * Wo / (tiles * 128.0);
      // prefer wider boxes on ties: longer contiguous runs for TMA
      const double score = eff + 1e-6 * bw + 1e-9 * bh;
      if (score > best) {
        best = score;
        b_w = bw;
        b_h = bh;
        b_n = bn;
      }
    }
  }
  *bw_o = b_w;
  *bh_o = b_h;
  *bn_o = b_n;
}

constexpr int kSmemBudget = 227 * 1024;
constexpr int kMaxDevices = 64;

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device and function: raised once to the whole budget.
template <typename K>
static cudaError_t raise_smem_once(K kern, bool (&done)[kMaxDevices], int device) {
  if (device >= 0 && device < kMaxDevices && done[device]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  if (e == cudaSuccess && device >= 0 && device < kMaxDevices) done[device] = true;
  return e;
}

template <int BLOCK_N>
static cudaError_t launch_igemm(const CUtensorMap maps[8], const IgemmParams& p, int grid,
                                size_t smem, int device, cudaStream_t stream) {
  auto kern = conv_igemm_kernel<BLOCK_N, false>;
  static bool done[kMaxDevices];
  cudaError_t e = raise_smem_once(kern, done, device);
  if (e != cudaSuccess) return e;
  mscnn::note_launch();
  kern<<<grid, kThreads, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p);
  return cudaGetLastError();
}

// CTA-pair launch: cluster dimension (2, 1, 1) as a launch attribute (the kernel itself serves both modes).
static cudaError_t launch_igemm_pair(const CUtensorMap maps[8], const IgemmParams& p, int grid, size_t smem,
                                     int device, cudaStream_t stream) {
  auto kern = conv_igemm_kernel<256, true>;
  static bool done[kMaxDevices];
  cudaError_t e = raise_smem_once(kern, done, device);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  mscnn::note_launch();
  e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p);
  return e != cudaSuccess ? e : cudaGetLastError();
}

// A launch plan: everything mscnn_conv_forward derives from a descriptor (tile shapes, ring sizes, the eight tensor
// maps).  Plans are cached per (descriptor, device, config epoch): a steady-state forward does one hash lookup per
// convolution instead of ~15 getenv() and up to 8 cuTensorMapEncodeTiled calls (27 convolutions per step).
struct ConvPlan {
  IgemmParams p;
  CUtensorMap maps[8];
  int grid;
  int BN;
  size_t smem;
};

struct PlanKey {
  const void* ptr[11];
  int v[15];
  bool operator==(const PlanKey& o) const { return memcmp(this, &o, sizeof(PlanKey)) == 0; }
};
struct PlanKeyHash {
  size_t operator()(const PlanKey& k) const {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
    uint64_t h = 1469598103934665603ull;  // FNV-1a
    for (size_t i = 0; i < sizeof(PlanKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static std::mutex g_plan_mu;
static std::unordered_map<PlanKey, ConvPlan, PlanKeyHash> g_plans;

void conv_plan_cache_clear() {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plans.clear();
}

static int build_plan(const mscnn_conv_desc* d, ConvPlan* plan, bool encode_maps = true);

}  // namespace mscnn

using namespace mscnn;

extern "C" int mscnn_conv_forward(const mscnn_conv_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!d) return MSCNN_ERR_INVALID;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return MSCNN_ERR_CUDA;
  PlanKey key;
  memset(&key, 0, sizeof(key));
  const void* ptrs[11] = {d->x_hi, d->x_lo, d->w_hi, d->w_lo, d->bias, d->y_hi, d->y_lo, d->y_f32, d->pool_hi, d->pool_lo, d->dyn_n};
  memcpy(key.ptr, ptrs, sizeof(ptrs));
  const int vals[15] = {d->N, d->H, d->W, d->C, d->Cout, d->Cout_pad, d->KH, d->KW, d->pad_h, d->pad_w, d->relu,
                        d->out_mode, device, static_cast<int>(config().epoch), 0};
  memcpy(key.v, vals, sizeof(vals));
  ConvPlan plan;
  bool hit = false;
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
      plan = it->second;
      hit = true;
    }
  }
  if (!hit) {
    const int rc = build_plan(d, &plan);
    if (rc != MSCNN_OK) return rc;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (g_plans.size() >= 4096) g_plans.clear();  // blobs that keep moving (re-allocation): bound the table
    g_plans.emplace(key, plan);
  }
  cudaError_t e;
  const IgemmParams& p = plan.p;
  switch (plan.BN) {
    case 256: e = p.pair ? launch_igemm_pair(plan.maps, p, plan.grid, plan.smem, device, stream)
                         : launch_igemm<256>(plan.maps, p, plan.grid, plan.smem, device, stream); break;
    case 128: e = launch_igemm<128>(plan.maps, p, plan.grid, plan.smem, device, stream); break;
    case 64: e = launch_igemm<64>(plan.maps, p, plan.grid, plan.smem, device, stream); break;
    default: e = launch_igemm<32>(plan.maps, p, plan.grid, plan.smem, device, stream); break;
  }
  if (e != cudaSuccess) {
    fprintf(stderr, "mscnn_conv_forward: launch failed: %s\n", cudaGetErrorString(e));
    return MSCNN_ERR_CUDA;
  }
  return MSCNN_OK;
}

namespace mscnn {

static int build_plan(const mscnn_conv_desc* d, ConvPlan* plan, bool encode_maps) {
  const Config& cfg = config();
  if (!d->x_hi || !d->w_hi || !d->bias) return MSCNN_ERR_INVALID;
  if (d->C % 64 != 0) return MSCNN_ERR_INVALID;
  const bool split = (d->x_lo != nullptr);
  if (split && !d->w_lo) return MSCNN_ERR_INVALID;
  const int Ho = d->H + 2 * d->pad_h - d->KH + 1;
  const int Wo = d->W + 2 * d->pad_w - d->KW + 1;
  if (Ho <= 0 || Wo <= 0 || d->N <= 0) return MSCNN_ERR_INVALID;
  const int BN = pick_block_n(d->Cout_pad);
  if (d->Cout_pad % BN != 0 || d->Cout > d->Cout_pad) return MSCNN_ERR_INVALID;
  const bool pool = (d->out_mode == MSCNN_OUT_NHWC_BF16 && d->pool_hi != nullptr);
  if (d->out_mode == MSCNN_OUT_NHWC_BF16 && (BN < 64 || (!d->y_hi && !pool))) return MSCNN_ERR_INVALID;
  if (pool && ((Ho & 1) || (Wo & 1))) return MSCNN_ERR_INVALID;  // fused pooling: even output extents only
  if ((d->out_mode == MSCNN_OUT_NCHW_F32 || d->out_mode == MSCNN_OUT_NHWC_F32) && !d->y_f32) return MSCNN_ERR_INVALID;

  IgemmParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.num_terms = split ? 3 : 1;
  p.taps_h = d->KH;
  p.taps_w = d->KW;
  p.pad_h = d->pad_h;
  p.pad_w = d->pad_w;
  p.cin_chunks = d->C / 64;
  pick_box(d->N, Ho, Wo, pool, &p.box_w, &p.box_h, &p.box_n);
  // Pooled narrow-N 3x3 layers of the fp32-faithful path (conv1_2, conv2_2): row-share mode over image-row pairs
  // with the pooling done in registers (p.vpool, see the kernel).  Tiles are 128 pixels x 2 rows.
  const bool vpair = pool && !d->y_hi && BN <= 128 && d->KW == 3 && d->KH == 3 && (Wo % 128 == 0) &&
                     !cfg.no_fat && !cfg.no_rowshare && !cfg.no_vpool &&
                     (split || !cfg.no_bf16_rings);
  if (vpair) {
    p.box_w = 128;
    p.box_h = 1;
    p.box_n = 1;
  }
  p.tiles_w = (Wo + p.box_w - 1) / p.box_w;
  p.tiles_h = vpair ? Ho / 2 : (Ho + p.box_h - 1) / p.box_h;
  p.tiles_n = (d->N + p.box_n - 1) / p.box_n;
  p.n_tiles = d->Cout_pad / BN;
  p.relu = d->relu;
  p.out_mode = d->out_mode;
  p.has_lo_out = (d->out_mode == MSCNN_OUT_NHWC_BF16 && (d->y_lo != nullptr || (pool && d->pool_lo != nullptr))) ? 1 : 0;
  p.pool = pool ? 1 : 0;
  p.store_full = (d->y_hi != nullptr) ? 1 : 0;
  if (pool && p.store_full && p.has_lo_out && !d->y_lo) return MSCNN_ERR_INVALID;
  if (pool && p.has_lo_out && !d->pool_lo) return MSCNN_ERR_INVALID;
  p.bias = d->bias;
  p.dyn_n = d->dyn_n;
  p.out_f32 = d->y_f32;
  p.out_n = d->N;
  p.out_c = d->Cout;
  p.out_h = Ho;
  p.out_w = Wo;
  p.out_ld = d->Cout_pad;

  // shared memory plan
  const int b_bytes = BN * kBlockK * 2;
  // fat stages (A_hi, A_lo, B_hi, B_lo of a k-block in one stage, three MMAs per K step) for the
  // narrow-N layers of the fp32-faithful path: A_hi / B_hi are fetched once instead of twice and a
  // barrier round trip covers 3x the MMAs (conv1_2: 6.3 -> 3.9 ms, profiles/r01_probe_layers*.log).
  p.fat = (split && BN <= 128 && !cfg.no_fat) ? 1 : 0;
  p.wide = (p.fat && !cfg.no_wide) ? 1 : 0;
  // vpair with BN = 128: two wide sub-tiles are 512 TMEM columns, so a single accumulator set (the epilogue of a
  // tile is ~7 % of its MMA time there); keeps the accumulation order of the un-pooled row-share path bit for bit
  p.acc_sets = (vpair && BN == 128 && p.wide) ? 1 : 2;
  const int acc_mul = p.wide ? 2 : 1;  // accumulator columns per sub-tile = acc_mul * BN
  // M sub-tiles per CTA tile: for narrow N one weight tile should feed several activation tiles
  // (fewer hot-line weight fetches and barrier round trips per MMA).  2 * mt * BN TMEM columns <= 512.
  const int m_tiles_total = p.tiles_w * p.tiles_h * p.tiles_n;
  int mt = 1;
  if (BN <= 64) mt = 4;
  else if (BN == 128) mt = 2;
  if (cfg.mt > 0) mt = cfg.mt;
  if (d->dyn_n) mt = 1;  // the tile count is only known on the device: no grouping of M sub-tiles
  if (mt > kMaxMt) mt = kMaxMt;
  if (mt < 1) mt = 1;
  while (mt > 1 && (2 * mt * BN * acc_mul > 512 || m_tiles_total % mt != 0 || m_tiles_total / mt * p.n_tiles < mscnn_sm_count())) mt >>= 1;
  const int epi_unit = (d->out_mode == MSCNN_OUT_NHWC_BF16) ? (kABytes + (pool ? kABytes / 4 : 0)) * (p.has_lo_out ? 2 : 1) : 0;
  const int misc = BN * 4 + 8 * (2 * 8 + 4) + 16 + 1024 /*alignment slack*/;
  const int budget = kSmemBudget;
  int epi_bufs = (epi_unit == 0) ? 0 : 2;
  int stage_bytes = 0, stages = 0;
  for (;; mt >>= 1) {
    stage_bytes = (mt * kABytes + b_bytes) * (p.fat ? 2 : 1);
    epi_bufs = (epi_unit == 0) ? 0 : 2;
    stages = (budget - misc - epi_bufs * epi_unit) / stage_bytes;
    if (stages < 4 && epi_bufs == 2) {
      epi_bufs = 1;
      stages = (budget - misc - epi_bufs * epi_unit) / stage_bytes;
    }
    if (stages >= 3 || mt == 1) break;  // keep at least a 3-deep TMA pipeline
  }
  if (stages > 8) stages = 8;
  if (stages < 2) return MSCNN_ERR_INVALID;
  // Row-share mode: wide mode over 128 x 1 pixel boxes with three horizontal taps.  The activation tile of a
  // (dy, channel chunk) is loaded ONCE with a one-pixel halo on each side (130 rows) and serves the three dx taps
  // through descriptor row shifts; narrow-N layers are bound by shared-memory bandwidth (TMA fill + operand
  // reads share 128 B/clk/SM, profiles/r01h_summary.md), and this removes two of three activation fills.
  size_t smem_rs = 0;
  if (vpair) {
    const int a_plane = ((p.box_w + 2) * 2 * 128 + 1023) / 1024 * 1024;  // two image rows of 130 pixels
    const int a_slot = (split ? 2 : 1) * a_plane, b_slot = (split ? 2 : 1) * b_bytes;
    const int vp_unit = 64 * 128 * (p.has_lo_out ? 2 : 1);
    for (int eb = 2; eb >= 1 && !p.rowshare; --eb) {
      int sb = (budget - misc - eb * vp_unit - 2 * a_slot) / b_slot;
      if (sb > 6) sb = 6;
      if (sb >= 2) {
        p.rowshare = 1;
        p.a_taps = 3;
        p.b_split = 0;
        p.b_slot = b_slot;
        p.vpool = 1;
        p.sa_slots = 2;
        p.sb_slots = sb;
        p.a_slot = a_slot;
        epi_bufs = eb;
        stages = 2 + sb;
        mt = 2;
        smem_rs = (size_t)2 * a_slot + (size_t)sb * b_slot + (size_t)eb * vp_unit + misc;
      }
    }
    if (!p.rowshare) return MSCNN_ERR_INVALID;
  }
  const bool bf16_rings = !split && !cfg.no_bf16_rings;  // plain bf16: same engine, single planes
  if (!vpair && (p.wide || bf16_rings) && d->KW == 3 && p.box_w == 128 && p.box_h == 1 && p.box_n == 1 && !pool &&
      !cfg.no_rowshare) {
    const int a_plane = ((p.box_w + 2) * 128 + 1023) / 1024 * 1024;
    const int a_slot = (split ? 2 : 1) * a_plane, b_slot = (split ? 2 : 1) * b_bytes;
    for (int eb = (epi_unit == 0 ? 0 : 2); eb >= (epi_unit == 0 ? 0 : 1) && !p.rowshare; --eb) {
      const int rings = budget - misc - eb * epi_unit;
      for (int sa = 3; sa >= 2 && !p.rowshare; --sa) {
        int sb = (rings - sa * a_slot) / b_slot;
        if (sb > 5) sb = 5;
        if (sa + sb > 8) sb = 8 - sa;
        if (sb >= 3) {
          p.rowshare = 1;
          p.a_taps = 3;
          p.b_split = 0;
          p.b_slot = b_slot;
          p.sa_slots = sa;
          p.sb_slots = sb;
          p.a_slot = a_slot;
          epi_bufs = eb;
          stages = sa + sb;
          mt = 1;
          smem_rs = (size_t)sa * a_slot + (size_t)sb * b_slot + (size_t)eb * epi_unit + misc;
        }
      }
    }
  }
  // BLOCK_N = 256, fp32-faithful: the same two-ring engine with single-tile weight slots.  The term-major loop of the
  // non-fat path fetches A_hi and B_hi twice per (tap, channel chunk) (144 KB of TMA fills); here the activation pair and
  // both weight tiles are fetched once (96 KB; 76 KB with the row-share halo on 128 x 1 boxes): a third less L2->SM and
  // shared-memory fill traffic on the layers that hold 70 % of the step and run against the power cap.
  if (split && BN == 256 && !cfg.no_ring256) {
    const bool halo = (d->KW == 3 && p.box_w == 128 && p.box_h == 1 && p.box_n == 1 && !pool && !cfg.no_rowshare);
    const int a_plane = halo ? ((p.box_w + 2) * 128 + 1023) / 1024 * 1024 : kABytes;
    // CTA pairs (MSCNN_NO_2CTA=1 turns them off): half weight tiles per CTA, MMAs of M = 256 issued by the leader;
    // conv3_2 2.13 -> 2.00 ms, the step 37.3 -> 35.5 ms on the same B200 (profiles/r01n_summary.md)
    // (pooled layers too: the 2x2 pooling of the staged tile is per CTA and does not care who issued the MMAs)
    p.pair = (!cfg.no_2cta && !(pool && cfg.no_2cta_pool) && d->out_mode == MSCNN_OUT_NHWC_BF16 && mscnn_sm_count() >= 2) ? 1 : 0;
    // pooled + an odd number of M tiles (a phantom tile in the last pair): that combination has not been run on the
    // device yet, so it keeps the validated single-CTA build (batch-1 toy sizes only; every BASELINE size is even)
    if (pool && (m_tiles_total & 1)) p.pair = 0;
    const int a_slot = 2 * a_plane, b_slot = p.pair ? b_bytes / 2 : b_bytes;
    const int eb = (epi_unit == 0) ? 0 : 1;
    const int rings = budget - misc - eb * epi_unit;
    for (int sa = 3; sa >= 2 && !p.rowshare; --sa) {
      int sb = (rings - sa * a_slot) / b_slot;
      if (sb > (p.pair ? 5 : 4)) sb = p.pair ? 5 : 4;
      if (sb >= 3) {
        p.rowshare = 1;
        p.a_taps = halo ? 3 : 1;
        p.b_split = 1;
        p.b_slot = b_slot;
        p.sa_slots = sa;
        p.sb_slots = sb;
        p.a_slot = a_slot;
        epi_bufs = eb;
        stages = sa + sb;
        mt = 1;
        smem_rs = (size_t)sa * a_slot + (size_t)sb * b_slot + (size_t)eb * epi_unit + misc;
      }
    }
    if (!p.rowshare) p.pair = 0;
  }
  p.mt = mt;
  int tcols = 32;
  while (tcols < p.acc_sets * mt * BN * acc_mul) tcols <<= 1;
  p.tmem_cols = tcols;
  p.stages = stages;
  p.epi_bufs = epi_bufs;  // 0 in fp32-output mode: no staging region is carved
  const size_t smem = p.rowshare ? smem_rs : (size_t)stages * stage_bytes + (size_t)epi_bufs * epi_unit + misc;
  if (cfg.verbose_conv)
    fprintf(stderr, "conv plan: N=%d H=%d W=%d C=%d Cout_pad=%d k=%dx%d BN=%d box=%dx%dx%d mt=%d fat=%d wide=%d rowshare=%d(%d+%d) vpool=%d a_taps=%d b_split=%d pair=%d terms=%d stages=%d epi_bufs=%d smem=%zu tiles=%d\n",
            d->N, d->H, d->W, d->C, d->Cout_pad, d->KH, d->KW, BN, p.box_w, p.box_h, p.box_n, p.mt, p.fat, p.wide, p.rowshare, p.sa_slots, p.sb_slots, p.vpool, p.a_taps, p.b_split, p.pair, p.num_terms, stages,
            epi_bufs, smem, m_tiles_total / p.mt * p.n_tiles);

  plan->BN = BN;
  plan->smem = smem;
  CUtensorMap* maps = plan->maps;
  memset(maps, 0, sizeof(plan->maps));
  if (!encode_maps) {  // mscnn_conv_plan_describe: the shape decisions only, no driver call
    plan->grid = 0;
    return MSCNN_OK;
  }
  const uint32_t obox[4] = {64u, (uint32_t)p.box_w, (uint32_t)p.box_h, (uint32_t)p.box_n};
  const uint32_t abox[4] = {64u, (uint32_t)(p.box_w + (p.a_taps == 3 ? 2 : 0)), (uint32_t)(p.vpool ? 2 : p.box_h), (uint32_t)p.box_n};
  const uint64_t adim[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
  int rc = tmap_nhwc_bf16(&maps[0], d->x_hi, adim, abox);
  if (rc) return rc;
  if (split) {
    rc = tmap_nhwc_bf16(&maps[1], d->x_lo, adim, abox);
    if (rc) return rc;
  } else {
    maps[1] = maps[0];
  }
  const uint64_t ktot = (uint64_t)d->KH * d->KW * d->C;
  const uint32_t b_box_rows = (uint32_t)(p.pair ? BN / 2 : BN);
  rc = tmap_2d_bf16(&maps[2], d->w_hi, ktot, (uint64_t)d->Cout_pad, 64u, b_box_rows);
  if (rc) return rc;
  if (split) {
    rc = tmap_2d_bf16(&maps[3], d->w_lo, ktot, (uint64_t)d->Cout_pad, 64u, b_box_rows);
    if (rc) return rc;
  } else {
    maps[3] = maps[2];
  }
  if (d->out_mode == MSCNN_OUT_NHWC_BF16) {
    const uint64_t odim[4] = {(uint64_t)d->Cout_pad, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->N};
    if (d->y_hi) {
      rc = tmap_nhwc_bf16(&maps[4], d->y_hi, odim, obox);
      if (rc) return rc;
    } else {
      maps[4] = maps[0];
    }
    if (p.has_lo_out && d->y_lo) {
      rc = tmap_nhwc_bf16(&maps[5], d->y_lo, odim, obox);
      if (rc) return rc;
    } else {
      maps[5] = maps[4];
    }
  } else {
    maps[4] = maps[0];
    maps[5] = maps[0];
  }
  maps[6] = maps[4];
  maps[7] = maps[5];
  if (p.pool) {
    const uint64_t pdim[4] = {(uint64_t)d->Cout_pad, (uint64_t)(Wo / 2), (uint64_t)(Ho / 2), (uint64_t)d->N};
    const uint32_t pbox[4] = {64u, (uint32_t)(p.box_w / 2), (uint32_t)(p.vpool ? 1 : p.box_h / 2), (uint32_t)p.box_n};
    rc = tmap_nhwc_bf16(&maps[6], d->pool_hi, pdim, pbox);
    if (rc) return rc;
    if (p.has_lo_out) {
      rc = tmap_nhwc_bf16(&maps[7], d->pool_lo, pdim, pbox);
      if (rc) return rc;
    } else {
      maps[7] = maps[6];
    }
  }

  const int total_tiles = p.vpool ? p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles
                                  : p.tiles_w * p.tiles_h * p.tiles_n / p.mt * p.n_tiles;
  int grid = mscnn_sm_count();
  if (grid > total_tiles) grid = total_tiles;
  if (p.pair) {  // whole clusters of two
    const int pair_tiles = (p.tiles_w * p.tiles_h * p.tiles_n + 1) / 2 * p.n_tiles;
    grid = mscnn_sm_count() & ~1;
    if (grid > 2 * pair_tiles) grid = 2 * pair_tiles;
  }
  plan->grid = grid;
  plan->BN = BN;
  plan->smem = smem;
  return MSCNN_OK;
}

}  // namespace mscnn

// The launch plan mscnn_conv_forward would use for `d`, as text, WITHOUT touching the device (no tensor map is encoded,
// the pointers in `d` only have to be non-NULL where the call would require them): which instantiation, pixel box, ring
// layout, CTA pairs, row-share halo, register pooling.  Host-side tests pin the variant every full-size layer takes.
extern "C" int mscnn_conv_plan_describe(const mscnn_conv_desc* d, char* buf, int cap) {
  if (!d || !buf || cap <= 0) return MSCNN_ERR_INVALID;
  mscnn::ConvPlan plan;
  const int rc = mscnn::build_plan(d, &plan, false);
  if (rc != MSCNN_OK) return rc;
  const mscnn::IgemmParams& p = plan.p;
  const int n = snprintf(buf, (size_t)cap,
                         "kernel=conv_igemm_kernel<%d, %s> box=%dx%dx%d mt=%d terms=%d fat=%d wide=%d rings=%d(%d+%d) a_taps=%d "
                         "b_split=%d vpool=%d pool=%d acc_sets=%d stages=%d epi_bufs=%d smem=%zu dyn=%d",
                         plan.BN, p.pair ? "true" : "false", p.box_w, p.box_h, p.box_n, p.mt, p.num_terms, p.fat, p.wide,
                         p.rowshare, p.sa_slots, p.sb_slots, p.a_taps, p.b_split, p.vpool, p.pool, p.acc_sets, p.stages,
                         p.epi_bufs, plan.smem, p.dyn_n ? 1 : 0);
  return (n > 0 && n < cap) ? MSCNN_OK : MSCNN_ERR_INVALID;
}
