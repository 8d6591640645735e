// Process-wide switches of the library, read from the environment ONCE (first use) instead of with getenv() on every
// launch.  mscnn_config_reload() re-reads them and drops every cached launch plan: it exists for the tests, which hold
// each operand-delivery variant of the convolution kernel to the others bit for bit.
#pragma once

namespace mscnn {

struct Config {
  bool no_fat, no_wide, no_rowshare, no_vpool, no_ring256, no_2cta, no_bf16_rings;
  bool no_head_taps;
  bool verbose_conv, c3_swap;
  bool no_2cta_pool;  // MSCNN_NO_2CTA_POOL: keep the pooled BLOCK_N = 256 layers (conv3_3) on single CTAs
  int mt;          // MSCNN_MT, 0 = per-layer default
  int conv1_mode;  // MSCNN_CONV1: 0 = tensor-core kernel (default), 1 = pair, 2 = direct, 3 = patch
  unsigned epoch;  // bumped by every reload: cached plans carry the epoch they were built under
};

const Config& config();

}  // namespace mscnn
