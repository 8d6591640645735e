// Environment switches, read once (config.h).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "config.h"
#include "mscnn_b200.h"

namespace mscnn {

static Config g_cfg;
static std::atomic<bool> g_loaded{false};
static std::mutex g_mu;

static void load_locked() {
  auto on = [](const char* k) { return std::getenv(k) != nullptr; };
  Config c;
  memset(&c, 0, sizeof(c));
  c.no_fat = on("MSCNN_NO_FAT");
  c.no_wide = on("MSCNN_NO_WIDE");
  c.no_rowshare = on("MSCNN_NO_ROWSHARE");
  c.no_vpool = on("MSCNN_NO_VPOOL");
  c.no_ring256 = on("MSCNN_NO_RING256");
  c.no_2cta = on("MSCNN_NO_2CTA");
  c.no_2cta_pool = on("MSCNN_NO_2CTA_POOL");
  c.no_bf16_rings = on("MSCNN_NO_BF16_RINGS");
  c.no_head_taps = on("MSCNN_NO_HEAD_TAPS");
  c.verbose_conv = on("MSCNN_VERBOSE_CONV");
  c.c3_swap = on("MSCNN_C3_SWAP");
  if (const char* e = std::getenv("MSCNN_MT")) c.mt = atoi(e);
  if (const char* e = std::getenv("MSCNN_CONV1")) {
    c.conv1_mode = !strcmp(e, "pair") ? 1 : !strcmp(e, "direct") ? 2 : 3;
  }
  c.epoch = g_cfg.epoch + 1;
  g_cfg = c;
  g_loaded.store(true, std::memory_order_release);
}

const Config& config() {
  if (!g_loaded.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_loaded.load(std::memory_order_relaxed)) load_locked();
  }
  return g_cfg;
}

void conv_plan_cache_clear();  // conv_igemm.cu

}  // namespace mscnn

extern "C" void mscnn_config_reload(void) {
  {
    std::lock_guard<std::mutex> lk(mscnn::g_mu);
    mscnn::load_locked();
  }
  mscnn::conv_plan_cache_clear();
}
