// Image-parallel MS-CNN inference on the GPUs of one box from ONE C++ process: one host thread per GPU, each with its
// own Caffe context (the mirror keeps the reference's thread-local context, /root/reference/src/caffe/common.cpp:13-22),
// its own net replica and stream, and one rank of the library's NCCL communicator (mscnn_comm_init_all).  The only
// exchange is ONE all-gather of the final detections per step, issued on the communicator's stream so that it
// overlaps the next step's trunk (include/mscnn_b200.h "Multi-GPU exchange").
//
//   g++ -std=c++17 -O2 -I include -I /usr/local/cuda/include examples/multi_gpu_driver.cpp \
//       -L mscnn_b200 -lmscnn_b200 -L /usr/local/cuda/lib64 -lcudart -lpthread -Wl,-rpath,$PWD/mscnn_b200 -o multi_gpu_driver
//   ./multi_gpu_driver deploy.prototxt --gpus 8 --steps 20 --warmup 5 [--exchange peer|nccl] [--no-gather] [--verify]
//
// Prints one line per rank (min / median / max step time on the device) and the whole-job images/s
// (max over ranks of the timed region).  --verify: every rank checks that the gathered buffer holds, for every
// other rank, exactly the packed detections that rank produced.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mscnn_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    const int rc_ = (x);                                                           \
    if (rc_ != 0) {                                                                \
      std::fprintf(stderr, "%s:%d: %s -> %d\n", __FILE__, __LINE__, #x, rc_);      \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

static void fill(float* p, long count, unsigned salt, float scale) {
  for (long e = 0; e < count; ++e) {
    const unsigned h = (static_cast<unsigned>(e) * 2654435761u + salt * 40503u) >> 8;
    p[e] = (static_cast<float>(h & 0xFFFFu) / 65536.0f - 0.5f) * scale;
  }
}

struct Barrier {  // C++17 has no std::barrier
  std::mutex m;
  std::condition_variable cv;
  int n, count = 0, gen = 0;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int g = gen;
    if (++count == n) {
      count = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

struct RankResult {
  std::vector<float> step_ms;
  float region_ms = 0.f;
  int proposals = 0;
  std::vector<float> own_payload;
  bool verified = true;
};

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s deploy.prototxt [--gpus N] [--steps K] [--warmup W] [--exchange peer|nccl] [--no-gather] [--verify]\n", argv[0]);
    return 2;
  }
  const std::string proto = argv[1];
  int gpus = 1, steps = 10, warmup = 3;
  bool gather = true, verify = false, peer = true;
  for (int i = 2; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--steps") && i + 1 < argc) steps = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--no-gather")) gather = false;
    else if (!std::strcmp(argv[i], "--exchange") && i + 1 < argc) peer = std::strcmp(argv[++i], "nccl") != 0;
    else if (!std::strcmp(argv[i], "--verify")) verify = true;
  }
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (gpus > ndev) {
    std::fprintf(stderr, "%d GPUs requested, %d visible\n", gpus, ndev);
    return 2;
  }
  std::vector<void*> comms(gpus, nullptr), xchgs(gpus, nullptr);
  if (gather && !peer) {
    CK(mscnn_comm_init_all(comms.data(), gpus, nullptr));
    std::printf("nccl %d, %d rank(s) in one process\n", mscnn_comm_nccl_version(), gpus);
  } else if (gather) {
    std::printf("peer-memory exchange, %d rank(s) in one process\n", gpus);
  }

  Barrier bar(gpus);
  std::vector<RankResult> res(gpus);
  std::vector<float*> payload_dev(gpus, nullptr);
  size_t per = 0;
  int batch = 0;
  mscnn_detect_cfg cfg;
  std::memset(&cfg, 0, sizeof(cfg));

  auto worker = [&](int rank) {
    CK(mscnn_set_device(rank));
    cudaStream_t stream;
    cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    CK(mscnn_set_stream(stream));
    void* net = mscnn_net_create(proto.c_str(), 1);
    if (!net) std::exit(1);
    // identical parameters on every rank (a replica), distinct images per rank (its shard of the global batch)
    const int L = mscnn_net_num_layers(net);
    for (int i = 0; i < L; ++i) {
      const char* name = mscnn_net_layer_name(net, i);
      const int np = mscnn_net_num_params(net, name);
      for (int j = 0; j < np; ++j) {
        int shp[4] = {1, 1, 1, 1};
        const int axes = mscnn_net_param_shape(net, name, j, shp);
        long count = 1;
        for (int a = 0; a < axes; ++a) count *= shp[a];
        const long fan = count / shp[0];
        // narrow heads (LFCN_*, cls_pred, bbox_pred) small so that scores straddle the thresholds
        const bool narrow = shp[0] <= 32;
        const float scale = j == 0 ? (narrow ? 0.35f : 3.4641f) / std::sqrt(static_cast<float>(fan > 0 ? fan : 1)) : 0.0f;
        std::vector<float> w(count);
        fill(w.data(), count, static_cast<unsigned>(i * 8 + j), scale);
        CK(mscnn_net_set_param(net, name, j, w.data(), count));
      }
    }
    const char* in_name = mscnn_net_input_name(net, 0);
    int ishp[4];
    mscnn_net_blob_shape(net, in_name, ishp);
    const long in_count = (long)ishp[0] * ishp[1] * ishp[2] * ishp[3];
    float* host_in = nullptr;
    cudaMallocHost(&host_in, in_count * sizeof(float));
    fill(host_in, in_count, 9999u + 131u * rank, 200.0f);
    if (rank == 0) {
      batch = ishp[0];
      cfg.num_cls = 5; cfg.cls_id = 2;
      for (int k = 0; k < 4; ++k) cfg.bbox_mean[k] = 0.f;
      cfg.bbox_std[0] = cfg.bbox_std[1] = 0.1f; cfg.bbox_std[2] = cfg.bbox_std[3] = 0.2f;
      cfg.proposal_thr = -10.f; cfg.nms_overlap = 0.5f;
      cfg.ratio_h = cfg.ratio_w = 1.f;
      cfg.org_h = (float)ishp[2]; cfg.org_w = (float)ishp[3];
      cfg.max_rois_per_image = 2000;
      per = mscnn_detect_payload_floats(batch, cfg.max_rois_per_image);
    }
    bar.wait();
    if (gather && peer) {
      CK(mscnn_xchg_create(&xchgs[rank], gpus, rank, per, 16));
      bar.wait();
      if (rank == 0) CK(mscnn_xchg_connect_local(xchgs.data(), gpus));
      bar.wait();
    }
    cudaMalloc(&payload_dev[rank], per * gpus * sizeof(float));
    cudaMemset(payload_dev[rank], 0, per * gpus * sizeof(float));
    float* dets = nullptr;
    int* cnt = nullptr;
    cudaMalloc(&dets, (size_t)batch * cfg.max_rois_per_image * 5 * sizeof(float));
    cudaMalloc(&cnt, batch * sizeof(int));

    auto step = [&]() {
      CK(mscnn_net_set_blob(net, in_name, host_in, in_count));   // pinned host -> device on the net's stream
      CK(mscnn_net_forward(net, 0, -1));
      if (gather && peer) CK(mscnn_net_detect_push(net, &cfg, xchgs[rank]));
      else if (gather) CK(mscnn_net_detect_gather(net, &cfg, comms[rank], payload_dev[rank]));
      else CK(mscnn_net_detect(net, &cfg, dets, cnt));
    };
    auto settle = [&]() {   // everything this rank sent and everything it should receive has landed
      if (gather && peer) CK(mscnn_xchg_wait(xchgs[rank], stream));
      cudaStreamSynchronize(stream);
      if (gather && !peer) CK(mscnn_comm_synchronize(comms[rank]));
    };
    for (int s = 0; s < warmup; ++s) step();
    settle();
    bar.wait();
    std::vector<cudaEvent_t> ev(steps + 1);
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], stream);
    for (int s = 0; s < steps; ++s) {
      step();
      cudaEventRecord(ev[s + 1], stream);
    }
    settle();
    RankResult& r = res[rank];
    for (int s = 0; s < steps; ++s) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[s], ev[s + 1]);
      r.step_ms.push_back(ms);
    }
    cudaEventElapsedTime(&r.region_ms, ev[0], ev[steps]);
    r.proposals = mscnn_net_num_proposals(net, -1);
    if (gather && verify) {
      // this rank's own packed detections, computed once more without the exchange ...
      std::vector<float> all(per * gpus);
      const float* gathered = peer ? mscnn_xchg_buffer(xchgs[rank]) : payload_dev[rank];
      cudaMemcpy(all.data(), gathered, all.size() * sizeof(float), cudaMemcpyDeviceToHost);
      r.own_payload.assign(all.begin() + per * rank, all.begin() + per * (rank + 1));
      bar.wait();
      // ... must be what every other rank received in slot `rank` (header + rows; the tail behind the rows is unused)
      for (int o = 0; o < gpus; ++o) {
        const int* head = reinterpret_cast<const int*>(res[o].own_payload.data());
        const size_t used = ((2 + batch + 3) & ~3) + (size_t)head[1] * 5;
        if (std::memcmp(all.data() + per * o, res[o].own_payload.data(), used * sizeof(float)) != 0) r.verified = false;
      }
    }
    bar.wait();
    if (xchgs[rank]) mscnn_xchg_destroy(xchgs[rank]);
    mscnn_net_destroy(net);
    cudaFree(dets);
    cudaFree(cnt);
    cudaFreeHost(host_in);
  };

  std::vector<std::thread> th;
  for (int r = 0; r < gpus; ++r) th.emplace_back(worker, r);
  for (auto& t : th) t.join();

  float worst = 0.f;
  bool ok = true;
  for (int r = 0; r < gpus; ++r) {
    std::vector<float> v = res[r].step_ms;
    std::sort(v.begin(), v.end());
    std::printf("rank %d: step ms min %.3f median %.3f max %.3f, region %.3f ms, %d proposals%s\n", r, v.front(),
                v[v.size() / 2], v.back(), res[r].region_ms, res[r].proposals,
                verify && gather ? (res[r].verified ? ", gathered payload verified" : ", GATHER MISMATCH") : "");
    worst = std::max(worst, res[r].region_ms);
    ok = ok && res[r].verified;
  }
  std::printf("images_per_s %.2f (batch %d x %d GPUs x %d steps / %.3f ms, max over ranks; gather %s)\n",
              1e3 * batch * gpus * steps / worst, batch, gpus, steps, worst, gather ? "on" : "off");
  for (void* c : comms) mscnn_comm_destroy(c);
  std::printf("exchange: %s\n", !gather ? "off" : peer ? "peer-memory push (no collective kernel)" : "ncclAllGather");
  return ok ? 0 : 1;
}
