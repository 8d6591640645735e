// SyncedMemory: lazily allocated host/device mirror with the reference's head state machine
// (/root/reference/include/caffe/syncedmem.hpp:45-83, src/caffe/syncedmem.cpp:25-139).
// Host memory is pinned (cudaMallocHost) as the reference does in GPU mode (syncedmem.hpp:15-26).
#pragma once
#include <cstdlib>

#include "caffe/common.hpp"

namespace caffe {

class CAFFE_API SyncedMemory {
 public:
  SyncedMemory() : cpu_ptr_(NULL), gpu_ptr_(NULL), size_(0), head_(UNINITIALIZED), own_cpu_data_(false),
                   own_gpu_data_(false) {}
  explicit SyncedMemory(size_t size) : cpu_ptr_(NULL), gpu_ptr_(NULL), size_(size), head_(UNINITIALIZED),
                                       own_cpu_data_(false), own_gpu_data_(false) {}
  ~SyncedMemory();
  const void* cpu_data();
  void set_cpu_data(void* data);
  const void* gpu_data();
  void set_gpu_data(void* data);
  void* mutable_cpu_data();
  void* mutable_gpu_data();
  enum SyncedHead { UNINITIALIZED, HEAD_AT_CPU, HEAD_AT_GPU, SYNCED };
  SyncedHead head() { return head_; }
  size_t size() { return size_; }

 private:
  void to_cpu();
  void to_gpu();
  void* cpu_ptr_;
  void* gpu_ptr_;
  size_t size_;
  SyncedHead head_;
  bool own_cpu_data_;
  bool own_gpu_data_;
  DISABLE_COPY_AND_ASSIGN(SyncedMemory);
};

}  // namespace caffe
