// Shim for the two uses in the reference's hot path: boost::mutex (src/caffe/layer.cpp:6-23)
// and boost::thread_specific_ptr (src/caffe/common.cpp:13-22).
#pragma once
#include <memory>
#include <mutex>
#include <unistd.h>  // real boost/thread.hpp pulls this in; common.cpp:36 relies on it for getpid()
namespace boost {
class mutex {
 public:
  void lock() { m_.lock(); }
  void unlock() { m_.unlock(); }
 private:
  std::mutex m_;
};
template <typename T>
class thread_specific_ptr {
 public:
  T* get() const { return slot().get(); }
  void reset(T* p = nullptr) { slot().reset(p); }
  T* operator->() const { return get(); }
 private:
  static std::unique_ptr<T>& slot() {
    static thread_local std::unique_ptr<T> s;
    return s;
  }
};
}  // namespace boost
