// glog-style CHECK / LOG macros with Caffe's abort-on-FATAL error convention
// (/root/reference/include/caffe/util/device_alternate.hpp:48-53 CUDA_CHECK -> CHECK_EQ -> abort).
// glog itself is not available in the image.  INFO lines are printed only when MSCNN_VERBOSE
// is set in the environment.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace caffe {
namespace logging {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
class Message {
 public:
  Message(const char* file, int line, int sev) : sev_(sev) {
    static const bool verbose = std::getenv("MSCNN_VERBOSE") != nullptr;
    on_ = sev >= WARNING || verbose;
    if (on_) ss_ << "IWEF"[sev] << " " << file << ":" << line << "] ";
  }
  ~Message() {
    if (on_) std::cerr << ss_.str() << std::endl;
    if (sev_ == FATAL) std::abort();
  }
  std::ostream& stream() { return ss_; }
 private:
  std::ostringstream ss_;
  int sev_;
  bool on_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
}  // namespace logging
}  // namespace caffe

#define LOG(sev) ::caffe::logging::Message(__FILE__, __LINE__, ::caffe::logging::sev).stream()
#define LOG_IF(sev, cond) !(cond) ? (void)0 : ::caffe::logging::Voidify() & LOG(sev)
#define DLOG(sev) LOG_IF(sev, false)
#define CHECK(cond) LOG_IF(FATAL, !(cond)) << "Check failed: " #cond " "
#define MSCNN_CHECK_OP(a, b, op) \
  LOG_IF(FATAL, !((a)op(b))) << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) MSCNN_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) MSCNN_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) MSCNN_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) MSCNN_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) MSCNN_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) MSCNN_CHECK_OP(a, b, >=)
#define DCHECK(c) CHECK(c)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define CHECK_NOTNULL(p) (p)
