// Minimal glog stand-in: CHECK*/LOG/DLOG/LOG_IF with glog's abort-on-FATAL behaviour
// (the reference's error convention).  INFO/WARNING output is dropped unless
// MSCNN_REF_VERBOSE is set in the environment.
#pragma once
#include <cstdlib>
#include <cstring>  // the real glog headers pull in <cstring>/<ctime>; math_functions.hpp:68 relies on memset
#include <ctime>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
enum LogSeverity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev) : sev_(sev) {
    static const bool verbose = std::getenv("MSCNN_REF_VERBOSE") != nullptr;
    on_ = sev >= GLOG_ERROR || verbose;
    if (on_) ss_ << "[ref " << "IWEF"[sev] << " " << file << ":" << line << "] ";
  }
  ~LogMessage() {
    if (on_) std::cerr << ss_.str() << std::endl;
    if (sev_ == GLOG_FATAL) std::abort();
  }
  std::ostream& stream() { return ss_; }
 private:
  std::ostringstream ss_;
  int sev_;
  bool on_;
};
struct LogVoidify {
  void operator&(std::ostream&) {}
};
template <typename T>
T* CheckNotNull(const char* file, int line, const char* names, T* t) {
  if (t == nullptr) LogMessage(file, line, GLOG_FATAL).stream() << names;
  return t;
}
}  // namespace google

#define MSCNN_GLOG_SEV_INFO ::google::GLOG_INFO
#define MSCNN_GLOG_SEV_WARNING ::google::GLOG_WARNING
#define MSCNN_GLOG_SEV_ERROR ::google::GLOG_ERROR
#define MSCNN_GLOG_SEV_FATAL ::google::GLOG_FATAL
#define LOG(sev) ::google::LogMessage(__FILE__, __LINE__, MSCNN_GLOG_SEV_##sev).stream()
#define LOG_IF(sev, cond) \
  !(cond) ? (void)0 : ::google::LogVoidify() & LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define LOG_EVERY_N(sev, n) LOG(sev)
#define VLOG(n) LOG_IF(INFO, false)
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
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define CHECK_NOTNULL(p) ::google::CheckNotNull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))
