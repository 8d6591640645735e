// Shim: common.hpp includes gflags only for GlobalInit (src/caffe/common.cpp:43-50).
#pragma once
#define GFLAGS_GFLAGS_H_
namespace gflags {
inline unsigned ParseCommandLineFlags(int*, char***, bool) { return 0; }
}
