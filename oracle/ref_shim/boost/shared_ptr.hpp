// Shim: the reference only needs boost::shared_ptr as a type (include/caffe/common.hpp:4,74).
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::weak_ptr;
}  // namespace boost
