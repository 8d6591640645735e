// Umbrella header, as /root/reference/include/caffe/caffe.hpp:1-21 (what matcaffe, pycaffe and tools/caffe.cpp include).
#pragma once
#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/layer.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/net.hpp"
#include "caffe/proto/caffe.pb.h"
