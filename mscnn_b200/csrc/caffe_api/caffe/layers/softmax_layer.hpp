// Source-compatibility header: see caffe/layers/mscnn_layers.hpp
#pragma once
#include "caffe/layers/mscnn_layers.hpp"
