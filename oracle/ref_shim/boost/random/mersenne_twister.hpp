#pragma once
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
}
