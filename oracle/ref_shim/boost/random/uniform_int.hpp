#pragma once
#include <random>
namespace boost {
template <typename T = int>
class uniform_int : public std::uniform_int_distribution<T> {
 public:
  uniform_int(T a, T b) : std::uniform_int_distribution<T>(a, b) {}
};
}  // namespace boost
