// Shim for boost/random.hpp as used by src/caffe/util/math_functions.cpp:304-401 (caffe_rng_*).
// The RNG paths are not on the inference hot path (fillers are bypassed: weights are injected),
// so distribution streams need not match boost bit-for-bit.
#pragma once
#include <random>
#include "boost/random/mersenne_twister.hpp"
#include "boost/random/uniform_int.hpp"
namespace boost {
template <typename T>
class uniform_real : public std::uniform_real_distribution<T> {
 public:
  uniform_real(T a, T b) : std::uniform_real_distribution<T>(a, b) {}
};
template <typename T>
class normal_distribution : public std::normal_distribution<T> {
 public:
  normal_distribution(T m, T s) : std::normal_distribution<T>(m, s) {}
};
template <typename T>
class bernoulli_distribution : public std::bernoulli_distribution {
 public:
  explicit bernoulli_distribution(T p) : std::bernoulli_distribution(static_cast<double>(p)) {}
};
template <typename Engine, typename Dist>
class variate_generator {
 public:
  variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
  auto operator()() { return d_(*e_); }
 private:
  Engine e_;  // a pointer type (caffe::rng_t*)
  Dist d_;
};
}  // namespace boost
