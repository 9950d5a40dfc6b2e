// Stand-in for boost/random.hpp on top of <random> -- TEST INFRASTRUCTURE for oracle/_ref.
// The engine is the same algorithm (mt19937); the distributions are libstdc++'s, so the stream of VALUES differs from
// boost's (the reference's boost version is unpinned anyway) -- RNG parity is statistical only.
#ifndef REF_SHIM_BOOST_RANDOM_HPP_
#define REF_SHIM_BOOST_RANDOM_HPP_
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
namespace random { typedef std::mt19937 mt19937; }
template <typename T = double> using uniform_real = std::uniform_real_distribution<T>;
template <typename T = int> using uniform_int = std::uniform_int_distribution<T>;
template <typename T = double> using normal_distribution = std::normal_distribution<T>;
template <typename T = double> class bernoulli_distribution {
 public:
  explicit bernoulli_distribution(T p = T(0.5)) : d_((double)p) {}
  template <class E> bool operator()(E& e) { return d_(e); }
 private:
  std::bernoulli_distribution d_;
};
template <class EnginePtr, class Dist> class variate_generator {
 public:
  variate_generator(EnginePtr e, Dist d) : e_(e), d_(d) {}
  auto operator()() -> decltype(std::declval<Dist&>()(*std::declval<EnginePtr>())) { return d_(*e_); }
 private:
  EnginePtr e_; Dist d_;
};
}
#endif
