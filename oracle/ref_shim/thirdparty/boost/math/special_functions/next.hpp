// Stand-in for boost/math/special_functions/next.hpp -- TEST INFRASTRUCTURE for oracle/_ref.
#ifndef REF_SHIM_BOOST_NEXT_HPP_
#define REF_SHIM_BOOST_NEXT_HPP_
#include <cmath>
namespace boost { namespace math { template <typename T> inline T nextafter(T a, T b) { return std::nextafter(a, b); } } }
#endif
