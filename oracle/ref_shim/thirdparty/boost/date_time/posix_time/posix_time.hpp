// Stand-in for boost::posix_time on top of <chrono> (util/benchmark.cpp uses it for CPU timing) -- TEST INFRASTRUCTURE.
#ifndef REF_SHIM_BOOST_POSIX_TIME_HPP_
#define REF_SHIM_BOOST_POSIX_TIME_HPP_
#include <chrono>
namespace boost { namespace posix_time {
struct time_duration {
  std::chrono::steady_clock::duration d;
  long long total_microseconds() const { return std::chrono::duration_cast<std::chrono::microseconds>(d).count(); }
  long long total_milliseconds() const { return std::chrono::duration_cast<std::chrono::milliseconds>(d).count(); }
};
struct ptime {
  std::chrono::steady_clock::time_point t;
  time_duration operator-(const ptime& o) const { return time_duration{t - o.t}; }
};
struct microsec_clock { static ptime local_time() { return ptime{std::chrono::steady_clock::now()}; } };
}}
#endif
