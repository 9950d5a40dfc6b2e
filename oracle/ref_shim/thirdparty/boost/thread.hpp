// Stand-in for boost/thread.hpp -- TEST INFRASTRUCTURE for oracle/_ref.
#ifndef REF_SHIM_BOOST_THREAD_HPP_
#define REF_SHIM_BOOST_THREAD_HPP_
#include <memory>
#include <mutex>
#include <unistd.h>   // boost.thread pulls it in; common.cpp:45 relies on that for getpid()
namespace boost {
class mutex : public std::mutex {};
// One instance per thread, like boost::thread_specific_ptr (common.cpp:22).
template <typename T> class thread_specific_ptr {
 public:
  T* get() const { return slot().get(); }
  void reset(T* p) { slot().reset(p); }
 private:
  static std::unique_ptr<T>& slot() { static thread_local std::unique_ptr<T> p; return p; }
};
}
#endif
