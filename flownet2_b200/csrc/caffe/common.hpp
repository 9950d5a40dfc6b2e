// Host-side common definitions of the Caffe-surface engine (reference: include/caffe/common.hpp).
// Errors are C++ exceptions (caught at the C-ABI and turned into fn2_status codes) instead of the
// reference's glog CHECK aborts (device_alternate.hpp:48-76).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/fn2.h"

namespace caffe {

using std::shared_ptr;
using std::string;
using std::vector;

enum Phase { TRAIN = 0, TEST = 1 };

struct CaffeError : std::runtime_error {
    explicit CaffeError(const std::string& m) : std::runtime_error(m) {}
};

struct CheckFail {
    std::ostringstream os;
    CheckFail(const char* file, int line, const char* cond) { os << file << ":" << line << " Check failed: " << cond << " "; }
    template <typename T> CheckFail& operator<<(const T& v) { os << v; return *this; }
    [[noreturn]] ~CheckFail() noexcept(false) { throw CaffeError(os.str()); }
};
#define CHECK(cond) if (cond) {} else ::caffe::CheckFail(__FILE__, __LINE__, #cond)
#define CHECK_EQ(a, b) CHECK((a) == (b)) << "(" << (a) << " vs " << (b) << ") "
#define CHECK_GE(a, b) CHECK((a) >= (b)) << "(" << (a) << " vs " << (b) << ") "
#define CHECK_LE(a, b) CHECK((a) <= (b)) << "(" << (a) << " vs " << (b) << ") "
#define CHECK_GT(a, b) CHECK((a) > (b)) << "(" << (a) << " vs " << (b) << ") "
#define NOT_IMPLEMENTED CHECK(false) << "Not Implemented Yet"

#define CUDA_CHECK(call)                                                              \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        CHECK(e__ == cudaSuccess) << #call << ": " << cudaGetErrorString(e__);         \
    } while (0)

// fn2_* C-ABI call -> exception on failure
#define FN2_CALL(call)                                                                \
    do {                                                                              \
        int rc__ = (call);                                                            \
        CHECK(rc__ == 0) << #call << " -> " << rc__ << ": " << fn2_last_error();       \
    } while (0)

// The reference's Caffe singleton carries mode/device/RNG (common.hpp:100-170); here the only
// thread-local state is the stream layers enqueue on (set by Net around Forward) and the seed.
class Caffe {
 public:
    static cudaStream_t& stream() { static thread_local cudaStream_t s = 0; return s; }
    static uint64_t& seed() { static thread_local uint64_t s = 1701; return s; }
};

}  // namespace caffe
