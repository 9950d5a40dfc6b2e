// Stand-in for libprotobuf's reflection headers -- TEST INFRASTRUCTURE for oracle/_ref (see miniproto_rt.hpp).
#include "miniproto_rt.hpp"
