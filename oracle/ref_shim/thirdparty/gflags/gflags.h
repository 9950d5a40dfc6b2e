// Stand-in for gflags (absent from the image) -- TEST INFRASTRUCTURE for oracle/_ref.
#ifndef GFLAGS_GFLAGS_H_
#define GFLAGS_GFLAGS_H_
namespace gflags { inline void ParseCommandLineFlags(int*, char***, bool) {} }
#endif
