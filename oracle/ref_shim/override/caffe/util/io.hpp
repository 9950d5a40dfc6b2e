// Shadows the reference's include/caffe/util/io.hpp (boost::filesystem, protobuf IO, OpenCV, HDF5: none in the image).
// TEST INFRASTRUCTURE for oracle/_ref.  The only call from the compiled layer sources is the debug dump at
// data_augmentation_layer.cpp:68 (a hard-coded lab path), which is a no-op here.
#ifndef CAFFE_UTIL_IO_H_
#define CAFFE_UTIL_IO_H_
#include <string>
#include "caffe/common.hpp"
#include "caffe/proto/caffe.pb.h"
namespace caffe {
inline void WriteProtoToTextFile(const ::google::protobuf::Message&, const char*) {}
inline void WriteProtoToTextFile(const ::google::protobuf::Message&, const std::string&) {}
}
#endif
