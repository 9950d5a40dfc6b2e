// Shadows the reference's umbrella header include/caffe/caffe.hpp, which also drags in net/solver/parallel (not part of
// the layer path compiled for oracle/_ref).  TEST INFRASTRUCTURE.
#ifndef CAFFE_CAFFE_HPP_
#define CAFFE_CAFFE_HPP_
#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/filler.hpp"
#include "caffe/layer.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/proto/caffe.pb.h"
#include "caffe/util/benchmark.hpp"
#include "caffe/util/io.hpp"
#endif
