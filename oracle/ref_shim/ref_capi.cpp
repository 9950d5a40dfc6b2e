// ref_capi.cpp -- C entry points around the REFERENCE's own layer classes (oracle/_ref/libref_caffe.so).
//
// TEST INFRASTRUCTURE ONLY.  oracle/ref_shim/build_ref.py compiles the reference's layer / blob / math sources where they
// lie under /root/reference (unmodified; only third-party libraries absent from the image are replaced by the stand-ins in
// ref_shim/thirdparty) and links them with this file.  Everything a test does goes through caffe::Layer<float>::SetUp /
// Forward / Backward of the reference (include/caffe/layer.hpp:67-81,483-535), i.e. the reference's own host code chooses
// launch configurations, scratch blobs and cuBLAS calls.
//
// Also here: the CBLAS symbols math_functions.cpp links against, forwarded to the OpenBLAS inside the image's scipy / numpy
// wheels (the reference's `BLAS := open` option, Makefile.config.example:46), and the two registrations that live in the
// reference's layer_factory.cpp (not compiled: it needs cuDNN and Python headers).
#include <dlfcn.h>

#include <cstring>
#include <string>
#include <vector>

#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/layer.hpp"
#include "caffe/layer_factory.hpp"
#include "caffe/layers/conv_layer.hpp"
#include "caffe/layers/relu_layer.hpp"
#include "caffe/proto/caffe.pb.h"
#include "caffe/util/benchmark.hpp"
// Test-only peek at DataAugmentationLayer's protected scratch (the chromatic eigenspace statistics), see
// ref_layer_debug_eigenspace below.  Access specifiers do not change the class layout.
#define protected public
#include "caffe/layers/data_augmentation_layer.hpp"
#undef protected

int FLAGS_logtostderr = 0, FLAGS_alsologtostderr = 0, FLAGS_minloglevel = 0, FLAGS_v = 0;

namespace caffe {
// layer_factory.cpp:37-71 picks ConvolutionLayer for `engine: CAFFE` (and DEFAULT without cuDNN), :149-170 ReLULayer.
REGISTER_LAYER_CLASS(Convolution);
REGISTER_LAYER_CLASS(ReLU);
}  // namespace caffe

// ---- CBLAS -> OpenBLAS ------------------------------------------------------------------------------------------------
namespace {
void* g_blas = nullptr;
bool g_ilp64 = false;
std::string g_err;

template <typename F> F sym(const char* name) {
  if (!g_blas) throw std::runtime_error("ref: no BLAS loaded (call ref_load_blas)");
  std::string n = std::string("scipy_") + name + (g_ilp64 ? "64_" : "");
  void* p = dlsym(g_blas, n.c_str());
  if (!p) throw std::runtime_error("ref: BLAS symbol missing: " + n);
  return (F)p;
}
typedef long long L;
}  // namespace

extern "C" {
void cblas_sgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K, float alpha, const float* A,
                 int lda, const float* B, int ldb, float beta, float* C, int ldc) {
  if (g_ilp64) {
    static auto f = sym<void (*)(int, int, int, L, L, L, float, const float*, L, const float*, L, float, float*, L)>("cblas_sgemm");
    f(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  } else {
    static auto f = sym<void (*)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int)>("cblas_sgemm");
    f(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  }
}
void cblas_dgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K, double alpha, const double* A,
                 int lda, const double* B, int ldb, double beta, double* C, int ldc) {
  if (g_ilp64) {
    static auto f = sym<void (*)(int, int, int, L, L, L, double, const double*, L, const double*, L, double, double*, L)>("cblas_dgemm");
    f(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  } else {
    static auto f = sym<void (*)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int)>("cblas_dgemm");
    f(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  }
}
void cblas_sgemv(CBLAS_ORDER o, CBLAS_TRANSPOSE t, int M, int N, float alpha, const float* A, int lda, const float* x, int incx,
                 float beta, float* y, int incy) {
  if (g_ilp64) {
    static auto f = sym<void (*)(int, int, L, L, float, const float*, L, const float*, L, float, float*, L)>("cblas_sgemv");
    f(o, t, M, N, alpha, A, lda, x, incx, beta, y, incy);
  } else {
    static auto f = sym<void (*)(int, int, int, int, float, const float*, int, const float*, int, float, float*, int)>("cblas_sgemv");
    f(o, t, M, N, alpha, A, lda, x, incx, beta, y, incy);
  }
}
void cblas_dgemv(CBLAS_ORDER o, CBLAS_TRANSPOSE t, int M, int N, double alpha, const double* A, int lda, const double* x,
                 int incx, double beta, double* y, int incy) {
  if (g_ilp64) {
    static auto f = sym<void (*)(int, int, L, L, double, const double*, L, const double*, L, double, double*, L)>("cblas_dgemv");
    f(o, t, M, N, alpha, A, lda, x, incx, beta, y, incy);
  } else {
    static auto f = sym<void (*)(int, int, int, int, double, const double*, int, const double*, int, double, double*, int)>("cblas_dgemv");
    f(o, t, M, N, alpha, A, lda, x, incx, beta, y, incy);
  }
}
// Level-1 routines: plain loops (bandwidth-trivial; keeps the ILP64/LP64 plumbing to the level-2/3 calls that matter).
void cblas_saxpy(int n, float a, const float* x, int incx, float* y, int incy) { for (int i = 0; i < n; ++i) y[(size_t)i * incy] += a * x[(size_t)i * incx]; }
void cblas_daxpy(int n, double a, const double* x, int incx, double* y, int incy) { for (int i = 0; i < n; ++i) y[(size_t)i * incy] += a * x[(size_t)i * incx]; }
void cblas_sscal(int n, float a, float* x, int incx) { for (int i = 0; i < n; ++i) x[(size_t)i * incx] *= a; }
void cblas_dscal(int n, double a, double* x, int incx) { for (int i = 0; i < n; ++i) x[(size_t)i * incx] *= a; }
void cblas_scopy(int n, const float* x, int incx, float* y, int incy) { for (int i = 0; i < n; ++i) y[(size_t)i * incy] = x[(size_t)i * incx]; }
void cblas_dcopy(int n, const double* x, int incx, double* y, int incy) { for (int i = 0; i < n; ++i) y[(size_t)i * incy] = x[(size_t)i * incx]; }
float cblas_sdot(int n, const float* x, int incx, const float* y, int incy) { float s = 0; for (int i = 0; i < n; ++i) s += x[(size_t)i * incx] * y[(size_t)i * incy]; return s; }
double cblas_ddot(int n, const double* x, int incx, const double* y, int incy) { double s = 0; for (int i = 0; i < n; ++i) s += x[(size_t)i * incx] * y[(size_t)i * incy]; return s; }
float cblas_sasum(int n, const float* x, int incx) { float s = 0; for (int i = 0; i < n; ++i) s += std::fabs(x[(size_t)i * incx]); return s; }
double cblas_dasum(int n, const double* x, int incx) { double s = 0; for (int i = 0; i < n; ++i) s += std::fabs(x[(size_t)i * incx]); return s; }
}

// ---- C API ------------------------------------------------------------------------------------------------------------
using caffe::Blob;
using caffe::Caffe;
using caffe::Layer;

struct RefLayer {
  boost::shared_ptr<Layer<float> > layer;
  std::vector<Blob<float>*> bottom, top;
};

#define REF_TRY try {
#define REF_CATCH(rv)                                   \
  }                                                     \
  catch (const std::exception& e) { g_err = e.what(); return rv; }

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char* ref_last_error() { return g_err.c_str(); }

REF_API int ref_load_blas(const char* path, int ilp64) {
  g_blas = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!g_blas) { g_err = dlerror(); return -1; }
  g_ilp64 = ilp64 != 0;
  return 0;
}
REF_API int ref_blas_set_threads(int n) {
  REF_TRY
  if (g_ilp64) sym<void (*)(int)>("openblas_set_num_threads")(n);
  else { void* p = dlsym(g_blas, "scipy_openblas_set_num_threads"); if (!p) throw std::runtime_error("no set_num_threads"); ((void (*)(int))p)(n); }
  return 0;
  REF_CATCH(-1)
}

// mode: 0 CPU, 1 GPU (Caffe::set_mode, common.hpp:139); device = cudaSetDevice ordinal.
REF_API int ref_set_mode(int gpu, int device) {
  REF_TRY
  if (gpu) { Caffe::SetDevice(device); Caffe::set_mode(Caffe::GPU); } else Caffe::set_mode(Caffe::CPU);
  return 0;
  REF_CATCH(-1)
}
REF_API int ref_set_seed(unsigned seed) { REF_TRY Caffe::set_random_seed(seed); return 0; REF_CATCH(-1) }

REF_API void* ref_blob_create() { return new Blob<float>(); }
REF_API void ref_blob_destroy(void* b) { delete (Blob<float>*)b; }
REF_API int ref_blob_reshape(void* b, int naxes, const int* shape) {
  REF_TRY ((Blob<float>*)b)->Reshape(std::vector<int>(shape, shape + naxes)); return 0; REF_CATCH(-1)
}
REF_API int ref_blob_num_axes(void* b) { return ((Blob<float>*)b)->num_axes(); }
REF_API int ref_blob_shape(void* b, int* shape) {
  auto* bl = (Blob<float>*)b;
  for (int i = 0; i < bl->num_axes(); ++i) shape[i] = bl->shape(i);
  return bl->num_axes();
}
REF_API long long ref_blob_count(void* b) { return ((Blob<float>*)b)->count(); }
REF_API int ref_blob_set(void* b, const float* host, int diff) {
  REF_TRY
  auto* bl = (Blob<float>*)b;
  float* dst = diff ? bl->mutable_cpu_diff() : bl->mutable_cpu_data();
  std::memcpy(dst, host, sizeof(float) * bl->count());
  return 0;
  REF_CATCH(-1)
}
REF_API int ref_blob_get(void* b, float* host, int diff) {
  REF_TRY
  auto* bl = (Blob<float>*)b;
  const float* src = diff ? bl->cpu_diff() : bl->cpu_data();
  std::memcpy(host, src, sizeof(float) * bl->count());
  return 0;
  REF_CATCH(-1)
}
// Device pointer of the blob's data (GPU mode), for timing / device-side comparisons.
REF_API const void* ref_blob_gpu_data(void* b) { REF_TRY return ((Blob<float>*)b)->gpu_data(); REF_CATCH(nullptr) }

// layer_prototxt: the text of ONE LayerParameter (the inside of a `layer { ... }` block).  phase: 0 TRAIN, 1 TEST.
REF_API void* ref_layer_create(const char* layer_prototxt, int phase) {
  REF_TRY
  caffe::LayerParameter p;
  ::google::protobuf::ParseTextInto(layer_prototxt, &p);
  p.set_phase(phase ? caffe::TEST : caffe::TRAIN);
  auto* h = new RefLayer();
  h->layer = caffe::LayerRegistry<float>::CreateLayer(p);
  return h;
  REF_CATCH(nullptr)
}
REF_API void ref_layer_destroy(void* h) { delete (RefLayer*)h; }
REF_API int ref_layer_setup(void* hv, int nb, void** bottoms, int nt, void** tops) {
  REF_TRY
  auto* h = (RefLayer*)hv;
  h->bottom.assign((Blob<float>**)bottoms, (Blob<float>**)bottoms + nb);
  h->top.assign((Blob<float>**)tops, (Blob<float>**)tops + nt);
  h->layer->SetUp(h->bottom, h->top);
  return 0;
  REF_CATCH(-1)
}
REF_API int ref_layer_reshape(void* hv) { REF_TRY auto* h = (RefLayer*)hv; h->layer->Reshape(h->bottom, h->top); return 0; REF_CATCH(-1) }
REF_API int ref_layer_forward(void* hv, float* loss) {
  REF_TRY
  auto* h = (RefLayer*)hv;
  float l = h->layer->Forward(h->bottom, h->top);
  if (loss) *loss = l;
  if (Caffe::mode() == Caffe::GPU) CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
  REF_CATCH(-1)
}
REF_API int ref_layer_backward(void* hv, const int* propagate_down) {
  REF_TRY
  auto* h = (RefLayer*)hv;
  std::vector<bool> pd(h->bottom.size());
  for (size_t i = 0; i < pd.size(); ++i) pd[i] = propagate_down ? propagate_down[i] != 0 : true;
  h->layer->Backward(h->top, pd, h->bottom);
  if (Caffe::mode() == Caffe::GPU) CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
  REF_CATCH(-1)
}
REF_API int ref_layer_num_params(void* hv) { return (int)((RefLayer*)hv)->layer->blobs().size(); }
REF_API void* ref_layer_param(void* hv, int i) { return ((RefLayer*)hv)->layer->blobs()[i].get(); }
REF_API const char* ref_layer_type(void* hv) { return ((RefLayer*)hv)->layer->type(); }
// Net::CopyTrainedLayersFrom's special case (net.cpp:769-778): layers with DoesUseCustomCopyBlobs() get the source blobs.
REF_API int ref_layer_uses_custom_copy(void* hv) { return ((RefLayer*)hv)->layer->DoesUseCustomCopyBlobs() ? 1 : 0; }
REF_API int ref_layer_custom_copy(void* hv, int n, void** blobs) {
  REF_TRY
  std::vector<Blob<float>*> v((Blob<float>**)blobs, (Blob<float>**)blobs + n);
  ((RefLayer*)hv)->layer->CustomCopyBlobs(v);
  return 0;
  REF_CATCH(-1)
}
// Average forward time over `iters` runs after one warm-up, with the reference's own Timer (util/benchmark.cpp: CUDA events
// in GPU mode, wall clock in CPU mode) -- the `caffe time` procedure (tools/caffe.cpp:346-385).
REF_API int ref_layer_time_forward(void* hv, int iters, float* ms) {
  REF_TRY
  auto* h = (RefLayer*)hv;
  h->layer->Forward(h->bottom, h->top);
  if (Caffe::mode() == Caffe::GPU) CUDA_CHECK(cudaDeviceSynchronize());
  caffe::Timer t;
  t.Start();
  for (int i = 0; i < iters; ++i) h->layer->Forward(h->bottom, h->top);
  *ms = t.MilliSeconds() / iters;
  return 0;
  REF_CATCH(-1)
}

// tChromaticEigenSpace (25 floats: mean_eig[3] mean_rgb[3] max_abs_eig[3] max_rgb[3] min_rgb[3] max_l eigvec[9],
// augmentation_layer_base.hpp:117-129) as the last Forward left it.  Needed because the reference's statistics are RACY:
// fatomicMax / fatomicMin (data_augmentation_layer.cu:117-143) assign atomicCAS's unsigned return value to a float (a numeric
// conversion, not a bit cast), so a thread whose first compare-and-swap loses against a concurrent writer gives up and its
// larger value is dropped -- max_abs_eig / max_rgb / min_rgb come out as "some pixel's value", not the extremum.  The golden
// vectors therefore carry the statistics the reference actually used, and the per-pixel transform is pinned against those.
REF_API int ref_layer_debug_eigenspace(void* hv, float* out25) {
  REF_TRY
  auto* l = dynamic_cast<caffe::DataAugmentationLayer<float>*>(((RefLayer*)hv)->layer.get());
  if (!l) throw std::runtime_error("not a DataAugmentation layer");
  if (!l->chromatic_eigenspace_) throw std::runtime_error("layer not set up");
  std::memcpy(out25, l->chromatic_eigenspace_->cpu_data(), 25 * sizeof(float));
  return 0;
  REF_CATCH(-1)
}
