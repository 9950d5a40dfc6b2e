// Stand-in for boost::shared_ptr (boost is absent from the image) -- TEST INFRASTRUCTURE for oracle/_ref.
#ifndef REF_SHIM_BOOST_SHARED_PTR_HPP_
#define REF_SHIM_BOOST_SHARED_PTR_HPP_
#include <cfloat>   // boost.config pulls <cfloat> in; data_augmentation_layer.cu:166 relies on that for FLT_MAX
#include <memory>
namespace boost {
using std::shared_ptr; using std::weak_ptr; using std::dynamic_pointer_cast; using std::static_pointer_cast;
}
#endif
