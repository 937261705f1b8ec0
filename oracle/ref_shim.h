// Force-included (nvcc -include) in front of every UNMODIFIED reference source when building
// oracle/_ref/vren.so from /root/reference/models/csrc (see oracle/build_ref.sh).
// TEST INFRASTRUCTURE ONLY -- nothing in the product path may load oracle/_ref.
//
// It makes the reference compile on torch 2.11 / CUDA 12.9 without editing a single reference line:
//  * AT_DISPATCH_*(x.type(), ...) expands to ::detail::scalar_type(the_type); torch 2.11 dropped the
//    DeprecatedTypeProperties overload, so it is supplied here (same value the old overload returned).
//  * thrust::device / thrust::reduce are used inside the reference kernels but their headers are no
//    longer pulled in transitively by <thrust/scan.h>.
// No arithmetic is touched.
#pragma once
#include <torch/extension.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
