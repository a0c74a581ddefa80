// TEST INFRASTRUCTURE (oracle/): see caffe2/core/operator.h of this shim.  math::Set is the one routine the ops call.
#pragma once
#include "caffe2/core/context.h"
#include <cstdint>
namespace caffe2 {
namespace math {
template <typename T, class Context>
void Set(const int64_t n, const T alpha, T* y, Context* context);
template <>
inline void Set<float, CPUContext>(const int64_t n, const float alpha, float* y, CPUContext*) {
  for (int64_t i = 0; i < n; ++i) y[i] = alpha;
}
}  // namespace math
}  // namespace caffe2
