// TEST INFRASTRUCTURE (oracle/): see caffe2/core/operator.h of this shim.
#pragma once
namespace caffe2 {
class CPUContext {
 public:
  void SetStream(void*) {}
};
}  // namespace caffe2
