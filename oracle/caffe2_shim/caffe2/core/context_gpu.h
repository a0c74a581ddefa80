// TEST INFRASTRUCTURE (oracle/): see caffe2/core/operator.h of this shim.  The launch geometry follows Caffe2's
// common_gpu.h of pytorch v1.0.1 (128 threads per block, at most 4096 blocks, grid-stride loop).
#pragma once
#include <cuda_runtime.h>

#include "caffe2/core/context.h"
#include "caffe2/core/operator.h"
#include "caffe2/utils/math.h"

namespace caffe2 {

class CUDAContext {
 public:
  void SetStream(void* s) { stream_ = static_cast<cudaStream_t>(s); }
  cudaStream_t cuda_stream() const { return stream_; }

 private:
  cudaStream_t stream_ = nullptr;
};

constexpr int CAFFE_CUDA_NUM_THREADS = 128;
constexpr int CAFFE_MAXIMUM_NUM_BLOCKS = 4096;
inline int CAFFE_GET_BLOCKS(const int N) {
  int b = (N + CAFFE_CUDA_NUM_THREADS - 1) / CAFFE_CUDA_NUM_THREADS;
  if (b > CAFFE_MAXIMUM_NUM_BLOCKS) b = CAFFE_MAXIMUM_NUM_BLOCKS;
  return b < 1 ? 1 : b;
}
#define CUDA_1D_KERNEL_LOOP(i, n) \
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += blockDim.x * gridDim.x)

#ifdef __CUDACC__
namespace shim_detail {
template <typename T>
__global__ void FillKernel(const int64_t n, const T alpha, T* y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)blockDim.x * gridDim.x) y[i] = alpha;
}
}  // namespace shim_detail
namespace math {
template <>
inline void Set<float, CUDAContext>(const int64_t n, const float alpha, float* y, CUDAContext* context) {
  if (n <= 0) return;
  if (alpha == 0.f) {
    cudaMemsetAsync(y, 0, (size_t)n * sizeof(float), context->cuda_stream());
  } else {
    shim_detail::FillKernel<float><<<CAFFE_GET_BLOCKS((int)(n > (1 << 30) ? (1 << 30) : n)), CAFFE_CUDA_NUM_THREADS, 0,
                                     context->cuda_stream()>>>(n, alpha, y);
  }
}
}  // namespace math
#endif

}  // namespace caffe2
