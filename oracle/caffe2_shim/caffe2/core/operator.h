// TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the parts of the Caffe2 operator interface (pytorch v1.0.1, absent
// from /root/reference and from this image) that the reference's re-ID custom ops touch, so that their UNMODIFIED sources
//   /root/reference/detectron/ops/batch_hard_op.cc            (BatchHard, BatchHardGradient: CPU)
//   /root/reference/detectron/ops/pairwise_distance_op.{cc,cu} (PairWiseDistance, PairWiseDistanceGradient: CUDA)
// compile where they lie (oracle/build_ref_ops.py -> oracle/_ref/libref_reid_ops.so) and can be run as the checker of
// pps_b200/csrc/triplet.cu and of the oracle's restatement.  Nothing here is Caffe2 code: it supplies containers
// (Tensor = dims + a caller-owned buffer), the Operator base with Input / Output, the enforce macros and registries.
// The arithmetic under test is entirely the reference's.
#pragma once

#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace caffe2 {

using std::string;
using std::vector;

struct OperatorDef {
  string type;
  vector<string> inputs, outputs;
};
struct Workspace {};

class EnforceNotMet : public std::runtime_error {
 public:
  explicit EnforceNotMet(const string& what) : std::runtime_error(what) {}
};

#define CAFFE_ENFORCE_EQ(a, b, ...)                                                                  \
  do {                                                                                               \
    if (!((a) == (b))) {                                                                             \
      std::ostringstream _os;                                                                        \
      _os << "Enforce failed: " #a " == " #b " (" << (a) << " vs " << (b) << ")";                    \
      throw ::caffe2::EnforceNotMet(_os.str());                                                      \
    }                                                                                                \
  } while (0)
#define CAFFE_ENFORCE(cond, ...)                                                                     \
  do {                                                                                               \
    if (!(cond)) throw ::caffe2::EnforceNotMet("Enforce failed: " #cond);                            \
  } while (0)

// dims + a buffer the harness owns (capacity checked when an op resizes an output)
class Tensor {
 public:
  Tensor() = default;
  Tensor(void* ptr, size_t capacity_bytes, const vector<int64_t>& dims) : ptr_(ptr), cap_(capacity_bytes), dims_(dims) {}
  int dim() const { return (int)dims_.size(); }
  int ndim() const { return (int)dims_.size(); }
  int dim32(int i) const { return (int)dims_.at((size_t)i); }
  int64_t numel() const {
    int64_t n = 1;
    for (int64_t d : dims_) n *= d;
    return n;
  }
  const vector<int64_t>& sizes() const { return dims_; }
  template <class... Ts>
  void Resize(Ts... d) { dims_ = {static_cast<int64_t>(d)...}; }
  void ResizeLike(const Tensor& o) { dims_ = o.dims_; }
  template <class T>
  const T* data() const { return static_cast<const T*>(ptr_); }
  template <class T>
  T* mutable_data() {
    if ((size_t)numel() * sizeof(T) > cap_) throw EnforceNotMet("shim: output buffer too small for the shape the op set");
    return static_cast<T*>(ptr_);
  }

 private:
  void* ptr_ = nullptr;
  size_t cap_ = 0;
  vector<int64_t> dims_;
};

class OperatorBase {
 public:
  virtual ~OperatorBase() {}
  virtual bool RunOnDevice() = 0;
  virtual void SetStream(void*) {}
  vector<const Tensor*> inputs_;
  vector<Tensor*> outputs_;
};

template <class Context>
class Operator : public OperatorBase {
 public:
  Operator(const OperatorDef&, Workspace*) {}
  const Tensor& Input(int i) { return *inputs_.at((size_t)i); }
  Tensor* Output(int i) { return outputs_.at((size_t)i); }
  void SetStream(void* s) override { context_.SetStream(s); }

 protected:
  Context context_;
};

#define USE_OPERATOR_CONTEXT_FUNCTIONS            \
  using Operator<Context>::Input;                 \
  using Operator<Context>::Output;                \
  using Operator<Context>::context_

// ---- registries: name -> factory, one per device type ----
using OpFactory = std::function<OperatorBase*()>;
inline std::map<string, OpFactory>& ShimRegistry(int device) {
  static std::map<string, OpFactory> reg[2];
  return reg[device];
}
struct ShimRegistrar {
  ShimRegistrar(int device, const char* name, OpFactory f) { ShimRegistry(device)[name] = f; }
};
#define SHIM_CAT2(a, b) a##b
#define SHIM_CAT(a, b) SHIM_CAT2(a, b)
#define REGISTER_CPU_OPERATOR(name, ...)                                                      \
  static ::caffe2::ShimRegistrar SHIM_CAT(shim_cpu_reg_##name, __LINE__)(                     \
      0, #name, []() -> ::caffe2::OperatorBase* { return new __VA_ARGS__(::caffe2::OperatorDef(), nullptr); })
#define REGISTER_CUDA_OPERATOR(name, ...)                                                     \
  static ::caffe2::ShimRegistrar SHIM_CAT(shim_cuda_reg_##name, __LINE__)(                    \
      1, #name, []() -> ::caffe2::OperatorBase* { return new __VA_ARGS__(::caffe2::OperatorDef(), nullptr); })

// ---- schema / gradient registration: recorded, not interpreted ----
class OpSchema {
 public:
  OpSchema& NumInputs(int n) { n_in = n; return *this; }
  OpSchema& NumOutputs(int n) { n_out = n; return *this; }
  OpSchema& IdenticalTypeAndShapeOfInputDim(int, int) { return *this; }
  OpSchema& SetDoc(const char*) { return *this; }
  OpSchema& Input(int, const char*, const char*) { return *this; }
  OpSchema& Output(int, const char*, const char*) { return *this; }
  int n_in = -1, n_out = -1;
};
inline std::map<string, OpSchema>& ShimSchemas() {
  static std::map<string, OpSchema> m;
  return m;
}
#define OPERATOR_SCHEMA(name) static ::caffe2::OpSchema& SHIM_CAT(shim_schema_##name, __LINE__) = ::caffe2::ShimSchemas()[#name]

class GradientMakerBase {
 public:
  GradientMakerBase(const OperatorDef& def, const vector<string>& g_output) : def_(def), g_output_(g_output) {}
  virtual ~GradientMakerBase() {}
  virtual vector<OperatorDef> GetGradientDefs() = 0;

 protected:
  string I(int i) const { return def_.inputs.at((size_t)i); }
  string O(int i) const { return def_.outputs.at((size_t)i); }
  string GI(int i) const { return I(i) + "_grad"; }
  string GO(int i) const { return g_output_.at((size_t)i); }
  static vector<OperatorDef> SingleGradientDef(const string& type, const string&, const vector<string>& in,
                                               const vector<string>& out) {
    OperatorDef d;
    d.type = type; d.inputs = in; d.outputs = out;
    return vector<OperatorDef>{d};
  }
  OperatorDef def_;
  vector<string> g_output_;
};
using GradFactory = std::function<GradientMakerBase*(const OperatorDef&, const vector<string>&)>;
inline std::map<string, GradFactory>& ShimGradients() {
  static std::map<string, GradFactory> m;
  return m;
}
struct ShimGradRegistrar {
  ShimGradRegistrar(const char* name, GradFactory f) { ShimGradients()[name] = f; }
};
#define REGISTER_GRADIENT(name, ...)                                                                         \
  static ::caffe2::ShimGradRegistrar SHIM_CAT(shim_grad_reg_##name, __LINE__)(                               \
      #name, [](const ::caffe2::OperatorDef& d, const ::caffe2::vector<::caffe2::string>& g) -> ::caffe2::GradientMakerBase* { \
        return new __VA_ARGS__(d, g);                                                                        \
      })

}  // namespace caffe2
