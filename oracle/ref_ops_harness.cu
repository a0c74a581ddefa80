// TEST INFRASTRUCTURE (oracle/): C entry point that instantiates one of the reference's re-ID ops - compiled UNMODIFIED from
// /root/reference/detectron/ops/ against oracle/caffe2_shim - by its registered name and runs it on caller-owned buffers.
// Built by oracle/build_ref_ops.py into oracle/_ref/libref_reid_ops.so; loaded only by tests/ and oracle/ scripts.
#include <cuda_runtime.h>

#include <cstdio>
#include <memory>

#include "caffe2/core/context_gpu.h"

namespace {
void set_err(char* err, int errlen, const char* msg) {
  if (err && errlen > 0) std::snprintf(err, (size_t)errlen, "%s", msg);
}
}  // namespace

// device: 0 = the CPU registry (host pointers), 1 = the CUDA registry (device pointers, launched on `stream`).
// in_dims / out_dims: 4 int64 per tensor.  Outputs: caller passes capacities in bytes; on return out_ndim / out_dims hold
// the shapes the op set.  Returns 0, or -1 with a message in err (unknown op, failed enforce, RunOnDevice false, CUDA error).
extern "C" int ref_op_run(const char* name, int device, int n_in, const void* const* in, const int* in_ndim,
                          const int64_t* in_dims, int n_out, void* const* out, const int64_t* out_capacity_bytes,
                          int* out_ndim, int64_t* out_dims, void* stream, char* err, int errlen) {
  try {
    if (device < 0 || device > 1) { set_err(err, errlen, "device must be 0 or 1"); return -1; }
    auto& reg = caffe2::ShimRegistry(device);
    auto it = reg.find(name);
    if (it == reg.end()) { set_err(err, errlen, "no such operator registered for this device"); return -1; }
    std::unique_ptr<caffe2::OperatorBase> op(it->second());
    std::vector<caffe2::Tensor> tin((size_t)n_in), tout((size_t)n_out);
    for (int i = 0; i < n_in; ++i) {
      std::vector<int64_t> d(in_dims + 4 * i, in_dims + 4 * i + in_ndim[i]);
      tin[(size_t)i] = caffe2::Tensor(const_cast<void*>(in[i]), (size_t)-1, d);
      op->inputs_.push_back(&tin[(size_t)i]);
    }
    for (int i = 0; i < n_out; ++i) {
      tout[(size_t)i] = caffe2::Tensor(out[i], (size_t)out_capacity_bytes[i], {});
      op->outputs_.push_back(&tout[(size_t)i]);
    }
    op->SetStream(stream);
    if (!op->RunOnDevice()) { set_err(err, errlen, "RunOnDevice returned false"); return -1; }
    if (device == 1) {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { set_err(err, errlen, cudaGetErrorString(e)); return -1; }
    }
    for (int i = 0; i < n_out; ++i) {
      const auto& d = tout[(size_t)i].sizes();
      out_ndim[i] = (int)d.size();
      for (size_t k = 0; k < d.size() && k < 4; ++k) out_dims[4 * i + (int)k] = d[k];
    }
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
}

// the gradient definition the reference registers for `name` (REGISTER_GRADIENT): "GradOpType|in0,in1,...|out0,..."
extern "C" int ref_op_gradient_def(const char* name, int n_in, int n_out, char* buf, int buflen) {
  auto& g = caffe2::ShimGradients();
  auto it = g.find(name);
  if (it == g.end()) return -1;
  caffe2::OperatorDef def;
  def.type = name;
  for (int i = 0; i < n_in; ++i) def.inputs.push_back("I" + std::to_string(i));
  for (int i = 0; i < n_out; ++i) def.outputs.push_back("O" + std::to_string(i));
  std::vector<std::string> go;
  for (int i = 0; i < n_out; ++i) go.push_back("GO" + std::to_string(i));
  std::unique_ptr<caffe2::GradientMakerBase> mk(it->second(def, go));
  auto defs = mk->GetGradientDefs();
  if (defs.size() != 1) return -1;
  std::string s = defs[0].type + "|";
  for (size_t i = 0; i < defs[0].inputs.size(); ++i) s += (i ? "," : "") + defs[0].inputs[i];
  s += "|";
  for (size_t i = 0; i < defs[0].outputs.size(); ++i) s += (i ? "," : "") + defs[0].outputs[i];
  std::snprintf(buf, (size_t)buflen, "%s", s.c_str());
  return 0;
}

extern "C" int ref_op_schema(const char* name, int* n_in, int* n_out) {
  auto& m = caffe2::ShimSchemas();
  auto it = m.find(name);
  if (it == m.end()) return -1;
  *n_in = it->second.n_in;
  *n_out = it->second.n_out;
  return 0;
}
