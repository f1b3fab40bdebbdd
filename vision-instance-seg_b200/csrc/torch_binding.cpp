// Thin torch extension over the C ABI of libmsda_b200.so (include/msda_b200.h).
//
// Same two functions, same argument order and same error behaviour as upstream's compiled module
// `MultiScaleDeformableAttention` (maskdino/modeling/pixel_decoder/ops/src/vision.cpp: `ms_deform_attn_forward`,
// `ms_deform_attn_backward`), SURVEY.md section 8b.  Everything that touches the GPU happens behind the C ABI; this file only
// does what upstream's ms_deform_attn.h / ms_deform_attn_cuda.cu host code does -- check the tensors, allocate the
// results (and the scratch the C ABI asks for) with torch's caching allocator, pass the current stream -- without
// the ~20 marshalled scalars per call of the ctypes route (vision-instance-seg_b200/MultiScaleDeformableAttention.py, which
// stays as the fallback and calls the very same entry points).
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <string>
#include <vector>

#include "../../include/msda_b200.h"

namespace {

void require(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_contiguous(), name, " tensor has to be contiguous");
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
}

int dtype_code(const at::Tensor& value, const char* what) {
  switch (value.scalar_type()) {
    case at::kFloat: return MSDA_F32;
    case at::kDouble: return MSDA_F64;
    case at::kBFloat16: return MSDA_BF16;
    case at::kHalf: return MSDA_F16;
    default: TORCH_CHECK(false, what, " not implemented for '", value.scalar_type(), "'");
  }
}

struct Dims { int N, S, M, D, Lq, L, P; };

Dims dims(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& loc) {
  TORCH_CHECK(value.dim() == 4 && loc.dim() == 6, "value must be (N, S, M, D) and sampling_loc (N, Lq, M, L, P, 2)");
  Dims d{static_cast<int>(value.size(0)), static_cast<int>(value.size(1)), static_cast<int>(value.size(2)),
         static_cast<int>(value.size(3)), static_cast<int>(loc.size(1)), static_cast<int>(loc.size(3)),
         static_cast<int>(loc.size(4))};
  TORCH_CHECK(loc.size(2) == d.M && loc.size(5) == 2 && shapes.size(0) == d.L && loc.size(0) == d.N,
              "inconsistent shapes between value, spatial_shapes and sampling_loc");
  return d;
}

at::Tensor meta(const at::Tensor& t, const at::Device& dev) {
  if (t.scalar_type() == at::kLong && t.device() == dev && t.is_contiguous()) return t;
  return t.to(dev, at::kLong).contiguous();
}

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, " failed: ", msda_error_string(rc), " (code ", rc, ")");
}

at::Tensor forward(const at::Tensor& value, const at::Tensor& spatial_shapes, const at::Tensor& level_start_index,
                   const at::Tensor& sampling_loc, const at::Tensor& attn_weight, int64_t im2col_step) {
  require(value, "value");
  require(spatial_shapes, "spatial_shapes");
  require(level_start_index, "level_start_index");
  require(sampling_loc, "sampling_loc");
  require(attn_weight, "attn_weight");
  const int code = dtype_code(value, "ms_deform_attn_forward");
  const Dims d = dims(value, spatial_shapes, sampling_loc);
  const auto aux = value.scalar_type() == at::kDouble ? at::kDouble : at::kFloat;
  const at::Tensor loc = sampling_loc.scalar_type() == aux ? sampling_loc : sampling_loc.to(aux);
  const at::Tensor attn = attn_weight.scalar_type() == aux ? attn_weight : attn_weight.to(aux);
  const at::Tensor shapes = meta(spatial_shapes, value.device());
  const at::Tensor lsi = meta(level_start_index, value.device());
  c10::cuda::CUDAGuard guard(value.device());
  at::Tensor out = at::empty({d.N, d.Lq, static_cast<int64_t>(d.M) * d.D}, value.options());
  if (out.numel() == 0) return out;
  const int rc = msda_forward(value.data_ptr(), shapes.data_ptr<int64_t>(), lsi.data_ptr<int64_t>(), loc.data_ptr(),
                              attn.data_ptr(), out.data_ptr(), d.N, d.S, d.M, d.D, d.Lq, d.L, d.P, code,
                              static_cast<int>(im2col_step), c10::cuda::getCurrentCUDAStream().stream());
  check(rc, "ms_deform_attn_forward");
  return out;
}

std::vector<at::Tensor> backward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                 const at::Tensor& level_start_index, const at::Tensor& sampling_loc,
                                 const at::Tensor& attn_weight, const at::Tensor& grad_output, int64_t im2col_step,
                                 int64_t flags) {
  require(value, "value");
  require(spatial_shapes, "spatial_shapes");
  require(level_start_index, "level_start_index");
  require(sampling_loc, "sampling_loc");
  require(attn_weight, "attn_weight");
  require(grad_output, "grad_output");
  const int code = dtype_code(value, "ms_deform_attn_backward");
  const Dims d = dims(value, spatial_shapes, sampling_loc);
  const auto aux = value.scalar_type() == at::kDouble ? at::kDouble : at::kFloat;
  const at::Tensor loc = sampling_loc.scalar_type() == aux ? sampling_loc : sampling_loc.to(aux);
  const at::Tensor attn = attn_weight.scalar_type() == aux ? attn_weight : attn_weight.to(aux);
  const at::Tensor go = grad_output.scalar_type() == value.scalar_type() ? grad_output : grad_output.to(value.scalar_type());
  const at::Tensor shapes = meta(spatial_shapes, value.device());
  const at::Tensor lsi = meta(level_start_index, value.device());
  c10::cuda::CUDAGuard guard(value.device());
  at::Tensor grad_value = at::empty_like(value);
  at::Tensor grad_loc = at::empty(sampling_loc.sizes(), value.options().dtype(aux));
  at::Tensor grad_attn = at::empty(attn_weight.sizes(), value.options().dtype(aux));
  if (grad_loc.numel() == 0 || value.numel() == 0) return {grad_value.zero_(), grad_loc.zero_(), grad_attn.zero_()};
  const size_t nbytes = msda_backward_scratch_bytes(d.N, d.S, d.M, d.D, d.Lq, d.L, d.P, code, static_cast<int>(flags));
  at::Tensor scratch;
  if (nbytes) scratch = at::empty({static_cast<int64_t>(nbytes)}, value.options().dtype(at::kByte));
  const int rc = msda_backward(value.data_ptr(), shapes.data_ptr<int64_t>(), lsi.data_ptr<int64_t>(), loc.data_ptr(),
                               attn.data_ptr(), go.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(),
                               grad_attn.data_ptr(), nbytes ? scratch.data_ptr() : nullptr, nbytes, d.N, d.S, d.M, d.D,
                               d.Lq, d.L, d.P, code, static_cast<int>(im2col_step), static_cast<int>(flags),
                               c10::cuda::getCurrentCUDAStream().stream());
  check(rc, "ms_deform_attn_backward");
  if (grad_loc.scalar_type() != sampling_loc.scalar_type()) grad_loc = grad_loc.to(sampling_loc.scalar_type());
  if (grad_attn.scalar_type() != attn_weight.scalar_type()) grad_attn = grad_attn.to(attn_weight.scalar_type());
  return {grad_value, grad_loc, grad_attn};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torch extension over the C ABI of libmsda_b200.so (B200 multi-scale deformable attention)";
  // no Python object is touched inside: run without the GIL, like the ctypes route does (a launch that blocks on a full
  // queue must not stall autograd's other thread or the bench's clock sampler)
  m.def("ms_deform_attn_forward", &forward, "ms_deform_attn_forward", pybind11::call_guard<pybind11::gil_scoped_release>());
  m.def("ms_deform_attn_backward", &backward, "ms_deform_attn_backward (flags: include/msda_b200.h MSDA_BWD_*)",
        pybind11::call_guard<pybind11::gil_scoped_release>());
  m.def("abi_version", []() { return msda_abi_version(); });
  m.def("raise_for_code", [](int code) { check(code, "raise_for_code"); },
        "turn a C-ABI return code into the RuntimeError the two functions raise (0: no error); used by the CPU tests");
}
