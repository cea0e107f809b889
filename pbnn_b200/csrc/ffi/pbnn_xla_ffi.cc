// pbnn_b200: XLA FFI handlers over the C ABI of include/pbnn_b200.h -- the "thin jax.ffi layer" of BASELINE.json's
// north_star.  One handler per sampler entry point; each one only unpacks XLA buffers / attributes into the C call.
//
// Build (needs the FFI headers that ship inside jaxlib >= 0.4.31, `python -c "import jaxlib, os;
// print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))"`):
//     g++ -O2 -std=c++17 -shared -fPIC -I$JAXLIB_INCLUDE -I include pbnn_b200/csrc/ffi/pbnn_xla_ffi.cc \
//         -L pbnn_b200 -lpbnn_b200 -o pbnn_b200/libpbnn_b200_xla_ffi.so
// __graft_entry__.build() compiles it when the headers are importable and skips it otherwise (this build image has no
// jaxlib: the file is compiled to an empty object here, and has NOT been run against XLA -- INTEGRATION.md section 3).
// Python side: jax.ffi.register_ffi_target("pbnn_sgld", jax.ffi.pycapsule(lib.PbnnSgld), platform="CUDA"), then
// jax.ffi.ffi_call("pbnn_sgld", (theta_sds, trace_sds))(X, y, keys, theta0, steps, workspace, batch_size=..., ...).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define PBNN_HAVE_XLA_FFI 1
#endif
#endif

#if defined(PBNN_HAVE_XLA_FFI)
#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstring>

#include "../../../include/pbnn_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static pbnn_mlp_spec MakeSpec(ffi::Span<const int32_t> widths, int32_t activation, float noise_level, float prior_scale) {
  pbnn_mlp_spec spec;
  std::memset(&spec, 0, sizeof(spec));
  spec.n_layers = static_cast<int32_t>(widths.size()) - 1;
  for (size_t i = 0; i < widths.size() && i < 9; ++i) spec.dims[i] = widths[i];
  spec.activation = activation;
  spec.noise_level = noise_level;
  spec.prior_scale = prior_scale;
  return spec;
}

static ffi::Error Status(int rc) { return rc ? ffi::Error::Internal(pbnn_last_error()) : ffi::Error::Success(); }

// pbnn.mcmc.sgld / scheduled_sgld (src/pbnn/mcmc/langevin.py:21-78, :143-204) for C chains
static ffi::Error SgldImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> X, ffi::Buffer<ffi::F32> y,
                           ffi::Buffer<ffi::U32> keys, ffi::Buffer<ffi::F32> theta0, ffi::Buffer<ffi::F32> steps,
                           ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F32> theta,
                           ffi::ResultBuffer<ffi::F32> trace, ffi::ResultBuffer<ffi::U32> flags, int32_t batch_size,
                           int32_t num_iterations, int32_t thin, int32_t burnin, int32_t activation, float noise_level,
                           float prior_scale, ffi::Span<const int32_t> widths) {
  const pbnn_mlp_spec spec = MakeSpec(widths, activation, noise_level, prior_scale);
  const int N = static_cast<int>(X.dimensions()[0]), C = static_cast<int>(keys.dimensions()[0]);
  cudaMemcpyAsync(theta->typed_data(), theta0.typed_data(), theta0.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  return Status(pbnn_sgld_run(&spec, X.typed_data(), y.typed_data(), N, nullptr, 0, PBNN_MB_REFERENCE_FIXED, batch_size,
                              keys.typed_data(), theta->typed_data(), steps.typed_data(),
                              static_cast<int>(steps.element_count()), 0, num_iterations, C, thin, burnin,
                              trace->typed_data(), flags->typed_data(), nullptr, nullptr, 1.0f,
                              workspace.size_bytes() ? workspace.typed_data() : nullptr,
                              static_cast<int64_t>(workspace.size_bytes()), stream));
}

// pbnn.mcmc.hmc (src/pbnn/mcmc/hamiltonian.py:18-77) for C chains
static ffi::Error HmcImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> X, ffi::Buffer<ffi::F32> y,
                          ffi::Buffer<ffi::U32> keys, ffi::Buffer<ffi::F32> theta0, ffi::Buffer<ffi::F32> inv_mass,
                          ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F32> theta,
                          ffi::ResultBuffer<ffi::F32> trace, ffi::ResultBuffer<ffi::F32> logp,
                          ffi::ResultBuffer<ffi::S32> accept_count, ffi::ResultBuffer<ffi::U32> flags, float step_size,
                          int32_t num_integration_steps, int32_t num_samples, int32_t thin, int32_t burnin,
                          int32_t activation, float noise_level, float prior_scale, ffi::Span<const int32_t> widths) {
  const pbnn_mlp_spec spec = MakeSpec(widths, activation, noise_level, prior_scale);
  const int N = static_cast<int>(X.dimensions()[0]), C = static_cast<int>(keys.dimensions()[0]);
  cudaMemcpyAsync(theta->typed_data(), theta0.typed_data(), theta0.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  return Status(pbnn_hmc_run(&spec, X.typed_data(), y.typed_data(), N, keys.typed_data(), theta->typed_data(),
                             inv_mass.typed_data(), step_size, num_integration_steps, 0, num_samples, C, thin, burnin,
                             trace->typed_data(), logp->typed_data(), accept_count->typed_data(), nullptr, nullptr,
                             flags->typed_data(), workspace.size_bytes() ? workspace.typed_data() : nullptr,
                             static_cast<int64_t>(workspace.size_bytes()), stream));
}

// predict_fn + moments (langevin.py:75-76; metrics_prediction_intervals.py:188)
static ffi::Error PredictImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> samples, ffi::Buffer<ffi::F32> x_test,
                              ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F32> sum,
                              ffi::ResultBuffer<ffi::F32> sumsq, int32_t activation, float noise_level,
                              ffi::Span<const int32_t> widths) {
  const pbnn_mlp_spec spec = MakeSpec(widths, activation, noise_level, 1.0f);
  const int S = static_cast<int>(samples.dimensions()[0]), M = static_cast<int>(x_test.dimensions()[0]);
  return Status(pbnn_predict(&spec, samples.typed_data(), S, x_test.typed_data(), M, nullptr, sum->typed_data(),
                             sumsq->typed_data(), reinterpret_cast<float*>(workspace.typed_data()),
                             static_cast<int64_t>(workspace.size_bytes()), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(PbnnSgld, SgldImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Attr<int32_t>("batch_size")
                                  .Attr<int32_t>("num_iterations")
                                  .Attr<int32_t>("thin")
                                  .Attr<int32_t>("burnin")
                                  .Attr<int32_t>("activation")
                                  .Attr<float>("noise_level")
                                  .Attr<float>("prior_scale")
                                  .Attr<ffi::Span<const int32_t>>("widths"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(PbnnHmc, HmcImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Attr<float>("step_size")
                                  .Attr<int32_t>("num_integration_steps")
                                  .Attr<int32_t>("num_samples")
                                  .Attr<int32_t>("thin")
                                  .Attr<int32_t>("burnin")
                                  .Attr<int32_t>("activation")
                                  .Attr<float>("noise_level")
                                  .Attr<float>("prior_scale")
                                  .Attr<ffi::Span<const int32_t>>("widths"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(PbnnPredict, PredictImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("activation")
                                  .Attr<float>("noise_level")
                                  .Attr<ffi::Span<const int32_t>>("widths"));
#else
// no XLA FFI headers at build time: nothing to compile (see the header comment)
extern "C" int pbnn_xla_ffi_available(void) { return 0; }
#endif
