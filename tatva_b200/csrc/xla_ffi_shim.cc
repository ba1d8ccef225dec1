// XLA FFI handlers for the C ABI of include/tatva_b200.h — the binding the reference anticipates
// (jax.ffi.ffi_call, tatva tests/test_sparse_tracer.py:582-598; tatva/sparse/tracer.py:2073-2097).
//
// Compiled only where jaxlib's headers exist.  This image has neither jax nor xla/ffi/api/ffi.h (SURVEY.md
// §8(b) states the gap), so here the translation unit is empty; INTEGRATION.md §2 shows the Python side.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define TATVA_HAVE_XLA_FFI 1
#endif
#endif

#ifdef TATVA_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/tatva_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// Plans are views over the caller's coords / connectivity buffers plus scratch; XLA may invoke handlers from
// any executor thread and for several devices at once, so the cache is keyed by (device pointers, element)
// and guarded by a mutex.  Everything enqueued is stream-ordered on the handler's stream.
struct PlanCache {
  std::mutex mu;
  std::map<std::tuple<const void*, const void*, int>, tatva_plan_t*> plans;
  tatva_plan_t* get(int element, ffi::Buffer<ffi::F64>& coords, ffi::Buffer<ffi::S32>& conn, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple((const void*)coords.typed_data(), (const void*)conn.typed_data(), element);
    auto it = plans.find(key);
    if (it != plans.end()) return it->second;
    tatva_plan_t* p = nullptr;
    const auto cd = coords.dimensions();
    const auto ed = conn.dimensions();
    if (tatva_plan_create(&p, element, cd[0], ed[0], coords.typed_data(), conn.typed_data(), 0, stream) != 0) return nullptr;
    plans.emplace(key, p);
    return p;
  }
};
PlanCache& cache() {
  static PlanCache c;
  return c;
}

ffi::Error status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(tatva_error_string(rc)); }

ffi::Error EnergyImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> u,
                      ffi::ResultBuffer<ffi::F64> out, int32_t element, int32_t material, double mu, double lmbda) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  const double prm[2] = {mu, lmbda};
  return status(tatva_energy(plan, material, prm, 2, u.typed_data(), out->typed_data(), stream));
}
ffi::Error ResidualImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> u,
                        ffi::ResultBuffer<ffi::F64> out, int32_t element, int32_t material, double mu, double lmbda) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  const double prm[2] = {mu, lmbda};
  return status(tatva_residual(plan, material, prm, 2, u.typed_data(), out->typed_data(), stream));
}
ffi::Error HvpImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> u,
                   ffi::Buffer<ffi::F64> v, ffi::ResultBuffer<ffi::F64> out, int32_t element, int32_t material, double mu,
                   double lmbda) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  const double prm[2] = {mu, lmbda};
  return status(tatva_hvp(plan, material, prm, 2, u.typed_data(), v.typed_data(), out->typed_data(), stream));
}
ffi::Error CsrAssembleImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn,
                           ffi::Buffer<ffi::F64> u, ffi::Buffer<ffi::S32> indptr, ffi::Buffer<ffi::S32> elem_pos,
                           ffi::ResultBuffer<ffi::F64> data, int32_t element, int32_t material, double mu, double lmbda) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  const double prm[2] = {mu, lmbda};
  return status(tatva_csr_assemble(plan, material, prm, 2, u.typed_data(), indptr.typed_data(), elem_pos.typed_data(),
                                   (int64_t)data->element_count(), data->typed_data(), stream));
}

}  // namespace

#define TATVA_COMMON_ATTRS .Attr<int32_t>("element").Attr<int32_t>("material").Attr<double>("mu").Attr<double>("lmbda")

XLA_FFI_DEFINE_HANDLER_SYMBOL(tatva_energy_ffi, EnergyImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                      TATVA_COMMON_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(tatva_residual_ffi, ResidualImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                      TATVA_COMMON_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(tatva_hvp_ffi, HvpImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>() TATVA_COMMON_ATTRS);
XLA_FFI_DEFINE_HANDLER_SYMBOL(tatva_csr_assemble_ffi, CsrAssembleImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F64>>() TATVA_COMMON_ATTRS);
#endif  // TATVA_HAVE_XLA_FFI
