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

#include <cstdint>
#include <map>
#include <mutex>
#include <tuple>

#include "../../include/tatva_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// A plan is a VIEW over the caller's coords / connectivity buffers plus plan-owned scratch.  XLA re-allocates buffers
// between executions, so plans are cached by what does not change — (device, stream, element, n_nodes, n_elems) — and
// re-pointed at the call's buffers with tatva_plan_rebind (no allocation, no synchronisation).  The stream is part of the
// key: XLA may run handlers from several executor threads, and two calls that share a plan must be stream-ordered.
// At most kMaxPlans plans are kept; the least recently used one is destroyed when a new key arrives.
struct PlanCache {
  static constexpr size_t kMaxPlans = 32;
  using Key = std::tuple<int, const void*, int, int64_t, int64_t>;
  struct Entry {
    tatva_plan_t* plan;
    uint64_t last_use;
  };
  std::mutex mu;
  std::map<Key, Entry> plans;
  uint64_t tick = 0;
  tatva_plan_t* get(int element, ffi::Buffer<ffi::F64>& coords, ffi::Buffer<ffi::S32>& conn, cudaStream_t stream) {
    const auto cd = coords.dimensions();
    const auto ed = conn.dimensions();
    if (cd.size() != 2 || ed.size() != 2) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    const Key key{dev, (const void*)stream, element, (int64_t)cd[0], (int64_t)ed[0]};
    auto it = plans.find(key);
    if (it == plans.end()) {
      if (plans.size() >= kMaxPlans) {
        auto lru = plans.begin();
        for (auto j = plans.begin(); j != plans.end(); ++j)
          if (j->second.last_use < lru->second.last_use) lru = j;
        tatva_plan_destroy(lru->second.plan);
        plans.erase(lru);
      }
      tatva_plan_t* p = nullptr;
      if (tatva_plan_create(&p, element, cd[0], ed[0], coords.typed_data(), conn.typed_data(), 0, stream) != 0) return nullptr;
      it = plans.emplace(key, Entry{p, 0}).first;
    }
    it->second.last_use = ++tick;
    if (tatva_plan_rebind(it->second.plan, coords.typed_data(), conn.typed_data()) != 0) return nullptr;
    return it->second.plan;
  }
};
PlanCache& cache() {
  static PlanCache c;
  return c;
}

ffi::Error status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(tatva_error_string(rc)); }

// `params` is the law's parameter vector: (mu, lambda) for linear elasticity and neo-Hooke, (mu, lambda, Gc, ell, k) for
// the two-field phase-field law, prm[0..n) for a law registered with tatva_law_register (material >= 1000).
ffi::Error EnergyImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> u,
                      ffi::ResultBuffer<ffi::F64> out, int32_t element, int32_t material, ffi::Span<const double> params) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  return status(tatva_energy(plan, material, params.begin(), (int)params.size(), u.typed_data(), out->typed_data(), stream));
}
ffi::Error ResidualImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> u,
                        ffi::ResultBuffer<ffi::F64> out, int32_t element, int32_t material, ffi::Span<const double> params) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  return status(tatva_residual(plan, material, params.begin(), (int)params.size(), u.typed_data(), out->typed_data(), stream));
}
ffi::Error HvpImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn, ffi::Buffer<ffi::F64> u,
                   ffi::Buffer<ffi::F64> v, ffi::ResultBuffer<ffi::F64> out, int32_t element, int32_t material,
                   ffi::Span<const double> params) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  return status(tatva_hvp(plan, material, params.begin(), (int)params.size(), u.typed_data(), v.typed_data(), out->typed_data(), stream));
}
ffi::Error CsrAssembleImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> coords, ffi::Buffer<ffi::S32> conn,
                           ffi::Buffer<ffi::F64> u, ffi::Buffer<ffi::S32> indptr, ffi::Buffer<ffi::S32> elem_pos,
                           ffi::ResultBuffer<ffi::F64> data, int32_t element, int32_t material, ffi::Span<const double> params) {
  tatva_plan_t* plan = cache().get(element, coords, conn, stream);
  if (!plan) return ffi::Error::Internal("tatva_plan_create failed");
  return status(tatva_csr_assemble(plan, material, params.begin(), (int)params.size(), u.typed_data(), indptr.typed_data(),
                                   elem_pos.typed_data(), (int64_t)data->element_count(), data->typed_data(), stream));
}

}  // namespace

#define TATVA_COMMON_ATTRS .Attr<int32_t>("element").Attr<int32_t>("material").Attr<ffi::Span<const double>>("params")

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
