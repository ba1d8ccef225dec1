// User-supplied constitutive laws, compiled at run time.
//
// In the reference the energy density is USER code that JAX differentiates and XLA compiles (README.md:93;
// tests/test_sparse_tracer.py:103-115 is one such density).  The B200 counterpart: the user (or tatva_b200/lawgen.py, from
// a density written once on symbols) supplies the CUDA source of a `UserLaw` struct with the `Mat` interface of common.cuh
// (psi / first / second); it is compiled by NVRTC into the SAME fused kernel templates the built-in laws use
// (fused.cuh: k_fused<El, UserLaw, MODE>, k_hvp_lifted, k_hessian_diag), loaded with the driver API and launched on the
// caller's stream.  The header text is embedded in the library at build time (build/embedded_sources.inc), so nothing but
// libnvrtc and the driver is needed at run time; both are opened lazily with dlopen so that the library still loads —
// and everything else works — on a machine without them.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "embedded_sources.inc"  // kSrcCommon, kSrcFused

namespace tatva {
namespace {

struct Api {
  void* nvrtc = nullptr;
  void* cuda = nullptr;
  decltype(&nvrtcCreateProgram) CreateProgram;
  decltype(&nvrtcDestroyProgram) DestroyProgram;
  decltype(&nvrtcAddNameExpression) AddNameExpression;
  decltype(&nvrtcCompileProgram) CompileProgram;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize;
  decltype(&nvrtcGetProgramLog) GetProgramLog;
  decltype(&nvrtcGetLoweredName) GetLoweredName;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize;
  decltype(&nvrtcGetCUBIN) GetCUBIN;
  decltype(&cuModuleLoadData) ModuleLoadData;
  decltype(&cuModuleGetFunction) ModuleGetFunction;
  decltype(&cuModuleGetGlobal) ModuleGetGlobal;
  decltype(&cuLaunchKernel) LaunchKernel;
  decltype(&cuMemcpyHtoDAsync) MemcpyHtoDAsync;
  decltype(&cuFuncSetAttribute) FuncSetAttribute;
  bool ok = false;
  std::string why;
};

template <class F>
bool load(void* lib, const char* name, F& out, std::string& why) {
  out = reinterpret_cast<F>(dlsym(lib, name));
  if (!out) why = std::string("missing symbol ") + name;
  return out != nullptr;
}

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so"})
      if ((a.nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    for (const char* n : {"libcuda.so.1", "libcuda.so"})
      if ((a.cuda = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!a.nvrtc) { a.why = "libnvrtc.so.12 not found"; return; }
    if (!a.cuda) { a.why = "libcuda.so.1 not found"; return; }
    a.ok = load(a.nvrtc, "nvrtcCreateProgram", a.CreateProgram, a.why) && load(a.nvrtc, "nvrtcDestroyProgram", a.DestroyProgram, a.why) &&
           load(a.nvrtc, "nvrtcAddNameExpression", a.AddNameExpression, a.why) && load(a.nvrtc, "nvrtcCompileProgram", a.CompileProgram, a.why) &&
           load(a.nvrtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize, a.why) && load(a.nvrtc, "nvrtcGetProgramLog", a.GetProgramLog, a.why) &&
           load(a.nvrtc, "nvrtcGetLoweredName", a.GetLoweredName, a.why) && load(a.nvrtc, "nvrtcGetCUBINSize", a.GetCUBINSize, a.why) &&
           load(a.nvrtc, "nvrtcGetCUBIN", a.GetCUBIN, a.why) && load(a.cuda, "cuModuleLoadData", a.ModuleLoadData, a.why) &&
           load(a.cuda, "cuModuleGetFunction", a.ModuleGetFunction, a.why) && load(a.cuda, "cuModuleGetGlobal_v2", a.ModuleGetGlobal, a.why) &&
           load(a.cuda, "cuLaunchKernel", a.LaunchKernel, a.why) && load(a.cuda, "cuMemcpyHtoDAsync_v2", a.MemcpyHtoDAsync, a.why) &&
           load(a.cuda, "cuFuncSetAttribute", a.FuncSetAttribute, a.why);
  });
  return a;
}

enum { kEnergy = 0, kResidual = 1, kHvp = 2, kDiag = 3, kLifted = 4, kNumKernels = 5 };

struct Module {
  CUmodule mod = nullptr;
  CUfunction fn[kNumKernels] = {};
  CUdeviceptr rule = 0;  // the module's own c_rule (custom quadrature)
};

struct Law {
  std::string source;
  int dim, dpn, n_params, uses_values;
  std::map<std::tuple<int, int, int>, Module> modules;  // (element, custom rule, device)
};

std::mutex g_mu;
std::vector<Law> g_laws;
std::string g_log;

const char* element_name(int el) {
  switch (el) {
    case TATVA_TRI3: return "tatva::Tri3";
    case TATVA_TET4: return "tatva::Tet4";
    case TATVA_HEX8: return "tatva::Hex8";
    case TATVA_QUAD4: return "tatva::Quad4";
    case TATVA_TRI6: return "tatva::Tri6";
    case TATVA_QUAD8: return "tatva::Quad8";
    default: return nullptr;
  }
}

// compile (law, element) for the current device; g_mu is held
int build_module(Law& law, int element, int custom, int dev, Module** out) {
  auto key = std::make_tuple(element, custom, dev);
  auto it = law.modules.find(key);
  if (it != law.modules.end()) { *out = &it->second; return TATVA_OK; }
  Api& a = api();
  if (!a.ok) { g_log = "run-time compilation unavailable: " + a.why; return TATVA_E_UNSUPPORTED; }
  const char* base = element_name(element);
  if (!base) return TATVA_E_UNSUPPORTED;
  const std::string el = custom ? std::string("tatva::Custom<") + base + ">" : std::string(base);
  const std::string src = std::string("#include \"fused.cuh\"\n") + law.source + "\n";
  const char* headers[] = {kSrcCommon, kSrcFused};
  const char* names[] = {"common.cuh", "fused.cuh"};
  nvrtcProgram prog;
  if (a.CreateProgram(&prog, src.c_str(), "user_law.cu", 2, headers, names) != NVRTC_SUCCESS) return TATVA_E_INVALID;
  const std::string expr[kNumKernels] = {
      "tatva::k_fused<" + el + ", UserLaw, 0>", "tatva::k_fused<" + el + ", UserLaw, 1>", "tatva::k_fused<" + el + ", UserLaw, 2>",
      "tatva::k_hessian_diag<" + el + ", UserLaw>", "tatva::k_hvp_lifted<" + el + ", UserLaw>"};
  for (const auto& e : expr) a.AddNameExpression(prog, e.c_str());
  if (custom) a.AddNameExpression(prog, "&tatva::c_rule");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  // sm_100a on B200 (the arch-specific target, like the ahead-of-time kernels); otherwise the device's own sm_XY
  const std::string arch = "--gpu-architecture=sm_" + std::to_string(major * 10 + minor) + ((major == 10 && minor == 0) ? "a" : "");
  const char* opts[] = {arch.c_str(), "-std=c++17", "-lineinfo", "--fmad=true"};
  const nvrtcResult rc = a.CompileProgram(prog, 4, opts);
  size_t n = 0;
  a.GetProgramLogSize(prog, &n);
  g_log.assign(n, '\0');
  if (n) a.GetProgramLog(prog, &g_log[0]);
  if (rc != NVRTC_SUCCESS) { a.DestroyProgram(&prog); return TATVA_E_INVALID; }
  size_t sz = 0;
  a.GetCUBINSize(prog, &sz);
  std::vector<char> cubin(sz);
  a.GetCUBIN(prog, cubin.data());
  Module m;
  cudaFree(nullptr);  // make sure the runtime's primary context is current
  if (a.ModuleLoadData(&m.mod, cubin.data()) != CUDA_SUCCESS) { a.DestroyProgram(&prog); g_log += "\ncuModuleLoadData failed"; return TATVA_E_INVALID; }
  for (int k = 0; k < kNumKernels; ++k) {
    const char* low = nullptr;
    if (a.GetLoweredName(prog, expr[k].c_str(), &low) != NVRTC_SUCCESS || a.ModuleGetFunction(&m.fn[k], m.mod, low) != CUDA_SUCCESS) {
      a.DestroyProgram(&prog);
      g_log += "\nkernel lookup failed: " + expr[k];
      return TATVA_E_INVALID;
    }
  }
  if (custom) {
    size_t bytes = 0;
    const char* low = nullptr;
    if (a.GetLoweredName(prog, "&tatva::c_rule", &low) != NVRTC_SUCCESS || a.ModuleGetGlobal(&m.rule, &bytes, m.mod, low) != CUDA_SUCCESS || bytes != sizeof(QuadRule)) {
      a.DestroyProgram(&prog);
      g_log += "\nc_rule not found in the module";
      return TATVA_E_INVALID;
    }
  }
  a.DestroyProgram(&prog);
  *out = &(law.modules[key] = m);
  return TATVA_OK;
}

}  // namespace

bool is_user_law(int material) { return material >= TATVA_USER_LAW_BASE; }

int user_law_info(int material, int* dpn) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int id = material - TATVA_USER_LAW_BASE;
  if (id < 0 || id >= (int)g_laws.size()) return TATVA_E_INVALID;
  if (dpn) *dpn = g_laws[id].dpn;
  return TATVA_OK;
}

// what: 0 energy (out = scalar), 1 residual, 2 HVP, 3 Hessian diagonal, 4 lifted HVP (v / out reduced, `map` given).
// The caller has zeroed `out` where its entry point says so for kLifted; the others follow p->zero_output here.
int user_law_launch(const tatva_plan* p, int material, int what, const double* prm, int n_params, const double* u,
                    const double* v, const int32_t* map, double* out, cudaStream_t st) {
  Module* m = nullptr;
  int dpn = 0, np = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const int id = material - TATVA_USER_LAW_BASE;
    if (id < 0 || id >= (int)g_laws.size()) return TATVA_E_INVALID;
    Law& law = g_laws[id];
    if (law.dim != p->dim || n_params != law.n_params) return TATVA_E_INVALID;
    int dev = 0;
    TATVA_CUDA_TRY(cudaGetDevice(&dev));
    const int rc = build_module(law, p->element, p->custom, dev, &m);
    if (rc != TATVA_OK) return rc;
    dpn = law.dpn;
    np = law.n_params > 0 ? law.n_params : 1;
  }
  Api& a = api();
  if (p->custom && a.MemcpyHtoDAsync(m->rule, &p->rule, sizeof(QuadRule), (CUstream)st) != CUDA_SUCCESS) return TATVA_E_INVALID;
  std::vector<double> blob(np, 0.0);  // the UserLaw kernel parameter: struct { double prm[np]; }
  for (int k = 0; k < n_params; ++k) blob[k] = prm[k];
  const int grid = grid_for(p->n_elems);
  const size_t scatter_smem = (size_t)(kBlock / 32) * (32 * ((p->npe * dpn) | 1) + 16 * (p->npe | 1)) * sizeof(double);
  long long E = p->n_elems;
  const double* coords = p->coords;
  const int32_t* conn = p->conn;
  double* partials = p->scratch;
  double* none = nullptr;
  CUresult rc;
  if (what == kEnergy) {
    if (grid > p->scratch_len) return TATVA_E_INVALID;
    void* args[] = {&coords, &conn, &E, blob.data(), &u, &v, &none, &partials};
    rc = a.LaunchKernel(m->fn[kEnergy], grid, 1, 1, kBlock, 1, 1, 0, (CUstream)st, args, nullptr);
    if (rc != CUDA_SUCCESS) return TATVA_E_INVALID;
    return sum_partials(p->scratch, grid, out, st);
  }
  if (what == kResidual || what == kHvp || what == kDiag) {
    if (p->zero_output) TATVA_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * p->n_nodes * dpn, st));
    if (scatter_smem > 48 * 1024) a.FuncSetAttribute(m->fn[what], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)scatter_smem);
    if (what == kDiag) {
      void* args[] = {&coords, &conn, &E, blob.data(), &u, &out};
      rc = a.LaunchKernel(m->fn[kDiag], grid, 1, 1, kBlock, 1, 1, (unsigned)scatter_smem, (CUstream)st, args, nullptr);
    } else {
      void* args[] = {&coords, &conn, &E, blob.data(), &u, &v, &out, &none};
      rc = a.LaunchKernel(m->fn[what], grid, 1, 1, kBlock, 1, 1, (unsigned)scatter_smem, (CUstream)st, args, nullptr);
    }
    return rc == CUDA_SUCCESS ? TATVA_OK : TATVA_E_INVALID;
  }
  if (what == kLifted) {
    void* args[] = {&coords, &conn, &E, blob.data(), &u, &v, &map, &out};
    rc = a.LaunchKernel(m->fn[kLifted], grid, 1, 1, kBlock, 1, 1, 0, (CUstream)st, args, nullptr);
    return rc == CUDA_SUCCESS ? TATVA_OK : TATVA_E_INVALID;
  }
  return TATVA_E_INVALID;
}

}  // namespace tatva

extern "C" {

int tatva_law_register(const char* cuda_source, int dim, int dofs_per_node, int n_params, int uses_values, int* material_id) {
  if (!cuda_source || !material_id || dim < 2 || dim > 3 || dofs_per_node < 1 || dofs_per_node > 8 || n_params < 0 || n_params > 64) return TATVA_E_INVALID;
  std::lock_guard<std::mutex> lk(tatva::g_mu);
  tatva::Law law;
  law.source = cuda_source;
  law.dim = dim;
  law.dpn = dofs_per_node;
  law.n_params = n_params;
  law.uses_values = uses_values;
  tatva::g_laws.push_back(std::move(law));
  *material_id = TATVA_USER_LAW_BASE + (int)tatva::g_laws.size() - 1;
  return TATVA_OK;
}

int tatva_law_compile_log(char* buf, int len) {
  if (!buf || len <= 0) return TATVA_E_INVALID;
  std::lock_guard<std::mutex> lk(tatva::g_mu);
  const size_t n = tatva::g_log.size() < (size_t)len - 1 ? tatva::g_log.size() : (size_t)len - 1;
  memcpy(buf, tatva::g_log.data(), n);
  buf[n] = '\0';
  return TATVA_OK;
}

}  // extern "C"
