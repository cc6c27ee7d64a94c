// ni_b200.cu -- C-ABI implementation (include/ni_b200.h) of the B200 ODE-ensemble engine.
//
// Host side only: handle management, HBM allocation, host<->device copies, kernel launches,
// NVRTC specialisation of user right-hand sides.  All arithmetic of the hot path lives in the
// kernels of ni_core.cuh / ni_large.cuh.  There is no CPU fallback: every compute entry point
// fails with NI_ERR_CUDA when no CUDA device is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ni_b200.h"
#include "ni_host.h"
#include "ni_large_host.h"
#include "ni_rhs_builtin.cuh"

#define NI_API extern "C" __attribute__((visibility("default")))
#ifndef NI_MAX_PEERS
#define NI_MAX_PEERS 16  // destinations of ni_ensemble_pack_terminal_peers (GPUs of one NVSwitch domain)
#endif

extern const char ni_core_source[];   // ni_core.cuh as a string (generated: ni_core_src.cpp)
extern const char ni_math_source[];   // ni_math.cuh as a string
extern const char ni_large_source[];  // ni_large.cuh as a string
extern const char ni_rhs_builtin_source[];  // ni_rhs_builtin.cuh as a string

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(NI_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

NI_API int ni_version(void) { return 100; }
NI_API const char *ni_last_error(void) { return g_err.c_str(); }
NI_API int ni_device_count(int *count) {
  if (!count) return fail(NI_ERR_INVALID, "count is NULL");
  *count = 0;
  CU(cudaGetDeviceCount(count));
  return NI_OK;
}

// ------------------------------------------------------------------ modules
struct ni_module {
  int kind = NI_KIND_SMALL;
  int n = 0, P = 0, Q = 0, method = NI_RK45, advanced = 0, fast = 0, scipy = 0, S = 6;
  const void *k_init = nullptr;
  const void *k_run[2] = {nullptr, nullptr};  // [record]
  int block = NI_BLOCK;
  size_t smem = 0;
  NiLargePlan large{};       // kernel-class specific launch plan (CTA-per-trajectory)
  cudaLibrary_t lib = nullptr;  // NVRTC modules only
  std::string log;
  std::string rhs_src;          // definition of `NiUserRhs` (thread-per-trajectory) / `NiUserStencil` (CTA-per-
                                // trajectory): recompiled with a user predicate by ni_ensemble_ff2cond_user
  bool component = false;       // CTA-per-trajectory: component-wise (dense) right-hand side
};

static const char *builtin_functor(int rhs_id) {
  switch (rhs_id) {
    case NI_RHS_HARMONIC: return "NiRhsHarmonic";
    case NI_RHS_LORENZ63: return "NiRhsLorenz63";
    case NI_RHS_VANDERPOL: return "NiRhsVanDerPol";
    case NI_RHS_EXPONENTIAL: return "NiRhsExponential<1>";
    case NI_RHS_RICCATI: return "NiRhsRiccati";
    case NI_RHS_README_ADV: return "NiRhsReadmeAdvanced";
    case NI_RHS_BLOWUP: return "NiRhsBlowup";
    default: return nullptr;
  }
}

static const NiKernelSet *builtin_small(int rhs_id, int n) {
  switch (rhs_id) {
    case NI_RHS_HARMONIC: return n == 2 ? ni_kset_harmonic() : nullptr;
    case NI_RHS_LORENZ63: return n == 3 ? ni_kset_lorenz63() : nullptr;
    case NI_RHS_VANDERPOL: return n == 2 ? ni_kset_vanderpol() : nullptr;
    case NI_RHS_EXPONENTIAL: return n == 1 ? ni_kset_exponential1() : nullptr;
    case NI_RHS_RICCATI: return n == 1 ? ni_kset_riccati() : nullptr;
    case NI_RHS_README_ADV: return n == 2 ? ni_kset_readme_adv() : nullptr;
    case NI_RHS_BLOWUP: return n == 1 ? ni_kset_blowup() : nullptr;
    default: return nullptr;
  }
}

static int compile_stencil_module(const std::string &functor_src, int n, int P, int Q, int method, int mode,
                                  const char *nvrtc_options, ni_module **out, bool component = false,
                                  const std::string &cond_src = std::string());

NI_API int ni_module_builtin(int rhs_id, int n, int method, int mode, ni_module **out) {
  if (!out) return fail(NI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (method != NI_RK23 && method != NI_RK45) return fail(NI_ERR_INVALID, "method must be NI_RK23 or NI_RK45");
  if (n < 1) return fail(NI_ERR_INVALID, "n must be >= 1");
  if (mode & ~(NI_MODE_ADVANCED | NI_MODE_FAST | NI_MODE_SCIPY)) return fail(NI_ERR_INVALID, "unknown mode bits 0x%x", mode);
  const int advanced = (mode & NI_MODE_ADVANCED) ? 1 : 0, fast = (mode & NI_MODE_FAST) ? 1 : 0,
            scipy = (mode & NI_MODE_SCIPY) ? 1 : 0;
  ni_module *m = new ni_module;
  m->method = method;
  m->advanced = advanced;
  m->fast = fast;
  m->scipy = scipy;
  m->S = method == NI_RK45 ? 6 : 3;
  if (const NiKernelSet *ks = builtin_small(rhs_id, n)) {
    m->kind = NI_KIND_SMALL;
    m->n = ks->n;
    m->P = ks->P;
    m->Q = ks->Q;
    m->k_init = ks->init[method];
    m->k_run[0] = ks->run[method][scipy ? 2 : advanced][0][fast];
    m->k_run[1] = ks->run[method][scipy ? 2 : advanced][1][fast];
    m->rhs_src = std::string("#include \"ni_rhs_builtin.cuh\"\nusing NiUserRhs = ") + builtin_functor(rhs_id) + ";\n";
    *out = m;
    return NI_OK;
  }
  if ((fast || scipy) && (rhs_id == NI_RHS_HEAT1D || rhs_id == NI_RHS_BURGERS1D)) {
    // the ahead-of-time library carries the reference-semantics strict kernels; the opt-in variants of the
    // CTA-per-trajectory kernels are specialised with NVRTC from the same templates
    delete m;
    const int P = rhs_id == NI_RHS_HEAT1D ? 1 : 2;
    return compile_stencil_module(std::string("using NiUserStencil = ") + (rhs_id == NI_RHS_HEAT1D ? "NiStencilHeat" : "NiStencilBurgers") + ";\n",
                                  n, P, 0, method, mode, nullptr, out);
  }
  if (ni_large_builtin(rhs_id, n, method, advanced, &m->large, &m->P, &m->Q)) {
    m->kind = NI_KIND_LARGE;
    m->n = n;
    m->k_init = m->large.k_init;
    m->k_run[0] = m->large.k_run[0];
    m->k_run[1] = m->large.k_run[1];
    m->block = m->large.threads;
    m->smem = m->large.smem_bytes;
    m->rhs_src = std::string("using NiUserStencil = ") + (rhs_id == NI_RHS_HEAT1D ? "NiStencilHeat" : "NiStencilBurgers") + ";\n";
    *out = m;
    return NI_OK;
  }
  delete m;
  return fail(NI_ERR_UNSUPPORTED, "no built-in kernel for rhs_id=%d with n=%d", rhs_id, n);
}

// ---- NVRTC (loaded lazily so that the library also loads on machines without it)
typedef struct _nvrtcProgram *nvrtcProgram;
struct NvrtcApi {
  void *h = nullptr;
  int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
  int (*DestroyProgram)(nvrtcProgram *);
  int (*CompileProgram)(nvrtcProgram, int, const char *const *);
  int (*GetProgramLogSize)(nvrtcProgram, size_t *);
  int (*GetProgramLog)(nvrtcProgram, char *);
  int (*GetCUBINSize)(nvrtcProgram, size_t *);
  int (*GetCUBIN)(nvrtcProgram, char *);
  const char *(*GetErrorString)(int);
};
static NvrtcApi g_nvrtc;
static std::mutex g_nvrtc_mu;

static int load_nvrtc() {
  std::lock_guard<std::mutex> lk(g_nvrtc_mu);
  if (g_nvrtc.h) return NI_OK;
  const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                         "/usr/local/cuda/lib64/libnvrtc.so"};
  void *h = nullptr;
  if (const char *env = getenv("NI_B200_NVRTC")) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
  for (size_t i = 0; !h && i < sizeof(names) / sizeof(names[0]); ++i) h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(NI_ERR_NVRTC, "cannot load libnvrtc (%s); set NI_B200_NVRTC to its path", dlerror());
#define NI_SYM(field, name)                                                       \
  *(void **)(&g_nvrtc.field) = dlsym(h, name);                                    \
  if (!g_nvrtc.field) {                                                           \
    dlclose(h);                                                                   \
    return fail(NI_ERR_NVRTC, "libnvrtc lacks %s", name);                         \
  }
  NI_SYM(CreateProgram, "nvrtcCreateProgram")
  NI_SYM(DestroyProgram, "nvrtcDestroyProgram")
  NI_SYM(CompileProgram, "nvrtcCompileProgram")
  NI_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  NI_SYM(GetProgramLog, "nvrtcGetProgramLog")
  NI_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  NI_SYM(GetCUBIN, "nvrtcGetCUBIN")
  NI_SYM(GetErrorString, "nvrtcGetErrorString")
#undef NI_SYM
  g_nvrtc.h = h;
  return NI_OK;
}

// ---- on-disk cubin cache (opt-in: NI_B200_CACHE_DIR): key = FNV-1a of source + headers + options + library version
static uint64_t fnv1a(uint64_t h, const char *p, size_t n) {
  for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
  return h;
}
static std::string cache_path(const std::string &src, const std::string &opts) {
  const char *dir = getenv("NI_B200_CACHE_DIR");
  if (!dir || !*dir) return std::string();
  uint64_t h = 1469598103934665603ull;
  h = fnv1a(h, src.data(), src.size());
  h = fnv1a(h, opts.data(), opts.size());
  for (const char *hdr : {ni_core_source, ni_math_source, ni_large_source, ni_rhs_builtin_source}) h = fnv1a(h, hdr, strlen(hdr));
  char name[64];
  snprintf(name, sizeof name, "/ni_b200_v%d_sm100a_%016llx.cubin", ni_version(), (unsigned long long)h);
  return std::string(dir) + name;
}
static bool read_file(const std::string &path, std::vector<char> *out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t)n : 0);
  const bool ok = n > 0 && fread(out->data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  return ok;
}
static void write_file_atomic(const std::string &path, const std::vector<char> &data) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (ok)
    rename(tmp.c_str(), path.c_str());
  else
    remove(tmp.c_str());
}

static int load_cubin(const std::vector<char> &cubin, const char *const names[3], ni_module *m) {
  cudaError_t ce = cudaLibraryLoadData(&m->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  cudaKernel_t k[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < 3 && ce == cudaSuccess; ++i) ce = cudaLibraryGetKernel(&k[i], m->lib, names[i]);
  if (ce != cudaSuccess) {
    int code = fail(NI_ERR_CUDA, "loading the compiled RHS module: %s", cudaGetErrorString(ce));
    if (m->lib) cudaLibraryUnload(m->lib);
    m->lib = nullptr;
    return code;
  }
  m->k_init = (const void *)k[0];
  m->k_run[0] = (const void *)k[1];
  m->k_run[1] = (const void *)k[2];
  return NI_OK;
}

// Compiles `src` (which includes the embedded kernel headers) for sm_100a and loads the three entry points.
static int compile_and_load(const std::string &src, const char *nvrtc_options, const char *const names[3], ni_module *m) {
  const std::string cpath = cache_path(src, nvrtc_options ? nvrtc_options : "");
  if (!cpath.empty()) {
    std::vector<char> cached;
    if (read_file(cpath, &cached) && load_cubin(cached, names, m) == NI_OK) {
      m->log = "cubin cache hit: " + cpath;
      return NI_OK;
    }
  }
  if (int rc = load_nvrtc()) return rc;
  nvrtcProgram prog = nullptr;
  const char *hdr_src[] = {ni_core_source, ni_math_source, ni_large_source, ni_rhs_builtin_source};
  const char *hdr_name[] = {"ni_core.cuh", "ni_math.cuh", "ni_large.cuh", "ni_rhs_builtin.cuh"};
  int rc = g_nvrtc.CreateProgram(&prog, src.c_str(), "ni_user_rhs.cu", 4, hdr_src, hdr_name);
  if (rc) return fail(NI_ERR_NVRTC, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(rc));
  std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
  if (nvrtc_options && *nvrtc_options) {  // space separated extra options
    std::string o(nvrtc_options);
    size_t pos = 0;
    while (pos < o.size()) {
      size_t e = o.find(' ', pos);
      if (e == std::string::npos) e = o.size();
      if (e > pos) opts.push_back(o.substr(pos, e - pos));
      pos = e + 1;
    }
  }
  std::vector<const char *> copts;
  for (auto &s : opts) copts.push_back(s.c_str());
  rc = g_nvrtc.CompileProgram(prog, (int)copts.size(), copts.data());
  size_t lsz = 0;
  if (g_nvrtc.GetProgramLogSize(prog, &lsz) == 0 && lsz > 1) {
    m->log.resize(lsz);
    g_nvrtc.GetProgramLog(prog, &m->log[0]);
  }
  if (rc) {
    int code = fail(NI_ERR_NVRTC, "nvrtcCompileProgram: %s\n%s", g_nvrtc.GetErrorString(rc), m->log.c_str());
    g_nvrtc.DestroyProgram(&prog);
    return code;
  }
  size_t csz = 0;
  g_nvrtc.GetCUBINSize(prog, &csz);
  std::vector<char> cubin(csz);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  g_nvrtc.DestroyProgram(&prog);
  if (int rc = load_cubin(cubin, names, m)) return rc;
  if (!cpath.empty()) write_file_atomic(cpath, cubin);
  return NI_OK;
}

NI_API int ni_module_compile(const char *rhs_body, int n, int P, int Q, int method, int mode,
                             const char *nvrtc_options, ni_module **out) {
  if (!out) return fail(NI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!rhs_body) return fail(NI_ERR_INVALID, "rhs_body is NULL");
  if (method != NI_RK23 && method != NI_RK45) return fail(NI_ERR_INVALID, "method must be NI_RK23 or NI_RK45");
  if (n < 1 || P < 0 || Q < 0) return fail(NI_ERR_INVALID, "need n >= 1, P >= 0, Q >= 0");
  if (n > NI_MAX_SMALL_N)
    return fail(NI_ERR_UNSUPPORTED, "dense user RHS snippets support n <= %d (got %d); larger systems must be "
                "3-point stencils (ni_module_compile_stencil)", NI_MAX_SMALL_N, n);
  if (mode & ~(NI_MODE_ADVANCED | NI_MODE_FAST | NI_MODE_SCIPY)) return fail(NI_ERR_INVALID, "unknown mode bits 0x%x", mode);
  const int advanced = (mode & NI_MODE_ADVANCED) ? 1 : 0, fast = (mode & NI_MODE_FAST) ? 1 : 0,
            scipy = (mode & NI_MODE_SCIPY) ? 1 : 0;
  std::string rhs_src;
  rhs_src += "struct NiUserRhs {\n  static constexpr int N = " + std::to_string(n) + ", P = " + std::to_string(P) +
             ", Q = " + std::to_string(Q) + ";\n";
  rhs_src += "  template <bool FAST>\n";
  rhs_src += "  NI_DEV static void eval(double t, const double* y, const double* p, double* dy, double* aux) {\n";
  rhs_src += "    (void)t; (void)y; (void)p; (void)dy; (void)aux;\n";
  rhs_src += rhs_body;
  rhs_src += "\n  }\n};\n";
  std::string src = "#include \"ni_core.cuh\"\n" + rhs_src;
  const std::string M = method == NI_RK45 ? "NI_RK45" : "NI_RK23", A = scipy ? "2" : advanced ? "1" : "0",
                    F = fast ? "true" : "false";
  src += "extern \"C\" __global__ void __launch_bounds__(NI_BLOCK) ni_user_init(const NiArgs a) { "
         "ni_small_init_body<NiUserRhs, " + M + ">(a); }\n";
  src += "extern \"C\" __global__ void __launch_bounds__(NI_BLOCK, NI_MIN_BLOCKS(NiUserRhs::N)) ni_user_run(const NiArgs a) { "
         "ni_small_run_body<NiUserRhs, " + M + ", " + A + ", false, " + F + ">(a); }\n";
  src += "extern \"C\" __global__ void __launch_bounds__(NI_BLOCK, NI_MIN_BLOCKS(NiUserRhs::N)) ni_user_run_rec(const NiArgs a) { "
         "ni_small_run_body<NiUserRhs, " + M + ", " + A + ", true, " + F + ">(a); }\n";
  ni_module *m = new ni_module;
  const char *names[3] = {"ni_user_init", "ni_user_run", "ni_user_run_rec"};
  // strict: the snippet's own a*b+c must not be contracted either; fast: let NVRTC contract it
  const std::string all_opts = std::string(fast ? "--fmad=true " : "--fmad=false ") + (nvrtc_options ? nvrtc_options : "");
  if (int rc = compile_and_load(src, all_opts.c_str(), names, m)) {
    delete m;
    return rc;
  }
  m->fast = fast;
  m->scipy = scipy;
  m->rhs_src = rhs_src;
  m->kind = NI_KIND_SMALL;
  m->n = n;
  m->P = P;
  m->Q = Q;
  m->method = method;
  m->advanced = advanced;
  m->S = method == NI_RK45 ? 6 : 3;
  *out = m;
  return NI_OK;
}

// NVRTC build of the CTA-per-trajectory kernels for a functor `NiUserStencil` defined by `functor_src`; with
// `cond_src` (the definition of ni_user_cond) the run kernels carry that ff2cond predicate
static int compile_stencil_module(const std::string &functor_src, int n, int P, int Q, int method, int mode,
                                  const char *nvrtc_options, ni_module **out, bool component, const std::string &cond_src) {
  if (method != NI_RK23 && method != NI_RK45) return fail(NI_ERR_INVALID, "method must be NI_RK23 or NI_RK45");
  if (P < 0 || Q < 0) return fail(NI_ERR_INVALID, "need P >= 0, Q >= 0");
  if (mode & ~(NI_MODE_ADVANCED | NI_MODE_FAST | NI_MODE_SCIPY)) return fail(NI_ERR_INVALID, "unknown mode bits 0x%x", mode);
  int lg = 0, threads = 0;
  size_t smem = 0;
  if (!ni_large_dims(n, method, &lg, &threads, &smem, component))
    return fail(NI_ERR_UNSUPPORTED, "CTA-per-trajectory systems support 1 <= n <= %d (got %d)", NI_LARGE_MAX_N, n);
  const int advanced = (mode & NI_MODE_ADVANCED) ? 1 : 0, fast = (mode & NI_MODE_FAST) ? 1 : 0,
            scipy = (mode & NI_MODE_SCIPY) ? 1 : 0;
  const std::string M = method == NI_RK45 ? "NI_RK45" : "NI_RK23", A = scipy ? "2" : advanced ? "1" : "0",
                    F = fast ? "true" : "false", C = std::to_string(1 << lg);
  std::string src = std::string(cond_src.empty() ? "" : "#define NI_HAVE_USER_COND 1\n") + "#include \"ni_large.cuh\"\n" +
                    functor_src + cond_src;
  const std::string lb = "extern \"C\" __global__ void __launch_bounds__(NI_LARGE_MAX_THREADS, 1) ";
  src += lb + "ni_user_init(const NiArgs a) { extern __shared__ double s[]; ni_large_init_body<NiUserStencil, " + M +
         ", " + C + ">(a, s); }\n";
  src += lb + "ni_user_run(const NiArgs a) { extern __shared__ double s[]; ni_large_run_body<NiUserStencil, " + M +
         ", " + A + ", false, " + C + ", " + F + ">(a, s); }\n";
  src += lb + "ni_user_run_rec(const NiArgs a) { extern __shared__ double s[]; ni_large_run_body<NiUserStencil, " + M +
         ", " + A + ", true, " + C + ", " + F + ">(a, s); }\n";
  ni_module *m = new ni_module;
  const char *names[3] = {"ni_user_init", "ni_user_run", "ni_user_run_rec"};
  const std::string all_opts = std::string(fast ? "--fmad=true " : "--fmad=false ") + (nvrtc_options ? nvrtc_options : "");
  if (int rc = compile_and_load(src, all_opts.c_str(), names, m)) {
    delete m;
    return rc;
  }
  m->kind = NI_KIND_LARGE;
  m->n = n;
  m->P = P;
  m->Q = Q;
  m->method = method;
  m->advanced = advanced;
  m->fast = fast;
  m->scipy = scipy;
  m->S = method == NI_RK45 ? 6 : 3;
  m->block = threads;
  m->smem = smem;
  m->rhs_src = functor_src;
  m->component = component;
  m->large.cpt = 1 << lg;
  m->large.threads = threads;
  m->large.smem_bytes = smem;
  m->large.k_init = m->k_init;
  m->large.k_run[0] = m->k_run[0];
  m->large.k_run[1] = m->k_run[1];
  *out = m;
  return NI_OK;
}

// Source of `NiUserStencil` for a stencil (component = 0) or component-wise (1) right-hand side body, with an optional
// auxiliary function body over the whole state vector
static std::string large_functor_src(int component, const char *rhs_body, const char *aux_body, int P, int Q) {
  std::string f;
  f += "struct NiUserStencil {\n  static constexpr int P = " + std::to_string(P < 0 ? 0 : P) + ", Q = " + std::to_string(Q) + ";\n";
  f += std::string("  static constexpr bool COMPONENT = ") + (component ? "true" : "false") + ";\n  template <bool FAST>\n";
  if (component) {
    f += "  NI_DEV static double eval(double t, const double* y, int j, int n, const double* p) {\n";
    f += "    (void)t; (void)y; (void)j; (void)n; (void)p;\n";
  } else {
    f += "  NI_DEV static double eval(double t, double ym, double yc, double yp, const double* p, int j, int n) {\n";
    f += "    (void)t; (void)ym; (void)yc; (void)yp; (void)p; (void)j; (void)n;\n";
  }
  f += rhs_body;
  f += "\n  }\n";
  if (Q > 0) {
    f += "  NI_DEV static void aux(double t, const double* y, int n, const double* p, double* aux) {\n";
    f += "    (void)t; (void)y; (void)n; (void)p; (void)aux;\n";
    f += aux_body;
    f += "\n  }\n";
  }
  f += "};\n";
  return f;
}

NI_API int ni_module_compile_large(int component, const char *rhs_body, const char *aux_body, int n, int P, int Q,
                                   int method, int mode, const char *nvrtc_options, ni_module **out) {
  if (!out) return fail(NI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!rhs_body) return fail(NI_ERR_INVALID, "rhs_body is NULL");
  if (Q < 0 || (Q > 0 && !aux_body)) return fail(NI_ERR_INVALID, "Q auxiliary outputs need an aux_body");
  return compile_stencil_module(large_functor_src(component, rhs_body, aux_body, P, Q), n, P, Q, method, mode, nvrtc_options,
                                out, component != 0);
}

NI_API int ni_module_compile_stencil(const char *stencil_body, int n, int P, int method, int mode,
                                     const char *nvrtc_options, ni_module **out) {
  return ni_module_compile_large(0, stencil_body, nullptr, n, P, 0, method, mode, nvrtc_options, out);
}

NI_API int ni_module_compile_component(const char *component_body, int n, int P, int method, int mode,
                                       const char *nvrtc_options, ni_module **out) {
  return ni_module_compile_large(1, component_body, nullptr, n, P, 0, method, mode, nvrtc_options, out);
}

NI_API int ni_module_info(const ni_module *m, int *n, int *P, int *Q, int *method, int *advanced, int *n_stages) {
  if (!m) return fail(NI_ERR_INVALID, "module is NULL");
  if (n) *n = m->n;
  if (P) *P = m->P;
  if (Q) *Q = m->Q;
  if (method) *method = m->method;
  if (advanced) *advanced = m->advanced | (m->fast ? NI_MODE_FAST : 0) | (m->scipy ? NI_MODE_SCIPY : 0);
  if (n_stages) *n_stages = m->S;
  return NI_OK;
}
NI_API int ni_module_layout(const ni_module *m, int *layout) {
  if (!m || !layout) return fail(NI_ERR_INVALID, "module / layout is NULL");
  *layout = m->kind;
  return NI_OK;
}
NI_API const char *ni_module_log(const ni_module *m) { return m ? m->log.c_str() : ""; }
NI_API int ni_module_destroy(ni_module *m) {
  if (!m) return NI_OK;
  if (m->lib) cudaLibraryUnload(m->lib);
  delete m;
  return NI_OK;
}

// ------------------------------------------------------------------ ensembles
// Pinned host slots for the 64-byte queue read-back of every launch.  cudaHostAlloc costs milliseconds, so one
// 32 KB slab per process is cut into 512 slots; an ensemble that finds none free reads back through pageable memory.
namespace {
struct PinnedSlots {
  std::mutex mu;
  unsigned long long *slab = nullptr;
  bool tried = false;
  std::vector<int> free_list;
  unsigned long long *get() {
    std::lock_guard<std::mutex> g(mu);
    if (!tried) {
      tried = true;
      if (cudaHostAlloc((void **)&slab, 512 * 64, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        slab = nullptr;
      } else {
        for (int i = 511; i >= 0; --i) free_list.push_back(i);
      }
    }
    if (!slab || free_list.empty()) return nullptr;
    const int i = free_list.back();
    free_list.pop_back();
    return slab + i * 8;
  }
  void put(unsigned long long *p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(mu);
    free_list.push_back((int)((p - slab) / 8));
  }
};
PinnedSlots g_pinned;
}  // namespace

#define NI_PASSES 3
struct ni_ensemble {
  ni_module mod;  // copy of the module description (kernels stay valid while the module lives)
  int device = 0;
  int64_t N = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sms = 148, blocks_per_sm = 1;
  NiArgs a{};
  // owned device allocations
  std::vector<void *> allocs;
  double *y0_rows = nullptr, *t0_d = nullptr, *tb0_d = nullptr, *params_rows = nullptr;
  double *tb_saved = nullptr;
  double *max_step_d = nullptr, *first_step_d = nullptr;  // per-trajectory step limits (allocated on first use)
  std::vector<double> tol_host;  // CTA-per-trajectory kernels: the rtol | atol vectors last copied to tol_d
  unsigned char *flag_d = nullptr;
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  // guard mode (environment NI_B200_GUARD=1 when the ensemble is created; compute-sanitizer is closed on the GPU pool):
  // every device allocation of the ensemble sits between two 64 KiB guard zones filled with 0xA5;
  // ni_ensemble_guard_check counts the guard bytes a kernel has overwritten
  // launch order (ni_ensemble_set_order): the work queue of a run hands out order_d[0], order_d[1], ... instead of
  // 0, 1, ...; NULL = identity
  unsigned int *order_d = nullptr;
  unsigned long long *order_n_d = nullptr;
  bool guard = false;
  std::vector<std::pair<void *, size_t>> guarded;  // (user pointer, user bytes) of every live allocation
  bool uploaded = false, initialised = false;
  int64_t max_attempts = 1ll << 40;
  // tail compaction (thread-per-trajectory kernels, fast-forward launches of ensembles that fill the GPU): the run is
  // NI_PASSES launches; pass k works on the trajectories pass k-1 parked (park_list[k-1], count = queue_all[8(k-1)+4])
  unsigned long long *queue_all = nullptr;  // NI_PASSES x 8 counters; a.queue points at the current pass
  unsigned int *park_list[2] = {nullptr, nullptr};
  int64_t park_cap = 0;
  int park_below = 0;                       // 0: no compaction (the default, see ni_ensemble_compaction); NI_B200_PARK_BELOW
  int last_passes = 1;                      // launches of the most recent run
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t n_launches = 0;
  bool have_init_t = false, have_run_t = false;
  struct CondKernels {
    cudaLibrary_t lib = nullptr;
    const void *k_run[2] = {nullptr, nullptr};
  };
  // lock-step ni.step(): {clear the queue, run kernel with max_accepts = 1, queue -> pinned host} as one CUDA graph
  // from the 65th call on (the call is launch-bound: 37 us per step with direct launches, 25 us with the graph).
  // Re-captured whenever the kernel arguments, the kernel or the stream differ from the captured ones.
  unsigned long long *q_host = nullptr;  // pinned slot (8 entries) from g_pinned, or nullptr
  unsigned long long q_pageable[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaGraphExec_t step_graph = nullptr;
  NiArgs step_args{};
  const void *step_kernel = nullptr;
  cudaStream_t step_stream = nullptr;
  int64_t step_calls = 0;  // capture + instantiation cost ~1.5 ms: only worth it for loops longer than ~100 steps
};

#define NI_GUARD_BYTES ((size_t)65536)
#define NI_GUARD_FILL 0xA5
// cudaMalloc / cudaFree of the ensemble's device buffers; in guard mode with a guard zone on either side
static int gmalloc(ni_ensemble *e, void **out, size_t bytes) {
  void *q = nullptr;
  const size_t user = (bytes + 255) / 256 * 256;  // (the trailing guard starts on a 256-byte boundary)
  cudaError_t ce = cudaMalloc(&q, e->guard ? user + 2 * NI_GUARD_BYTES : bytes);
  if (ce != cudaSuccess) return fail(ce == cudaErrorMemoryAllocation ? NI_ERR_NOMEM : NI_ERR_CUDA, "cudaMalloc(%zu bytes): %s",
                                     bytes, cudaGetErrorString(ce));
  if (e->guard) {
    if (cudaMemset(q, NI_GUARD_FILL, NI_GUARD_BYTES) != cudaSuccess ||
        cudaMemset((char *)q + NI_GUARD_BYTES + user, NI_GUARD_FILL, NI_GUARD_BYTES) != cudaSuccess) {
      cudaFree(q);
      return fail(NI_ERR_CUDA, "cudaMemset of the guard zones failed");
    }
    q = (char *)q + NI_GUARD_BYTES;
    e->guarded.emplace_back(q, user);
  }
  *out = q;
  return NI_OK;
}
static void gfree(ni_ensemble *e, void *q) {
  if (!q) return;
  if (e->guard) {
    for (size_t k = 0; k < e->guarded.size(); ++k)
      if (e->guarded[k].first == q) {
        e->guarded.erase(e->guarded.begin() + (long)k);
        break;
      }
    q = (char *)q - NI_GUARD_BYTES;
  }
  cudaFree(q);
}

template <class T>
static int dalloc(ni_ensemble *e, T **p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  void *q = nullptr;
  if (int rc = gmalloc(e, &q, count * sizeof(T))) return rc;
  e->allocs.push_back(q);
  *p = (T *)q;
  return NI_OK;
}

static int need_scratch(ni_ensemble *e, size_t bytes) {
  if (bytes <= e->scratch_bytes) return NI_OK;
  void *q = nullptr;
  if (int rc = gmalloc(e, &q, bytes)) return rc;
  if (e->scratch) {
    cudaStreamSynchronize(e->stream);
    gfree(e, e->scratch);
  }
  e->scratch = q;
  e->scratch_bytes = bytes;
  return NI_OK;
}

NI_API int ni_ensemble_destroy(ni_ensemble *e) {
  if (!e) return NI_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void *p : e->allocs) gfree(e, p);
  if (e->scratch) gfree(e, e->scratch);
  if (e->step_graph) cudaGraphExecDestroy(e->step_graph);
  g_pinned.put(e->q_host);
  for (auto &ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return NI_OK;
}

NI_API int ni_ensemble_create(const ni_module *m, int device, int64_t N, ni_ensemble **out) {
  if (!out) return fail(NI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!m) return fail(NI_ERR_INVALID, "module is NULL");
  if (N < 1) return fail(NI_ERR_INVALID, "N must be >= 1");
  if (N > 0xffffffffll) return fail(NI_ERR_INVALID, "N must fit 32 bits per GPU");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(NI_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  ni_ensemble *e = new ni_ensemble;
  if (const char *g = getenv("NI_B200_GUARD")) e->guard = atoi(g) != 0;
  e->mod = *m;
  e->mod.lib = nullptr;  // not owned
  e->device = device;
  e->N = N;
  const int n = m->n, P = m->P, Q = m->Q;
  int rc = NI_OK;
  cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) {
    delete e;
    return fail(NI_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(ce));
  }
  e->own_stream = true;
  for (auto &ev : e->ev) cudaEventCreate(&ev);
  cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
  NiArgs &a = e->a;
  a.N = N;
  const size_t sN = (size_t)N;
#define NI_TRY(x)                \
  if (rc == NI_OK) rc = (x);
  NI_TRY(dalloc(e, &a.t, sN))
  NI_TRY(dalloc(e, &a.h_abs, sN))
  NI_TRY(dalloc(e, &a.step_size, sN))
  NI_TRY(dalloc(e, &a.y, sN * n))
  NI_TRY(dalloc(e, &a.klast, sN * n))
  a.y_rows = nullptr;
  if (m->kind == NI_KIND_SMALL && n > 1) NI_TRY(dalloc(e, &a.y_rows, sN * n))  // row-major mirror of y for the getters
  NI_TRY(dalloc(e, &a.aux, sN * Q))
  NI_TRY(dalloc(e, &a.params, sN * P))
  NI_TRY(dalloc(e, &a.t_bound, sN))
  NI_TRY(dalloc(e, &a.direction, sN))
  NI_TRY(dalloc(e, &a.status, sN))
  NI_TRY(dalloc(e, &a.n_acc, sN))
  NI_TRY(dalloc(e, &a.n_rej, sN))
  NI_TRY(dalloc(e, &a.running, sN))
  NI_TRY(dalloc(e, &a.cond_hit, sN))
  NI_TRY(dalloc(e, &a.step_rejected, sN))
  NI_TRY(dalloc(e, &e->queue_all, 8 * NI_PASSES))
  a.queue = e->queue_all;
  NI_TRY(dalloc(e, &a.rec_cursor, 1))
  NI_TRY(dalloc(e, &e->y0_rows, sN * n))
  NI_TRY(dalloc(e, &e->t0_d, sN))
  NI_TRY(dalloc(e, &e->tb0_d, sN))
  NI_TRY(dalloc(e, &e->params_rows, sN * P))
  NI_TRY(dalloc(e, &e->tb_saved, sN))
  NI_TRY(dalloc(e, &e->flag_d, sN))
  if (rc == NI_OK && m->kind == NI_KIND_LARGE) {
    rc = dalloc(e, &e->mod.large.tol_d, 2 * (size_t)n);
    a.n = n;
    a.rtol_d = e->mod.large.tol_d;
    a.atol_d = e->mod.large.tol_d + n;
    for (const void *k : {m->k_init, m->k_run[0], m->k_run[1]}) {  // opt in to > 48 KB of dynamic shared memory
      cudaError_t ce2 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem);
      if (rc == NI_OK && ce2 != cudaSuccess)
        rc = fail(NI_ERR_CUDA, "cudaFuncSetAttribute(%zu B shared memory): %s", m->smem, cudaGetErrorString(ce2));
    }
  }
#undef NI_TRY
  if (rc != NI_OK) {
    std::string keep = g_err;
    ni_ensemble_destroy(e);
    g_err = keep;
    return rc;
  }
  cudaMemsetAsync(a.rec_cursor, 0, sizeof(unsigned long long), e->stream);
  cudaMemsetAsync(a.cond_hit, 0, sN, e->stream);
  a.rec_cap = 0;
  a.compat = m->scipy;
  a.max_step = __builtin_inf();
  a.first_step = 0.0;
  a.y0_rows = e->y0_rows;
  a.t0 = e->t0_d;
  a.tb0 = e->tb0_d;
  a.params_rows = e->params_rows;
  // persistent grid: resident blocks per SM x SMs
  int nb = 0;
  ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, m->k_run[0], m->block, m->smem);
  if (ce != cudaSuccess || nb < 1) {
    cudaGetLastError();
    nb = m->kind == NI_KIND_LARGE ? 1 : 4;
  }
  if (const char *env = getenv("NI_B200_BLOCKS_PER_SM")) {  // experiments: fewer resident CTAs than fit
    const int want = atoi(env);
    if (want >= 1 && want < nb) nb = want;
  }
  e->blocks_per_sm = nb;
  if (const char *env = getenv("NI_B200_PARK_BELOW")) {
    const int want = atoi(env);
    if (want >= 0 && want <= 32) e->park_below = want;
  }
  if (m->kind == NI_KIND_SMALL) {
    // a warp parks at most 31 trajectories per pass
    const int64_t warps = (int64_t)e->sms * nb * ((m->block + 31) / 32);
    e->park_cap = warps * 31 < N ? warps * 31 : N;
    for (int k = 0; k < 2 && rc == NI_OK; ++k) rc = dalloc(e, &e->park_list[k], (size_t)e->park_cap);
    if (rc != NI_OK) {
      std::string keep = g_err;
      ni_ensemble_destroy(e);
      g_err = keep;
      return rc;
    }
  }
  e->q_host = g_pinned.get();
  *out = e;
  return NI_OK;
}

NI_API int ni_ensemble_set_stream(ni_ensemble *e, void *cuda_stream) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  if (e->own_stream) cudaStreamDestroy(e->stream);
  e->stream = (cudaStream_t)cuda_stream;
  e->own_stream = false;
  return NI_OK;
}

NI_API int ni_ensemble_set_max_attempts(ni_ensemble *e, int64_t per_launch) {
  if (!e || per_launch < 1) return fail(NI_ERR_INVALID, "need an ensemble and per_launch >= 1");
  e->max_attempts = per_launch;
  return NI_OK;
}

// Per-trajectory max_step / first_step (the reference takes them per solver object, _basic.py:292-299 / 323-330).
// NULL keeps / restores the scalar of ni_ensemble_reset.  Host arrays of N doubles, copied asynchronously: they must
// stay valid until the next synchronising call; takes effect at the next ni_ensemble_reset / run.
NI_API int ni_ensemble_set_step_limits(ni_ensemble *e, const double *max_step, const double *first_step) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  const size_t N = (size_t)e->N;
  if (max_step) {
    if (!e->max_step_d)
      if (int rc = dalloc(e, &e->max_step_d, N)) return rc;
    CU(cudaMemcpyAsync(e->max_step_d, max_step, sizeof(double) * N, cudaMemcpyHostToDevice, e->stream));
  }
  if (first_step) {
    if (!e->first_step_d)
      if (int rc = dalloc(e, &e->first_step_d, N)) return rc;
    CU(cudaMemcpyAsync(e->first_step_d, first_step, sizeof(double) * N, cudaMemcpyHostToDevice, e->stream));
  }
  e->a.max_step_v = max_step ? e->max_step_d : nullptr;
  e->a.first_step_v = first_step ? e->first_step_d : nullptr;
  return NI_OK;
}

NI_API int ni_ensemble_compaction(ni_ensemble *e, int park_below, int64_t *parked_last_run) {
  if (!e || park_below > 32) return fail(NI_ERR_INVALID, "need an ensemble and park_below <= 32");
  CU(cudaSetDevice(e->device));
  if (park_below >= 0) e->park_below = park_below;
  if (parked_last_run) {
    unsigned long long all[8 * NI_PASSES];
    CU(cudaMemcpyAsync(all, e->queue_all, sizeof all, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    *parked_last_run = 0;
    for (int k = 0; k < e->last_passes; ++k) *parked_last_run += (int64_t)all[8 * k + 4];
  }
  return NI_OK;
}

NI_API int ni_ensemble_upload(ni_ensemble *e, const double *t0, int t0_per_traj, const double *y0, const double *params,
                              const double *t_bound, int t_bound_per_traj) {
  if (!e || !t0 || !y0 || !t_bound) return fail(NI_ERR_INVALID, "ensemble, t0, y0 and t_bound are required");
  if (e->mod.P > 0 && !params) return fail(NI_ERR_INVALID, "this right-hand side needs %d parameters per trajectory", e->mod.P);
  CU(cudaSetDevice(e->device));
  const size_t N = (size_t)e->N;
  CU(cudaMemcpyAsync(e->t0_d, t0, sizeof(double) * (t0_per_traj ? N : 1), cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->tb0_d, t_bound, sizeof(double) * (t_bound_per_traj ? N : 1), cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->y0_rows, y0, sizeof(double) * N * e->mod.n, cudaMemcpyHostToDevice, e->stream));
  if (e->mod.P > 0)
    CU(cudaMemcpyAsync(e->params_rows, params, sizeof(double) * N * e->mod.P, cudaMemcpyHostToDevice, e->stream));
  e->a.t0_stride = t0_per_traj ? 1 : 0;
  e->a.tb_stride = t_bound_per_traj ? 1 : 0;
  e->uploaded = true;
  return NI_OK;
}

static int launch(ni_ensemble *e, const void *k, int grid, int block, size_t smem) {
  void *args[] = {(void *)&e->a};
  CU(cudaLaunchKernel(k, dim3((unsigned)grid), dim3((unsigned)block), args, smem, e->stream));
  e->n_launches++;
  return NI_OK;
}

static int run_grid(const ni_ensemble *e) {
  if (e->mod.kind == NI_KIND_LARGE) {
    long long g = (long long)e->sms * e->blocks_per_sm;
    return (int)(e->N < g ? e->N : g);
  }
  long long need = (e->N + e->mod.block - 1) / e->mod.block;
  long long cap = (long long)e->sms * e->blocks_per_sm;
  return (int)(need < cap ? need : cap);
}

NI_API int ni_ensemble_reset(ni_ensemble *e, const double *rtol, const double *atol, double max_step, double first_step) {
  if (!e || !rtol || !atol) return fail(NI_ERR_INVALID, "ensemble, rtol and atol are required");
  if (!e->uploaded) return fail(NI_ERR_INVALID, "ni_ensemble_upload has not been called");
  CU(cudaSetDevice(e->device));
  const int n = e->mod.n;
  if (e->mod.kind == NI_KIND_SMALL) {
    bool floor_ok = true;
    for (int i = 0; i < n; ++i) {
      e->a.rtol[i] = rtol[i];
      e->a.atol[i] = atol[i];
      floor_ok = floor_ok && atol[i] >= 6.223015277861142e-61 && atol[i] < INFINITY;  // 2^-200 (ni_guard_lite_out)
    }
    e->a.atol_floor = floor_ok ? 1 : 0;
  } else {
    bool uni = true;
    for (int i = 1; i < n; ++i) uni = uni && rtol[i] == rtol[0] && atol[i] == atol[0];
    e->a.tol_uniform = uni ? 1 : 0;
    e->a.rtol[0] = rtol[0];
    e->a.atol[0] = atol[0];
    // rtol / atol are caller-owned: copy them through an ensemble-owned staging vector, and only when they changed
    // (a stream synchronise here cost the double-buffered end-to-end loop of C5 7 %)
    bool same = e->tol_host.size() == 2 * (size_t)n;
    for (int i = 0; same && i < n; ++i) same = e->tol_host[i] == rtol[i] && e->tol_host[n + i] == atol[i];
    if (!same) {
      CU(cudaStreamSynchronize(e->stream));  // an earlier copy may still read the staging vector
      e->tol_host.assign(2 * (size_t)n, 0.0);
      for (int i = 0; i < n; ++i) {
        e->tol_host[i] = rtol[i];
        e->tol_host[n + i] = atol[i];
      }
      CU(cudaMemcpyAsync(e->mod.large.tol_d, e->tol_host.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, e->stream));
    }
  }
  e->a.max_step = max_step;
  e->a.first_step = first_step;
  CU(cudaMemsetAsync(e->a.cond_hit, 0, (size_t)e->N, e->stream));
  CU(cudaMemsetAsync(e->a.step_rejected, 0, (size_t)e->N, e->stream));
  CU(cudaMemsetAsync(e->a.rec_cursor, 0, sizeof(unsigned long long), e->stream));
  CU(cudaEventRecord(e->ev[0], e->stream));
  int grid;
  if (e->mod.kind == NI_KIND_LARGE) {
    grid = run_grid(e);
  } else {
    long long need = (e->N + NI_BLOCK - 1) / NI_BLOCK, cap = (long long)e->sms * 16;
    grid = (int)(need < cap ? need : cap);
  }
  if (int rc = launch(e, e->mod.k_init, grid, e->mod.block, e->mod.kind == NI_KIND_LARGE ? e->mod.smem : 0)) return rc;
  CU(cudaEventRecord(e->ev[1], e->stream));
  e->have_init_t = true;
  e->initialised = true;
  return NI_OK;
}

NI_API int ni_ensemble_sync(ni_ensemble *e) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  return NI_OK;
}

NI_API int ni_ensemble_init(ni_ensemble *e, const double *t0, int t0_per_traj, const double *y0, const double *params,
                            const double *t_bound, int t_bound_per_traj, const double *rtol, const double *atol,
                            double max_step, double first_step) {
  if (int rc = ni_ensemble_upload(e, t0, t0_per_traj, y0, params, t_bound, t_bound_per_traj)) return rc;
  if (int rc = ni_ensemble_reset(e, rtol, atol, max_step, first_step)) return rc;
  return ni_ensemble_sync(e);
}

// Number of launches of one fast-forward run: 1, or NI_PASSES when the tail is compacted (the ensemble fills the
// resident grid at least twice -- smaller ones have no steady state to protect -- and nothing bounds the launch).
static int run_passes(const ni_ensemble *e, int64_t max_accepts) {
  if (e->mod.kind != NI_KIND_SMALL || e->park_below <= 0 || max_accepts < 0x7fffffffLL) return 1;
  const int64_t lanes = (int64_t)e->sms * e->blocks_per_sm * e->mod.block;
  return e->N >= 2 * lanes ? NI_PASSES : 1;
}

// one run of the run kernel (a single launch, or the passes of a compacted run); counters are left in e->queue_all
static int launch_run(ni_ensemble *e, int64_t max_accepts, bool timed) {
  if (!e->initialised) return fail(NI_ERR_INVALID, "ensemble is not initialised (call ni_ensemble_init)");
  e->a.max_accepts = max_accepts;
  e->a.max_attempts = e->max_attempts;
  // record slots: warps of a long fast-forward reserve 512 at a time (one global cursor cannot serve ~1e9
  // reservations/s); lock-step calls and small ensembles reserve exactly what they write
  e->a.rec_chunk = (max_accepts == 1 || e->N < 65536 || e->mod.kind != NI_KIND_SMALL) ? 0 : 512;
  CU(cudaMemsetAsync(e->queue_all, 0, sizeof(unsigned long long) * 8 * NI_PASSES, e->stream));
  const bool rec = e->a.rec_cap > 0;
  const int passes = run_passes(e, max_accepts);
  e->last_passes = passes;
  if (timed) CU(cudaEventRecord(e->ev[2], e->stream));
  int rc = NI_OK;
  for (int k = 0; k < passes && rc == NI_OK; ++k) {
    e->a.queue = e->queue_all + 8 * k;
    e->a.work_list = k == 0 ? e->order_d : e->park_list[(k - 1) & 1];
    e->a.work_count = k == 0 ? e->order_n_d : e->queue_all + 8 * (k - 1) + 4;
    e->a.park_list = e->park_list[k & 1];
    e->a.park_below = k + 1 < passes ? e->park_below : 0;  // the last pass finishes everything it holds
    rc = launch(e, e->mod.k_run[rec ? 1 : 0], run_grid(e), e->mod.block, e->mod.smem);
  }
  e->a.queue = e->queue_all;
  e->a.work_list = nullptr;
  e->a.work_count = nullptr;
  e->a.park_below = 0;
  if (rc != NI_OK) return rc;
  if (timed) {
    CU(cudaEventRecord(e->ev[3], e->stream));
    e->have_run_t = true;
  }
  return NI_OK;
}

// counters of the last run: [0] cursor of the first pass, [1] flags (OR over the passes), [2] trajectories that left a
// pass still RUNNING without being parked for the next one (budget exhausted, record log full), [3] accepted lock-steps
static int read_queue(ni_ensemble *e, unsigned long long q[4]) {
  unsigned long long *dst = e->q_host ? e->q_host : e->q_pageable;
  unsigned long long more[8 * (NI_PASSES - 1)];
  const int passes = e->last_passes;
  CU(cudaMemcpyAsync(dst, e->queue_all, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, e->stream));
  if (passes > 1)
    CU(cudaMemcpyAsync(more, e->queue_all + 8, sizeof(unsigned long long) * 8 * (passes - 1), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < 4; ++i) q[i] = dst[i];
  for (int k = 1; k < passes; ++k) {
    q[1] |= more[8 * (k - 1) + 1];
    q[2] += more[8 * (k - 1) + 2];
    q[3] += more[8 * (k - 1) + 3];
  }
  return NI_OK;
}

// One lock-step launch as a CUDA graph.  Returns NI_OK with *used = false when graphs cannot be used on this stream
// (the legacy default stream cannot be captured) -- the caller then launches directly.
static int step_with_graph(ni_ensemble *e, unsigned long long q[4], bool *used) {
  *used = false;
  if (!e->initialised) return fail(NI_ERR_INVALID, "ensemble is not initialised (call ni_ensemble_init)");
  if (e->stream == nullptr || e->stream == cudaStreamLegacy) return NI_OK;
  if (++e->step_calls <= 64) return NI_OK;  // short loops (the README's 11 steps) launch directly
  e->a.max_accepts = 1;
  e->a.max_attempts = e->max_attempts;
  e->a.rec_chunk = 0;
  const void *kernel = e->mod.k_run[e->a.rec_cap > 0 ? 1 : 0];
  if (!e->q_host) return NI_OK;  // no pinned slot for the read-back node
  if (!e->step_graph || kernel != e->step_kernel || e->stream != e->step_stream ||
      memcmp(&e->a, &e->step_args, sizeof(NiArgs)) != 0) {
    if (e->step_graph) {
      cudaGraphExecDestroy(e->step_graph);
      e->step_graph = nullptr;
    }
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return NI_OK;  // not capturable: direct launches
    }
    cudaMemsetAsync(e->queue_all, 0, sizeof(unsigned long long) * 8 * NI_PASSES, e->stream);
    void *args[] = {(void *)&e->a};
    cudaLaunchKernel(kernel, dim3((unsigned)run_grid(e)), dim3((unsigned)e->mod.block), args, e->mod.smem, e->stream);
    cudaMemcpyAsync(e->q_host, e->a.queue, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, e->stream);
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    if (ce != cudaSuccess || !g) {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      return NI_OK;
    }
    const cudaError_t ci = cudaGraphInstantiate(&e->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (ci != cudaSuccess) {
      cudaGetLastError();
      e->step_graph = nullptr;
      return NI_OK;
    }
    memcpy(&e->step_args, &e->a, sizeof(NiArgs));
    e->step_kernel = kernel;
    e->step_stream = e->stream;
  }
  CU(cudaGraphLaunch(e->step_graph, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < 4; ++i) q[i] = e->q_host[i];
  e->n_launches++;
  e->have_run_t = false;  // no timing events inside the graph
  *used = true;
  return NI_OK;
}

NI_API int ni_ensemble_step(ni_ensemble *e, int64_t *n_true) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  unsigned long long q[4];
  bool used = false;
  if (int rc = step_with_graph(e, q, &used)) return rc;
  if (!used) {
    if (int rc = launch_run(e, 1, true)) return rc;
    if (int rc = read_queue(e, q)) return rc;
  }
  if (n_true) *n_true = (int64_t)q[3];
  if (q[1] & NI_FLAG_REC_OVERFLOW) return fail(NI_ERR_REC_OVERFLOW, "record log full (%lld records): drain it and step again", (long long)e->a.rec_cap);
  return NI_OK;
}

NI_API int ni_ensemble_run_async(ni_ensemble *e) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  return launch_run(e, 1ll << 62, true);
}

// Outcome of the most recent run / run_async / step: synchronises, then reports what the launch(es) left behind.
NI_API int ni_ensemble_poll(ni_ensemble *e, int64_t *n_unfinished, int64_t *n_failed) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  unsigned long long all[8 * NI_PASSES];
  CU(cudaMemcpyAsync(all, e->queue_all, sizeof all, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  unsigned long long flags = 0, unfinished = 0, failed = 0;
  for (int k = 0; k < e->last_passes; ++k) {
    flags |= all[8 * k + 1];
    unfinished += all[8 * k + 2];
    failed += all[8 * k + 5];
  }
  if (n_unfinished) *n_unfinished = (int64_t)unfinished;
  if (n_failed) *n_failed = (int64_t)failed;
  if (flags & NI_FLAG_REC_OVERFLOW)
    return fail(NI_ERR_REC_OVERFLOW, "record log full (%lld records): drain it and run again", (long long)e->a.rec_cap);
  return NI_OK;
}

NI_API int ni_ensemble_run(ni_ensemble *e, int64_t *n_unfinished) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  for (;;) {
    if (int rc = launch_run(e, 1ll << 62, true)) return rc;
    unsigned long long q[4];
    if (int rc = read_queue(e, q)) return rc;
    if (n_unfinished) *n_unfinished = (int64_t)q[2];
    if (q[1] & NI_FLAG_REC_OVERFLOW) return fail(NI_ERR_REC_OVERFLOW, "record log full (%lld records): drain it and run again", (long long)e->a.rec_cap);
    if (q[2] == 0) return NI_OK;  // otherwise max_attempts per launch was exhausted: relaunch
  }
}

// ---- small utility kernels (host-facing plumbing, not on the hot path)
__global__ void ni_k_tbound(NiArgs a, int mode, double t_end, double *saved, unsigned char *flag, const double *newtb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
    double tb = a.t_bound[i];
    if (mode == 0) {  // set from array
      tb = newtb[i];
    } else if (mode == 1) {  // ff2t begin (_API.py:30-33)
      saved[i] = tb;
      const bool not_last = tb > t_end;
      flag[i] = not_last ? 1 : 0;
      if (not_last) tb = t_end;
    } else if (mode == 2) {  // ff2t end (_API.py:37)
      tb = saved[i];
    }  // mode 3: t / direction were overwritten, only the finished <-> running status below is re-evaluated
    a.t_bound[i] = tb;
    if (a.status[i] == NI_ST_FINISHED && !(a.direction[i] * (a.t[i] - tb) >= 0.0)) a.status[i] = NI_ST_RUNNING;
  }
}

// component-major [c][N] <-> row-major [N][c]
// row-major [N][c] -> component-major [c][N] (field writes of the thread-per-trajectory layout)
__global__ void ni_k_from_rows(const double *src, double *dst, long long N, int c) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < N * c; k += (long long)gridDim.x * blockDim.x) {
    const long long i = k / c;
    const int j = (int)(k - i * c);
    dst[(long long)j * N + i] = src[k];
  }
}
__global__ void ni_k_to_rows(const double *src, double *dst, long long N, int c) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < N * c; k += (long long)gridDim.x * blockDim.x) {
    const long long i = k / c;
    const int j = (int)(k - i * c);
    dst[k] = src[(long long)j * N + i];
  }
}
// ---- draining the record log: order-preserving compaction of a window of slots into row-major staging buffers
// (slots reserved in chunks but never written carry NI_REC_INVALID).  Three small kernels per window: valid slots per
// block of NI_DRAIN_BLOCK slots, exclusive scan of the block counts (one CTA), scatter.
#define NI_DRAIN_BLOCK 4096
__global__ void __launch_bounds__(256) ni_k_drain_count(const unsigned *rec_traj, long long s0, long long count, unsigned *block_count) {
  __shared__ unsigned warp_sums[8];
  const long long b0 = s0 + (long long)blockIdx.x * NI_DRAIN_BLOCK;
  unsigned c = 0;
  for (int k = threadIdx.x; k < NI_DRAIN_BLOCK; k += 256) {
    const long long sl = b0 + k;
    c += (sl < s0 + count && rec_traj[sl] != NI_REC_INVALID) ? 1u : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int w = 0; w < 8; ++w) tot += warp_sums[w];
    block_count[blockIdx.x] = tot;
  }
}
__global__ void __launch_bounds__(1024) ni_k_drain_scan(unsigned *block_count, int nblocks, unsigned long long *total) {
  // exclusive scan in place (nblocks <= a few thousand: one CTA, sequential over tiles of 1024)
  __shared__ unsigned tile[1024];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + (int)threadIdx.x;
    const unsigned v = i < nblocks ? block_count[i] : 0u;
    tile[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned add = threadIdx.x >= (unsigned)o ? tile[threadIdx.x - o] : 0u;
      __syncthreads();
      tile[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < nblocks) block_count[i] = carry + tile[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += tile[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
// small != 0: rec_y / rec_aux are component-major [c][cap]; else row-major [cap][c]
__global__ void __launch_bounds__(256) ni_k_drain_pack(NiArgs a, int small, int n, int Q, long long s0, long long count,
                                                       const unsigned *block_offset, unsigned *o_traj, double *o_t, double *o_y,
                                                       double *o_aux) {
  __shared__ unsigned warp_base[8];
  __shared__ unsigned running;
  if (threadIdx.x == 0) running = block_offset[blockIdx.x];
  const long long b0 = s0 + (long long)blockIdx.x * NI_DRAIN_BLOCK;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  __syncthreads();
  for (int k = threadIdx.x; k < NI_DRAIN_BLOCK; k += 256) {  // consecutive tiles of 256 slots, in slot order
    const long long sl = b0 + k;
    const unsigned tr = sl < s0 + count ? a.rec_traj[sl] : NI_REC_INVALID;
    const bool valid = tr != NI_REC_INVALID;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_base[warp] = __popc(m);
    __syncthreads();
    unsigned before = running;
    for (unsigned w = 0; w < warp; ++w) before += warp_base[w];
    if (valid) {
      const size_t o = before + __popc(m & ((1u << lane) - 1u));
      o_traj[o] = tr;
      o_t[o] = a.rec_t[sl];
      for (int i = 0; i < n; ++i) o_y[o * n + i] = small ? a.rec_y[(long long)i * a.rec_cap + sl] : a.rec_y[sl * n + i];
      for (int i = 0; i < Q; ++i) o_aux[o * Q + i] = small ? a.rec_aux[(long long)i * a.rec_cap + sl] : a.rec_aux[sl * Q + i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_base[w];
      running += tot;
    }
    __syncthreads();
  }
}

// ---- terminal records (the final gather moves terminal t, y, aux and the step statistics)
// One packed row per trajectory: double t, y[n], aux[Q]; int32 n_accepted, n_rejected, status -- 8 (1 + n + Q) + 12
// bytes (52 for Lorenz-63 via Advanced).  One thread per 32-bit word of the output, so the stores are coalesced; the
// loads hit the row-major mirror of y (small kernels) or the trajectory-major state (large kernels).
__global__ void ni_k_pack_terminal(NiArgs a, int small, int n, int Q, unsigned *dst) {
  const int dwords = 2 * (1 + n + Q), W = dwords + 3;
  const long long total = a.N * W;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long i = g / W;
    const int w = (int)(g - i * W);
    unsigned out;
    if (w < dwords) {
      const int d = w >> 1;
      double v;
      if (d == 0)
        v = a.t[i];
      else if (d <= n)
        v = small ? (a.y_rows ? a.y_rows[i * n + (d - 1)] : a.y[(long long)(d - 1) * a.N + i]) : a.y[i * n + (d - 1)];
      else
        v = small ? a.aux[(long long)(d - 1 - n) * a.N + i] : a.aux[i * Q + (d - 1 - n)];
      out = (w & 1) ? (unsigned)__double2hiint(v) : (unsigned)__double2loint(v);
    } else {
      const int k = w - dwords;
      out = (unsigned)(k == 0 ? a.n_acc[i] : k == 1 ? a.n_rej[i] : a.status[i]);
    }
    dst[g] = out;
  }
}

// The same packing with the gather fused in: every word is stored to `nd` destinations -- this GPU's slot in the
// receive buffer of every rank, peer memory reached over NVLink / NVSwitch (symmetric-memory allocations whose peer
// addresses the caller obtained at rendezvous).  The transfer overlaps the packing word by word; no staging copy, no
// library collective.  Completion on the peers is the caller's barrier.
struct NiPeerDst {
  unsigned *p[NI_MAX_PEERS];
  int nd;
};
// one 32-bit word of the packed terminal records: word w of trajectory i
__device__ __forceinline__ unsigned ni_terminal_word(const NiArgs &a, int small, int n, int Q, int dwords, long long i, int w) {
  if (w < dwords) {
    const int c = w >> 1;
    double v;
    if (c == 0)
      v = a.t[i];
    else if (c <= n)
      v = small ? (a.y_rows ? a.y_rows[i * n + (c - 1)] : a.y[(long long)(c - 1) * a.N + i]) : a.y[i * n + (c - 1)];
    else
      v = small ? a.aux[(long long)(c - 1 - n) * a.N + i] : a.aux[i * Q + (c - 1 - n)];
    return (w & 1) ? (unsigned)__double2hiint(v) : (unsigned)__double2loint(v);
  }
  const int k = w - dwords;
  return (unsigned)(k == 0 ? a.n_acc[i] : k == 1 ? a.n_rej[i] : a.status[i]);
}
// Tiles of NI_PACK_TILE trajectories (a multiple of 4, so a tile starts on a 16-byte boundary of the packed buffer);
// inside a tile the (trajectory, word) split is a 32-bit division.  16-byte stores: a warp sends 512 contiguous bytes
// per destination and instruction over NVLink.
#define NI_PACK_TILE 1024
__global__ void ni_k_pack_terminal_peers(NiArgs a, int small, int n, int Q, NiPeerDst d) {
  const int dwords = 2 * (1 + n + Q), W = dwords + 3;
  const long long tiles = (a.N + NI_PACK_TILE - 1) / NI_PACK_TILE;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long i0 = tile * NI_PACK_TILE;
    const unsigned rows = (unsigned)(a.N - i0 < NI_PACK_TILE ? a.N - i0 : NI_PACK_TILE);
    const unsigned words = rows * (unsigned)W, quads = words >> 2;
    const long long g0 = i0 * W;  // a multiple of 4
    for (unsigned q = threadIdx.x; q < quads; q += blockDim.x) {
      unsigned o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned gl = 4u * q + k, il = gl / (unsigned)W;
        o[k] = ni_terminal_word(a, small, n, Q, dwords, i0 + il, (int)(gl - il * (unsigned)W));
      }
      const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
      for (int r = 0; r < d.nd; ++r) ((uint4 *)(d.p[r] + g0))[q] = v;
    }
    if (threadIdx.x < (words & 3u)) {  // the last 1..3 words of the buffer (only the last tile can have them)
      const unsigned gl = (quads << 2) + threadIdx.x, il = gl / (unsigned)W;
      const unsigned out = ni_terminal_word(a, small, n, Q, dwords, i0 + il, (int)(gl - il * (unsigned)W));
      for (int r = 0; r < d.nd; ++r) d.p[r][g0 + gl] = out;
    }
  }
}

static int grid_for(long long work) {
  long long g = (work + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

NI_API int ni_ensemble_ff2t(ni_ensemble *e, double t_end, uint8_t *is_not_last) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  ni_k_tbound<<<grid_for(e->N), 256, 0, e->stream>>>(e->a, 1, t_end, e->tb_saved, e->flag_d, nullptr);
  CU(cudaGetLastError());
  int64_t unfinished = 0;
  int rc = ni_ensemble_run(e, &unfinished);
  ni_k_tbound<<<grid_for(e->N), 256, 0, e->stream>>>(e->a, 2, t_end, e->tb_saved, e->flag_d, nullptr);
  CU(cudaGetLastError());
  if (is_not_last) CU(cudaMemcpyAsync(is_not_last, e->flag_d, (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return rc;
}

NI_API int ni_ensemble_ff2cond_user(ni_ensemble *e, const char *cond_body, const double *cond_params, int n_params,
                                    uint8_t *hit, int64_t *n_hit);

// Built-in threshold predicates (ni.y_above(i) ...): generated as a one-line predicate body and compiled into a copy
// of the run kernel like a user predicate; the threshold travels as cp[0], so one kernel per (kind, index) serves
// every value.  The ahead-of-time kernels carry no predicate code (ni_core.cuh).
NI_API int ni_ensemble_ff2cond(ni_ensemble *e, int cond_kind, int index, double value, uint8_t *hit, int64_t *n_hit) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  if (cond_kind < NI_COND_Y_GT || cond_kind > NI_COND_AUX_LT) return fail(NI_ERR_INVALID, "unknown predicate kind %d", cond_kind);
  const int lim = cond_kind <= NI_COND_Y_LT ? e->mod.n : e->mod.Q;
  if (index < 0 || index >= lim) return fail(NI_ERR_INVALID, "predicate index %d out of range [0, %d)", index, lim);
  const bool on_y = cond_kind <= NI_COND_Y_LT, gt = cond_kind == NI_COND_Y_GT || cond_kind == NI_COND_AUX_GT;
  const std::string body = std::string("  return ") + (on_y ? "y[" : "aux[") + std::to_string(index) + "] " + (gt ? ">" : "<") + " cp[0];";
  return ni_ensemble_ff2cond_user(e, body.c_str(), &value, 1, hit, n_hit);
}

// ni.ff2cond with a user predicate: the body of
//   __device__ bool cond(double t, const double* y, const double* aux, const double* cp)
// is compiled (NVRTC) into a copy of this ensemble's run kernel and evaluated after every accepted step.
NI_API int ni_ensemble_ff2cond_user(ni_ensemble *e, const char *cond_body, const double *cond_params, int n_params,
                                    uint8_t *hit, int64_t *n_hit) {
  if (!e || !cond_body) return fail(NI_ERR_INVALID, "ensemble / cond_body is NULL");
  if (n_params < 0 || n_params > 8 || (n_params > 0 && !cond_params)) return fail(NI_ERR_INVALID, "a predicate takes 0..8 parameters");
  if (e->mod.rhs_src.empty()) return fail(NI_ERR_UNSUPPORTED, "this module does not carry the source of its right-hand side");
  CU(cudaSetDevice(e->device));
  // run kernels with a predicate compiled in, shared by every ensemble of the process (key: generated source +
  // options); never unloaded -- a few hundred KB per distinct (RHS, method, predicate)
  static std::mutex cond_mu;
  static std::vector<std::pair<std::string, ni_ensemble::CondKernels>> cond_cache;
  ni_ensemble::CondKernels found, *ck = nullptr;
  {
    const ni_module &m = e->mod;
    const std::string M = m.method == NI_RK45 ? "NI_RK45" : "NI_RK23", A = m.scipy ? "2" : m.advanced ? "1" : "0",
                      F = m.fast ? "true" : "false";
    std::string src = "#define NI_HAVE_USER_COND 1\n#include \"ni_core.cuh\"\n" + m.rhs_src;
    src += "__device__ __forceinline__ bool ni_user_cond(double t, const double* y, const double* aux, const double* cp) {\n";
    src += "  (void)t; (void)y; (void)aux; (void)cp;\n";
    src += cond_body;
    src += "\n}\n";
    src += "extern \"C\" __global__ void __launch_bounds__(NI_BLOCK) ni_user_init(const NiArgs a) { (void)a; }\n";
    src += "extern \"C\" __global__ void __launch_bounds__(NI_BLOCK, NI_MIN_BLOCKS(NiUserRhs::N)) ni_user_run(const NiArgs a) { "
           "ni_small_run_body<NiUserRhs, " + M + ", " + A + ", false, " + F + ">(a); }\n";
    src += "extern \"C\" __global__ void __launch_bounds__(NI_BLOCK, NI_MIN_BLOCKS(NiUserRhs::N)) ni_user_run_rec(const NiArgs a) { "
           "ni_small_run_body<NiUserRhs, " + M + ", " + A + ", true, " + F + ">(a); }\n";
    const std::string opts = m.fast ? "--fmad=true" : "--fmad=false";
    std::string cond_src;
    if (m.kind == NI_KIND_LARGE) {  // CTA-per-trajectory: the same templates as the ensemble's module + the predicate
      cond_src = "__device__ __forceinline__ bool ni_user_cond(double t, const double* y, const double* aux, const double* cp) {\n"
                 "  (void)t; (void)y; (void)aux; (void)cp;\n" + std::string(cond_body) + "\n}\n";
      src = "large n=" + std::to_string(m.n) + " " + M + A + F + (m.component ? "c" : "s") + "\n" + m.rhs_src + cond_src;
    }
    const std::string key = opts + "\n" + src;
    std::lock_guard<std::mutex> lock(cond_mu);
    for (auto &c : cond_cache)
      if (c.first == key) {
        found = c.second;
        ck = &found;
      }
    if (!ck) {
      if (m.kind == NI_KIND_LARGE) {
        ni_module *tmp = nullptr;
        const int mode = (m.advanced ? NI_MODE_ADVANCED : 0) | (m.fast ? NI_MODE_FAST : 0) | (m.scipy ? NI_MODE_SCIPY : 0);
        if (int rc = compile_stencil_module(m.rhs_src, m.n, m.P, m.Q, m.method, mode, nullptr, &tmp, m.component, cond_src)) return rc;
        found.lib = tmp->lib;
        found.k_run[0] = tmp->k_run[0];
        found.k_run[1] = tmp->k_run[1];
        for (const void *k : {tmp->k_run[0], tmp->k_run[1]}) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.smem);
        tmp->lib = nullptr;  // the kernels stay loaded for the process (cond_cache)
        delete tmp;
      } else {
        ni_module tmp;
        const char *names[3] = {"ni_user_init", "ni_user_run", "ni_user_run_rec"};
        if (int rc = compile_and_load(src, opts.c_str(), names, &tmp)) return rc;
        found.lib = tmp.lib;
        found.k_run[0] = tmp.k_run[0];
        found.k_run[1] = tmp.k_run[1];
      }
      cond_cache.emplace_back(key, found);
      ck = &found;
    }
  }
  CU(cudaMemsetAsync(e->a.cond_hit, 0, (size_t)e->N, e->stream));
  e->a.cond_kind = 5;
  for (int i = 0; i < 8; ++i) e->a.cond_params[i] = i < n_params ? cond_params[i] : 0.0;
  const void *saved[2] = {e->mod.k_run[0], e->mod.k_run[1]};
  e->mod.k_run[0] = ck->k_run[0];
  e->mod.k_run[1] = ck->k_run[1];
  int64_t unfinished = 0;
  int rc = ni_ensemble_run(e, &unfinished);
  e->mod.k_run[0] = saved[0];
  e->mod.k_run[1] = saved[1];
  e->a.cond_kind = 0;
  if (rc != NI_OK) return rc;
  if (hit || n_hit) {
    std::vector<uint8_t> tmpv;
    uint8_t *dst = hit;
    if (!dst) {
      tmpv.resize((size_t)e->N);
      dst = tmpv.data();
    }
    CU(cudaMemcpyAsync(dst, e->a.cond_hit, (size_t)e->N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (n_hit) {
      int64_t c = 0;
      for (int64_t i = 0; i < e->N; ++i) c += dst[i] != 0;
      *n_hit = c;
    }
  }
  return NI_OK;
}

struct FieldInfo {
  void *ptr;
  int comps;      // components per trajectory
  size_t elem;    // bytes per element
  bool is_double;
};

static bool field_info(ni_ensemble *e, int field, FieldInfo *f) {
  NiArgs &a = e->a;
  switch (field) {
    case NI_F_T: *f = {a.t, 1, 8, true}; return true;
    case NI_F_Y: *f = {a.y, e->mod.n, 8, true}; return true;
    case NI_F_AUX: *f = {a.aux, e->mod.Q, 8, true}; return true;
    case NI_F_H_ABS: *f = {a.h_abs, 1, 8, true}; return true;
    case NI_F_STEP_SIZE: *f = {a.step_size, 1, 8, true}; return true;
    case NI_F_KLAST: *f = {a.klast, e->mod.n, 8, true}; return true;
    case NI_F_STATUS: *f = {a.status, 1, 4, false}; return true;
    case NI_F_N_ACCEPTED: *f = {a.n_acc, 1, 4, false}; return true;
    case NI_F_N_REJECTED: *f = {a.n_rej, 1, 4, false}; return true;
    case NI_F_RUNNING: *f = {a.running, 1, 1, false}; return true;
    case NI_F_T_BOUND: *f = {a.t_bound, 1, 8, true}; return true;
    case NI_F_DIRECTION: *f = {a.direction, 1, 8, true}; return true;
    case NI_F_PARAMS: *f = {a.params, e->mod.P, 8, true}; return true;
    case NI_F_COND_HIT: *f = {a.cond_hit, 1, 1, false}; return true;
    default: return false;
  }
}

NI_API int ni_ensemble_device_ptr(ni_ensemble *e, int field, void **dptr) {
  FieldInfo f;
  if (e && dptr && field >= NI_F_REC_TRAJ && field <= NI_F_REC_AUX) {
    if (e->a.rec_cap <= 0) return fail(NI_ERR_INVALID, "recording is not enabled");
    if (field == NI_F_REC_STEP)
      return fail(NI_ERR_UNSUPPORTED, "the log stores no step numbers: the records of a trajectory appear in step order "
                  "(ni_ensemble_record_drain numbers them)");
    void *ptrs[] = {e->a.rec_traj, nullptr, e->a.rec_t, e->a.rec_y, e->a.rec_aux};
    *dptr = ptrs[field - NI_F_REC_TRAJ];
    return NI_OK;
  }
  if (!e || !dptr || !field_info(e, field, &f)) return fail(NI_ERR_INVALID, "bad ensemble / field %d", field);
  *dptr = f.ptr;
  return NI_OK;
}

NI_API int ni_ensemble_get(ni_ensemble *e, int field, void *host_out) {
  FieldInfo f;
  if (!e || !host_out || !field_info(e, field, &f)) return fail(NI_ERR_INVALID, "bad ensemble / field %d / NULL buffer", field);
  CU(cudaSetDevice(e->device));
  const size_t N = (size_t)e->N;
  if (f.comps == 0) return NI_OK;
  if (field == NI_F_Y && e->mod.kind == NI_KIND_SMALL && e->a.y_rows) {  // kept row-major by the kernels
    CU(cudaMemcpyAsync(host_out, e->a.y_rows, N * f.comps * 8, cudaMemcpyDeviceToHost, e->stream));
  } else if (f.comps > 1 && e->mod.kind == NI_KIND_SMALL) {  // component-major on device -> rows
    if (int rc = need_scratch(e, N * f.comps * 8)) return rc;
    ni_k_to_rows<<<grid_for((long long)N * f.comps), 256, 0, e->stream>>>((const double *)f.ptr, (double *)e->scratch, (long long)N, f.comps);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_out, e->scratch, N * f.comps * 8, cudaMemcpyDeviceToHost, e->stream));
  } else {
    CU(cudaMemcpyAsync(host_out, f.ptr, N * f.comps * f.elem, cudaMemcpyDeviceToHost, e->stream));
  }
  CU(cudaStreamSynchronize(e->stream));
  return NI_OK;
}

NI_API int ni_ensemble_set(ni_ensemble *e, int field, const void *host_in) {
  if (!e || !host_in) return fail(NI_ERR_INVALID, "ensemble / buffer is NULL");
  CU(cudaSetDevice(e->device));
  const size_t N = (size_t)e->N;
  if (field == NI_F_T_BOUND) {
    if (int rc = need_scratch(e, N * 8)) return rc;
    CU(cudaMemcpyAsync(e->scratch, host_in, N * 8, cudaMemcpyHostToDevice, e->stream));
    ni_k_tbound<<<grid_for(e->N), 256, 0, e->stream>>>(e->a, 0, 0.0, e->tb_saved, e->flag_d, (const double *)e->scratch);
    CU(cudaGetLastError());
  } else if (field == NI_F_STATUS || field == NI_F_N_ACCEPTED || field == NI_F_N_REJECTED || field == NI_F_RUNNING ||
             field == NI_F_COND_HIT || field == NI_F_PARAMS) {
    return fail(NI_ERR_INVALID, "field %d is read-only", field);
  } else {
    // the state fields of the reference's jitclass are plain attributes (_basic.py:182-199): t, y, h_abs, step_size,
    // direction, K[-1] and the auxiliary output can be overwritten between calls; the caller keeps them consistent
    FieldInfo f;
    if (!field_info(e, field, &f) || !f.is_double) return fail(NI_ERR_INVALID, "field %d cannot be written", field);
    if (f.comps == 0) return NI_OK;
    if (f.comps > 1 && e->mod.kind == NI_KIND_SMALL) {  // rows on the host, component-major on the device
      if (int rc = need_scratch(e, N * f.comps * 8)) return rc;
      CU(cudaMemcpyAsync(e->scratch, host_in, N * f.comps * 8, cudaMemcpyHostToDevice, e->stream));
      ni_k_from_rows<<<grid_for((long long)N * f.comps), 256, 0, e->stream>>>((const double *)e->scratch, (double *)f.ptr, (long long)N, f.comps);
      CU(cudaGetLastError());
      if (field == NI_F_Y && e->a.y_rows)
        CU(cudaMemcpyAsync(e->a.y_rows, e->scratch, N * f.comps * 8, cudaMemcpyDeviceToDevice, e->stream));
    } else {
      CU(cudaMemcpyAsync(f.ptr, host_in, N * f.comps * 8, cudaMemcpyHostToDevice, e->stream));
      if (field == NI_F_Y && e->a.y_rows)
        CU(cudaMemcpyAsync(e->a.y_rows, host_in, N * f.comps * 8, cudaMemcpyHostToDevice, e->stream));
    }
    if (field == NI_F_T || field == NI_F_DIRECTION) {  // a finished trajectory may be running again (and vice versa)
      ni_k_tbound<<<grid_for(e->N), 256, 0, e->stream>>>(e->a, 3, 0.0, e->tb_saved, e->flag_d, nullptr);
      CU(cudaGetLastError());
    }
  }
  CU(cudaStreamSynchronize(e->stream));
  return NI_OK;
}

// ---- terminal records for the final gather
NI_API int ni_ensemble_terminal_bytes(const ni_ensemble *e, int64_t *bytes_per_trajectory) {
  if (!e || !bytes_per_trajectory) return fail(NI_ERR_INVALID, "ensemble / bytes_per_trajectory is NULL");
  *bytes_per_trajectory = 8ll * (1 + e->mod.n + e->mod.Q) + 12;
  return NI_OK;
}

// ---- launch order by cost on the device: counting sort of the trajectories by their attempts so far, descending.
// Keys are bucketed (NI_ORDER_BINS buckets of 2^shift attempts): longest-first scheduling needs no finer order, and the
// order inside a bucket may differ from run to run (atomics) -- results never depend on the launch order.
#define NI_ORDER_BINS 2048
__global__ void __launch_bounds__(256) ni_k_order_max(const int *n_acc, const int *n_rej, long long N, unsigned *mx) {
  unsigned m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const unsigned k = (unsigned)(n_acc[i] + n_rej[i]);
    m = k > m ? k : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned v = __shfl_xor_sync(0xffffffffu, m, o);
    m = v > m ? v : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(mx, m);
}
__device__ __forceinline__ unsigned ni_order_bin(int acc, int rej, int shift) {
  const unsigned k = (unsigned)(acc + rej) >> shift;
  return NI_ORDER_BINS - 1 - (k < NI_ORDER_BINS ? k : NI_ORDER_BINS - 1);  // expensive trajectories first
}
__global__ void __launch_bounds__(256) ni_k_order_hist(const int *n_acc, const int *n_rej, long long N, int shift, unsigned *hist) {
  __shared__ unsigned h[NI_ORDER_BINS];
  for (int b = threadIdx.x; b < NI_ORDER_BINS; b += blockDim.x) h[b] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&h[ni_order_bin(n_acc[i], n_rej[i], shift)], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < NI_ORDER_BINS; b += blockDim.x)
    if (h[b]) atomicAdd(&hist[b], h[b]);
}
__global__ void __launch_bounds__(1024) ni_k_order_scan(unsigned *hist) {  // exclusive scan in place, one block
  __shared__ unsigned part[1024];
  constexpr int PER = NI_ORDER_BINS / 1024;
  unsigned v[PER], sum = 0;
  for (int k = 0; k < PER; ++k) {
    v[k] = hist[threadIdx.x * PER + k];
    sum += v[k];
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned add = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0u;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  unsigned base = part[threadIdx.x] - sum;
  for (int k = 0; k < PER; ++k) {
    hist[threadIdx.x * PER + k] = base;
    base += v[k];
  }
}
__global__ void __launch_bounds__(256) ni_k_order_scatter(const int *n_acc, const int *n_rej, long long N, int shift,
                                                           unsigned *cursor, unsigned *order) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    order[atomicAdd(&cursor[ni_order_bin(n_acc[i], n_rej[i], shift)], 1u)] = (unsigned)i;
}

NI_API int ni_ensemble_order_by_attempts(ni_ensemble *e) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  if (!e->initialised) return fail(NI_ERR_INVALID, "ensemble is not initialised (call ni_ensemble_init)");
  CU(cudaSetDevice(e->device));
  const size_t N = (size_t)e->N;
  if (!e->order_d) {
    if (int rc = dalloc(e, &e->order_d, N)) return rc;
    if (int rc = dalloc(e, &e->order_n_d, 1)) return rc;
  }
  if (int rc = need_scratch(e, (NI_ORDER_BINS + 1) * sizeof(unsigned))) return rc;
  unsigned *hist = (unsigned *)e->scratch, *mx = hist + NI_ORDER_BINS;
  CU(cudaMemsetAsync(hist, 0, (NI_ORDER_BINS + 1) * sizeof(unsigned), e->stream));
  const int grid = grid_for((long long)N);
  ni_k_order_max<<<grid, 256, 0, e->stream>>>(e->a.n_acc, e->a.n_rej, (long long)N, mx);
  unsigned mx_h = 0;
  CU(cudaMemcpyAsync(&mx_h, mx, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  int shift = 0;
  while ((mx_h >> shift) >= NI_ORDER_BINS) ++shift;
  ni_k_order_hist<<<grid, 256, 0, e->stream>>>(e->a.n_acc, e->a.n_rej, (long long)N, shift, hist);
  ni_k_order_scan<<<1, 1024, 0, e->stream>>>(hist);
  ni_k_order_scatter<<<grid, 256, 0, e->stream>>>(e->a.n_acc, e->a.n_rej, (long long)N, shift, hist, e->order_d);
  CU(cudaGetLastError());
  const unsigned long long n64 = (unsigned long long)N;
  CU(cudaMemcpyAsync(e->order_n_d, &n64, sizeof(n64), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->n_launches += 4;
  return NI_OK;
}

NI_API int ni_ensemble_get_order(ni_ensemble *e, uint32_t *order_host) {
  if (!e || !order_host) return fail(NI_ERR_INVALID, "ensemble / order_host is NULL");
  CU(cudaSetDevice(e->device));
  if (!e->order_d) {
    for (int64_t i = 0; i < e->N; ++i) order_host[i] = (uint32_t)i;
    return NI_OK;
  }
  CU(cudaMemcpyAsync(order_host, e->order_d, (size_t)e->N * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return NI_OK;
}

// Launch order of the work queue (longest-processing-time-first scheduling against the tail of a run).
NI_API int ni_ensemble_set_order(ni_ensemble *e, const uint32_t *order_host) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  if (!order_host) {  // back to 0, 1, 2, ...
    CU(cudaStreamSynchronize(e->stream));
    auto drop = [&](void *q) {
      if (!q) return;
      for (size_t k = 0; k < e->allocs.size(); ++k)
        if (e->allocs[k] == q) {
          e->allocs.erase(e->allocs.begin() + (long)k);
          break;
        }
      gfree(e, q);
    };
    drop(e->order_d);
    drop(e->order_n_d);
    e->order_d = nullptr;
    e->order_n_d = nullptr;
    return NI_OK;
  }
  const size_t N = (size_t)e->N;
  std::vector<bool> seen(N, false);
  for (size_t i = 0; i < N; ++i) {
    const uint32_t v = order_host[i];
    if (v >= N || seen[v]) return fail(NI_ERR_INVALID, "order is not a permutation of 0..N-1 (entry %zu = %u)", i, v);
    seen[v] = true;
  }
  if (!e->order_d) {
    if (int rc = dalloc(e, &e->order_d, N)) return rc;
    if (int rc = dalloc(e, &e->order_n_d, 1)) return rc;
  }
  const unsigned long long n64 = (unsigned long long)N;
  CU(cudaMemcpyAsync(e->order_d, order_host, N * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->order_n_d, &n64, sizeof(n64), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));  // (order_host and n64 are the caller's / this frame's)
  return NI_OK;
}

// Guard mode: how many guard bytes around the ensemble's device buffers have been overwritten (0 expected).
NI_API int ni_ensemble_guard_check(ni_ensemble *e, int64_t *n_buffers, int64_t *bad_bytes) {
  if (!e || !n_buffers || !bad_bytes) return fail(NI_ERR_INVALID, "ensemble / outputs are NULL");
  *n_buffers = (int64_t)e->guarded.size();
  *bad_bytes = 0;
  if (!e->guard) return NI_OK;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  std::vector<unsigned char> h(NI_GUARD_BYTES);
  for (const auto &g : e->guarded) {
    const char *zones[2] = {(const char *)g.first - NI_GUARD_BYTES, (const char *)g.first + g.second};
    for (const char *z : zones) {
      CU(cudaMemcpy(h.data(), z, NI_GUARD_BYTES, cudaMemcpyDeviceToHost));
      for (unsigned char c : h) *bad_bytes += c != NI_GUARD_FILL;
    }
  }
  return NI_OK;
}

// Stiffness probe: spectral-radius estimate of df/dy at every trajectory's current state.
NI_API int ni_ensemble_stiffness(ni_ensemble *e, int iterations, double *rho_host) {
  if (!e || !rho_host || iterations < 1 || iterations > 1000)
    return fail(NI_ERR_INVALID, "need an ensemble, an output buffer and 1..1000 iterations");
  if (!e->initialised) return fail(NI_ERR_INVALID, "ensemble is not initialised (call ni_ensemble_init)");
  CU(cudaSetDevice(e->device));
  if (int rc = need_scratch(e, (size_t)e->N * 8)) return rc;
  e->a.probe_iters = iterations;
  e->a.probe_out = (double *)e->scratch;
  int grid;
  if (e->mod.kind == NI_KIND_LARGE) {
    grid = run_grid(e);
  } else {
    long long need = (e->N + NI_BLOCK - 1) / NI_BLOCK, cap = (long long)e->sms * 16;
    grid = (int)(need < cap ? need : cap);
  }
  const int rc = launch(e, e->mod.k_init, grid, e->mod.block, e->mod.kind == NI_KIND_LARGE ? e->mod.smem : 0);
  e->a.probe_iters = 0;  // the init kernel initialises again
  e->a.probe_out = nullptr;
  if (rc) return rc;
  CU(cudaMemcpyAsync(rho_host, e->scratch, (size_t)e->N * 8, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return NI_OK;
}

NI_API int ni_ensemble_terminal_sink_bytes(const ni_ensemble *e, int64_t *bytes_per_trajectory) {
  if (!e || !bytes_per_trajectory) return fail(NI_ERR_INVALID, "ensemble / bytes_per_trajectory is NULL");
  *bytes_per_trajectory = 16 * (int64_t)ni_sink_vectors(e->mod.n, e->mod.Q);
  return NI_OK;
}

NI_API int ni_ensemble_set_terminal_sink(ni_ensemble *e, void *const *dst_device, int n_dst, int64_t first_row) {
  if (!e || n_dst < 0 || n_dst > NI_SINK_MAX || (n_dst > 0 && !dst_device) || first_row < 0)
    return fail(NI_ERR_INVALID, "need an ensemble, 0..%d destination buffers and a first row >= 0", NI_SINK_MAX);
  if (n_dst > 0 && e->mod.kind != NI_KIND_SMALL)
    return fail(NI_ERR_UNSUPPORTED, "the in-kernel terminal sink belongs to the thread-per-trajectory kernels "
                                    "(CTA-per-trajectory systems: ni_ensemble_pack_terminal_peers after the run)");
  for (int r = 0; r < n_dst; ++r)
    if (!dst_device[r] || ((uintptr_t)dst_device[r] & 15)) return fail(NI_ERR_INVALID, "destination %d is NULL or not 16-byte aligned", r);
  for (int r = 0; r < NI_SINK_MAX; ++r) e->a.sink[r] = r < n_dst ? (unsigned int *)dst_device[r] : nullptr;
  e->a.sink_n = n_dst;
  e->a.sink_row0 = first_row;
  return NI_OK;
}

NI_API int ni_ensemble_pack_terminal_peers(ni_ensemble *e, void *const *dst_device, int n_dst);
NI_API int ni_ensemble_pack_terminal(ni_ensemble *e, void *dst_device) {
  if (!e || !dst_device) return fail(NI_ERR_INVALID, "ensemble / dst_device is NULL");
  if (!e->initialised) return fail(NI_ERR_INVALID, "ensemble is not initialised (call ni_ensemble_init)");
  CU(cudaSetDevice(e->device));
  if (((uintptr_t)dst_device & 15) == 0) {  // the tiled kernel with 16-byte stores (one destination)
    void *one[1] = {dst_device};
    return ni_ensemble_pack_terminal_peers(e, one, 1);
  }
  const long long words = (long long)e->N * (2 * (1 + e->mod.n + e->mod.Q) + 3);
  ni_k_pack_terminal<<<grid_for(words), 256, 0, e->stream>>>(e->a, e->mod.kind == NI_KIND_SMALL ? 1 : 0, e->mod.n, e->mod.Q,
                                                             (unsigned *)dst_device);
  CU(cudaGetLastError());
  e->n_launches++;
  return NI_OK;
}

NI_API int ni_ensemble_pack_terminal_peers(ni_ensemble *e, void *const *dst_device, int n_dst) {
  if (!e || !dst_device || n_dst < 1 || n_dst > NI_MAX_PEERS)
    return fail(NI_ERR_INVALID, "need an ensemble and 1..%d destination buffers", NI_MAX_PEERS);
  if (!e->initialised) return fail(NI_ERR_INVALID, "ensemble is not initialised (call ni_ensemble_init)");
  CU(cudaSetDevice(e->device));
  NiPeerDst d;
  d.nd = n_dst;
  for (int r = 0; r < NI_MAX_PEERS; ++r) d.p[r] = r < n_dst ? (unsigned *)dst_device[r] : nullptr;
  for (int r = 0; r < n_dst; ++r)
    if (!d.p[r]) return fail(NI_ERR_INVALID, "destination %d is NULL", r);
  for (int r = 0; r < n_dst; ++r)
    if ((uintptr_t)d.p[r] & 15) return fail(NI_ERR_INVALID, "destination %d is not 16-byte aligned", r);
  const long long words = (long long)e->N * (2 * (1 + e->mod.n + e->mod.Q) + 3);
  (void)words;
  const long long tiles = (e->N + NI_PACK_TILE - 1) / NI_PACK_TILE;
  ni_k_pack_terminal_peers<<<(int)(tiles < 148 * 8 ? tiles : 148 * 8), 256, 0, e->stream>>>(e->a, e->mod.kind == NI_KIND_SMALL ? 1 : 0, e->mod.n,
                                                                   e->mod.Q, d);
  CU(cudaGetLastError());
  e->n_launches++;
  return NI_OK;
}

NI_API int ni_ensemble_get_terminal(ni_ensemble *e, void *host_out) {
  if (!e || !host_out) return fail(NI_ERR_INVALID, "ensemble / host_out is NULL");
  int64_t rb = 0;
  ni_ensemble_terminal_bytes(e, &rb);
  const size_t bytes = (size_t)e->N * (size_t)rb;
  CU(cudaSetDevice(e->device));
  if (e->mod.kind == NI_KIND_LARGE) {
    // A CTA-per-trajectory run kernel fills every SM completely (255 registers x 256 threads), and when a second
    // ensemble's run is already queued on another stream (double buffering) its CTAs move in as this ensemble's drain:
    // a pack KERNEL would wait behind that whole run, and with it the download and the host (measured on C5: end to end
    // 90 instead of 79 ms per step; a high-priority stream does not help, the SMs are simply taken).  The state of these
    // ensembles is trajectory-major already, so the copy engine assembles the records itself: one strided
    // device-to-host copy per field, straight into the caller's record array -- no kernel on the path.
    const size_t n = (size_t)e->mod.n, Q = (size_t)e->mod.Q, N = (size_t)e->N, pitch = (size_t)rb;
    char *out = (char *)host_out;
    const NiArgs &a = e->a;
    CU(cudaMemcpy2DAsync(out, pitch, a.t, 8, 8, N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpy2DAsync(out + 8, pitch, a.y, 8 * n, 8 * n, N, cudaMemcpyDeviceToHost, e->stream));
    if (Q) CU(cudaMemcpy2DAsync(out + 8 * (1 + n), pitch, a.aux, 8 * Q, 8 * Q, N, cudaMemcpyDeviceToHost, e->stream));
    const size_t k0 = 8 * (1 + n + Q);
    CU(cudaMemcpy2DAsync(out + k0, pitch, a.n_acc, 4, 4, N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpy2DAsync(out + k0 + 4, pitch, a.n_rej, 4, 4, N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpy2DAsync(out + k0 + 8, pitch, a.status, 4, 4, N, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return NI_OK;
  }
  if (int rc = need_scratch(e, bytes)) return rc;
  if (int rc = ni_ensemble_pack_terminal(e, e->scratch)) return rc;
  CU(cudaMemcpyAsync(host_out, e->scratch, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return NI_OK;
}

// ---- recording
NI_API int ni_ensemble_record_enable(ni_ensemble *e, int64_t capacity) {
  if (!e || capacity < 0) return fail(NI_ERR_INVALID, "need an ensemble and capacity >= 0");
  CU(cudaSetDevice(e->device));
  NiArgs &a = e->a;
  if (capacity > a.rec_cap) {
    const size_t c = (size_t)capacity;
    int rc = NI_OK;
    if (a.rec_cap > 0 || a.rec_traj) {  // growing: the old log goes (undrained records with it, as documented)
      CU(cudaStreamSynchronize(e->stream));
      void *old[] = {a.rec_traj, a.rec_t, a.rec_y, a.rec_aux};
      for (void *q : old) {
        if (!q) continue;
        for (size_t k = 0; k < e->allocs.size(); ++k)
          if (e->allocs[k] == q) {
            e->allocs.erase(e->allocs.begin() + (long)k);
            break;
          }
        gfree(e, q);
      }
      a.rec_traj = nullptr;
      a.rec_t = a.rec_y = a.rec_aux = nullptr;
      a.rec_cap = 0;
    }
    if (rc == NI_OK) rc = dalloc(e, &a.rec_traj, c);
    if (rc == NI_OK) rc = dalloc(e, &a.rec_t, c);
    if (rc == NI_OK) rc = dalloc(e, &a.rec_y, c * e->mod.n);
    if (rc == NI_OK) rc = dalloc(e, &a.rec_aux, c * e->mod.Q);
    if (rc != NI_OK) {
      a.rec_cap = 0;
      return rc;
    }
  }
  a.rec_cap = capacity;
  CU(cudaMemsetAsync(a.rec_cursor, 0, sizeof(unsigned long long), e->stream));
  return NI_OK;
}

NI_API int ni_ensemble_record_count(ni_ensemble *e, int64_t *count) {
  if (!e || !count) return fail(NI_ERR_INVALID, "ensemble / count is NULL");
  CU(cudaSetDevice(e->device));
  unsigned long long c = 0;
  CU(cudaMemcpyAsync(&c, e->a.rec_cursor, sizeof c, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *count = (int64_t)(c < (unsigned long long)e->a.rec_cap ? c : (unsigned long long)e->a.rec_cap);
  return NI_OK;
}

NI_API int ni_ensemble_record_drain(ni_ensemble *e, uint32_t *traj, uint32_t *step, double *t, double *y, double *aux,
                                    int64_t max_records, int64_t *n_out) {
  if (!e || !n_out) return fail(NI_ERR_INVALID, "ensemble / n_out is NULL");
  if (!traj) return fail(NI_ERR_INVALID, "the traj buffer is required");
  int64_t count = 0;
  if (int rc = ni_ensemble_record_count(e, &count)) return rc;
  if (count > max_records) return fail(NI_ERR_INVALID, "host buffers hold %lld records, the log has %lld", (long long)max_records, (long long)count);
  NiArgs &a = e->a;
  const int n = e->mod.n, Q = e->mod.Q, small = e->mod.kind == NI_KIND_SMALL ? 1 : 0;
  // Windows of at most `window` slots are compacted on the device (unused slots dropped, order kept) into row-major
  // staging buffers of fixed size -- the scratch does not grow with the log -- and copied out one after the other.
  const long long rec_bytes = 4 + 8 + 8ll * n + 8ll * Q;
  long long window = (256ll << 20) / rec_bytes / NI_DRAIN_BLOCK * NI_DRAIN_BLOCK;  // ~256 MB of staging
  if (window < NI_DRAIN_BLOCK) window = NI_DRAIN_BLOCK;
  if (window > count) window = (count + NI_DRAIN_BLOCK - 1) / NI_DRAIN_BLOCK * NI_DRAIN_BLOCK;
  int64_t kept = 0;
  if (count > 0) {
    const long long max_blocks = window / NI_DRAIN_BLOCK;
    const size_t off_t = ((size_t)window * 4 + 15) / 16 * 16, off_y = off_t + (size_t)window * 8,
                 off_aux = off_y + (size_t)window * 8 * n, off_cnt = off_aux + (size_t)window * 8 * Q,
                 off_tot = (off_cnt + (size_t)max_blocks * 4 + 15) / 16 * 16;
    if (int rc = need_scratch(e, off_tot + 16)) return rc;
    char *sc = (char *)e->scratch;
    unsigned *o_traj = (unsigned *)sc, *blk = (unsigned *)(sc + off_cnt);
    double *o_t = (double *)(sc + off_t), *o_y = (double *)(sc + off_y), *o_aux = (double *)(sc + off_aux);
    unsigned long long *tot_d = (unsigned long long *)(sc + off_tot);
    for (long long s0 = 0; s0 < count; s0 += window) {
      const long long cnt = count - s0 < window ? count - s0 : window;
      const int nb = (int)((cnt + NI_DRAIN_BLOCK - 1) / NI_DRAIN_BLOCK);
      ni_k_drain_count<<<nb, 256, 0, e->stream>>>(a.rec_traj, s0, cnt, blk);
      ni_k_drain_scan<<<1, 1024, 0, e->stream>>>(blk, nb, tot_d);
      ni_k_drain_pack<<<nb, 256, 0, e->stream>>>(a, small, n, Q, s0, cnt, blk, o_traj, o_t, o_y, o_aux);
      CU(cudaGetLastError());
      unsigned long long k = 0;
      CU(cudaMemcpyAsync(&k, tot_d, sizeof k, cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      if (k > 0) {
        CU(cudaMemcpyAsync(traj + kept, o_traj, (size_t)k * 4, cudaMemcpyDeviceToHost, e->stream));
        if (t) CU(cudaMemcpyAsync(t + kept, o_t, (size_t)k * 8, cudaMemcpyDeviceToHost, e->stream));
        if (y) CU(cudaMemcpyAsync(y + kept * n, o_y, (size_t)k * 8 * n, cudaMemcpyDeviceToHost, e->stream));
        if (aux && Q > 0) CU(cudaMemcpyAsync(aux + kept * Q, o_aux, (size_t)k * 8 * Q, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));  // the staging buffers are reused by the next window
        kept += (int64_t)k;
      }
    }
  }
  CU(cudaMemsetAsync(a.rec_cursor, 0, sizeof(unsigned long long), e->stream));
  if (step && kept > 0) {
    // The log stores no step numbers: the records of one trajectory appear in step order, and the last one of this
    // batch is step n_accepted of the trajectory as it stands now (no kernel is running).
    const size_t N = (size_t)e->N;
    std::vector<int> nacc(N), seen(N, 0);
    CU(cudaMemcpyAsync(nacc.data(), a.n_acc, N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (int64_t i = 0; i < kept; ++i) ++seen[traj[i]];
    for (size_t i = 0; i < N; ++i) {  // first step number of the batch, per trajectory
      nacc[i] -= seen[i];
      seen[i] = 0;
    }
    for (int64_t i = 0; i < kept; ++i) step[i] = (uint32_t)(nacc[traj[i]] + ++seen[traj[i]]);
  }
  CU(cudaStreamSynchronize(e->stream));
  *n_out = kept;
  return NI_OK;
}

NI_API int ni_ensemble_timing(ni_ensemble *e, float *ms_init, float *ms_run, int64_t *n_launches) {
  if (!e) return fail(NI_ERR_INVALID, "ensemble is NULL");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  if (ms_init) {
    *ms_init = -1.f;
    if (e->have_init_t) CU(cudaEventElapsedTime(ms_init, e->ev[0], e->ev[1]));
  }
  if (ms_run) {
    *ms_run = -1.f;
    if (e->have_run_t) CU(cudaEventElapsedTime(ms_run, e->ev[2], e->ev[3]));
  }
  if (n_launches) *n_launches = e->n_launches;
  return NI_OK;
}
