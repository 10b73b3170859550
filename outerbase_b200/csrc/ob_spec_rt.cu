/*
 * ob_spec_rt.cu -- run-time side of the terms-specialised Phi kernels: NVRTC binding,
 * module cache, geometry and launchers.  The kernels themselves are the hand-written frame
 * ob_spec_scaffold.inc with the switch bodies ob_spec.hpp generates from a compiled terms
 * table; see those two files for what runs on the GPU.
 *
 * NVRTC is bound with dlopen (like NCCL) so that the library still loads where it is
 * absent; the compiled cubin is loaded through the CUDA runtime's library API
 * (cudaLibraryLoadData / cudaLibraryGetKernel) -- no driver-API linkage.  Compiled modules
 * are cached on disk by source hash ($OB_SPEC_CACHE, default ~/.cache/outerbase_b200;
 * "0" disables) because a compile takes seconds.
 */
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <fstream>

#include "ob_device.cuh"

namespace obd {

namespace {

struct NvrtcApi {
  void* handle = nullptr;
  int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(void**) = nullptr;
  int (*CompileProgram)(void*, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(void*, size_t*) = nullptr;
  int (*GetProgramLog)(void*, char*) = nullptr;
  int (*GetCUBINSize)(void*, size_t*) = nullptr;
  int (*GetCUBIN)(void*, char*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  bool ok() const { return handle && CreateProgram && CompileProgram && GetCUBIN && GetCUBINSize; }
  static NvrtcApi& get() {
    static NvrtcApi api;
    static bool tried = false;
    if (!tried) {
      tried = true;
      const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
      for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (api.handle) break; }
      if (api.handle) {
#define OB_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
        OB_SYM(CreateProgram, "nvrtcCreateProgram");
        OB_SYM(DestroyProgram, "nvrtcDestroyProgram");
        OB_SYM(CompileProgram, "nvrtcCompileProgram");
        OB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        OB_SYM(GetProgramLog, "nvrtcGetProgramLog");
        OB_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        OB_SYM(GetCUBIN, "nvrtcGetCUBIN");
        OB_SYM(Version, "nvrtcVersion");
#undef OB_SYM
      }
    }
    return api;
  }
};

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

std::string cache_dir() {
  const char* e = getenv("OB_SPEC_CACHE");
  if (e) return std::string(e) == "0" ? std::string() : std::string(e);
  const char* home = getenv("HOME");
  if (!home || !*home) return std::string();
  return std::string(home) + "/.cache/outerbase_b200";
}

} // namespace

bool spec_compiler_available() { return NvrtcApi::get().ok(); }

/* source -> cubin for sm_100a (NVRTC), with the on-disk cache */
std::string spec_compile_source(const std::string& src, double* seconds, bool* from_cache, bool only_if_cached, bool use_cache = true) {
  if (seconds) *seconds = 0;
  if (from_cache) *from_cache = false;
  NvrtcApi& api = NvrtcApi::get();
  int vmaj = 0, vmin = 0;
  if (api.ok() && api.Version) api.Version(&vmaj, &vmin);
  char key[64];
  std::snprintf(key, sizeof key, "%016llx_%d_%d.cubin", (unsigned long long)fnv1a(src), vmaj, vmin);
  const std::string dir = use_cache ? cache_dir() : std::string();
  if (!dir.empty()) {
    std::ifstream in(dir + "/" + key, std::ios::binary);
    if (in) {
      std::string cubin((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      if (cubin.size() > 64) { if (from_cache) *from_cache = true; return cubin; }
    }
  }
  if (only_if_cached) return std::string();
  if (!api.ok()) throw std::runtime_error("libnvrtc.so.12 not found: the terms-specialised kernels need NVRTC at run time");
  const auto t0 = std::chrono::steady_clock::now();
  void* prog = nullptr;
  if (api.CreateProgram(&prog, src.c_str(), "ob_spec.cu", 0, nullptr, nullptr) != 0) throw std::runtime_error("nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const int rc = api.CompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t ls = 0;
    std::string log;
    if (api.GetProgramLogSize && api.GetProgramLogSize(prog, &ls) == 0 && ls > 1) { log.resize(ls); api.GetProgramLog(prog, &log[0]); }
    api.DestroyProgram(&prog);
    throw std::runtime_error("NVRTC could not compile the specialised kernels: " + log.substr(0, 2000));
  }
  size_t cs = 0;
  api.GetCUBINSize(prog, &cs);
  std::string cubin(cs, '\0');
  api.GetCUBIN(prog, &cubin[0]);
  api.DestroyProgram(&prog);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!dir.empty()) { /* best effort */
    mkdir(dir.c_str(), 0755);
    const std::string tmp = dir + "/" + key + ".tmp" + std::to_string((long)getpid());
    std::ofstream out(tmp, std::ios::binary);
    if (out) { out.write(cubin.data(), (std::streamsize)cubin.size()); out.close(); rename(tmp.c_str(), (dir + "/" + key).c_str()); }
  }
  return cubin;
}

std::string spec_compile_nocache(const std::string& src, double* seconds) {
  return spec_compile_source(src, seconds, nullptr, false, /*use_cache=*/false);
}

struct SpecKernels {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t ka = nullptr, kt = nullptr;
  obs::SpecOptions opt;
  int types = 0, tr_a = 0, tr_t = 0;
  size_t vec_bytes_a = 0; /* shared-memory copy of the coefficients, slot order */
  size_t smem_a_set = 0, smem_t_set = 0;
  double compile_seconds = 0;
  bool from_cache = false;
  ~SpecKernels() { if (lib) cudaLibraryUnload(lib); }
};

obs::SpecOptions spec_default_options() {
  obs::SpecOptions o;
  if (const char* e = getenv("OB_SPEC_OPTS")) { /* wa,ra,pa,cache_a,wt,rt,pt,cache_t,acc_cap,qa -- tuning only */
    int* f[] = {&o.wa, &o.ra, &o.pa, &o.cache_a, &o.wt, &o.rt, &o.pt, &o.cache_t, &o.acc_cap, &o.qa};
    int i = 0;
    for (const char* p = e; *p && i < 10; ++i) { *f[i] = atoi(p); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
  }
  return o;
}

std::shared_ptr<SpecKernels> spec_build(Ctx& c, const obt::Program& pa, const obt::Program& pt, int types, const obs::SpecOptions& opt,
                                        bool only_if_cached) {
  obs::SpecSource S = obs::generate(&pa, &pt, types, opt);
  if (!S.ok) throw std::runtime_error("specialised kernel generation failed: " + S.why);
  auto k = std::make_shared<SpecKernels>();
  k->opt = opt; k->types = types; k->tr_a = S.tr_a; k->tr_t = S.tr_t;
  k->vec_bytes_a = ((pa.nslots() * sizeof(double) + 127) / 128) * 128;
  const std::string cubin = spec_compile_source(S.src, &k->compile_seconds, &k->from_cache, only_if_cached);
  if (cubin.empty()) return nullptr;
  OB_CUDA(cudaSetDevice(c.device));
  OB_CUDA(cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  OB_CUDA(cudaLibraryGetKernel(&k->ka, k->lib, "phi_a_spec"));
  OB_CUDA(cudaLibraryGetKernel(&k->kt, k->lib, "phi_t_spec"));
  return k;
}

double spec_compile_seconds(const SpecKernels& k) { return k.from_cache ? 0.0 : k.compile_seconds; }

static void spec_fill(obs::SpecParams& p, const PhiPlan& pl, int TR) {
  p.load_src = pl.cols->load_src.p; p.col_op = pl.cols->col_op.p;
  p.scale = pl.scale; p.sq = pl.sq; p.N = pl.N;
  p.ncol = pl.cols->ncol; p.has_ops = pl.cols->has_ops ? 1 : 0;
  p.ntiles = (int)((pl.N + TR - 1) / TR);
}

static void spec_launch(Ctx& c, cudaKernel_t kern, size_t& smem_set, int grid, int threads, size_t smem, obs::SpecParams& p, const char* what) {
  if (smem > smem_set) {
    OB_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  void* args[] = {&p};
  const cudaError_t e = cudaLaunchKernel((const void*)kern, dim3(grid), dim3(threads), args, smem, c.stream);
  if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e));
  c.launches++;
}

bool spec_fits(const Ctx& c, const SpecKernels& k, int ncol) {
  const size_t a1 = 128 + (size_t)ncol * k.tr_a * 8 + 2 * (size_t)k.opt.wa * k.tr_a * 8 + k.vec_bytes_a;
  const size_t t1 = 128 + (size_t)ncol * k.tr_t * 8 + 2 * (size_t)k.tr_t * 8;
  return a1 <= c.smem_optin && t1 <= c.smem_optin;
}

void launch_phi_a_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, const PhiAArgs& a, Workspace& ws, int* grid_out) {
  if (pl.N == 0) { if (grid_out) *grid_out = 0; return; }
  if (pl.cols->nload != pl.cols->ncol) throw std::logic_error("specialised kernels take plain column tables only");
  const int TR = k.tr_a;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  p.tile_doubles = (unsigned)(p.ncol * TR);
  const size_t tile_bytes = (size_t)p.tile_doubles * 8, part_bytes = 2 * (size_t)k.opt.wa * TR * 8;
  p.nbuf = (128 + 2 * tile_bytes + part_bytes + k.vec_bytes_a <= c.smem_optin) ? 2 : 1;
  p.off_tile = 128;
  p.off_part = (unsigned)(128 + p.nbuf * tile_bytes);
  p.off_vec = (unsigned)(p.off_part + part_bytes);
  const size_t smem = p.off_vec + k.vec_bytes_a;
  if (smem > c.smem_optin) throw std::logic_error("specialised Phi a does not fit in shared memory");
  p.out = a.out; p.w = a.w; p.y = a.y; p.sd = a.sd; p.mode = a.mode;
  const int grid = std::max(1, std::min(p.ntiles, c.sms));
  if (a.mode == PHI_UPDATE) p.ssq_partial = a.ssq_partial ? a.ssq_partial : ws.ssq.ensure(c.sms);
  p.a = a.a; p.slot_term = pl.prog->slot_term.p; p.nslots = (int)pl.prog->host.nslots();
  spec_launch(c, k.ka, k.smem_a_set, grid, 32 * (k.opt.wa * k.opt.qa + 1), smem, p, "phi_a_spec");
  if (grid_out) *grid_out = grid;
}

void launch_phi_t_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, const double* w, double* out, Workspace& ws) {
  const DevProgram& pr = *pl.prog;
  const int K = (int)pr.host.K;
  if (K == 0) return;
  if (pl.N == 0) { launch_fill(c, out, K, 0.0); return; }
  if (pl.cols->nload != pl.cols->ncol) throw std::logic_error("specialised kernels take plain column tables only");
  const int TR = k.tr_t;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  p.tile_doubles = (unsigned)(p.ncol * TR);
  const size_t tile_bytes = (size_t)p.tile_doubles * 8, b_bytes = 2 * (size_t)TR * 8;
  p.nbuf = (128 + 2 * tile_bytes + b_bytes <= c.smem_optin) ? 2 : 1;
  p.off_tile = 128;
  p.off_part = (unsigned)(128 + p.nbuf * tile_bytes);
  const size_t smem = p.off_part + b_bytes;
  if (smem > c.smem_optin) throw std::logic_error("specialised Phi^T does not fit in shared memory");
  const int J = std::max(1, std::min(p.ntiles, c.sms / k.types));
  const int grid = J * k.types;
  p.win = w;
  p.nslots = (int)pr.host.nslots();
  p.partial = ws.partial.ensure((size_t)J * p.nslots);
  spec_launch(c, k.kt, k.smem_t_set, grid, 32 * (k.opt.wt + 1), smem, p, "phi_t_spec");
  launch_phi_t_reduce(c, p.partial, J, p.nslots, pr.slot_term.p, out);
}

} // namespace obd
