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
#include <cuda.h>
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

/* forget a cached module that the driver refused to load (truncated or stale file) */
static void spec_cache_evict(const std::string& src) {
  NvrtcApi& api = NvrtcApi::get();
  int vmaj = 0, vmin = 0;
  if (api.ok() && api.Version) api.Version(&vmaj, &vmin);
  char key[64];
  std::snprintf(key, sizeof key, "%016llx_%d_%d.cubin", (unsigned long long)fnv1a(src), vmaj, vmin);
  const std::string dir = cache_dir();
  if (!dir.empty()) unlink((dir + "/" + key).c_str());
}

/* source -> loaded library.  A cubin that came from the disk cache and does not load is evicted and compiled once
 * more; every other failure throws (callers of the lazily built modules catch and fall back). */
static cudaLibrary_t spec_load_library(Ctx& c, const std::string& src, bool only_if_cached, double* seconds, bool* from_cache) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    bool cached = false;
    const std::string cubin = spec_compile_source(src, seconds, &cached, only_if_cached);
    if (from_cache) *from_cache = cached;
    if (cubin.empty()) return nullptr;
    OB_CUDA(cudaSetDevice(c.device));
    cudaLibrary_t lib = nullptr;
    const cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) return lib;
    (void)cudaGetLastError();
    if (!cached || attempt == 1) throw CudaError(std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
    spec_cache_evict(src);
  }
  return nullptr;
}

struct SpecKernels {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t ka = nullptr, kt = nullptr;
  obs::SpecOptions opt;
  int types = 0, tr_a = 0, tr_t = 0, maxcols_t = 0, cluster = 1;
  size_t vec_bytes_a = 0; /* shared-memory copy of the coefficients, slot order */
  size_t smem_a_set = 0, smem_t_set = 0;
  int max_clusters = 0;
  bool fuse_ok = true; /* phi_t_spec may reduce in its own tail (cooperative launch accepted so far) */
  /* multi right-hand-side module (phi_am_spec), built at first use */
  cudaLibrary_t libm = nullptr;
  cudaKernel_t km = nullptr;
  int nblk = 0, mat_state = 0; /* 0 not built, 1 ready, -1 failed */
  size_t smem_m_set = 0;
  DevBuf<double> aperm;
  /* transposed multi right-hand-side module (phi_tm_spec), built at first use */
  cudaLibrary_t libtm = nullptr;
  cudaKernel_t ktm = nullptr;
  int tm_state = 0, tm_maxcols = 0; /* 0 not built, 1 ready, -1 failed */
  size_t smem_tm_set = 0;
  DevBuf<double> tm_partial, tm_stage;
  /* hyper-gradient sweep module (phi_d_spec), built at first use */
  cudaLibrary_t libd = nullptr;
  cudaKernel_t kd = nullptr;
  int dot_state = 0, tr_d = 0; /* 0 not built, 1 ready, -1 failed (too many columns for the register accumulators) */
  size_t smem_d_set = 0;
  DevBuf<double> dpart;
  obt::Program pa_host; /* copy of the G = 1 program the module is generated from */
  double compile_seconds = 0;
  bool from_cache = false;
  std::string lazy_why; /* why a lazily built module (phi_am_spec, phi_d_spec) is unavailable */
  ~SpecKernels() { if (lib) cudaLibraryUnload(lib); if (libm) cudaLibraryUnload(libm); if (libd) cudaLibraryUnload(libd); if (libtm) cudaLibraryUnload(libtm); }
};

obs::SpecOptions spec_default_options() {
  obs::SpecOptions o;
  if (const char* e = getenv("OB_SPEC_OPTS")) { /* ra,qa,tga,cache_a,wt,rt,pt,cache_t,acc_cap,np,mc,ut,mw,kc,qd,tgd,cache_d,maxcols_d,nreg_c,nreg_p,psleep,tmap -- tuning only */
    int* f[] = {&o.ra, &o.qa, &o.tga, &o.cache_a, &o.wt, &o.rt, &o.pt, &o.cache_t, &o.acc_cap, &o.np, &o.mc, &o.ut, &o.mw, &o.kc,
                &o.qd, &o.tgd, &o.cache_d, &o.maxcols_d, &o.nreg_c, &o.nreg_p, &o.psleep, &o.tmap};
    int i = 0;
    for (const char* p = e; *p && i < 22; ++i) { *f[i] = atoi(p); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
  }
  return o;
}

std::shared_ptr<SpecKernels> spec_build(Ctx& c, const obt::Program& pa, const obt::Program& pt, int types, const obs::SpecOptions& opt,
                                        bool only_if_cached) {
  obs::SpecSource S = obs::generate(&pa, &pt, types, opt);
  if (!S.ok) throw std::runtime_error("specialised kernel generation failed: " + S.why);
  auto k = std::make_shared<SpecKernels>();
  k->opt = opt; k->types = types; k->tr_a = S.tr_a; k->tr_t = S.tr_t; k->maxcols_t = S.maxcols_t; k->cluster = S.cluster;
  k->vec_bytes_a = ((pa.nslots() * sizeof(double) + 127) / 128) * 128;
  k->pa_host = pa;
  k->lib = spec_load_library(c, S.src, only_if_cached, &k->compile_seconds, &k->from_cache);
  if (!k->lib) return nullptr;
  OB_CUDA(cudaLibraryGetKernel(&k->ka, k->lib, "phi_a_spec"));
  OB_CUDA(cudaLibraryGetKernel(&k->kt, k->lib, "phi_t_spec"));
  return k;
}

bool spec_uses_cluster(const SpecKernels& k) { return k.cluster > 1; }

bool device_range_readable(const void* p, size_t bytes) {
  using Fn = int (*)(unsigned long long*, size_t*, unsigned long long);
  static Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<Fn>(f);
  }
  if (!fn) return false;
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return false;
  const unsigned long long lo = (unsigned long long)(uintptr_t)p;
  return lo >= base && lo + bytes <= base + size;
}
double spec_compile_seconds(const SpecKernels& k) { return k.from_cache ? 0.0 : k.compile_seconds; }

/* ---- 2-D tensor maps of phi_a_spec's tile (option tmap).  The tile's columns are ordered by (dimension, level); the
 * levels 1..max of a dimension are adjacent columns of the (compact) basis matrix, so each such run is ONE tiled tensor
 * copy of TR rows x run-length columns instead of run-length bulk copies.  Encoded through the driver entry point (no
 * link against libcuda); any table whose runs are not evenly strided in memory keeps the bulk copies. */
using TensorMap128 = CUtensorMap;
static_assert(sizeof(CUtensorMap) == 128, "the kernel indexes the maps 128 bytes apart");
static bool encode_tile_map(TensorMap128* out, const double* base, u64 rows, u64 width, u64 stride_bytes, int TR) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<Fn>(f);
  }
  if (!fn) return false;
  const cuuint64_t gdim[2] = {rows, width};
  const cuuint64_t gstride[1] = {stride_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)TR, (cuuint32_t)width};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2u, const_cast<double*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static const void* tile_tensor_maps(Ctx& c, const SpecKernels& k, const PhiPlan& pl, int TR) {
  const ColTable& ct = *pl.cols;
  if (ct.tmap_state == 1 && ct.tmap_tr == TR) return ct.tmaps.p;
  if (ct.tmap_state == -1) return nullptr;
  ct.tmap_state = -1;
  const obt::Program& P = k.pa_host;
  if (pl.ld == 0 || pl.ld % (u64)TR || pl.N >= (1ull << 31) || ct.src_host.size() < P.cols.size() || ct.nload != ct.ncol) return nullptr;
  if (!k.opt.tmap) return nullptr; /* the module's tile is in the program's column order */
  const std::vector<int> order = obs::tile_order(P), runs = obs::tile_runs(P, order);
  if (runs.empty() || runs.size() > 31) return nullptr;
  std::vector<TensorMap128> maps(runs.size());
  for (size_t r = 0; r < runs.size(); ++r) {
    const int c0 = runs[r], c1 = r + 1 < runs.size() ? runs[r + 1] : (int)order.size();
    const double* base = ct.src_host[order[c0]];
    for (int cc = c0 + 1; cc < c1; ++cc) /* adjacent columns, ld rows apart */
      if (ct.src_host[order[cc]] != base + (size_t)(cc - c0) * pl.ld) return nullptr;
    if ((uintptr_t)base % 16 || c1 - c0 > 256) return nullptr;
    if (!encode_tile_map(&maps[r], base, pl.ld, (u64)(c1 - c0), pl.ld * sizeof(double), TR)) return nullptr;
  }
  ct.tmaps.ensure(maps.size() * sizeof(TensorMap128));
  OB_CUDA(cudaMemcpyAsync(ct.tmaps.p, maps.data(), maps.size() * sizeof(TensorMap128), cudaMemcpyHostToDevice, c.stream));
  OB_CUDA(cudaStreamSynchronize(c.stream)); /* `maps` is a local */
  ct.tmap_tr = TR;
  ct.tmap_state = 1;
  return ct.tmaps.p;
}

static void spec_fill(obs::SpecParams& p, const PhiPlan& pl, int TR) {
  p.load_src = pl.cols->load_src.p; p.col_op = pl.cols->col_op.p;
  p.scale = pl.scale; p.sq = pl.sq; p.N = pl.N;
  p.ncol = pl.cols->ncol; p.has_ops = pl.cols->has_ops ? 1 : 0;
  p.ntiles = (int)((pl.N + TR - 1) / TR);
}

static cudaError_t spec_launch_try(Ctx& c, cudaKernel_t kern, int grid, int threads, size_t smem, obs::SpecParams& p, bool cooperative) {
  void* args[] = {&p};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = c.stream;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeCooperative;
  at.val.cooperative = 1;
  cfg.attrs = &at; cfg.numAttrs = cooperative ? 1 : 0;
  return cudaLaunchKernelExC(&cfg, (const void*)kern, args);
}

static void spec_launch(Ctx& c, cudaKernel_t kern, size_t& smem_set, int grid, int threads, size_t smem, obs::SpecParams& p, const char* what,
                        int cluster = 1) {
  if (smem > smem_set) {
    OB_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  void* args[] = {&p};
  cudaError_t e;
  if (cluster > 1) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = c.stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = cluster; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
  } else e = cudaLaunchKernel((const void*)kern, dim3(grid), dim3(threads), args, smem, c.stream);
  if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e));
  c.launches++;
}

/* shared-memory layout: 16 mbarriers | coefficient copy (Phi a) | nstage tiles of (ncol + extra) columns */
struct SpecGeom { int nstage = 0; unsigned off_vec = 0, off_tile = 0, tile_doubles = 0; size_t smem = 0; };
/* `group` = consumer groups that take the tiles round-robin in a kernel WITHOUT the rounds[] guard (phi_d_spec): the
 * stage count is then a multiple of it, so that every round of one stage goes to the SAME group -- a group that met a
 * stage of another group's could test its full barrier two phases late, and an mbarrier parity wait cannot tell phase
 * r from phase r + 2 (seen as wrong rows / a launch failure once in a few runs with 4 groups on 5 or 6 stages).
 * phi_a_spec keeps a per-stage count of completed rounds in shared memory instead (ob_spec_scaffold.inc) and takes
 * any stage count. */
static SpecGeom spec_geometry(const Ctx& c, int ncol, int nextra, int TR, size_t vec_bytes, int want_stages, int group = 1) {
  SpecGeom g;
  g.off_vec = 128 + 256 + 128; /* 16 mbarriers | 256 square flags | completed rounds per stage | ... */
  g.off_tile = (unsigned)(g.off_vec + vec_bytes);
  g.tile_doubles = (unsigned)((ncol + nextra) * TR);
  const size_t tile_bytes = (size_t)g.tile_doubles * 8;
  const size_t room = c.smem_optin > g.off_tile ? c.smem_optin - g.off_tile : 0;
  g.nstage = (int)std::min<size_t>({(size_t)want_stages, room / std::max<size_t>(tile_bytes, 1), (size_t)8});
  g.nstage -= g.nstage % std::max(group, 1);
  g.smem = g.off_tile + (size_t)g.nstage * tile_bytes;
  return g;
}

/* shrink the tile geometry until a table with `ncol` columns and `nslots` coefficient slots fits the
 * shared memory of the SM: fewer tiles in work first, then fewer rows per tile */
obs::SpecOptions spec_adapt_options(const Ctx& c, obs::SpecOptions o, int ncol, size_t nslots) {
  /* Phi a: four 64-row tiles in work (2 warps each) want a fifth stage to load ahead; a table whose tiles leave room
   * for four only (C4: 81 columns, 4000 coefficients) runs 10 % faster on two 128-row tiles of 4 warps
   * (profiles/r02_spec_sweeps.txt) */
  if (o.ra == 1 && o.qa == 2 && o.tga == 4 && !getenv("OB_SPEC_OPTS")) {
    const size_t vec = std::max<size_t>(((nslots * 8 + 127) / 128) * 128, 8 * 32 * (size_t)(o.qa * o.tga + o.np));
    if (spec_geometry(c, ncol, 1, 64, vec, 8).nstage < 5 && spec_geometry(c, ncol, 1, 128, vec, 8).nstage >= 2) { o.qa = 4; o.tga = 2; }
  }
  for (;;) {
    const size_t vec = std::max<size_t>(((nslots * 8 + 127) / 128) * 128, 8 * 32 * (size_t)(o.qa * o.tga + o.np));
    if (spec_geometry(c, ncol, 1, 32 * o.ra * o.qa, vec, 8).nstage >= o.tga) break;
    if (o.tga > 1) --o.tga;
    else if (o.qa > 1) o.qa /= 2;
    else if (o.ra > 1) o.ra /= 2;
    else break;
  }
  while (spec_geometry(c, ncol, 2, 32 * o.rt * o.pt, 0, 8).nstage < 2 && o.pt > 1) o.pt /= 2;
  return o;
}

bool spec_fits(const Ctx& c, const SpecKernels& k, int ncol) {
  if (ncol + 2 > 256) return false; /* one square flag per tile column in a 256-byte block */
  return spec_geometry(c, ncol, 1, k.tr_a, std::max<size_t>(k.vec_bytes_a, 8 * 32 * (k.opt.qa * k.opt.tga + k.opt.np)), 8).nstage >= k.opt.tga &&
         spec_geometry(c, k.maxcols_t, 2, k.tr_t, 0, 8).nstage >= 1;
}

void launch_phi_a_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, const PhiAArgs& a, Workspace& ws, int* grid_out) {
  if (pl.N == 0) { if (grid_out) *grid_out = 0; return; }
  if (pl.cols->nload != pl.cols->ncol) throw std::logic_error("specialised kernels take plain column tables only");
  const int TR = k.tr_a, warps = k.opt.qa * k.opt.tga;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  /* the coefficient copy doubles as the scratch of the final residual reduction (one double per thread) */
  const SpecGeom g = spec_geometry(c, p.ncol, 1, TR, std::max<size_t>(k.vec_bytes_a, 8 * 32 * (warps + k.opt.np)), 8);
  if (g.nstage < k.opt.tga) throw std::logic_error("specialised Phi a does not fit in shared memory");
  p.nstage = g.nstage; p.off_vec = g.off_vec; p.off_tile = g.off_tile; p.tile_doubles = g.tile_doubles; p.off_flags = 128;
  p.out = a.out; p.w = a.w; p.y = a.y; p.sd = a.sd; p.mode = a.mode;
  const int grid = std::max(1, std::min(p.ntiles, c.sms));
  if (a.mode == PHI_UPDATE || a.mode == PHI_DOT) p.ssq_partial = a.ssq_partial ? a.ssq_partial : ws.ssq.ensure(c.sms);
  if (a.mode == PHI_DOT) { p.win = a.wdot; p.yh = a.yh; }
  p.a = a.a; p.slot_term = pl.prog->slot_term.p; p.nslots = (int)pl.prog->host.nslots();
  p.tmaps = (k.opt.tmap && c.tmap) ? tile_tensor_maps(c, k, pl, TR) : nullptr;
  spec_launch(c, k.ka, k.smem_a_set, grid, 32 * (warps + k.opt.np), g.smem, p, "phi_a_spec");
  if (grid_out) *grid_out = grid;
}

void launch_phi_t_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, const double* w, double* out, Workspace& ws, const unsigned* ready) {
  const DevProgram& pr = *pl.prog;
  const int K = (int)pr.host.K;
  if (K == 0) return;
  if (pl.N == 0) { launch_fill(c, out, K, 0.0); return; }
  if (pl.cols->nload != pl.cols->ncol || pl.cols->has_ops) throw std::logic_error("specialised Phi^T takes stored columns only (no per-column ops)");
  if (pl.N >= (1ull << 31)) throw std::range_error("specialised Phi^T: more than 2^31 - 1 rows per GPU");
  const int TR = k.tr_t;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  const SpecGeom g = spec_geometry(c, k.maxcols_t, 2, TR, 0, 4);
  if (g.nstage < 1) throw std::logic_error("specialised Phi^T does not fit in shared memory");
  p.nstage = g.nstage; p.off_vec = g.off_vec; p.off_tile = g.off_tile; p.tile_doubles = g.tile_doubles; p.off_flags = 128;
  int jmax = c.sms / k.types;
  if (k.cluster > 1) { /* clusters that can be resident at once: a second, partial wave would double the time */
    if (k.max_clusters == 0 || g.smem > k.smem_t_set) {
      if (g.smem > k.smem_t_set) {
        OB_CUDA(cudaFuncSetAttribute((const void*)k.kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        k.smem_t_set = g.smem;
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(c.sms / k.cluster * k.cluster); cfg.blockDim = dim3(32 * (k.opt.wt + k.opt.np)); cfg.dynamicSmemBytes = g.smem;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = k.cluster; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      int nc = 0;
      OB_CUDA(cudaOccupancyMaxActiveClusters(&nc, (const void*)k.kt, &cfg));
      if (nc < 1) throw std::logic_error("the Phi^T cluster does not fit on this GPU");
      k.max_clusters = nc;
    }
    jmax = std::min(jmax, k.max_clusters);
  }
  const int J = std::max(1, std::min(p.ntiles, jmax));
  const int grid = J * k.types;
  p.win = w;
  p.ready = k.cluster == 1 ? ready : nullptr;
  if (ready && k.cluster != 1) throw std::logic_error("streamed input vector: not with the cluster variant");
  p.nslots = (int)pr.host.nslots();
  p.partial = ws.partial.ensure((size_t)J * p.nslots);
  /* The cross-CTA sum (and, when the caller's allreduce rides along, the cross-GPU sum) runs in the tail of the same
   * launch: needs every CTA resident at once, which a cooperative launch guarantees.  OB_FUSE_TAIL=0 or a refused
   * cooperative launch (SMs held by somebody else) fall back to the separate reduction kernel. */
  static const bool fuse_env = !(getenv("OB_FUSE_TAIL") && std::string(getenv("OB_FUSE_TAIL")) == "0");
  if (fuse_env && k.cluster == 1 && k.fuse_ok && grid <= c.sms) {
    const int nextra = c.fuse_extra && c.fuse_extra_n > 0 ? 1 : 0;
    const bool ranks = c.fuse_n == (size_t)(K + nextra) && c.p2p_ok((size_t)(K + nextra));
    p.extra = nextra ? c.fuse_extra : nullptr; p.extra_n = nextra ? c.fuse_extra_n : 0;
    if (g.smem > k.smem_t_set) {
      OB_CUDA(cudaFuncSetAttribute((const void*)k.kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
      k.smem_t_set = g.smem;
    }
    p.out = out; p.slot_term = pr.slot_term.p;
    p.sync_ctr = c.grid_sync_counter();
    p.sync_target = c.sync_count + (unsigned)grid;
    p.fuse_tail = 1; p.G = 1; p.rank = 0; p.seq = 0;
    Ctx::P2PCall call{};
    if (ranks) {
      call = c.p2p_next_call();
      for (int r = 0; r < call.G; ++r) p.peer[r] = call.slots[r];
      p.G = call.G; p.rank = call.rank; p.seq = call.seq; p.timeout_ns = call.timeout_ns;
    }
    const cudaError_t e = spec_launch_try(c, k.kt, grid, 32 * (k.opt.wt + k.opt.np), g.smem, p, /*cooperative=*/true);
    if (e == cudaSuccess) {
      c.sync_count += (unsigned)grid;
      c.launches++;
      if (nextra) c.extra_done = true;
      if (ranks) { c.fuse_n = 0; c.fused = true; }
      return;
    }
    (void)cudaGetLastError();
    if (ranks) throw CudaError(std::string("phi_t_spec (fused tail): ") + cudaGetErrorString(e)); /* the sequence number is spent */
    k.fuse_ok = false; /* this GPU cannot hold the grid at once (shared with other work): separate reduction from now on */
    p.fuse_tail = 0; p.extra = nullptr; p.extra_n = 0;
  }
  spec_launch(c, k.kt, k.smem_t_set, grid, 32 * (k.opt.wt + k.opt.np), g.smem, p, "phi_t_spec", k.cluster);
  launch_phi_t_reduce(c, p.partial, J, p.nslots, pr.slot_term.p, out);
}

bool launch_phi_am_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, const double* A, u64 C, double* out, u64 ldo) {
  if (pl.N == 0 || C == 0) return true;
  if (pl.cols->nload != pl.cols->ncol) return false;
  if (k.mat_state == 0) {
    k.mat_state = -1;
    obs::SpecSource S = obs::generate_mat(k.pa_host, k.opt);
    if (!S.ok) return false;
    try { /* a failed build leaves state -1: the caller runs the column loop, now and later */
      k.libm = spec_load_library(c, S.src, false, nullptr, nullptr);
      OB_CUDA(cudaLibraryGetKernel(&k.km, k.libm, "phi_am_spec"));
    } catch (const std::exception& ex) {
      k.lazy_why = std::string("phi_am_spec: ") + ex.what();
      return false;
    }
    k.nblk = S.nacc;
    k.mat_state = 1;
  }
  if (k.mat_state != 1) return false;
  const int MW = k.opt.mw, KC = k.opt.kc, TR = 32 * MW;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  obs::MatParams q{};
  p.off_flags = 128;
  q.off_coef = 128 + 256;
  q.off_phi = q.off_coef + 2 * KC * 64 * 8;
  p.off_tile = q.off_phi + MW * KC * 36 * 8; /* OBS_PHI_STRIDE */
  p.tile_doubles = (unsigned)((p.ncol + 1) * TR);
  p.nstage = 1;
  const size_t smem = p.off_tile + (size_t)p.tile_doubles * 8;
  if (smem > c.smem_optin) return false;
  const DevProgram& pr = *pl.prog;
  const int nslots = (int)pr.host.nslots();
  k.aperm.ensure((size_t)k.nblk * KC * 64);
  const int grid = std::max(1, std::min(p.ntiles, c.sms));
  for (u64 c0 = 0; c0 < C; c0 += 64) {
    const int nc = (int)std::min<u64>(64, C - c0);
    launch_gather_coef_blocks(c, A, pr.host.K, c0, nc, pr.slot_term.p, nslots, k.nblk * KC, k.aperm.p);
    q.aperm = k.aperm.p; q.out = out + c0 * ldo; q.ldo = ldo; q.ncols = nc; q.nblocks = k.nblk;
    if (smem > k.smem_m_set) {
      OB_CUDA(cudaFuncSetAttribute((const void*)k.km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k.smem_m_set = smem;
    }
    void* args[] = {&p, &q};
    const cudaError_t e = cudaLaunchKernel((const void*)k.km, dim3(grid), dim3(32 * MW), args, smem, c.stream);
    if (e != cudaSuccess) throw CudaError(std::string("phi_am_spec: ") + cudaGetErrorString(e));
    c.launches++;
  }
  return true;
}

/* out(term(slot), c) = sum over the J row groups of partial[(j * nslots + slot) * 64 + c], fixed order */
__global__ void phi_tm_reduce_kernel(const double* __restrict__ partial, int J, int nslots, const int* __restrict__ slot_term, int ncols,
                                     unsigned long long K, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x; /* one thread per (slot, column) */
  if (idx >= nslots * 64) return;
  const int i = idx >> 6, c = idx & 63;
  const int t = slot_term[i];
  if (t < 0 || c >= ncols) return;
  double s = 0.0;
  for (int j = 0; j < J; ++j) s += partial[((size_t)j * nslots + i) * 64 + c];
  out[(unsigned long long)t + (unsigned long long)c * K] = s;
}

/* rows [0, N) of `ncols` columns into a buffer of leading dimension ld whose pad rows are zero (the tensor-core kernel
 * multiplies pad rows by Phi = 0: they must be finite) */
__global__ void pad_columns_kernel(const double* __restrict__ A, unsigned long long lda, unsigned long long N, unsigned long long ld, int ncols,
                                   double* __restrict__ out) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ld * (unsigned long long)ncols) return;
  const unsigned long long n = idx % ld, c = idx / ld;
  out[idx] = n < N ? A[n + c * lda] : 0.0;
}

/* Phi^T . A for C >= 8 columns as dense contractions on the FP64 tensor cores (phi_tm_spec, 64 columns per launch).
 * pl.prog = the G = types * 8 program with streams of <= 32 terms.  out: K x C, column-major (local sums: the caller
 * adds the ranks).  false: not available (module build failed, geometry) -- the caller runs the column loop. */
bool launch_phi_tm_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, int types, const double* A, u64 lda, u64 C, double* out) {
  const DevProgram& pr = *pl.prog;
  const u64 K = pr.host.K;
  if (K == 0 || C == 0) return true;
  if (pl.cols->nload != pl.cols->ncol || pl.cols->has_ops || pl.N >= (1ull << 31)) return false;
  if (pl.N == 0) { launch_fill(c, out, K * C, 0.0); return true; }
  if (k.tm_state == 0) {
    k.tm_state = -1;
    obs::SpecSource S = obs::generate_tmat(pr.host, types, k.opt);
    if (!S.ok) { k.lazy_why = "phi_tm_spec: " + S.why; return false; }
    try {
      k.libtm = spec_load_library(c, S.src, false, nullptr, nullptr);
      OB_CUDA(cudaLibraryGetKernel(&k.ktm, k.libtm, "phi_tm_spec"));
    } catch (const std::exception& ex) {
      k.lazy_why = std::string("phi_tm_spec: ") + ex.what();
      return false;
    }
    k.tm_maxcols = S.maxcols_t;
    k.tm_state = 1;
  }
  if (k.tm_state != 1) return false;
  constexpr int TR = 64, W = obs::kTmWarps, T = obs::kTmTerms, PS = 36, RS = 68, NP = 4;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  obs::TMatParams q{};
  p.off_flags = 128;
  q.off_park = 128 + 256;
  p.off_tile = q.off_park + (unsigned)(W * T * PS * 8);
  q.off_r = (unsigned)((k.tm_maxcols + 1) * TR * 8);
  p.tile_doubles = (unsigned)((k.tm_maxcols + 1) * TR + 64 * RS);
  const size_t stage_bytes = (size_t)p.tile_doubles * 8;
  const size_t room = c.smem_optin > p.off_tile ? c.smem_optin - p.off_tile : 0;
  p.nstage = (int)std::min<size_t>({(size_t)4, room / stage_bytes, (size_t)8});
  if (p.nstage < 1) return false;
  const size_t smem = p.off_tile + (size_t)p.nstage * stage_bytes;
  const int J = std::max(1, std::min(p.ntiles, c.sms / types));
  const int grid = J * types;
  const int nslots = (int)pr.host.nslots();
  const u64 ld = ((pl.N + 255) / 256) * 256;
  q.partial = k.tm_partial.ensure((size_t)J * nslots * 64);
  q.nslots = nslots;
  if (smem > k.smem_tm_set) {
    OB_CUDA(cudaFuncSetAttribute((const void*)k.ktm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k.smem_tm_set = smem;
  }
  for (u64 c0 = 0; c0 < C; c0 += 64) {
    const int nc = (int)std::min<u64>(64, C - c0);
    /* the tile reads whole 64-row blocks of every column: stage unpadded / unaligned callers into a zero-padded copy */
    const double* Ac = A + c0 * lda;
    u64 ldc = lda;
    if (lda < ld || (reinterpret_cast<uintptr_t>(Ac) & 15u) || (lda & 1u)) {
      k.tm_stage.ensure(ld * 64);
      const u64 n = ld * (u64)nc;
      pad_columns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(Ac, lda, pl.N, ld, nc, k.tm_stage.p);
      c.launches++;
      Ac = k.tm_stage.p; ldc = ld;
    }
    q.A = Ac; q.lda = ldc; q.ncols = nc;
    void* args[] = {&p, &q};
    const cudaError_t e = cudaLaunchKernel((const void*)k.ktm, dim3(grid), dim3(32 * (W + NP)), args, smem, c.stream);
    if (e != cudaSuccess) throw CudaError(std::string("phi_tm_spec: ") + cudaGetErrorString(e));
    c.launches++;
    phi_tm_reduce_kernel<<<(nslots * 64 + 255) / 256, 256, 0, c.stream>>>(q.partial, J, nslots, pr.slot_term.p, nc, K, out + c0 * K);
    const cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) throw CudaError(std::string("phi_tm_reduce_kernel: ") + cudaGetErrorString(e2));
    c.launches++;
  }
  return true;
}

/* out[h] = sum over CTAs of partial[h * n + cta], fixed order */
__global__ void sum_rows_kernel(const double* __restrict__ partial, int rows, int n, double* __restrict__ out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= rows) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += partial[(size_t)h * n + i];
  out[h] = s;
}

/* gradhyp[h] = sum_n w_n * d(Phi a)[n]/d hyp_h for every hyper-parameter from ONE sweep over the rows (phi_d_spec,
 * ob_spec_scaffold.inc): prodmmge_'s outge (src/linalg.cpp:225-277) contracted with the row weights, never stored.
 * false: this table / model cannot use the kernel (too many basis columns for the register accumulators, more than
 * 64 hyper-parameters or 32 dimensions, a column table with ops) -- the caller falls back to one product per hyper. */
bool launch_phi_d_spec(Ctx& c, SpecKernels& k, const PhiPlan& pl, const double* a, const double* w, const DotArgs& g, double* out) {
  if (g.H == 0) return true;
  if (g.H > 64 || g.d > 32) return false;
  if (pl.cols->nload != pl.cols->ncol || pl.cols->has_ops) return false; /* stored columns only (basemat / basematsq) */
  if (k.dot_state == 0) {
    k.dot_state = -1;
    obs::SpecSource S = obs::generate_dot(k.pa_host, k.opt);
    if (!S.ok) return false;
    try { /* a failed build leaves state -1: the caller runs one product per hyper-parameter, now and later */
      k.libd = spec_load_library(c, S.src, false, nullptr, nullptr);
      OB_CUDA(cudaLibraryGetKernel(&k.kd, k.libd, "phi_d_spec"));
    } catch (const std::exception& ex) {
      k.lazy_why = std::string("phi_d_spec: ") + ex.what();
      return false;
    }
    k.tr_d = S.tr_a;
    k.dot_state = 1;
  }
  if (k.dot_state != 1) return false;
  if (pl.N == 0) { launch_fill(c, out, (u64)g.H, 0.0); return true; }
  const int TR = k.tr_d, warps = k.opt.qd * k.opt.tgd;
  obs::SpecParams p{};
  spec_fill(p, pl, TR);
  obs::DotParams q{};
  const size_t vec_bytes = ((k.pa_host.K * sizeof(double) + 127) / 128) * 128;
  const size_t hsm_bytes = (((size_t)warps * g.H * sizeof(double) + 127) / 128) * 128;
  p.off_flags = 128;
  p.off_vec = 128 + 256;
  q.off_hsm = (unsigned)(p.off_vec + vec_bytes);
  p.off_tile = (unsigned)(q.off_hsm + hsm_bytes);
  p.tile_doubles = (unsigned)((p.ncol + 1) * TR);
  const size_t tile_bytes = (size_t)p.tile_doubles * 8;
  const size_t room = c.smem_optin > p.off_tile ? c.smem_optin - p.off_tile : 0;
  p.nstage = (int)std::min<size_t>({(size_t)std::max(2 * k.opt.tgd, 3), room / tile_bytes, (size_t)8});
  p.nstage -= p.nstage % k.opt.tgd; /* every round of a stage to the same consumer group: see spec_geometry */
  if (p.nstage < k.opt.tgd) return false;
  const size_t smem = p.off_tile + (size_t)p.nstage * tile_bytes;
  p.a = a;
  const int grid = std::max(1, std::min(p.ntiles, c.sms));
  q.gmat = g.gmat; q.bmat = g.bmat; q.wdot = w; q.ld = g.ld; q.H = g.H; q.d = g.d;
  for (int l = 0; l <= g.d; ++l) { q.hst[l] = (int)g.hypst[l]; q.kst[l] = (int)g.knotptst[l]; }
  for (int h = 0; h <= g.H; ++h) q.gest[h] = (int)g.gest[h];
  q.partial = k.dpart.ensure((size_t)g.H * grid);
  if (smem > k.smem_d_set) {
    OB_CUDA(cudaFuncSetAttribute((const void*)k.kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k.smem_d_set = smem;
  }
  void* args[] = {&p, &q};
  const cudaError_t e = cudaLaunchKernel((const void*)k.kd, dim3(grid), dim3(32 * (warps + k.opt.npd)), args, smem, c.stream);
  if (e != cudaSuccess) throw CudaError(std::string("phi_d_spec: ") + cudaGetErrorString(e));
  c.launches++;
  sum_rows_kernel<<<(g.H + 63) / 64, 64, 0, c.stream>>>(q.partial, g.H, grid, out);
  const cudaError_t e2 = cudaGetLastError();
  if (e2 != cudaSuccess) throw CudaError(std::string("sum_rows_kernel: ") + cudaGetErrorString(e2));
  c.launches++;
  return true;
}

} // namespace obd
