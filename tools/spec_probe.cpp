/* spec_probe -- offline check of the terms-specialised kernels: generate the source for a
 * terms table, compile it with NVRTC for sm_100a (no GPU needed) and report ptxas' resource
 * usage.  usage: spec_probe terms.bin out_prefix [ra qa tga cache_a wt rt pt cache_t acc_cap]
 * terms.bin = u64 K, u64 d, then K*d u64 column-major. */
#include <nvrtc.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../outerbase_b200/csrc/ob_spec.hpp"

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage\n"); return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  uint64_t K, d;
  in.read((char*)&K, 8); in.read((char*)&d, 8);
  std::vector<uint64_t> terms(K * d);
  in.read((char*)terms.data(), K * d * 8);
  obs::SpecOptions o;
  int* f[] = {&o.ra, &o.qa, &o.tga, &o.cache_a, &o.wt, &o.rt, &o.pt, &o.cache_t, &o.acc_cap, &o.np, &o.mc, &o.ut};
  for (int i = 0; i < 12 && 3 + i < argc; ++i) *f[i] = std::atoi(argv[3 + i]);
  auto t0 = std::chrono::steady_clock::now();
  const int types = obs::choose_types(terms.data(), K, d, o);
  if (!types) { std::fprintf(stderr, "terms table is not trie-compilable\n"); return 1; }
  const obt::Program pa = obt::compile(terms.data(), K, d, 1), pt = obt::compile(terms.data(), K, d, types * o.wt);
  obs::SpecSource S = obs::generate(&pa, &pt, types, o);
  auto t1 = std::chrono::steady_clock::now();
  if (!S.ok) { std::fprintf(stderr, "generate failed: %s\n", S.why.c_str()); return 1; }
  std::string pre = argv[2];
  { std::ofstream f2(pre + ".cu"); f2 << S.src; }
  std::printf("generated %zu bytes in %.3f s; types %d nacc %d; prog_a words %zu prog_t words %zu\n", S.src.size(),
              std::chrono::duration<double>(t1 - t0).count(), S.types, S.nacc, pa.bwd.size(), pt.fwd.size());
  nvrtcProgram prog;
  nvrtcCreateProgram(&prog, S.src.c_str(), "spec.cu", 0, nullptr, nullptr);
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--ptxas-options=-v"};
  auto t2 = std::chrono::steady_clock::now();
  nvrtcResult rc = nvrtcCompileProgram(prog, 4, opts);
  auto t3 = std::chrono::steady_clock::now();
  size_t ls = 0;
  nvrtcGetProgramLogSize(prog, &ls);
  std::string log(ls, 0);
  nvrtcGetProgramLog(prog, log.data());
  std::printf("nvrtc rc=%d in %.2f s\n%s\n", (int)rc, std::chrono::duration<double>(t3 - t2).count(), log.c_str());
  if (rc != NVRTC_SUCCESS) return 1;
  size_t cs = 0;
  nvrtcGetCUBINSize(prog, &cs);
  std::string cubin(cs, 0);
  nvrtcGetCUBIN(prog, cubin.data());
  { std::ofstream f3(pre + ".cubin", std::ios::binary); f3.write(cubin.data(), cs); }
  std::printf("cubin %zu bytes\n", cs);
  return 0;
}
