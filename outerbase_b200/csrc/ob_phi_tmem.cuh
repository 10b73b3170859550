/*
 * ob_phi_tmem.cuh -- Phi a / Phi^T r, second generation: factor columns resident in TENSOR MEMORY.
 * (included by ob_kernels.cu after the first-generation kernels; same word streams, ob_terms.hpp)
 *
 * profiles/r01: the shared-memory kernels are bound by LDS wavefronts (one 8-byte factor per
 * (row, term) through a 128 B/clk/SM pipe).  tools/microbench.cu measured that tcgen05.ld reads
 * TMEM at >= 177 B/clk/SM *in addition* to LDS.  So the hot factor columns of a row tile are kept
 * in TMEM -- lane = row, 2R consecutive 32-bit columns = the R rows a lane owns of one factor, one
 * tcgen05.ld.32x32b.x8 per word -- and only the cold columns go through shared memory.
 *
 * A warp reaches only the 32 TMEM lanes of its quadrant (warp % 4), so a CTA is split into two
 * TEAMS by quadrant pair: team = (warp%4)/2 owns 64 lanes x R rows = a 64R-row tile and all 512
 * columns of those lanes (64 factors at R=4).  Its 8 warps = 2 quadrants x 4 term groups.  The
 * teams run out of phase on interleaved tiles, synchronised with named barriers: while one team
 * refills TMEM (ld.global -> tcgen05.st, op applied in registers) the other one computes.
 */
#pragma once

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int R>
__device__ __forceinline__ void tmem_load(double (&f)[R], uint32_t taddr) {
  if constexpr (R == 4) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
        : "r"(taddr)
        : "memory");
    f[0] = __hiloint2double(r1, r0); f[1] = __hiloint2double(r3, r2);
    f[2] = __hiloint2double(r5, r4); f[3] = __hiloint2double(r7, r6);
  } else {
    uint32_t r0, r1, r2, r3;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "r"(taddr)
        : "memory");
    f[0] = __hiloint2double(r1, r0); f[1] = __hiloint2double(r3, r2);
  }
}
template <int R>
__device__ __forceinline__ void tmem_store(uint32_t taddr, const double (&f)[R]) {
  if constexpr (R == 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__double2loint(f[0])), "r"(__double2hiint(f[0])), "r"(__double2loint(f[1])), "r"(__double2hiint(f[1])),
                 "r"(__double2loint(f[2])), "r"(__double2hiint(f[2])), "r"(__double2loint(f[3])), "r"(__double2hiint(f[3]))
                 : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(f[0])),
                 "r"(__double2hiint(f[0])), "r"(__double2loint(f[1])), "r"(__double2hiint(f[1]))
                 : "memory");
  }
}

/* rows of the team tile owned by a lane: (2l, 2l+1) and, for R=4, (128+2l, 128+2l+1), l = 0..63 */
template <int R>
__device__ __forceinline__ int team_row(int ell, int r) { return (r >> 1) * 128 + 2 * ell + (r & 1); }

/* cold factor from the team's shared-memory tile ([column][64R rows]) */
template <int R>
__device__ __forceinline__ void smem_load(double (&f)[R], uint32_t tls, uint32_t c) {
  const uint32_t addr = tls + c * (uint32_t)(64 * R * sizeof(double));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(f[0]), "=d"(f[1]) : "r"(addr));
  if constexpr (R == 4) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+1024];" : "=d"(f[2]), "=d"(f[3]) : "r"(addr));
}

/* bit 15 of a word's column field: 0 = TMEM slot, 1 = shared-memory column (set by the host
 * compiler, ob_terms.hpp tmem_cap) */
template <int R>
__device__ __forceinline__ void factor_load(double (&f)[R], uint32_t w, uint32_t tbase, uint32_t tls) {
  const uint32_t idx = w & 0x7FFFu;
  if (!(w & 0x8000u)) tmem_load<R>(f, tbase + idx * (2 * R));
  else smem_load<R>(f, tls, idx);
}

/* split form for two factors in flight: issue both loads, then ONE wait that the compiler must
 * order before any use (the registers are in/out operands of the wait) */
template <int R>
struct Fac { uint32_t r[2 * R]; };
template <int R>
__device__ __forceinline__ void factor_issue(Fac<R>& f, uint32_t w, uint32_t tbase, uint32_t tls) {
  const uint32_t idx = w & 0x7FFFu;
  if (!(w & 0x8000u)) {
    const uint32_t taddr = tbase + idx * (2 * R);
    if constexpr (R == 4)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(f.r[0]), "=r"(f.r[1]), "=r"(f.r[2]), "=r"(f.r[3]), "=r"(f.r[4]), "=r"(f.r[5]), "=r"(f.r[6]), "=r"(f.r[7])
                   : "r"(taddr) : "memory");
    else
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(f.r[0]), "=r"(f.r[1]), "=r"(f.r[2]), "=r"(f.r[3]) : "r"(taddr) : "memory");
  } else {
    const uint32_t addr = tls + idx * (uint32_t)(64 * R * sizeof(double));
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(f.r[0]), "=r"(f.r[1]), "=r"(f.r[2]), "=r"(f.r[3]) : "r"(addr));
    if constexpr (R == 4)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+1024];" : "=r"(f.r[4]), "=r"(f.r[5]), "=r"(f.r[6]), "=r"(f.r[7]) : "r"(addr));
  }
}
/* pins registers behind the preceding (volatile) wait without emitting an instruction */
template <int R>
__device__ __forceinline__ void factor_pin2(Fac<R>& a, Fac<R>& b) {
  if constexpr (R == 4)
    asm volatile("" : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(a.r[4]), "+r"(a.r[5]), "+r"(a.r[6]), "+r"(a.r[7]),
                      "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7]) :: "memory");
  else
    asm volatile("" : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]) :: "memory");
}
template <int R>
__device__ __forceinline__ void factor_wait2(Fac<R>& a, Fac<R>& b) {
  if constexpr (R == 4)
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(a.r[4]), "+r"(a.r[5]), "+r"(a.r[6]), "+r"(a.r[7]),
                   "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7])
                 :: "memory");
  else
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3])
                 :: "memory");
}
template <int R>
__device__ __forceinline__ double fac_get(const Fac<R>& f, int r) { return __hiloint2double((int)f.r[2 * r + 1], (int)f.r[2 * r]); }

__device__ __forceinline__ double apply_col_op(const PhiKParams& p, int col, double x, unsigned long long row) {
  const int op = p.col_op[col];
  if ((op & 255) == COL_SQUARE) return x * x;
  if ((op & 255) == COL_TWO_G_B) return 2.0 * (x * p.load_src[op >> 8][row]);
  return x;
}

/* team-wide refill: TMEM slots [0, lt) by the quadrant's four warps, cold columns into shared memory */
template <int R>
__device__ __forceinline__ void team_fill(const PhiKParams& p, unsigned long long row0, int g, int ell, int ttid,
                                          uint32_t tbase, double* Ts, uint64_t* bar) {
  constexpr int TRT = 64 * R;
  if (p.ncs > 0) {
    if (!p.has_ops) {
      if (ttid == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, (uint32_t)(TRT * sizeof(double)) * (uint32_t)p.ncs);
#pragma unroll 1
        for (int c = 0; c < p.ncs; ++c) bulk_g2s(Ts + (size_t)c * TRT, p.load_src[p.lt + c] + row0, TRT * sizeof(double), bar);
      }
    } else {
      for (int idx = ttid; idx < p.ncs * TRT; idx += 256) {
        const int c = idx / TRT, r = idx - c * TRT;
        Ts[idx] = apply_col_op(p, p.lt + c, p.load_src[p.lt + c][row0 + r], row0 + r);
      }
    }
  }
#pragma unroll 1
  for (int s0 = g; s0 < p.lt; s0 += 16) { /* four slots in flight per warp */
    double v[4][R];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + 4 * u;
      if (s < p.lt) {
        const double* src = p.load_src[s] + row0;
        const double2 a = __ldg(reinterpret_cast<const double2*>(src + 2 * ell));
        v[u][0] = a.x; v[u][1] = a.y;
        if constexpr (R == 4) {
          const double2 b = __ldg(reinterpret_cast<const double2*>(src + 128 + 2 * ell));
          v[u][2] = b.x; v[u][3] = b.y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + 4 * u;
      if (s < p.lt) {
        if (p.has_ops) {
#pragma unroll
          for (int r = 0; r < R; ++r) v[u][r] = apply_col_op(p, s, v[u][r], row0 + team_row<R>(ell, r));
        }
        tmem_store<R>(tbase + (uint32_t)s * (2 * R), v[u]);
      }
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

/* ------------------------------------------------------------------ Phi a (TMEM) */
template <int R, bool PS>
__global__ void __launch_bounds__(512, 1) phi_a2_kernel(const PhiKParams p) {
  constexpr int TRT = 64 * R;
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16);
  double* a_sm = reinterpret_cast<double*>(smem + p.off_vec);
  uint32_t* prog_sm = reinterpret_cast<uint32_t*>(smem + p.off_prog);
  double* part = reinterpret_cast<double*>(smem + p.off_part);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, team = q >> 1, g = warp >> 2;
  const int ell = 32 * (q & 1) + lane, ttid = g * 64 + ell;
  double* Ts = reinterpret_cast<double*>(smem + p.off_tile) + (size_t)team * p.tile_doubles;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  for (int i = tid; i < p.nslots; i += 512) { const int t = p.slot_term[i]; a_sm[i] = t >= 0 ? p.a[t] : 0.0; }
  if (PS) for (int i = tid; i < p.nwords; i += 512) prog_sm[i] = p.prog[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * q) << 16);
  const uint32_t tls = smem_u32(Ts) + 16u * (uint32_t)ell;
  const uint32_t* pw0 = (PS ? prog_sm : p.prog) + p.prog_off[g];
  const int slot_hi = (int)p.slot_base[g] + (int)p.slot_real[g] - 1;

  uint32_t phase = 0;
  double ssq_local = 0.0;
  for (int tile = 2 * blockIdx.x + team; tile < p.ntiles; tile += 2 * gridDim.x) {
    const unsigned long long row0 = (unsigned long long)tile * TRT;
    team_fill<R>(p, row0, g, ell, ttid, tbase, Ts, &bars[team]);
    tc_fence_before();
    named_bar_sync(1 + team, 256);
    tc_fence_after();
    if (p.ncs > 0 && !p.has_ops) { mbar_wait(&bars[team], phase); phase ^= 1u; }

    double cur[R], stk[4][R];
#pragma unroll
    for (int r = 0; r < R; ++r) cur[r] = 0.0;
    const uint32_t* pw = pw0;
    uint32_t w = pw[0], w1 = pw[1], w2 = pw[2], w3 = pw[3];
    int slot = slot_hi;
#pragma unroll 1
    for (;;) {
      if ((int32_t)(w & w1 & w2 & w3) < 0) { /* four B_LEAF words: four factor loads in flight */
        const uint32_t n0 = pw[4], n1 = pw[5], n2 = pw[6], n3 = pw[7];
        pw += 4;
        const double av0 = a_sm[slot], av1 = a_sm[slot - 1], av2 = a_sm[slot - 2], av3 = a_sm[slot - 3];
        slot -= 4;
        Fac<R> fa, fb, fc, fd;
        factor_issue<R>(fa, w, tbase, tls);
        factor_issue<R>(fb, w1, tbase, tls);
        factor_issue<R>(fc, w2, tbase, tls);
        factor_issue<R>(fd, w3, tbase, tls);
        if ((w & w1 & w2 & w3 & 0x8000u) == 0) { factor_wait2<R>(fa, fb); factor_pin2<R>(fc, fd); }
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(fac_get<R>(fa, r), av0, cur[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(fac_get<R>(fb, r), av1, cur[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(fac_get<R>(fc, r), av2, cur[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(fac_get<R>(fd, r), av3, cur[r]);
        w = n0; w1 = n1; w2 = n2; w3 = n3;
        continue;
      }
      if ((int32_t)(w & w1) < 0) { /* two B_LEAF words: both factor loads in flight together */
        const uint32_t n0 = pw[4], n1 = pw[5];
        pw += 2;
        const double av0 = a_sm[slot], av1 = a_sm[slot - 1];
        slot -= 2;
        Fac<R> fa, fb;
        factor_issue<R>(fa, w, tbase, tls);
        factor_issue<R>(fb, w1, tbase, tls);
        if ((w & w1 & 0x8000u) == 0) factor_wait2<R>(fa, fb);
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(fac_get<R>(fa, r), av0, cur[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(fac_get<R>(fb, r), av1, cur[r]);
        w = w2; w1 = w3; w2 = n0; w3 = n1;
        continue;
      }
      const uint32_t n0 = pw[4];
      ++pw;
      if ((int32_t)w < 0) { /* single B_LEAF */
        double f[R];
        const double av = a_sm[slot--];
        factor_load<R>(f, w, tbase, tls);
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = fma(f[r], av, cur[r]);
      } else {
        const uint32_t op = w >> 28;
        if (op == B_END) break;
        const uint32_t e = (w >> 24) & 15u;
        if (op == B_CLOSE_FRESH) {
          double f[R];
          const double av = (w & (FLAG_HAS_A << 20)) ? a_sm[slot--] : 0.0;
          factor_load<R>(f, w, tbase, tls);
#pragma unroll
          for (int r = 0; r < R; ++r) cur[r] = f[r] * (av + cur[r]);
        } else if (op == B_CLOSE_LOAD) {
          double f[R];
          const double av = (w & (FLAG_HAS_A << 20)) ? a_sm[slot--] : 0.0;
          factor_load<R>(f, w, tbase, tls);
#define X(i) case i + 1: _Pragma("unroll") for (int r = 0; r < R; ++r) cur[r] = fma(f[r], av + cur[r], stk[i][r]); break;
          switch (e) { OB_CASES4(X) default: break; }
#undef X
        } else if (op == B_SAVE) {
#define X(i) case i: _Pragma("unroll") for (int r = 0; r < R; ++r) { stk[i][r] = cur[r]; cur[r] = 0.0; } break;
          switch (e) { OB_CASES4(X) default: break; }
#undef X
        } else { /* B_ROOT */
          const double av = a_sm[slot--];
#pragma unroll
          for (int r = 0; r < R; ++r) cur[r] += av;
        }
      }
      w = w1; w1 = w2; w2 = w3; w3 = n0;
    }
    double* mypart = part + (size_t)(team * 4 + g) * TRT;
#pragma unroll
    for (int r = 0; r < R; ++r) mypart[team_row<R>(ell, r)] = cur[r];
    named_bar_sync(1 + team, 256);
    if (ttid < TRT) {
      const unsigned long long row = row0 + ttid;
      if (row < p.N) {
        const double* tp = part + (size_t)(team * 4) * TRT + ttid;
        const double s = ((tp[0] + tp[TRT]) + tp[2 * TRT]) + tp[3 * TRT];
        double sc = p.scale[row];
        if (p.sq) sc = sc * sc;
        const double yv = s * sc;
        if (p.mode == PHI_PLAIN) p.out[row] = yv;
        else if (p.mode == PHI_UPDATE) {
          p.out[row] = yv;
          const double rt = (yv - p.y[row]) / p.sd;
          ssq_local += rt * rt;
          p.w[row] = -1. * (rt / p.sd);
        } else {
          p.w[row] = (yv / p.sd) / p.sd;
        }
      }
    }
    named_bar_sync(1 + team, 256);
  }
  __syncthreads();
  if (p.mode == PHI_UPDATE) {
    part[tid] = ssq_local;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < 512; ++i) s += part[i];
      p.ssq_partial[blockIdx.x] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_slot) : "memory");
}

/* ------------------------------------------------------------------ Phi^T r (TMEM) */
template <int R, bool PS>
__global__ void __launch_bounds__(512, 1) phi_t2_kernel(const PhiKParams p) {
  constexpr int TRT = 64 * R;
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16);
  double* acc_sm = reinterpret_cast<double*>(smem + p.off_vec); /* 2 teams x nslots */
  uint32_t* prog_sm = reinterpret_cast<uint32_t*>(smem + p.off_prog);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, team = q >> 1, g = warp >> 2;
  const int ell = 32 * (q & 1) + lane, ttid = g * 64 + ell;
  double* Ts = reinterpret_cast<double*>(smem + p.off_tile) + (size_t)team * p.tile_doubles;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  for (int i = tid; i < 4 * p.nslots; i += 512) acc_sm[i] = 0.0;
  if (PS) for (int i = tid; i < p.nwords; i += 512) prog_sm[i] = p.prog[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * q) << 16);
  const uint32_t tls = smem_u32(Ts) + 16u * (uint32_t)ell;
  const uint32_t* pw0 = (PS ? prog_sm : p.prog) + p.prog_off[g];
  /* the same term group runs in both quadrants of both teams: one accumulator set per (team, quadrant) */
  double* my_acc = acc_sm + (size_t)(2 * team + (q & 1)) * p.nslots;
  const int slot_lo = (int)p.slot_base[g];
  double* stage = reinterpret_cast<double*>(smem + p.off_part) + (size_t)warp * (16 * 33);

  uint32_t phase = 0;
  for (int tile = 2 * blockIdx.x + team; tile < p.ntiles; tile += 2 * gridDim.x) {
    const unsigned long long row0 = (unsigned long long)tile * TRT;
    double cur[R], stk[4][R];
#pragma unroll
    for (int r = 0; r < R; ++r) { /* b = basescale % a (linalg.cpp:312) */
      const unsigned long long row = row0 + team_row<R>(ell, r);
      double b = 0.0;
      if (row < p.N) { double sc = p.scale[row]; if (p.sq) sc = sc * sc; b = sc * p.win[row]; }
      stk[0][r] = b;
      cur[r] = b;
    }
    team_fill<R>(p, row0, g, ell, ttid, tbase, Ts, &bars[team]);
    tc_fence_before();
    named_bar_sync(1 + team, 256);
    tc_fence_after();
    if (p.ncs > 0 && !p.has_ops) { mbar_wait(&bars[team], phase); phase ^= 1u; }

    const uint32_t* pw = pw0;
    uint32_t w = pw[0], w1 = pw[1], w2 = pw[2], w3 = pw[3];
    int slot = slot_lo;
    /* Row sums across the warp go through a per-warp shared-memory staging block (the LSU is idle
     * in this kernel: factors come from TMEM): every lane stores its partial of emit e at
     * stage[e][lane]; after 16 emits lanes 0..15 each add one row of 32 partials in lane order
     * (deterministic) into the slot accumulators.  ~7 instructions per emit instead of ~24 for the
     * shuffle butterfly. */
    uint32_t ecnt = 0;
    auto emit = [&](double acc) {
      stage[(ecnt & 15u) * 33u + (uint32_t)lane] = acc;
      ++ecnt;
      if ((ecnt & 15u) == 0) {
        __syncwarp();
        if (lane < 16) {
          const double* row = stage + lane * 33;
          double sacc = row[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) sacc += row[j];
          my_acc[slot + lane] += sacc;
        }
        __syncwarp();
        slot += kEmitBatch;
      }
    };
#pragma unroll 1
    for (;;) {
      if ((int32_t)(w & w1 & w2 & w3) < 0) { /* four F_LEAF words of the same parent */
        const uint32_t n0 = pw[4], n1 = pw[5], n2 = pw[6], n3 = pw[7];
        pw += 4;
        Fac<R> fa, fb, fc, fd;
        factor_issue<R>(fa, w, tbase, tls);
        factor_issue<R>(fb, w1, tbase, tls);
        factor_issue<R>(fc, w2, tbase, tls);
        factor_issue<R>(fd, w3, tbase, tls);
        if ((w & w1 & w2 & w3 & 0x8000u) == 0) { factor_wait2<R>(fa, fb); factor_pin2<R>(fc, fd); }
        double acc0 = cur[0] * fac_get<R>(fa, 0), acc1 = cur[0] * fac_get<R>(fb, 0);
        double acc2 = cur[0] * fac_get<R>(fc, 0), acc3 = cur[0] * fac_get<R>(fd, 0);
#pragma unroll
        for (int r = 1; r < R; ++r) {
          acc0 = fma(cur[r], fac_get<R>(fa, r), acc0); acc1 = fma(cur[r], fac_get<R>(fb, r), acc1);
          acc2 = fma(cur[r], fac_get<R>(fc, r), acc2); acc3 = fma(cur[r], fac_get<R>(fd, r), acc3);
        }
        emit(acc0); emit(acc1); emit(acc2); emit(acc3);
        w = n0; w1 = n1; w2 = n2; w3 = n3;
        continue;
      }
      if ((int32_t)(w & w1) < 0) { /* two F_LEAF words of the same parent */
        const uint32_t n0 = pw[4], n1 = pw[5];
        pw += 2;
        Fac<R> fa, fb;
        factor_issue<R>(fa, w, tbase, tls);
        factor_issue<R>(fb, w1, tbase, tls);
        if ((w & w1 & 0x8000u) == 0) factor_wait2<R>(fa, fb);
        double acc0 = cur[0] * fac_get<R>(fa, 0), acc1 = cur[0] * fac_get<R>(fb, 0);
#pragma unroll
        for (int r = 1; r < R; ++r) { acc0 = fma(cur[r], fac_get<R>(fa, r), acc0); acc1 = fma(cur[r], fac_get<R>(fb, r), acc1); }
        emit(acc0);
        emit(acc1);
        w = w2; w1 = w3; w2 = n0; w3 = n1;
        continue;
      }
      const uint32_t n0 = pw[4];
      ++pw;
      double acc;
      if ((int32_t)w < 0) { /* single F_LEAF */
        double f[R];
        factor_load<R>(f, w, tbase, tls);
        acc = cur[0] * f[0];
#pragma unroll
        for (int r = 1; r < R; ++r) acc = fma(cur[r], f[r], acc);
      } else {
        const uint32_t op = w >> 28, d = (w >> 24) & 15u;
        if (op == F_END) break;
        acc = 0.0;
        if (op == F_DESC_CUR || op == F_DESC_STK) {
          double f[R];
          factor_load<R>(f, w, tbase, tls);
          if (op == F_DESC_STK) {
#define X(i) case i + 1: _Pragma("unroll") for (int r = 0; r < R; ++r) cur[r] = stk[i][r] * f[r]; break;
            switch (d) { OB_CASES4(X) default: break; }
#undef X
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) cur[r] *= f[r];
          }
          if (w & (FLAG_SAVE << 20)) {
#define X(i) case i: _Pragma("unroll") for (int r = 0; r < R; ++r) stk[i][r] = cur[r]; break;
            switch (d) { OB_CASES4(X) default: break; }
#undef X
          }
          acc = cur[0];
#pragma unroll
          for (int r = 1; r < R; ++r) acc += cur[r];
        } else if (op == F_ROOT) {
          acc = stk[0][0];
#pragma unroll
          for (int r = 1; r < R; ++r) acc += stk[0][r];
        } else if (op == F_LOADCUR) {
#define X(i) case i: _Pragma("unroll") for (int r = 0; r < R; ++r) cur[r] = stk[i][r]; break;
          switch (d) { OB_CASES4(X) default: break; }
#undef X
        }
      }
      if (w & (FLAG_EMIT << 20)) emit(acc);
      w = w1; w1 = w2; w2 = w3; w3 = n0;
    }
    named_bar_sync(1 + team, 256); /* TMEM / shared tile may be refilled */
  }
  __syncthreads();
  for (int i = tid; i < p.nslots; i += 512) /* fixed order: (team0,q0) + (team0,q1) + (team1,q0) + (team1,q1) */
    p.partial[(size_t)blockIdx.x * p.nslots + i] =
        ((acc_sm[i] + acc_sm[p.nslots + i]) + acc_sm[2 * p.nslots + i]) + acc_sm[3 * p.nslots + i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_slot) : "memory");
}
