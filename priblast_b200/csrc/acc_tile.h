// Tile-persistent span march (kernel set v3).
//
// One CTA owns TX consecutive columns of the batch and walks ALL spans for them, keeping the source rows
// of the 2-D stencils in shared-memory ring buffers, so nothing is re-read from HBM.  Because a cell (i, d)
// only depends on sub-intervals of [i, i+d] (inside) or on super-intervals within the band (outside), a CTA
// that also carries a halo of W+1 columns on the right (inside) or on the left (outside) can recompute
// everything its owned cells need without talking to its neighbours: no grid-wide synchronisation.
//
// Time tiling.  The long-range sums of a cell (generic interior loops: ~435 terms, multibranch splits,
// long bulges) only read rows that are at least kTT spans away from the cell, so the sums of kTT
// consecutive spans can be evaluated TOGETHER from the rows that exist before the first of them is
// produced: one "deep" step loads every source element once and feeds kTT accumulators (kTT FMAs per
// shared-memory load instead of one), then kTT "shallow" steps finish the cells span by span (stack,
// 1-nt bulges, 1x1/1x2/2x1/2x2 loops, multiloop bookkeeping) with a __syncthreads in between.
//
// Synchronisation.  A cell only reads ring rows of its own warp's columns and of the columns of ONE neighbouring
// warp (inside: up to 29 columns to the right; outside: up to 30 to the left), so the kernels do not use a
// CTA-wide barrier per span: every warp publishes a progress counter and waits only for its neighbour
// ("forward": the neighbour has produced the rows I read; "back-pressure": the other neighbour has finished
// reading the ring slots I am about to overwrite).  The ring sizes below leave 3 steps of slack between
// neighbouring warps, so the warps drift apart and the latency-bound shallow steps of one warp overlap with the
// issue-bound deep steps of others.  The order of additions per cell is not affected.
//
// The functions below are the per-thread bodies (thread t = local column t of the tile); the CUDA kernels
// and the host emulation (tests/hostemu) both call them.  The order of additions inside a cell is
// identical to acc_core.h's one-cell-at-a-time functions, so both give the same bits in the same
// precision (checked by tests/test_hostemu.py).
#pragma once
#include <string.h>
#include "acc_core.h"

// 1 = the FP32 tile kernels evaluate the generic interior-loop sums by the centre-line chain (below); needs 8 more
// shared-memory rows and ~54 more registers per thread, which forces a narrower tile — measured on B200: no net gain
// (DESIGN.md §8), so the product build leaves it off; tests/hostemu builds with 1 to keep the formulation tested.
#ifndef PRIB_CHAIN
#define PRIB_CHAIN 0
#endif

namespace prib {

enum {
  kTT = 4,         // spans per deep step
  kRingIn = 32,    // inside: source rows d-30..d-1 plus the row being written
  kRingOut = 34,   // outside: source rows d+1..d+32 plus the row being written
  kRingStem = 8,
  kRingSE = 8,
  kRingMu = 4,     // Alpha/Beta_multi and _multi2: the previous row is read, 4 slots leave 3 steps of slack between
                   // neighbouring warps (the kernels synchronise warp to warp, not CTA-wide)
  kXchRows = 2 * 4, // generic-loop sums handed from the centre-line chain to the column threads, double buffered
  kTileRows = 88 + (PRIB_CHAIN ? kXchRows : 0),  // shared-memory rows of TC reals per CTA (both passes)
  kTilePad = 64,   // zeroed reals in front of the rings: the outside stencils look up to 33 columns to the left
  kOutBaseLead = 2,  // outside pass: base codes staged in shared memory start 2 columns left of the tile ...
  kOutBaseTail = 8,  // ... and reach W + 4 columns past its right edge (s[d + 3] of the last column)
};

// L1 prefetch of a global line a few loop rounds ahead (multibranch sums: every round touches one new line per array).
// Measured on B200 (profiles/r2/experiments.md): slightly SLOWER than without (the extra issue slots cost more than the
// L2 latency they hide), so it stays an opt-in experiment switch.
PRIB_HD void prefetch_l1(const void *p) {
#if defined(__CUDA_ARCH__) && defined(PRIB_PREFETCH)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
enum { kPfDist = 6 };  // rows ahead

#if defined(__CUDA_ARCH__)
// Blackwell packed FP32: one FFMA2 = two FMAs with the same multiplicand v (scalar operand, broadcast to both
// halves) and a coefficient PAIR from the constant bank (uniform-register operand): acc.lo += v * g.x, acc.hi += v * g.y.
__device__ __forceinline__ void ffma2_bcast(unsigned long long &acc, float v, float2 g) {
  unsigned long long vv, gg;
  asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
  asm("mov.b64 %0, {%1, %2};" : "=l"(gg) : "f"(g.x), "f"(g.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(vv), "l"(gg));
}
__device__ __forceinline__ void unpack2(unsigned long long p, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
  unsigned long long p;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
  return p;
}
// (d0, d1) += (a0, a1) * (b0, b1) as one FFMA2
__device__ __forceinline__ void fma2_into(float &d0, float &d1, float a0, float a1, float b0, float b1) {
  unsigned long long acc = pack2f(d0, d1);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(pack2f(a0, a1)), "l"(pack2f(b0, b1)));
  unpack2(acc, d0, d1);
}
// (s0, s1) = (a0, a1) + (b, b) as one FADD2
__device__ __forceinline__ void add2(float &s0, float &s1, float a0, float a1, float b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2f(a0, a1)), "l"(pack2f(b, b)));
  unpack2(r, s0, s1);
}
#endif

// Called by the shallow steps right after their shared-memory / scratch stores: the CUDA kernels arrive at a
// split barrier there, so the persistent global stores and range checks that follow overlap the barrier skew.
struct NoStepHook {
  PRIB_HD void operator()() const {}
};

template <typename real>
struct Tile {
  typedef Core<real> K;
  typedef typename K::Ctx Ctx;
  typedef typename K::SmallTables ST;

  struct Geo {
    long long g0;  // first owned global column
    int TC, TX, H; // columns in smem, owned columns, halo = W + 1
  };

  struct ColState {  // per thread, fixed for the whole tile
    int sq;          // sequence id (-1: padding)
    int L, i;        // sequence length and left index of this column (i < 0: padding / outside batch)
    long long zcol;  // column holding log Z of this sequence (seq_off + L)
  };

  static PRIB_HD void col_state(const Ctx &c, long long g, ColState &cs) {
    cs.sq = -1;
    cs.L = 0;
    cs.i = -1;
    cs.zcol = 0;
    if (g < 0 || g >= c.NC) return;
    const int sq = c.col_seq[g];
    if (sq < 0) return;
    cs.sq = sq;
    cs.L = c.seq_len[sq];
    cs.i = (int)(g - c.seq_off[sq]);
    cs.zcol = c.seq_off[sq] + cs.L;
  }

  // index into the coefficient-pair table of the packed stencils: ninio step of (u1, sum), 7 = "not a term"
  static PRIB_HD constexpr int in_cidx(int u1, int sum) {
    if (!(sum >= 4 && sum <= kMaxLoop && u1 <= sum - 1 && !(sum == 4 && u1 == 2))) return 7;
    const int k = 2 * u1 - sum, a = k < 0 ? -k : k;
    return a > 6 ? 6 : a;
  }
  // One product of the generic-loop stencils.  The host emulation (tests/hostemu) can switch it to the arithmetic of a
  // tensor-core formulation -- both operands split into TF32 halves, x_hi w_hi + x_lo w_hi + x_hi w_lo, FP32 sums
  // (profiles/tc_probe.cu measures that formulation's rate) -- to see what it would do to the parity statement.
#if !defined(__CUDA_ARCH__)
  static int &emu_tf32() {
    static int on = 0;
    return on;
  }
  static float tf32_trunc(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    b &= 0xFFFFE000u;
    memcpy(&v, &b, 4);
    return v;
  }
#endif
  static PRIB_HD real stencil_mul(real w, real v) {
#if !defined(__CUDA_ARCH__)
    if (sizeof(real) == 4 && emu_tf32()) {
      const float wh = tf32_trunc((float)w), wl = tf32_trunc((float)w - wh);
      const float vh = tf32_trunc((float)v), vl = tf32_trunc((float)v - vh);
      return (real)((wh * vh + wl * vh) + wh * vl);
    }
#endif
    return w * v;
  }
  static PRIB_HD real gsel(int k, real g0, real g1, real g2, real g3, real g4, real g5, real g6) {
    return k == 0 ? g0 : k == 1 ? g1 : k == 2 ? g2 : k == 3 ? g3 : k == 4 ? g4 : k == 5 ? g5 : g6;
  }

  // first span of the first deep step: groups of kTT spans that END exactly at W + 1 (spans below kTurn
  // are skipped by the callers)
  static PRIB_HD int first_group(int W) { return W + 2 - ((W + 2 - kTurn + kTT - 1) / kTT) * kTT; }

  // ---- shared-memory carve-up (same kTileRows rows for both passes; base points past the zero pad) ----
  struct InSmem {
    real *stemI, *stemB, *stem, *se, *mu, *m2, *xch;
    const uint8_t *S;  // bases of local columns 0 .. TC+3
  };
  static PRIB_HD InSmem carve_in(real *base, int TC, const uint8_t *S) {
    InSmem s;
    s.stemI = base;
    s.stemB = s.stemI + kRingIn * TC;
    s.stem = s.stemB + kRingIn * TC;
    s.se = s.stem + kRingStem * TC;
    s.mu = s.se + kRingSE * TC;
    s.m2 = s.mu + kRingMu * TC;
    s.xch = s.m2 + kRingMu * TC;
    s.S = S;
    return s;
  }
  struct OutSmem {
    real *stemO, *stemB, *stem, *mu, *m2, *xch;
    const uint8_t *S;  // bases of global columns g0 - H - kOutBaseLead .. (TC + W + kOutBaseTail of them): S[kOutBaseLead + t] = column of thread t
  };
  static PRIB_HD OutSmem carve_out(real *base, int TC, const uint8_t *S) {
    OutSmem s;
    s.S = S;
    s.stemO = base;
    s.stemB = s.stemO + kRingOut * TC;
    s.stem = s.stemB + kRingOut * TC;
    s.mu = s.stem + kRingStem * TC;
    s.m2 = s.mu + kRingMu * TC;
    s.xch = s.m2 + kRingMu * TC;
    return s;
  }

  // ---------------------------------------------------------------------------------------------
  // Centre-line chain for the generic interior-loop sums (FP32 engine).
  //
  // For a target cell (i, j) and loop size s the deep steps need the row sum
  //     H_s(i, j) = sum over u1 = 1 .. s-1 of cg[|2 u1 - s|] * X[r][i + u1]        (r = source row, X = the ...I / ...O ring)
  // The cell (i - 1, j + 1) with loop size s + 2 reads the SAME source row, the same elements with the same
  // coefficients (u1 and u2 both grow by one, |u1 - u2| stays), plus the two new end elements u1 = 1 and u1 = s + 1:
  //     H_{s+2}(i - 1, j + 1) = H_s(i, j) + cg[min(s, 6)] * (two end elements).
  // So a thread that follows one CENTRE LINE (cells (x, d), (x - 1, d + 2), ... in the inside pass; (x, d),
  // (x + 1, d - 2), ... in the outside pass) keeps the 27 running row sums of each span parity in registers and pays
  // 2 loads + 2 flops per (target, loop size) instead of s - 1 of each: ~60 loads + ~100 flops per cell instead of
  // ~130 loads + ~520 FMA slots with the time-tiled direct sums.  All terms are positive, so the chained sums carry no
  // cancellation; the 2x2 loop (s = 4, u1 = 2) is excluded from the target's sum but kept in the chain state.
  // Thread tau owns the centre line of column tau - floor(d / 2) (inside) / tau + floor((W + 1 - d) / 2) (outside);
  // every exact cell of the rectangular tile has its centre thread inside the CTA.  The finished sums gs[k] go to the
  // column-mapped shallow steps through a double-buffered [kTT][TC] exchange array in shared memory.
  // ---------------------------------------------------------------------------------------------
  enum { kChainN = kMaxLoop - 4 + 1 };  // loop sizes 4 .. 30
  struct Chain {
    real f[2][kChainN];  // [parity of the target's step][s - 4]
  };
  static PRIB_HD void clear(Chain &ch) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < kChainN; ++b) ch.f[a][b] = 0;
  }
  // e[u1] = ring element of the loop with left strand u1 (DIR = +1: inside, column x + u1; DIR = -1: outside, column
  // x - 1 - u1, the caller passes e = row + x - 1); returns H_s and updates the chain state
  template <int s, int FILE_, int DIR>
  static PRIB_HD real chain_step(const real *e, const real *cg, Chain &ch) {
    real v;
    if constexpr (s == 4) {
      v = cg[2] * (e[DIR * 1] + e[DIR * 3]);
      ch.f[FILE_][0] = v + cg[0] * e[DIR * 2];
    } else if constexpr (s == 5) {
      v = cg[3] * (e[DIR * 1] + e[DIR * 4]) + cg[1] * (e[DIR * 2] + e[DIR * 3]);
      ch.f[FILE_][1] = v;
    } else {
      v = ch.f[FILE_][s - 6] + cg[s - 2 < 6 ? s - 2 : 6] * (e[DIR * 1] + e[DIR * (s - 1)]);
      ch.f[FILE_][s - 4] = v;
    }
    return v;
  }

  // One source row of the inside stencils (row d0 - S of the Alpha_stemI / Alpha_stemB rings) for all kTT targets;
  // S is a template parameter so that every coefficient choice and column offset is resolved at compile
  // time (a plain `#pragma unroll` nest of this size is not unrolled by nvcc).
  template <int S, int TCC>
  static PRIB_HD void in_rows(const InSmem &sm, int TC, int t, int d0, const real *cf, const real *bu, real g0, real g1,
                              real g2, real g3, real g4, real g5, real g6, real (&gs)[kTT], real (&bs)[kTT]) {
    if (d0 - S >= 5) {  // rows below span 5 hold no stems; same cut as `sum <= min(30, d - 5)` per target
      const real *row = sm.stemI + ((d0 - S) & (kRingIn - 1)) * (TCC > 0 ? TCC : TC) + t;
      real rs[kTT];
#pragma unroll
      for (int k = 0; k < kTT; ++k) rs[k] = 0;
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4 && kTT == 4) {
        // FP32 engine: targets (0, 1) and (2, 3) share one FFMA2 each per source element; a target that does not
        // take the element gets the coefficient 0 (index 7 of the pair table) — the same sums, two FMAs per issue slot
        unsigned long long r01 = 0, r23 = 0;
#if defined(PRIB_SPLIT_ACC)
        unsigned long long r01b = 0, r23b = 0;  // even / odd elements in separate chains: half the dependent-FFMA2 depth
#endif
#pragma unroll
        for (int x = 1; x <= S + kTT - 2; ++x) {
#if defined(PRIB_EXP_HALFLDS)
          const float v = row[x | 1];
#else
          const float v = row[x];
#endif
          const int a0 = in_cidx(x, S), a1 = in_cidx(x, S + 1), a2 = in_cidx(x, S + 2), a3 = in_cidx(x, S + 3);
#if defined(PRIB_SPLIT_ACC)
          if (x & 1) {
            if (a0 != 7 || a1 != 7) ffma2_bcast(r01b, v, g_cgpair_f[a0 * 8 + a1]);
            if (a2 != 7 || a3 != 7) ffma2_bcast(r23b, v, g_cgpair_f[a2 * 8 + a3]);
            continue;
          }
#endif
          if (a0 != 7 || a1 != 7) ffma2_bcast(r01, v, g_cgpair_f[a0 * 8 + a1]);
#if !defined(PRIB_EXP_HALFFMA)
          if (a2 != 7 || a3 != 7) ffma2_bcast(r23, v, g_cgpair_f[a2 * 8 + a3]);
#endif
        }
        float q0, q1, q2, q3;
        unpack2(r01, q0, q1);
        unpack2(r23, q2, q3);
#if defined(PRIB_SPLIT_ACC)
        {
          float p0, p1, p2, p3;
          unpack2(r01b, p0, p1);
          unpack2(r23b, p2, p3);
          q0 += p0, q1 += p1, q2 += p2, q3 += p3;
        }
#endif
        rs[0] = q0, rs[1] = q1, rs[2] = q2, rs[3] = q3;
      } else
#endif
      {
#pragma unroll
        for (int x = 1; x <= S + kTT - 2; ++x) {  // x = u1; target k takes it while x <= (S + k) - 1
          const real v = row[x];
#pragma unroll
          for (int k = 0; k < kTT; ++k) {
            const int sum = S + k;
            if (sum >= 4 && sum <= kMaxLoop && x <= sum - 1 && !(sum == 4 && x == 2))
              rs[k] += stencil_mul(gsel(K::gidx(x, sum), g0, g1, g2, g3, g4, g5, g6), v);
          }
        }
      }
      // bulges of length u = S + k >= 4 out of the Alpha_stemB ring (same slot): inner cells (i + u, j) and (i, j - u)
      const real *rowB = row + kRingIn * (TCC > 0 ? TCC : TC);
      const real b0 = rowB[0];
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4 && kTT == 4) {  // packed: targets in pairs, a target outside 4..30 gets factor 0
#pragma unroll
        for (int k = 0; k < kTT; k += 2) {
          const bool v0 = S + k >= 4 && S + k <= kMaxLoop, v1 = S + k + 1 >= 4 && S + k + 1 <= kMaxLoop;
          if (v0 || v1) {
            fma2_into(gs[k], gs[k + 1], rs[k], rs[k + 1], v0 ? cf[S + k] : 0.f, v1 ? cf[S + k + 1] : 0.f);
            float t0, t1;
            add2(t0, t1, v0 ? rowB[S + k] : 0.f, v1 ? rowB[S + k + 1] : 0.f, b0);
            fma2_into(bs[k], bs[k + 1], t0, t1, v0 ? bu[S + k] : 0.f, v1 ? bu[S + k + 1] : 0.f);
          }
        }
      } else
#endif
      {
#pragma unroll
        for (int k = 0; k < kTT; ++k)
          if (S + k >= 4 && S + k <= kMaxLoop) gs[k] += cf[S + k] * rs[k];
#pragma unroll
        for (int k = 0; k < kTT; ++k)
          if (S + k >= 4 && S + k <= kMaxLoop) bs[k] += bu[S + k] * (rowB[S + k] + b0);
      }
    }
    if constexpr (S < kMaxLoop) in_rows<S + 1, TCC>(sm, TC, t, d0, cf, bu, g0, g1, g2, g3, g4, g5, g6, gs, bs);
  }

  // multibif: mb[k] = sum over m = 5 .. d0+k-5 of multi1[m][t] * multi2[d0+k-m][t+m]; a multi1 element
  // serves all kTT targets.  Two running pointers, every other offset is an immediate.
  static PRIB_HD void in_bif_all(const real *scrM1, const real *scrM2, int TC, int t, int d0, real (&mb)[kTT]) {
    {
      const real *pa = scrM1 + 5 * TC + t;                       // multi1[m][t]
      const real *pb = scrM2 + (long long)(d0 - 5) * TC + t + 5;  // multi2[d0 - m][t + m]; target k: pb[k * TC]
      int m = 5;
#if defined(PRIB_EXP_NOBIF)
      if (d0 == 1000)
#endif
      for (; m + 1 <= d0 - 5; m += 2) {  // two m per round, 10 independent loads in flight
        if (m + kPfDist + 1 <= d0 - 5) {
          prefetch_l1(pa + (kPfDist + 1) * TC);
          prefetch_l1(pb - (kPfDist + 1) * (TC - 1));
        }
        const real a0 = pa[0], a1 = pa[TC];
        real b0[kTT], b1[kTT];
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          b0[k] = pb[k * TC];
          b1[k] = pb[k * TC - (TC - 1)];
        }
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          mb[k] += a0 * b0[k];
          mb[k] += a1 * b1[k];
        }
        pa += 2 * TC;
        pb -= 2 * (TC - 1);
      }
      {  // tail (<= kTT rows): targets drop out one by one (rows below 5 do not exist).  All operands are loaded
         // first: one L2 latency instead of one per row.
        real ta[kTT], tb[kTT][kTT];
#pragma unroll
        for (int j = 0; j < kTT; ++j) {
          const bool ok = m + j <= d0 + kTT - 1 - 5;
          ta[j] = ok ? pa[j * TC] : (real)0;
#pragma unroll
          for (int k = 0; k < kTT; ++k) tb[j][k] = (ok && m + j <= d0 + k - 5) ? pb[k * TC - j * (TC - 1)] : (real)0;
        }
#pragma unroll
        for (int j = 0; j < kTT; ++j) {
          if (m + j <= d0 + kTT - 1 - 5) {
#pragma unroll
            for (int k = 0; k < kTT; ++k)
              if (m + j <= d0 + k - 5) mb[k] += ta[j] * tb[j][k];
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // inside, deep step: thread t = local column t; targets are the cells (i, i + d0 + k), k = 0..kTT-1.
  //   gs[k] = generic interior loops of Alpha_stemend (raccess.cpp:201-215, 808-812) as a stencil over the
  //           Alpha_stemI ring: source row d0 - s serves target k with loop size s + k;
  //   bs[k] = bulges of length >= 4 (:788-795) over the Alpha_stemB ring, same rows;
  //   mb[k] = Alpha_multibif (:131-143) from the per-CTA scratch rows scrM1 / scrM2 ([(W+4)][TC]).
  // Only rows <= d0 - 1 are read, so all kTT targets are legal at once.  TCC: compile-time tile width
  // (0 = ge.TC at run time, host emulation): every ring row offset becomes an immediate.
  // ---------------------------------------------------------------------------------------------
  template <int TCC = 0>
  static PRIB_HD void inside_deep(const ST &T, const Geo &ge, const InSmem &sm, const real *scrM1, const real *scrM2,
                                  int t, int d0, real (&gs)[kTT], real (&mb)[kTT], real (&bs)[kTT]) {
    const int TC = TCC > 0 ? TCC : ge.TC;
    const real *cf = K::cf_tab(T), *bu = K::bulge_tab(T);
    const real g0 = T.cg[0], g1 = T.cg[1], g2 = T.cg[2], g3 = T.cg[3], g4 = T.cg[4], g5 = T.cg[5], g6 = T.cg[6];
#pragma unroll
    for (int k = 0; k < kTT; ++k) gs[k] = mb[k] = bs[k] = 0;
#if defined(PRIB_EXP_NODEEP)
    if (d0 == 1000)
#endif
    in_rows<1, TCC>(sm, TC, t, d0, cf, bu, g0, g1, g2, g3, g4, g5, g6, gs, bs);
    in_bif_all(scrM1, scrM2, TC, t, d0, mb);
  }

  // One source row (d0 - S) in the chain formulation: generic sums through chain_step (thread = centre line, column
  // xb - ((PAR + k) >> 1) for target k), bulges >= 4 as in in_rows (thread = column t).  Rows are walked from S = 30
  // DOWN to 1: target k reads the chain entry s - 2 that target k - 2 (or the previous group) left for the same row.
  template <int S, int K, int PAR>
  static PRIB_HD void in_chain_cell(const real *rowI, int xb, const real *cf, const real *cg, Chain &ch, real (&gs)[kTT]) {
    constexpr int s = S + K;
    if constexpr (s >= 4 && s <= kMaxLoop) {
      const real v = chain_step<s, (PAR + K) & 1, 1>(rowI + xb - ((PAR + K) >> 1), cg, ch);
      gs[K] += cf[s] * v;
    }
  }
  template <int S, int PAR, int TCC>
  static PRIB_HD void in_chain_rows(const InSmem &sm, int TC, int t, int xb, int d0, const real *cf, const real *bu,
                                    const real *cg, Chain &ch, real (&gs)[kTT], real (&bs)[kTT]) {
    if (d0 - S >= 5) {  // rows below span 5 hold no stems (their chain entries are still zero)
      const real *rowI = sm.stemI + ((d0 - S) & (kRingIn - 1)) * (TCC > 0 ? TCC : TC);
      in_chain_cell<S, 0, PAR>(rowI, xb, cf, cg, ch, gs);
      in_chain_cell<S, 1, PAR>(rowI, xb, cf, cg, ch, gs);
      in_chain_cell<S, 2, PAR>(rowI, xb, cf, cg, ch, gs);
      in_chain_cell<S, 3, PAR>(rowI, xb, cf, cg, ch, gs);
      const real *rowB = rowI + kRingIn * (TCC > 0 ? TCC : TC) + t;
      const real b0 = rowB[0];
#pragma unroll
      for (int k = 0; k < kTT; ++k)
        if (S + k >= 4 && S + k <= kMaxLoop) bs[k] += bu[S + k] * (rowB[S + k] + b0);
    }
    if constexpr (S > 1) in_chain_rows<S - 1, PAR, TCC>(sm, TC, t, xb, d0, cf, bu, cg, ch, gs, bs);
  }

  // Deep step in the chain formulation; PAR = d0 & 1 (the same for every group of a kernel launch).  The generic sums
  // of the kTT targets are written to xch[k * TC + column] (one group's half of the exchange array); mb / bs come back
  // to the caller as in inside_deep.
  template <int PAR, int TCC = 0>
  static PRIB_HD void inside_deep_chain(const ST &T, const Geo &ge, const InSmem &sm, const real *scrM1,
                                        const real *scrM2, int t, int d0, Chain &ch, real *xch, real (&mb)[kTT],
                                        real (&bs)[kTT]) {
    const int TC = TCC > 0 ? TCC : ge.TC;
    const real *cf = K::cf_tab(T), *bu = K::bulge_tab(T), *cg = K::cg_tab(T);
    real gs[kTT];
#pragma unroll
    for (int k = 0; k < kTT; ++k) gs[k] = mb[k] = bs[k] = 0;
    const int xb = t - ((d0 - PAR) >> 1);
    in_chain_rows<kMaxLoop, PAR, TCC>(sm, TC, t, xb, d0, cf, bu, cg, ch, gs, bs);
#pragma unroll
    for (int k = 0; k < kTT; ++k) {
      const int x = xb - ((PAR + k) >> 1);
      if (x >= 0) xch[k * TC + x] = gs[k];
    }
    in_bif_all(scrM1, scrM2, TC, t, d0, mb);
  }

  // ---------------------------------------------------------------------------------------------
  // inside, shallow step: finishes cell (i, i + d) of column t given its deep sums gs / mb.
  // ---------------------------------------------------------------------------------------------
  template <int TCC = 0, typename Hook = NoStepHook>
  static PRIB_HD void inside_shallow(const Ctx &c, const ST &T, const Geo &ge, const InSmem &sm, real *scrM1,
                                     real *scrM2, int t, const ColState &cs, int d, real gs, real mb, real bs,
                                     Hook stores_done = Hook()) {
    const int TC = TCC > 0 ? TCC : ge.TC;
    const real *bu = K::bulge_tab(T);
    real stem = 0, stemI = 0, stemB = 0, stemD = 0, se = 0, mu = 0, m1 = 0, m2 = 0;
    const int smax = imin(kMaxLoop, d - 5);  // u1 + u2 <= smax keeps the inner span >= 5
    const int L = cs.L, i = cs.i, j = i + d;
#if defined(PRIB_EXP_NOSHALLOW)
    const bool live = i >= 0 && j <= L && t + d <= TC - 1 && d == 1000;
#else
    const bool live = i >= 0 && j <= L && t + d <= TC - 1;
#endif
    if (live) {
      const uint8_t *s = sm.S + t;
      const int si = s[0], si1 = s[1], sj = s[d], sj1 = s[d + 1];
      const int te = (j != L) ? T.bp[si][sj1] : 0;
      // Special-loop table entries (1x1, 1x2, 2x1, 2x2: 6.4 KB + 32 KB + 160 KB tables, L2-resident) are fetched
      // first: their indices depend on the bases only, and the rest of the step hides the L2 latency.
      real p11 = 0, p21a = 0, p21b = 0, p22 = 0;
#if defined(PRIB_EXP_NOTAB)
      if (te && d == 1000) {
#else
      if (te) {
#endif
        const int sd1 = s[d - 1], sd2 = s[d - 2], s2 = s[2], s3 = s[3];
        if (smax >= 2) p11 = c.e_int11[idx11(te, T.bpr[s2][sd1], si1, sj)];
        if (smax >= 3) {
          p21a = c.e_int21[idx21(te, T.bpr[s2][sd2], si1, sd1, sj)];
          p21b = c.e_int21[idx21(T.bpr[s3][sd1], te, sj, si1, s2)];
        }
        if (smax >= 4) p22 = c.e_int22[idx22(te, T.bpr[s3][sd2], si1, s2, sd1, sj)];
      }
      const int tp = T.bp[si1][sj];
      if (tp) {
        const int t2r = T.bpr[s[2]][s[d - 1]];
        stem = K::template scal<K::kScK2>(T) * (sm.se[((d - 2) & (kRingSE - 1)) * TC + t + 1] +
                       sm.stem[((d - 2) & (kRingStem - 1)) * TC + t + 1] * T.e_stack[tp][t2r]);
      }
      mb *= K::template scal<K::kScInvCA>(T);
      stemD = tp ? stem * K::e_dangle(T, tp, i > 0, si, j < L, sj1) : 0;
      m2 = stemD * K::template scal<K::kScMlIntern>(T) + sm.m2[((d - 1) & (kRingMu - 1)) * TC + t] * K::template scal<K::kScMlBase>(T);
      m1 = m2 + mb;
      mu = sm.mu[((d - 1) & (kRingMu - 1)) * TC + t + 1] * K::template scal<K::kScMlBase>(T) + mb;
      if (tp) {
        stemI = stem * T.e_mmI[T.rt[tp]][sj1][si];
        stemB = stem * T.tau[tp];
      }
      if (te) {
        real a = T.e_hairpin[d] * (d != 3 ? T.e_mmH[te][si1][sj] : T.tau[te]);
        const real *st1 = sm.stem + ((d - 1) & (kRingStem - 1)) * TC + t;
        const real *st2 = sm.stem + ((d - 2) & (kRingStem - 1)) * TC + t;
        const real *st3 = sm.stem + ((d - 3) & (kRingStem - 1)) * TC + t;
        const real *st4 = sm.stem + ((d - 4) & (kRingStem - 1)) * TC + t;
        if (smax >= 1) {
          a += bu[1] * (st1[1] * T.e_stack[te][T.bpr[s[2]][sj]] +
                        st1[0] * T.e_stack[te][T.bpr[si1][s[d - 1]]]);
        }
        if (smax >= 2) {
          a += st2[1] * p11;
          // long bulges: lengths >= 4 were summed by the deep step, 2 and 3 (rows d-2, d-3) follow here
#pragma unroll
          for (int u = 2; u <= 3; ++u) {
            if (u <= smax) {
              const real *row = sm.stemB + ((d - u) & (kRingIn - 1)) * TC + t;
              bs += bu[u] * (row[u] + row[0]);
            }
          }
          a += T.tau[te] * bs;
        }
        if (smax >= 3) {
          a += st3[1] * p21a;
          a += st3[2] * p21b;
        }
        if (smax >= 4) {
          a += st4[2] * p22;
          a += T.e_mmI[te][si1][sj] * gs;
        }
        const int tt = T.rt[te];
        a += mu * K::template scal<K::kScMlClose>(T) * T.e_d3[tt][si1] * T.e_d5[tt][sj];
        se = a;
      }
    }
    // every column refreshes its ring slots every span (zeros where the cell does not exist)
    sm.stemI[(d & (kRingIn - 1)) * TC + t] = stemI;
    sm.stemB[(d & (kRingIn - 1)) * TC + t] = stemB;
    sm.stem[(d & (kRingStem - 1)) * TC + t] = stem;
    sm.se[(d & (kRingSE - 1)) * TC + t] = se;
    sm.mu[(d & (kRingMu - 1)) * TC + t] = mu;
    sm.m2[(d & (kRingMu - 1)) * TC + t] = m2;
#if defined(PRIB_SCR_BEFORE)
    scrM1[d * TC + t] = m1;
    scrM2[d * TC + t] = m2;
#endif
    stores_done();
#if !defined(PRIB_SCR_BEFORE)
    // The multibranch scratch rows go to global memory AFTER the release of this step's progress counter: the first
    // reader of row d is a deep step that waits for a LATER event of this warp (it reads rows <= its d0 - 2), and that
    // event's release covers these stores; a release right behind them would wait for their L2 round trip every step.
    scrM1[d * TC + t] = m1;
    scrM2[d * TC + t] = m2;
#endif
    // persistent outputs: owned columns only
#if defined(PRIB_EXP_NOSTG)
    if (t < ge.TX && i >= 0 && j <= L && d == 1000) {
#else
    if (t < ge.TX && i >= 0 && j <= L) {
#endif
      const long long g = ge.g0 + t;
      if (c.delta == 2) c.put(A_STEM, d, g, stem);  // only the 2x1 / 2x2 loops of delta == 2 read it (BiTile)
      c.put(A_STEMI, d, g, stemI);
      c.put(A_STEMB, d, g, stemB);
      c.put(A_STEMD, d, g, stemD);
      c.put(A_MULTI, d, g, mu);
      c.put(A_MULTI1, d, g, m1);
      c.put(A_MULTI2, d, g, m2);
      if (!(K::in_safe_range(stem) && K::in_safe_range(se) && K::in_safe_range(mu) && K::in_safe_range(m1) &&
            K::in_safe_range(m2)))
        K::raise_flag(&c.flags[cs.sq], K::range_bits(stem) | K::range_bits(se) | K::range_bits(mu) | K::range_bits(m1) |
                                           K::range_bits(m2));
    }
  }

  // ---------------------------------------------------------------------------------------------
  // outside: thread t = local column t, global column g0 - H + t; cell (p, p + d).  Spans descend.
  // Deep step at d0 serves the targets d0 - k, k = 0..kTT-1, from rows >= d0 + 1:
  //   gs[k]  generic interior loops of Beta_stem (raccess.cpp:373-386) over the Beta_stemO ring: source row
  //          d0 + s serves target k with loop size s + k - 2;
  //   bs[k]  bulges longer than 1 nt over the Beta_stemB ring (:788-795 seen from the inner pair);
  //   bm1[k] Beta_multi1 (:310-324) and ks[k], the k-loop of Beta_multi2 (:339-349), from the per-CTA
  //          Beta_multibif scratch scrBif ([(W+4)][TC]) and the Alpha arrays in HBM.
  // ---------------------------------------------------------------------------------------------
  static PRIB_HD int wrap_out(int r) { return r >= kRingOut ? r - kRingOut : r; }

  struct OutDeep {
    real gs[kTT], bs[kTT], bm1[kTT], ks[kTT];
  };

  // One source row (d0 + S of the Beta_stemO / Beta_stemB rings) of the outside stencils for all kTT targets.
  template <int S, int TCC>
  static PRIB_HD void out_rows(const OutSmem &sm, int TC, int t, int d0, int slot_d0, int W, const real *bu,
                               const real *cf, real g0, real g1, real g2, real g3, real g4, real g5, real g6,
                               OutDeep &o) {
    if (d0 + S <= W + 1) {  // same cut as `sum <= min(30, W - 1 - d)` per target
      const int slot = wrap_out(slot_d0 + S);
      const real *rowB = sm.stemB + slot * (TCC > 0 ? TCC : TC) + t - 1;
      const real *rowO = sm.stemO + slot * (TCC > 0 ? TCC : TC) + t - 1;
      const real b0 = rowB[0];
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4 && kTT == 4) {  // packed (see in_rows)
#pragma unroll
        for (int k = 0; k < kTT; k += 2) {
          const int u0 = S + k - 2, u1 = u0 + 1;
          const bool v0 = u0 >= 2 && u0 <= kMaxLoop, v1 = u1 >= 2 && u1 <= kMaxLoop;
          if (v0 || v1) {
            float t0, t1;
            add2(t0, t1, v0 ? rowB[-u0] : 0.f, v1 ? rowB[-u1] : 0.f, b0);
            fma2_into(o.bs[k], o.bs[k + 1], t0, t1, v0 ? bu[u0] : 0.f, v1 ? bu[u1] : 0.f);
          }
        }
      } else
#endif
      {
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          const int u = S + k - 2;
          if (u >= 2 && u <= kMaxLoop) o.bs[k] += bu[u] * (rowB[-u] + b0);
        }
      }
      real rs[kTT];
#pragma unroll
      for (int k = 0; k < kTT; ++k) rs[k] = 0;
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4 && kTT == 4) {  // packed FP32 (see in_rows)
        unsigned long long r01 = 0, r23 = 0;
#pragma unroll
        for (int y = 1; y <= S + kTT - 4; ++y) {
#if defined(PRIB_EXP_HALFLDS)
          const float v = rowO[-(y | 1)];
#else
          const float v = rowO[-y];
#endif
          const int a0 = in_cidx(y, S - 2), a1 = in_cidx(y, S - 1), a2 = in_cidx(y, S), a3 = in_cidx(y, S + 1);
          if (a0 != 7 || a1 != 7) ffma2_bcast(r01, v, g_cgpair_f[a0 * 8 + a1]);
#if !defined(PRIB_EXP_HALFFMA)
          if (a2 != 7 || a3 != 7) ffma2_bcast(r23, v, g_cgpair_f[a2 * 8 + a3]);
#endif
        }
        float q0, q1, q2, q3;
        unpack2(r01, q0, q1);
        unpack2(r23, q2, q3);
        rs[0] = q0, rs[1] = q1, rs[2] = q2, rs[3] = q3;
      } else
#endif
      {
#pragma unroll
        for (int y = 1; y <= S + kTT - 4; ++y) {  // y = u1: source column t - 1 - u1
          const real v = rowO[-y];
#pragma unroll
          for (int k = 0; k < kTT; ++k) {
            const int sum = S + k - 2;
            if (sum >= 4 && sum <= kMaxLoop && y <= sum - 1 && !(sum == 4 && y == 2))
              rs[k] += stencil_mul(gsel(K::gidx(y, sum), g0, g1, g2, g3, g4, g5, g6), v);
          }
        }
      }
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4 && kTT == 4) {
#pragma unroll
        for (int k = 0; k < kTT; k += 2) {
          const int s0 = S + k - 2, s1 = s0 + 1;
          const bool v0 = s0 >= 4 && s0 <= kMaxLoop, v1 = s1 >= 4 && s1 <= kMaxLoop;
          if (v0 || v1) fma2_into(o.gs[k], o.gs[k + 1], rs[k], rs[k + 1], v0 ? cf[s0] : 0.f, v1 ? cf[s1] : 0.f);
        }
      } else
#endif
      {
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          const int sum = S + k - 2;
          if (sum >= 4 && sum <= kMaxLoop) o.gs[k] += cf[sum] * rs[k];
        }
      }
    }
    if constexpr (S < kMaxLoop + 2) out_rows<S + 1, TCC>(sm, TC, t, d0, slot_d0, W, bu, cf, g0, g1, g2, g3, g4, g5, g6, o);
  }

  // Multiloop sums of the outside deep step (see outside_deep): bm1[k] and ks[k] from the per-CTA Beta_multibif
  // scratch and the Alpha arrays in HBM.
  static PRIB_HD void out_multi_all(const Ctx &c, const Geo &ge, const real *scrBif, int TC, int t, const ColState &cs,
                                    int d0, OutDeep &o) {
    const int W = c.W;
    // Multiloop sums.  Only cells strictly inside the sequence (p >= 1, q < L) use them (outside_shallow
    // ignores the sums of all others), and for those the term ranges of the kTT targets line up:
    //   bm1[k]: m = 5 .. min(L - q_k, W - d_k)  <=>  bif row d0 + s with s = m - k = 5 - k .. min(L - p, W) - d0
    //   ks[k] : m = 5 .. min(p, W - d_k)
    // so the loops run on common bounds and every load stays inside the sequence / the scratch rows.
    const long long g = ge.g0 - ge.H + t;
    const int L = cs.L, p = cs.i;
    const long long nc = c.NC;
    {  // bm1[k] += bif[d0 + s][t] * Alpha_multi2[s + k][g + d0 - k]; a bif element serves all targets
      const int shi = p >= 1 ? imin(L - p, W) - d0 : -1;
      const real *pa = scrBif + (long long)(d0 + 5 - (kTT - 1)) * TC + t;                  // bif row d0 + s
      const real *pb = c.arr[A_MULTI2] + (long long)(5 - (kTT - 1)) * nc + g + d0;          // row s, column q_0
      {  // head: target k joins at s = 5 - k.  All operands are loaded first (one L2 latency instead of three).
        real ha[kTT - 1], hb[kTT - 1][kTT];
#pragma unroll
        for (int h = 0; h < kTT - 1; ++h) {
          const bool ok = 5 - (kTT - 1) + h <= shi;
          ha[h] = ok ? pa[h * TC] : (real)0;
#pragma unroll
          for (int k = 0; k < kTT; ++k)
            hb[h][k] = (ok && h + k >= kTT - 1) ? pb[(long long)h * nc + (long long)k * (nc - 1)] : (real)0;
        }
#pragma unroll
        for (int h = 0; h < kTT - 1; ++h) {
          if (5 - (kTT - 1) + h <= shi) {
#pragma unroll
            for (int k = 0; k < kTT; ++k)
              if (h + k >= kTT - 1) o.bm1[k] += ha[h] * hb[h][k];
          }
        }
        pa += (kTT - 1) * TC;
        pb += (long long)(kTT - 1) * nc;
      }
      int s = 5;
#if defined(PRIB_EXP_NOBIF)
      if (d0 == 1000)
#endif
      for (; s + 1 <= shi; s += 2) {  // two rows per round, 10 independent loads in flight
        if (s + kPfDist + 1 <= shi) {
          prefetch_l1(pa + (kPfDist + 1) * TC);
          prefetch_l1(pb + (long long)(kPfDist + 1) * nc + (long long)(kTT - 1) * (nc - 1));
        }
        const real a0 = pa[0], a1 = pa[TC];
        real b0[kTT], b1[kTT];
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          b0[k] = pb[(long long)k * (nc - 1)];
          b1[k] = pb[(long long)k * (nc - 1) + nc];
        }
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          o.bm1[k] += a0 * b0[k];
          o.bm1[k] += a1 * b1[k];
        }
        pa += 2 * TC;
        pb += 2 * nc;
      }
      if (s <= shi) {
        const real a = pa[0];
#pragma unroll
        for (int k = 0; k < kTT; ++k) o.bm1[k] += a * pb[(long long)k * (nc - 1)];
      }
    }
    {  // ks[k] += bif[d0 - k + m][t - m] * Alpha_multi1[m][g - m]; an Alpha element serves all targets
      const int mhi = p >= 1 ? imin(p, W - d0) : -1;  // common part: all targets take m <= W - d0
      const real *pa = scrBif + (long long)(d0 + 5) * TC + t - 5;  // bif[d0 + m][t - m]; target k: pa[-k * TC]
      const real *pb = c.arr[A_MULTI1] + 5 * (nc - 1) + g;         // Alpha_multi1[m][g - m]
      int m = 5;
#if defined(PRIB_EXP_NOBIF)
      if (d0 == 1000)
#endif
      for (; m + 1 <= mhi; m += 2) {
        if (m + kPfDist + 1 <= mhi) {
          prefetch_l1(pb + (long long)(kPfDist + 1) * (nc - 1));
          prefetch_l1(pa + (kPfDist + 1) * (TC - 1));
        }
        const real b0 = pb[0], b1 = pb[nc - 1];
        real a0[kTT], a1[kTT];
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          a0[k] = pa[-k * TC];
          a1[k] = pa[-k * TC + (TC - 1)];
        }
#pragma unroll
        for (int k = 0; k < kTT; ++k) {
          o.ks[k] += a0[k] * b0;
          o.ks[k] += a1[k] * b1;
        }
        pa += 2 * (TC - 1);
        pb += 2 * (nc - 1);
      }
      if (m <= mhi) {
        const real b = pb[0];
#pragma unroll
        for (int k = 0; k < kTT; ++k) o.ks[k] += pa[-k * TC] * b;
      }
      // tail: m = W - d0 + e, e = 1 .. kTT-1, exists only for the targets with k >= e
      {  // (all operands loaded first: one L2 latency instead of three)
        real tb[kTT - 1], tq[kTT - 1][kTT];
#pragma unroll
        for (int e = 1; e < kTT; ++e) {
          const int mm = W - d0 + e;
          const bool ok = p >= 1 && mm >= 5 && mm <= p;
          tb[e - 1] = ok ? c.arr[A_MULTI1][(long long)mm * (nc - 1) + g] : (real)0;
#pragma unroll
          for (int k = 0; k < kTT; ++k)
            tq[e - 1][k] = (ok && k >= e) ? scrBif[(long long)(d0 + mm - k) * TC + t - mm] : (real)0;
        }
#pragma unroll
        for (int e = 1; e < kTT; ++e) {
          const int mm = W - d0 + e;
          if (p >= 1 && mm >= 5 && mm <= p) {
#pragma unroll
            for (int k = 0; k < kTT; ++k)
              if (k >= e) o.ks[k] += tq[e - 1][k] * tb[e - 1];
          }
        }
      }
    }
  }

  template <int TCC = 0>
  static PRIB_HD void outside_deep(const Ctx &c, const ST &T, const Geo &ge, const OutSmem &sm, const real *scrBif,
                                   int t, const ColState &cs, int d0, int slot_d0 /* = d0 % kRingOut */, OutDeep &o) {
    const int TC = TCC > 0 ? TCC : ge.TC, W = c.W;
    const real *bu = K::bulge_tab(T), *cf = K::cf_tab(T);
    const real g0 = T.cg[0], g1 = T.cg[1], g2 = T.cg[2], g3 = T.cg[3], g4 = T.cg[4], g5 = T.cg[5], g6 = T.cg[6];
#pragma unroll
    for (int k = 0; k < kTT; ++k) o.gs[k] = o.bs[k] = o.bm1[k] = o.ks[k] = 0;
    // stencils: source row d0 + s; target k: bulge length / loop size s + k - 2
    out_rows<1, TCC>(sm, TC, t, d0, slot_d0, W, bu, cf, g0, g1, g2, g3, g4, g5, g6, o);
    out_multi_all(c, ge, scrBif, TC, t, cs, d0, o);
  }

  // One source row (d0 + S) of the outside pass in the chain formulation (see "Centre-line chain"): target k has
  // span d0 - k and loop size S + k - 2; its centre-line thread sits at column xb + (k >> 1) (the groups start at
  // W + 1 and have kTT = 4 spans, so the step parity of target k is k & 1).  Bulges as in out_rows (thread = column).
  template <int S, int K>
  static PRIB_HD void out_chain_cell(const real *rowO, int xb, const real *cf, const real *cg, Chain &ch, real (&gs)[kTT]) {
    constexpr int s = S + K - 2;
    if constexpr (s >= 4 && s <= kMaxLoop) {
      const real v = chain_step<s, K & 1, -1>(rowO + xb + (K >> 1) - 1, cg, ch);
      gs[K] += cf[s] * v;
    }
  }
  template <int S, int TCC>
  static PRIB_HD void out_chain_rows(const OutSmem &sm, int TC, int t, int xb, int d0, int slot_d0, int W, const real *bu,
                                     const real *cf, const real *cg, Chain &ch, OutDeep &o) {
    if (d0 + S <= W + 1) {  // rows above W + 1 do not exist (their chain entries are still zero)
      const int slot = wrap_out(slot_d0 + S);
      const real *rowO = sm.stemO + slot * (TCC > 0 ? TCC : TC);
      out_chain_cell<S, 0>(rowO, xb, cf, cg, ch, o.gs);
      out_chain_cell<S, 1>(rowO, xb, cf, cg, ch, o.gs);
      out_chain_cell<S, 2>(rowO, xb, cf, cg, ch, o.gs);
      out_chain_cell<S, 3>(rowO, xb, cf, cg, ch, o.gs);
      const real *rowB = sm.stemB + slot * (TCC > 0 ? TCC : TC) + t - 1;
      const real b0 = rowB[0];
#pragma unroll
      for (int k = 0; k < kTT; ++k) {
        const int u = S + k - 2;
        if (u >= 2 && u <= kMaxLoop) o.bs[k] += bu[u] * (rowB[-u] + b0);
      }
    }
    if constexpr (S > 1) out_chain_rows<S - 1, TCC>(sm, TC, t, xb, d0, slot_d0, W, bu, cf, cg, ch, o);
  }

  // Deep step in the chain formulation: o.gs is written to xch[k * TC + column] instead of being returned.
  template <int TCC = 0>
  static PRIB_HD void outside_deep_chain(const Ctx &c, const ST &T, const Geo &ge, const OutSmem &sm, const real *scrBif,
                                         int t, const ColState &cs, int d0, int slot_d0, Chain &ch, real *xch,
                                         OutDeep &o) {
    const int TC = TCC > 0 ? TCC : ge.TC, W = c.W;
    const real *bu = K::bulge_tab(T), *cf = K::cf_tab(T), *cg = K::cg_tab(T);
#pragma unroll
    for (int k = 0; k < kTT; ++k) o.gs[k] = o.bs[k] = o.bm1[k] = o.ks[k] = 0;
    const int xb = t + ((W + 1 - d0) >> 1);
    out_chain_rows<kMaxLoop + 2, TCC>(sm, TC, t, xb, d0, slot_d0, W, bu, cf, cg, ch, o);
#pragma unroll
    for (int k = 0; k < kTT; ++k) {
      const int x = xb + (k >> 1);
      if (x < TC) xch[k * TC + x] = o.gs[k];
    }
    out_multi_all(c, ge, scrBif, TC, t, cs, d0, o);
  }

  template <int TCC = 0, typename Hook = NoStepHook>
  static PRIB_HD void outside_shallow(const Ctx &c, const ST &T, const Geo &ge, const OutSmem &sm, real *scrBif, int t,
                                      const ColState &cs, int d, int slot_d /* = d % kRingOut */, real gs, real bs,
                                      real bm1, real ks, Hook stores_done = Hook()) {
    const int TC = TCC > 0 ? TCC : ge.TC, W = c.W;
    const real *bu = K::bulge_tab(T);
    real bstem = 0, bstemO = 0, bstemB = 0, bmulti = 0, bmulti2 = 0, bmbif = 0;
    const int smax = imin(kMaxLoop, W - 1 - d);  // source row d + sum + 2 <= W + 1
    const long long g = ge.g0 - ge.H + t;
    const int L = cs.L, p = cs.i, q = p + d;
    // a halo cell is exact iff its end reaches the owned region (all its super-intervals are in the tile)
#if defined(PRIB_EXP_NOSHALLOW)
    const bool live = p >= 0 && q <= L && t + d >= ge.H && d == 1000;
#else
    const bool live = p >= 0 && q <= L && t + d >= ge.H;
#endif
    if (live) {
      const uint8_t *s = sm.S + kOutBaseLead + t;  // staged for columns g0-H-2 .. g0-H+TC+W+5 (the right end q = p + d lies beyond the tile)
      const int sp = s[0], sp1 = s[1], sq_ = s[d], sq1 = s[d + 1];
      const bool inner = (p != 0 && q != L);
      const int te = inner ? T.bp[sp][sq1] : 0;
      // Global operands first (their addresses depend on the bases only; the rest of the step hides the latency):
      // the exponent of the base term and the special-loop table entries.
      const int t2_ = T.bp[sp1][sq_];
      double la = 0, lb = 0, lz = 0;  // loaded here, combined where the base term is formed (in-order issue: a
                                      // dependent add up here would stall the warp on the loads right away)
      real p11 = 0, p21a = 0, p21b = 0, p22 = 0;
#if defined(PRIB_EXP_NOTAB)
      if (t2_ && d == 1000) {
#else
      if (t2_) {
#endif
        la = c.lao[g];
        lb = c.lbo[g + d];
        lz = c.lao[cs.zcol];
        const int t2r_ = T.bpr[sp1][sq_];
        const int sm1 = s[-1], sm2 = s[-2], sd2 = s[d + 2], sd3 = s[d + 3];
        if (smax >= 2) p11 = c.e_int11[idx11(T.bp[sm1][sd2], t2r_, sp, sq1)];
        if (smax >= 3) {
          p21a = c.e_int21[idx21(T.bp[sm1][sd3], t2r_, sp, sq1, sd2)];
          p21b = c.e_int21[idx21(t2r_, T.bp[sm2][sd2], sq1, sm1, sp)];
        }
        if (smax >= 4) p22 = c.e_int22[idx22(T.bp[sm2][sd3], t2r_, sm1, sp, sq1, sd2)];
      }
      const real *b2 = sm.stem + ((d + 2) & (kRingStem - 1)) * TC + t;
      const real bse = (inner && d + 2 <= W + 1) ? b2[-1] : 0;  // Beta_stemend(p,q), :277-279
      if (inner) {
        const int tt = T.rt[te];
        bmulti = (d + 1 <= W + 1 ? sm.mu[((d + 1) & (kRingMu - 1)) * TC + t - 1] * K::template scal<K::kScMlBase>(T) : (real)0) +
                 K::template scal<K::kScK2>(T) * bse * K::template scal<K::kScMlClose>(T) * T.e_d3[tt][sp1] * T.e_d5[tt][sq_];
        bm1 *= K::template scal<K::kScInvCA>(T);
        bmulti2 = bm1 + sm.m2[((d + 1) & (kRingMu - 1)) * TC + t] * K::template scal<K::kScMlBase>(T) + ks * K::template scal<K::kScInvCA>(T);
        bmbif = bm1 + bmulti;
      }
      const int t2 = T.bp[sp1][sq_];
      if (t2) {
        const int t2r = T.bpr[sp1][sq_];
        const real dang = K::e_dangle(T, t2, p > 0, sp, q < L, sq1);
        real l = 0;
        if (smax >= 0) l += bse * T.e_stack[te][t2r];
        const real *b3 = sm.stem + ((d + 3) & (kRingStem - 1)) * TC + t;
        const real *b4 = sm.stem + ((d + 4) & (kRingStem - 1)) * TC + t;
        const real *b5 = sm.stem + ((d + 5) & (kRingStem - 1)) * TC + t;
        const real *b6 = sm.stem + ((d + 6) & (kRingStem - 1)) * TC + t;
        if (smax >= 1) {
          const int ta = T.bp[s[-1]][sq1];
          const int tb = T.bp[sp][s[d + 2]];
          l += bu[1] * (b3[-2] * T.e_stack[ta][t2r] + b3[-1] * T.e_stack[tb][t2r]);
        }
        if (smax >= 2) {
          l += b4[-2] * p11;
          l += T.tau[t2r] * bs;
        }
        if (smax >= 3) {
          l += b5[-2] * p21a;
          l += b5[-3] * p21b;
        }
        if (smax >= 4) {
          l += b6[-3] * p22;
          l += T.e_mmI[t2r][sq1][sp] * gs;
        }
        const real base = (real)exp(la + lb - lz) * dang * T.sB[d];  // raccess.cpp:370, divided by Z
        bstem = base + K::template scal<K::kScK2>(T) * l + bmulti2 * K::template scal<K::kScMlIntern>(T) * dang;
        bstemO = bstem * T.e_mmI[t2][s[2]][s[d - 1]];
        bstemB = bstem * T.tau[t2];
      }
    }
    sm.stemO[slot_d * TC + t] = bstemO;
    sm.stemB[slot_d * TC + t] = bstemB;
    sm.stem[(d & (kRingStem - 1)) * TC + t] = bstem;
    sm.mu[(d & (kRingMu - 1)) * TC + t] = bmulti;
    sm.m2[(d & (kRingMu - 1)) * TC + t] = bmulti2;
#if defined(PRIB_SCR_BEFORE)
    scrBif[d * TC + t] = bmbif;
#endif
    stores_done();
#if !defined(PRIB_SCR_BEFORE)
    scrBif[d * TC + t] = bmbif;  // after the release, see inside_shallow (deep steps read bif rows >= their d0 + 2)
#endif
#if defined(PRIB_EXP_NOSTG)
    if (t >= ge.H && p >= 0 && q <= L && d == 1000) {
#else
    if (t >= ge.H && p >= 0 && q <= L) {
#endif
      c.put(B_STEM, d, g, bstem);
      c.put(B_STEMO, d, g, bstemO);
      c.put(B_STEMB, d, g, bstemB);
      c.put(B_MULTI, d, g, bmulti);
      c.put(B_MULTI2, d, g, bmulti2);
      if (!(K::in_safe_range(bstem) && K::in_safe_range(bmulti) && K::in_safe_range(bmulti2) &&
            K::in_safe_range(bmbif)))
        K::raise_flag(&c.flags[cs.sq], 4 * (K::range_bits(bstem) | K::range_bits(bmulti) | K::range_bits(bmulti2) |
                                                K::range_bits(bmbif)));
    }
  }
};

}  // namespace prib

// ------------------------------------------------------------------------------------------------
// Interior-loop strand weights (restructured raccess.cpp:614-771), tile version.
//
// ML[u1][i] / MR[u2][j'] are sums over the outer cells (i, j' = i + dp) that close a loop.  Only ~3/8 of
// the cells can (Beta_stemend != 0 needs a pair), so each thread first lists ITS valid spans and then
// walks only those: lanes sit at different spans, which is fine because there is no wavefront here and
// the Alpha_stemI tile in shared memory has a row stride that is a multiple of 32 words (bank = column).
// The left kernel reads the tile start-indexed, the right kernel end-indexed (element (r, q) = cell with
// span r ENDING at column q), so that a lane's column never depends on its span.
// ------------------------------------------------------------------------------------------------
namespace prib {

template <typename real>
struct BiTile {
  typedef Core<real> K;
  typedef typename K::Ctx Ctx;
  typedef typename K::SmallTables ST;
  // outer spans handled together by the dense pass (a multiple of 4: the right tile's row offsets repeat every 4 spans)
#ifndef PRIB_BI_TT
#define PRIB_BI_TT 4
#endif
  static constexpr int kTB = PRIB_BI_TT;
  static constexpr int kSmin = 5 - kTB;  // first source-row offset: loop size 4 of the LAST span of a group

  struct Geo {
    long long g0;  // first owned column
    int TXb;       // owned columns = threads
    int cols;      // tile_cols(): row stride of the tile
    int rows;      // W - 5: spans 5 .. W-1
  };

  // Tiles.  Left kernel: start-indexed, element (r, x) = cell of span r starting at column g0 + x.  Right
  // kernel: end-indexed, element (r, x) = cell of span r ENDING at column g0 - 31 + x.  The dense generic
  // pass multiplies every tile element by a (possibly zero) coefficient, so elements of cells that do not
  // exist must read as 0 rather than as whatever an earlier batch left in the DP state.  A cell exists iff
  // its span is <= the limit of its tile column: L - i of the start column (left), left index of the end
  // column (right); -1 for padding columns and columns outside the batch.
  //
  // Row r of the right kernel's tile is the contiguous run of start columns g0 - 31 - r .. of the band array, i.e. it
  // begins one element further left per span.  So that the bulk-copy engine can fetch it (16-byte aligned source,
  // destination and size) the row is copied from the aligned address below its first element: its element x then
  // sits row_off(r) = (g0 - 31 - r) mod kAl elements into the shared-memory row, and the row stride has kAl spare
  // elements.  The offset repeats every 4 spans, which is the step of the dense pass.
  static constexpr int kAl = 16 / (int)sizeof(real);
  static PRIB_HD int tile_cols(bool left_side, int TXb) { return TXb + 32 + (left_side ? 0 : kAl); }
  static PRIB_HD int row_off(const Geo &ge, bool left_side, int r) {
    return left_side ? 0 : (int)((ge.g0 - 31 - r) & (long long)(kAl - 1));
  }
  static PRIB_HD int tile_col_limit(const Ctx &c, const Geo &ge, bool left_side, int x) {
    const long long col = left_side ? ge.g0 + x : ge.g0 - 31 + x;
    if (col < 0 || col >= c.NC) return -1;
    const int sq = c.col_seq[col];
    if (sq < 0) return -1;
    const int i = (int)(col - c.seq_off[sq]);
    return left_side ? c.seq_len[sq] - i : i;
  }
  static PRIB_HD real tile_elem(const Ctx &c, const Geo &ge, bool left_side, int arr, int r, int x, int limit) {
    if (r > limit) return (real)0;
    return c.ld(arr, r, left_side ? ge.g0 + x : ge.g0 - 31 + x - r);
  }

  // ---------------------------------------------------------------------------------------------
  // Dense, time-tiled generic pass.  Outer spans dp0 + k (k = 0..kTB-1) of one column are handled together:
  // tile row dp0 - S (inner span) holds, at offset DIR * u, the inner cell of the loops with kept strand
  // length u and loop size S + k for target k, so
  //     m[u] += row[DIR * u] * sum_k w[k] * conv[u][S + k - u]
  // with w[k] = Beta_stemO of the outer pair (0 when it does not close).  Where all kTB coefficients sit in
  // the saturated ninio zone (|u1 - u2| >= 6) the inner sum does not depend on u and is formed once per row.
  // All indices are compile-time: conv[][] entries are constant-bank operands.
  // ---------------------------------------------------------------------------------------------
  static PRIB_HD constexpr bool dn_valid(int u, int sum) { return sum <= kMaxLoop && sum - u >= 1 && sum >= 4 && !(u == 2 && sum == 4); }
  static PRIB_HD constexpr bool dn_sat(int u, int sum) { return (2 * u - sum >= 6) || (sum - 2 * u >= 6); }
  static PRIB_HD constexpr bool dn_allsat(int u, int S) {
    for (int k = 0; k < kTB; ++k)
      if (!dn_valid(u, S + k) || !dn_sat(u, S + k)) return false;
    return true;
  }
  static PRIB_HD constexpr bool dn_row_has_sat(int S, int ulo) {
    for (int u = ulo; u <= kMaxLoop; ++u)
      if (dn_allsat(u, S)) return true;
    return false;
  }
  static PRIB_HD constexpr bool dn_any(int u, int S) {
    for (int k = 0; k < kTB; ++k)
      if (dn_valid(u, S + k)) return true;
    return false;
  }

  template <int S, int U, int ULO, int DIR>
  static PRIB_HD void dense_cols(const real *row, const real *cv, const real (&w)[kTB], real qsat,
                                 real (&m)[kMaxLoop + 1]) {
    if constexpr (dn_any(U, S)) {
      real q;
      if constexpr (dn_allsat(U, S)) {
        q = qsat;
      } else {
        q = 0;
#pragma unroll
        for (int k = 0; k < kTB; ++k)
          if (dn_valid(U, S + k)) q += w[k] * cv[U * 32 + (S + k - U)];
      }
      m[U] += row[DIR * U] * q;
    }
    if constexpr (U < kMaxLoop) dense_cols<S, U + 1, ULO, DIR>(row, cv, w, qsat, m);
  }

  template <int S, int COLS, int ULO, int DIR>
  static PRIB_HD void dense_rows(const real *base, int cols, int dp0, int hi, const real *cv, const real (&w)[kTB],
                                 real (&m)[kMaxLoop + 1], const int (&o)[4]) {
    if (dp0 - S >= 5 && (S >= 1 || dp0 - S <= hi)) {  // uniform: inner spans below 5 hold no stems; the tile ends at span hi
      const real *row = base - S * (COLS > 0 ? COLS : cols);
      if constexpr (DIR < 0) row += o[S & 3];  // right tile: row_off() of span dp0 - S
      real qsat = 0;
      if constexpr (dn_row_has_sat(S, ULO)) {
#pragma unroll
        for (int k = 0; k < kTB; ++k) qsat += w[k] * cv[(S + k - 1) * 32 + 1];  // conv[sum - 1][1]: saturated, sum >= 8
      }
      dense_cols<S, ULO, ULO, DIR>(row, cv, w, qsat, m);
    }
    if constexpr (S < kMaxLoop) dense_rows<S + 1, COLS, ULO, DIR>(base, cols, dp0, hi, cv, w, m, o);
  }

#if defined(__CUDA_ARCH__)
  // ---- packed FP32 variant of the dense pass (FFMA2): strand lengths (U, U + 1) share one accumulator pair ----
  static __device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
    return p;
  }
  static __device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
  }
  static PRIB_HD constexpr bool dn_use(int u, int S, int ULO) { return u >= ULO && u <= kMaxLoop && dn_any(u, S); }

  // coefficient q of strand length U for source row S (see dense_cols), scalar
  template <int S, int U>
  static __device__ __forceinline__ float dense_q(const float *cv, const float (&w)[kTB], float qsat) {
    if constexpr (dn_allsat(U, S)) {
      return qsat;
    } else {
      float q = 0;
#pragma unroll
      for (int k = 0; k < kTB; ++k)
        if (dn_valid(U, S + k)) q += w[k] * cv[U * 32 + (S + k - U)];
      return q;
    }
  }

  template <int S, int U, int ULO, int DIR>
  static __device__ __forceinline__ void dense_cols2(const float *row, const float *cv, const float (&w)[kTB], float qsat,
                                                     unsigned long long (&mp)[kMaxLoop / 2 + 1]) {
    constexpr bool v0 = dn_use(U, S, ULO), v1 = dn_use(U + 1, S, ULO);
    if constexpr (v0 || v1) {
      unsigned long long qq;
      if constexpr (v0 && v1 && !dn_allsat(U, S) && !dn_allsat(U + 1, S)) {
        // both coefficients are 4-term sums over the targets: one FFMA2 per target, coefficient pairs from the
        // constant bank
        qq = 0;
#pragma unroll
        for (int k = 0; k < kTB; ++k) {
          const bool a = dn_valid(U, S + k), b = dn_valid(U + 1, S + k);
          if (a || b) ffma2_bcast(qq, w[k], g_convpair_f[U * 32 + S + k]);  // the table holds 0 where a loop does not exist
        }
      } else {
        float q0 = 0, q1 = 0;
        if constexpr (v0) q0 = dense_q<S, U>(cv, w, qsat);
        if constexpr (v1) q1 = dense_q<S, U + 1>(cv, w, qsat);
        qq = pack2(q0, q1);
      }
      ffma2(mp[U / 2], pack2(v0 ? row[DIR * U] : 0.f, v1 ? row[DIR * (U + 1)] : 0.f), qq);
    }
    if constexpr (U + 2 <= kMaxLoop) dense_cols2<S, U + 2, ULO, DIR>(row, cv, w, qsat, mp);
  }

  template <int S, int COLS, int ULO, int DIR>
  static __device__ __forceinline__ void dense_rows2(const float *base, int cols, int dp0, int hi, const float *cv,
                                                     const float (&w)[kTB], unsigned long long (&mp)[kMaxLoop / 2 + 1],
                                                     const int (&o)[4]) {
    if (dp0 - S >= 5 && (S >= 1 || dp0 - S <= hi)) {
      const float *row = base - S * (COLS > 0 ? COLS : cols);
      if constexpr (DIR < 0) row += o[S & 3];
      float qsat = 0;
      if constexpr (dn_row_has_sat(S, ULO)) {
#pragma unroll
        for (int k = 0; k < kTB; ++k) qsat += w[k] * cv[(S + k - 1) * 32 + 1];
      }
      dense_cols2<S, (ULO / 2) * 2, ULO, DIR>(row, cv, w, qsat, mp);
    }
    if constexpr (S < kMaxLoop) dense_rows2<S + 1, COLS, ULO, DIR>(base, cols, dp0, hi, cv, w, mp, o);
  }
#endif

  // Per-thread state carried from the generic pass (Alpha_stemI tile) to the bulge pass (Alpha_stemB tile).
  struct Strand {
    real w[kMaxLoop + 1];  // strand weights by strand length
    int cnt;               // number of closing spans listed by this thread
    bool active;
  };

  // thread t = left index i (column g0 + t); list: uint8 [W][TXb] scratch in shared memory.
  // COLS: compile-time row stride of the tile (0 = ge.cols at run time, host emulation); ULO: smallest strand
  // length that is accumulated (min(delta, 5): strands shorter than delta are never read).
  template <int COLS, int ULO>
  static PRIB_HD void left(const Ctx &c, const Geo &ge, const real *tile, uint8_t *list, int t, Strand &st) {
    const long long g = ge.g0 + t;
    st.cnt = 0;
    st.active = false;
    if (g >= c.NC) return;
    typename K::ColInfo ci;
    if (!K::col_info(c, g, ci)) return;
    st.active = true;
    const ST &T = *c.T;
    const real *cv = K::conv_tab(T);
    const int L = ci.L, i = ci.i, W = c.W, delta = c.delta, TXb = ge.TXb, cols = COLS > 0 ? COLS : ge.cols;
    const uint8_t *s = c.S + g;
    real (&ml)[kMaxLoop + 1] = st.w;
#pragma unroll
    for (int u = 0; u <= kMaxLoop; ++u) ml[u] = 0;
    int cnt = 0;
    // pass A (all lanes at the same span, widest first): the suffix sums over span of the hairpin-loop weights
    // (X_SUFH, restructured raccess.cpp:546-561 -- same Beta_stem loads as the list), the list of the closing spans
    // and, for delta == 2, the 2x1 / 2x2 special loops
    const int dpmax = i >= 1 ? imin(W - 1, L - 1 - i) : -1;
    const uint8_t s0 = s[0], s1 = s[1];  // a sequence column always has a successor (padding)
    real suf = 0;
    for (int dd0 = W; dd0 >= 4; dd0 -= 8) {
      real bv[8];  // 8 independent loads in flight (the list update is a serial chain on the loaded values)
      uint8_t sv[9];  // bases dd0 .. dd0 - 8 (a closing span needs its own and the next one)
#pragma unroll
      for (int k = 0; k < 8; ++k) bv[k] = (dd0 - k >= 4 && dd0 - k - 1 <= dpmax) ? c.ld(B_STEM, dd0 - k + 1, g - 1) : (real)0;
#pragma unroll
      for (int k = 0; k < 9; ++k) sv[k] = (dd0 - k >= 3 && dd0 - k - 1 <= dpmax) ? s[dd0 - k] : (uint8_t)0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int dd = dd0 - k, dp = dd - 1;
        if (dd < 4) break;
        const real bse = bv[k];
        if (bse != 0) {
          const int te = T.bp[s0][sv[k]];
          suf += bse * T.hpB[dd] * (dd - 1 != 3 ? T.e_mmH[te][s1][sv[k + 1]] : T.tau[te]);
          if (dp >= delta + 5) {
            list[cnt * TXb + t] = (uint8_t)dp;
            ++cnt;
            if (delta == 2) {
              const real bseO = c.ld(B_STEMO, dp + 2, g - 1), bseB = c.ld(B_STEMB, dp + 2, g - 1);
              if (dp - 5 >= 3) ml[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 1);
              if (dp - 5 >= 4) ml[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 2);
            }
          }
        }
        c.at(X_SUFH, dd, g) = suf;
      }
    }
    if (i >= 1) {
      const int kNoOff[4] = {0, 0, 0, 0};
      // pass B: generic interior loops out of the shared-memory tile, dense over groups of kTB outer spans
      // (the outer-pair weights of the next group are loaded while this group is being evaluated)
      real wn[kTB];
#pragma unroll
      for (int k = 0; k < kTB; ++k) wn[k] = (delta + 5 + k <= dpmax) ? c.ld(B_STEMO, delta + 5 + k + 2, g - 1) : (real)0;
#if defined(__CUDA_ARCH__)
      unsigned long long mp[kMaxLoop / 2 + 1];  // FP32 engine: packed accumulators (m[2j], m[2j + 1])
#pragma unroll
      for (int u = 0; u <= kMaxLoop / 2; ++u) mp[u] = 0;
#endif
      for (int dp0 = delta + 5; dp0 <= dpmax; dp0 += kTB) {
        real w[kTB];
#pragma unroll
        for (int k = 0; k < kTB; ++k) w[k] = wn[k];
#pragma unroll
        for (int k = 0; k < kTB; ++k) wn[k] = (dp0 + kTB + k <= dpmax) ? c.ld(B_STEMO, dp0 + kTB + k + 2, g - 1) : (real)0;
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(real) == 4) dense_rows2<kSmin, COLS, ULO, 1>(tile + (dp0 - 5) * cols + t, cols, dp0, W - 1, cv, w, mp, kNoOff);
        else
#endif
          dense_rows<kSmin, COLS, ULO, 1>(tile + (dp0 - 5) * cols + t, cols, dp0, W - 1, cv, w, ml, kNoOff);
      }
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4) {
#pragma unroll
        for (int u = 0; u <= kMaxLoop / 2; ++u) {
          float lo, hi;
          unpack2(mp[u], lo, hi);
          ml[2 * u] += lo;
          if (2 * u + 1 <= kMaxLoop) ml[2 * u + 1] += hi;
        }
      }
#endif
    }
    st.cnt = cnt;
  }

  // bulges (u2 = 0) of the left strand out of the Alpha_stemB tile, then the store of ML
  template <int COLS, int ULO>
  static PRIB_HD void left_bulge(const Ctx &c, const Geo &ge, const real *tile, const uint8_t *list, int t, Strand &st) {
    if (!st.active) return;
    const long long g = ge.g0 + t;
    const real *bu = K::bulge_tab(*c.T);
    const int delta = c.delta, TXb = ge.TXb, cols = COLS > 0 ? COLS : ge.cols;
    real (&ml)[kMaxLoop + 1] = st.w;
    real bseB_next = st.cnt > 0 ? c.ld(B_STEMB, list[t] + 2, g - 1) : (real)0;
    // sum over the closing spans first, the bulge factor bu[u1] once at the end (it does not depend on the span)
    real bsum[kMaxLoop + 1];
#pragma unroll
    for (int u1 = 0; u1 <= kMaxLoop; ++u1) bsum[u1] = 0;
    for (int k = 0; k < st.cnt; ++k) {
      const int dp = list[k * TXb + t];
      const real bseB = bseB_next;
      if (k + 1 < st.cnt) bseB_next = c.ld(B_STEMB, list[(k + 1) * TXb + t] + 2, g - 1);
      const int umax = imin(kMaxLoop, dp - 5);
      const real *base = tile + (dp - 5) * cols + t;
#pragma unroll
      for (int u1 = ULO; u1 <= kMaxLoop; ++u1)
        if (u1 <= umax) bsum[u1] += bseB * base[u1 - u1 * cols];  // cell (i+u1, j'), span dp-u1
    }
#pragma unroll
    for (int u1 = ULO; u1 <= kMaxLoop; ++u1)
      if (u1 >= delta) ml[u1] += bu[u1] * bsum[u1];
    real suf = 0;
#pragma unroll
    for (int u1 = kMaxLoop; u1 >= 2; --u1) {
      if (u1 >= delta) {
        c.at(X_ML, u1, g) = ml[u1];
        suf += ml[u1];
        c.at(X_MLS, u1, g) = suf;
      }
    }
  }

  // thread t = right end j' of the outer cell (column g0 + t)
  template <int COLS, int ULO>
  static PRIB_HD void right(const Ctx &c, const Geo &ge, const real *tile, uint8_t *list, int t, Strand &st) {
    const long long g2 = ge.g0 + t;
    st.cnt = 0;
    st.active = false;
    if (g2 >= c.NC) return;
    typename K::ColInfo ci;
    if (!K::col_info(c, g2, ci)) return;
    st.active = true;
    const ST &T = *c.T;
    const real *cv = K::conv_tab(T);
    const int L = ci.L, jp = ci.i, W = c.W, delta = c.delta, TXb = ge.TXb, cols = COLS > 0 ? COLS : ge.cols;
    real (&mr)[kMaxLoop + 1] = st.w;
#pragma unroll
    for (int u = 0; u <= kMaxLoop; ++u) mr[u] = 0;
    int cnt = 0;
    if (jp <= L - 1) {
      const int dpmax = imin(W - 1, jp - 1);  // i = jp - dp >= 1
      for (int dp0 = delta + 5; dp0 <= dpmax; dp0 += 8) {
        real bv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) bv[k] = (dp0 + k <= dpmax) ? c.ld(B_STEM, dp0 + k + 2, g2 - dp0 - k - 1) : (real)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int dp = dp0 + k;
          const long long g = g2 - dp;  // column of i
          const real bse = bv[k];
          if (bse == 0) continue;
          list[cnt * TXb + t] = (uint8_t)dp;
          ++cnt;
          if (delta == 2) {
            const uint8_t *s = c.S + g;
            const int te = T.bp[s[0]][s[dp + 1]];
            const real bseO = c.ld(B_STEMO, dp + 2, g - 1), bseB = c.ld(B_STEMB, dp + 2, g - 1);
            if (dp - 5 >= 3) mr[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 1, 2);
            if (dp - 5 >= 4) mr[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 2);
          }
        }
      }
      // row_off() of the inner rows dp0 - S by S mod 4 (dp0 advances in steps of kTB = 4: the same for every group)
      int o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = row_off(ge, false, delta + 5 - k);
      real wn[kTB];
#pragma unroll
      for (int k = 0; k < kTB; ++k)
        wn[k] = (delta + 5 + k <= dpmax) ? c.ld(B_STEMO, delta + 5 + k + 2, g2 - (delta + 5) - k - 1) : (real)0;
#if defined(__CUDA_ARCH__)
      unsigned long long mp[kMaxLoop / 2 + 1];
#pragma unroll
      for (int u = 0; u <= kMaxLoop / 2; ++u) mp[u] = 0;
#endif
      for (int dp0 = delta + 5; dp0 <= dpmax; dp0 += kTB) {
        real w[kTB];
#pragma unroll
        for (int k = 0; k < kTB; ++k) w[k] = wn[k];
#pragma unroll
        for (int k = 0; k < kTB; ++k)
          wn[k] = (dp0 + kTB + k <= dpmax) ? c.ld(B_STEMO, dp0 + kTB + k + 2, g2 - dp0 - kTB - k - 1) : (real)0;
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(real) == 4)
          dense_rows2<kSmin, COLS, ULO, -1>(tile + (dp0 - 5) * cols + t + 31, cols, dp0, W - 1, cv, w, mp, o);
        else
#endif
          dense_rows<kSmin, COLS, ULO, -1>(tile + (dp0 - 5) * cols + t + 31, cols, dp0, W - 1, cv, w, mr, o);
      }
#if defined(__CUDA_ARCH__)
      if constexpr (sizeof(real) == 4) {
#pragma unroll
        for (int u = 0; u <= kMaxLoop / 2; ++u) {
          float lo, hi;
          unpack2(mp[u], lo, hi);
          mr[2 * u] += lo;
          if (2 * u + 1 <= kMaxLoop) mr[2 * u + 1] += hi;
        }
      }
#endif
    }
    st.cnt = cnt;
  }

  // bulges (u1 = 0) of the right strand out of the end-indexed Alpha_stemB tile, then the store of MR
  template <int COLS, int ULO>
  static PRIB_HD void right_bulge(const Ctx &c, const Geo &ge, const real *tile, const uint8_t *list, int t, Strand &st) {
    if (!st.active) return;
    const long long g2 = ge.g0 + t;
    const real *bu = K::bulge_tab(*c.T);
    const int delta = c.delta, TXb = ge.TXb, cols = COLS > 0 ? COLS : ge.cols;
    real (&mr)[kMaxLoop + 1] = st.w;
    real bseB_next = st.cnt > 0 ? c.ld(B_STEMB, list[t] + 2, g2 - list[t] - 1) : (real)0;
    real bsum[kMaxLoop + 1];
#pragma unroll
    for (int u2 = 0; u2 <= kMaxLoop; ++u2) bsum[u2] = 0;
    for (int k = 0; k < st.cnt; ++k) {
      const int dp = list[k * TXb + t];
      const real bseB = bseB_next;
      if (k + 1 < st.cnt) {
        const int dn = list[(k + 1) * TXb + t];
        bseB_next = c.ld(B_STEMB, dn + 2, g2 - dn - 1);
      }
      const int umax = imin(kMaxLoop, dp - 5);
      const real *base = tile + (dp - 5) * cols + t + 31;
      const real *bo[4];  // + row_off() of span dp - u2, by u2 mod 4
#pragma unroll
      for (int k = 0; k < 4; ++k) bo[k] = base + row_off(ge, false, dp - k);
#pragma unroll
      for (int u2 = ULO; u2 <= kMaxLoop; ++u2)
        if (u2 <= umax) bsum[u2] += bseB * bo[u2 & 3][-u2 - u2 * cols];  // cell (i, j'-u2), span dp-u2
    }
#pragma unroll
    for (int u2 = ULO; u2 <= kMaxLoop; ++u2)
      if (u2 >= delta) mr[u2] += bu[u2] * bsum[u2];
    real suf = 0;
#pragma unroll
    for (int u2 = kMaxLoop; u2 >= 2; --u2) {
      if (u2 >= delta) {
        c.at(X_MR, u2, g2) = mr[u2];
        suf += mr[u2];
        c.at(X_MRS, u2, g2) = suf;
      }
    }
  }
};

}  // namespace prib
