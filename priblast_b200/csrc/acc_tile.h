// Tile-persistent span march (kernel set v2).
//
// One CTA owns TX consecutive columns of the batch and walks ALL spans for them, keeping the source rows
// of the 2-D stencils in shared-memory ring buffers, so a stencil term is one LDS + one FMA and nothing
// is re-read from HBM.  Because a cell (i, d) only depends on sub-intervals of [i, i+d] (inside) or on
// super-intervals within the band (outside), a CTA that also carries a halo of W+1 columns on the right
// (inside) or on the left (outside) can recompute everything its owned cells need without talking to
// its neighbours: no grid-wide synchronisation, one __syncthreads per span.
//
// The functions below are the per-thread, per-span bodies (thread t = local column t of the tile); the
// CUDA kernels and the host emulation both call them span by span with a barrier in between.
// Summation order inside a cell is identical to acc_core.h's v1 cell functions, so both give the same
// bits in the same precision (checked by tests/test_hostemu.py).
#pragma once
#include "acc_core.h"

namespace prib {

enum {
  kRingIn = 32,    // inside: source rows d-30..d-1 plus the row being written
  kRingOut = 34,   // outside: source rows d+1..d+32 plus the row being written
  kRingStem = 8,
  kRingSE = 4,
  kTileRows = 80,  // shared-memory rows of TC reals per CTA (both passes)
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

  // ---- shared-memory carve-up (same 80 rows for both passes) -----------------------------------
  struct InSmem {
    real *stemI, *stemB, *stem, *se, *mu, *m2;
    const uint8_t *S;  // bases of local columns 0 .. TC+3
  };
  static PRIB_HD InSmem carve_in(real *base, int TC, const uint8_t *S) {
    InSmem s;
    s.stemI = base;
    s.stemB = s.stemI + kRingIn * TC;
    s.stem = s.stemB + kRingIn * TC;
    s.se = s.stem + kRingStem * TC;
    s.mu = s.se + kRingSE * TC;
    s.m2 = s.mu + 2 * TC;
    s.S = S;
    return s;
  }
  struct OutSmem {
    real *stemO, *stemB, *stem, *mu, *m2;
  };
  static PRIB_HD OutSmem carve_out(real *base, int TC) {
    OutSmem s;
    s.stemO = base;
    s.stemB = s.stemO + kRingOut * TC;
    s.stem = s.stemB + kRingOut * TC;
    s.mu = s.stem + kRingStem * TC;
    s.m2 = s.mu + 2 * TC;
    return s;
  }

  // ---------------------------------------------------------------------------------------------
  // inside: cell (i, i + d) of local column t.  scrM1/scrM2: per-CTA global scratch, [(W+2)][TC].
  // ---------------------------------------------------------------------------------------------
  static PRIB_HD void inside_span(const Ctx &c, const Geo &ge, const InSmem &sm, real *scrM1, real *scrM2,
                                  int t, const ColState &cs, int d) {
    const int TC = ge.TC;
    const ST &T = *c.T;
    const real *cv = K::conv_tab(T), *bu = K::bulge_tab(T);
    const int L = cs.L, i = cs.i, j = i + d;
    real stem = 0, stemI = 0, stemB = 0, stemD = 0, se = 0, mu = 0, m1 = 0, m2 = 0;
    const bool live = i >= 0 && j <= L && t + d <= TC - 1;
    if (live) {
      const uint8_t *s = sm.S + t;
      const int si = s[0], si1 = s[1], sj = s[d], sj1 = s[d + 1];
      const int tp = T.bp[si1][sj];
      if (tp) {
        const int t2 = T.bp[s[2]][s[d - 1]];
        stem = T.k2 * (sm.se[((d - 2) & (kRingSE - 1)) * TC + t + 1] +
                       sm.stem[((d - 2) & (kRingStem - 1)) * TC + t + 1] * T.e_stack[tp][T.rt[t2]]);
      }
      real mb = 0;
      for (int m = 5; m <= d - 5; ++m) mb += scrM1[m * TC + t] * scrM2[(d - m) * TC + t + m];
      mb *= T.inv_cA;
      stemD = tp ? stem * K::e_dangle(T, tp, i > 0, si, j < L, sj1) : 0;
      m2 = stemD * T.e_mlintern + sm.m2[((d - 1) & 1) * TC + t] * T.e_mlbase;
      m1 = m2 + mb;
      mu = sm.mu[((d - 1) & 1) * TC + t + 1] * T.e_mlbase + mb;

      const int te = (j != L) ? T.bp[si][sj1] : 0;
      if (te) {
        real acc = T.e_hairpin[d] * (d != 3 ? T.e_mmH[te][si1][sj] : T.tau[te]);
        const int smax = imin(kMaxLoop, d - 5);
        const real *st1 = sm.stem + ((d - 1) & (kRingStem - 1)) * TC + t;
        const real *st2 = sm.stem + ((d - 2) & (kRingStem - 1)) * TC + t;
        const real *st3 = sm.stem + ((d - 3) & (kRingStem - 1)) * TC + t;
        const real *st4 = sm.stem + ((d - 4) & (kRingStem - 1)) * TC + t;
        if (smax >= 1) {
          acc += bu[1] * (st1[1] * T.e_stack[te][T.rt[T.bp[s[2]][sj]]] +
                                 st1[0] * T.e_stack[te][T.rt[T.bp[si1][s[d - 1]]]]);
        }
        if (smax >= 2) {
          const int t2 = T.rt[T.bp[s[2]][s[d - 1]]];
          acc += st2[1] * c.e_int11[idx11(te, t2, si1, sj)];
          real bs = 0;
#pragma unroll
          for (int u = 2; u <= kMaxLoop; ++u) {
            if (u <= smax) {
              const real *row = sm.stemB + ((d - u) & (kRingIn - 1)) * TC + t;
              bs += bu[u] * (row[u] + row[0]);
            }
          }
          acc += T.tau[te] * bs;
        }
        if (smax >= 3) {
          const int ta = T.rt[T.bp[s[2]][s[d - 2]]];
          acc += st3[1] * c.e_int21[idx21(te, ta, si1, s[d - 1], sj)];
          const int tb = T.rt[T.bp[s[3]][s[d - 1]]];
          acc += st3[2] * c.e_int21[idx21(tb, te, sj, si1, s[2])];
        }
        if (smax >= 4) {
          const int tc = T.rt[T.bp[s[3]][s[d - 2]]];
          acc += st4[2] * c.e_int22[idx22(te, tc, si1, s[2], s[d - 1], sj)];
          // fully unrolled: every coefficient is a compile-time offset into constant memory, so a
          // stencil term is exactly one LDS and one FMA
          real gs = 0;
#pragma unroll
          for (int sum = 4; sum <= kMaxLoop; ++sum) {
            if (sum <= smax) {
              const real *row = sm.stemI + ((d - sum) & (kRingIn - 1)) * TC + t;
#pragma unroll
              for (int u1 = 1; u1 < sum; ++u1) gs += cv[u1 * 32 + sum - u1] * row[u1];
            }
          }
          acc += T.e_mmI[te][si1][sj] * gs;
        }
        const int tt = T.rt[te];
        acc += mu * T.e_mlclose * T.e_d3[tt][si1] * T.e_d5[tt][sj];
        se = acc;
      }
      if (tp) {
        stemI = stem * T.e_mmI[T.rt[tp]][sj1][si];
        stemB = stem * T.tau[tp];
      }
    }
    // every thread refreshes its ring slots every span (zeros where the cell does not exist)
    sm.stemI[(d & (kRingIn - 1)) * TC + t] = stemI;
    sm.stemB[(d & (kRingIn - 1)) * TC + t] = stemB;
    sm.stem[(d & (kRingStem - 1)) * TC + t] = stem;
    sm.se[(d & (kRingSE - 1)) * TC + t] = se;
    sm.mu[(d & 1) * TC + t] = mu;
    sm.m2[(d & 1) * TC + t] = m2;
    scrM1[d * TC + t] = m1;
    scrM2[d * TC + t] = m2;
    // persistent outputs: owned columns only
    if (t < ge.TX && i >= 0 && j <= L) {
      const long long g = ge.g0 + t;
      c.at(A_STEM, d, g) = stem;
      c.at(A_STEMI, d, g) = stemI;
      c.at(A_STEMB, d, g) = stemB;
      c.at(A_STEMD, d, g) = stemD;
      c.at(A_STEMDE, d, g + d) = stemD;
      c.at(A_MULTI, d, g) = mu;
      c.at(A_MULTI1, d, g) = m1;
      c.at(A_MULTI2, d, g) = m2;
      if (!(K::in_safe_range(stem) && K::in_safe_range(se) && K::in_safe_range(mu) && K::in_safe_range(m1) &&
            K::in_safe_range(m2)))
        c.flags[cs.sq] = 1;
    }
  }

  // ---------------------------------------------------------------------------------------------
  // outside: cell (p, p + d) of local column t; global column g = g0 - H + t.
  // scrBif: per-CTA global scratch for Beta_multibif, [(W+4)][TC].
  // ---------------------------------------------------------------------------------------------
  static PRIB_HD int wrap_out(int r) { return r >= kRingOut ? r - kRingOut : r; }

  static PRIB_HD void outside_span(const Ctx &c, const Geo &ge, const OutSmem &sm, real *scrBif, int t,
                                   const ColState &cs, int d, int slot_d /* = d % kRingOut */) {
    const int TC = ge.TC, W = c.W;
    const ST &T = *c.T;
    const real *cv = K::conv_tab(T), *bu = K::bulge_tab(T);
    const long long g = ge.g0 - ge.H + t;
    const int L = cs.L, p = cs.i, q = p + d;
    real bstem = 0, bstemO = 0, bstemB = 0, bmulti = 0, bmulti2 = 0, bmbif = 0;
    // a halo cell is exact iff its end reaches the owned region (all its super-intervals are in the tile)
    const bool live = p >= 0 && q <= L && t + d >= ge.H;
    if (live) {
      const uint8_t *s = c.S + g;  // the right end q = p + d can lie beyond the tile: bases come from global
      const int sp = s[0], sp1 = s[1], sq_ = s[d], sq1 = s[d + 1];
      const bool inner = (p != 0 && q != L);
      const int te = inner ? T.bp[sp][sq1] : 0;
      const real *b2 = sm.stem + ((d + 2) & (kRingStem - 1)) * TC + t;
      const real bse = (inner && d + 2 <= W + 1) ? b2[-1] : 0;  // Beta_stemend(p,q), :277-279
      if (inner) {
        const int tt = T.rt[te];
        bmulti = (d + 1 <= W + 1 ? sm.mu[((d + 1) & 1) * TC + t - 1] * T.e_mlbase : (real)0) +
                 T.k2 * bse * T.e_mlclose * T.e_d3[tt][sp1] * T.e_d5[tt][sq_];
        real bm1 = 0;
        const int m1max = imin(L - q, W - d);
        for (int m = 5; m <= m1max; ++m) bm1 += scrBif[(d + m) * TC + t] * c.ld(A_MULTI2, m, g + d);
        bm1 *= T.inv_cA;
        real ks = 0;
        const int m2max = imin(p, W - d);
        for (int m = 5; m <= m2max; ++m) ks += scrBif[(d + m) * TC + t - m] * c.ld(A_MULTI1, m, g - m);
        bmulti2 = bm1 + sm.m2[((d + 1) & 1) * TC + t] * T.e_mlbase + ks * T.inv_cA;
        bmbif = bm1 + bmulti;
      }
      const int t2 = T.bp[sp1][sq_];
      if (t2) {
        const int t2r = T.rt[t2];
        const real dang = K::e_dangle(T, t2, p > 0, sp, q < L, sq1);
        const real base = (real)exp(c.lao[g] + c.lbo[g + d] - c.lao[cs.zcol]) * dang * T.sB[d];
        real ls = 0;
        const int smax = imin(kMaxLoop, W - 1 - d);
        if (smax >= 0) ls += bse * T.e_stack[te][t2r];
        const real *b3 = sm.stem + ((d + 3) & (kRingStem - 1)) * TC + t;
        const real *b4 = sm.stem + ((d + 4) & (kRingStem - 1)) * TC + t;
        const real *b5 = sm.stem + ((d + 5) & (kRingStem - 1)) * TC + t;
        const real *b6 = sm.stem + ((d + 6) & (kRingStem - 1)) * TC + t;
        if (smax >= 1) {
          const int ta = T.bp[s[-1]][sq1];
          const int tb = T.bp[sp][s[d + 2]];
          ls += bu[1] * (b3[-2] * T.e_stack[ta][t2r] + b3[-1] * T.e_stack[tb][t2r]);
        }
        if (smax >= 2) {
          const int to = T.bp[s[-1]][s[d + 2]];
          ls += b4[-2] * c.e_int11[idx11(to, t2r, sp, sq1)];
          real bs = 0;
          int slot = wrap_out(slot_d + 4);
#pragma unroll
          for (int u = 2; u <= kMaxLoop; ++u) {
            if (u <= smax) {
              const real *row = sm.stemB + slot * TC + t - 1;
              bs += bu[u] * (row[-u] + row[0]);
              slot = wrap_out(slot + 1);
            }
          }
          ls += T.tau[t2r] * bs;
        }
        if (smax >= 3) {
          const int ta = T.bp[s[-1]][s[d + 3]];
          ls += b5[-2] * c.e_int21[idx21(ta, t2r, sp, sq1, s[d + 2])];
          const int tb = T.bp[s[-2]][s[d + 2]];
          ls += b5[-3] * c.e_int21[idx21(t2r, tb, sq1, s[-1], sp)];
        }
        if (smax >= 4) {
          const int tc = T.bp[s[-2]][s[d + 3]];
          ls += b6[-3] * c.e_int22[idx22(tc, t2r, s[-1], sp, sq1, s[d + 2])];
          real gs = 0;
          int slot = wrap_out(slot_d + 6);
#pragma unroll
          for (int sum = 4; sum <= kMaxLoop; ++sum) {
            if (sum <= smax) {
              const real *row = sm.stemO + slot * TC + t - 1;
#pragma unroll
              for (int u1 = 1; u1 < sum; ++u1) gs += cv[u1 * 32 + sum - u1] * row[-u1];
              slot = wrap_out(slot + 1);
            }
          }
          ls += T.e_mmI[t2r][sq1][sp] * gs;
        }
        bstem = base + T.k2 * ls + bmulti2 * T.e_mlintern * dang;
        bstemO = bstem * T.e_mmI[t2][s[2]][s[d - 1]];
        bstemB = bstem * T.tau[t2];
      }
    }
    sm.stemO[slot_d * TC + t] = bstemO;
    sm.stemB[slot_d * TC + t] = bstemB;
    sm.stem[(d & (kRingStem - 1)) * TC + t] = bstem;
    sm.mu[(d & 1) * TC + t] = bmulti;
    sm.m2[(d & 1) * TC + t] = bmulti2;
    scrBif[d * TC + t] = bmbif;
    if (t >= ge.H && p >= 0 && q <= L) {
      c.at(B_STEM, d, g) = bstem;
      c.at(B_STEMO, d, g) = bstemO;
      c.at(B_STEMB, d, g) = bstemB;
      c.at(B_MULTI, d, g) = bmulti;
      c.at(B_MULTI2, d, g) = bmulti2;
      if (!(K::in_safe_range(bstem) && K::in_safe_range(bmulti) && K::in_safe_range(bmulti2) &&
            K::in_safe_range(bmbif)))
        c.flags[cs.sq] = 1;
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

  struct Geo {
    long long g0;  // first owned column
    int TXb;       // owned columns = threads
    int cols;      // TXb + 32 (row stride of the tile; multiple of 32)
    int rows;      // W - 5: spans 5 .. W-1
  };

  // element of the start-indexed tile: span r, global column g0 + x        (left kernel)
  static PRIB_HD real load_left(const Ctx &c, const Geo &ge, int r, int x) {
    const long long col = ge.g0 + x;
    return col < c.NC ? c.ld(A_STEMI, r, col) : (real)0;
  }
  // element of the end-indexed tile: span r, END column g0 - 31 + x        (right kernel)
  static PRIB_HD real load_right(const Ctx &c, const Geo &ge, int r, int x) {
    const long long col = ge.g0 - 31 + x - r;
    return (col >= 0 && col < c.NC) ? c.ld(A_STEMI, r, col) : (real)0;
  }

  // thread t = left index i (column g0 + t); list: uint8 [W][TXb] scratch in shared memory
  static PRIB_HD void left(const Ctx &c, const Geo &ge, const real *tile, uint8_t *list, int t) {
    const long long g = ge.g0 + t;
    if (g >= c.NC) return;
    typename K::ColInfo ci;
    if (!K::col_info(c, g, ci)) return;
    const ST &T = *c.T;
    const real *cv = K::conv_tab(T), *bu = K::bulge_tab(T);
    const int L = ci.L, i = ci.i, W = c.W, delta = c.delta, TXb = ge.TXb, cols = ge.cols;
    const uint8_t *s = c.S + g;
    real ml[kMaxLoop + 1];
#pragma unroll
    for (int u = 0; u <= kMaxLoop; ++u) ml[u] = 0;
    int cnt = 0;
    if (i >= 1) {
      const int dpmax = imin(W - 1, L - 1 - i);
      // pass A (all lanes at the same span): list the closing spans, add bulges (u2 = 0) and, for
      // delta == 2, the 2x1 / 2x2 special loops
      for (int dp = delta + 5; dp <= dpmax; ++dp) {
        const real bse = c.ld(B_STEM, dp + 2, g - 1);
        if (bse == 0) continue;
        list[cnt * TXb + t] = (uint8_t)dp;
        ++cnt;
        const real bseB = c.ld(B_STEMB, dp + 2, g - 1);
        const int umax = imin(kMaxLoop, dp - 5);
#pragma unroll
        for (int u1 = 2; u1 <= kMaxLoop; ++u1)
          if (u1 >= delta && u1 <= umax) ml[u1] += bseB * bu[u1] * c.ld(A_STEMB, dp - u1, g + u1);
        if (delta == 2) {
          const int te = T.bp[s[0]][s[dp + 1]];
          const real bseO = c.ld(B_STEMO, dp + 2, g - 1);
          if (dp - 5 >= 3) ml[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 1);
          if (dp - 5 >= 4) ml[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 2);
        }
      }
      // pass B (each lane at its own span): generic interior loops out of the shared-memory tile
      for (int k = 0; k < cnt; ++k) {
        const int dp = list[k * TXb + t];
        const real bseO = c.ld(B_STEMO, dp + 2, g - 1);
        const int smax = imin(kMaxLoop, dp - 5);
#pragma unroll
        for (int u1 = 2; u1 <= kMaxLoop - 1; ++u1) {
          if (u1 >= delta && u1 <= smax - 1) {
            real a = 0;
            const real *p = tile + (dp - u1 - 1 - 5) * cols + t + u1;  // inner cell (i+u1, j'-u2) at u2 = 1
            const int n2 = smax - u1;
            for (int u2 = 1; u2 <= n2; ++u2) {
              a += cv[u1 * 32 + u2] * *p;
              p -= cols;
            }
            ml[u1] += bseO * a;
          }
        }
      }
    }
#pragma unroll
    for (int u1 = 2; u1 <= kMaxLoop; ++u1)
      if (u1 >= delta) c.at(X_ML, u1, g) = ml[u1];
  }

  // thread t = right end j' of the outer cell (column g0 + t)
  static PRIB_HD void right(const Ctx &c, const Geo &ge, const real *tile, uint8_t *list, int t) {
    const long long g2 = ge.g0 + t;
    if (g2 >= c.NC) return;
    typename K::ColInfo ci;
    if (!K::col_info(c, g2, ci)) return;
    const ST &T = *c.T;
    const real *cv = K::conv_tab(T), *bu = K::bulge_tab(T);
    const int L = ci.L, jp = ci.i, W = c.W, delta = c.delta, TXb = ge.TXb, cols = ge.cols;
    real mr[kMaxLoop + 1];
#pragma unroll
    for (int u = 0; u <= kMaxLoop; ++u) mr[u] = 0;
    int cnt = 0;
    if (jp <= L - 1) {
      const int dpmax = imin(W - 1, jp - 1);  // i = jp - dp >= 1
      for (int dp = delta + 5; dp <= dpmax; ++dp) {
        const long long g = g2 - dp;  // column of i
        const real bse = c.ld(B_STEM, dp + 2, g - 1);
        if (bse == 0) continue;
        list[cnt * TXb + t] = (uint8_t)dp;
        ++cnt;
        const real bseB = c.ld(B_STEMB, dp + 2, g - 1);
        const int umax = imin(kMaxLoop, dp - 5);
#pragma unroll
        for (int u2 = 2; u2 <= kMaxLoop; ++u2)
          if (u2 >= delta && u2 <= umax) mr[u2] += bseB * bu[u2] * c.ld(A_STEMB, dp - u2, g);
        if (delta == 2) {
          const uint8_t *s = c.S + g;
          const int te = T.bp[s[0]][s[dp + 1]];
          const real bseO = c.ld(B_STEMO, dp + 2, g - 1);
          if (dp - 5 >= 3) mr[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 1, 2);
          if (dp - 5 >= 4) mr[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 2);
        }
      }
      for (int k = 0; k < cnt; ++k) {
        const int dp = list[k * TXb + t];
        const real bseO = c.ld(B_STEMO, dp + 2, g2 - dp - 1);
        const int smax = imin(kMaxLoop, dp - 5);
#pragma unroll
        for (int u2 = 2; u2 <= kMaxLoop - 1; ++u2) {
          if (u2 >= delta && u2 <= smax - 1) {
            real a = 0;
            const real *p = tile + (dp - u2 - 1 - 5) * cols + t + 31 - u2;  // inner cell ends at j' - u2; u1 = 1
            const int n1 = smax - u2;
            for (int u1 = 1; u1 <= n1; ++u1) {
              a += cv[u2 * 32 + u1] * *p;  // conv is symmetric: contiguous walk through constant memory
              p -= cols;
            }
            mr[u2] += bseO * a;
          }
        }
      }
    }
#pragma unroll
    for (int u2 = 2; u2 <= kMaxLoop; ++u2)
      if (u2 >= delta) c.at(X_MR, u2, g2) = mr[u2];
  }
};

}  // namespace prib
