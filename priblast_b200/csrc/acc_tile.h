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

  static PRIB_HD real gsel(int k, real g0, real g1, real g2, real g3, real g4, real g5, real g6) {
    return k == 0 ? g0 : k == 1 ? g1 : k == 2 ? g2 : k == 3 ? g3 : k == 4 ? g4 : k == 5 ? g5 : g6;
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

  // R consecutive reals from shared memory with one vector load (p aligned to R * sizeof(real))
  template <int R>
  static PRIB_HD void load_vec(const real *p, real (&v)[R]) {
#if defined(__CUDA_ARCH__)
    if (R == 4 && sizeof(real) == 4) {
      const float4 q = *reinterpret_cast<const float4 *>(p);
      v[0] = (real)q.x; v[1] = (real)q.y; v[2 % R] = (real)q.z; v[3 % R] = (real)q.w;
      return;
    }
    if (R == 2 && sizeof(real) == 4) {
      const float2 q = *reinterpret_cast<const float2 *>(p);
      v[0] = (real)q.x; v[1 % R] = (real)q.y;
      return;
    }
    if (R == 2 && sizeof(real) == 8) {
      const double2 q = *reinterpret_cast<const double2 *>(p);
      v[0] = (real)q.x; v[1 % R] = (real)q.y;
      return;
    }
#endif
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = p[r];
  }

  // ---------------------------------------------------------------------------------------------
  // inside: thread tq owns the R consecutive local columns tq*R .. tq*R+R-1 (cells (i, i+d)).
  // scrM1/scrM2: per-CTA global scratch, [(W+4)][TC].  The generic-loop stencil is evaluated jointly for
  // the R cells: every source element is loaded once (vector LDS) and used by up to R targets, each
  // target still adds its terms in ascending u1, so the result is bit-identical for every R.
  // ---------------------------------------------------------------------------------------------
  // TCC: compile-time tile width (0 = take ge.TC at run time, used by the host emulation): with a constant
  // row stride every ring / scratch row offset becomes an immediate of the load instruction.
  template <int R, int TCC = 0>
  static PRIB_HD void inside_span(const Ctx &c, const ST &T, const Geo &ge, const InSmem &sm, real *scrM1,
                                  real *scrM2, int tq, const ColState (&cs)[R], int d) {
    const int TC = TCC > 0 ? TCC : ge.TC, c0 = tq * R;
    const real *bu = K::bulge_tab(T), *cf = K::cf_tab(T);
    const real g0 = T.cg[0], g1 = T.cg[1], g2 = T.cg[2], g3 = T.cg[3], g4 = T.cg[4], g5 = T.cg[5], g6 = T.cg[6];
    real stem[R], stemI[R], stemB[R], stemD[R], se[R], mu[R], m1[R], m2[R], acc[R], gs[R];
    int te[R];
    bool any = false;
    const int smax = imin(kMaxLoop, d - 5);  // u1 + u2 <= smax keeps the inner span >= 5
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = c0 + r;
      const int L = cs[r].L, i = cs[r].i, j = i + d;
      stem[r] = stemI[r] = stemB[r] = stemD[r] = se[r] = mu[r] = m1[r] = m2[r] = acc[r] = gs[r] = 0;
      te[r] = 0;
      const bool live = i >= 0 && j <= L && t + d <= TC - 1;
      if (!live) continue;
      const uint8_t *s = sm.S + t;
      const int si = s[0], si1 = s[1], sj = s[d], sj1 = s[d + 1];
      const int tp = T.bp[si1][sj];
      if (tp) {
        const int t2 = T.bp[s[2]][s[d - 1]];
        stem[r] = T.k2 * (sm.se[((d - 2) & (kRingSE - 1)) * TC + t + 1] +
                          sm.stem[((d - 2) & (kRingStem - 1)) * TC + t + 1] * T.e_stack[tp][T.rt[t2]]);
      }
      real mb = 0;
      {
        const real *pa = scrM1 + 5 * TC + t, *pb = scrM2 + (d - 5) * TC + t + 5;
        int m = 5;
        for (; m + 7 <= d - 5; m += 8) {  // 16 loads in flight; same order of additions as the plain loop
          real av[8], bv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            av[k] = pa[k * TC];
            bv[k] = pb[-k * (TC - 1)];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) mb += av[k] * bv[k];
          pa += 8 * TC;
          pb -= 8 * (TC - 1);
        }
        for (; m <= d - 5; ++m) {
          mb += pa[0] * pb[0];
          pa += TC;
          pb -= TC - 1;
        }
      }
      mb *= T.inv_cA;
      stemD[r] = tp ? stem[r] * K::e_dangle(T, tp, i > 0, si, j < L, sj1) : 0;
      m2[r] = stemD[r] * T.e_mlintern + sm.m2[((d - 1) & 1) * TC + t] * T.e_mlbase;
      m1[r] = m2[r] + mb;
      mu[r] = sm.mu[((d - 1) & 1) * TC + t + 1] * T.e_mlbase + mb;
      if (tp) {
        stemI[r] = stem[r] * T.e_mmI[T.rt[tp]][sj1][si];
        stemB[r] = stem[r] * T.tau[tp];
      }
      te[r] = (j != L) ? T.bp[si][sj1] : 0;
      if (te[r]) {
        any = true;
        real a = T.e_hairpin[d] * (d != 3 ? T.e_mmH[te[r]][si1][sj] : T.tau[te[r]]);
        const real *st1 = sm.stem + ((d - 1) & (kRingStem - 1)) * TC + t;
        const real *st2 = sm.stem + ((d - 2) & (kRingStem - 1)) * TC + t;
        const real *st3 = sm.stem + ((d - 3) & (kRingStem - 1)) * TC + t;
        const real *st4 = sm.stem + ((d - 4) & (kRingStem - 1)) * TC + t;
        if (smax >= 1) {
          a += bu[1] * (st1[1] * T.e_stack[te[r]][T.rt[T.bp[s[2]][sj]]] +
                        st1[0] * T.e_stack[te[r]][T.rt[T.bp[si1][s[d - 1]]]]);
        }
        if (smax >= 2) {
          const int t2 = T.rt[T.bp[s[2]][s[d - 1]]];
          a += st2[1] * c.e_int11[idx11(te[r], t2, si1, sj)];
          real bs = 0;
#pragma unroll
          for (int u = 2; u <= kMaxLoop; ++u) {
            if (u <= smax) {
              const real *row = sm.stemB + ((d - u) & (kRingIn - 1)) * TC + t;
              bs += bu[u] * (row[u] + row[0]);
            }
          }
          a += T.tau[te[r]] * bs;
        }
        if (smax >= 3) {
          const int ta = T.rt[T.bp[s[2]][s[d - 2]]];
          a += st3[1] * c.e_int21[idx21(te[r], ta, si1, s[d - 1], sj)];
          const int tb = T.rt[T.bp[s[3]][s[d - 1]]];
          a += st3[2] * c.e_int21[idx21(tb, te[r], sj, si1, s[2])];
        }
        if (smax >= 4) {
          const int tc = T.rt[T.bp[s[3]][s[d - 2]]];
          a += st4[2] * c.e_int22[idx22(te[r], tc, si1, s[2], s[d - 1], sj)];
        }
        acc[r] = a;
      }
    }
    // generic interior loops: joint stencil over A_STEMI, fully unrolled.  Local column c0 + x of row
    // d - sum is source u1 = x - r of target r; weights depend on |u1 - u2| only (7 register values)
    // times one factor per loop size, so a term is one FMA and 1/R of a vector LDS.
    if (any && smax >= 4) {
#pragma unroll
      for (int sum = 4; sum <= kMaxLoop; ++sum) {
        if (sum <= smax) {
          const real *row = sm.stemI + ((d - sum) & (kRingIn - 1)) * TC + c0;
          real rs[R];
#pragma unroll
          for (int r = 0; r < R; ++r) rs[r] = 0;
#pragma unroll
          for (int xb = 0; xb <= sum + R - 2; xb += R) {
            real v[R];
            load_vec<R>(row + xb, v);
#pragma unroll
            for (int k = 0; k < R; ++k) {
              const int x = xb + k;
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const int u1 = x - r;
                if (u1 >= 1 && u1 <= sum - 1 && !(sum == 4 && u1 == 2))
                  rs[r] += gsel(K::gidx(u1, sum), g0, g1, g2, g3, g4, g5, g6) * v[k];
              }
            }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) gs[r] += cf[sum] * rs[r];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = c0 + r;
      const int L = cs[r].L, i = cs[r].i, j = i + d;
      if (te[r]) {
        const uint8_t *s = sm.S + t;
        const int si1 = s[1], sj = s[d];
        real a = acc[r];
        if (smax >= 4) a += T.e_mmI[te[r]][si1][sj] * gs[r];
        const int tt = T.rt[te[r]];
        a += mu[r] * T.e_mlclose * T.e_d3[tt][si1] * T.e_d5[tt][sj];
        se[r] = a;
      }
      // every column refreshes its ring slots every span (zeros where the cell does not exist)
      sm.stemI[(d & (kRingIn - 1)) * TC + t] = stemI[r];
      sm.stemB[(d & (kRingIn - 1)) * TC + t] = stemB[r];
      sm.stem[(d & (kRingStem - 1)) * TC + t] = stem[r];
      sm.se[(d & (kRingSE - 1)) * TC + t] = se[r];
      sm.mu[(d & 1) * TC + t] = mu[r];
      sm.m2[(d & 1) * TC + t] = m2[r];
      scrM1[d * TC + t] = m1[r];
      scrM2[d * TC + t] = m2[r];
      // persistent outputs: owned columns only
      if (t < ge.TX && i >= 0 && j <= L) {
        const long long g = ge.g0 + t;
        c.at(A_STEM, d, g) = stem[r];
        c.at(A_STEMI, d, g) = stemI[r];
        c.at(A_STEMB, d, g) = stemB[r];
        c.at(A_STEMD, d, g) = stemD[r];
        c.at(A_STEMDE, d, g + d) = stemD[r];
        c.at(A_MULTI, d, g) = mu[r];
        c.at(A_MULTI1, d, g) = m1[r];
        c.at(A_MULTI2, d, g) = m2[r];
        if (!(K::in_safe_range(stem[r]) && K::in_safe_range(se[r]) && K::in_safe_range(mu[r]) &&
              K::in_safe_range(m1[r]) && K::in_safe_range(m2[r])))
          c.flags[cs[r].sq] = 1;
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // outside: thread tq owns local columns tq*R .. tq*R+R-1; cell (p, p + d); global column g0 - H + t.
  // scrBif: per-CTA global scratch for Beta_multibif, [(W+4)][TC].
  // ---------------------------------------------------------------------------------------------
  static PRIB_HD int wrap_out(int r) { return r >= kRingOut ? r - kRingOut : r; }

  template <int R, int TCC = 0>
  static PRIB_HD void outside_span(const Ctx &c, const ST &T, const Geo &ge, const OutSmem &sm, real *scrBif,
                                   int tq, const ColState (&cs)[R], int d, int slot_d /* = d % kRingOut */) {
    const int TC = TCC > 0 ? TCC : ge.TC, W = c.W, c0 = tq * R;
    const real *bu = K::bulge_tab(T), *cf = K::cf_tab(T);
    const real g0 = T.cg[0], g1 = T.cg[1], g2 = T.cg[2], g3 = T.cg[3], g4 = T.cg[4], g5 = T.cg[5], g6 = T.cg[6];
    real bstem[R], bstemO[R], bstemB[R], bmulti[R], bmulti2[R], bmbif[R], base[R], ls[R], gs[R], dang[R];
    int t2v[R];
    bool any = false;
    const int smax = imin(kMaxLoop, W - 1 - d);  // source row d + sum + 2 <= W + 1
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = c0 + r;
      const long long g = ge.g0 - ge.H + t;
      const int L = cs[r].L, p = cs[r].i, q = p + d;
      bstem[r] = bstemO[r] = bstemB[r] = bmulti[r] = bmulti2[r] = bmbif[r] = base[r] = ls[r] = gs[r] = dang[r] = 0;
      t2v[r] = 0;
      // a halo cell is exact iff its end reaches the owned region (all its super-intervals are in the tile)
      const bool live = p >= 0 && q <= L && t + d >= ge.H;
      if (!live) continue;
      const uint8_t *s = c.S + g;  // the right end q = p + d can lie beyond the tile: bases come from global
      const int sp = s[0], sp1 = s[1], sq_ = s[d], sq1 = s[d + 1];
      const bool inner = (p != 0 && q != L);
      const int te = inner ? T.bp[sp][sq1] : 0;
      const real *b2 = sm.stem + ((d + 2) & (kRingStem - 1)) * TC + t;
      const real bse = (inner && d + 2 <= W + 1) ? b2[-1] : 0;  // Beta_stemend(p,q), :277-279
      if (inner) {
        const int tt = T.rt[te];
        bmulti[r] = (d + 1 <= W + 1 ? sm.mu[((d + 1) & 1) * TC + t - 1] * T.e_mlbase : (real)0) +
                    T.k2 * bse * T.e_mlclose * T.e_d3[tt][sp1] * T.e_d5[tt][sq_];
        real bm1 = 0;
        const int m1max = imin(L - q, W - d);
        {
          const real *pa = scrBif + (d + 5) * TC + t;
          const real *pb = c.arr[A_MULTI2] + 5 * c.NC + g + d;
          const long long nc = c.NC;
          int m = 5;
          for (; m + 7 <= m1max; m += 8) {  // 16 loads in flight; additions in the plain order
            real av[8], bv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              av[k] = pa[k * TC];
              bv[k] = pb[k * nc];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) bm1 += av[k] * bv[k];
            pa += 8 * TC;
            pb += 8 * nc;
          }
          for (; m <= m1max; ++m) {
            bm1 += pa[0] * pb[0];
            pa += TC;
            pb += nc;
          }
        }
        bm1 *= T.inv_cA;
        real ks = 0;
        const int m2max = imin(p, W - d);
        {
          const real *pa = scrBif + (d + 5) * TC + t - 5;
          const real *pb = c.arr[A_MULTI1] + 5 * c.NC + g - 5;
          const long long nc1 = c.NC - 1;
          int m = 5;
          for (; m + 7 <= m2max; m += 8) {
            real av[8], bv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              av[k] = pa[k * (TC - 1)];
              bv[k] = pb[k * nc1];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) ks += av[k] * bv[k];
            pa += 8 * (TC - 1);
            pb += 8 * nc1;
          }
          for (; m <= m2max; ++m) {
            ks += pa[0] * pb[0];
            pa += TC - 1;
            pb += nc1;
          }
        }
        bmulti2[r] = bm1 + sm.m2[((d + 1) & 1) * TC + t] * T.e_mlbase + ks * T.inv_cA;
        bmbif[r] = bm1 + bmulti[r];
      }
      const int t2 = T.bp[sp1][sq_];
      t2v[r] = t2;
      if (t2) {
        any = true;
        const int t2r = T.rt[t2];
        dang[r] = K::e_dangle(T, t2, p > 0, sp, q < L, sq1);
        base[r] = (real)exp(c.lao[g] + c.lbo[g + d] - c.lao[cs[r].zcol]) * dang[r] * T.sB[d];
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
          const int to = T.bp[s[-1]][s[d + 2]];
          l += b4[-2] * c.e_int11[idx11(to, t2r, sp, sq1)];
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
          l += T.tau[t2r] * bs;
        }
        if (smax >= 3) {
          const int ta = T.bp[s[-1]][s[d + 3]];
          l += b5[-2] * c.e_int21[idx21(ta, t2r, sp, sq1, s[d + 2])];
          const int tb = T.bp[s[-2]][s[d + 2]];
          l += b5[-3] * c.e_int21[idx21(t2r, tb, sq1, s[-1], sp)];
        }
        if (smax >= 4) {
          const int tc = T.bp[s[-2]][s[d + 3]];
          l += b6[-3] * c.e_int22[idx22(tc, t2r, s[-1], sp, sq1, s[d + 2])];
        }
        ls[r] = l;
      }
    }
    // generic loops: joint stencil over B_STEMO.  Target r (column c0 + r) takes source u1 from local
    // column c0 + r - 1 - u1; the elements are read as aligned vectors going left from c0 + R - 1.
    if (any && smax >= 4) {
      int slot = wrap_out(slot_d + 6);
#pragma unroll
      for (int sum = 4; sum <= kMaxLoop; ++sum) {
        if (sum <= smax) {
          const real *row = sm.stemO + slot * TC + c0;
          real rs[R];
#pragma unroll
          for (int r = 0; r < R; ++r) rs[r] = 0;
          // aligned blocks [c0 - yb - R, c0 - yb - 1], yb = 0, R, ...; for ascending u1 per target the
          // blocks are visited right to left and their elements right to left
#pragma unroll
          for (int yb = -R; yb <= sum - 1; yb += R) {
            real v[R];
            if (c0 - yb - R >= 0) {
              load_vec<R>(row - yb - R, v);
            } else {  // left of the tile: only columns that are not live could ask for it
#pragma unroll
              for (int k = 0; k < R; ++k) v[k] = 0;
            }
#pragma unroll
            for (int k = R - 1; k >= 0; --k) {
              const int e = -yb - R + k;  // column offset from c0
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const int u1 = r - 1 - e;
                if (u1 >= 1 && u1 <= sum - 1 && !(sum == 4 && u1 == 2))
                  rs[r] += gsel(K::gidx(u1, sum), g0, g1, g2, g3, g4, g5, g6) * v[k];
              }
            }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) gs[r] += cf[sum] * rs[r];
          slot = wrap_out(slot + 1);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = c0 + r;
      const long long g = ge.g0 - ge.H + t;
      const int L = cs[r].L, p = cs[r].i, q = p + d;
      if (t2v[r]) {
        const uint8_t *s = c.S + g;
        const int t2 = t2v[r], t2r = T.rt[t2];
        real l = ls[r];
        if (smax >= 4) l += T.e_mmI[t2r][s[d + 1]][s[0]] * gs[r];
        bstem[r] = base[r] + T.k2 * l + bmulti2[r] * T.e_mlintern * dang[r];
        bstemO[r] = bstem[r] * T.e_mmI[t2][s[2]][s[d - 1]];
        bstemB[r] = bstem[r] * T.tau[t2];
      }
      sm.stemO[slot_d * TC + t] = bstemO[r];
      sm.stemB[slot_d * TC + t] = bstemB[r];
      sm.stem[(d & (kRingStem - 1)) * TC + t] = bstem[r];
      sm.mu[(d & 1) * TC + t] = bmulti[r];
      sm.m2[(d & 1) * TC + t] = bmulti2[r];
      scrBif[d * TC + t] = bmbif[r];
      if (t >= ge.H && p >= 0 && q <= L) {
        c.at(B_STEM, d, g) = bstem[r];
        c.at(B_STEMO, d, g) = bstemO[r];
        c.at(B_STEMB, d, g) = bstemB[r];
        c.at(B_MULTI, d, g) = bmulti[r];
        c.at(B_MULTI2, d, g) = bmulti2[r];
        if (!(K::in_safe_range(bstem[r]) && K::in_safe_range(bmulti[r]) && K::in_safe_range(bmulti2[r]) &&
              K::in_safe_range(bmbif[r])))
          c.flags[cs[r].sq] = 1;
      }
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
  static PRIB_HD real load_left(const Ctx &c, const Geo &ge, int r, int x, int arr = A_STEMI) {
    const long long col = ge.g0 + x;
    return col < c.NC ? c.ld(arr, r, col) : (real)0;
  }
  // element of the end-indexed tile: span r, END column g0 - 31 + x        (right kernel)
  static PRIB_HD real load_right(const Ctx &c, const Geo &ge, int r, int x, int arr = A_STEMI) {
    const long long col = ge.g0 - 31 + x - r;
    return (col >= 0 && col < c.NC) ? c.ld(arr, r, col) : (real)0;
  }

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
    const real *cv = K::conv_tab(T), *bu = K::bulge_tab(T);
    const int L = ci.L, i = ci.i, W = c.W, delta = c.delta, TXb = ge.TXb, cols = COLS > 0 ? COLS : ge.cols;
    const uint8_t *s = c.S + g;
    real (&ml)[kMaxLoop + 1] = st.w;
#pragma unroll
    for (int u = 0; u <= kMaxLoop; ++u) ml[u] = 0;
    int cnt = 0;
    if (i >= 1) {
      const int dpmax = imin(W - 1, L - 1 - i);
      // pass A (all lanes at the same span): list the closing spans, add bulges (u2 = 0) and, for
      // delta == 2, the 2x1 / 2x2 special loops
      for (int dp0 = delta + 5; dp0 <= dpmax; dp0 += 8) {
        real bv[8];  // 8 independent loads in flight (the list update is a serial chain on the loaded values)
#pragma unroll
        for (int k = 0; k < 8; ++k) bv[k] = (dp0 + k <= dpmax) ? c.ld(B_STEM, dp0 + k + 2, g - 1) : (real)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int dp = dp0 + k;
          const real bse = bv[k];
          if (bse == 0) continue;
          list[cnt * TXb + t] = (uint8_t)dp;
          ++cnt;
          if (delta == 2) {
            const int te = T.bp[s[0]][s[dp + 1]];
            const real bseO = c.ld(B_STEMO, dp + 2, g - 1), bseB = c.ld(B_STEMB, dp + 2, g - 1);
            if (dp - 5 >= 3) ml[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 1);
            if (dp - 5 >= 4) ml[2] += K::loop_weight(c, T, s, g, dp, te, bse, bseO, bseB, 2, 2);
          }
        }
      }
      // pass B (each lane at its own span): generic interior loops out of the shared-memory tile, walked by
      // loop size: all terms of one size share a tile row, every offset and coefficient is compile-time
      real bseO_next = cnt > 0 ? c.ld(B_STEMO, list[t] + 2, g - 1) : (real)0;
      for (int k = 0; k < cnt; ++k) {
        const int dp = list[k * TXb + t];
        const real bseO = bseO_next;  // fetched one iteration ahead: the gather latency hides behind the stencil
        if (k + 1 < cnt) bseO_next = c.ld(B_STEMO, list[(k + 1) * TXb + t] + 2, g - 1);
        const int smax = imin(kMaxLoop, dp - 5);
        const real *base = tile + (dp - 5) * cols + t;
        real a[kMaxLoop];
#pragma unroll
        for (int u1 = 0; u1 < kMaxLoop; ++u1) a[u1] = 0;
#pragma unroll
        for (int sum = ULO + 1; sum <= kMaxLoop; ++sum) {
          if (sum <= smax) {
            const real *row = base - sum * cols;  // span dp - sum
#pragma unroll
            for (int u1 = ULO; u1 < sum; ++u1) a[u1] += cv[u1 * 32 + sum - u1] * row[u1];
          }
        }
#pragma unroll
        for (int u1 = ULO; u1 < kMaxLoop; ++u1) ml[u1] += bseO * a[u1];
      }
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
    for (int k = 0; k < st.cnt; ++k) {
      const int dp = list[k * TXb + t];
      const real bseB = bseB_next;
      if (k + 1 < st.cnt) bseB_next = c.ld(B_STEMB, list[(k + 1) * TXb + t] + 2, g - 1);
      const int umax = imin(kMaxLoop, dp - 5);
      const real *base = tile + (dp - 5) * cols + t;
#pragma unroll
      for (int u1 = ULO; u1 <= kMaxLoop; ++u1)
        if (u1 >= delta && u1 <= umax) ml[u1] += bseB * bu[u1] * base[u1 - u1 * cols];  // cell (i+u1, j'), span dp-u1
    }
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
    const real *cv = K::conv_tab(T), *bu = K::bulge_tab(T);
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
      real bseO_next = cnt > 0 ? c.ld(B_STEMO, list[t] + 2, g2 - list[t] - 1) : (real)0;
      for (int k = 0; k < cnt; ++k) {
        const int dp = list[k * TXb + t];
        const real bseO = bseO_next;
        if (k + 1 < cnt) {
          const int dn = list[(k + 1) * TXb + t];
          bseO_next = c.ld(B_STEMO, dn + 2, g2 - dn - 1);
        }
        const int smax = imin(kMaxLoop, dp - 5);
        const real *base = tile + (dp - 5) * cols + t + 31;
        real a[kMaxLoop];
#pragma unroll
        for (int u2 = 0; u2 < kMaxLoop; ++u2) a[u2] = 0;
#pragma unroll
        for (int sum = ULO + 1; sum <= kMaxLoop; ++sum) {
          if (sum <= smax) {
            const real *row = base - sum * cols;  // span dp - sum, end-indexed: the inner cell ends at j' - u2
#pragma unroll
            for (int u2 = ULO; u2 < sum; ++u2) a[u2] += cv[u2 * 32 + sum - u2] * row[-u2];
          }
        }
#pragma unroll
        for (int u2 = ULO; u2 < kMaxLoop; ++u2) mr[u2] += bseO * a[u2];
      }
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
    for (int k = 0; k < st.cnt; ++k) {
      const int dp = list[k * TXb + t];
      const real bseB = bseB_next;
      if (k + 1 < st.cnt) {
        const int dn = list[(k + 1) * TXb + t];
        bseB_next = c.ld(B_STEMB, dn + 2, g2 - dn - 1);
      }
      const int umax = imin(kMaxLoop, dp - 5);
      const real *base = tile + (dp - 5) * cols + t + 31;
#pragma unroll
      for (int u2 = ULO; u2 <= kMaxLoop; ++u2)
        if (u2 >= delta && u2 <= umax) mr[u2] += bseB * bu[u2] * base[-u2 - u2 * cols];  // cell (i, j'-u2), span dp-u2
    }
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
