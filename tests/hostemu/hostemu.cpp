// TEST INFRASTRUCTURE — CPU emulation of the CUDA kernels' per-thread bodies.
//
// The device code in priblast_b200/csrc/acc_kernels.cu is a set of "one thread = one column" kernels
// whose bodies are the PRIB_HD functions of acc_core.h.  This harness runs the very same functions in
// plain loops (same launch order as the device driver) so the restructured mathematics can be checked
// against the oracle in a container without a GPU.  It is NOT linked into libpriblast_acc.so, is not
// reachable from the C ABI, and is not a fallback: the product fails loudly without CUDA.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../priblast_b200/csrc/acc_core.h"
#include "../../priblast_b200/csrc/acc_tables.h"
#include "../../priblast_b200/csrc/acc_tile.h"

using namespace prib;
static int g_poison = 0;
static int g_chain = 0;  // 1: the tile emulation uses the centre-line chain formulation of the deep steps (FP32 device engine)

namespace {

template <typename real>
struct EmuT {
  HostTablesT<real> tab;
  BatchLayout lay;
  std::vector<std::vector<real>> arr;
  std::vector<double> lao, lbo;
  std::vector<int32_t> flags;
  typename Core<real>::Ctx c;
};

template <typename real>
bool setup(EmuT<real> &e, int n, const char *const *seqs, const int32_t *lens, int W, int delta, float *out,
           const int64_t *acc_off, const int64_t *cond_off, const ScaleSpec &spec = ScaleSpec()) {
  std::string err;
  if (!build_tables(W, delta, spec, e.tab, err)) return false;
  build_layout(n, seqs, lens, e.lay);
  typename Core<real>::Ctx &c = e.c;
  std::memset(&c, 0, sizeof(c));
  c.NC = e.lay.NC;
  c.W = W;
  c.delta = delta;
  c.rows = W + 4;
  c.nseq = n;
  c.S = e.lay.S.data();
  c.col_seq = e.lay.col_seq.data();
  c.seq_len = e.lay.seq_len.data();
  c.seq_off = e.lay.seq_off.data();
  c.T = &e.tab.small;
  c.e_int11 = e.tab.e_int11.data();
  c.e_int21 = e.tab.e_int21.data();
  c.e_int22 = e.tab.e_int22.data();
  c.log_tbl = e.tab.log_tbl.data();
  e.arr.resize(kNumArr);
  for (int a = 0; a < kNumArr; a++) {
    int rows = (a == X_ML || a == X_MR || a == X_MLS || a == X_MRS) ? 32 : c.rows;
    e.arr[a].assign((size_t)rows * (size_t)c.NC, (real)0);
    c.arr[a] = e.arr[a].data();
  }
  e.lao.assign((size_t)c.NC, 0.0);
  e.lbo.assign((size_t)c.NC, 0.0);
  c.lao = e.lao.data();
  c.lbo = e.lbo.data();
  c.acc_off = (const long long *)acc_off;
  c.cond_off = (const long long *)cond_off;
  c.out = out;
  e.flags.assign((size_t)n + 1, 0);
  c.flags = e.flags.data();
  return true;
}

template <typename real>
void run_dp(EmuT<real> &e) {
  typedef Core<real> K;
  const typename K::Ctx &c = e.c;
  double ring[256];
  for (int d = kTurn; d <= c.W + 1; d++)
    for (long long g = 0; g < c.NC; g++) K::inside_cell(c, g, d);
  for (int k = 0; k < c.nseq; k++) {
    K::scan_alpha_outer(c, k, ring);
    K::scan_beta_outer(c, k, ring);
  }
  for (int d = c.W + 1; d >= kTurn; d--)
    for (long long g = 0; g < c.NC; g++) K::outside_cell(c, g, d);
}

// Emulation of the tile-persistent, time-tiled kernels (acc_tile.h): one "CTA" per tile, TC "threads"; a
// loop over t = everything the threads do between two barriers.
template <typename real>
void run_dp_tiled(EmuT<real> &e, int TC) {
  typedef Tile<real> TL;
  typedef Core<real> K;
  const typename K::Ctx &c = e.c;
  const int W = c.W, H = W + 1, TX = TC - H;
  const long long ntiles = (c.NC + TX - 1) / TX;
  std::vector<real> smem((size_t)kTilePad + (size_t)kTileRows * TC), scr((size_t)2 * (W + 4) * TC);
  std::vector<uint8_t> sS((size_t)TC + 16);
  std::vector<typename TL::ColState> cs(TC);
  real *base = smem.data() + kTilePad;
  const int dfirst = TL::first_group(W);
  struct InDeep { real gs[kTT], mb[kTT], bs[kTT]; };
  std::vector<InDeep> din(TC);
  std::vector<typename TL::OutDeep> dout(TC);
  std::vector<typename TL::Chain> chains(TC);
  real *scrM1 = scr.data(), *scrM2 = scr.data() + (size_t)(W + 4) * TC;
  for (long long tile = 0; tile < ntiles; tile++) {
    typename TL::Geo ge{tile * TX, TC, TX, H};
    std::fill(smem.begin(), smem.end(), (real)0);
    if (g_poison) std::fill(scr.begin(), scr.end(), std::numeric_limits<real>::quiet_NaN());
    for (int k = 0; k < TC + 8; k++) sS[k] = (ge.g0 + k < c.NC) ? c.S[ge.g0 + k] : 0;
    for (int t = 0; t < TC; t++) TL::col_state(c, ge.g0 + t, cs[t]);
    typename TL::InSmem sm = TL::carve_in(base, TC, sS.data());
    for (auto &ch : chains) TL::clear(ch);
    int grp = 0;
    for (int d0 = dfirst; d0 <= W + 1; d0 += kTT, ++grp) {
      real *xch = sm.xch + (size_t)(grp & 1) * kTT * TC;
      for (int t = 0; t < TC; t++) {
        if (!g_chain) TL::template inside_deep<0>(*c.T, ge, sm, scrM1, scrM2, t, d0, din[t].gs, din[t].mb, din[t].bs);
        else if (dfirst & 1) TL::template inside_deep_chain<1, 0>(*c.T, ge, sm, scrM1, scrM2, t, d0, chains[t], xch, din[t].mb, din[t].bs);
        else TL::template inside_deep_chain<0, 0>(*c.T, ge, sm, scrM1, scrM2, t, d0, chains[t], xch, din[t].mb, din[t].bs);
      }
      for (int k = 0; k < kTT; k++) {
        if (d0 + k < kTurn) continue;
        for (int t = 0; t < TC; t++)
          TL::template inside_shallow<0>(c, *c.T, ge, sm, scrM1, scrM2, t, cs[t], d0 + k,
                                         g_chain ? xch[(size_t)k * TC + t] : din[t].gs[k], din[t].mb[k], din[t].bs[k]);
      }
    }
  }
  double ring[256];
  for (int k = 0; k < c.nseq; k++) {
    K::scan_alpha_outer(c, k, ring);
    K::scan_beta_outer(c, k, ring);
  }
  for (long long tile = 0; tile < ntiles; tile++) {
    typename TL::Geo ge{tile * TX, TC, TX, H};
    std::fill(smem.begin(), smem.end(), (real)0);
    if (g_poison) std::fill(scr.begin(), scr.end(), std::numeric_limits<real>::quiet_NaN());
    for (int t = 0; t < TC; t++) TL::col_state(c, ge.g0 - H + t, cs[t]);
    std::vector<uint8_t> sSo((size_t)TC + c.W + kOutBaseTail);
    for (int k = 0; k < TC + c.W + kOutBaseTail; k++) {
      const long long col = ge.g0 - ge.H - kOutBaseLead + k;
      sSo[k] = (col >= 0 && col < c.NC) ? c.S[col] : 0;
    }
    typename TL::OutSmem sm = TL::carve_out(base, TC, sSo.data());
    for (auto &ch : chains) TL::clear(ch);
    int grp = 0;
    for (int d0 = W + 1; d0 >= dfirst + kTT - 1; d0 -= kTT, ++grp) {
      real *xch = sm.xch + (size_t)(grp & 1) * kTT * TC;
      const int slot0 = ((d0 % kRingOut) + kRingOut) % kRingOut;
      for (int t = 0; t < TC; t++) {
        if (!g_chain) TL::template outside_deep<0>(c, *c.T, ge, sm, scr.data(), t, cs[t], d0, slot0, dout[t]);
        else TL::template outside_deep_chain<0>(c, *c.T, ge, sm, scr.data(), t, cs[t], d0, slot0, chains[t], xch, dout[t]);
      }
      for (int k = 0; k < kTT; k++) {
        const int d = d0 - k;
        if (d < kTurn) continue;
        for (int t = 0; t < TC; t++)
          TL::template outside_shallow<0>(c, *c.T, ge, sm, scr.data(), t, cs[t], d, d % kRingOut,
                                          g_chain ? xch[(size_t)k * TC + t] : dout[t].gs[k], dout[t].bs[k], dout[t].bm1[k],
                                          dout[t].ks[k]);
      }
    }
  }
}

template <typename real>
void run_biloop_tiled(EmuT<real> &e, int TXb) {
  typedef BiTile<real> BT;
  const typename Core<real>::Ctx &c = e.c;
  typename BT::Geo ge;
  ge.TXb = TXb;
  ge.rows = c.W - 5 > 0 ? c.W - 5 : 0;
  std::vector<real> tile((size_t)ge.rows * BT::tile_cols(false, TXb) + 1);
  std::vector<uint8_t> list((size_t)(c.W + 1) * TXb);
  std::vector<typename BT::Strand> st(TXb);
  for (int side = 0; side < 2; side++)
    for (long long g0 = 0; g0 < c.NC; g0 += TXb) {
      ge.g0 = g0;
      ge.cols = BT::tile_cols(side == 0, TXb);
      auto fill = [&](int arr) {  // what the bulk copy + masking of the device kernel leave in shared memory
        std::fill(tile.begin(), tile.end(), std::numeric_limits<real>::quiet_NaN());  // slack elements are never read
        for (int x = 0; x < TXb + 32; x++) {
          const int lim = BT::tile_col_limit(c, ge, side == 0, x);
          for (int r = 5; r < 5 + ge.rows; r++)
            tile[(size_t)(r - 5) * ge.cols + x + BT::row_off(ge, side == 0, r)] = BT::tile_elem(c, ge, side == 0, arr, r, x, lim);
        }
      };
      fill(A_STEMI);
      for (int t = 0; t < TXb; t++) {
        if (side == 0) {
          if (c.delta >= 5) BT::template left<0, 5>(c, ge, tile.data(), list.data(), t, st[t]);
          else BT::template left<0, 2>(c, ge, tile.data(), list.data(), t, st[t]);
        } else {
          if (c.delta >= 5) BT::template right<0, 5>(c, ge, tile.data(), list.data(), t, st[t]);
          else BT::template right<0, 2>(c, ge, tile.data(), list.data(), t, st[t]);
        }
      }
      fill(A_STEMB);
      for (int t = 0; t < TXb; t++) {
        if (side == 0) {
          if (c.delta >= 5) BT::template left_bulge<0, 5>(c, ge, tile.data(), list.data(), t, st[t]);
          else BT::template left_bulge<0, 2>(c, ge, tile.data(), list.data(), t, st[t]);
        } else {
          if (c.delta >= 5) BT::template right_bulge<0, 5>(c, ge, tile.data(), list.data(), t, st[t]);
          else BT::template right_bulge<0, 2>(c, ge, tile.data(), list.data(), t, st[t]);
        }
      }
    }
}

template <typename real>
void run_acc(EmuT<real> &e, int TXb = 0) {
  typedef Core<real> K;
  const typename K::Ctx &c = e.c;
  if (TXb > 0) run_biloop_tiled(e, TXb);
  for (long long g = 0; g < c.NC; g++) {
    if (TXb == 0) {
      K::biloop_left(c, g);
      K::biloop_right(c, g);
      K::hairpin_suffix(c, g);  // the tiled left strand-weight body produces X_SUFH itself
    }
  }
  for (long long g = 0; g < c.NC; g++) K::finalize_position(c, g);
}

}  // namespace

extern "C" {

int hostemu_run_batch(int n, const char *const *seqs, const int32_t *lens, int W, int delta, float *out,
                      const int64_t *acc_off, const int64_t *cond_off, int /*nthreads*/) {
  EmuT<double> e;
  for (int k = 0; k < n; k++) {
    std::memset(out + acc_off[k], 0, sizeof(float) * (size_t)lens[k]);
    std::memset(out + cond_off[k], 0, sizeof(float) * (size_t)lens[k]);
  }
  if (!setup(e, n, seqs, lens, W, delta, out, acc_off, cond_off)) return -1;
  run_dp(e);
  run_acc(e);
  return 1;
}

}  // extern "C"

namespace {
template <typename real>
int run_tiled(int n, const char *const *seqs, const int32_t *lens, int W, int delta, float *out,
              const int64_t *acc_off, const int64_t *cond_off, int TC, const ScaleSpec &spec, int32_t *flags_out,
              int /*unused*/ = 1) {
  EmuT<real> e;
  for (int k = 0; k < n; k++) {
    std::memset(out + acc_off[k], 0, sizeof(float) * (size_t)lens[k]);
    std::memset(out + cond_off[k], 0, sizeof(float) * (size_t)lens[k]);
  }
  if (!setup(e, n, seqs, lens, W, delta, out, acc_off, cond_off, spec)) return -1;
  if (TC <= W + 2 || TC % 4 != 0) return -2;
  if (g_poison) {
    for (auto &a : e.arr) std::fill(a.begin(), a.end(), std::numeric_limits<real>::quiet_NaN());
    std::fill(e.lao.begin(), e.lao.end(), std::numeric_limits<double>::quiet_NaN());
    std::fill(e.lbo.begin(), e.lbo.end(), std::numeric_limits<double>::quiet_NaN());
  }
  run_dp_tiled<real>(e, TC);
  run_acc(e, TC >= 256 ? 128 : 64);
  if (flags_out) std::memcpy(flags_out, e.flags.data(), sizeof(int32_t) * (size_t)n);
  return 1;
}
}  // namespace

extern "C" {

// poison = 1: fill every DP array with NaN before the run (emulates a device that does not zero its state):
// any read of a never-written cell then shows up in the output
void hostemu_set_poison(int on) { g_poison = on; }
// chain = 1: deep steps in the centre-line chain formulation (what the FP32 device engine runs)
void hostemu_set_chain(int on) { g_chain = on; }
// tf32 = 1: the FP32 engine's generic-loop stencil products in the 3-product TF32 split of a tensor-core formulation
void hostemu_set_tf32_split(int on) { Tile<float>::emu_tf32() = on; }

// FP32 band arithmetic with span scaling; flags_out[k] != 0 marks sequences whose stored values left
// the safe range (the product re-runs those in FP64).
int hostemu_run_batch_tiled_f32(int n, const char *const *seqs, const int32_t *lens, int W, int delta, float *out,
                                const int64_t *acc_off, const int64_t *cond_off, int TC, double klog2,
                                double alog2, double blog2, int32_t *flags_out) {
  ScaleSpec spec;
  spec.klog2 = klog2;
  spec.alog2 = alog2;
  spec.blog2 = blog2;
  return run_tiled<float>(n, seqs, lens, W, delta, out, acc_off, cond_off, TC, spec, flags_out);
}

// Same batch through the tile-persistent formulation; TC = emulated CTA width.
int hostemu_run_batch_tiled(int n, const char *const *seqs, const int32_t *lens, int W, int delta, float *out,
                            const int64_t *acc_off, const int64_t *cond_off, int TC, double klog2, double alog2,
                            double blog2) {
  ScaleSpec spec;
  spec.klog2 = klog2;
  spec.alog2 = alog2;
  spec.blog2 = blog2;
  return run_tiled<double>(n, seqs, lens, W, delta, out, acc_off, cond_off, TC, spec, nullptr);
}

int hostemu_run(const char *seq, int L, int W, int delta, float *acc, float *cond) {
  std::vector<float> out(2 * (size_t)L + 2, 0.f);
  int64_t ao = 0, co = L;
  int32_t len = L;
  const char *sp = seq;
  int rc = hostemu_run_batch(1, &sp, &len, W, delta, out.data(), &ao, &co, 1);
  std::memcpy(acc, out.data(), sizeof(float) * (size_t)L);
  std::memcpy(cond, out.data() + L, sizeof(float) * (size_t)L);
  return rc > 0 ? 0 : rc;
}

// Debug: band state of one sequence, arrays as [kNumArr][W+4][L+1] (left index fastest), plus logs.
int hostemu_dump(const char *seq, int L, int W, int delta, double *band, double *lao, double *lbo) {
  EmuT<double> e;
  std::vector<float> out(2 * (size_t)L + 2, 0.f);
  int64_t ao = 0, co = L;
  int32_t len = L;
  const char *sp = seq;
  if (!setup(e, 1, &sp, &len, W, delta, out.data(), &ao, &co)) return -1;
  run_dp(e);
  const long long off = e.lay.seq_off[0];
  for (int a = 0; a < X_ML; a++)
    for (int d = 0; d < W + 4; d++)
      for (int i = 0; i <= L; i++)
        band[((size_t)a * (W + 4) + d) * (L + 1) + i] = e.arr[a][(size_t)d * e.c.NC + off + i];
  for (int i = 0; i <= L; i++) {
    lao[i] = e.lao[(size_t)(off + i)];
    lbo[i] = e.lbo[(size_t)(off + i)];
  }
  return 0;
}

int hostemu_num_arrays() { return X_ML; }

}  // extern "C"
