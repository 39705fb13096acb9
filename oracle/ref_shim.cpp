// TEST INFRASTRUCTURE — not part of the product.
//
// Thin C-ABI harness around the UNMODIFIED reference `Raccess` class.  It is compiled by
// oracle/Makefile together with /root/reference/src/raccess.cpp *where that file lies* (no reference
// source is copied into this repository); the result goes to oracle/_ref/ (git-ignored).
//
// The only symbol the reference translation unit needs from outside is `MyAccFile`
// (reference: utils.cpp:62-72, which pulls in <mpi.h>); a single-rank stand-in is defined here.
//
// `#define private public` is used so the harness can dump the reference's DP state (band arrays,
// outer arrays) for stage-by-stage debugging of the restatement and of the CUDA path.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#include "raccess.hpp"
#undef private
#include "utils.hpp"

static thread_local std::string g_tmp_dir = "/tmp";

std::string MyAccFile(const std::string &path, int idx) {
  std::stringstream s;
  if (path != "") s << path << "/";
  s << "priblast_tmp_acc0_" << idx << ".acc";
  return s.str();
}

extern "C" {

// Reference `Raccess::Run(seq, acc, cond)` (raccess.cpp:42-50).  acc/cond receive L floats each.
int ref_raccess_run(const char *seq, int L, int W, int delta, float *acc, float *cond) {
  Raccess r(W, delta);
  std::string s(seq, (size_t)L);
  std::vector<float> a, c;
  r.Run(s, a, c);
  if ((int)a.size() != L || (int)c.size() != L) return -1;
  std::memcpy(acc, a.data(), sizeof(float) * (size_t)L);
  std::memcpy(cond, c.data(), sizeof(float) * (size_t)L);
  return 0;
}

// Reference `Raccess::Run(seq, idx)` (raccess.cpp:34-40): writes <dir>/priblast_tmp_acc0_<idx>.acc
int ref_raccess_run_file(const char *seq, int L, int W, int delta, const char *dir, int idx) {
  Raccess r("db", W, delta, dir);
  std::string s(seq, (size_t)L);
  r.Run(s, idx);
  return 0;
}

// Batch over sequences with OpenMP the way DbConstruction::CalculateAccessibility does
// (db_construction.cpp:182-223): thread-private Raccess, dynamic counter, caller passes the
// longest-first order.  out receives, per sequence, acc at acc_off[k] (L floats) and cond at
// cond_off[k] (L floats).  Returns the number of threads used.
int ref_raccess_batch(int n, const char *const *seqs, const int32_t *lens, int W, int delta,
                      float *out, const int64_t *acc_off, const int64_t *cond_off, int nthreads) {
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int next = 0;
#pragma omp parallel
  {
#ifdef _OPENMP
#pragma omp single
    used = omp_get_num_threads();
#endif
    Raccess r(W, delta);
    std::vector<float> a, c;
    for (;;) {
      int k;
#pragma omp atomic capture
      k = next++;
      if (k >= n) break;
      std::string s(seqs[k], (size_t)lens[k]);
      r.Run(s, a, c);
      std::memcpy(out + acc_off[k], a.data(), sizeof(float) * (size_t)lens[k]);
      std::memcpy(out + cond_off[k], c.data(), sizeof(float) * (size_t)lens[k]);
    }
  }
  return used;
}

// Debug dump of the reference DP state after inside+outside.  Every band array is written as
// (L+1)*(W+2) doubles, row = left index, column = span (the reference's own layout,
// raccess.cpp:70-96).  which: 0 stem 1 stemend 2 multi 3 multibif 4 multi1 5 multi2 (Alpha),
// 6..11 same for Beta; outer arrays are (L+1) doubles.
int ref_raccess_dump(const char *seq, int L, int W, int delta, double *band12, double *alpha_outer,
                     double *beta_outer) {
  Raccess r(W, delta);
  std::string s(seq, (size_t)L);
  r.Initiallize(s);
  r.CalcInsideVariable();
  r.CalcOutsideVariable();
  const std::vector<std::vector<double>> *arrs[12] = {
      &r._Alpha_stem, &r._Alpha_stemend, &r._Alpha_multi, &r._Alpha_multibif, &r._Alpha_multi1,
      &r._Alpha_multi2, &r._Beta_stem,  &r._Beta_stemend, &r._Beta_multi,  &r._Beta_multibif,
      &r._Beta_multi1,  &r._Beta_multi2};
  size_t plane = (size_t)(L + 1) * (size_t)(W + 2);
  for (int a = 0; a < 12; a++)
    for (int i = 0; i <= L; i++)
      std::memcpy(band12 + a * plane + (size_t)i * (W + 2), (*arrs[a])[i].data(),
                  sizeof(double) * (size_t)(W + 2));
  std::memcpy(alpha_outer, r._Alpha_outer.data(), sizeof(double) * (size_t)(L + 1));
  std::memcpy(beta_outer, r._Beta_outer.data(), sizeof(double) * (size_t)(L + 1));
  r.Clear();
  return 0;
}

}  // extern "C"
