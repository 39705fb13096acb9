// TEST/BUILD INFRASTRUCTURE — prints the reference's compiled-in energy parameters into the POD blob
// priblast_b200/data/turner99.bin.  Compiled with -I/root/reference/src so the numbers come from the
// reference headers where they lie (energy_par.hpp, intloops.hpp); only the resulting data file is
// committed.  Run by `make -C oracle params` in the build container (the GPU box never needs it).
#include <cstdio>
#include <cstring>

#include "energy_par.hpp"
#include "intloops.hpp"

#include "../priblast_b200/csrc/turner_params.h"

int main(int argc, char **argv) {
  if (argc != 2) {
    std::fprintf(stderr, "usage: %s out.bin\n", argv[0]);
    return 2;
  }
  static prib_turner_params p;
  std::memset(&p, 0, sizeof(p));
  p.magic = PRIB_TURNER_MAGIC;
  p.version = PRIB_TURNER_VERSION;
  p.inf = INF;
  p.turn = TURN;
  p.maxloop = MAXLOOP;
  p.temperature_c = temperature;
  p.terminal_au = TerminalAU;
  p.ml_closing = ML_closing37;
  p.ml_intern = ML_intern37;
  p.ml_base = ML_BASE37;
  p.max_ninio = MAX_NINIO;
  p.f_ninio = F_ninio37;
  p.gasconst = GASCONST;
  p.k0 = K0;
  p.lxc37 = lxc37;
  static_assert(sizeof(p.bp_pair) == sizeof(BP_pair), "bp");
  static_assert(sizeof(p.rtype) == sizeof(rtype), "rtype");
  static_assert(sizeof(p.hairpin) == sizeof(hairpin37), "hairpin");
  static_assert(sizeof(p.bulge) == sizeof(bulge37), "bulge");
  static_assert(sizeof(p.internal_loop) == sizeof(internal_loop37), "internal");
  static_assert(sizeof(p.mismatch_h) == sizeof(mismatchH37), "mmH");
  static_assert(sizeof(p.mismatch_i) == sizeof(mismatchI37), "mmI");
  static_assert(sizeof(p.stack) == sizeof(stack37), "stack");
  static_assert(sizeof(p.dangle5) == sizeof(dangle5_37), "d5");
  static_assert(sizeof(p.dangle3) == sizeof(dangle3_37), "d3");
  static_assert(sizeof(p.int11) == sizeof(int11_37), "int11");
  static_assert(sizeof(p.int21) == sizeof(int21_37), "int21");
  static_assert(sizeof(p.int22) == sizeof(int22_37), "int22");
  std::memcpy(p.bp_pair, BP_pair, sizeof(BP_pair));
  std::memcpy(p.rtype, rtype, sizeof(rtype));
  std::memcpy(p.hairpin, hairpin37, sizeof(hairpin37));
  std::memcpy(p.bulge, bulge37, sizeof(bulge37));
  std::memcpy(p.internal_loop, internal_loop37, sizeof(internal_loop37));
  std::memcpy(p.mismatch_h, mismatchH37, sizeof(mismatchH37));
  std::memcpy(p.mismatch_i, mismatchI37, sizeof(mismatchI37));
  std::memcpy(p.stack, stack37, sizeof(stack37));
  std::memcpy(p.dangle5, dangle5_37, sizeof(dangle5_37));
  std::memcpy(p.dangle3, dangle3_37, sizeof(dangle3_37));
  std::memcpy(p.int11, int11_37, sizeof(int11_37));
  std::memcpy(p.int21, int21_37, sizeof(int21_37));
  std::memcpy(p.int22, int22_37, sizeof(int22_37));
  FILE *f = std::fopen(argv[1], "wb");
  if (!f) return 1;
  std::fwrite(&p, sizeof(p), 1, f);
  std::fclose(f);
  std::printf("wrote %zu bytes\n", sizeof(p));
  return 0;
}
