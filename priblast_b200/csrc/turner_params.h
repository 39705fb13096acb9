// Turner-1999-style nearest-neighbour parameters as one POD blob.
//
// The VALUES are the ones the reference compiles in (reference: energy_par.hpp:6-174 and
// intloops.hpp:6,309,1788).  They are data, not code: oracle/dump_params.cpp prints them from the
// reference headers into priblast_b200/data/turner99.bin, which is embedded in the library with
// .incbin (turner_blob.cpp).  Units: 0.01 kcal/mol as int32; `inf` (1000000) marks forbidden entries
// and is kept verbatim (the reference scales it like any other number).
#pragma once
#include <stdint.h>

#define PRIB_TURNER_MAGIC 0x39395254u /* "TR99" */
#define PRIB_TURNER_VERSION 1

#ifdef __cplusplus
extern "C" {
#endif

typedef struct prib_turner_params {
  uint32_t magic;
  uint32_t version;
  int32_t inf;            /* energy_par.hpp:8  */
  int32_t turn;           /* energy_par.hpp:9  */
  int32_t maxloop;        /* energy_par.hpp:10 */
  int32_t temperature_c;  /* energy_par.hpp:12 */
  int32_t terminal_au;    /* energy_par.hpp:92 */
  int32_t ml_closing;     /* energy_par.hpp:144 */
  int32_t ml_intern;      /* energy_par.hpp:145 */
  int32_t ml_base;        /* energy_par.hpp:146 */
  int32_t max_ninio;      /* energy_par.hpp:173 */
  int32_t f_ninio;        /* energy_par.hpp:174 */
  double gasconst;        /* energy_par.hpp:6  */
  double k0;              /* energy_par.hpp:7  */
  double lxc37;           /* energy_par.hpp:14 */
  int32_t bp_pair[5][5];  /* energy_par.hpp:17 */
  int32_t rtype[7];       /* energy_par.hpp:26 */
  int32_t pad0;
  int32_t hairpin[31];
  int32_t bulge[31];
  int32_t internal_loop[31];
  int32_t pad1;
  int32_t mismatch_h[7][5][5];
  int32_t mismatch_i[7][5][5];
  int32_t stack[7][7];
  int32_t pad2;
  int32_t dangle5[8][5];
  int32_t dangle3[8][5];
  int32_t int11[8][8][5][5];
  int32_t int21[8][8][5][5][5];
  int32_t int22[8][8][5][5][5][5];
} prib_turner_params;

/* Returns the embedded blob (validated magic/version), or NULL. */
const prib_turner_params *prib_turner_embedded(void);

#ifdef __cplusplus
}
#endif
