/* TEST INFRASTRUCTURE: exports the reference's own suffix-array builder (sais.cpp:656, called at
 * db_construction.cpp:334) from oracle/_ref/libsais_ref.so so tests/test_gpu_sa.py can compare with it.
 * Compiled as C++ together with the reference's sais.cpp (oracle/Makefile target refsais). */
#include "sais.hpp"
extern "C" int ref_sais(const unsigned char *T, int *SA, int n) { return sais(T, SA, n); }
