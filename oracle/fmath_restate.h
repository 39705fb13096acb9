/* TEST INFRASTRUCTURE — restatement of the two fmath primitives pRIblast uses.
 *
 * fmath (Mitsunari Shigeo, BSD-3) is vendored in the reference as fmath.hpp; pRIblast calls only
 *   fmath::expd(double)  reference: fmath.hpp:439-479   (2^11-entry table, cubic correction)
 *   fmath::log(float)    reference: fmath.hpp:738-752   (2^11-entry table, linear interpolation)
 * Both tables are built at static-init from the host libm (fmath.hpp:161-177, 193-207), so this
 * restatement builds them the same way from the same libm calls and is bit-identical on the same host.
 */
#ifndef FMATH_RESTATE_H
#define FMATH_RESTATE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void fmr_init(void);                 /* idempotent, thread-safe after first call returns */
double fmr_expd(double x);           /* == fmath::expd */
float fmr_logf(float x);             /* == fmath::log(float) */

/* Raw log table for upload to the device: 2048 x {app, rev} floats, plus c_log2. */
const float *fmr_log_table(void);    /* 4096 floats: app0, rev0, app1, rev1, ... */
float fmr_log_c_log2(void);

#ifdef __cplusplus
}
#endif
#endif
