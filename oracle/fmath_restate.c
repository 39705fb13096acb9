/* TEST INFRASTRUCTURE — see fmath_restate.h. */
#include "fmath_restate.h"

#include <math.h>
#include <string.h>

#define EXPD_SBIT 11                       /* fmath.hpp:81  EXPD_TABLE_SIZE */
#define EXPD_S (1u << EXPD_SBIT)
#define LOG_LEN 11                         /* fmath.hpp:82,182  LOG_TABLE_SIZE - 1 */
#define LOG_N (1u << LOG_LEN)

static uint64_t g_expd_tbl[EXPD_S];
static double g_expd_a, g_expd_ra;
static float g_log_tbl[2 * LOG_N];         /* app, rev interleaved */
static float g_c_log2;
static volatile int g_ready = 0;

void fmr_init(void) {
  if (g_ready) return;
#pragma omp critical(fmr_init_lock)
  {
    if (!g_ready) {
      /* ExpdVar ctor, fmath.hpp:161-177 */
      g_expd_a = (double)EXPD_S / log(2.0);
      g_expd_ra = 1 / g_expd_a;
      for (unsigned i = 0; i < EXPD_S; i++) {
        double v = pow(2.0, i * (1.0 / EXPD_S));
        uint64_t bits;
        memcpy(&bits, &v, 8);
        g_expd_tbl[i] = bits & ((1ULL << 52) - 1);
      }
      /* LogVar ctor, fmath.hpp:193-207 */
      g_c_log2 = logf(2.0f) / (1 << 23);
      const double e = 1 / (double)(1 << 24);
      const double h = 1 / (double)(1 << LOG_LEN);
      for (unsigned i = 0; i < LOG_N; i++) {
        double x = 1 + (double)i / LOG_N;
        double a = log(x);
        g_log_tbl[2 * i] = (float)a;
        if (i < LOG_N - 1) {
          double b = log(x + h - e);
          g_log_tbl[2 * i + 1] = (float)((b - a) / ((h - e) * (1 << 23)));
        } else {
          g_log_tbl[2 * i + 1] = (float)(1 / (x * (1 << 23)));
        }
      }
      g_ready = 1;
    }
  }
}

double fmr_expd(double x) {
  /* fmath.hpp:439-462 (the SSE2 branch; arithmetic is plain IEEE double, one rounding per op) */
  if (x <= -708.39641853226408) return 0;
  if (x >= 709.78271289338397) return INFINITY;
  const double b = (double)(3ULL << 51);
  volatile double prod = x * g_expd_a;     /* volatile: forbid contraction whatever the flags */
  double d = prod + b;
  uint64_t dbits;
  memcpy(&dbits, &d, 8);
  uint64_t di = (uint64_t)(int64_t)(int32_t)(uint32_t)dbits;   /* _mm_cvtsi128_si32, sign-extended */
  uint64_t iax = g_expd_tbl[di & (EXPD_S - 1)];
  volatile double back = (d - b) * g_expd_ra;
  double t = back - x;
  const uint64_t adj = (1ULL << (EXPD_SBIT + 10)) - (1ULL << EXPD_SBIT);
  uint64_t u = ((di + adj) >> EXPD_SBIT) << 52;
  volatile double tt = t * t;
  volatile double p1 = (3.0000000027955394 - t) * tt;
  volatile double p2 = p1 * 0.16666666685227835064;
  double y = p2 - t + 1.0;
  u |= iax;
  double scale;
  memcpy(&scale, &u, 8);
  return y * scale;
}

float fmr_logf(float x) {
  /* fmath.hpp:738-752 */
  uint32_t bits;
  memcpy(&bits, &x, 4);
  int a = (int)(bits & (0xFFu << 23));
  uint32_t b1 = bits & (((1u << LOG_LEN) - 1) << (23 - LOG_LEN));
  uint32_t b2 = bits & ((1u << (23 - LOG_LEN)) - 1);
  uint32_t idx = b1 >> (23 - LOG_LEN);
  volatile float t1 = (float)(a - (127 << 23)) * g_c_log2;
  volatile float t2 = t1 + g_log_tbl[2 * idx];
  volatile float t3 = (float)b2 * g_log_tbl[2 * idx + 1];
  return t2 + t3;
}

const float *fmr_log_table(void) { return g_log_tbl; }
float fmr_log_c_log2(void) { return g_c_log2; }
