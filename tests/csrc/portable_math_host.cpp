// TEST HARNESS: exposes include/mas_b200/portable_math.h to ctypes for tests/test_portable_math.py.
#include "mas_b200/portable_math.h"

extern "C" void pm_eval(const double* x, int n, double* s, double* c, double* t) {
  for (int i = 0; i < n; ++i) {
    mas_b200::pm::sincos_(x[i], &s[i], &c[i]);
    t[i] = mas_b200::pm::tan_(x[i]);
  }
}

// div_const against the division instruction: returns the number of inputs where the bits differ.
// `which` selects the compile-time divisor exactly as the kernels spell it.
extern "C" long pm_div_const_mismatches(int which, const double* a, long n) {
  const double e5 = 1e-5, e6 = 1e-6;
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    double got, want;
    switch (which) {
      case 0: got = MAS_DIV_CONST(a[i], 2.5); want = a[i] / 2.5; break;
      case 1: got = MAS_DIV_CONST(a[i], 6.0); want = a[i] / 6.0; break;
      case 2: got = MAS_DIV_CONST(a[i], 2 * e6); want = a[i] / (2 * e6); break;
      case 3: got = MAS_DIV_CONST(a[i], e5 * e5); want = a[i] / (e5 * e5); break;
      case 4: got = MAS_DIV_CONST(a[i], 4 * e5 * e5); want = a[i] / (4 * e5 * e5); break;
      case 5: got = MAS_DIV_CONST(a[i], 4 * e6 * e6); want = a[i] / (4 * e6 * e6); break;
      case 6: got = MAS_DIV_CONST(a[i], e6 * e6); want = a[i] / (e6 * e6); break;
      default: return -1;
    }
    if (__builtin_memcmp(&got, &want, sizeof(double)) != 0 && !(got != got && want != want)) ++bad;
  }
  return bad;
}
