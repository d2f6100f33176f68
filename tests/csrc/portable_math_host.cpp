// TEST HARNESS: exposes include/mas_b200/portable_math.h to ctypes for tests/test_portable_math.py.
#include "mas_b200/portable_math.h"

extern "C" void pm_eval(const double* x, int n, double* s, double* c, double* t) {
  for (int i = 0; i < n; ++i) {
    mas_b200::pm::sincos_(x[i], &s[i], &c[i]);
    t[i] = mas_b200::pm::tan_(x[i]);
  }
}
