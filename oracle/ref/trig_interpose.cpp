// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference calls std::sin / std::cos / std::tan (examples/models/single_track_model.hpp:38-40,
// pendulum_model.hpp:18, ...), i.e. glibc's libm.  libref.so is linked with -Bsymbolic-functions and
// defines sin / cos / tan / sincos itself, so the reference's unmodified calls land here:
//   mode 0 (TRIG_GLIBC)    -> forwarded to the next definition in link order (glibc), the reference's
//                             own arithmetic;
//   mode 1 (TRIG_PORTABLE) -> include/mas_b200/portable_math.h, the bit-reproducible implementation
//                             the GPU kernels use.  This is the only substitution made to the
//                             reference build; it lets "GPU == oracle(portable) == reference(portable)"
//                             be asserted bit for bit, while mode 0 shows what libm's last bit moves.
// (gcc -O3 merges sin(x) and cos(x) of one argument into a sincos call, hence sincos.)
#include <dlfcn.h>

#include "mas_b200/portable_math.h"

namespace {
int g_mode = 0;
// per-thread override (-1 = none): the example builders compute their initial guesses (pendulum_swing_up.cpp:110-113) with
// libm when the program starts; the wrapper builds OCPs inside parallel loops, so it pins the builder call to libm per thread
thread_local int t_override = -1;
inline int mode() { return t_override >= 0 ? t_override : g_mode; }
typedef double (*fn1)(double);
typedef void (*fn_sc)(double, double*, double*);
fn1 next_sin() {
  static fn1 f = reinterpret_cast<fn1>(dlsym(RTLD_NEXT, "sin"));
  return f;
}
fn1 next_cos() {
  static fn1 f = reinterpret_cast<fn1>(dlsym(RTLD_NEXT, "cos"));
  return f;
}
fn1 next_tan() {
  static fn1 f = reinterpret_cast<fn1>(dlsym(RTLD_NEXT, "tan"));
  return f;
}
fn_sc next_sincos() {
  static fn_sc f = reinterpret_cast<fn_sc>(dlsym(RTLD_NEXT, "sincos"));
  return f;
}
}  // namespace

extern "C" {
void ref_set_trig_mode(int mode) { g_mode = mode; }
void ref_set_thread_trig_override(int mode) { t_override = mode; }
int ref_get_trig_mode() { return g_mode; }

double sin(double x) { return mode() ? mas_b200::pm::sin_(x) : next_sin()(x); }
double cos(double x) { return mode() ? mas_b200::pm::cos_(x) : next_cos()(x); }
double tan(double x) { return mode() ? mas_b200::pm::tan_(x) : next_tan()(x); }
void sincos(double x, double* s, double* c) {
  if (mode())
    mas_b200::pm::sincos_(x, s, c);
  else
    next_sincos()(x, s, c);
}
}
