// ORACLE -- TEST INFRASTRUCTURE ONLY.  A minimal stand-in for <gtest/gtest.h> (googletest is fetched
// from the network by the reference's CMakeLists.txt:91-97 and is absent here): just enough for the
// reference's tests/ocp_tests.cpp to compile unmodified and run.  Provides its own main().
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace testing_shim {
struct Case {
  std::string name;
  std::function<void()> fn;
};
inline std::vector<Case>& cases() {
  static std::vector<Case> c;
  return c;
}
inline int& failures() {
  static int f = 0;
  return f;
}
struct Registrar {
  Registrar(const char* suite, const char* name, void (*fn)()) { cases().push_back({std::string(suite) + "." + name, fn}); }
};
inline void fail(const char* file, int line, const char* what) {
  std::printf("  FAILED %s:%d: %s\n", file, line, what);
  ++failures();
}
// googletest's EXPECT_DOUBLE_EQ: within 4 units in the last place
inline bool almost_equal_4ulp(double a, double b) {
  if (std::isnan(a) || std::isnan(b)) return false;
  if (a == b) return true;
  std::int64_t ia, ib;
  std::memcpy(&ia, &a, 8);
  std::memcpy(&ib, &b, 8);
  if ((ia < 0) != (ib < 0)) return false;
  const std::int64_t d = ia > ib ? ia - ib : ib - ia;
  return d <= 4;
}
}  // namespace testing_shim

#define TEST(suite, name)                                                                  \
  static void suite##_##name##_body();                                                     \
  static testing_shim::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
  static void suite##_##name##_body()

#define MAS_SHIM_CHECK(cond, text, fatal)                      \
  do {                                                         \
    if (!(cond)) {                                             \
      testing_shim::fail(__FILE__, __LINE__, text);            \
      if (fatal) return;                                       \
    }                                                          \
  } while (0)

#define EXPECT_TRUE(c) MAS_SHIM_CHECK((c), "EXPECT_TRUE(" #c ")", false)
#define ASSERT_TRUE(c) MAS_SHIM_CHECK((c), "ASSERT_TRUE(" #c ")", true)
#define EXPECT_FALSE(c) MAS_SHIM_CHECK(!(c), "EXPECT_FALSE(" #c ")", false)
#define EXPECT_EQ(a, b) MAS_SHIM_CHECK((a) == (b), "EXPECT_EQ(" #a ", " #b ")", false)
#define ASSERT_EQ(a, b) MAS_SHIM_CHECK((a) == (b), "ASSERT_EQ(" #a ", " #b ")", true)
#define EXPECT_DOUBLE_EQ(a, b) MAS_SHIM_CHECK(testing_shim::almost_equal_4ulp((a), (b)), "EXPECT_DOUBLE_EQ(" #a ", " #b ")", false)
#define EXPECT_NEAR(a, b, tol) MAS_SHIM_CHECK(std::fabs((a) - (b)) <= (tol), "EXPECT_NEAR(" #a ", " #b ", " #tol ")", false)

int main() {
  int failed_cases = 0;
  for (auto& c : testing_shim::cases()) {
    const int before = testing_shim::failures();
    c.fn();
    const bool ok = testing_shim::failures() == before;
    std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", c.name.c_str());
    failed_cases += !ok;
  }
  std::printf("%zu tests, %d failed\n", testing_shim::cases().size(), failed_cases);
  return failed_cases ? 1 : 0;
}
