// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product
// (multi_agent_solver_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// PARITY PIN: this restatement is checked BIT FOR BIT against a build of the reference's own unmodified sources --
// oracle/_ref/libref.so = /root/reference/include/multi_agent_solver/** + examples/*.cpp compiled by oracle/ref/Makefile
// against oracle/eigen_shim (the image has no Eigen 3.4) -- by tests/test_ref_pin.py (configs 1-5, every example
// model, all four strategies, constraints, mixed agents; both libm modes; trajectories, costs, iteration counts, status,
// line-search candidates, regularisation retries) and tests/test_golden.py.  It is also checked against the values the
// reference's own tests hold for the layers under iLQR (tests/ocp_tests.cpp:21-154 -> oracle_selftest.cpp; the
// reference's test file itself runs unmodified on the shim) and closed-form anchors (tests/test_oracle.py).
// What no build in this image can pin: real Eigen's SIMD summation order above its small-size thresholds and its
// blocked LLT at size >= 32; the shim and this file use the sequential order documented below.
//
// dense.hpp: the handful of dense operations the reference takes from Eigen 3.4 (un-vendored system
// package, CMakeLists.txt:12), restated as scalar loops with a fixed, documented operation order:
//   * storage is column-major like Eigen::MatrixXd;
//   * every product coefficient is a k-ascending sequential sum that starts from the first product
//     (Eigen's coefficient-based lazy product, used below its GEMM threshold);
//   * no fused multiply-add anywhere (the reference Release build targets baseline x86-64,
//     scripts/build.sh:99) -- build with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace oracle {

using Vec = std::vector<double>;

struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> d;  // column-major
  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), d(static_cast<std::size_t>(r) * c, 0.0) {}
  double& operator()(int i, int j) { return d[static_cast<std::size_t>(j) * rows + i]; }
  double operator()(int i, int j) const { return d[static_cast<std::size_t>(j) * rows + i]; }
  void set_zero() { for (auto& v : d) v = 0.0; }
  Vec col(int j) const { return Vec(d.begin() + static_cast<std::size_t>(j) * rows, d.begin() + static_cast<std::size_t>(j + 1) * rows); }
  void set_col(int j, const Vec& v) { for (int i = 0; i < rows; ++i) (*this)(i, j) = v[i]; }
  static Mat identity(int n) { Mat m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
};

inline Vec zeros(int n) { return Vec(static_cast<std::size_t>(n), 0.0); }

// a + b, a - b, s * a   (coefficient-wise)
inline Vec add(const Vec& a, const Vec& b) { Vec r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i]; return r; }
inline Vec sub(const Vec& a, const Vec& b) { Vec r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = a[i] - b[i]; return r; }
inline Vec scale(double s, const Vec& a) { Vec r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = s * a[i]; return r; }
inline Mat add(const Mat& a, const Mat& b) { Mat r(a.rows, a.cols); for (std::size_t i = 0; i < a.d.size(); ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
inline Mat sub(const Mat& a, const Mat& b) { Mat r(a.rows, a.cols); for (std::size_t i = 0; i < a.d.size(); ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
inline Mat scale(double s, const Mat& a) { Mat r(a.rows, a.cols); for (std::size_t i = 0; i < a.d.size(); ++i) r.d[i] = s * a.d[i]; return r; }
inline Mat neg(const Mat& a) { Mat r(a.rows, a.cols); for (std::size_t i = 0; i < a.d.size(); ++i) r.d[i] = -a.d[i]; return r; }

inline Mat transpose(const Mat& a) {
  Mat r(a.cols, a.rows);
  for (int j = 0; j < a.cols; ++j) for (int i = 0; i < a.rows; ++i) r(j, i) = a(i, j);
  return r;
}

// C = A * B, each coefficient = ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + ...
inline Mat matmul(const Mat& a, const Mat& b) {
  Mat c(a.rows, b.cols);
  for (int j = 0; j < b.cols; ++j)
    for (int i = 0; i < a.rows; ++i) {
      double s = a(i, 0) * b(0, j);
      for (int k = 1; k < a.cols; ++k) s = s + a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}

// y = A * x
inline Vec matvec(const Mat& a, const Vec& x) {
  Vec y(static_cast<std::size_t>(a.rows));
  for (int i = 0; i < a.rows; ++i) {
    double s = a(i, 0) * x[0];
    for (int k = 1; k < a.cols; ++k) s = s + a(i, k) * x[k];
    y[i] = s;
  }
  return y;
}

// y = A^T * x
inline Vec matTvec(const Mat& a, const Vec& x) {
  Vec y(static_cast<std::size_t>(a.cols));
  for (int i = 0; i < a.cols; ++i) {
    double s = a(0, i) * x[0];
    for (int k = 1; k < a.rows; ++k) s = s + a(k, i) * x[k];
    y[i] = s;
  }
  return y;
}

// C = A^T * B
inline Mat matTmul(const Mat& a, const Mat& b) {
  Mat c(a.cols, b.cols);
  for (int j = 0; j < b.cols; ++j)
    for (int i = 0; i < a.cols; ++i) {
      double s = a(0, i) * b(0, j);
      for (int k = 1; k < a.rows; ++k) s = s + a(k, i) * b(k, j);
      c(i, j) = s;
    }
  return c;
}

inline double dot(const Vec& a, const Vec& b) {
  double s = 0.0;
  for (std::size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}

// Frobenius norm, column-major sequential sum of squares (Eigen: MatrixBase::norm, nash.hpp:219)
inline double frobenius_norm(const Mat& a) {
  double s = 0.0;
  for (std::size_t i = 0; i < a.d.size(); ++i) s += a.d[i] * a.d[i];
  return std::sqrt(s);
}

// In-place `m = 0.5 * (m + m.transpose())` exactly as the reference writes it (ilqr.hpp:102,192).
// The expression aliases its destination; Eigen's NDEBUG dense assignment walks columns then rows
// without a temporary, so the strict lower triangle becomes 0.5*(a_ij + a_ji) and the strict upper
// triangle then reads the *already updated* lower entry: 0.5*(a_ij + new a_ji)  (SURVEY 8a quirk 3).
inline void symmetrize_aliased(Mat& m) {
  for (int j = 0; j < m.cols; ++j)
    for (int i = 0; i < m.rows; ++i) m(i, j) = 0.5 * (m(i, j) + m(j, i));
}

// Eigen::LLT<MatrixXd, Lower> restated (unblocked algorithm, used for sizes < 32): reads the lower
// triangle only; fails at column k iff  x = a_kk - sum_j L_kj^2  satisfies x <= 0 (a NaN passes).
struct LLT {
  Mat L;
  bool ok = false;
  bool compute(const Mat& a) {
    const int n = a.rows;
    L = a;
    ok = true;
    for (int k = 0; k < n; ++k) {
      double x = L(k, k);
      if (k > 0) {
        double sq = 0.0;
        for (int j = 0; j < k; ++j) sq += L(k, j) * L(k, j);
        x -= sq;
      }
      if (x <= 0.0) { ok = false; return false; }
      x = std::sqrt(x);
      L(k, k) = x;
      for (int i = k + 1; i < n; ++i) {
        double s = L(i, k);
        if (k > 0) {
          double acc = L(i, 0) * L(k, 0);
          for (int j = 1; j < k; ++j) acc = acc + L(i, j) * L(k, j);
          s -= acc;
        }
        L(i, k) = s / x;
      }
    }
    return true;
  }
  // X = A^{-1} B via L y = b (forward), L^T x = y (backward), column by column.
  Mat solve(const Mat& b) const {
    const int n = L.rows;
    Mat x = b;
    for (int c = 0; c < b.cols; ++c) {
      for (int i = 0; i < n; ++i) {
        double s = x(i, c);
        for (int j = 0; j < i; ++j) s -= L(i, j) * x(j, c);
        x(i, c) = s / L(i, i);
      }
      for (int i = n - 1; i >= 0; --i) {
        double s = x(i, c);
        for (int j = i + 1; j < n; ++j) s -= L(j, i) * x(j, c);
        x(i, c) = s / L(i, i);
      }
    }
    return x;
  }
};

}  // namespace oracle
