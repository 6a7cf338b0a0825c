#include <qmf/Matrix.h>

#include <cmath>
#include <utility>

namespace qmf {

namespace {

// unblocked Bunch-Kaufman on the upper triangle; A is symmetric so row-major == column-major
struct SymSolver {
  Matrix& A;
  const long n;
  std::vector<long> piv;

  explicit SymSolver(Matrix& a) : A(a), n(static_cast<long>(a.nrows())), piv(a.nrows(), 0) {}
  Double& at(long i, long j) { return A(static_cast<size_t>(j), static_cast<size_t>(i)); }  // element (i,j), i <= j

  long argmaxCol(long j, long hi) {
    long best = 0;
    Double bv = -1.0;
    for (long i = 0; i < hi; ++i) {
      const Double v = std::fabs(at(i, j));
      if (v > bv) {
        bv = v;
        best = i;
      }
    }
    return best;
  }

  int factor() {
    const Double alpha = (1.0 + std::sqrt(17.0)) / 8.0;
    int info = 0;
    for (long k = n - 1; k >= 0;) {
      long step = 1, kp = k;
      const Double akk = std::fabs(at(k, k));
      long imax = 0;
      Double colmax = 0.0;
      if (k > 0) {
        imax = argmaxCol(k, k);
        colmax = std::fabs(at(imax, k));
      }
      if (std::fmax(akk, colmax) == 0.0 || std::isnan(akk)) {
        if (info == 0) info = static_cast<int>(k + 1);
      } else {
        if (akk < alpha * colmax) {
          Double rowmax = 0.0;
          for (long j = imax + 1; j <= k; ++j) rowmax = std::fmax(rowmax, std::fabs(at(imax, j)));
          if (imax > 0) rowmax = std::fmax(rowmax, std::fabs(at(argmaxCol(imax, imax), imax)));
          if (akk >= alpha * colmax * (colmax / rowmax)) {
            kp = k;
          } else if (std::fabs(at(imax, imax)) >= alpha * rowmax) {
            kp = imax;
          } else {
            kp = imax;
            step = 2;
          }
        }
        const long kk = k - step + 1;
        if (kp != kk) {
          for (long i = 0; i < kp; ++i) std::swap(at(i, kk), at(i, kp));
          for (long j = kp + 1; j < kk; ++j) std::swap(at(j, kk), at(kp, j));
          std::swap(at(kk, kk), at(kp, kp));
          if (step == 2) std::swap(at(k - 1, k), at(kp, k));
        }
        if (step == 1) {
          const Double r = 1.0 / at(k, k);
          for (long j = 0; j < k; ++j) {
            const Double t = -r * at(j, k);
            if (t != 0.0) {
              for (long i = 0; i <= j; ++i) at(i, j) += at(i, k) * t;
            }
          }
          for (long i = 0; i < k; ++i) at(i, k) *= r;
        } else if (k > 1) {
          Double d12 = at(k - 1, k);
          const Double d22 = at(k - 1, k - 1) / d12, d11 = at(k, k) / d12;
          d12 = (1.0 / (d11 * d22 - 1.0)) / d12;
          for (long j = k - 2; j >= 0; --j) {
            const Double wm = d12 * (d11 * at(j, k - 1) - at(j, k));
            const Double wk = d12 * (d22 * at(j, k) - at(j, k - 1));
            for (long i = j; i >= 0; --i) at(i, j) = at(i, j) - at(i, k) * wk - at(i, k - 1) * wm;
            at(j, k) = wk;
            at(j, k - 1) = wm;
          }
        }
      }
      if (step == 1) {
        piv[k] = kp + 1;
      } else {
        piv[k] = piv[k - 1] = -(kp + 1);
      }
      k -= step;
    }
    return info;
  }

  void solve(Vector& b) {
    for (long k = n - 1; k >= 0;) {
      if (piv[k] > 0) {
        if (piv[k] - 1 != k) std::swap(b(k), b(piv[k] - 1));
        for (long i = 0; i < k; ++i) b(i) -= at(i, k) * b(k);
        b(k) /= at(k, k);
        k -= 1;
      } else {
        if (-piv[k] - 1 != k - 1) std::swap(b(k - 1), b(-piv[k] - 1));
        for (long i = 0; i < k - 1; ++i) b(i) -= at(i, k) * b(k) + at(i, k - 1) * b(k - 1);
        const Double e = at(k - 1, k), am = at(k - 1, k - 1) / e, ak = at(k, k) / e, den = am * ak - 1.0;
        const Double bm = b(k - 1) / e, bk = b(k) / e;
        b(k - 1) = (ak * bm - bk) / den;
        b(k) = (am * bk - bm) / den;
        k -= 2;
      }
    }
    for (long k = 0; k < n;) {
      const long w = piv[k] > 0 ? 1 : 2;
      for (long c = 0; c < w; ++c) {
        Double s = 0.0;
        for (long i = 0; i < k; ++i) s += at(i, k + c) * b(i);
        b(k + c) -= s;
      }
      const long kp = (piv[k] > 0 ? piv[k] : -piv[k]) - 1;
      if (kp != k) std::swap(b(k), b(kp));
      k += w;
    }
  }
};

}  // namespace

Vector linearSymmetricSolve(Matrix A, Vector b) {
  CHECK_EQ(A.nrows(), A.ncols()) << "A should be squared";
  CHECK_EQ(A.nrows(), b.size()) << "b should have the same number of rows as A";
  SymSolver s(A);
  const int info = s.factor();
  CHECK_EQ(info, 0) << "symmetric solve failed, code " << info;
  s.solve(b);
  return b;
}

}  // namespace qmf
