// Row-major FP64 matrix / vector containers with the reference's surface (qmf/Matrix.h:27-93,
// qmf/Vector.h).  They are the HOST mirror of the device buffers the engines own; the WALS solve
// itself runs on the GPU.  linearSymmetricSolve stays a host routine for API completeness.
#pragma once
#include <vector>

#include <qmf/Types.h>
#include <qmf/utils/Log.h>

namespace qmf {

class Vector {
 public:
  using value_type = Double;
  explicit Vector(size_t n) : v_(n, 0.0) {}
  Double operator()(size_t i) const { return v_[i]; }
  Double& operator()(size_t i) { return v_[i]; }
  size_t size() const { return v_.size(); }
  Double* data() { return v_.data(); }
  const Double* data() const { return v_.data(); }

 private:
  std::vector<Double> v_;
};

class Matrix {
 public:
  using value_type = Double;
  Matrix(size_t nrows, size_t ncols) : nrows_(nrows), ncols_(ncols), v_(nrows * ncols, 0.0) {
    CHECK_GT(nrows * ncols, 0u) << "matrix's dimensions should be positive";
  }
  Double operator()(size_t r, size_t c) const { return v_[r * ncols_ + c]; }
  Double& operator()(size_t r, size_t c) { return v_[r * ncols_ + c]; }
  size_t nrows() const { return nrows_; }
  size_t ncols() const { return ncols_; }
  void clear() { v_.assign(v_.size(), 0.0); }
  Double* data() { return v_.data(); }
  const Double* data() const { return v_.data(); }
  Double* data(size_t r) { return v_.data() + r * ncols_; }

  Matrix transpose() const {
    Matrix t(ncols_, nrows_);
    for (size_t r = 0; r < nrows_; ++r) {
      for (size_t c = 0; c < ncols_; ++c) t(c, r) = (*this)(r, c);
    }
    return t;
  }

  Matrix operator+(const Matrix& o) const {
    CHECK_EQ(nrows_, o.nrows_);
    CHECK_EQ(ncols_, o.ncols_);
    Matrix s(nrows_, ncols_);
    for (size_t e = 0; e < v_.size(); ++e) s.v_[e] = v_[e] + o.v_[e];
    return s;
  }

 private:
  size_t nrows_, ncols_;
  std::vector<Double> v_;
};

// Solves A x = b for symmetric (possibly indefinite) A: Bunch-Kaufman U D U^T, the algorithm
// behind the LAPACK dsysv_ call of the reference (qmf/Matrix.cpp:81-96).
Vector linearSymmetricSolve(Matrix A, Vector b);

}  // namespace qmf
