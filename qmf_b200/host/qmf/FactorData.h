// Factors (+ optional biases) of one side; host mirror of the device-resident buffers.
// Surface of the reference's qmf/FactorData.h:28-142.
#pragma once
#include <cstdio>
#include <fstream>
#include <string>

#include <qmf/Matrix.h>

namespace qmf {

class FactorData {
 public:
  FactorData(size_t nelems, size_t nfactors, bool withBiases = false)
    : withBiases_(withBiases), factors_(nelems, nfactors), biases_(withBiases ? nelems : 0) {}

  Double at(size_t idx, size_t f) const { return factors_(idx, f); }
  Double& at(size_t idx, size_t f) { return factors_(idx, f); }
  Double biasAt(size_t idx) const { return withBiases_ ? biases_(idx) : 0.0; }
  Double& biasAt(size_t idx) {
    CHECK(withBiases_) << "can't access bias when withBiases = false";
    return biases_(idx);
  }

  template <typename Fn>
  void setFactors(Fn fn) {
    for (size_t i = 0; i < nelems(); ++i) {
      for (size_t f = 0; f < nfactors(); ++f) factors_(i, f) = fn(i, f);
    }
  }
  void setFactors() { factors_.clear(); }

  // one value per line, row-major (idx, factor) order: the --distribution_file of the reference
  // (qmf/FactorData.h:74-100); a short file leaves the remaining entries untouched.  A regular file is mapped and its
  // line ranges are parsed on all host threads (at 1 M items x 128 factors the getline + sscanf
  // loop reads 128 M lines); setFactorsSequential is the reference's loop, kept for streams and as the equality check.
  void setFactors(const std::string& fileName);
  void setFactorsSequential(const std::string& fileName);

  template <typename Fn>
  void setBiases(Fn fn) {
    for (size_t i = 0; i < biases_.size(); ++i) biases_(i) = fn(i);
  }

  size_t nelems() const { return factors_.nrows(); }
  size_t nfactors() const { return factors_.ncols(); }
  bool withBiases() const { return withBiases_; }
  const Matrix& getFactors() const { return factors_; }
  Matrix& getFactors() { return factors_; }
  const Vector& getBiases() const { return biases_; }
  Vector& getBiases() { return biases_; }

 private:
  const bool withBiases_;
  Matrix factors_;
  Vector biases_;
};

}  // namespace qmf
