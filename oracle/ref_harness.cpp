// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// C-ABI harness around the UNMODIFIED reference sources in /root/reference (compiled where they
// lie by oracle/Makefile into oracle/_ref/libqmf_ref.so).  It exposes the reference's own
// implementation of every function on the hot path so that tests/ and
// bench.py's cpu_baseline / --impl reference leg can (a) validate the C restatement in
// oracle/qmf_oracle.c and (b) time the reference on the GPU box's host cores.
//
// Private members are reached through the friendships the reference declares for its own
// gtests: FRIEND_TEST(WALSEngine, init) (qmf/wals/WALSEngine.h:138) expands to
// `friend class WALSEngine_init_Test;`, so defining a class of that name in namespace qmf
// grants access without touching reference sources (same for BPREngine.h:154 and
// Engine.h:93-95).  No reference code is copied here: every numeric result below is produced
// by calling the reference's functions.
#include <qmf/DatasetReader.h>
#include <qmf/Engine.h>
#include <qmf/bpr/BPREngine.h>
#include <qmf/metrics/MetricsEngine.h>
#include <qmf/metrics/MetricsManager.h>
#include <qmf/wals/WALSEngine.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace qmf {

// ---- friend of WALSEngine (WALSEngine.h:138) -------------------------------------------------
class WALSEngine_init_Test {
 public:
  static FactorData& userFactors(WALSEngine& e) { return *e.userFactors_; }
  static FactorData& itemFactors(WALSEngine& e) { return *e.itemFactors_; }
  static const IdIndex& userIndex(WALSEngine& e) { return e.userIndex_; }
  static const IdIndex& itemIndex(WALSEngine& e) { return e.itemIndex_; }
  // one half-step through the reference's private iterate() (WALSEngine.cpp:165-218)
  static double halfStep(WALSEngine& e, int side) {
    if (side == 0) {
      return e.iterate(*e.userFactors_, e.userIndex_, e.userSignals_, *e.itemFactors_, e.itemIndex_);
    }
    return e.iterate(*e.itemFactors_, e.itemIndex_, e.itemSignals_, *e.userFactors_, e.userIndex_);
  }
  static int64_t nnz(WALSEngine& e, int side) {
    const auto& sig = side == 0 ? e.userSignals_ : e.itemSignals_;
    int64_t n = 0;
    for (const auto& g : sig) n += static_cast<int64_t>(g.group.size());
    return n;
  }
  // dump the AoS signal groups (WALSEngine.h:67-75) as CSR with dense column idx
  static void csr(WALSEngine& e, int side, int64_t* rowPtr, int64_t* rowId, int32_t* colIdx, int64_t* colId,
                  double* val) {
    const auto& sig = side == 0 ? e.userSignals_ : e.itemSignals_;
    const IdIndex& right = side == 0 ? e.itemIndex_ : e.userIndex_;
    int64_t p = 0;
    for (size_t r = 0; r < sig.size(); ++r) {
      rowPtr[r] = p;
      rowId[r] = sig[r].sourceId;
      for (const auto& s : sig[r].group) {
        colIdx[p] = static_cast<int32_t>(right.idx(s.id));
        colId[p] = s.id;
        val[p] = s.value;
        ++p;
      }
    }
    rowPtr[sig.size()] = p;
  }
  static Matrix gramRaceFree(WALSEngine& e, const Matrix& X) { return e.computeXtX(X); }
  static void gramUsed(WALSEngine& e, const Matrix& X, Matrix* out) { e.computeXtX(X, out); }
  // the reference's static per-row build+solve+loss (WALSEngine.cpp:266-310) on caller data
  static double updateOne(Matrix& X, size_t leftIdx, const Matrix& Y, const int32_t* cols, const double* vals,
                          int64_t nnz, const Matrix& YtY, double alpha, double lambda) {
    IdIndex leftIndex, rightIndex;
    for (size_t i = 0; i < X.nrows(); ++i) leftIndex.getOrSetIdx(static_cast<int64_t>(i));
    for (size_t i = 0; i < Y.nrows(); ++i) rightIndex.getOrSetIdx(static_cast<int64_t>(i));
    WALSEngine::SignalGroup g;
    g.sourceId = static_cast<int64_t>(leftIdx);
    for (int64_t s = 0; s < nnz; ++s) g.group.push_back(WALSEngine::Signal{cols[s], vals[s]});
    return WALSEngine::updateFactorsForOne(X, leftIndex, Y, rightIndex, g, YtY, alpha, lambda);
  }
  // many rows with shared (identity) indexes, spread over the reference's own ParallelExecutor
  static double updateRows(WALSEngine& e, Matrix& X, const Matrix& Y, const int64_t* rowPtr, const int32_t* cols,
                           const double* vals, int64_t nrows, const Matrix& YtY, double alpha, double lambda) {
    IdIndex leftIndex, rightIndex;
    for (size_t i = 0; i < X.nrows(); ++i) leftIndex.getOrSetIdx(static_cast<int64_t>(i));
    for (size_t i = 0; i < Y.nrows(); ++i) rightIndex.getOrSetIdx(static_cast<int64_t>(i));
    std::vector<WALSEngine::SignalGroup> groups(nrows);
    for (int64_t r = 0; r < nrows; ++r) {
      groups[r].sourceId = r;
      for (int64_t s = rowPtr[r]; s < rowPtr[r + 1]; ++s) {
        groups[r].group.push_back(WALSEngine::Signal{cols[s], vals[s]});
      }
    }
    auto map = [&](const size_t t) {
      return WALSEngine::updateFactorsForOne(X, leftIndex, Y, rightIndex, groups[t], YtY, alpha, lambda);
    };
    auto reduce = [](double a, double b) { return a + b; };
    return e.parallel_.mapReduce(static_cast<size_t>(nrows), map, reduce, 0.0);
  }
  // same, and the loss term of every row is kept (rowLoss[t]); for the sampled-row parity checks at
  // the BASELINE-sized configs
  static double updateRowsLosses(WALSEngine& e, Matrix& X, const Matrix& Y, const int64_t* rowPtr, const int32_t* cols,
                                 const double* vals, int64_t nrows, const Matrix& YtY, double alpha, double lambda,
                                 double* rowLoss) {
    IdIndex leftIndex, rightIndex;
    for (size_t i = 0; i < X.nrows(); ++i) leftIndex.getOrSetIdx(static_cast<int64_t>(i));
    for (size_t i = 0; i < Y.nrows(); ++i) rightIndex.getOrSetIdx(static_cast<int64_t>(i));
    std::vector<WALSEngine::SignalGroup> groups(nrows);
    for (int64_t r = 0; r < nrows; ++r) {
      groups[r].sourceId = r;
      for (int64_t s = rowPtr[r]; s < rowPtr[r + 1]; ++s) {
        groups[r].group.push_back(WALSEngine::Signal{cols[s], vals[s]});
      }
    }
    auto map = [&](const size_t t) {
      rowLoss[t] = WALSEngine::updateFactorsForOne(X, leftIndex, Y, rightIndex, groups[t], YtY, alpha, lambda);
      return rowLoss[t];
    };
    auto reduce = [](double a, double b) { return a + b; };
    return e.parallel_.mapReduce(static_cast<size_t>(nrows), map, reduce, 0.0);
  }
};

// ---- friend of BPREngine (BPREngine.h:154) ---------------------------------------------------
class BPREngine_init_Test {
 public:
  static void seed(BPREngine& e, uint32_t s) { e.gen_.seed(s); }
  static FactorData& userFactors(BPREngine& e) { return *e.userFactors_; }
  static FactorData& itemFactors(BPREngine& e) { return *e.itemFactors_; }
  static const IdIndex& userIndex(BPREngine& e) { return e.userIndex_; }
  static const IdIndex& itemIndex(BPREngine& e) { return e.itemIndex_; }
  static void update(BPREngine& e, size_t u, size_t i, size_t j) { e.update(BPREngine::PosNegTriplet{u, i, j}); }
  static double predictDifference(BPREngine& e, size_t u, size_t i, size_t j) { return e.predictDifference(u, i, j); }
  static double learningRate(BPREngine& e) { return e.learningRate_; }
  static void setLearningRate(BPREngine& e, double lr) { e.learningRate_ = lr; }
  static int64_t ndata(BPREngine& e) { return static_cast<int64_t>(e.data_.size()); }
  static void data(BPREngine& e, int64_t* u, int64_t* i) {
    for (size_t p = 0; p < e.data_.size(); ++p) {
      u[p] = static_cast<int64_t>(e.data_[p].userIdx);
      i[p] = static_cast<int64_t>(e.data_[p].posItemIdx);
    }
  }
  static int64_t evalSize(BPREngine& e, int test) {
    return static_cast<int64_t>((test ? e.testEvalSet_ : e.evalSet_).size());
  }
  static void evalSet(BPREngine& e, int test, int64_t* u, int64_t* i, int64_t* j) {
    const auto& s = test ? e.testEvalSet_ : e.evalSet_;
    for (size_t p = 0; p < s.size(); ++p) {
      u[p] = static_cast<int64_t>(s[p].userIdx);
      i[p] = static_cast<int64_t>(s[p].posItemIdx);
      j[p] = static_cast<int64_t>(s[p].negItemIdx);
    }
  }
  // the loss half of BPREngine::evaluate (BPREngine.cpp:246-261) composed from the reference's
  // own loss(), predictDifference() and ParallelExecutor::mapReduce(elems) (tail-drop included)
  static double evalLoss(BPREngine& e, int test) {
    const auto& s = test ? e.testEvalSet_ : e.evalSet_;
    if (s.empty()) return -1.0;
    auto f = [&e](const BPREngine::PosNegTriplet& t) {
      return e.loss(e.predictDifference(t.userIdx, t.posItemIdx, t.negItemIdx));
    };
    return e.parallel_.mapReduce(s, f, std::plus<double>(), 0.0) / s.size();
  }
  static int64_t numTestUsers(BPREngine& e) { return static_cast<int64_t>(e.testUsers_.size()); }
  static void testUsers(BPREngine& e, int64_t* out) {
    for (size_t p = 0; p < e.testUsers_.size(); ++p) out[p] = static_cast<int64_t>(e.testUsers_[p]);
  }
};

// ---- friend of Engine (Engine.h:93-95) -------------------------------------------------------
class Engine_computeTestScores_Test {
 public:
  static void scores(std::vector<std::vector<double>>& out, const std::vector<size_t>& users, const FactorData& U,
                     const FactorData& V, ParallelExecutor& p) {
    Engine::computeTestScores(out, users, U, V, p);
  }
  static void initAvg(std::vector<size_t>& users, std::vector<std::vector<double>>& labels,
                      std::vector<std::vector<double>>& scores, const std::vector<DatasetElem>& test,
                      const IdIndex& ui, const IdIndex& ii, size_t n, int32_t seed) {
    Engine::initAvgTestData(users, labels, scores, test, ui, ii, n, seed);
  }
  static void save(const FactorData& f, const IdIndex& idx, std::ostream& os) { Engine::saveFactors(f, idx, os); }
};

}  // namespace qmf

using namespace qmf;

namespace {

std::vector<DatasetElem> makeDataset(const int64_t* u, const int64_t* i, const double* v, int64_t n) {
  std::vector<DatasetElem> d(static_cast<size_t>(n));
  for (int64_t p = 0; p < n; ++p) {
    d[p].userId = u[p];
    d[p].itemId = i[p];
    d[p].value = v[p];
  }
  return d;
}

void fillMatrix(Matrix& M, const double* src) { std::memcpy(M.data(), src, sizeof(double) * M.nrows() * M.ncols()); }
void dumpMatrix(const Matrix& M, double* dst) {
  std::memcpy(dst, const_cast<Matrix&>(M).data(), sizeof(double) * M.nrows() * M.ncols());
}

struct WalsHandle {
  WALSConfig config;
  MetricsConfig metricsConfig;
  std::unique_ptr<MetricsEngine> metrics;
  std::unique_ptr<WALSEngine> engine;
};

struct BprHandle {
  BPRConfig config;
  MetricsConfig metricsConfig;
  std::unique_ptr<MetricsEngine> metrics;
  std::unique_ptr<BPREngine> engine;
};

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------- WALS --
void* ref_wals_create(int64_t nfactors, int64_t nepochs, double lambda, double alpha, int nthreads,
                      const char* testAvgMetrics, int64_t numTestUsers, int testAlways, int32_t evalSeed) {
  auto* h = new WalsHandle{WALSConfig{static_cast<size_t>(nepochs), static_cast<size_t>(nfactors), lambda, alpha, 0.01, ""},
                           MetricsConfig{static_cast<size_t>(numTestUsers), testAlways != 0, evalSeed}, nullptr, nullptr};
  h->metrics = std::make_unique<MetricsEngine>(h->metricsConfig);
  if (testAvgMetrics != nullptr) {
    std::stringstream ss(testAvgMetrics);
    std::string m;
    while (std::getline(ss, m, ',')) {
      if (!m.empty()) CHECK(h->metrics->addTestAvgMetric(m)) << "metric " << m << " is not available";
    }
  }
  h->engine = std::make_unique<WALSEngine>(h->config, h->metrics, static_cast<size_t>(nthreads));
  return h;
}
void ref_wals_destroy(void* hp) { delete static_cast<WalsHandle*>(hp); }

void ref_wals_init(void* hp, const int64_t* u, const int64_t* i, const double* v, int64_t n) {
  static_cast<WalsHandle*>(hp)->engine->init(makeDataset(u, i, v, n));
}
void ref_wals_init_test(void* hp, const int64_t* u, const int64_t* i, const double* v, int64_t n) {
  static_cast<WalsHandle*>(hp)->engine->initTest(makeDataset(u, i, v, n));
}
int64_t ref_wals_nusers(void* hp) { return static_cast<int64_t>(static_cast<WalsHandle*>(hp)->engine->nusers()); }
int64_t ref_wals_nitems(void* hp) { return static_cast<int64_t>(static_cast<WalsHandle*>(hp)->engine->nitems()); }
int64_t ref_wals_nnz(void* hp, int side) { return WALSEngine_init_Test::nnz(*static_cast<WalsHandle*>(hp)->engine, side); }
void ref_wals_ids(void* hp, int side, int64_t* out) {
  auto& e = *static_cast<WalsHandle*>(hp)->engine;
  const IdIndex& idx = side == 0 ? WALSEngine_init_Test::userIndex(e) : WALSEngine_init_Test::itemIndex(e);
  for (size_t p = 0; p < idx.size(); ++p) out[p] = idx.id(p);
}
void ref_wals_csr(void* hp, int side, int64_t* rowPtr, int64_t* rowId, int32_t* colIdx, int64_t* colId, double* val) {
  WALSEngine_init_Test::csr(*static_cast<WalsHandle*>(hp)->engine, side, rowPtr, rowId, colIdx, colId, val);
}
void ref_wals_set_factors(void* hp, int side, const double* src) {
  auto& e = *static_cast<WalsHandle*>(hp)->engine;
  fillMatrix((side == 0 ? WALSEngine_init_Test::userFactors(e) : WALSEngine_init_Test::itemFactors(e)).getFactors(), src);
}
void ref_wals_get_factors(void* hp, int side, double* dst) {
  auto& e = *static_cast<WalsHandle*>(hp)->engine;
  dumpMatrix((side == 0 ? WALSEngine_init_Test::userFactors(e) : WALSEngine_init_Test::itemFactors(e)).getFactors(), dst);
}
double ref_wals_half_step(void* hp, int side) {
  return WALSEngine_init_Test::halfStep(*static_cast<WalsHandle*>(hp)->engine, side);
}
void ref_wals_evaluate(void* hp, int64_t epoch) { static_cast<WalsHandle*>(hp)->engine->evaluate(static_cast<size_t>(epoch)); }
void ref_wals_optimize(void* hp) { static_cast<WalsHandle*>(hp)->engine->optimize(); }
void ref_wals_save(void* hp, const char* userFile, const char* itemFile) {
  auto& e = *static_cast<WalsHandle*>(hp)->engine;
  e.saveUserFactors(userFile);
  e.saveItemFactors(itemFile);
}

// Gram: variant 0 = computeXtX(const Matrix&) (race-free, WALSEngine.cpp:220-244),
//       variant 1 = computeXtX(const Matrix&, Matrix*) (the one iterate() uses, :246-264)
void ref_gram(const double* Y, int64_t n, int64_t k, int nthreads, int variant, double* out) {
  WALSConfig cfg{1, static_cast<size_t>(k), 0.0, 0.0, 0.01, ""};
  std::unique_ptr<MetricsEngine> none;
  WALSEngine e(cfg, none, static_cast<size_t>(nthreads));
  Matrix M(static_cast<size_t>(n), static_cast<size_t>(k));
  fillMatrix(M, Y);
  if (variant == 0) {
    Matrix G = WALSEngine_init_Test::gramRaceFree(e, M);
    dumpMatrix(G, out);
  } else {
    Matrix G(static_cast<size_t>(k), static_cast<size_t>(k));
    WALSEngine_init_Test::gramUsed(e, M, &G);
    dumpMatrix(G, out);
  }
}

// one row of updateFactorsForOne; X (nleft x k) is updated in place at row leftIdx, returns loss term
double ref_wals_update_one(double* X, int64_t nleft, int64_t leftIdx, const double* Y, int64_t nright, int64_t k,
                           const int32_t* cols, const double* vals, int64_t nnz, const double* YtY, double alpha,
                           double lambda) {
  Matrix Xm(static_cast<size_t>(nleft), static_cast<size_t>(k)), Ym(static_cast<size_t>(nright), static_cast<size_t>(k)),
    G(static_cast<size_t>(k), static_cast<size_t>(k));
  fillMatrix(Xm, X);
  fillMatrix(Ym, Y);
  fillMatrix(G, YtY);
  const double loss =
    WALSEngine_init_Test::updateOne(Xm, static_cast<size_t>(leftIdx), Ym, cols, vals, nnz, G, alpha, lambda);
  dumpMatrix(Xm, X);
  return loss;
}

// many rows (CSR over dense right idx) with the reference's thread pool; returns the summed loss
// terms (NOT divided by nusers*nitems) and the wall seconds spent in the row loop.
double ref_wals_update_rows(double* X, int64_t nrows, const double* Y, int64_t nright, int64_t k,
                            const int64_t* rowPtr, const int32_t* cols, const double* vals, const double* YtY,
                            double alpha, double lambda, int nthreads, double* seconds) {
  WALSConfig cfg{1, static_cast<size_t>(k), lambda, alpha, 0.01, ""};
  std::unique_ptr<MetricsEngine> none;
  WALSEngine e(cfg, none, static_cast<size_t>(nthreads));
  Matrix Xm(static_cast<size_t>(nrows), static_cast<size_t>(k)), Ym(static_cast<size_t>(nright), static_cast<size_t>(k)),
    G(static_cast<size_t>(k), static_cast<size_t>(k));
  fillMatrix(Ym, Y);
  fillMatrix(G, YtY);
  const auto t0 = std::chrono::steady_clock::now();
  const double loss = WALSEngine_init_Test::updateRows(e, Xm, Ym, rowPtr, cols, vals, nrows, G, alpha, lambda);
  const auto t1 = std::chrono::steady_clock::now();
  if (seconds != nullptr) *seconds = std::chrono::duration<double>(t1 - t0).count();
  dumpMatrix(Xm, X);
  return loss;
}

// as ref_wals_update_rows, also returning every row's loss term in rowLoss[nrows]
double ref_wals_update_rows_losses(double* X, int64_t nrows, const double* Y, int64_t nright, int64_t k,
                                   const int64_t* rowPtr, const int32_t* cols, const double* vals, const double* YtY,
                                   double alpha, double lambda, int nthreads, double* rowLoss) {
  WALSConfig cfg{1, static_cast<size_t>(k), lambda, alpha, 0.01, ""};
  std::unique_ptr<MetricsEngine> none;
  WALSEngine e(cfg, none, static_cast<size_t>(nthreads));
  Matrix Xm(static_cast<size_t>(nrows), static_cast<size_t>(k)), Ym(static_cast<size_t>(nright), static_cast<size_t>(k)),
    G(static_cast<size_t>(k), static_cast<size_t>(k));
  fillMatrix(Ym, Y);
  fillMatrix(G, YtY);
  const double loss = WALSEngine_init_Test::updateRowsLosses(e, Xm, Ym, rowPtr, cols, vals, nrows, G, alpha, lambda, rowLoss);
  dumpMatrix(Xm, X);
  return loss;
}

// linearSymmetricSolve (Matrix.cpp:81-96): A row-major n x n, b length n -> x
void ref_linear_symmetric_solve(const double* A, const double* b, int64_t n, double* x) {
  Matrix Am(static_cast<size_t>(n), static_cast<size_t>(n));
  fillMatrix(Am, A);
  Vector bv(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) bv(i) = b[i];
  Vector r = linearSymmetricSolve(Am, bv);
  for (int64_t i = 0; i < n; ++i) x[i] = r(i);
}

// -------------------------------------------------------------------------------------- BPR --
void* ref_bpr_create(int64_t nfactors, int64_t nepochs, double lr, double biasLambda, double userLambda,
                     double itemLambda, double decayRate, int useBiases, double initBound, int64_t numNeg,
                     int64_t numHogwild, int shuffle, int64_t evalNumNeg, int32_t evalSeed, int nthreads,
                     const char* testAvgMetrics, int64_t numTestUsers, int testAlways, int64_t genSeed) {
  auto* h = new BprHandle{BPRConfig{static_cast<size_t>(nepochs), static_cast<size_t>(nfactors), lr, biasLambda, userLambda,
                                    itemLambda, decayRate, useBiases != 0, initBound, static_cast<size_t>(numNeg),
                                    static_cast<size_t>(numHogwild), shuffle != 0},
                          MetricsConfig{static_cast<size_t>(numTestUsers), testAlways != 0, evalSeed}, nullptr, nullptr};
  h->metrics = std::make_unique<MetricsEngine>(h->metricsConfig);
  if (testAvgMetrics != nullptr) {
    std::stringstream ss(testAvgMetrics);
    std::string m;
    while (std::getline(ss, m, ',')) {
      if (!m.empty()) CHECK(h->metrics->addTestAvgMetric(m)) << "metric " << m << " is not available";
    }
  }
  h->engine = std::make_unique<BPREngine>(h->config, h->metrics, static_cast<size_t>(evalNumNeg), evalSeed,
                                          static_cast<size_t>(nthreads));
  if (genSeed >= 0) BPREngine_init_Test::seed(*h->engine, static_cast<uint32_t>(genSeed));
  return h;
}
void ref_bpr_destroy(void* hp) { delete static_cast<BprHandle*>(hp); }
void ref_bpr_init(void* hp, const int64_t* u, const int64_t* i, const double* v, int64_t n) {
  static_cast<BprHandle*>(hp)->engine->init(makeDataset(u, i, v, n));
}
void ref_bpr_init_test(void* hp, const int64_t* u, const int64_t* i, const double* v, int64_t n) {
  static_cast<BprHandle*>(hp)->engine->initTest(makeDataset(u, i, v, n));
}
int64_t ref_bpr_nusers(void* hp) { return static_cast<int64_t>(static_cast<BprHandle*>(hp)->engine->nusers()); }
int64_t ref_bpr_nitems(void* hp) { return static_cast<int64_t>(static_cast<BprHandle*>(hp)->engine->nitems()); }
void ref_bpr_ids(void* hp, int side, int64_t* out) {
  auto& e = *static_cast<BprHandle*>(hp)->engine;
  const IdIndex& idx = side == 0 ? BPREngine_init_Test::userIndex(e) : BPREngine_init_Test::itemIndex(e);
  for (size_t p = 0; p < idx.size(); ++p) out[p] = idx.id(p);
}
int64_t ref_bpr_ndata(void* hp) { return BPREngine_init_Test::ndata(*static_cast<BprHandle*>(hp)->engine); }
void ref_bpr_data(void* hp, int64_t* u, int64_t* i) { BPREngine_init_Test::data(*static_cast<BprHandle*>(hp)->engine, u, i); }
int64_t ref_bpr_eval_size(void* hp, int test) { return BPREngine_init_Test::evalSize(*static_cast<BprHandle*>(hp)->engine, test); }
void ref_bpr_eval_set(void* hp, int test, int64_t* u, int64_t* i, int64_t* j) {
  BPREngine_init_Test::evalSet(*static_cast<BprHandle*>(hp)->engine, test, u, i, j);
}
void ref_bpr_get_factors(void* hp, int side, double* dst) {
  auto& e = *static_cast<BprHandle*>(hp)->engine;
  dumpMatrix((side == 0 ? BPREngine_init_Test::userFactors(e) : BPREngine_init_Test::itemFactors(e)).getFactors(), dst);
}
void ref_bpr_set_factors(void* hp, int side, const double* src) {
  auto& e = *static_cast<BprHandle*>(hp)->engine;
  fillMatrix((side == 0 ? BPREngine_init_Test::userFactors(e) : BPREngine_init_Test::itemFactors(e)).getFactors(), src);
}
void ref_bpr_get_biases(void* hp, double* dst) {
  auto& f = BPREngine_init_Test::itemFactors(*static_cast<BprHandle*>(hp)->engine);
  for (size_t p = 0; p < f.nelems(); ++p) dst[p] = static_cast<const FactorData&>(f).biasAt(p);
}
void ref_bpr_set_biases(void* hp, const double* src) {
  auto& f = BPREngine_init_Test::itemFactors(*static_cast<BprHandle*>(hp)->engine);
  for (size_t p = 0; p < f.nelems(); ++p) f.biasAt(p) = src[p];
}
void ref_bpr_update(void* hp, int64_t u, int64_t i, int64_t j) {
  BPREngine_init_Test::update(*static_cast<BprHandle*>(hp)->engine, static_cast<size_t>(u), static_cast<size_t>(i),
                              static_cast<size_t>(j));
}
double ref_bpr_predict_difference(void* hp, int64_t u, int64_t i, int64_t j) {
  return BPREngine_init_Test::predictDifference(*static_cast<BprHandle*>(hp)->engine, static_cast<size_t>(u),
                                                static_cast<size_t>(i), static_cast<size_t>(j));
}
double ref_bpr_learning_rate(void* hp) { return BPREngine_init_Test::learningRate(*static_cast<BprHandle*>(hp)->engine); }
void ref_bpr_set_learning_rate(void* hp, double lr) {
  BPREngine_init_Test::setLearningRate(*static_cast<BprHandle*>(hp)->engine, lr);
}
double ref_bpr_eval_loss(void* hp, int test) { return BPREngine_init_Test::evalLoss(*static_cast<BprHandle*>(hp)->engine, test); }
// BPREngine::optimize() runs config.nepochs epochs (SGD pass, evaluate, lr decay, shuffle;
// BPREngine.cpp:146-176); state carries over between calls, so create with nepochs=1 to step.
double ref_bpr_optimize(void* hp) {
  const auto t0 = std::chrono::steady_clock::now();
  static_cast<BprHandle*>(hp)->engine->optimize();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int64_t ref_bpr_num_test_users(void* hp) { return BPREngine_init_Test::numTestUsers(*static_cast<BprHandle*>(hp)->engine); }
void ref_bpr_test_users(void* hp, int64_t* out) { BPREngine_init_Test::testUsers(*static_cast<BprHandle*>(hp)->engine, out); }

// ------------------------------------------------------------------------------- evaluation --
// Engine::computeTestScores (Engine.cpp:73-96); biases may be null
void ref_compute_test_scores(const double* U, int64_t nu, const double* V, int64_t ni, int64_t k,
                             const double* biases, const int64_t* testUsers, int64_t nT, int nthreads, double* out) {
  FactorData Uf(static_cast<size_t>(nu), static_cast<size_t>(k));
  FactorData Vf(static_cast<size_t>(ni), static_cast<size_t>(k), biases != nullptr);
  fillMatrix(Uf.getFactors(), U);
  fillMatrix(Vf.getFactors(), V);
  if (biases != nullptr) {
    for (int64_t p = 0; p < ni; ++p) Vf.biasAt(p) = biases[p];
  }
  std::vector<size_t> users(static_cast<size_t>(nT));
  for (int64_t t = 0; t < nT; ++t) users[t] = static_cast<size_t>(testUsers[t]);
  std::vector<std::vector<double>> scores(static_cast<size_t>(nT), std::vector<double>(static_cast<size_t>(ni)));
  ParallelExecutor p(static_cast<size_t>(nthreads));
  Engine_computeTestScores_Test::scores(scores, users, Uf, Vf, p);
  for (int64_t t = 0; t < nT; ++t) std::memcpy(out + t * ni, scores[t].data(), sizeof(double) * ni);
}

// Engine::initAvgTestData (Engine.cpp:27-71): which users are test users, in what order, and
// the dense label rows.  userIds/itemIds give the train index order (idx -> raw id).
int64_t ref_init_avg_test_data(const int64_t* userIds, int64_t nu, const int64_t* itemIds, int64_t ni,
                               const int64_t* tu, const int64_t* ti, const double* tv, int64_t nt,
                               int64_t numTestUsers, int32_t seed, int64_t* testUsersOut, double* labelsOut) {
  IdIndex ui, ii;
  for (int64_t p = 0; p < nu; ++p) ui.getOrSetIdx(userIds[p]);
  for (int64_t p = 0; p < ni; ++p) ii.getOrSetIdx(itemIds[p]);
  std::vector<size_t> users;
  std::vector<std::vector<double>> labels, scores;
  Engine_computeTestScores_Test::initAvg(users, labels, scores, makeDataset(tu, ti, tv, nt), ui, ii,
                                         static_cast<size_t>(numTestUsers), seed);
  if (testUsersOut != nullptr) {
    for (size_t t = 0; t < users.size(); ++t) testUsersOut[t] = static_cast<int64_t>(users[t]);
  }
  if (labelsOut != nullptr) {
    for (size_t t = 0; t < users.size(); ++t) std::memcpy(labelsOut + t * ni, labels[t].data(), sizeof(double) * ni);
  }
  return static_cast<int64_t>(users.size());
}

// Metric::compute for one user (Metrics.cpp:54-164); returns NaN if the metric name is unknown
double ref_metric_one(const char* name, const double* labels, const double* scores, int64_t n) {
  if (!MetricsManager::get().exists(name)) return std::nan("");
  std::vector<double> l(labels, labels + n), s(scores, scores + n);
  return MetricsManager::get().getMetric(name)->compute(l, s);
}
// per-user average (Metrics.cpp:27-52): nthreads==0 -> serial overload, else the parallel one
double ref_metric_avg(const char* name, const double* labels, const double* scores, int64_t nT, int64_t ni,
                      int nthreads) {
  if (!MetricsManager::get().exists(name)) return std::nan("");
  std::vector<std::vector<double>> l(static_cast<size_t>(nT)), s(static_cast<size_t>(nT));
  for (int64_t t = 0; t < nT; ++t) {
    l[t].assign(labels + t * ni, labels + (t + 1) * ni);
    s[t].assign(scores + t * ni, scores + (t + 1) * ni);
  }
  const auto& m = MetricsManager::get().getMetric(name);
  if (nthreads <= 0) return m->compute(l, s);
  ParallelExecutor p(static_cast<size_t>(nthreads));
  return m->compute(l, s, p);
}

// Engine::saveFactors (Engine.cpp:105-122) into a caller buffer; returns bytes needed
int64_t ref_save_factors(const double* F, const double* biases, const int64_t* ids, int64_t n, int64_t k, char* out,
                         int64_t cap) {
  FactorData f(static_cast<size_t>(n), static_cast<size_t>(k), biases != nullptr);
  fillMatrix(f.getFactors(), F);
  IdIndex idx;
  for (int64_t p = 0; p < n; ++p) {
    idx.getOrSetIdx(ids[p]);
    if (biases != nullptr) f.biasAt(p) = biases[p];
  }
  std::ostringstream os;
  Engine_computeTestScores_Test::save(f, idx, os);
  const std::string s = os.str();
  if (out != nullptr && static_cast<int64_t>(s.size()) <= cap) std::memcpy(out, s.data(), s.size());
  return static_cast<int64_t>(s.size());
}

// DatasetReader::readAll (DatasetReader.cpp:44-51); call with null outputs to get the count
int64_t ref_read_dataset(const char* fileName, int64_t* u, int64_t* i, double* v, int64_t cap) {
  DatasetReader r(fileName);
  const auto d = r.readAll();
  if (u != nullptr) {
    for (size_t p = 0; p < d.size() && static_cast<int64_t>(p) < cap; ++p) {
      u[p] = d[p].userId;
      i[p] = d[p].itemId;
      v[p] = d[p].value;
    }
  }
  return static_cast<int64_t>(d.size());
}

void ref_set_min_log_level(int level) { FLAGS_minloglevel = level; }

}  // extern "C"
