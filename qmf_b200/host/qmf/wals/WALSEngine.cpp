#include <qmf/wals/WALSEngine.h>

#include <algorithm>
#include <random>
#include <thread>

#include "qmf_b200.h"

namespace qmf {

namespace {

#define QMFB_OK_OR_DIE(call)                                        \
  do {                                                              \
    const int qmfb_rc_ = (call);                                    \
    CHECK_EQ(qmfb_rc_, 0) << #call << ": " << qmfb_last_error();    \
  } while (0)

}  // namespace

WALSEngine::WALSEngine(const WALSConfig& config, const std::unique_ptr<MetricsEngine>& metricsEngine, const size_t nthreads)
  : config_(config), metricsEngine_(metricsEngine), nthreads_(nthreads) {
  if (metricsEngine_ && !metricsEngine_->testAvgMetrics().empty() && metricsEngine_->config().numTestUsers == 0) {
    LOG(WARNING) << "computing average test metrics on all users can be slow! "
                    "Set numTestUsers > 0 to sample some of them";
  }
}

WALSEngine::~WALSEngine() {
  if (dev_ != nullptr) qmfb_wals_destroy(dev_);
  if (sharded_ != nullptr) qmfb_wals_sharded_destroy(sharded_);
}

void WALSEngine::init(const std::vector<DatasetElem>& dataset) {
  CHECK(!userFactors_ && !itemFactors_) << "engine was already initialized with train data";
  CHECK(!dataset.empty()) << "empty training dataset";
  // The GPU row solve is a Cholesky factorisation: every row's A = Y^T Y + sum alpha r y y^T + lambda I must be positive
  // definite.  The reference's dsysv (Bunch-Kaufman, qmf/Matrix.cpp:81-96) also solves the indefinite systems that
  // lambda <= 0 on rank-deficient rows or negative confidences (alpha r < 0) can produce; this engine states the
  // restriction BEFORE training instead of failing with QMFB_ERR_NOT_SPD in the middle of it.
  CHECK_GT(config_.regularizationLambda, 0.0)
    << "qmf_b200 requires --regularization_lambda > 0 (per-row Cholesky solve; the reference's dsysv accepts lambda <= 0)";
  {
    bool negative = false;
    for (const auto& e : dataset) negative |= config_.confidenceWeight * e.value < 0.0;
    CHECK(!negative) << "qmf_b200 requires confidence_weight * value >= 0 for every training signal (per-row Cholesky solve; "
                        "the reference's dsysv accepts indefinite rows)";
  }
  // Dense indices + both CSR orientations are built on the GPU (qmfb_signals_*): idx = rank of the
  // raw id among the distinct ids, rows by row id, cells by column id - what IdIndex +
  // groupSignals / sortDataset produce (qmf/wals/WALSEngine.cpp:130-163), without the two host
  // sorts of the whole dataset.
  qmfb_signals_t* signals = nullptr;
  {
    // struct-of-arrays copy of the dataset, first-touched and filled by all host threads
    const size_t n = dataset.size();
    std::unique_ptr<int64_t[]> uid(new int64_t[n]), iid(new int64_t[n]);
    std::unique_ptr<double[]> val(new double[n]);
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), n / 65536 + 1));
    std::vector<std::thread> pool;
    auto fill = [&](size_t t) {
      for (size_t p = n * t / nthreads, e = n * (t + 1) / nthreads; p < e; ++p) {
        uid[p] = dataset[p].userId;
        iid[p] = dataset[p].itemId;
        val[p] = dataset[p].value;
      }
    };
    for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(fill, t);
    fill(0);
    for (auto& th : pool) th.join();
    QMFB_OK_OR_DIE(qmfb_signals_build(config_.device, int64_t(n), uid.get(), iid.get(), val.get(), &signals));
  }
  int64_t nu = 0, ni = 0;
  QMFB_OK_OR_DIE(qmfb_signals_dims(signals, &nu, &ni, nullptr));
  {
    std::vector<int64_t> ids(static_cast<size_t>(std::max(nu, ni)));
    QMFB_OK_OR_DIE(qmfb_signals_ids(signals, QMFB_SIDE_USER, ids.data()));
    for (int64_t r = 0; r < nu; ++r) CHECK_EQ(userIndex_.getOrSetIdx(ids[size_t(r)]), size_t(r));
    QMFB_OK_OR_DIE(qmfb_signals_ids(signals, QMFB_SIDE_ITEM, ids.data()));
    for (int64_t r = 0; r < ni; ++r) CHECK_EQ(itemIndex_.getOrSetIdx(ids[size_t(r)]), size_t(r));
  }

  userFactors_ = std::make_unique<FactorData>(nusers(), config_.nfactors);
  itemFactors_ = std::make_unique<FactorData>(nitems(), config_.nfactors);
  if (config_.DistributionFile.empty()) {
    // user factors need no initial value: the first half-step overwrites them
    std::mt19937 gen(config_.seed >= 0 ? static_cast<uint32_t>(config_.seed) : std::random_device()());
    std::uniform_real_distribution<Double> dist(-config_.initDistributionBound, config_.initDistributionBound);
    itemFactors_->setFactors([&](size_t, size_t) { return dist(gen); });
  } else {
    itemFactors_->setFactors(config_.DistributionFile);
  }

  if (config_.ngpus > 1) {
    // users and items row-partitioned over the GPUs of this box, full factor replicas on each
    // (SURVEY.md 8e); factors and loss are bit-identical to the one-GPU engine
    std::vector<int> devices(size_t(config_.ngpus));
    for (int d = 0; d < config_.ngpus; ++d) devices[size_t(d)] = config_.device + d;
    QMFB_OK_OR_DIE(qmfb_wals_sharded_create(config_.ngpus, devices.data(), int64_t(nusers()), int64_t(nitems()),
                                            int(config_.nfactors), &sharded_));
    QMFB_OK_OR_DIE(qmfb_wals_sharded_set_signals(sharded_, signals));
    qmfb_signals_destroy(signals);
    QMFB_OK_OR_DIE(qmfb_wals_sharded_set_factors(sharded_, QMFB_SIDE_ITEM, itemFactors_->getFactors().data()));
    return;
  }
  QMFB_OK_OR_DIE(qmfb_wals_create(config_.device, int64_t(nusers()), int64_t(nitems()), int(config_.nfactors), &dev_));
  QMFB_OK_OR_DIE(qmfb_wals_set_signals(dev_, signals));
  qmfb_signals_destroy(signals);
  QMFB_OK_OR_DIE(qmfb_wals_set_factors(dev_, QMFB_SIDE_ITEM, itemFactors_->getFactors().data()));
}

void WALSEngine::initTest(const std::vector<DatasetElem>& testDataset) {
  CHECK(test_.empty()) << "engine was already initialized with test data";
  if (metricsEngine_ && !metricsEngine_->testAvgMetrics().empty()) {
    initAvgTestData(test_, testDataset, userIndex_, itemIndex_, metricsEngine_->config().numTestUsers,
                    metricsEngine_->config().seed);
  }
}

Double WALSEngine::iterate(int side) {
  double lossSum = 0.0;
  if (sharded_ != nullptr) {
    QMFB_OK_OR_DIE(qmfb_wals_sharded_half_step(sharded_, side, config_.confidenceWeight, config_.regularizationLambda, &lossSum));
  } else {
    QMFB_OK_OR_DIE(qmfb_wals_half_step(dev_, side, config_.confidenceWeight, config_.regularizationLambda, &lossSum));
  }
  hostStale_ = true;
  return lossSum / nusers() / nitems();
}

void WALSEngine::optimize() {
  CHECK(userFactors_ && itemFactors_) << "no factor data, have you initialized the engine?";
  for (size_t epoch = 1; epoch <= config_.nepochs; ++epoch) {
    iterate(QMFB_SIDE_USER);                      // fix item factors, update user factors
    const Double loss = iterate(QMFB_SIDE_ITEM);  // fix user factors, update item factors
    LOG(INFO) << "epoch " << epoch << ": train loss = " << loss;
    evaluate(epoch);
  }
  syncFactorsToHost();
}

void WALSEngine::syncFactorsToHost() const {
  if (!hostStale_ || (dev_ == nullptr && sharded_ == nullptr)) return;
  if (sharded_ != nullptr) {
    QMFB_OK_OR_DIE(qmfb_wals_sharded_get_factors(sharded_, QMFB_SIDE_USER, 0, userFactors_->getFactors().data()));
    QMFB_OK_OR_DIE(qmfb_wals_sharded_get_factors(sharded_, QMFB_SIDE_ITEM, 0, itemFactors_->getFactors().data()));
  } else {
    QMFB_OK_OR_DIE(qmfb_wals_get_factors(dev_, QMFB_SIDE_USER, userFactors_->getFactors().data()));
    QMFB_OK_OR_DIE(qmfb_wals_get_factors(dev_, QMFB_SIDE_ITEM, itemFactors_->getFactors().data()));
  }
  hostStale_ = false;
}

void WALSEngine::evaluate(const size_t epoch) {
  if (metricsEngine_ && !metricsEngine_->testAvgMetrics().empty() && !test_.empty() &&
      (metricsEngine_->config().alwaysCompute || epoch == config_.nepochs)) {
    LOG(INFO) << "do compute evaluate ...";
    // on the resident factors (no host round trip); with --ngpus the test users are cut over the GPUs
    computeAndRecordTestAvgMetrics(*metricsEngine_, epoch, test_, nitems(), nthreads_,
                                   [this](const int32_t* users, int64_t nT, const int64_t* lp, const int32_t* li, int32_t* cnt, double* ps) {
                                     return sharded_ != nullptr ? qmfb_wals_sharded_eval_rank(sharded_, users, nT, lp, li, cnt, ps)
                                                                : qmfb_wals_eval_rank(dev_, users, nT, lp, li, cnt, ps);
                                   });
  }
}

void WALSEngine::saveUserFactors(const std::string& fileName) const {
  CHECK(userFactors_) << "user factors wasn't initialized";
  syncFactorsToHost();
  saveFactors(*userFactors_, userIndex_, fileName);
}

void WALSEngine::saveItemFactors(const std::string& fileName) const {
  CHECK(itemFactors_) << "item factors wasn't initialized";
  syncFactorsToHost();
  saveFactors(*itemFactors_, itemIndex_, fileName);
}

}  // namespace qmf
