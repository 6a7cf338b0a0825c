#include <qmf/bpr/BPREngine.h>

#include <utility>

#include "qmf_b200.h"

namespace qmf {

#define QMFB_OK_OR_DIE(call)                                        \
  do {                                                              \
    const int qmfb_rc_ = (call);                                    \
    CHECK_EQ(qmfb_rc_, 0) << #call << ": " << qmfb_last_error();    \
  } while (0)

BPREngine::BPREngine(const BPRConfig& config, const std::unique_ptr<MetricsEngine>& metricsEngine, const size_t evalNumNeg,
                     const int32_t evalSeed, const size_t nthreads)
  : config_(config),
    metricsEngine_(metricsEngine),
    evalNumNeg_(evalNumNeg),
    evalSeed_(evalSeed),
    nthreads_(nthreads),
    gen_(config.seed >= 0 ? static_cast<uint32_t>(config.seed) : std::random_device()()) {
  deviceSeed_ = (uint64_t(gen_()) << 32) | gen_();
  if (config_.numHogwildThreads > nthreads) {
    LOG(WARNING) << "number of hogwild threads should be smaller than number of threads in the threadpool";
  }
  if (metricsEngine_ && !metricsEngine_->testAvgMetrics().empty() && metricsEngine_->config().numTestUsers == 0) {
    LOG(WARNING) << "computing average test metrics on all users can be slow! "
                    "Set numTestUsers > 0 to sample some of them";
  }
}

BPREngine::~BPREngine() {
  if (dev_ != nullptr) qmfb_bpr_destroy(dev_);
}

template <typename Gen>
size_t BPREngine::sampleNegative(const ItemSets& sets, size_t userIdx, Gen& gen) const {
  const auto& positives = sets.at(userIdx);
  std::uniform_int_distribution<> pick(0, static_cast<int>(nitems()) - 1);
  size_t j;
  do {
    j = size_t(pick(gen));
  } while (positives.count(j) > 0);
  return j;
}

void BPREngine::init(const std::vector<DatasetElem>& dataset) {
  CHECK(!userFactors_ && !itemFactors_) << "engine was already initialized with train data";
  // dense idx in first-appearance order; lines with value < 1 are not positives (BPREngine.cpp:69-77)
  for (const auto& e : dataset) {
    if (e.value < 1.0) continue;
    dataUser_.push_back(int32_t(userIndex_.getOrSetIdx(e.userId)));
    dataItem_.push_back(int32_t(itemIndex_.getOrSetIdx(e.itemId)));
  }
  CHECK(!dataUser_.empty()) << "no positive pairs in the training dataset";
  itemMap_.resize(nusers());
  for (size_t p = 0; p < dataUser_.size(); ++p) itemMap_[size_t(dataUser_[p])].insert(size_t(dataItem_[p]));

  // fixed evaluation triplets: evalNumNeg negatives per pair from mt19937(evalSeed) (:85-87)
  std::mt19937 evalGen(evalSeed_);
  for (size_t p = 0; p < dataUser_.size(); ++p) {
    for (size_t n = 0; n < evalNumNeg_; ++n) {
      evalSet_.user.push_back(dataUser_[p]);
      evalSet_.pos.push_back(dataItem_[p]);
      evalSet_.neg.push_back(int32_t(sampleNegative(itemMap_, size_t(dataUser_[p]), evalGen)));
    }
  }

  learningRate_ = config_.initLearningRate;
  userFactors_ = std::make_unique<FactorData>(nusers(), config_.nfactors);
  itemFactors_ = std::make_unique<FactorData>(nitems(), config_.nfactors, config_.useBiases);
  std::uniform_real_distribution<Double> dist(-config_.initDistributionBound, config_.initDistributionBound);
  auto draw = [&](auto...) { return dist(gen_); };
  userFactors_->setFactors(draw);
  itemFactors_->setFactors(draw);
  if (config_.useBiases) itemFactors_->setBiases(draw);

  QMFB_OK_OR_DIE(qmfb_bpr_create(config_.device, int64_t(nusers()), int64_t(nitems()), int(config_.nfactors),
                                 config_.useBiases ? 1 : 0, &dev_));
  QMFB_OK_OR_DIE(qmfb_bpr_set_data(dev_, dataUser_.data(), dataItem_.data(), int64_t(dataUser_.size())));
  QMFB_OK_OR_DIE(qmfb_bpr_set_factors(dev_, QMFB_SIDE_USER, userFactors_->getFactors().data()));
  QMFB_OK_OR_DIE(qmfb_bpr_set_factors(dev_, QMFB_SIDE_ITEM, itemFactors_->getFactors().data()));
  if (config_.useBiases) QMFB_OK_OR_DIE(qmfb_bpr_set_biases(dev_, itemFactors_->getBiases().data()));
  // Hogwild block split: the tail of ndata mod numHogwildThreads pairs is never visited (BPREngine.cpp:156-160)
  QMFB_OK_OR_DIE(qmfb_bpr_set_hogwild_blocks(dev_, int64_t(config_.numHogwildThreads)));
}

void BPREngine::initTest(const std::vector<DatasetElem>& testDataset) {
  CHECK(testEvalSet_.size() == 0) << "engine was already initialzied with test data";
  std::vector<std::pair<size_t, size_t>> valid;
  testItemMap_.resize(nusers());
  for (const auto& e : testDataset) {
    if (e.value < 1.0) continue;
    const size_t u = userIndex_.idx(e.userId), p = itemIndex_.idx(e.itemId);
    if (u == IdIndex::missingIdx || p == IdIndex::missingIdx) continue;
    testItemMap_[u].insert(p);
    valid.emplace_back(u, p);
  }
  // negatives of the test evaluation set avoid only the TEST positives (BPREngine.cpp:127-136)
  std::mt19937 evalGen(evalSeed_);
  for (const auto& up : valid) {
    for (size_t n = 0; n < evalNumNeg_; ++n) {
      testEvalSet_.user.push_back(int32_t(up.first));
      testEvalSet_.pos.push_back(int32_t(up.second));
      testEvalSet_.neg.push_back(int32_t(sampleNegative(testItemMap_, up.first, evalGen)));
    }
  }
  if (metricsEngine_ && !metricsEngine_->testAvgMetrics().empty()) {
    initAvgTestData(test_, testDataset, userIndex_, itemIndex_, metricsEngine_->config().numTestUsers,
                    metricsEngine_->config().seed);
  }
}

void BPREngine::optimize() {
  CHECK(userFactors_ && itemFactors_) << "no factor data, have you initialized the engine?";
  for (size_t epoch = 1; epoch <= config_.nepochs; ++epoch) {
    int64_t nUpdates = 0;
    // the reference shuffles AFTER evaluate() (BPREngine.cpp:172-174): epoch 1 runs in file order
    QMFB_OK_OR_DIE(qmfb_bpr_epoch(dev_, learningRate_, config_.userLambda, config_.itemLambda, config_.biasLambda,
                                  int(config_.numNegativeSamples), deviceSeed_, uint64_t(epoch),
                                  (config_.shuffleTrainingSet && epoch > 1) ? 1 : 0, &nUpdates));
    hostStale_ = true;
    evaluate(epoch);
    if (config_.decayRate < 1.0) learningRate_ *= config_.decayRate;
  }
  syncFactorsToHost();
}

Double BPREngine::evalLoss(const Triplets& set) const {
  if (set.size() == 0) return -1.0;
  // the reference sums nthreads blocks of floor(n / nthreads) triplets and divides by n
  // (ParallelExecutor::mapReduce(elems), qmf/utils/ParallelExecutor-inl.h:60-85)
  const size_t used = (set.size() / nthreads_) * nthreads_;
  double sum = 0.0;
  QMFB_OK_OR_DIE(qmfb_bpr_eval_loss(dev_, set.user.data(), set.pos.data(), set.neg.data(), int64_t(used), &sum));
  return sum / set.size();
}

void BPREngine::evaluate(const size_t epoch) {
  lastTrainLoss_ = evalLoss(evalSet_);
  lastTestLoss_ = evalLoss(testEvalSet_);
  LOG(INFO) << "epoch " << epoch << ": train loss = " << lastTrainLoss_ << ", test loss = " << lastTestLoss_;
  if (metricsEngine_ && !metricsEngine_->testAvgMetrics().empty() && !test_.empty() &&
      (metricsEngine_->config().alwaysCompute || epoch == config_.nepochs)) {
    computeAndRecordTestAvgMetrics(*metricsEngine_, epoch, test_, nitems(), nthreads_,
                                   [this](const int32_t* users, int64_t nT, const int64_t* lp, const int32_t* li, int32_t* cnt, double* ps) {
                                     return qmfb_bpr_eval_rank(dev_, users, nT, lp, li, cnt, ps);
                                   });
  }
}

void BPREngine::syncFactorsToHost() const {
  if (!hostStale_ || dev_ == nullptr) return;
  QMFB_OK_OR_DIE(qmfb_bpr_get_factors(dev_, QMFB_SIDE_USER, userFactors_->getFactors().data()));
  QMFB_OK_OR_DIE(qmfb_bpr_get_factors(dev_, QMFB_SIDE_ITEM, itemFactors_->getFactors().data()));
  if (config_.useBiases) QMFB_OK_OR_DIE(qmfb_bpr_get_biases(dev_, itemFactors_->getBiases().data()));
  hostStale_ = false;
}

void BPREngine::saveUserFactors(const std::string& fileName) const {
  CHECK(userFactors_) << "user factors wasn't initialized";
  syncFactorsToHost();
  saveFactors(*userFactors_, userIndex_, fileName);
}

void BPREngine::saveItemFactors(const std::string& fileName) const {
  CHECK(itemFactors_) << "item factors wasn't initialized";
  syncFactorsToHost();
  saveFactors(*itemFactors_, itemIndex_, fileName);
}

}  // namespace qmf
