// BPR engine with the reference's public surface (qmf/bpr/BPREngine.h:38-76); the Hogwild SGD
// pass and the evaluation losses run on the GPU through the C ABI.
#pragma once
#include <memory>
#include <random>
#include <string>
#include <unordered_set>
#include <vector>

#include <qmf/Engine.h>

struct qmfb_bpr;

namespace qmf {

struct BPRConfig {
  size_t nepochs;
  size_t nfactors;
  Double initLearningRate;
  Double biasLambda;
  Double userLambda;
  Double itemLambda;
  Double decayRate;
  bool useBiases;
  Double initDistributionBound;
  size_t numNegativeSamples;
  size_t numHogwildThreads;  // kept for the CLI; the GPU pass is always lock-free parallel
  bool shuffleTrainingSet;
  int64_t seed = -1;         // additive: >= 0 seeds the initial factors and the device sampler
  int device = 0;            // additive: CUDA device ordinal
};

class BPREngine : public Engine {
 public:
  explicit BPREngine(const BPRConfig& config, const std::unique_ptr<MetricsEngine>& metricsEngine,
                     const size_t evalNumNeg = 3, const int32_t evalSeed = 42, const size_t nthreads = 16);
  ~BPREngine() override;

  void init(const std::vector<DatasetElem>& dataset) override;
  void initTest(const std::vector<DatasetElem>& testDataset) override;
  void optimize() override;
  void evaluate(const size_t epoch) override;
  void saveUserFactors(const std::string& fileName) const override;
  void saveItemFactors(const std::string& fileName) const override;

  size_t nusers() const { return userIndex_.size(); }
  size_t nitems() const { return itemIndex_.size(); }

  struct Triplets {
    std::vector<int32_t> user, pos, neg;
    size_t size() const { return user.size(); }
  };
  const Triplets& evalSet() const { return evalSet_; }
  const Triplets& testEvalSet() const { return testEvalSet_; }
  const FactorData& userFactors() const { return *userFactors_; }
  const FactorData& itemFactors() const { return *itemFactors_; }
  Double lastTrainLoss() const { return lastTrainLoss_; }
  Double lastTestLoss() const { return lastTestLoss_; }

 private:
  using ItemSets = std::vector<std::unordered_set<size_t>>;
  // rejection sampling with std::uniform_int_distribution<> exactly as
  // BPREngine::sampleRandomNegative (qmf/bpr/BPREngine-inl.h:48-60)
  template <typename Gen>
  size_t sampleNegative(const ItemSets& sets, size_t userIdx, Gen& gen) const;
  Double evalLoss(const Triplets& set) const;
  void syncFactorsToHost() const;

  const BPRConfig& config_;
  const std::unique_ptr<MetricsEngine>& metricsEngine_;
  const size_t evalNumNeg_;
  const int32_t evalSeed_;
  const size_t nthreads_;
  std::mt19937 gen_;
  uint64_t deviceSeed_ = 0;
  Double learningRate_ = 0.0;
  Double lastTrainLoss_ = -1.0, lastTestLoss_ = -1.0;

  IdIndex userIndex_, itemIndex_;
  std::vector<int32_t> dataUser_, dataItem_;  // data_ (BPREngine.cpp:76)
  ItemSets itemMap_, testItemMap_;
  Triplets evalSet_, testEvalSet_;
  std::unique_ptr<FactorData> userFactors_, itemFactors_;  // host mirrors
  mutable bool hostStale_ = false;
  qmfb_bpr* dev_ = nullptr;
  TestData test_;
};

}  // namespace qmf
