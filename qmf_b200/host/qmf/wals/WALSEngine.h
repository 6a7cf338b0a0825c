// Weighted ALS engine with the reference's public surface (qmf/wals/WALSEngine.h:35-60); the
// half-step (Gram, per-row normal equations, solve, loss) runs on the GPU through the C ABI.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include <qmf/Engine.h>

struct qmfb_wals;
struct qmfb_wals_sharded;

namespace qmf {

struct WALSConfig {
  size_t nepochs;
  size_t nfactors;
  Double regularizationLambda;
  Double confidenceWeight;
  Double initDistributionBound;
  std::string DistributionFile;
  int64_t seed = -1;   // additive: >= 0 seeds the initial item factors (the reference uses random_device)
  int device = 0;      // additive: CUDA device ordinal (the first one when ngpus > 1)
  int ngpus = 1;       // additive: row-partition the half-steps over devices device .. device+ngpus-1 of this box
};

class WALSEngine : public Engine {
 public:
  explicit WALSEngine(const WALSConfig& config, const std::unique_ptr<MetricsEngine>& metricsEngine,
                      const size_t nthreads = 16);
  ~WALSEngine() override;

  void init(const std::vector<DatasetElem>& dataset) override;
  void initTest(const std::vector<DatasetElem>& testDataset) override;
  void optimize() override;
  void evaluate(const size_t epoch) override;
  void saveUserFactors(const std::string& fileName) const override;
  void saveItemFactors(const std::string& fileName) const override;

  size_t nusers() const { return userIndex_.size(); }
  size_t nitems() const { return itemIndex_.size(); }

  const FactorData& userFactors() const { return *userFactors_; }
  const FactorData& itemFactors() const { return *itemFactors_; }

 private:
  // one half-step on the device; returns the loss divided by nusers and nitems (WALSEngine.cpp:215)
  Double iterate(int side);
  void syncFactorsToHost() const;

  const WALSConfig& config_;
  const std::unique_ptr<MetricsEngine>& metricsEngine_;
  const size_t nthreads_;

  IdIndex userIndex_, itemIndex_;
  std::unique_ptr<FactorData> userFactors_, itemFactors_;  // host mirrors
  mutable bool hostStale_ = false;
  qmfb_wals* dev_ = nullptr;                // ngpus == 1
  qmfb_wals_sharded* sharded_ = nullptr;    // ngpus > 1: one process, all GPUs (qmfb_wals_sharded_*)
  TestData test_;
};

}  // namespace qmf
