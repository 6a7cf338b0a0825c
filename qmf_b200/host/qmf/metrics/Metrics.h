// Per-user-averaged ranking metrics.  The heavy part (scoring all items and ranking) runs on the
// GPU (qmfb_eval_rank); this header holds the registry / recorder surface of the reference
// (qmf/metrics/MetricsEngine.h:29-135, MetricsManager.h) and the arithmetic that turns the GPU's
// integer rank statistics into the metric values with the reference's own formulas
// (qmf/metrics/Metrics.cpp:65-164).
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include <qmf/Types.h>
#include <qmf/utils/Log.h>
#include <qmf/utils/ParallelExecutor.h>

namespace qmf {

struct MetricsConfig {
  size_t numTestUsers;
  bool alwaysCompute;
  int32_t seed;
};

enum class MetricKind { kMse, kAuc, kAp, kPrecision, kRecall };

struct MetricSpec {
  MetricKind kind;
  size_t k;  // for p@k / r@k
};

namespace detail {
// "p@5" -> ("p", 5); false if the name has no '@' or no number after it
bool parseAtKMetric(const std::string& name, std::string& metricName, size_t& k);
}

// The reference's metric objects (qmf/metrics/Metrics.h:26-93): compute(labels, scores) for one user,
// compute(vector<labels>, vector<scores>[, parallel]) = mean over users (serial / the executor's strided
// partial sums, Metrics.cpp:27-52).  All of them evaluate through computeMetric below, i.e. the same
// count-based arithmetic the GPU path feeds (computeMetricFromCounts).
class Metric {
 public:
  explicit Metric(MetricSpec spec) : spec_(spec) {}
  virtual ~Metric() = default;
  virtual Double compute(const std::vector<Double>& labels, const std::vector<Double>& scores) const;
  virtual Double compute(const std::vector<std::vector<Double>>& labels, const std::vector<std::vector<Double>>& scores) const;
  virtual Double compute(const std::vector<std::vector<Double>>& labels, const std::vector<std::vector<Double>>& scores,
                         ParallelExecutor& parallel) const;
  const MetricSpec& spec() const { return spec_; }

 private:
  const MetricSpec spec_;
};
class MeanSquaredError : public Metric {
 public:
  MeanSquaredError() : Metric({MetricKind::kMse, 0}) {}
};
class AUC : public Metric {
 public:
  AUC() : Metric({MetricKind::kAuc, 0}) {}
};
class AveragePrecision : public Metric {
 public:
  AveragePrecision() : Metric({MetricKind::kAp, 0}) {}
};
class Precision : public Metric {
 public:
  explicit Precision(const size_t k) : Metric({MetricKind::kPrecision, k}) {}
};
class Recall : public Metric {
 public:
  explicit Recall(const size_t k) : Metric({MetricKind::kRecall, k}) {}
};

// name -> metric; knows mse, auc, ap and p@k / r@k for any k (qmf/metrics/MetricsManager.h:37-70)
class MetricsManager {
 public:
  MetricsManager() = default;
  MetricsManager(const MetricsManager&) = delete;
  MetricsManager& operator=(const MetricsManager&) = delete;

  static const MetricsManager& get();
  void init() {}  // the fixed names need no registration; "p@k" / "r@k" are created on first use (initFromName)
  bool initFromName(const std::string& name) const;
  template <typename MetricT, typename... Args>
  void registerMetric(const std::string& name, Args&&... args) const {
    std::lock_guard<std::mutex> lock(mu_);
    metrics_.emplace(name, std::make_unique<MetricT>(std::forward<Args>(args)...));
  }
  // null pointer if the name is unknown
  const std::unique_ptr<Metric>& getMetric(const std::string& name) const;
  bool exists(const std::string& name) const;
  bool lookup(const std::string& name, MetricSpec& spec) const;

 private:
  mutable std::mutex mu_;
  mutable std::unordered_map<std::string, std::unique_ptr<Metric>> metrics_;
};

// one user, dense vectors: the reference's Metric::compute(labels, scores)
Double computeMetric(const MetricSpec& spec, const std::vector<Double>& labels, const std::vector<Double>& scores);
// one user, from the GPU's bucket counts cnt[0..nP] (see include/qmf_b200.h, qmfb_eval_rank)
Double computeMetricFromCounts(const MetricSpec& spec, const int32_t* cnt, size_t nPos, size_t nItems);
// mean over users with the summation order of Metric::compute(labels, scores, parallel):
// strided per-thread partial sums folded in thread order (nthreads == 0: plain serial sum)
Double averageOverUsers(const std::vector<Double>& perUser, size_t nthreads);

class MetricsEngine {
 public:
  explicit MetricsEngine(const MetricsConfig& config = MetricsConfig{0, false, 0}, bool log = true) : config_(config), log_(log) {}

  const MetricsConfig& config() const { return config_; }
  bool addTrainMetric(const std::string& m) { return add(trainMetrics_, m); }
  bool addTestMetric(const std::string& m) { return add(testMetrics_, m); }
  bool addTrainAvgMetric(const std::string& m) { return add(trainAvgMetrics_, m); }
  bool addTestAvgMetric(const std::string& m) { return add(testAvgMetrics_, m); }
  const std::vector<std::string>& trainMetrics() const { return trainMetrics_; }
  const std::vector<std::string>& testMetrics() const { return testMetrics_; }
  const std::vector<std::string>& trainAvgMetrics() const { return trainAvgMetrics_; }
  const std::vector<std::string>& testAvgMetrics() const { return testAvgMetrics_; }

  // compute every registered metric on (labels, scores[, parallel]) and record it under its prefix
  // (qmf/metrics/MetricsEngine.h:59-122)
  void computeAndRecordTrainMetrics(const size_t epoch, const std::vector<Double>& labels, const std::vector<Double>& scores) {
    computeAndRecordMetrics(trainMetrics_, "train_", epoch, labels, scores);
  }
  void computeAndRecordTestMetrics(const size_t epoch, const std::vector<Double>& labels, const std::vector<Double>& scores) {
    computeAndRecordMetrics(testMetrics_, "test_", epoch, labels, scores);
  }
  template <typename... ComputeArgs>
  void computeAndRecordTrainAvgMetrics(const size_t epoch, ComputeArgs&... args) {
    computeAndRecordMetrics(trainAvgMetrics_, "train_avg_", epoch, args...);
  }
  template <typename... ComputeArgs>
  void computeAndRecordTestAvgMetrics(const size_t epoch, ComputeArgs&... args) {
    computeAndRecordMetrics(testAvgMetrics_, "test_avg_", epoch, args...);
  }

  using MetricVector = std::vector<std::pair<size_t, Double>>;
  // stores the value and logs "epoch N: recorded metric <key> = v" (MetricsEngine.cpp:36-44)
  void recordMetric(const std::string& key, size_t epoch, Double value);
  const MetricVector* recorded(const std::string& key) const;

 private:
  bool add(std::vector<std::string>& list, const std::string& m);
  template <typename... ComputeArgs>
  void computeAndRecordMetrics(const std::vector<std::string>& names, const std::string& prefix, const size_t epoch,
                               ComputeArgs&... args) {
    for (const auto& name : names) {
      const auto& m = MetricsManager::get().getMetric(name);
      CHECK(m) << "missing metric " << prefix + name;
      recordMetric(prefix + name, epoch, m->compute(args...));
    }
  }

  const MetricsConfig config_;
  const bool log_;
  std::vector<std::string> trainMetrics_, trainAvgMetrics_, testMetrics_, testAvgMetrics_;
  std::unordered_map<std::string, MetricVector> metricsMap_;
};

}  // namespace qmf
