// Per-user-averaged ranking metrics.  The heavy part (scoring all items and ranking) runs on the
// GPU (qmfb_eval_rank); this header holds the registry / recorder surface of the reference
// (qmf/metrics/MetricsEngine.h:29-135, MetricsManager.h) and the arithmetic that turns the GPU's
// integer rank statistics into the metric values with the reference's own formulas
// (qmf/metrics/Metrics.cpp:65-164).
#pragma once
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include <qmf/Types.h>

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

// name -> metric; knows mse, auc, ap and p@k / r@k for any k
class MetricsManager {
 public:
  static const MetricsManager& get();
  bool exists(const std::string& name) const;
  bool lookup(const std::string& name, MetricSpec& spec) const;
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

  using MetricVector = std::vector<std::pair<size_t, Double>>;
  // stores the value and logs "epoch N: recorded metric <key> = v" (MetricsEngine.cpp:36-44)
  void recordMetric(const std::string& key, size_t epoch, Double value);
  const MetricVector* recorded(const std::string& key) const;

 private:
  bool add(std::vector<std::string>& list, const std::string& m);

  const MetricsConfig config_;
  const bool log_;
  std::vector<std::string> trainMetrics_, trainAvgMetrics_, testMetrics_, testAvgMetrics_;
  std::unordered_map<std::string, MetricVector> metricsMap_;
};

}  // namespace qmf
