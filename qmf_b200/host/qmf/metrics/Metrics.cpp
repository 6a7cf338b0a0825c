#include <qmf/metrics/Metrics.h>

#include <algorithm>
#include <cmath>
#include <functional>

#include <qmf/utils/Log.h>

#include "qmf_b200.h"

namespace qmf {

namespace detail {
bool parseAtKMetric(const std::string& name, std::string& metricName, size_t& k) {
  const size_t at = name.find('@');
  if (at == std::string::npos || at == 0 || at + 1 >= name.size()) return false;
  size_t value = 0;
  for (size_t p = at + 1; p < name.size(); ++p) {
    if (name[p] < '0' || name[p] > '9') return false;
    value = value * 10 + size_t(name[p] - '0');
  }
  metricName = name.substr(0, at);
  k = value;
  return true;
}
}  // namespace detail

const MetricsManager& MetricsManager::get() {
  static const MetricsManager instance;
  return instance;
}

bool MetricsManager::lookup(const std::string& name, MetricSpec& spec) const {
  if (name == "mse") { spec = {MetricKind::kMse, 0}; return true; }
  if (name == "auc") { spec = {MetricKind::kAuc, 0}; return true; }
  if (name == "ap") { spec = {MetricKind::kAp, 0}; return true; }
  std::string base;
  size_t k = 0;
  if (!detail::parseAtKMetric(name, base, k)) return false;
  if (base == "p") { spec = {MetricKind::kPrecision, k}; return true; }
  if (base == "r") { spec = {MetricKind::kRecall, k}; return true; }
  return false;
}

bool MetricsManager::exists(const std::string& name) const {
  MetricSpec s;
  return lookup(name, s);
}

bool MetricsManager::initFromName(const std::string& name) const {
  MetricSpec spec;
  if (!lookup(name, spec)) return false;
  std::lock_guard<std::mutex> lock(mu_);
  if (metrics_.find(name) == metrics_.end()) metrics_.emplace(name, std::make_unique<Metric>(spec));
  return true;
}

const std::unique_ptr<Metric>& MetricsManager::getMetric(const std::string& name) const {
  static const std::unique_ptr<Metric> none;
  if (!initFromName(name)) return none;
  std::lock_guard<std::mutex> lock(mu_);
  return metrics_.find(name)->second;  // node-based map: the reference stays valid across later insertions
}

Double Metric::compute(const std::vector<Double>& labels, const std::vector<Double>& scores) const {
  return computeMetric(spec_, labels, scores);
}

Double Metric::compute(const std::vector<std::vector<Double>>& labels, const std::vector<std::vector<Double>>& scores) const {
  CHECK_EQ(labels.size(), scores.size());
  std::vector<Double> perUser(labels.size());
  for (size_t t = 0; t < labels.size(); ++t) perUser[t] = compute(labels[t], scores[t]);
  return averageOverUsers(perUser, 0);
}

Double Metric::compute(const std::vector<std::vector<Double>>& labels, const std::vector<std::vector<Double>>& scores,
                       ParallelExecutor& parallel) const {
  CHECK_EQ(labels.size(), scores.size());
  std::vector<Double> perUser(labels.size());
  parallel.execute(labels.size(), [&](const size_t t) { perUser[t] = compute(labels[t], scores[t]); });
  return averageOverUsers(perUser, parallel.nthreads());
}

namespace {
// position (0-based) in the reference's ranking (score descending, positives first on ties) of
// the q-th positive in ascending-score order: negatives that outscore it + positives above it
void positivePositions(const int32_t* cnt, size_t nPos, std::vector<size_t>& pos) {
  pos.resize(nPos);
  size_t greater = 0;
  for (size_t q = nPos; q-- > 0;) {
    greater += size_t(cnt[q + 1]);
    pos[q] = greater + (nPos - 1 - q);
  }
}
}  // namespace

Double computeMetricFromCounts(const MetricSpec& spec, const int32_t* cnt, size_t nPos, size_t nItems) {
  const size_t nNeg = nItems - nPos;
  if (spec.kind == MetricKind::kAuc) {
    if (nPos == 0 || nNeg == 0) {
      LOG(ERROR) << "AUC needs at least 1 example in each class";
      return 1.0;
    }
    // one addition of tp / pos / neg per negative, in rank order (buckets nP .. 0)
    const int32_t p = int32_t(nPos), n = int32_t(nNeg);
    Double auc = 0;
    for (size_t i = nPos + 1; i-- > 0;) {
      const Double term = static_cast<Double>(int(nPos - i)) / p / n;
      // cnt[i] additions of the same term, with the exact result of the one-by-one loop (qmfb_repeated_add)
      auc = qmfb_repeated_add(auc, term, cnt[i]);
    }
    return auc;
  }
  CHECK(spec.kind != MetricKind::kMse) << "mse is not a ranking metric";
  std::vector<size_t> pos;
  positivePositions(cnt, nPos, pos);
  if (spec.kind == MetricKind::kAp) {
    CHECK_GT(nPos, 0u) << "AP needs at least 1 positive";
    Double ap = 0.0;
    int32_t seen = 0;
    for (size_t q = nPos; q-- > 0;) {
      ++seen;
      ap += static_cast<Double>(seen) / (pos[q] + 1);
    }
    return ap / int32_t(nPos);
  }
  CHECK_GE(nItems, spec.k) << "P@k / R@k need at least k ranked elements";
  long hits = 0;
  for (size_t q = 0; q < nPos; ++q) hits += pos[q] < spec.k ? 1 : 0;
  if (spec.kind == MetricKind::kPrecision) return static_cast<Double>(hits) / spec.k;
  CHECK_GT(nPos, 0u) << "R@k needs at least 1 positive";
  return static_cast<Double>(hits) / int32_t(nPos);
}

Double computeMetric(const MetricSpec& spec, const std::vector<Double>& labels, const std::vector<Double>& scores) {
  CHECK_EQ(labels.size(), scores.size());
  if (spec.kind == MetricKind::kMse) {
    CHECK_GT(labels.size(), 0u);
    Double sum = 0.0;
    for (size_t i = 0; i < labels.size(); ++i) sum += std::pow(labels[i] - scores[i], 2);
    return sum / labels.size();
  }
  // bucket the negatives against the ascending positives, then share the count-based formulas
  std::vector<Double> posScores;
  for (size_t i = 0; i < labels.size(); ++i) {
    if (labels[i] > 0.0) posScores.push_back(scores[i]);
  }
  std::sort(posScores.begin(), posScores.end());
  std::vector<int32_t> cnt(posScores.size() + 1, 0);
  for (size_t i = 0; i < labels.size(); ++i) {
    if (!(labels[i] > 0.0)) {
      ++cnt[size_t(std::lower_bound(posScores.begin(), posScores.end(), scores[i]) - posScores.begin())];
    }
  }
  return computeMetricFromCounts(spec, cnt.data(), posScores.size(), labels.size());
}

Double averageOverUsers(const std::vector<Double>& perUser, size_t nthreads) {
  CHECK_GT(perUser.size(), 0u);
  Double total = 0.0;
  if (nthreads == 0) {
    for (const Double v : perUser) total += v;
    return total / perUser.size();
  }
  for (size_t th = 0; th < nthreads; ++th) {
    Double part = 0.0;
    for (size_t t = th; t < perUser.size(); t += nthreads) part = part + perUser[t];
    total = total + part;
  }
  return total / perUser.size();
}

bool MetricsEngine::add(std::vector<std::string>& list, const std::string& m) {
  if (!MetricsManager::get().exists(m)) return false;
  list.push_back(m);
  return true;
}

void MetricsEngine::recordMetric(const std::string& key, size_t epoch, Double value) {
  metricsMap_[key].emplace_back(epoch, value);
  if (log_) LOG(INFO) << "epoch " << epoch << ": recorded metric " << key << " = " << value;
}

const MetricsEngine::MetricVector* MetricsEngine::recorded(const std::string& key) const {
  const auto it = metricsMap_.find(key);
  return it == metricsMap_.end() ? nullptr : &it->second;
}

}  // namespace qmf
