#include <qmf/Engine.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <functional>
#include <random>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "qmf_b200.h"

namespace qmf {

void Engine::initAvgTestData(TestData& out, const std::vector<DatasetElem>& testDataset, const IdIndex& userIndex,
                             const IdIndex& itemIndex, size_t numTestUsers, int32_t seed) {
  std::unordered_set<size_t> seen;
  for (const auto& e : testDataset) {
    const size_t u = userIndex.idx(e.userId), p = itemIndex.idx(e.itemId);
    if (u != IdIndex::missingIdx && p != IdIndex::missingIdx) seen.insert(u);
  }
  out.users.assign(seen.begin(), seen.end());
  if (numTestUsers > 0 && numTestUsers < out.users.size()) {
    std::shuffle(out.users.begin(), out.users.end(), std::mt19937(seed));
    out.users.resize(numTestUsers);
  }
  std::unordered_map<size_t, size_t> slot;
  for (size_t t = 0; t < out.users.size(); ++t) slot[out.users[t]] = t;
  // last value wins per (user, item); positive <=> value > 0 (Metrics.cpp:72)
  std::vector<std::unordered_map<int32_t, Double>> labels(out.users.size());
  for (const auto& e : testDataset) {
    const size_t u = userIndex.idx(e.userId), p = itemIndex.idx(e.itemId);
    if (u == IdIndex::missingIdx || p == IdIndex::missingIdx) continue;
    const auto it = slot.find(u);
    if (it != slot.end()) labels[it->second][int32_t(p)] = e.value;
  }
  out.labelPtr.assign(1, 0);
  out.labelItems.clear();
  for (const auto& row : labels) {
    const size_t begin = out.labelItems.size();
    for (const auto& kv : row) {
      if (kv.second > 0.0) out.labelItems.push_back(kv.first);
    }
    std::sort(out.labelItems.begin() + begin, out.labelItems.end());
    out.labelPtr.push_back(int64_t(out.labelItems.size()));
  }
}

void Engine::computeAndRecordTestAvgMetrics(MetricsEngine& metrics, size_t epoch, const TestData& test,
                                            const FactorData& userFactors, const FactorData& itemFactors,
                                            size_t nthreads, int device) {
  const size_t nT = test.users.size(), nItems = itemFactors.nelems();
  std::vector<int32_t> users(nT);
  for (size_t t = 0; t < nT; ++t) users[t] = int32_t(test.users[t]);
  std::vector<int32_t> cnt(test.labelItems.size() + nT, 0);
  std::vector<Double> posScores(std::max<size_t>(test.labelItems.size(), 1));
  const int rc = qmfb_eval_rank(device, userFactors.getFactors().data(), int64_t(userFactors.nelems()),
                                itemFactors.getFactors().data(), int64_t(nItems), int(userFactors.nfactors()),
                                itemFactors.withBiases() ? itemFactors.getBiases().data() : nullptr, users.data(),
                                int64_t(nT), test.labelPtr.data(), test.labelItems.data(), cnt.data(), posScores.data());
  CHECK_EQ(rc, 0) << "qmfb_eval_rank: " << qmfb_last_error();
  for (const auto& name : metrics.testAvgMetrics()) {
    MetricSpec spec;
    CHECK(MetricsManager::get().lookup(name, spec)) << "missing metric test_avg_" << name;
    std::vector<Double> perUser(nT);
    for (size_t t = 0; t < nT; ++t) {
      const size_t nPos = size_t(test.labelPtr[t + 1] - test.labelPtr[t]);
      perUser[t] = computeMetricFromCounts(spec, cnt.data() + test.labelPtr[t] + int64_t(t), nPos, nItems);
    }
    metrics.recordMetric("test_avg_" + name, epoch, averageOverUsers(perUser, nthreads));
  }
}

void Engine::saveFactors(const FactorData& factorData, const IdIndex& index, const std::string& fileName) {
  std::ofstream out(fileName);
  saveFactors(factorData, index, out);
}

void Engine::saveFactors(const FactorData& factorData, const IdIndex& index, std::ostream& out) {
  CHECK_EQ(factorData.nelems(), index.size());
  // printf("%.9f") rounds exactly like std::fixed << std::setprecision(9) (Engine.cpp:98-122 of the
  // reference).  Rows are formatted by all host threads, a block of rows per thread and wave, and the
  // blocks are written in order: the bytes are those of the sequential writer
  const size_t n = factorData.nelems();
  const size_t nthreads = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), n / 256 + 1));
  const size_t block = 4096;
  auto formatRows = [&](size_t begin, size_t end, std::string& buf) {
    char num[64];
    buf.clear();
    for (size_t idx = begin; idx < end; ++idx) {
      buf += std::to_string(index.id(idx));
      if (factorData.withBiases()) {
        buf.append(num, size_t(std::snprintf(num, sizeof num, " %.9f", factorData.biasAt(idx))));
      }
      for (size_t f = 0; f < factorData.nfactors(); ++f) {
        buf.append(num, size_t(std::snprintf(num, sizeof num, " %.9f", factorData.at(idx, f))));
      }
      buf += '\n';
    }
  };
  std::vector<std::string> bufs(nthreads);
  for (size_t wave = 0; wave < n; wave += nthreads * block) {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nthreads; ++t) {
      const size_t b = std::min(n, wave + t * block), e = std::min(n, b + block);
      if (t == 0 || b >= e) continue;
      pool.emplace_back(formatRows, b, e, std::ref(bufs[t]));
    }
    formatRows(wave, std::min(n, wave + block), bufs[0]);
    for (auto& th : pool) th.join();
    for (size_t t = 0; t < nthreads; ++t) {
      if (wave + t * block < n) out.write(bufs[t].data(), std::streamsize(bufs[t].size()));
    }
  }
}

}  // namespace qmf
