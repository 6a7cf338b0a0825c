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

void Engine::selectTestUsers(std::vector<size_t>& users, const std::vector<DatasetElem>& testDataset, const IdIndex& userIndex,
                             const IdIndex& itemIndex, size_t numTestUsers, int32_t seed) {
  // Engine.cpp:35-50 of the reference: unordered_set iteration order, then mt19937(seed) shuffle + truncation
  std::unordered_set<size_t> seen;
  for (const auto& e : testDataset) {
    const size_t u = userIndex.idx(e.userId), p = itemIndex.idx(e.itemId);
    if (u != IdIndex::missingIdx && p != IdIndex::missingIdx) seen.insert(u);
  }
  users.assign(seen.begin(), seen.end());
  if (numTestUsers > 0 && numTestUsers < users.size()) {
    std::shuffle(users.begin(), users.end(), std::mt19937(seed));
    users.resize(numTestUsers);
  }
}

void Engine::initAvgTestData(std::vector<size_t>& testUsers, std::vector<std::vector<Double>>& testLabels,
                             std::vector<std::vector<Double>>& testScores, const std::vector<DatasetElem>& testDataset,
                             const IdIndex& userIndex, const IdIndex& itemIndex, const size_t numTestUsers, const int32_t seed) {
  selectTestUsers(testUsers, testDataset, userIndex, itemIndex, numTestUsers, seed);
  std::vector<int64_t> slot(userIndex.size(), -1);
  for (size_t t = 0; t < testUsers.size(); ++t) slot[testUsers[t]] = int64_t(t);
  testLabels.assign(testUsers.size(), std::vector<Double>(itemIndex.size(), 0.0));
  testScores.assign(testUsers.size(), std::vector<Double>(itemIndex.size(), 0.0));
  for (const auto& e : testDataset) {
    const size_t u = userIndex.idx(e.userId), i = itemIndex.idx(e.itemId);
    if (u == IdIndex::missingIdx || i == IdIndex::missingIdx || slot[u] < 0) continue;
    testLabels[size_t(slot[u])][i] = e.value;
  }
}

void Engine::computeTestScores(std::vector<std::vector<Double>>& testScores, const std::vector<size_t>& testUsers,
                               const FactorData& userFactors, const FactorData& itemFactors, ParallelExecutor& parallel) {
  // score = bias + sum_f p_uf q_if, accumulated in factor order (Engine.cpp:85-91)
  parallel.execute(testUsers.size(), [&](const size_t t) {
    const size_t u = testUsers[t];
    std::vector<Double>& row = testScores[t];
    for (size_t i = 0; i < itemFactors.nelems(); ++i) {
      Double acc = itemFactors.withBiases() ? itemFactors.biasAt(i) : 0.0;
      for (size_t f = 0; f < userFactors.nfactors(); ++f) acc += userFactors.at(u, f) * itemFactors.at(i, f);
      row[i] = acc;
    }
  });
}

void Engine::initAvgTestData(TestData& out, const std::vector<DatasetElem>& testDataset, const IdIndex& userIndex,
                             const IdIndex& itemIndex, size_t numTestUsers, int32_t seed) {
  selectTestUsers(out.users, testDataset, userIndex, itemIndex, numTestUsers, seed);
  // slot of a user idx in `users` (dense table instead of the reference's unordered_map, Engine.cpp:53-60)
  std::vector<int64_t> slot(userIndex.size(), -1);
  for (size_t t = 0; t < out.users.size(); ++t) slot[out.users[t]] = int64_t(t);
  // The reference writes testLabels[slot][item] = value line by line into dense nT x nitems vectors
  // (Engine.cpp:62-70): the LAST line of a (user, item) wins, positive <=> value > 0 (Metrics.cpp:72).
  // Same result without dense rows or per-user hash maps: the valid lines as (slot, item, line) keys,
  // sorted; the last line of each (slot, item) run decides.
  struct Cell {
    int64_t slot;
    int32_t item;
    uint32_t hi;   // line number (split so that the struct stays 24 bytes)
    uint32_t lo;
    Double value;
  };
  std::vector<Cell> cells;
  for (size_t p = 0; p < testDataset.size(); ++p) {
    const auto& e = testDataset[p];
    const size_t u = userIndex.idx(e.userId), i = itemIndex.idx(e.itemId);
    if (u == IdIndex::missingIdx || i == IdIndex::missingIdx || slot[u] < 0) continue;
    cells.push_back(Cell{slot[u], int32_t(i), uint32_t(uint64_t(p) >> 32), uint32_t(p), e.value});
  }
  std::sort(cells.begin(), cells.end(), [](const Cell& a, const Cell& b) {
    if (a.slot != b.slot) return a.slot < b.slot;
    if (a.item != b.item) return a.item < b.item;
    return (uint64_t(a.hi) << 32 | a.lo) < (uint64_t(b.hi) << 32 | b.lo);
  });
  out.labelPtr.assign(out.users.size() + 1, 0);
  out.labelItems.clear();
  for (size_t p = 0; p < cells.size(); ++p) {
    const bool last = p + 1 == cells.size() || cells[p + 1].slot != cells[p].slot || cells[p + 1].item != cells[p].item;
    if (last && cells[p].value > 0.0) {
      out.labelItems.push_back(cells[p].item);
      ++out.labelPtr[size_t(cells[p].slot) + 1];
    }
  }
  for (size_t t = 0; t < out.users.size(); ++t) out.labelPtr[t + 1] += out.labelPtr[t];
}

void Engine::computeAndRecordTestAvgMetrics(MetricsEngine& metrics, size_t epoch, const TestData& test, size_t nItems,
                                            size_t nthreads, const RankFn& rank) {
  const size_t nT = test.users.size();
  std::vector<int32_t> users(nT);
  for (size_t t = 0; t < nT; ++t) users[t] = int32_t(test.users[t]);
  std::vector<int32_t> cnt(test.labelItems.size() + nT, 0);
  std::vector<Double> posScores(std::max<size_t>(test.labelItems.size(), 1));
  // all-item scores + rank statistics on the GPU, against the engine's RESIDENT factors
  const int rc = rank(users.data(), int64_t(nT), test.labelPtr.data(), test.labelItems.data(), cnt.data(), posScores.data());
  CHECK_EQ(rc, 0) << "ranking evaluation: " << qmfb_last_error();
  for (const auto& name : metrics.testAvgMetrics()) {
    MetricSpec spec;
    CHECK(MetricsManager::get().lookup(name, spec)) << "missing metric test_avg_" << name;
    std::vector<Double> perUser(nT);
    // per-user values on all host threads (independent users); the average keeps the reference's order
    const size_t nth = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), nT / 512 + 1));
    auto work = [&](size_t th) {
      for (size_t t = nT * th / nth, e = nT * (th + 1) / nth; t < e; ++t) {
        const size_t nPos = size_t(test.labelPtr[t + 1] - test.labelPtr[t]);
        perUser[t] = computeMetricFromCounts(spec, cnt.data() + test.labelPtr[t] + int64_t(t), nPos, nItems);
      }
    };
    std::vector<std::thread> pool;
    for (size_t th = 1; th < nth; ++th) pool.emplace_back(work, th);
    work(0);
    for (auto& th : pool) th.join();
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
