// Engine base class with the reference's virtual surface (qmf/Engine.h:32-64) and the shared
// evaluation / output helpers.  Evaluation data is kept as a CSR of positive test items per test
// user (the reference materialises dense nT x nitems label and score matrices, Engine.cpp:52-69;
// the GPU path never needs them).
#pragma once
#include <functional>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#include <qmf/DatasetReader.h>
#include <qmf/FactorData.h>
#include <qmf/metrics/Metrics.h>
#include <qmf/utils/FriendTest.h>
#include <qmf/utils/IdIndex.h>
#include <qmf/utils/ParallelExecutor.h>

namespace qmf {

class Engine {
 public:
  Engine() = default;
  virtual ~Engine() = default;
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  virtual void init(const std::vector<DatasetElem>& /*dataset*/) {}
  virtual void initTest(const std::vector<DatasetElem>& /*testDataset*/) {}
  virtual void optimize() {}
  virtual void evaluate(const size_t /*epoch*/) {}
  virtual void saveUserFactors(const std::string& /*fileName*/) const {}
  virtual void saveItemFactors(const std::string& /*fileName*/) const {}

  struct TestData {
    std::vector<size_t> users;        // test user idx, in the reference's order
    std::vector<int64_t> labelPtr;    // users.size() + 1
    std::vector<int32_t> labelItems;  // ascending item idx with test label > 0
    bool empty() const { return users.empty(); }
  };

  // Which users are evaluated and in what order: identical to Engine::initAvgTestData
  // (qmf/Engine.cpp:27-71) - users with >= 1 test line whose user AND item are known from
  // training, iterated in std::unordered_set order, optionally std::shuffle'd with mt19937(seed)
  // and truncated to numTestUsers.  A label is the value of the LAST test line of (user, item).
  static void initAvgTestData(TestData& out, const std::vector<DatasetElem>& testDataset, const IdIndex& userIndex,
                              const IdIndex& itemIndex, size_t numTestUsers = 0, int32_t seed = 0);

  // scores of every item for every test user + rank statistics on the GPU, then every requested
  // test-average metric recorded with the reference's averaging order.  `rank` runs the engine's
  // qmfb_*_eval_rank on its resident factors (signature of qmfb_wals_eval_rank without the handle).
  using RankFn = std::function<int(const int32_t* users, int64_t nT, const int64_t* labelPtr, const int32_t* labelItems,
                                   int32_t* cnt, double* posScores)>;
  static void computeAndRecordTestAvgMetrics(MetricsEngine& metrics, size_t epoch, const TestData& test, size_t nItems,
                                             size_t nthreads, const RankFn& rank);

  // The reference's own (protected, static) helpers with their dense nT x nitems vectors
  // (qmf/Engine.h:66-82, Engine.cpp:27-96), kept so that code written against the reference compiles.  Same test
  // users in the same order as the TestData overload; labels hold the value of the LAST test line of a
  // (user, item).  The engines here never call them: evaluation runs on the GPU against resident factors
  // (computeAndRecordTestAvgMetrics).
  static void initAvgTestData(std::vector<size_t>& testUsers, std::vector<std::vector<Double>>& testLabels,
                              std::vector<std::vector<Double>>& testScores, const std::vector<DatasetElem>& testDataset,
                              const IdIndex& userIndex, const IdIndex& itemIndex, const size_t numTestUsers = 0,
                              const int32_t seed = 0);
  static void computeTestScores(std::vector<std::vector<Double>>& testScores, const std::vector<size_t>& testUsers,
                                const FactorData& userFactors, const FactorData& itemFactors, ParallelExecutor& parallel);

  // "<id>[ <bias>] <f0> ... <fk-1>\n", fixed, 9 decimals (qmf/Engine.cpp:98-122)
  static void saveFactors(const FactorData& factorData, const IdIndex& index, const std::string& fileName);
  static void saveFactors(const FactorData& factorData, const IdIndex& index, std::ostream& out);

 private:
  static void selectTestUsers(std::vector<size_t>& users, const std::vector<DatasetElem>& testDataset, const IdIndex& userIndex,
                              const IdIndex& itemIndex, size_t numTestUsers, int32_t seed);
  FRIEND_TEST(Engine, initAvgTestData);
  FRIEND_TEST(Engine, computeTestScores);
  FRIEND_TEST(Engine, saveFactors);
};

}  // namespace qmf
