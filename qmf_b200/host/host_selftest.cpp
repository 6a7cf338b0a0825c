// CPU self-test of the host-side mirror (no GPU): the known-answer vectors of the reference's
// gtests for the classes whose surface is mirrored here.  Exit code 0 = all good.
#include <cmath>
#include <cstdio>
#include <memory>
#include <sstream>

#include <cstring>
#include <fstream>
#include <random>
#include <algorithm>
#include <qmf/DatasetReader.h>
#include <qmf/Engine.h>
#include <qmf/FactorData.h>
#include <qmf/Matrix.h>
#include <qmf/metrics/Metrics.h>
#include <qmf/utils/Util.h>

static int failures = 0;
#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) {                                                        \
      std::fprintf(stderr, "FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

using qmf::Double;

static Double metric(const char* name, std::vector<Double> labels, std::vector<Double> scores) {
  qmf::MetricSpec spec;
  EXPECT(qmf::MetricsManager::get().lookup(name, spec));
  return qmf::computeMetric(spec, labels, scores);
}

int main() {
  // qmf/test/MetricsTest.cpp:35-88
  EXPECT(metric("mse", {1, 0}, {0.5, 0.5}) == 0.25);
  EXPECT(metric("auc", {1, 0}, {3, 2}) == 1.0);
  EXPECT(metric("auc", {0, 1}, {3, 2}) == 0.0);
  EXPECT(metric("auc", {1, 1, 0}, {3, 2, 0}) == 1.0);
  EXPECT(metric("auc", {1, 0, 1}, {3, 2, 0}) == 0.5);
  EXPECT(metric("auc", {0, 1, 1}, {3, 2, 0}) == 0.0);
  EXPECT(metric("p@1", {1, 1}, {3, 2}) == 1.0);
  EXPECT(metric("p@2", {0, 1, 0}, {3, 2, 1}) == 0.5);
  EXPECT(metric("p@2", {0, 1, 0}, {3, 1, 2}) == 0.0);
  EXPECT(metric("r@1", {1, 1}, {3, 2}) == 0.5);
  EXPECT(metric("r@2", {0, 1, 0}, {3, 2, 1}) == 1.0);
  EXPECT(metric("r@2", {0, 1, 0}, {3, 1, 2}) == 0.0);
  EXPECT(metric("ap", {0, 1}, {3, 2}) == 0.5);
  EXPECT(metric("ap", {0, 1, 0}, {3, 1, 2}) == 1.0 / 3);
  // qmf/test/MetricsManagerTest.cpp
  std::string base;
  size_t k = 0;
  EXPECT(qmf::detail::parseAtKMetric("p@5", base, k) && base == "p" && k == 5);
  EXPECT(!qmf::detail::parseAtKMetric("auc", base, k) && !qmf::detail::parseAtKMetric("@5", base, k));
  EXPECT(qmf::MetricsManager::get().exists("r@10") && !qmf::MetricsManager::get().exists("foo") &&
         !qmf::MetricsManager::get().exists("x@3"));
  // qmf/test/EngineTest.cpp:113-139 (exact output text)
  {
    qmf::IdIndex index;
    index.getOrSetIdx(3);
    index.getOrSetIdx(5);
    qmf::FactorData f(2, 3);
    f.setFactors([](size_t i, size_t j) { return Double(i * 3 + j); });
    std::ostringstream out;
    qmf::Engine::saveFactors(f, index, out);
    EXPECT(out.str() == "3 0.000000000 1.000000000 2.000000000\n5 3.000000000 4.000000000 5.000000000\n");
    qmf::FactorData fb(2, 3, true);
    fb.setFactors([](size_t i, size_t j) { return Double(i * 3 + j); });
    fb.setBiases([](size_t i) { return Double(5 + i); });
    std::ostringstream outb;
    qmf::Engine::saveFactors(fb, index, outb);
    EXPECT(outb.str() == "3 5.000000000 0.000000000 1.000000000 2.000000000\n5 6.000000000 3.000000000 4.000000000 5.000000000\n");
  }
  // qmf/test/EngineTest.cpp:23-73 (which users / labels are evaluated)
  {
    qmf::IdIndex users, items;
    for (int64_t u : {1, 2, 3}) users.getOrSetIdx(u);
    for (int64_t i : {10, 20, 30, 40}) items.getOrSetIdx(i);
    std::vector<qmf::DatasetElem> test = {{1, 20, 1.0}, {1, 99, 1.0}, {7, 10, 1.0}, {3, 40, 2.0}, {3, 10, 0.0}, {3, 40, 3.0}};
    qmf::Engine::TestData td;
    qmf::Engine::initAvgTestData(td, test, users, items);
    EXPECT(td.users.size() == 2 && td.labelPtr.size() == 3);
    for (size_t t = 0; t < td.users.size(); ++t) {
      const size_t n = size_t(td.labelPtr[t + 1] - td.labelPtr[t]);
      if (td.users[t] == 0) EXPECT(n == 1 && td.labelItems[size_t(td.labelPtr[t])] == 1);
      if (td.users[t] == 2) EXPECT(n == 1 && td.labelItems[size_t(td.labelPtr[t])] == 3);
    }
  }
  // qmf/test/MatrixTest.cpp:92-116: symmetric INDEFINITE 50 x 50, residual <= 1e-8
  {
    const size_t n = 50;
    qmf::Matrix A(n, n);
    qmf::Vector b(n);
    unsigned s = 12345;
    auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return (s >> 8) / double(1 << 24) - 0.5; };
    for (size_t i = 0; i < n; ++i) {
      b(i) = rnd();
      for (size_t j = i; j < n; ++j) A(i, j) = A(j, i) = rnd();
    }
    qmf::Vector x = qmf::linearSymmetricSolve(A, b);
    double worst = 0.0;
    for (size_t i = 0; i < n; ++i) {
      double r = -b(i);
      for (size_t j = 0; j < n; ++j) r += A(i, j) * x(j);
      worst = std::fmax(worst, std::fabs(r));
    }
    EXPECT(worst <= 1e-8);
    qmf::Matrix T = A.transpose();
    EXPECT(T(3, 7) == A(7, 3) && (A + T)(2, 9) == 2 * A(2, 9));
  }
  // qmf/test/DatasetReaderTest.cpp:25-58, UtilTest.cpp
  {
    auto in = std::make_unique<std::istringstream>("1 2 3\n-4 5 0.5\n");
    qmf::DatasetReader reader(std::move(in));
    const auto d = reader.readAll();
    EXPECT(d.size() == 2 && d[0].userId == 1 && d[0].itemId == 2 && d[0].value == 3.0 && d[1].userId == -4 && d[1].value == 0.5);
    const auto parts = qmf::split("auc,,p@10,", ',');
    EXPECT(parts.size() == 2 && parts[0] == "auc" && parts[1] == "p@10");
  }
  // the block-parallel saveFactors writes the bytes of the sequential "id [bias] f0 f1 ...\n" writer
  {
    const size_t n = 70001, k = 5;
    qmf::FactorData f(n, k, true);
    qmf::IdIndex index;
    std::mt19937_64 gen(3);
    std::uniform_real_distribution<double> dist(-3.0, 3.0);
    f.setFactors([&](size_t, size_t) { return dist(gen); });
    std::ostringstream want;
    for (size_t r = 0; r < n; ++r) {
      index.getOrSetIdx(int64_t(r) * 3 - 100);
      f.biasAt(r) = dist(gen);
    }
    for (size_t r = 0; r < n; ++r) {
      char num[64];
      want << (int64_t(r) * 3 - 100);
      std::snprintf(num, sizeof num, " %.9f", f.biasAt(r));
      want << num;
      for (size_t c = 0; c < k; ++c) {
        std::snprintf(num, sizeof num, " %.9f", f.at(r, c));
        want << num;
      }
      want << '\n';
    }
    std::ostringstream got;
    qmf::Engine::saveFactors(f, index, got);
    EXPECT(got.str() == want.str());
  }
  // the mapped multi-threaded readAll() returns exactly what the getline + sscanf loop returns
  {
    const std::string path = "/tmp/qmf_b200_selftest_dataset.txt";
    std::mt19937_64 gen(7);
    const char* weights[] = {"1", "0.5", "3.25", "1e-3", "2.5E+2", "-0", "007.1250", ".5", "5.", "0.1", "0.30000000000000004",
                             "123456789012345678", "1.7976931348623157e308", "4.9e-324", "inf", "nan", "0x1.8p1", "1e",
                             "17.000000000000000000001", "0.000000000000000000000012345", "9007199254740993", "3.14abc"};
    {
      std::ofstream out(path);
      for (int n = 0; n < 300000; ++n) {
        const long long u = (long long)(gen() % 2000003) - 1000, i = (long long)(gen() % 50021);
        const char* sep1 = (n % 7 == 0) ? "\t" : " ";
        const char* sep2 = (n % 11 == 0) ? "   " : " ";
        if (n % 13 == 0) out << "  ";
        if (n % 17 == 0 && u >= 0) out << '+';
        out << u << sep1 << i << sep2 << weights[gen() % (sizeof(weights) / sizeof(weights[0]))];
        if (n % 19 == 0) out << " trailing";
        if (n % 23 == 0) out << '\r';
        out << '\n';
      }
      out << "9223372036854775808 -9223372036854775809 1\n";  // clamped like strtoll
      out << "5 6 7";                                           // last line without a newline
    }
    std::vector<qmf::DatasetElem> slow, fast;
    {
      qmf::DatasetReader reader(path);
      qmf::DatasetElem e;
      while (reader.readOne(e)) slow.push_back(e);
    }
    {
      qmf::DatasetReader reader(path);
      reader.readAll(fast);
    }
    EXPECT(slow.size() == 300002 && fast.size() == slow.size());
    size_t diff = 0;
    for (size_t n = 0; n < std::min(slow.size(), fast.size()); ++n) {
      diff += slow[n].userId != fast[n].userId || slow[n].itemId != fast[n].itemId ||
              std::memcmp(&slow[n].value, &fast[n].value, sizeof(double)) != 0;
    }
    EXPECT(diff == 0);
    EXPECT(fast[300000].userId == INT64_MAX && fast[300000].itemId == INT64_MIN);
    qmf::DatasetElem e;
    EXPECT(!qmf::DatasetReader::parseLine("1 2", "1 2" + 3, e));
    EXPECT(!qmf::DatasetReader::parseLine("", "", e));
    EXPECT(!qmf::DatasetReader::parseLine("1 x 3", "1 x 3" + 5, e));
    EXPECT(!qmf::DatasetReader::parseLine("1 2 .", "1 2 ." + 5, e));
    {  // embedded NUL bytes end the line for sscanf (it sees the C string): same verdicts and values
#define QMF_LIT(x) std::string(x, sizeof(x) - 1)
      const std::string cases[] = {QMF_LIT("1 2 3\0junk"), QMF_LIT("1 2\0 3"), QMF_LIT("1 2 3.\0005"), QMF_LIT("1\0 2 3"),
                                   QMF_LIT("7 8 9e\0002")};
#undef QMF_LIT
      for (const auto& c : cases) {
        long long u = 0, i = 0;
        double w = 0.0;
        const bool want = std::sscanf(c.c_str(), "%lld %lld %lf", &u, &i, &w) == 3;
        qmf::DatasetElem g;
        const bool got = qmf::DatasetReader::parseLine(c.data(), c.data() + c.size(), g);
        EXPECT(want == got);
        if (want && got) EXPECT(g.userId == u && g.itemId == i && std::memcmp(&g.value, &w, sizeof w) == 0);
      }
    }
    std::remove(path.c_str());
  }
  // FactorData::setFactors(file): the mapped multi-threaded reader == the reference's getline + sscanf("%lf") loop
  // (qmf/FactorData.h:74-100), bit for bit, on a 3 MB file of mixed number formats; extra lines are ignored, a short file
  // sets what it has and leaves the rest untouched
  {
    const std::string path = "/tmp/qmf_b200_selftest_dist.txt";
    const size_t n = 1500, k = 96;  // 144 000 values
    {
      std::mt19937_64 g(11);
      std::uniform_real_distribution<double> U(-0.01, 0.01);
      std::ofstream out(path);
      char buf[128];
      for (size_t i = 0; i < n * k + 37; ++i) {   // 37 extra lines
        const double v = U(g);
        switch (i % 9) {
          case 0: std::snprintf(buf, sizeof buf, "%.9f\n", v); break;                 // gen_uniform's format
          case 1: std::snprintf(buf, sizeof buf, "  %.17g\r\n", v); break;             // leading blanks, CRLF, 17 digits
          case 2: std::snprintf(buf, sizeof buf, "%.3e trailing text\n", v); break;     // exponent + ignored tail
          case 3: std::snprintf(buf, sizeof buf, "\t%d\n", int(i % 1000) - 500); break;  // integers
          case 4: std::snprintf(buf, sizeof buf, "%.12f 7.5\n", v); break;             // only the first field counts
          case 5: std::snprintf(buf, sizeof buf, "+%.6f\n", std::fabs(v)); break;
          case 6: std::snprintf(buf, sizeof buf, "0x1.8p-%d\n", int(i % 20)); break;    // hex float
          case 7: std::snprintf(buf, sizeof buf, ".%05d\n", int(i % 100000)); break;
          default: std::snprintf(buf, sizeof buf, "%.20f\n", v); break;               // long mantissa
        }
        out << buf;
      }
    }
    qmf::FactorData a(n, k), b(n, k);
    a.setFactorsSequential(path);
    b.setFactors(path);
    EXPECT(std::memcmp(a.getFactors().data(), b.getFactors().data(), n * k * sizeof(double)) == 0);
    // short file: 3 rows more than the file holds
    qmf::FactorData c(n + 3, k), d(n + 3, k);
    c.setFactors([](size_t, size_t) { return 7.0; });
    d.setFactors([](size_t, size_t) { return 7.0; });
    c.setFactorsSequential(path);
    d.setFactors(path);
    EXPECT(std::memcmp(c.getFactors().data(), d.getFactors().data(), (n + 3) * k * sizeof(double)) == 0);
    EXPECT(d.at(n + 2, k - 1) == 7.0 && d.at(0, 0) == a.at(0, 0));
    std::remove(path.c_str());
  }
  // averaging order of Metric::compute(labels, scores, parallel)
  EXPECT(qmf::averageOverUsers({1.0, 2.0, 3.0, 4.0, 5.0}, 2) == ((1.0 + 3.0 + 5.0) + (2.0 + 4.0)) / 5);
  if (failures == 0) std::printf("host selftest ok\n");
  return failures == 0 ? 0 : 1;
}
