// FactorData::setFactors(file): the --distribution_file reader of the reference (qmf/FactorData.h:74-100), parallel.
#include <qmf/FactorData.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include <qmf/DatasetReader.h>
#include <qmf/utils/Log.h>

namespace qmf {

void FactorData::setFactorsSequential(const std::string& fileName) {
  std::ifstream in(fileName);
  std::string line;
  size_t count = 0;
  for (size_t i = 0; i < nelems(); ++i) {
    for (size_t f = 0; f < nfactors(); ++f) {
      if (!std::getline(in, line)) {
        LOG(ERROR) << "read uniform data from " << fileName << " failed.";
        return;
      }
      double v = 0.0;
      CHECK_EQ(std::sscanf(line.c_str(), "%lf", &v), 1) << "the file format is incorrect: " << line;
      factors_(i, f) = v;
      ++count;
    }
  }
  LOG(INFO) << "initialized factor from file size: " << count;
}

void FactorData::setFactors(const std::string& fileName) {
  const size_t want = nelems() * nfactors();
  const int fd = ::open(fileName.c_str(), O_RDONLY);
  struct stat st;
  if (fd < 0 || ::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size == 0 || want == 0) {
    if (fd >= 0) ::close(fd);
    setFactorsSequential(fileName);  // streams, missing or empty files: the reference's own loop and messages
    return;
  }
  const size_t size = size_t(st.st_size);
  void* map = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (map == MAP_FAILED) {
    setFactorsSequential(fileName);
    return;
  }
  ::madvise(map, size, MADV_SEQUENTIAL);
  const char* base = static_cast<const char*>(map);
  const size_t nthreads = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), size / (1 << 20) + 1));
  // chunk t covers the lines that START in [cut[t], cut[t+1])
  std::vector<size_t> cut(nthreads + 1, size);
  cut[0] = 0;
  for (size_t t = 1; t < nthreads; ++t) {
    size_t pos = std::max(size / nthreads * t, cut[t - 1]);
    const void* nl = pos < size ? std::memchr(base + pos, '\n', size - pos) : nullptr;
    cut[t] = nl ? size_t(static_cast<const char*>(nl) - base) + 1 : size;
  }
  auto for_each_line = [&](size_t t, auto&& fn) {  // fn(begin, end) -> false stops
    size_t pos = cut[t];
    while (pos < cut[t + 1]) {
      const void* nl = std::memchr(base + pos, '\n', cut[t + 1] - pos);
      const size_t end = nl ? size_t(static_cast<const char*>(nl) - base) : cut[t + 1];
      if (!fn(pos, end)) return;
      pos = end + 1;
    }
  };
  std::vector<size_t> nlines(nthreads, 0);
  std::vector<std::thread> pool;
  auto count = [&](size_t t) {
    size_t n = 0;
    for_each_line(t, [&](size_t, size_t) { ++n; return true; });
    nlines[t] = n;
  };
  for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(count, t);
  count(0);
  for (auto& th : pool) th.join();
  pool.clear();
  std::vector<size_t> first(nthreads + 1, 0);  // index of the first line of chunk t
  for (size_t t = 0; t < nthreads; ++t) first[t + 1] = first[t] + nlines[t];
  const size_t nread = std::min(first[nthreads], want);
  double* out = factors_.data();
  std::vector<size_t> bad(nthreads, SIZE_MAX);  // offset of the first malformed line (among the lines read) per chunk
  auto work = [&](size_t t) {
    size_t idx = first[t];
    for_each_line(t, [&](size_t b, size_t e) {
      if (idx >= nread) return false;
      // getline + c_str(): the line ends at the newline; an embedded NUL ends what sscanf sees, and the scanner stops there too
      if (!DatasetReader::parseDoubleField(base + b, base + e, out[idx])) {
        bad[t] = b;
        return false;
      }
      ++idx;
      return true;
    });
  };
  for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (size_t t = 0; t < nthreads; ++t) {
    if (bad[t] != SIZE_MAX) {  // the first malformed line in file order, as the sequential reader reports it
      const void* nl = std::memchr(base + bad[t], '\n', size - bad[t]);
      const size_t end = nl ? size_t(static_cast<const char*>(nl) - base) : size;
      const std::string line(base + bad[t], end - bad[t]);
      ::munmap(map, size);
      CHECK_EQ(0, 1) << "the file format is incorrect: " << line;
    }
  }
  ::munmap(map, size);
  if (nread < want) {
    LOG(ERROR) << "read uniform data from " << fileName << " failed.";
    return;
  }
  LOG(INFO) << "initialized factor from file size: " << nread;
}

}  // namespace qmf
