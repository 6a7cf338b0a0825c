#include <qmf/DatasetReader.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>

#include <qmf/utils/Log.h>

namespace qmf {

DatasetReader::DatasetReader(const std::string& fileName) : stream_(new std::ifstream(fileName)), fileName_(fileName) {}

bool DatasetReader::readOne(DatasetElem& elem) {
  CHECK(stream_ != nullptr);
  touched_ = true;
  if (!std::getline(*stream_, line_)) return false;
  long long u = 0, i = 0;
  double w = 0.0;
  // same conversion as the reference (sscanf "%lld %lld %lf", DatasetReader.cpp:37-41): a line
  // that does not carry all three fields is fatal
  const int got = std::sscanf(line_.c_str(), "%lld %lld %lf", &u, &i, &w);
  CHECK_EQ(got, 3) << "the file format is incorrect: " << line_;
  elem.userId = u;
  elem.itemId = i;
  elem.value = w;
  return true;
}

namespace {

// the characters scanf's white-space directive and strtoll/strtod skip in the "C" locale
inline bool isSpace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// %lld: optional sign, at least one digit, clamped like strtoll
inline bool parseInt64(const char*& p, const char* e, int64_t& out) {
  while (p < e && isSpace(*p)) ++p;
  bool neg = false;
  if (p < e && (*p == '-' || *p == '+')) {
    neg = *p == '-';
    ++p;
  }
  if (p >= e || unsigned(*p - '0') > 9u) return false;
  uint64_t v = 0;
  int nd = 0;
  for (; p < e && unsigned(*p - '0') <= 9u && nd < 18; ++p, ++nd) v = v * 10 + uint64_t(*p - '0');  // 18 digits cannot overflow
  if (p < e && unsigned(*p - '0') <= 9u) {  // longer: continue with overflow checks
    bool overflow = false;
    for (; p < e && unsigned(*p - '0') <= 9u; ++p) {
      const uint64_t d = uint64_t(*p - '0');
      if (v > (UINT64_MAX - d) / 10) overflow = true;
      if (!overflow) v = v * 10 + d;
    }
    const uint64_t lim = neg ? uint64_t(LLONG_MAX) + 1 : uint64_t(LLONG_MAX);
    if (overflow || v > lim) v = lim;
  }
  out = neg ? int64_t(0 - v) : int64_t(v);
  return true;
}

const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                         1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// %lf.  Plain decimals with <= 15 significant digits are converted exactly (integer mantissa and
// power of ten both representable: one correctly rounded division or product, Clinger's fast path);
// everything else (exponents, inf/nan, hex floats, long mantissas) goes through strtod on a
// NUL-terminated copy, which is what sscanf itself uses.
inline bool parseDouble(const char*& p, const char* e, double& out) {
  while (p < e && isSpace(*p)) ++p;
  const char* s = p;
  bool neg = false;
  if (s < e && (*s == '-' || *s == '+')) {
    neg = *s == '-';
    ++s;
  }
  uint64_t mant = 0;
  int digits = 0, frac = 0;
  const char* q = s;
  for (; q < e && *q >= '0' && *q <= '9'; ++q) {
    mant = mant * 10 + uint64_t(*q - '0');
    digits += (mant != 0);
    if (digits > 15) break;
  }
  if (digits <= 15 && q < e && *q == '.') {
    ++q;
    for (; q < e && *q >= '0' && *q <= '9'; ++q) {
      mant = mant * 10 + uint64_t(*q - '0');
      digits += (mant != 0);
      ++frac;
      if (digits > 15 || frac > 22) break;
    }
  }
  const bool simple = digits <= 15 && frac <= 22 && q > s && !(q == s + 1 && s[0] == '.') &&
                      (q >= e || !(*q == 'e' || *q == 'E' || *q == 'x' || *q == 'X' || *q == 'p' || *q == 'P' || *q == '.' ||
                                   (*q >= '0' && *q <= '9')));
  if (simple) {
    const double v = frac ? double(mant) / kPow10[frac] : double(mant);
    out = neg ? -v : v;
    p = q;
    return true;
  }
  char small[128];
  std::string big;
  const size_t n = size_t(e - p);
  char* buf = small;
  if (n >= sizeof(small)) {
    big.assign(p, n);
    buf = &big[0];
  } else {
    std::memcpy(small, p, n);
    small[n] = 0;
  }
  char* end = nullptr;
  const double v = std::strtod(buf, &end);
  if (end == buf) return false;
  out = v;
  p += end - buf;
  return true;
}

}  // namespace

bool DatasetReader::parseLine(const char* b, const char* e, DatasetElem& elem) {
  // sscanf works on the C string: an embedded NUL ends the line.  No pre-scan is needed for that: a NUL
  // is neither white space nor part of a number, so every scanner below stops at it exactly like sscanf,
  // and the strtod fallback works on a NUL-terminated copy.
  const char* p = b;
  int64_t u, i;
  double w;
  if (!parseInt64(p, e, u) || !parseInt64(p, e, i) || !parseDouble(p, e, w)) return false;
  elem.userId = u;
  elem.itemId = i;
  elem.value = w;
  return true;
}

bool DatasetReader::parseDoubleField(const char* b, const char* e, double& out) {
  const char* p = b;
  return parseDouble(p, e, out);
}

bool DatasetReader::readAllMapped(std::vector<DatasetElem>& dataset) {
  if (fileName_.empty() || touched_) return false;
  const int fd = ::open(fileName_.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
    ::close(fd);
    return false;
  }
  const size_t size = size_t(st.st_size);
  if (size == 0) {
    ::close(fd);
    dataset.clear();
    touched_ = true;
    return true;
  }
  void* map = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (map == MAP_FAILED) return false;
  ::madvise(map, size, MADV_SEQUENTIAL);
  const char* base = static_cast<const char*>(map);

  const size_t nthreads = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), size / (1 << 20) + 1));
  // chunk t covers the lines that START in [cut[t], cut[t+1])
  std::vector<size_t> cut(nthreads + 1, size);
  cut[0] = 0;
  for (size_t t = 1; t < nthreads; ++t) {
    size_t pos = size / nthreads * t;
    if (pos < cut[t - 1]) pos = cut[t - 1];
    const void* nl = pos < size ? std::memchr(base + pos, '\n', size - pos) : nullptr;
    cut[t] = nl ? size_t(static_cast<const char*>(nl) - base) + 1 : size;
  }
  // pass 1: lines per chunk (so that every thread parses straight into its slice of the result)
  std::vector<size_t> nlines(nthreads, 0);
  auto count = [&](size_t t) {
    size_t n = 0, pos = cut[t];
    while (pos < cut[t + 1]) {
      const void* nl = std::memchr(base + pos, '\n', cut[t + 1] - pos);
      ++n;
      if (!nl) break;  // last line of the file without a newline
      pos = size_t(static_cast<const char*>(nl) - base) + 1;
    }
    nlines[t] = n;
  };
  std::vector<std::thread> pool;
  for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(count, t);
  count(0);
  for (auto& th : pool) th.join();
  pool.clear();
  size_t total = 0;
  std::vector<size_t> off(nthreads);
  for (size_t t = 0; t < nthreads; ++t) {
    off[t] = total;
    total += nlines[t];
  }
  dataset.resize(total);
  // pass 2: parse
  std::vector<size_t> bad(nthreads, SIZE_MAX);  // offset of the first malformed line of each chunk
  auto work = [&](size_t t) {
    DatasetElem* out = dataset.data() + off[t];
    size_t pos = cut[t];
    while (pos < cut[t + 1]) {
      const void* nl = std::memchr(base + pos, '\n', cut[t + 1] - pos);
      const size_t end = nl ? size_t(static_cast<const char*>(nl) - base) : cut[t + 1];
      if (!parseLine(base + pos, base + end, *out++)) {
        bad[t] = pos;
        return;
      }
      pos = end + 1;
    }
  };
  for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (size_t t = 0; t < nthreads; ++t) {
    if (bad[t] != SIZE_MAX) {  // the first malformed line in file order, as the sequential reader would report it
      const void* nl = std::memchr(base + bad[t], '\n', size - bad[t]);
      const size_t end = nl ? size_t(static_cast<const char*>(nl) - base) : size;
      const std::string line(base + bad[t], end - bad[t]);
      ::munmap(map, size);
      CHECK_EQ(0, 3) << "the file format is incorrect: " << line;
    }
  }
  ::munmap(map, size);
  touched_ = true;
  stream_->setstate(std::ios::eofbit | std::ios::failbit);  // a later readOne() sees the end of the data
  return true;
}

void DatasetReader::readAll(std::vector<DatasetElem>& dataset) {
  if (readAllMapped(dataset)) return;
  dataset.clear();
  DatasetElem e;
  while (readOne(e)) dataset.push_back(e);
}

std::vector<DatasetElem> DatasetReader::readAll() {
  std::vector<DatasetElem> d;
  readAll(d);
  return d;
}

}  // namespace qmf
