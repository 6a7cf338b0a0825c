// Minimal logging/CHECK facility with glog's call syntax (the reference logs through glog; this
// image has none).  Observable format kept: "epoch N: train loss = ..." lines go to stderr with
// the ostream default precision, CHECK failures abort like glog's LOG(FATAL).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace qmf {
namespace logging {

enum Level { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };

inline int& minLevel() {
  static int level = INFO;
  return level;
}

class Line {
 public:
  Line(Level lvl, const char* file, int line) : lvl_(lvl) {
    static const char tags[] = "IWEF";
    const char* base = file;
    for (const char* p = file; *p != '\0'; ++p) {
      if (*p == '/') base = p + 1;
    }
    buf_ << tags[lvl] << ' ' << base << ':' << line << "] ";
    // additive: QMF_LOG_PRECISION=17 prints doubles with full precision (default: ostream's 6 digits)
    static const int precision = std::getenv("QMF_LOG_PRECISION") ? std::atoi(std::getenv("QMF_LOG_PRECISION")) : 0;
    if (precision > 0) buf_.precision(precision);
  }
  ~Line() {
    if (lvl_ >= minLevel() || lvl_ == FATAL) {
      buf_ << '\n';
      std::cerr << buf_.str() << std::flush;
    }
    if (lvl_ == FATAL) std::abort();
  }
  std::ostream& out() { return buf_; }

 private:
  Level lvl_;
  std::ostringstream buf_;
};

struct Sink {
  void operator&(std::ostream&) const {}
};

}  // namespace logging
}  // namespace qmf

#define LOG(level) ::qmf::logging::Line(::qmf::logging::level, __FILE__, __LINE__).out()
#define CHECK(cond) (cond) ? (void)0 : ::qmf::logging::Sink() & LOG(FATAL) << "Check failed: " #cond " "
#define QMF_CHECK_BINARY(a, b, op) ((a)op(b)) ? (void)0 : ::qmf::logging::Sink() & LOG(FATAL) << "Check failed: " #a " " #op " " #b " "
#define CHECK_EQ(a, b) QMF_CHECK_BINARY(a, b, ==)
#define CHECK_GT(a, b) QMF_CHECK_BINARY(a, b, >)
#define CHECK_GE(a, b) QMF_CHECK_BINARY(a, b, >=)
#define CHECK_LT(a, b) QMF_CHECK_BINARY(a, b, <)
