// TEST INFRASTRUCTURE — minimal stand-in for <glog/logging.h> so the UNMODIFIED reference
// sources under /root/reference compile in an image that has no glog. Only the macros the
// reference actually uses are provided (LOG, VLOG, CHECK, CHECK_EQ/NE/GT/GE/LT/LE, DCHECK).
// Doubles are streamed with 17 significant digits so the reference's own log lines
// ("epoch N: train loss = ...", qmf/wals/WALSEngine.cpp:92) carry full FP64 precision.
// Not part of the product; used only by oracle/Makefile (the recipe that compiles the unmodified reference into oracle/_ref/).
#pragma once
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
enum Severity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
inline void InitGoogleLogging(const char*) {}
}  // namespace google

extern int FLAGS_logtostderr;
extern int FLAGS_minloglevel;
extern int FLAGS_v;

namespace qmf_shim {
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev) : sev_(sev) {
    static const char tag[] = {'I', 'W', 'E', 'F'};
    os_ << std::setprecision(17);
    const char* base = file;
    for (const char* p = file; *p; ++p) if (*p == '/') base = p + 1;
    os_ << tag[sev] << ' ' << base << ':' << line << "] ";
  }
  ~LogMessage() {
    if (sev_ >= FLAGS_minloglevel || sev_ == google::GLOG_FATAL) {
      os_ << '\n';
      std::cerr << os_.str() << std::flush;
    }
    if (sev_ == google::GLOG_FATAL) std::abort();
  }
  std::ostream& stream() { return os_; }
 private:
  int sev_;
  std::ostringstream os_;
};
struct Voidify { void operator&(std::ostream&) {} };
}  // namespace qmf_shim

#define QMF_SHIM_SEV_INFO google::GLOG_INFO
#define QMF_SHIM_SEV_WARNING google::GLOG_WARNING
#define QMF_SHIM_SEV_ERROR google::GLOG_ERROR
#define QMF_SHIM_SEV_FATAL google::GLOG_FATAL
#define LOG(sev) qmf_shim::LogMessage(__FILE__, __LINE__, QMF_SHIM_SEV_##sev).stream()
#define VLOG(n) (FLAGS_v < (n)) ? (void)0 : qmf_shim::Voidify() & LOG(INFO)
#define LOG_IF(sev, cond) !(cond) ? (void)0 : qmf_shim::Voidify() & LOG(sev)
#define CHECK(cond) (cond) ? (void)0 : qmf_shim::Voidify() & LOG(FATAL) << "Check failed: " #cond " "
#define QMF_SHIM_CHECK_OP(a, b, op) \
  ((a)op(b)) ? (void)0 : qmf_shim::Voidify() & LOG(FATAL) << "Check failed: " #a " " #op " " #b " "
#define CHECK_EQ(a, b) QMF_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) QMF_SHIM_CHECK_OP(a, b, !=)
#define CHECK_GT(a, b) QMF_SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) QMF_SHIM_CHECK_OP(a, b, >=)
#define CHECK_LT(a, b) QMF_SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) QMF_SHIM_CHECK_OP(a, b, <=)
#define DCHECK(cond) CHECK(cond)
