// C-ABI (include/qmf_b200.h), HOST side of the ranking evaluation: the reference's metric arithmetic
// (qmf/metrics/Metrics.cpp:65-164) on the integer bucket counts the GPU produces.  No device code here.
//
// The reference accumulates AUC with ONE floating-point addition per negative, in rank order
// (Metrics.cpp:87-95: auc += (double)tp / pos / neg); every negative of one bucket adds the same
// term.  Replaying that literally costs O(nitems) host additions per test user (10^13 for 10 M users x
// 1 M items).  qmfb_repeated_add performs `count` additions of the same term with the exact
// round-to-nearest-even result of the sequential loop, in O(binades crossed) integer steps.
#include "qmfb_common.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

inline uint64_t bits_of(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return u;
}

// s <- fl(s + t), `count` times, s >= 0, t > 0, both finite and normal (or s == 0)
double repeated_add(double s, double t, int64_t count) {
  if (count <= 0) return s;
  if (!(t > 0.0) || !(s >= 0.0) || !std::isfinite(s) || !std::isfinite(t) || bits_of(t) < (uint64_t(1) << 52)) {
    for (int64_t c = 0; c < count; ++c) s += t;  // outside the fast path's preconditions: literal loop
    return s;
  }
  const uint64_t tb = bits_of(t);
  const int et = int(tb >> 52) - 1075;                       // t = mt * 2^et, mt a 53-bit integer
  const uint64_t mt = (tb & ((uint64_t(1) << 52) - 1)) | (uint64_t(1) << 52);
  while (count > 0) {
    const uint64_t sb = bits_of(s);
    if (sb < (uint64_t(1) << 52)) {  // zero or subnormal accumulator: one literal step
      s += t;
      --count;
      continue;
    }
    const int es = int(sb >> 52) - 1075;                     // s = S * 2^es, S in [2^52, 2^53), ulp(s) = 2^es
    const int shift = es - et;
    if (shift <= 0) {  // t is a multiple of ulp(s) and at least as large as s's binade: literal step (s is still small)
      s += t;
      --count;
      continue;
    }
    uint64_t S = (sb & ((uint64_t(1) << 52) - 1)) | (uint64_t(1) << 52);
    uint64_t q, D;
    bool tie = false;
    if (shift > 53) {
      return s;  // t < ulp(s) / 2: every remaining addition is absorbed
    } else {
      q = mt >> shift;                                         // t = (q + r / 2^shift) ulp
      const uint64_t r = mt & ((uint64_t(1) << shift) - 1), half = uint64_t(1) << (shift - 1);
      tie = r == half;
      D = q + (r > half ? 1 : 0);
    }
    if (tie) {
      // round-half-to-even: after one literal step S is even; from then on the increment is the even one of q, q + 1
      if (S & 1) {
        s += t;
        --count;
        continue;
      }
      D = q + (q & 1);
    }
    if (D == 0) return s;  // absorbed
    const uint64_t L = uint64_t(1) << 53;                      // the binade ends at S == 2^53 (exactly representable)
    const uint64_t n = (L - S) / D;                            // steps that stay at or below the end of the binade
    if (n == 0) {
      s += t;  // the step that crosses into the next binade: literal
      --count;
      continue;
    }
    const uint64_t m = std::min<uint64_t>(n, uint64_t(count));
    S += m * D;
    count -= int64_t(m);
    s = std::ldexp(double(S), es);                             // S <= 2^53: exact
  }
  return s;
}

bool parse_metric(const char* name, int* kind, int64_t* at_k) {
  const std::string n(name ? name : "");
  *at_k = 0;
  if (n == "auc") { *kind = 1; return true; }
  if (n == "ap") { *kind = 2; return true; }
  const size_t at = n.find('@');
  if (at == std::string::npos || at == 0 || at + 1 >= n.size()) return false;
  int64_t k = 0;
  for (size_t p = at + 1; p < n.size(); ++p) {
    if (n[p] < '0' || n[p] > '9') return false;
    k = k * 10 + (n[p] - '0');
  }
  *at_k = k;
  if (n.substr(0, at) == "p") { *kind = 3; return true; }
  if (n.substr(0, at) == "r") { *kind = 4; return true; }
  return false;
}

// one user; returns false where the reference CHECK-fails (Metrics.cpp:104,136,144)
bool metric_from_counts(int kind, int64_t at_k, const int32_t* cnt, int64_t nPos, int64_t nItems, double* out) {
  const int64_t nNeg = nItems - nPos;
  if (kind == 1) {
    if (nPos == 0 || nNeg == 0) {  // LOG(ERROR) + return 1.0 in the reference (Metrics.cpp:80-83)
      *out = 1.0;
      return true;
    }
    const int32_t p = int32_t(nPos), n = int32_t(nNeg);
    double auc = 0;
    for (int64_t i = nPos; i >= 0; --i) {  // rank order: bucket nP first (above every positive)
      const double term = static_cast<double>(int(nPos - i)) / p / n;
      auc = repeated_add(auc, term, cnt[i]);
    }
    *out = auc;
    return true;
  }
  // position (0-based, reference order) of the q-th positive in ascending-score order
  if (kind == 2) {
    if (nPos <= 0) return false;
    double ap = 0.0;
    int64_t greater = 0;
    int32_t seen = 0;
    for (int64_t q = nPos - 1; q >= 0; --q) {
      greater += cnt[q + 1];
      const int64_t pos = greater + (nPos - 1 - q);
      ++seen;
      ap += static_cast<double>(seen) / double(pos + 1);
    }
    *out = ap / int32_t(nPos);
    return true;
  }
  if (nItems < at_k) return false;
  int64_t hits = 0, greater = 0;
  for (int64_t q = nPos - 1; q >= 0; --q) {
    greater += cnt[q + 1];
    hits += (greater + (nPos - 1 - q)) < at_k ? 1 : 0;
  }
  if (kind == 3) {
    *out = static_cast<double>(hits) / double(at_k);
    return true;
  }
  if (nPos <= 0) return false;
  *out = static_cast<double>(hits) / int32_t(nPos);
  return true;
}

}  // namespace

extern "C" {

double qmfb_repeated_add(double s, double t, int64_t count) { return repeated_add(s, t, count); }

int qmfb_rank_metrics(const char* metric, const int32_t* cnt, const int64_t* label_ptr, int64_t nT, int64_t nitems, int host_threads,
                      double* per_user) {
  int kind = 0;
  int64_t at_k = 0;
  if (!parse_metric(metric, &kind, &at_k)) return qmfb::set_error(QMFB_ERR_INVALID, "qmfb_rank_metrics: unknown ranking metric '%s'", metric ? metric : "");
  if (!cnt || !label_ptr || !per_user || nT < 0 || nitems < 1) return qmfb::set_error(QMFB_ERR_INVALID, "qmfb_rank_metrics: bad argument");
  const int nth = int(std::max<int64_t>(1, std::min<int64_t>(host_threads > 0 ? host_threads : int(std::thread::hardware_concurrency()), nT / 1024 + 1)));
  std::vector<int> bad(size_t(nth), 0);
  auto work = [&](int th) {
    for (int64_t t = nT * th / nth, e = nT * (th + 1) / nth; t < e; ++t) {
      const int64_t nPos = label_ptr[t + 1] - label_ptr[t];
      if (!metric_from_counts(kind, at_k, cnt + label_ptr[t] + t, nPos, nitems, per_user + t)) bad[size_t(th)] = 1;
    }
  };
  std::vector<std::thread> pool;
  for (int th = 1; th < nth; ++th) pool.emplace_back(work, th);
  work(0);
  for (auto& th : pool) th.join();
  for (int b : bad) {
    if (b) return qmfb::set_error(QMFB_ERR_INVALID, "qmfb_rank_metrics: '%s' is undefined for a test user (no positive item, or fewer items than k)", metric);
  }
  return QMFB_OK;
}

}  // extern "C"
