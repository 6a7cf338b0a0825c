// Times the ingest path of this repo on one dataset file and prints one JSON line:
//   read    : DatasetReader::readAll (mapped, all host threads)   [reference: getline + sscanf loop]
//   read1   : the same file through the readOne() loop (the reference's algorithm, this build)
//   signals : qmfb_signals_build (dense indexing + both CSR orientations on the GPU, host arrays in)
// usage: ingest_bench file [device] [--no-slow]
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include <qmf/DatasetReader.h>

#include "qmf_b200.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); };
  const int device = argc > 2 ? std::atoi(argv[2]) : 0;
  const bool slow = !(argc > 3 && std::strcmp(argv[3], "--no-slow") == 0);
  std::vector<qmf::DatasetElem> d;
  auto t0 = clk::now();
  {
    qmf::DatasetReader r(argv[1]);
    r.readAll(d);
  }
  const double tRead = secs(t0);
  double tRead1 = -1.0;
  if (slow) {
    std::vector<qmf::DatasetElem> d1;
    t0 = clk::now();
    qmf::DatasetReader r(argv[1]);
    qmf::DatasetElem e;
    while (r.readOne(e)) d1.push_back(e);
    tRead1 = secs(t0);
    if (d1.size() != d.size()) return 3;
  }
  std::vector<int64_t> u(d.size()), i(d.size());
  std::vector<double> v(d.size());
  for (size_t p = 0; p < d.size(); ++p) {
    u[p] = d[p].userId;
    i[p] = d[p].itemId;
    v[p] = d[p].value;
  }
  qmfb_signals_t* s = nullptr;
  if (qmfb_signals_build(device, 1, u.data(), i.data(), v.data(), &s) != 0) {  // warm-up: context creation
    std::fprintf(stderr, "%s\n", qmfb_last_error());
    return 1;
  }
  qmfb_signals_destroy(s);
  t0 = clk::now();
  if (qmfb_signals_build(device, int64_t(d.size()), u.data(), i.data(), v.data(), &s) != 0) {
    std::fprintf(stderr, "%s\n", qmfb_last_error());
    return 1;
  }
  const double tSig = secs(t0);
  int64_t nu = 0, ni = 0;
  qmfb_signals_dims(s, &nu, &ni, nullptr);
  qmfb_signals_destroy(s);
  std::printf("{\"lines\": %zu, \"nusers\": %lld, \"nitems\": %lld, \"read_s\": %.3f, \"read_getline_sscanf_s\": %.3f, \"signals_gpu_s\": %.3f}\n",
              d.size(), (long long)nu, (long long)ni, tRead, tRead1, tSig);
  return 0;
}
