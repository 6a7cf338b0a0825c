// Writes N uniform(-bound, bound) doubles, one per line with 9 decimals: the --distribution_file
// consumed by `wals` (role of the reference's qmf/gen_uniform.cpp:7-30, plus a seed so that runs
// are repeatable).  usage: gen_uniform [count=1000000] [file=uniform.dat] [seed] [bound=0.01]
#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char** argv) {
  const long count = argc > 1 ? std::atol(argv[1]) : 1000000;
  const char* file = argc > 2 ? argv[2] : "uniform.dat";
  std::mt19937 gen(argc > 3 ? static_cast<unsigned>(std::atol(argv[3])) : std::random_device()());
  const double bound = argc > 4 ? std::atof(argv[4]) : 0.01;
  std::uniform_real_distribution<double> dist(-bound, bound);
  std::FILE* out = std::fopen(file, "w");
  if (out == nullptr) {
    std::perror(file);
    return 1;
  }
  for (long n = 0; n < count; ++n) std::fprintf(out, "%.9f\n", dist(gen));
  std::fclose(out);
  return 0;
}
