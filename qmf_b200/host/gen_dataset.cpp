// Writes a synthetic "<user> <item> <weight>" dataset (uniform users/items, integer weights 1..5):
// the input of the ingest measurements.  usage: gen_dataset nnz nusers nitems file [seed]
#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: gen_dataset nnz nusers nitems file [seed]\n");
    return 2;
  }
  const long nnz = std::atol(argv[1]), nu = std::atol(argv[2]), ni = std::atol(argv[3]);
  std::mt19937_64 gen(argc > 5 ? std::atol(argv[5]) : 1);
  std::FILE* out = std::fopen(argv[4], "w");
  if (out == nullptr) {
    std::perror(argv[4]);
    return 1;
  }
  static char buf[1 << 20];
  std::setvbuf(out, buf, _IOFBF, sizeof(buf));
  for (long n = 0; n < nnz; ++n) {
    const unsigned long long r = gen();
    std::fprintf(out, "%llu %llu %llu\n", (r >> 20) % nu, (r >> 3) % ni, 1 + r % 5);
  }
  std::fclose(out);
  return 0;
}
