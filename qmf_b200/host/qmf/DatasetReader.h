// "<int64 user> <int64 item> <double weight>" text datasets (README.md:68-74 of the reference;
// surface of qmf/DatasetReader.h:29-58)
#pragma once
#include <istream>
#include <memory>
#include <string>
#include <vector>

#include <qmf/Types.h>

namespace qmf {

struct DatasetElem {
  int64_t userId;
  int64_t itemId;
  Double value = 1.0;
};

class DatasetReader {
 public:
  DatasetReader() = default;
  explicit DatasetReader(const std::string& fileName);
  explicit DatasetReader(std::unique_ptr<std::istream> stream) : stream_(std::move(stream)) {}

  bool readOne(DatasetElem& elem);
  std::vector<DatasetElem> readAll();
  void readAll(std::vector<DatasetElem>& dataset);

 private:
  std::unique_ptr<std::istream> stream_;
  std::string line_;
};

}  // namespace qmf
