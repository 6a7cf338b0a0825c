#include <qmf/DatasetReader.h>

#include <cstdio>
#include <fstream>

#include <qmf/utils/Log.h>

namespace qmf {

DatasetReader::DatasetReader(const std::string& fileName) : stream_(new std::ifstream(fileName)) {}

bool DatasetReader::readOne(DatasetElem& elem) {
  CHECK(stream_ != nullptr);
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

void DatasetReader::readAll(std::vector<DatasetElem>& dataset) {
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
