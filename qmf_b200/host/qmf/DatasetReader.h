// "<int64 user> <int64 item> <double weight>" text datasets (README.md:68-74 of the reference;
// surface of qmf/DatasetReader.h:29-58)
#pragma once
#include <istream>
#include <memory>
#include <string>
#include <vector>

#include <qmf/Types.h>
#include <qmf/utils/FriendTest.h>

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
  // Same result as a readOne() loop.  For a regular file that has not been read from yet this
  // maps the file and parses line ranges on all host threads (at 100M+
  // lines the getline + sscanf loop dominates the wall time of a run).
  void readAll(std::vector<DatasetElem>& dataset);

  // Parses one line [b, e) exactly like sscanf(line, "%lld %lld %lf") == 3 does (same integer
  // clamping, same correctly-rounded double); false if the line does not carry three fields.
  static bool parseLine(const char* b, const char* e, DatasetElem& elem);
  // The first number of [b, e) exactly like sscanf(line, "%lf") == 1 does (leading white space skipped, same correctly
  // rounded value, trailing text ignored); false if there is none.  Used by FactorData::setFactors(file).
  static bool parseDoubleField(const char* b, const char* e, double& out);

 private:
  bool readAllMapped(std::vector<DatasetElem>& dataset);

  std::unique_ptr<std::istream> stream_;
  std::string fileName_;
  bool touched_ = false;  // readOne() has consumed part of the stream
  std::string line_;

  FRIEND_TEST(DatasetReader, readOne);
  FRIEND_TEST(DatasetReader, readOneBadFormat);
  FRIEND_TEST(DatasetReader, readAll);
};

}  // namespace qmf
