#pragma once
#include <string>
#include <vector>

namespace qmf {

// split on a delimiter, dropping empty pieces (flag lists such as "auc,p@10")
inline std::vector<std::string> split(const std::string& text, char delim) {
  std::vector<std::string> parts;
  std::string cur;
  for (const char c : text) {
    if (c == delim) {
      if (!cur.empty()) parts.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  if (!cur.empty()) parts.push_back(cur);
  return parts;
}

}  // namespace qmf
