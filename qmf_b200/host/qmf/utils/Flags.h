// Tiny command-line flag registry with the gflags surface the reference's mains use
// (DEFINE_* + ParseCommandLineFlags; accepts -f=v, --f=v, --f v, --flag, --noflag).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <map>
#include <string>

namespace qmf {
namespace flags {

struct Entry {
  std::function<void(const std::string&)> set;
  bool isBool;
};

inline std::map<std::string, Entry>& table() {
  static std::map<std::string, Entry> t;
  return t;
}

struct Register {
  Register(const char* name, bool isBool, std::function<void(const std::string&)> setter) {
    table()[name] = Entry{std::move(setter), isBool};
  }
};

inline bool toBool(const std::string& v) { return !(v == "false" || v == "0" || v == "no" || v == "f" || v == "n"); }

inline void parse(int argc, char** argv) {
  for (int a = 1; a < argc; ++a) {
    std::string arg = argv[a];
    if (arg.size() < 2 || arg[0] != '-') continue;
    arg = arg.substr(arg[1] == '-' ? 2 : 1);
    std::string value;
    bool hasValue = false;
    const size_t eq = arg.find('=');
    if (eq != std::string::npos) {
      value = arg.substr(eq + 1);
      arg = arg.substr(0, eq);
      hasValue = true;
    }
    auto it = table().find(arg);
    if (it == table().end() && arg.compare(0, 2, "no") == 0) {
      auto neg = table().find(arg.substr(2));
      if (neg != table().end() && neg->second.isBool) {
        neg->second.set("false");
        continue;
      }
    }
    if (it == table().end()) {
      std::cerr << "ERROR: unknown command line flag '" << arg << "'\n";
      std::exit(1);
    }
    if (!hasValue) {
      if (it->second.isBool) {
        value = "true";
      } else if (a + 1 < argc) {
        value = argv[++a];
      } else {
        std::cerr << "ERROR: flag '" << arg << "' is missing its argument\n";
        std::exit(1);
      }
    }
    it->second.set(value);
  }
}

}  // namespace flags
}  // namespace qmf

#define QMF_DEFINE_FLAG(type, name, dflt, isBool, conv)                                       \
  type FLAGS_##name = dflt;                                                                   \
  static ::qmf::flags::Register qmf_flag_##name(#name, isBool, [](const std::string& v) { FLAGS_##name = conv; })
#define DEFINE_uint64(name, dflt, help) QMF_DEFINE_FLAG(uint64_t, name, dflt, false, std::strtoull(v.c_str(), nullptr, 10))
#define DEFINE_int32(name, dflt, help) QMF_DEFINE_FLAG(int32_t, name, dflt, false, static_cast<int32_t>(std::strtol(v.c_str(), nullptr, 10)))
#define DEFINE_int64(name, dflt, help) QMF_DEFINE_FLAG(int64_t, name, dflt, false, std::strtoll(v.c_str(), nullptr, 10))
#define DEFINE_double(name, dflt, help) QMF_DEFINE_FLAG(double, name, dflt, false, std::strtod(v.c_str(), nullptr))
#define DEFINE_string(name, dflt, help) QMF_DEFINE_FLAG(std::string, name, dflt, false, v)
#define DEFINE_bool(name, dflt, help) QMF_DEFINE_FLAG(bool, name, dflt, true, ::qmf::flags::toBool(v))
