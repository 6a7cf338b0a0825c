// TEST INFRASTRUCTURE — minimal stand-in for <gflags/gflags.h> (absent from this image) so the
// UNMODIFIED reference mains (qmf/wals.cpp:26-50, qmf/bpr.cpp:28-59) compile. Supports
// DEFINE_{bool,int32,uint64,double,string}, -f=v / --f=v / --f v / --flag / --noflag.
// Not part of the product; used only by oracle/Makefile (the recipe that compiles the unmodified reference into oracle/_ref/).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>

namespace gflags {
struct FlagReg {
  enum Kind { kBool, kInt32, kUint64, kDouble, kString } kind;
  void* ptr;
};
inline std::map<std::string, FlagReg>& registry() {
  static std::map<std::string, FlagReg> r;
  return r;
}
struct Registrar {
  Registrar(const char* name, FlagReg::Kind kind, void* ptr) { registry()[name] = FlagReg{kind, ptr}; }
};
inline void SetUsageMessage(const std::string&) {}
inline bool setFlag(const FlagReg& f, const std::string& v) {
  switch (f.kind) {
    case FlagReg::kBool: {
      bool b = !(v == "false" || v == "0" || v == "no" || v == "f" || v == "n");
      *static_cast<bool*>(f.ptr) = b;
      return true;
    }
    case FlagReg::kInt32: *static_cast<int32_t*>(f.ptr) = static_cast<int32_t>(std::strtol(v.c_str(), nullptr, 10)); return true;
    case FlagReg::kUint64: *static_cast<uint64_t*>(f.ptr) = std::strtoull(v.c_str(), nullptr, 10); return true;
    case FlagReg::kDouble: *static_cast<double*>(f.ptr) = std::strtod(v.c_str(), nullptr); return true;
    case FlagReg::kString: *static_cast<std::string*>(f.ptr) = v; return true;
  }
  return false;
}
inline uint32_t ParseCommandLineFlags(int* argc, char*** argv, bool removeFlags) {
  int out = 1;
  char** av = *argv;
  for (int i = 1; i < *argc; ++i) {
    std::string a = av[i];
    if (a.size() < 2 || a[0] != '-') { av[out++] = av[i]; continue; }
    size_t p = (a[1] == '-') ? 2 : 1;
    std::string name = a.substr(p), val;
    bool hasVal = false;
    size_t eq = name.find('=');
    if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); hasVal = true; }
    auto it = registry().find(name);
    if (it == registry().end() && name.compare(0, 2, "no") == 0) {
      auto it2 = registry().find(name.substr(2));
      if (it2 != registry().end() && it2->second.kind == FlagReg::kBool) { *static_cast<bool*>(it2->second.ptr) = false; continue; }
    }
    if (it == registry().end()) { std::cerr << "ERROR: unknown command line flag '" << name << "'\n"; std::exit(1); }
    if (!hasVal) {
      if (it->second.kind == FlagReg::kBool) { *static_cast<bool*>(it->second.ptr) = true; continue; }
      if (i + 1 >= *argc) { std::cerr << "ERROR: flag '" << name << "' is missing its argument\n"; std::exit(1); }
      val = av[++i];
    }
    setFlag(it->second, val);
  }
  if (removeFlags) { *argc = out; av[out] = nullptr; }
  return out;
}
}  // namespace gflags


#define QMF_SHIM_DEFINE(type, kind, name, dflt) \
  type FLAGS_##name = dflt;                     \
  static gflags::Registrar qmf_shim_reg_##name(#name, gflags::FlagReg::kind, &FLAGS_##name)
#define DEFINE_bool(name, dflt, help) QMF_SHIM_DEFINE(bool, kBool, name, dflt)
#define DEFINE_int32(name, dflt, help) QMF_SHIM_DEFINE(int32_t, kInt32, name, dflt)
#define DEFINE_uint64(name, dflt, help) QMF_SHIM_DEFINE(uint64_t, kUint64, name, dflt)
#define DEFINE_double(name, dflt, help) QMF_SHIM_DEFINE(double, kDouble, name, dflt)
#define DEFINE_string(name, dflt, help) QMF_SHIM_DEFINE(std::string, kString, name, dflt)
