// Tiny gflags-compatible flag layer (the reference fetches gflags over the
// network at configure time; there is no network here).  Supports the syntax
// the reference's scripts use: --name=value, --name value, --name / --noname
// for booleans.  Unknown flags are an error, as with gflags.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace miniflags {
struct Entry {
    std::string type, help, def;
    std::function<void(const std::string &)> set;
    bool is_bool = false;
};
inline std::map<std::string, Entry> &registry()
{
    static std::map<std::string, Entry> r;
    return r;
}
template <typename T>
inline T parse(const std::string &s);
template <>
inline bool parse<bool>(const std::string &s)
{
    if (s == "true" || s == "1" || s == "yes" || s == "t" || s == "y") return true;
    if (s == "false" || s == "0" || s == "no" || s == "f" || s == "n") return false;
    throw std::runtime_error("illegal value '" + s + "' for a bool flag");
}
template <>
inline std::string parse<std::string>(const std::string &s)
{
    return s;
}
template <>
inline double parse<double>(const std::string &s)
{
    return std::stod(s);
}
template <>
inline uint32_t parse<uint32_t>(const std::string &s)
{
    return (uint32_t)std::stoul(s);
}
template <>
inline int32_t parse<int32_t>(const std::string &s)
{
    return (int32_t)std::stol(s);
}
template <typename T>
struct Registrar {
    Registrar(const char *name, T *var, const char *type, const char *help, const std::string &def)
    {
        Entry e;
        e.type = type;
        e.help = help;
        e.def = def;
        e.is_bool = std::string(type) == "bool";
        e.set = [var](const std::string &s) { *var = parse<T>(s); };
        registry()[name] = e;
    }
};
inline void usage(const std::string &msg)
{
    std::cout << msg << "\n";
    for (auto &kv : registry())
        std::cout << "    --" << kv.first << " (" << kv.second.help << ")  type: " << kv.second.type
                  << "  default: " << kv.second.def << "\n";
}
inline void ParseCommandLineFlags(int *argc, char ***argv, const std::string &usage_msg)
{
    for (int i = 1; i < *argc; ++i) {
        std::string a = (*argv)[i];
        if (a == "--help" || a == "-help" || a == "--helpfull") {
            usage(usage_msg);
            std::exit(0);
        }
        if (a.rfind("--", 0) != 0 && a.rfind("-", 0) == 0) a = "-" + a;   // -flag == --flag
        if (a.rfind("--", 0) != 0) throw std::runtime_error("unexpected argument '" + a + "'");
        std::string name = a.substr(2), val;
        bool has_val = false;
        auto eq = name.find('=');
        if (eq != std::string::npos) {
            val = name.substr(eq + 1);
            name = name.substr(0, eq);
            has_val = true;
        }
        auto it = registry().find(name);
        if (it == registry().end() && name.rfind("no", 0) == 0 && !has_val) {
            auto it2 = registry().find(name.substr(2));
            if (it2 != registry().end() && it2->second.is_bool) {
                it2->second.set("false");
                continue;
            }
        }
        if (it == registry().end())
            throw std::runtime_error("unknown command line flag '" + name + "'");
        if (!has_val) {
            if (it->second.is_bool) {
                val = "true";
            } else if (i + 1 < *argc) {
                val = (*argv)[++i];
            } else {
                throw std::runtime_error("flag '--" + name + "' is missing its argument");
            }
        }
        it->second.set(val);
    }
}
}  // namespace miniflags

#define MINIFLAGS_DEFINE(ctype, tname, name, def, help)                                        \
    ctype FLAGS_##name = def;                                                                  \
    static miniflags::Registrar<ctype> miniflags_reg_##name(#name, &FLAGS_##name, tname, help, \
                                                            #def)
#define DEFINE_bool(name, def, help) MINIFLAGS_DEFINE(bool, "bool", name, def, help)
#define DEFINE_uint32(name, def, help) MINIFLAGS_DEFINE(uint32_t, "uint32", name, def, help)
#define DEFINE_int32(name, def, help) MINIFLAGS_DEFINE(int32_t, "int32", name, def, help)
#define DEFINE_double(name, def, help) MINIFLAGS_DEFINE(double, "double", name, def, help)
#define DEFINE_string(name, def, help) MINIFLAGS_DEFINE(std::string, "string", name, def, help)
