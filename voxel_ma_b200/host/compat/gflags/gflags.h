// Minimal gflags-compatible command-line flag layer (header-only).
//
// The reference CLI (src/voroUtility.cpp:4,58-142,262-267) is written against gflags 2.2.2, which
// is fetched by Bazel (WORKSPACE:28-49) and is neither vendored in the reference tree nor present
// in this image.  This header provides exactly the surface that file uses, with gflags' observable
// behaviour: DEFINE_{string,bool,int32,double}, DEFINE_validator, ParseCommandLineFlags (flags
// removed from argv, "-f=v" / "--f=v" / "-f v" / "-boolflag" / "-noboolflag", "--" terminator,
// validators also run on untouched defaults), SetUsageMessage / ProgramUsage,
// ProgramInvocationShortName, GetCommandLineFlagInfoOrDie and DescribeOneFlag.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace google
{
struct CommandLineFlagInfo
{
    std::string name, type, description, current_value, default_value, filename;
    bool has_validator_fn = false;
    bool is_default = true;
    const void* flag_ptr = nullptr;
};

namespace detail
{
struct Flag
{
    std::string name, type, help, defval;
    void* ptr = nullptr;
    bool modified = false;
    bool (*vs)(const char*, const std::string&) = nullptr;
    bool (*vb)(const char*, bool) = nullptr;
    bool (*vi)(const char*, int32_t) = nullptr;
    bool (*vd)(const char*, double) = nullptr;

    std::string current() const
    {
        std::ostringstream o;
        if (type == "string") o << *static_cast<std::string*>(ptr);
        else if (type == "bool") o << (*static_cast<bool*>(ptr) ? "true" : "false");
        else if (type == "int32") o << *static_cast<int32_t*>(ptr);
        else if (type == "double") o << *static_cast<double*>(ptr);
        return o.str();
    }
    bool set(const std::string& v)
    {
        if (type == "string") { *static_cast<std::string*>(ptr) = v; return true; }
        if (type == "bool")
        {
            static const char* t[] = {"1", "t", "true", "y", "yes"};
            static const char* f[] = {"0", "f", "false", "n", "no"};
            for (auto s : t) if (v == s) { *static_cast<bool*>(ptr) = true; return true; }
            for (auto s : f) if (v == s) { *static_cast<bool*>(ptr) = false; return true; }
            return false;
        }
        char* end = nullptr;
        if (type == "int32")
        {
            long x = strtol(v.c_str(), &end, 0);
            if (v.empty() || *end) return false;
            *static_cast<int32_t*>(ptr) = (int32_t)x;
            return true;
        }
        if (type == "double")
        {
            double x = strtod(v.c_str(), &end);
            if (v.empty() || *end) return false;
            *static_cast<double*>(ptr) = x;
            return true;
        }
        return false;
    }
    bool validate() const
    {
        if (vs) return vs(name.c_str(), *static_cast<std::string*>(ptr));
        if (vb) return vb(name.c_str(), *static_cast<bool*>(ptr));
        if (vi) return vi(name.c_str(), *static_cast<int32_t*>(ptr));
        if (vd) return vd(name.c_str(), *static_cast<double*>(ptr));
        return true;
    }
    bool has_validator() const { return vs || vb || vi || vd; }
};

struct Registry
{
    std::map<std::string, Flag> flags;
    std::string usage, argv0;
    static Registry& get()
    {
        static Registry r;
        return r;
    }
    Flag* by_ptr(const void* p)
    {
        for (auto& kv : flags)
            if (kv.second.ptr == p) return &kv.second;
        return nullptr;
    }
};

struct Registrar
{
    Registrar(const char* name, const char* type, const char* help, void* ptr)
    {
        Flag f;
        f.name = name;
        f.type = type;
        f.help = help;
        f.ptr = ptr;
        f.defval = f.current();
        Registry::get().flags[name] = f;
    }
};

[[noreturn]] inline void die(const std::string& msg)
{
    fprintf(stderr, "ERROR: %s\n", msg.c_str());
    exit(1);
}
} // namespace detail

inline void SetUsageMessage(const std::string& u) { detail::Registry::get().usage = u; }
inline const char* ProgramUsage() { return detail::Registry::get().usage.c_str(); }
inline const char* ProgramInvocationShortName()
{
    const std::string& a = detail::Registry::get().argv0;
    size_t p = a.find_last_of("/\\");
    return p == std::string::npos ? a.c_str() : a.c_str() + p + 1;
}

inline bool RegisterFlagValidator(const std::string* f, bool (*fn)(const char*, const std::string&))
{
    auto* fl = detail::Registry::get().by_ptr(f);
    if (!fl) return false;
    fl->vs = fn;
    return true;
}
inline bool RegisterFlagValidator(const bool* f, bool (*fn)(const char*, bool))
{
    auto* fl = detail::Registry::get().by_ptr(f);
    if (!fl) return false;
    fl->vb = fn;
    return true;
}
inline bool RegisterFlagValidator(const int32_t* f, bool (*fn)(const char*, int32_t))
{
    auto* fl = detail::Registry::get().by_ptr(f);
    if (!fl) return false;
    fl->vi = fn;
    return true;
}
inline bool RegisterFlagValidator(const double* f, bool (*fn)(const char*, double))
{
    auto* fl = detail::Registry::get().by_ptr(f);
    if (!fl) return false;
    fl->vd = fn;
    return true;
}

inline CommandLineFlagInfo GetCommandLineFlagInfoOrDie(const char* name)
{
    auto& r = detail::Registry::get();
    auto it = r.flags.find(name);
    if (it == r.flags.end())
    {
        fprintf(stderr, "FATAL ERROR: flag name '%s' doesn't exist\n", name);
        exit(1);
    }
    const auto& f = it->second;
    CommandLineFlagInfo i;
    i.name = f.name;
    i.type = f.type;
    i.description = f.help;
    i.current_value = f.current();
    i.default_value = f.defval;
    i.has_validator_fn = f.has_validator();
    i.is_default = !f.modified;
    i.flag_ptr = f.ptr;
    return i;
}

inline std::string DescribeOneFlag(const CommandLineFlagInfo& f)
{
    std::string s = "    -" + f.name + " (" + f.description + ") type: " + f.type + " default: ";
    s += f.type == "string" ? "\"" + f.default_value + "\"" : f.default_value;
    if (f.current_value != f.default_value)
        s += " currently: " + (f.type == "string" ? "\"" + f.current_value + "\"" : f.current_value);
    return s + "\n";
}

// returns the index of the first non-flag argument (always 1 when remove_flags is true)
inline uint32_t ParseCommandLineFlags(int* argc, char*** argv, bool remove_flags)
{
    using namespace detail;
    auto& r = Registry::get();
    char** av = *argv;
    r.argv0 = av[0] ? av[0] : "";
    std::vector<char*> rest;
    bool stop = false;
    for (int i = 1; i < *argc; ++i)
    {
        char* a = av[i];
        if (stop || a[0] != '-' || a[1] == '\0') { rest.push_back(a); continue; }
        const char* p = a + 1;
        if (*p == '-') ++p;
        if (*p == '\0') { stop = true; continue; } // "--"
        std::string name(p), val;
        bool has_val = false;
        size_t eq = name.find('=');
        if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
        auto it = r.flags.find(name);
        if (it == r.flags.end() && name.compare(0, 2, "no") == 0 && !has_val)
        {
            auto it2 = r.flags.find(name.substr(2));
            if (it2 != r.flags.end() && it2->second.type == "bool") { it = it2; val = "false"; has_val = true; }
        }
        if (it == r.flags.end()) die("unknown command line flag '" + name + "'");
        Flag& f = it->second;
        if (!has_val)
        {
            if (f.type == "bool") val = "true";
            else
            {
                if (i + 1 >= *argc) die("flag '-" + name + "' is missing its argument");
                val = av[++i];
            }
        }
        std::string before = f.current();
        if (!f.set(val)) die("illegal value '" + val + "' specified for " + f.type + " flag '" + name + "'");
        if (!f.validate())
        {
            f.set(before);
            die("failed validation of new value '" + val + "' for flag '" + name + "'");
        }
        f.modified = true;
    }
    for (auto& kv : r.flags)
        if (!kv.second.modified && !kv.second.validate())
            die("--" + kv.first + " must be set on the commandline (default value fails validation)");
    if (remove_flags)
    {
        for (size_t k = 0; k < rest.size(); ++k) av[1 + k] = rest[k];
        *argc = 1 + (int)rest.size();
        av[*argc] = nullptr;
        return 1;
    }
    return 1;
}
inline void ShutDownCommandLineFlags() {}
} // namespace google
namespace gflags = google;

#define VC_GFLAGS_DEFINE(ctype, tname, name, val, txt)                                   \
    ctype FLAGS_##name = val;                                                            \
    static ::google::detail::Registrar vc_flag_registrar_##name(#name, tname, txt, &FLAGS_##name)
#define DEFINE_string(name, val, txt) VC_GFLAGS_DEFINE(std::string, "string", name, val, txt)
#define DEFINE_bool(name, val, txt) VC_GFLAGS_DEFINE(bool, "bool", name, val, txt)
#define DEFINE_int32(name, val, txt) VC_GFLAGS_DEFINE(int32_t, "int32", name, val, txt)
#define DEFINE_double(name, val, txt) VC_GFLAGS_DEFINE(double, "double", name, val, txt)
#define DECLARE_string(name) extern std::string FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
#define DEFINE_validator(name, fn) \
    static const bool name##_validator_registered = ::google::RegisterFlagValidator(&FLAGS_##name, fn)
