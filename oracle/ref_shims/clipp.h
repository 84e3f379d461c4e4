// Minimal work-alike of muellan/clipp (v1.2.3 is an un-vendored dependency of the reference, pkg.yaml:22-24) covering
// exactly what cli/main.cpp:36-45 and benchmark/main.cpp use:
//   clipp::value(label, target), clipp::option(flags...).doc(text) & clipp::value(label, target), the comma operator
//   that concatenates parameters into a group, parse(argc, argv, group) and clipp::make_man_page(group, program).
// Used ONLY to compile the reference's unchanged drivers for the drop-in check. Not part of the product.
#ifndef SPMV_B200_CLIPP_SHIM_H
#define SPMV_B200_CLIPP_SHIM_H

#include <ostream>
#include <string>
#include <vector>

namespace clipp {

struct parameter {
  std::vector<std::string> flags; // empty: positional value
  std::string label, doc_text;
  std::string *target = nullptr;  // where a value goes (positional, or the value following an option)
  bool has_value = false;
  std::string value_label;

  parameter &doc(const std::string &d) {
    doc_text = d;
    return *this;
  }
};

struct group {
  std::vector<parameter> params;
};

inline parameter value(const std::string &label, std::string &target) {
  parameter p;
  p.label = label;
  p.target = &target;
  p.has_value = true;
  return p;
}

template <typename... S> inline parameter option(S... names) {
  parameter p;
  p.flags = {std::string(names)...};
  return p;
}

// option & value : the option takes the value as its argument
inline parameter operator&(parameter opt, const parameter &val) {
  opt.target = val.target;
  opt.has_value = true;
  opt.value_label = val.label;
  return opt;
}

inline group operator,(const parameter &a, const parameter &b) {
  group g;
  g.params = {a, b};
  return g;
}

inline group operator,(group g, const parameter &b) {
  g.params.push_back(b);
  return g;
}

inline bool parse(int argc, char **argv, const group &g) {
  std::vector<bool> positional_done(g.params.size(), false);
  for (int i = 1; i < argc; ++i) {
    const std::string arg = argv[i];
    bool matched = false;
    for (const auto &p : g.params) {
      for (const auto &f : p.flags) {
        if (arg == f) {
          matched = true;
          if (p.has_value) {
            if (i + 1 >= argc)
              return false;
            *p.target = argv[++i];
          }
        }
      }
    }
    if (matched)
      continue;
    if (!arg.empty() && arg[0] == '-')
      return false; // unknown option
    for (std::size_t k = 0; k < g.params.size(); ++k) {
      if (g.params[k].flags.empty() && !positional_done[k]) {
        *g.params[k].target = arg;
        positional_done[k] = true;
        matched = true;
        break;
      }
    }
    if (!matched)
      return false;
  }
  for (std::size_t k = 0; k < g.params.size(); ++k)
    if (g.params[k].flags.empty() && !positional_done[k])
      return false; // required positional value missing
  return true;
}

struct man_page {
  std::string text;
};

inline man_page make_man_page(const group &g, const std::string &program) {
  man_page m;
  m.text = "SYNOPSIS\n        " + program;
  for (const auto &p : g.params) {
    if (p.flags.empty())
      m.text += " <" + p.label + ">";
    else
      m.text += " [" + p.flags[0] + (p.has_value ? " <" + p.value_label + ">" : "") + "]";
  }
  m.text += "\n\nOPTIONS\n";
  for (const auto &p : g.params) {
    if (p.flags.empty())
      continue;
    std::string names;
    for (std::size_t i = 0; i < p.flags.size(); ++i)
      names += (i ? ", " : "") + p.flags[i];
    m.text += "        " + names + (p.has_value ? " <" + p.value_label + ">" : "") + "\n                    " + p.doc_text + "\n";
  }
  return m;
}

inline std::ostream &operator<<(std::ostream &os, const man_page &m) { return os << m.text; }

} // namespace clipp

#endif
