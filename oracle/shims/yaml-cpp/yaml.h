// oracle/shims/yaml-cpp/yaml.h -- functional stand-in for the part of yaml-cpp 0.5 the reference's parameter parser uses
// (src/types/parameters.cpp:36-44, 272-440).  TEST INFRASTRUCTURE ONLY (see oracle/shims/Eigen/Core): it lets the
// reference's own, unmodified ParameterCollection::parseFromFile read configurations/*.yaml here, so that the effective
// parameter values the product uses (vslam-pose-estimation-framework_b200/configs.py) are pinned against the reference's
// own parser, parse quirks included.  It is not yaml-cpp and shares no code with it.
//
// Subset: block mappings nested by indentation, plain / quoted scalars, `#` comments, empty values (null).  That is all
// configurations/*.yaml contain (no sequences, no flow collections, no anchors).
// Semantics kept from yaml-cpp 0.5:
//   * node["missing"] yields an undefined node; as<T>() of an undefined, null or map node throws TypedBadConversion<T>
//     (the reference catches exactly that and keeps the struct default, parameters.cpp:42)
//   * the first of two equal keys wins on look-up
//   * integers: the whole scalar must parse as the integer type ("25.0".as<int>() throws; unsigned rejects "-1")
//   * bool: true/false, yes/no, on/off, y/n in lower, UPPER or Capitalised spelling
//   * a missing file throws BadFile
#pragma once
#include <cstdint>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace YAML {

struct Mark { int pos = 0, line = 0, column = 0; };

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class BadFile : public Exception {
 public:
  BadFile() : Exception("bad file") {}
};
class ParserException : public Exception {
 public:
  explicit ParserException(const std::string& m) : Exception(m) {}
};
class BadConversion : public Exception {
 public:
  BadConversion() : Exception("bad conversion") {}
};
template <class T>
class TypedBadConversion : public BadConversion {};

class Node {
 public:
  enum Kind { Undefined, Null, Scalar, Map };
  Node() : _d(std::make_shared<Data>()) {}
  bool IsDefined() const { return _d->kind != Undefined; }
  bool IsNull() const { return _d->kind == Null; }
  bool IsScalar() const { return _d->kind == Scalar; }
  bool IsMap() const { return _d->kind == Map; }
  std::size_t size() const { return _d->children.size(); }
  const std::string& Scalar_() const { return _d->scalar; }
  Node operator[](const std::string& key) const {
    if (_d->kind == Map)
      for (const auto& kv : _d->children)
        if (kv.first == key) return kv.second;
    return Node();
  }
  template <class T> T as() const;

  // construction (used by the loader)
  void setScalar(const std::string& s) { _d->kind = Scalar; _d->scalar = s; }
  void setNull() { _d->kind = Null; }
  void setMap() { _d->kind = Map; }
  void add(const std::string& key, const Node& child) { _d->children.emplace_back(key, child); }

 private:
  struct Data {
    Kind kind = Undefined;
    std::string scalar;
    std::vector<std::pair<std::string, Node>> children;
  };
  std::shared_ptr<Data> _d;
};

namespace detail {
inline bool decode(const Node& n, std::string& out) {
  if (!n.IsScalar()) return false;
  out = n.Scalar_();
  return true;
}
inline bool decode(const Node& n, bool& out) {
  if (!n.IsScalar()) return false;
  static const char* const names[4][2] = {{"y", "n"}, {"yes", "no"}, {"true", "false"}, {"on", "off"}};
  const std::string& s = n.Scalar_();
  auto lower = [](std::string v) { for (auto& c : v) c = (char)std::tolower((unsigned char)c); return v; };
  auto flexible = [&](const std::string& v) {          // lower, UPPER or Capitalised
    if (v.empty()) return false;
    bool all_lower = true, all_upper = true, rest_lower = true;
    for (std::size_t i = 0; i < v.size(); ++i) {
      const bool up = std::isupper((unsigned char)v[i]), lo = std::islower((unsigned char)v[i]);
      all_lower &= lo; all_upper &= up;
      if (i) rest_lower &= lo;
    }
    return all_lower || all_upper || (std::isupper((unsigned char)v[0]) && rest_lower);
  };
  if (!flexible(s)) return false;
  for (const auto& pair : names) {
    if (lower(s) == pair[0]) { out = true; return true; }
    if (lower(s) == pair[1]) { out = false; return true; }
  }
  return false;
}
template <class T>
typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value, bool>::type decode(const Node& n, T& out) {
  if (!n.IsScalar()) return false;
  const std::string& s = n.Scalar_();
  if (std::is_floating_point<T>::value) {
    if (s == ".inf" || s == ".Inf" || s == ".INF" || s == "+.inf") { out = (T)INFINITY; return true; }
    if (s == "-.inf" || s == "-.Inf" || s == "-.INF") { out = (T)-INFINITY; return true; }
    if (s == ".nan" || s == ".NaN" || s == ".NAN") { out = (T)NAN; return true; }
  }
  if (std::is_unsigned<T>::value && !s.empty() && s[0] == '-') return false;
  std::stringstream stream(s);
  stream.unsetf(std::ios::dec);                          // hex / octal literals as yaml-cpp accepts them
  if ((stream >> std::noskipws >> out) && (stream >> std::ws).eof()) return true;
  return false;
}
}  // namespace detail

template <class T>
T Node::as() const {
  T value{};
  if (!detail::decode(*this, value)) throw TypedBadConversion<T>();
  return value;
}

namespace detail {
inline std::string strip_comment_and_space(const std::string& line) {
  std::string out;
  char quote = 0;
  for (std::size_t i = 0; i < line.size(); ++i) {
    const char c = line[i];
    if (quote) {
      if (c == quote) quote = 0;
    } else if (c == '"' || c == '\'') {
      quote = c;
    } else if (c == '#' && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) {
      break;
    }
    out.push_back(c);
  }
  while (!out.empty() && (out.back() == ' ' || out.back() == '\t' || out.back() == '\r' || out.back() == '\n')) out.pop_back();
  return out;
}
inline std::string trim(const std::string& s) {
  std::size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t')) --b;
  return s.substr(a, b - a);
}
struct Line { int indent; std::string key, value; bool has_value; };

inline Node build(const std::vector<Line>& lines, std::size_t& i, int indent) {
  Node map;
  map.setMap();
  while (i < lines.size() && lines[i].indent == indent) {
    const Line& l = lines[i++];
    Node child;
    if (l.has_value) {
      child.setScalar(l.value);
    } else if (i < lines.size() && lines[i].indent > indent) {
      child = build(lines, i, lines[i].indent);
    } else {
      child.setNull();
    }
    map.add(l.key, child);
  }
  if (i < lines.size() && lines[i].indent > indent) throw ParserException("bad indentation of a mapping entry");
  return map;
}
}  // namespace detail

inline Node Load(std::istream& in) {
  std::vector<detail::Line> lines;
  std::string raw;
  while (std::getline(in, raw)) {
    const std::string text = detail::strip_comment_and_space(raw);
    std::size_t indent = 0;
    while (indent < text.size() && text[indent] == ' ') ++indent;
    if (indent == text.size()) continue;
    if (text.compare(indent, 3, "---") == 0 || text.compare(indent, 3, "...") == 0) continue;
    // key ends at the first ": " or at a trailing ':' ("aligner->damping: 0" keeps "aligner->damping" as the key)
    std::size_t colon = std::string::npos;
    for (std::size_t k = indent; k < text.size(); ++k)
      if (text[k] == ':' && (k + 1 == text.size() || text[k + 1] == ' ' || text[k + 1] == '\t')) { colon = k; break; }
    if (colon == std::string::npos) throw ParserException("oracle/shims yaml: only block mappings are supported: " + raw);
    detail::Line l;
    l.indent = (int)indent;
    l.key = detail::trim(text.substr(indent, colon - indent));
    l.value = detail::trim(text.substr(colon + 1));
    if (l.value.size() >= 2 && (l.value.front() == '"' || l.value.front() == '\'') && l.value.back() == l.value.front())
      l.value = l.value.substr(1, l.value.size() - 2);
    l.has_value = !l.value.empty() && l.value != "~" && l.value != "null";
    lines.push_back(l);
  }
  std::size_t i = 0;
  if (lines.empty()) { Node n; n.setNull(); return n; }
  Node root = detail::build(lines, i, lines[0].indent);
  if (i != lines.size()) throw ParserException("bad indentation");
  return root;
}
inline Node Load(const std::string& text) {
  std::istringstream in(text);
  return Load(in);
}
inline Node LoadFile(const std::string& filename) {
  std::ifstream in(filename.c_str());
  if (!in) throw BadFile();
  return Load(in);
}

}  // namespace YAML
