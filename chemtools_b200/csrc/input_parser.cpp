// YAML case-set front end of the reactor application (host side, C++14).
//
// Mirrors Input::Parser of the reference: src/input_parser.cpp:27-122 (decode of a case set), :147-170 (Parse),
// :172-351 (Validate) and include/input/types.hpp:99-111, :158-202 (unit conversion, composition types), with the
// same schema: a top-level sequence of case sets, each with `Reactor`, `Mechanism{Path,
// Description}`, `Description`, and a `Cases` mapping whose REPEATED `Case` keys are all significant.
// yaml-cpp and the LLNL units library are not available, so a small block/flow YAML reader and a unit table for the
// units the schema uses live here.  Error classes follow what the reference throws (tests/src/input_parser.cpp:27-94):
//   missing file -> YAML::BadFile, missing mandatory key -> YAML::InvalidNode, non-numeric value -> YAML::BadConversion,
//   missing mechanism -> filesystem_error, physical validation -> std::runtime_error, bad unit -> std::invalid_argument.
// Added over the reference (its gap): conversion of a case to (T, p, Y[K]).
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <set>
#include <sstream>

#include "input_parser.hpp"

namespace ctb {
namespace {

// ------------------------------------------------------------------ a small YAML reader
struct Line { int indent; std::string text; int no; };

std::string strip(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string strip_comment(const std::string& s) {
  bool sq = false, dq = false;
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '\'' && !dq) sq = !sq;
    else if (c == '"' && !sq) dq = !dq;
    else if (c == '#' && !sq && !dq && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
  }
  return s;
}
std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
  return s;
}

class Reader {
 public:
  explicit Reader(std::vector<Line> lines) : L(std::move(lines)) {}
  YamlNode parse_document() {
    if (L.empty()) return YamlNode();
    return parse_block(L[0].indent);
  }

 private:
  std::vector<Line> L;
  size_t pos = 0;

  static YamlNode scalar(const std::string& s) {
    YamlNode n;
    n.kind = YamlNode::Scalar;
    n.scalar = unquote(strip(s));
    return n;
  }
  // flow collections: [a, [b, c]]  {k: v}
  static YamlNode parse_flow(const std::string& s, size_t& i) {
    auto skip = [&]() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; };
    skip();
    YamlNode n;
    if (i < s.size() && s[i] == '[') {
      n.kind = YamlNode::Sequence;
      ++i;
      for (;;) {
        skip();
        if (i >= s.size()) throw InputError(InputError::InvalidNode, "unterminated flow sequence");
        if (s[i] == ']') { ++i; break; }
        n.items.push_back(parse_flow(s, i));
        skip();
        if (i < s.size() && s[i] == ',') ++i;
      }
      return n;
    }
    if (i < s.size() && s[i] == '{') {
      n.kind = YamlNode::Mapping;
      ++i;
      for (;;) {
        skip();
        if (i >= s.size()) throw InputError(InputError::InvalidNode, "unterminated flow mapping");
        if (s[i] == '}') { ++i; break; }
        size_t c = s.find(':', i);
        if (c == std::string::npos) throw InputError(InputError::InvalidNode, "bad flow mapping");
        std::string key = unquote(strip(s.substr(i, c - i)));
        i = c + 1;
        n.entries.emplace_back(key, parse_flow(s, i));
        skip();
        if (i < s.size() && s[i] == ',') ++i;
      }
      return n;
    }
    size_t b = i;
    bool sq = false, dq = false;
    while (i < s.size()) {
      char c = s[i];
      if (c == '\'' && !dq) sq = !sq;
      else if (c == '"' && !sq) dq = !dq;
      else if (!sq && !dq && (c == ',' || c == ']' || c == '}')) break;
      ++i;
    }
    return scalar(s.substr(b, i - b));
  }
  static YamlNode value_of(const std::string& v) {
    std::string t = strip(v);
    if (!t.empty() && (t[0] == '[' || t[0] == '{')) { size_t i = 0; return parse_flow(t, i); }
    return scalar(t);
  }
  static bool split_key(const std::string& t, std::string& key, std::string& rest) {
    bool sq = false, dq = false;
    for (size_t i = 0; i < t.size(); ++i) {
      char c = t[i];
      if (c == '\'' && !dq) sq = !sq;
      else if (c == '"' && !sq) dq = !dq;
      else if (c == ':' && !sq && !dq && (i + 1 == t.size() || std::isspace((unsigned char)t[i + 1]))) {
        key = unquote(strip(t.substr(0, i)));
        rest = strip(t.substr(i + 1));
        return true;
      }
    }
    return false;
  }

  YamlNode parse_block(int indent) {
    if (pos >= L.size()) return YamlNode();
    if (L[pos].text.compare(0, 1, "-") == 0 && (L[pos].text.size() == 1 || L[pos].text[1] == ' ')) return parse_sequence(indent);
    return parse_mapping(indent);
  }
  YamlNode parse_sequence(int indent) {
    YamlNode n;
    n.kind = YamlNode::Sequence;
    while (pos < L.size() && L[pos].indent == indent && L[pos].text[0] == '-' && (L[pos].text.size() == 1 || L[pos].text[1] == ' ')) {
      std::string rest = L[pos].text.substr(1);
      size_t lead = rest.find_first_not_of(' ');
      if (lead == std::string::npos) {  // "-" alone: nested block
        ++pos;
        n.items.push_back(pos < L.size() && L[pos].indent > indent ? parse_block(L[pos].indent) : YamlNode());
        continue;
      }
      // "- key: value" starts an inline mapping whose further keys are indented to the key column
      int child_indent = indent + 1 + (int)lead;
      std::string body = rest.substr(lead), key, val;
      if (split_key(body, key, val)) {
        L[pos].indent = child_indent;
        L[pos].text = body;
        n.items.push_back(parse_mapping(child_indent));
      } else {
        n.items.push_back(value_of(body));
        ++pos;
      }
    }
    return n;
  }
  YamlNode parse_mapping(int indent) {
    YamlNode n;
    n.kind = YamlNode::Mapping;
    while (pos < L.size() && L[pos].indent == indent) {
      std::string key, rest;
      if (!split_key(L[pos].text, key, rest)) throw InputError(InputError::InvalidNode, "line " + std::to_string(L[pos].no) + ": expected 'key: value'");
      ++pos;
      if (!rest.empty()) n.entries.emplace_back(key, value_of(rest));
      else if (pos < L.size() && L[pos].indent > indent) n.entries.emplace_back(key, parse_block(L[pos].indent));
      else n.entries.emplace_back(key, YamlNode());  // null
    }
    if (pos < L.size() && L[pos].indent > indent) throw InputError(InputError::InvalidNode, "line " + std::to_string(L[pos].no) + ": bad indentation");
    return n;
  }
};

// ------------------------------------------------------------------ units (include/input/types.hpp:99-111)
struct UnitDef { const char* name; int dim; double scale, offset; };  // value_SI = (value + offset) * scale
// A closed table, not the LLNL units library the reference links (units::unit_from_string): the unit strings of the
// reference's fixtures and documentation plus the common SI-prefixed forms; anything else is "not a valid unit".  Bare "C"
// and "F" are coulomb and farad there (not temperature scales), so they are known but not convertible to K.
enum { DIM_OTHER = 0, DIM_PRESSURE = 1, DIM_TEMPERATURE, DIM_TIME, DIM_CONCENTRATION };
const UnitDef kUnits[] = {
    {"Pa", DIM_PRESSURE, 1.0, 0}, {"Pascal", DIM_PRESSURE, 1.0, 0}, {"pascal", DIM_PRESSURE, 1.0, 0}, {"kPa", DIM_PRESSURE, 1e3, 0},
    {"MPa", DIM_PRESSURE, 1e6, 0}, {"atm", DIM_PRESSURE, 101325.0, 0}, {"bar", DIM_PRESSURE, 1e5, 0}, {"mbar", DIM_PRESSURE, 1e2, 0},
    {"torr", DIM_PRESSURE, 101325.0 / 760.0, 0}, {"Torr", DIM_PRESSURE, 101325.0 / 760.0, 0}, {"psi", DIM_PRESSURE, 6894.757293168, 0},
    {"K", DIM_TEMPERATURE, 1.0, 0}, {"Kelvin", DIM_TEMPERATURE, 1.0, 0}, {"kelvin", DIM_TEMPERATURE, 1.0, 0},
    {"degC", DIM_TEMPERATURE, 1.0, 273.15}, {"Celsius", DIM_TEMPERATURE, 1.0, 273.15},
    {"degF", DIM_TEMPERATURE, 5.0 / 9.0, 459.67}, {"degR", DIM_TEMPERATURE, 5.0 / 9.0, 0},
    {"C", DIM_OTHER, 1.0, 0}, {"F", DIM_OTHER, 1.0, 0}, {"m", DIM_OTHER, 1.0, 0}, {"kg", DIM_OTHER, 1.0, 0}, {"J", DIM_OTHER, 1.0, 0}, {"mol", DIM_OTHER, 1.0, 0},
    {"hPa", DIM_PRESSURE, 1e2, 0}, {"GPa", DIM_PRESSURE, 1e9, 0}, {"mmHg", DIM_PRESSURE, 133.322387415, 0}, {"N/m^2", DIM_PRESSURE, 1.0, 0},
    {"s", DIM_TIME, 1.0, 0}, {"sec", DIM_TIME, 1.0, 0}, {"second", DIM_TIME, 1.0, 0}, {"ms", DIM_TIME, 1e-3, 0}, {"us", DIM_TIME, 1e-6, 0},
    {"ns", DIM_TIME, 1e-9, 0}, {"min", DIM_TIME, 60.0, 0}, {"h", DIM_TIME, 3600.0, 0}, {"hr", DIM_TIME, 3600.0, 0}, {"day", DIM_TIME, 86400.0, 0},
    {"mol/m^3", DIM_CONCENTRATION, 1.0, 0}, {"mol/m3", DIM_CONCENTRATION, 1.0, 0}, {"mmol/L", DIM_CONCENTRATION, 1.0, 0}, {"mol/L", DIM_CONCENTRATION, 1e3, 0},
    {"kmol/m^3", DIM_CONCENTRATION, 1e3, 0}, {"mol/cm^3", DIM_CONCENTRATION, 1e6, 0}, {"kmol/cm^3", DIM_CONCENTRATION, 1e9, 0}, {"mol/dm^3", DIM_CONCENTRATION, 1e3, 0},
};
double convert_unit(double v, const std::string& unit, int dim, const char* target) {
  for (auto& u : kUnits)
    if (unit == u.name) {
      if (u.dim != dim) throw InputError(InputError::InvalidArgument, unit + " is not convertible to " + target + ".");
      return (v + u.offset) * u.scale;
    }
  throw InputError(InputError::InvalidArgument, unit + " is not a valid unit.");
}

const YamlNode& need(const YamlNode& map, const char* key, const char* ctx) {
  const YamlNode* n = map.find(key);
  if (!n || n->kind == YamlNode::Null) throw InputError(InputError::InvalidNode, "missing mandatory key '" + std::string(key) + "' in " + std::string(ctx));
  return *n;
}
double as_double(const YamlNode& n, const std::string& ctx) {
  if (n.kind != YamlNode::Scalar) throw InputError(InputError::BadConversion, ctx + ": expected a number");
  const char* b = n.scalar.c_str();
  char* e = nullptr;
  double v = std::strtod(b, &e);
  if (e == b || *e) throw InputError(InputError::BadConversion, ctx + ": '" + n.scalar + "' is not a number");
  return v;
}
std::string as_string(const YamlNode& n, const std::string& ctx) {
  if (n.kind != YamlNode::Scalar) throw InputError(InputError::BadConversion, ctx + ": expected a scalar");
  return n.scalar;
}
std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

InputCase decode_case(const YamlNode& v) {
  InputCase c;
  if (const YamlNode* d = v.find("Description")) { c.description = as_string(*d, "Case.Description"); c.has_description = true; }
  {
    const YamlNode& n = need(v, "Pressure", "Case");
    c.pressure = convert_unit(as_double(need(n, "Value", "Pressure"), "Pressure.Value"), as_string(need(n, "Unit", "Pressure"), "Pressure.Unit"), DIM_PRESSURE, "Pa");
  }
  {
    const YamlNode& n = need(v, "Temperature", "Case");
    c.temperature = convert_unit(as_double(need(n, "Value", "Temperature"), "Temperature.Value"), as_string(need(n, "Unit", "Temperature"), "Temperature.Unit"), DIM_TEMPERATURE, "K");
  }
  {
    const YamlNode& n = need(v, "Composition", "Case");
    const std::string type = as_string(need(n, "Type", "Composition"), "Composition.Type");
    const YamlNode& val = need(n, "Value", "Composition");
    if (val.kind != YamlNode::Sequence) throw InputError(InputError::BadConversion, "Composition.Value: expected [[name, value], ...]");
    for (auto& it : val.items) {
      if (it.kind != YamlNode::Sequence || it.items.size() != 2) throw InputError(InputError::BadConversion, "Composition.Value: expected [name, value] pairs");
      c.composition.emplace_back(as_string(it.items[0], "Composition species"), as_double(it.items[1], "Composition value"));
    }
    const std::string t = lower(type);  // Composition::Set lower-cases (types.hpp:163)
    if (t == "concentration") {
      // the decoder's own test is case-sensitive (src/input_parser.cpp:99-107): any other spelling reaches Set() without a unit
      const std::string unit = type == "concentration" ? as_string(need(n, "Unit", "Composition"), "Composition.Unit") : std::string();
      const double f = convert_unit(1.0, unit, DIM_CONCENTRATION, "mol/m^3");
      for (auto& p : c.composition) p.second *= f;
      c.composition_type = InputCase::Concentration;
    } else if (t == "mass fraction") c.composition_type = InputCase::MassFraction;
    else if (t == "mole fraction") c.composition_type = InputCase::MoleFraction;
    else throw InputError(InputError::InvalidArgument, "Composition type is set to \"" + t + "\". The valid types are: \"concentration\", \"mass fraction\" or \"mole fraction\".");
  }
  {
    const YamlNode& n = need(v, "Time", "Case");
    c.time = convert_unit(as_double(need(n, "Value", "Time"), "Time.Value"), as_string(need(n, "Unit", "Time"), "Time.Unit"), DIM_TIME, "s");
  }
  return c;
}

InputCaseSet decode_case_set(const YamlNode& node) {
  InputCaseSet s;
  const YamlNode& mech = need(node, "Mechanism", "case set");
  s.mechanism_path = as_string(need(mech, "Path", "Mechanism"), "Mechanism.Path");
  if (const YamlNode* d = mech.find("Description")) { s.mechanism_description = as_string(*d, "Mechanism.Description"); s.has_mechanism_description = true; }
  if (const YamlNode* d = node.find("Description")) { s.description = as_string(*d, "Description"); s.has_description = true; }
  const YamlNode& cases = need(node, "Cases", "case set");
  if (cases.kind != YamlNode::Mapping) throw InputError(InputError::InvalidNode, "Cases must be a mapping of Case entries");
  for (auto& kv : cases.entries)
    if (kv.first == "Case" && kv.second.kind == YamlNode::Mapping) s.cases.push_back(decode_case(kv.second));
  return s;
}

}  // namespace

const YamlNode* YamlNode::find(const std::string& key) const {
  if (kind != Mapping) return nullptr;
  for (auto& kv : entries) if (kv.first == key) return &kv.second;
  return nullptr;
}

YamlNode load_yaml(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw InputError(InputError::BadFile, "bad file: " + path);
  std::vector<Line> lines;
  std::string raw;
  int no = 0;
  while (std::getline(f, raw)) {
    ++no;
    std::string t = strip_comment(raw);
    size_t a = t.find_first_not_of(' ');
    if (a == std::string::npos || strip(t).empty()) continue;
    if (t.find('\t') != std::string::npos && t.find_first_not_of(" \t") > t.find('\t')) throw InputError(InputError::InvalidNode, "line " + std::to_string(no) + ": tab indentation");
    std::string body = strip(t);
    if (body == "---" || body == "...") continue;
    lines.push_back({(int)a, body, no});
  }
  return Reader(std::move(lines)).parse_document();
}

InputParameters parse_input(const std::string& config_file) {
  InputParameters P;
  const YamlNode root = load_yaml(config_file);
  if (root.kind != YamlNode::Sequence) throw InputError(InputError::InvalidNode, "top level must be a sequence of case sets");
  for (auto& node : root.items) {
    const std::string r = as_string(need(node, "Reactor", "case set"), "Reactor");
    // src/input_parser.cpp:153-168: unknown reactor names are silently skipped
    if (r == "CONSTANT_PRESSURE") P.const_P.push_back(decode_case_set(node));
    if (r == "CONSTANT_VOLUME") P.const_V.push_back(decode_case_set(node));
    if (r == "CONSTANT_TEMPERATURE_PRESSURE") P.const_TP.push_back(decode_case_set(node));
    if (r == "CONSTANT_TEMPERATURE_VOLUME") P.const_TV.push_back(decode_case_set(node));
  }
  return P;
}

void validate_case(const InputCase& c, const CompiledMech& mech, double Tmin, double Tmax, std::vector<std::string>* warnings) {
  if (c.temperature < 0)
    throw InputError(InputError::Runtime, "Input error: the specified teperature is below absolute zero. T = " + std::to_string(c.temperature) + " K.");
  if ((c.temperature < Tmin || c.temperature > Tmax) && warnings)
    warnings->push_back("Input warning: the specified temperature T = " + std::to_string(c.temperature) + " K is outside the mechanism's applicable thermodynamic range [" +
                        std::to_string(Tmin) + ", " + std::to_string(Tmax) + "] K.");
  if (c.pressure < 0.0) throw InputError(InputError::Runtime, "Input error: the specified pressure is negative. P = " + std::to_string(c.pressure) + " Pa.");
  std::set<std::string> seen;
  std::string dup, invalid, negative;
  for (auto& p : c.composition) {
    if (!seen.insert(p.first).second) dup += "' '" + p.first + "' '";
  }
  if (!dup.empty()) throw InputError(InputError::Runtime, "Input error: found duplicate species: [" + dup + "].");
  for (auto& p : c.composition) {
    bool ok = false;
    for (auto& n : mech.species_names) if (n == p.first) ok = true;
    if (!ok) invalid += "' '" + p.first + "' '";
  }
  if (!invalid.empty()) throw InputError(InputError::Runtime, "Input error: found invalid species: [" + invalid + "].");
  for (auto& p : c.composition) if (p.second < 0) negative += "' '" + p.first + ":" + std::to_string(p.second) + "' '";
  if (!negative.empty()) throw InputError(InputError::Runtime, "Input error: found negative species: [" + negative + "].");
  if (c.composition_type != InputCase::Concentration) {
    double sum = 0.0;
    for (auto& p : c.composition) sum += p.second;
    if (std::fabs(1.0 - sum) > 1.0e-12)
      throw InputError(InputError::Runtime, std::string("Input error: the ") + (c.composition_type == InputCase::MassFraction ? "mass" : "mole") +
                                                " fractions do not sum up to unity. The sum is " + std::to_string(sum) + ".");
  }
}

std::vector<double> case_mass_fractions(const InputCase& c, const CompiledMech& mech) {
  std::vector<double> v(mech.K, 0.0);
  for (auto& p : c.composition) {
    int k = -1;
    for (int i = 0; i < mech.K; ++i) if (mech.species_names[i] == p.first) k = i;
    if (k < 0) throw InputError(InputError::Runtime, "Input error: found invalid species: [' '" + p.first + "' '].");
    v[k] += p.second;
  }
  // mole fractions and concentrations are both proportional to moles: Y_k ~ n_k W_k
  if (c.composition_type != InputCase::MassFraction) for (int k = 0; k < mech.K; ++k) v[k] *= mech.W[k];
  double s = 0.0;
  for (double x : v) s += x;
  if (!(s > 0.0)) throw InputError(InputError::Runtime, "Input error: empty composition.");
  for (double& x : v) x /= s;
  return v;
}

}  // namespace ctb
