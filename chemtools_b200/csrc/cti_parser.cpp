// Recursive-descent parser for the Python-call subset used by Cantera .cti
// mechanism files (GRI-Mech 1.2 / 2.11 / 3.0), plus the
// ".ctm" compiled-text format.  No Python, no Cantera.
//
// Reference behaviour mirrored: Chemistry::Thermo::Create loads the FIRST
// ideal_gas entry when no phase id is given
// (/root/reference/include/chemistry.hpp:71-78).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>

#include "mechanism.hpp"

namespace ctb {
namespace {

struct Value;
using ValuePtr = std::shared_ptr<Value>;
struct Value {
  enum Kind { Number, String, Ident, List, Call } kind = Number;
  double num = 0;
  std::string str;               // String / Ident / Call name
  std::vector<ValuePtr> items;   // List items or positional args
  std::vector<std::pair<std::string, ValuePtr>> kwargs;
  int line = 0;
};

struct Tok {
  enum Kind { End, Num, Str, Id, Punct } kind = End;
  std::string text;
  double num = 0;
  int line = 0;
};

class Lexer {
 public:
  Lexer(const std::string& src, const std::string& path) : s_(src), path_(path) {}
  Tok next() {
    skip_ws();
    Tok t;
    t.line = line_;
    if (i_ >= s_.size()) return t;
    char c = s_[i_];
    // string literal with optional u/r/b prefixes
    size_t j = i_;
    while (j < s_.size() && (s_[j] == 'u' || s_[j] == 'U' || s_[j] == 'r' || s_[j] == 'R' || s_[j] == 'b')) ++j;
    if (j < s_.size() && (s_[j] == '"' || s_[j] == '\'') && j - i_ <= 2) {
      i_ = j;
      t.kind = Tok::Str;
      t.text = read_string();
      return t;
    }
    if (std::isdigit((unsigned char)c) || (c == '.' && i_ + 1 < s_.size() && std::isdigit((unsigned char)s_[i_ + 1]))) {
      const char* b = s_.c_str() + i_;
      char* e = nullptr;
      t.num = std::strtod(b, &e);
      t.kind = Tok::Num;
      t.text.assign(b, (size_t)(e - b));
      i_ += (size_t)(e - b);
      return t;
    }
    if (std::isalpha((unsigned char)c) || c == '_') {
      size_t b = i_;
      while (i_ < s_.size() && (std::isalnum((unsigned char)s_[i_]) || s_[i_] == '_')) ++i_;
      t.kind = Tok::Id;
      t.text = s_.substr(b, i_ - b);
      return t;
    }
    t.kind = Tok::Punct;
    t.text = std::string(1, c);
    ++i_;
    return t;
  }
  [[noreturn]] void fail(const std::string& msg, int line) const {
    throw MechError(path_ + ":" + std::to_string(line) + ": " + msg);
  }

 private:
  void skip_ws() {
    while (i_ < s_.size()) {
      char c = s_[i_];
      if (c == '\n') { ++line_; ++i_; }
      else if (std::isspace((unsigned char)c) || c == '\\') ++i_;
      else if (c == '#') { while (i_ < s_.size() && s_[i_] != '\n') ++i_; }
      else break;
    }
  }
  std::string read_string() {
    char q = s_[i_];
    bool triple = i_ + 2 < s_.size() && s_[i_ + 1] == q && s_[i_ + 2] == q;
    std::string out;
    int start_line = line_;
    i_ += triple ? 3 : 1;
    for (;;) {
      if (i_ >= s_.size()) fail("unterminated string", start_line);
      char c = s_[i_];
      if (triple) {
        if (c == q && i_ + 2 < s_.size() + 0 && s_.compare(i_, 3, std::string(3, q)) == 0) { i_ += 3; break; }
      } else {
        if (c == q) { ++i_; break; }
        if (c == '\n') fail("newline in string", start_line);
      }
      if (c == '\\' && i_ + 1 < s_.size()) {
        char n = s_[i_ + 1];
        if (n == 'n') out += '\n'; else if (n == 't') out += '\t'; else if (n == '\n') { ++line_; } else out += n;
        i_ += 2;
        continue;
      }
      if (c == '\n') ++line_;
      out += c;
      ++i_;
    }
    return out;
  }
  const std::string& s_;
  std::string path_;
  size_t i_ = 0;
  int line_ = 1;
};

class Parser {
 public:
  Parser(const std::string& src, const std::string& path) : lex_(src, path) { advance(); }
  bool at_end() const { return cur_.kind == Tok::End; }
  ValuePtr statement() { return expr(); }

 private:
  void advance() { cur_ = lex_.next(); }
  bool is_punct(const char* p) const { return cur_.kind == Tok::Punct && cur_.text == p; }
  void expect(const char* p) {
    if (!is_punct(p)) lex_.fail(std::string("expected '") + p + "' but found '" + cur_.text + "'", cur_.line);
    advance();
  }
  ValuePtr expr() {
    auto v = std::make_shared<Value>();
    v->line = cur_.line;
    if (is_punct("-") || is_punct("+")) {
      bool neg = is_punct("-");
      advance();
      auto inner = expr();
      if (inner->kind != Value::Number) lex_.fail("unary sign on non-number", v->line);
      if (neg) inner->num = -inner->num;
      return inner;
    }
    switch (cur_.kind) {
      case Tok::Num: v->kind = Value::Number; v->num = cur_.num; advance(); return v;
      case Tok::Str: {
        v->kind = Value::String;
        while (cur_.kind == Tok::Str) { v->str += cur_.text; advance(); }  // implicit concatenation
        return v;
      }
      case Tok::Id: {
        v->str = cur_.text;
        advance();
        if (is_punct("(")) {
          v->kind = Value::Call;
          advance();
          parse_args(*v, ")");
          return v;
        }
        v->kind = Value::Ident;
        return v;
      }
      case Tok::Punct:
        if (is_punct("[")) { advance(); v->kind = Value::List; parse_args(*v, "]"); return v; }
        if (is_punct("(")) { advance(); v->kind = Value::List; parse_args(*v, ")"); return v; }
        break;
      default: break;
    }
    lex_.fail("unexpected token '" + cur_.text + "'", cur_.line);
  }
  void parse_args(Value& v, const char* close) {
    while (!is_punct(close)) {
      if (cur_.kind == Tok::End) lex_.fail("unbalanced bracket", v.line);
      if (cur_.kind == Tok::Id) {
        Tok id = cur_;
        advance();
        if (is_punct("=")) {
          advance();
          v.kwargs.emplace_back(id.text, expr());
        } else {
          auto iv = std::make_shared<Value>();
          iv->line = id.line;
          iv->str = id.text;
          if (is_punct("(")) { iv->kind = Value::Call; advance(); parse_args(*iv, ")"); }
          else iv->kind = Value::Ident;
          v.items.push_back(iv);
        }
      } else {
        v.items.push_back(expr());
      }
      if (is_punct(",")) advance();
      else if (!is_punct(close)) lex_.fail("expected ',' in argument list", cur_.line);
    }
    advance();
  }
  Lexer lex_;
  Tok cur_;
};

std::vector<std::string> split_ws(const std::string& s) {
  std::istringstream is(s);
  std::vector<std::string> out;
  std::string t;
  while (is >> t) out.push_back(t);
  return out;
}

std::vector<std::pair<std::string, double>> parse_pairs(const std::string& s, const std::string& ctx) {
  std::vector<std::pair<std::string, double>> out;
  std::string t = s;
  for (auto& c : t) if (c == ',') c = ' ';
  for (auto& tok : split_ws(t)) {
    size_t p = tok.rfind(':');
    if (p == std::string::npos || p == 0 || p + 1 >= tok.size()) throw MechError("bad 'name:value' token '" + tok + "' in " + ctx);
    char* e = nullptr;
    double v = std::strtod(tok.c_str() + p + 1, &e);
    if (*e) throw MechError("bad number in '" + tok + "' in " + ctx);
    out.emplace_back(tok.substr(0, p), v);
  }
  return out;
}

const Value* kw(const Value& call, const char* name) {
  for (auto& k : call.kwargs) if (k.first == name) return k.second.get();
  return nullptr;
}
double as_num(const Value* v, const std::string& ctx) {
  if (!v) throw MechError("missing number: " + ctx);
  if (v->kind == Value::Number) return v->num;
  if (v->kind == Value::Ident && v->str == "OneAtm") return 101325.0;
  throw MechError("expected a number: " + ctx);
}
std::string as_str(const Value* v, const std::string& ctx) {
  if (!v || v->kind != Value::String) throw MechError("expected a string: " + ctx);
  return v->str;
}
void as_arrhenius(const Value* v, double out[3], const std::string& ctx) {
  if (!v || v->kind != Value::List || v->items.size() != 3) throw MechError("expected [A, b, E]: " + ctx);
  for (int i = 0; i < 3; ++i) out[i] = as_num(v->items[i].get(), ctx);
}

void parse_side(const std::string& side_in, std::vector<std::pair<std::string, int>>& terms, bool& has_M, bool& fall_M,
                const std::string& eq) {
  std::string side = side_in;
  has_M = fall_M = false;
  // "(+ M)" / "(+M)"
  size_t p = side.find('(');
  while (p != std::string::npos) {
    size_t q = side.find(')', p);
    if (q == std::string::npos) break;
    std::string inner;
    for (size_t i = p + 1; i < q; ++i) if (!std::isspace((unsigned char)side[i])) inner += side[i];
    if (inner == "+M") { fall_M = true; side.erase(p, q - p + 1); p = side.find('('); }
    else p = side.find('(', q);
  }
  // split on '+' that is followed or preceded by whitespace (species names keep embedded '+')
  std::vector<std::string> parts;
  std::string cur;
  for (size_t i = 0; i < side.size(); ++i) {
    char c = side[i];
    bool sep = c == '+' && ((i + 1 < side.size() && std::isspace((unsigned char)side[i + 1])) ||
                            (i > 0 && std::isspace((unsigned char)side[i - 1])));
    if (sep) { parts.push_back(cur); cur.clear(); } else cur += c;
  }
  parts.push_back(cur);
  for (auto& part : parts) {
    auto toks = split_ws(part);
    if (toks.empty()) continue;
    if (toks.size() == 1 && toks[0] == "M") { has_M = true; continue; }
    int coeff = 1;
    std::string name;
    if (toks.size() == 2) {
      char* e = nullptr;
      double c = std::strtod(toks[0].c_str(), &e);
      if (*e || c != std::floor(c) || c < 1) throw MechError("unsupported (non-integer) stoichiometric coefficient in: " + eq);
      coeff = (int)c;
      name = toks[1];
    } else if (toks.size() == 1) {
      name = toks[0];
    } else {
      throw MechError("cannot parse term '" + part + "' in: " + eq);
    }
    terms.emplace_back(name, coeff);
  }
}

}  // namespace

void parse_equation(const std::string& eq, ReactionDef& out, bool& has_M, bool& has_falloff_M) {
  size_t p;
  std::string lhs, rhs;
  if ((p = eq.find("<=>")) != std::string::npos) { lhs = eq.substr(0, p); rhs = eq.substr(p + 3); out.reversible = true; }
  else if ((p = eq.find("=>")) != std::string::npos) { lhs = eq.substr(0, p); rhs = eq.substr(p + 2); out.reversible = false; }
  else if ((p = eq.find('=')) != std::string::npos) { lhs = eq.substr(0, p); rhs = eq.substr(p + 1); out.reversible = true; }
  else throw MechError("no '<=>', '=>' or '=' in reaction equation: " + eq);
  bool m1, f1, m2, f2;
  out.reactants.clear(); out.products.clear();
  parse_side(lhs, out.reactants, m1, f1, eq);
  parse_side(rhs, out.products, m2, f2, eq);
  if (m1 != m2 || f1 != f2) throw MechError("third body appears on one side only: " + eq);
  if (out.reactants.empty() || out.products.empty()) throw MechError("empty side in: " + eq);
  has_M = m1; has_falloff_M = f1;
  out.equation = eq;
}

MechanismDef parse_cti(const std::string& path, const char* phase_id) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw MechError("cannot open mechanism file: " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  std::string src = ss.str();
  Parser P(src, path);

  struct Phase { std::string name; std::vector<std::string> elements, species; };
  std::vector<Phase> phases;
  std::map<std::string, SpeciesDef> species;
  MechanismDef M;

  while (!P.at_end()) {
    ValuePtr st = P.statement();
    if (st->kind == Value::String) continue;  // module docstring
    if (st->kind != Value::Call) throw MechError(path + ":" + std::to_string(st->line) + ": unsupported top-level statement");
    const Value& c = *st;
    const std::string ctx = path + ":" + std::to_string(c.line) + " " + c.str + "()";
    if (c.str == "units") {
      if (auto v = kw(c, "length")) M.u_length = as_str(v, ctx);
      if (auto v = kw(c, "time")) M.u_time = as_str(v, ctx);
      if (auto v = kw(c, "quantity")) M.u_quantity = as_str(v, ctx);
      if (auto v = kw(c, "act_energy")) M.u_act_energy = as_str(v, ctx);
    } else if (c.str == "ideal_gas") {
      Phase ph;
      ph.name = as_str(kw(c, "name"), ctx);
      ph.elements = split_ws(as_str(kw(c, "elements"), ctx));
      ph.species = split_ws(as_str(kw(c, "species"), ctx));
      if (auto r = kw(c, "reactions"))
        if (as_str(r, ctx) != "all") throw MechError(ctx + ": only reactions='all' is supported");
      phases.push_back(ph);
    } else if (c.str == "species") {
      SpeciesDef s;
      s.name = as_str(kw(c, "name"), ctx);
      s.atoms = parse_pairs(as_str(kw(c, "atoms"), ctx), ctx);
      const Value* th = kw(c, "thermo");
      if (!th || th->kind != Value::List || th->items.size() != 2) throw MechError(ctx + ": two NASA ranges expected");
      struct Rng { double lo, hi; double a[7]; } r[2];
      for (int i = 0; i < 2; ++i) {
        const Value& n = *th->items[i];
        if (n.kind != Value::Call || n.str != "NASA" || n.items.size() != 2) throw MechError(ctx + ": only NASA-7 thermo is supported");
        const Value& tr = *n.items[0];
        const Value& co = *n.items[1];
        if (tr.kind != Value::List || tr.items.size() != 2 || co.kind != Value::List || co.items.size() != 7)
          throw MechError(ctx + ": NASA([Tlo,Thi],[7 coefficients]) expected");
        r[i].lo = as_num(tr.items[0].get(), ctx);
        r[i].hi = as_num(tr.items[1].get(), ctx);
        for (int k = 0; k < 7; ++k) r[i].a[k] = as_num(co.items[k].get(), ctx);
      }
      int lo = r[0].lo <= r[1].lo ? 0 : 1, hi = 1 - lo;
      s.Tlo = r[lo].lo; s.Tmid = r[lo].hi; s.Thi = r[hi].hi;
      for (int k = 0; k < 7; ++k) { s.a_lo[k] = r[lo].a[k]; s.a_hi[k] = r[hi].a[k]; }
      species[s.name] = s;
    } else if (c.str == "reaction" || c.str == "three_body_reaction" || c.str == "falloff_reaction") {
      ReactionDef r;
      r.kind = c.str == "reaction" ? RxnKind::Elementary : c.str == "three_body_reaction" ? RxnKind::ThreeBody : RxnKind::Falloff;
      if (c.items.empty()) throw MechError(ctx + ": missing equation");
      bool has_M, fall_M;
      parse_equation(as_str(c.items[0].get(), ctx), r, has_M, fall_M);
      if (r.kind == RxnKind::Elementary && (has_M || fall_M)) throw MechError(ctx + ": reaction() with a third body");
      if (r.kind == RxnKind::ThreeBody && !has_M) throw MechError(ctx + ": three_body_reaction() without M");
      if (r.kind == RxnKind::Falloff && !fall_M) throw MechError(ctx + ": falloff_reaction() without (+ M)");
      const Value* kfv = c.items.size() > 1 ? c.items[1].get() : kw(c, "kf");
      as_arrhenius(kfv, r.kf, ctx);
      for (auto& k : c.kwargs) {
        if (k.first == "kf") continue;
        else if (k.first == "kf0") as_arrhenius(k.second.get(), r.kf0, ctx);
        else if (k.first == "efficiencies") r.efficiencies = parse_pairs(as_str(k.second.get(), ctx), ctx);
        else if (k.first == "options") {
          const Value& o = *k.second;
          auto has_dup = [](const std::string& s) { return s.find("duplicate") != std::string::npos; };
          if (o.kind == Value::String) r.duplicate = has_dup(o.str);
          else if (o.kind == Value::List) for (auto& it : o.items) if (it->kind == Value::String && has_dup(it->str)) r.duplicate = true;
        } else if (k.first == "falloff") {
          const Value& fo = *k.second;
          if (fo.kind != Value::Call || fo.str != "Troe") throw MechError(ctx + ": unsupported falloff function '" + fo.str + "' (only Troe/Lindemann)");
          r.troe = true;
          const char* names[4] = {"A", "T3", "T1", "T2"};
          for (int i = 0; i < 4; ++i) {
            const Value* v = kw(fo, names[i]);
            if (!v && (int)fo.items.size() > i) v = fo.items[i].get();
            if (!v) { if (i == 3) { r.troe_p[3] = 0.0; continue; } throw MechError(ctx + ": Troe needs A, T3, T1"); }
            r.troe_p[i] = as_num(v, ctx);
          }
        } else {
          throw MechError(ctx + ": unsupported keyword '" + k.first + "'");
        }
      }
      if (r.kind == RxnKind::Falloff && !kw(c, "kf0")) throw MechError(ctx + ": falloff_reaction() without kf0");
      M.reactions.push_back(r);
    } else if (c.str == "validate" || c.str == "dataset" || c.str == "element") {
      continue;
    } else {
      throw MechError(ctx + ": unsupported CTI entry (supported: units, ideal_gas, species, reaction, three_body_reaction, falloff_reaction)");
    }
  }
  if (phases.empty()) throw MechError(path + ": no ideal_gas phase");
  const Phase* ph = &phases[0];
  if (phase_id && *phase_id) {
    ph = nullptr;
    for (auto& p : phases) if (p.name == phase_id) ph = &p;
    if (!ph) throw MechError(path + ": no phase named '" + phase_id + "'");
  }
  M.phase = ph->name;
  M.elements = ph->elements;
  M.n_phases = (int)phases.size();
  for (auto& n : ph->species) {
    auto it = species.find(n);
    if (it == species.end()) throw MechError(path + ": phase lists undeclared species '" + n + "'");
    M.species.push_back(it->second);
  }
  return M;
}

// ---------------------------------------------------------------- .ctm I/O
namespace {
std::string g17(double v) {
  char buf[40];
  std::snprintf(buf, sizeof buf, "%.17g", v);
  return buf;
}
}  // namespace

void write_ctm(const MechanismDef& M, const std::string& path) {
  std::ofstream o(path);
  if (!o) throw MechError("cannot write " + path);
  o << "CTM 1\nPHASE " << M.phase << "\nUNITS " << M.u_length << ' ' << M.u_time << ' ' << M.u_quantity << ' ' << M.u_act_energy << "\nELEMENTS";
  for (auto& e : M.elements) o << ' ' << e;
  o << "\nSPECIES " << M.species.size() << "\n";
  for (auto& s : M.species) {
    o << "S " << s.name;
    for (auto& a : s.atoms) o << ' ' << a.first << ':' << g17(a.second);
    o << "\n T " << g17(s.Tlo) << ' ' << g17(s.Tmid) << ' ' << g17(s.Thi) << "\n LO";
    for (double a : s.a_lo) o << ' ' << g17(a);
    o << "\n HI";
    for (double a : s.a_hi) o << ' ' << g17(a);
    o << "\n";
  }
  o << "REACTIONS " << M.reactions.size() << "\n";
  for (auto& r : M.reactions) {
    o << "R " << (r.kind == RxnKind::Elementary ? 'E' : r.kind == RxnKind::ThreeBody ? 'T' : 'F') << ' ' << (r.duplicate ? 1 : 0) << ' ' << r.equation << "\n";
    o << " KF " << g17(r.kf[0]) << ' ' << g17(r.kf[1]) << ' ' << g17(r.kf[2]) << "\n";
    if (r.kind == RxnKind::Falloff) {
      o << " KF0 " << g17(r.kf0[0]) << ' ' << g17(r.kf0[1]) << ' ' << g17(r.kf0[2]) << "\n";
      if (r.troe) o << " TROE " << g17(r.troe_p[0]) << ' ' << g17(r.troe_p[1]) << ' ' << g17(r.troe_p[2]) << ' ' << g17(r.troe_p[3]) << "\n";
    }
    if (!r.efficiencies.empty()) {
      o << " EFF";
      for (auto& e : r.efficiencies) o << ' ' << e.first << ':' << g17(e.second);
      o << "\n";
    }
  }
  o << "END\n";
}

MechanismDef read_ctm(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw MechError("cannot open mechanism file: " + path);
  MechanismDef M;
  std::string line;
  int ln = 0;
  auto fail = [&](const std::string& m) -> void { throw MechError(path + ":" + std::to_string(ln) + ": " + m); };
  auto nums = [&](std::istringstream& is, double* out, int n) { for (int i = 0; i < n; ++i) if (!(is >> out[i])) fail("expected " + std::to_string(n) + " numbers"); };
  bool header = false;
  while (std::getline(f, line)) {
    ++ln;
    std::istringstream is(line);
    std::string key;
    if (!(is >> key)) continue;
    if (key == "CTM") { int v; is >> v; if (v != 1) fail("unsupported CTM version"); header = true; }
    else if (!header) fail("not a CTM file");
    else if (key == "PHASE") is >> M.phase;
    else if (key == "UNITS") is >> M.u_length >> M.u_time >> M.u_quantity >> M.u_act_energy;
    else if (key == "ELEMENTS") { std::string e; while (is >> e) M.elements.push_back(e); }
    else if (key == "SPECIES" || key == "REACTIONS" || key == "END") continue;
    else if (key == "S") {
      SpeciesDef s;
      is >> s.name;
      std::string rest, tok;
      while (is >> tok) rest += tok + " ";
      s.atoms = parse_pairs(rest, path);
      M.species.push_back(s);
    } else if (key == "T") { if (M.species.empty()) fail("T before S"); double t[3]; nums(is, t, 3); auto& s = M.species.back(); s.Tlo = t[0]; s.Tmid = t[1]; s.Thi = t[2]; }
    else if (key == "LO") { if (M.species.empty()) fail("LO before S"); nums(is, M.species.back().a_lo, 7); }
    else if (key == "HI") { if (M.species.empty()) fail("HI before S"); nums(is, M.species.back().a_hi, 7); }
    else if (key == "R") {
      ReactionDef r;
      char k; int dup;
      is >> k >> dup;
      std::string eq;
      std::getline(is, eq);
      size_t b = eq.find_first_not_of(' ');
      eq = b == std::string::npos ? "" : eq.substr(b);
      r.kind = k == 'E' ? RxnKind::Elementary : k == 'T' ? RxnKind::ThreeBody : RxnKind::Falloff;
      r.duplicate = dup != 0;
      bool m, fm;
      parse_equation(eq, r, m, fm);
      M.reactions.push_back(r);
    } else if (key == "KF") { if (M.reactions.empty()) fail("KF before R"); nums(is, M.reactions.back().kf, 3); }
    else if (key == "KF0") { if (M.reactions.empty()) fail("KF0 before R"); nums(is, M.reactions.back().kf0, 3); }
    else if (key == "TROE") { if (M.reactions.empty()) fail("TROE before R"); M.reactions.back().troe = true; nums(is, M.reactions.back().troe_p, 4); }
    else if (key == "EFF") {
      if (M.reactions.empty()) fail("EFF before R");
      std::string rest, tok;
      while (is >> tok) rest += tok + " ";
      M.reactions.back().efficiencies = parse_pairs(rest, path);
    } else fail("unknown record '" + key + "'");
  }
  if (!header || M.species.empty()) throw MechError(path + ": empty or invalid CTM file");
  return M;
}

}  // namespace ctb
